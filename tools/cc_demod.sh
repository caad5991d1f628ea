#!/bin/bash
# Compile ldd_demod.cu alone and print the register / spill lines of the compile-time-plan kernels.
cd /root/repo/lddecode_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v -c ldd_demod.cu -o /tmp/ldd_demod_check.o 2>&1 | grep -E "error|Li8192" -A2 | grep -E "error|Compiling|Used|spill"

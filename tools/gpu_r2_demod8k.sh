#!/bin/bash
# in-place float32 block (ldd_demod8k.cuh) against the Stockham block: parity tests, kernel time, per-step cycles
O=gpurun_out/${1:-d8a}
mkdir -p $O
( time python -m pytest tests/test_demod.py tests/test_parity_gpu.py -m gpu -q -x ) > $O/pytest_demod.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_demod.log
python tools/gpu_demod_only.py mixed f32 PAL > $O/demod_only_new.log 2>&1
LDD_STOCKHAM_BLOCK=1 python tools/gpu_demod_only.py mixed f32 PAL > $O/demod_only_old.log 2>&1
python tools/gpu_demod_only.py mixed NTSC audio > $O/demod_only_ntsc_new.log 2>&1
LDD_STOCKHAM_BLOCK=1 python tools/gpu_demod_only.py mixed NTSC audio > $O/demod_only_ntsc_old.log 2>&1
python tools/gpu_demod_phases.py mixed PAL > $O/phases_pal.txt 2>&1
python tools/gpu_demod_phases.py mixed NTSC audio > $O/phases_ntsc.txt 2>&1
tail -5 $O/pytest_demod.log; cat $O/demod_only_*.log; cat $O/phases_pal.txt | head -24; head -24 $O/phases_ntsc.txt

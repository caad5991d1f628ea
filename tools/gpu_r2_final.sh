#!/bin/bash
# round 2 closing visit (one GPU): smoke, all GPU tests, bench, launch list, ncu capture of the fused demodulation kernel
O=gpurun_out/${1:-fin}
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"
( time python -m pytest tests -m gpu -q -x ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python tools/kernel_bench.py PAL > $O/kernel_bench_pal.log 2>&1; cp gpurun_out/kernel_bench_PAL.json $O/ 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-extra > $O/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_mixed_kernel -c 1 -o $O/demod_mixed python tools/gpu_demod_only.py mixed > $O/ncu_full.log 2>&1
python tools/ncu_summary.py $O/demod_mixed.ncu-rep > $O/ncu_demod_mixed.csv 2>/dev/null
tail -2 $O/smoke.log; tail -4 $O/pytest_gpu.log; cut -c1-300 $O/bench.json; tail -2 $O/bench.err; tail -16 $O/kernel_bench_pal.log | cut -c1-200

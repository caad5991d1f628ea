#!/bin/bash
# tests + per-kernel bench + bench, one GPU
O=gpurun_out/${1:-r2q}
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
( time python -m pytest tests -m gpu -q -x --durations=5 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
python tools/kernel_bench.py PAL > $O/kernel_bench_pal.log 2>&1; cp gpurun_out/kernel_bench_PAL.json $O/ 2>/dev/null
python tools/kernel_bench.py NTSC > $O/kernel_bench_ntsc.log 2>&1; cp gpurun_out/kernel_bench_NTSC.json $O/ 2>/dev/null
python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -4 $O/pytest_gpu.log; tail -14 $O/kernel_bench_pal.log | cut -c1-220; tail -8 $O/kernel_bench_ntsc.log | cut -c1-220; cut -c1-300 $O/bench.json; tail -3 $O/bench.err

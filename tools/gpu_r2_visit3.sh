#!/bin/bash
# round 2, visit 3: all GPU tests, bench, demod-only timings, ncu source-level capture of the fused mixed kernel
O=gpurun_out/${1:-v3}
mkdir -p $O
( time python -m pytest tests -m gpu -q -x --durations=5 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python tools/gpu_demod_only.py mixed f32 PAL > $O/demod_only_pal.log 2>&1
python tools/gpu_demod_only.py mixed NTSC audio > $O/demod_only_ntsc.log 2>&1
python tools/gpu_demod_phases.py mixed NTSC audio > $O/phases_ntsc.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_mixed_kernel -c 1 -o $O/demod_mixed python tools/gpu_demod_only.py mixed > $O/ncu_full.log 2>&1
python tools/ncu_summary.py $O/demod_mixed.ncu-rep > $O/ncu_demod_mixed.csv 2>/dev/null
tail -4 $O/pytest_gpu.log; cut -c1-400 $O/bench.json; tail -3 $O/bench.err; cat $O/demod_only_*.log; grep -E "audio|float|tangle|stores" $O/phases_ntsc.txt

#!/bin/bash
# tests + demod-only timing (fused vs two-launch mixed lane) + bench, one GPU
O=gpurun_out/${1:-r2q}
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
( time python -m pytest tests -m gpu -q -x --durations=5 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
python tools/gpu_demod_only.py f32 mixed > $O/demod_only.log 2>&1
LDD_MIXED_TWO_LAUNCH=1 python tools/gpu_demod_only.py mixed > $O/demod_only_two_launch.log 2>&1
LDD_SPARE_SMS=0 python tools/gpu_demod_only.py mixed > $O/demod_only_spare0.log 2>&1
LDD_FLAG_MARGIN_HZ=6 python tools/gpu_demod_only.py mixed > $O/demod_only_margin6.log 2>&1
python tools/gpu_demod_only.py mixed NTSC audio > $O/demod_only_ntsc_audio.log 2>&1
python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
LDD_SPARE_SMS=0 python bench.py --steps 20 --warmup 3 --skip-cpu --no-extra > $O/bench_spare0.json 2> $O/bench_spare0.err
LDD_SPARE_SMS=8 python bench.py --steps 20 --warmup 3 --skip-cpu --no-extra > $O/bench_spare8.json 2> $O/bench_spare8.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_ntsc_audio_lds.csv python bench.py --system NTSC --audio --fmt lds --steps 2 --warmup 3 --skip-cpu --no-extra > $O/ncu_launch_ntsc.log 2>&1
tail -4 $O/pytest_gpu.log; cat $O/demod_only*.log; cut -c1-200 $O/bench.json; cut -c1-200 $O/bench_spare0.json;  cut -c1-200 $O/bench_spare8.json; tail -3 $O/bench.err

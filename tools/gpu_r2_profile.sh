#!/bin/bash
# Round-2 GPU visit: tests, bench, host-side profile, ncu launch lists and one full capture of the demodulation kernel.
O=gpurun_out/${1:-r2}
mkdir -p $O
( time python -m pytest tests -m gpu -q -s --durations=10 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python tools/gpu_e2e_profile.py > $O/e2e_profile.log 2>&1
python tools/gpu_demod_only.py f32 f64 mixed > $O/demod_only.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_pal.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-extra > $O/ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_ntsc_audio_lds.csv python bench.py --system NTSC --audio --fmt lds --steps 2 --warmup 3 --skip-cpu --no-extra > $O/ncu_launch_ntsc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_kernel -c 2 -o $O/demod_mixed python tools/gpu_demod_only.py mixed > $O/ncu_full.log 2>&1
python tools/ncu_summary.py $O/demod_mixed.ncu-rep > $O/ncu_demod_mixed.csv 2>/dev/null
tail -4 $O/pytest_gpu.log; cut -c1-1500 $O/bench.json; echo; cat $O/e2e_profile.log; cat $O/demod_only.log

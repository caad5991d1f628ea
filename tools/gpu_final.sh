#!/bin/bash
# The measurements quoted in DESIGN.md section 7 / committed under profiles/, one GPU-box visit.
O=gpurun_out/${1:-final}
mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest_gpu.log
python bench.py > $O/bench_pal.json 2> $O/bench_pal.err; echo "bench exit $?"
python bench.py --precision f64 --skip-cpu > $O/bench_pal_f64.json 2>&1
python bench.py --system NTSC --audio --skip-cpu > $O/bench_ntsc_audio.json 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>&1
python tools/kernel_bench.py > $O/kernel_bench_pal.log 2>&1; cp gpurun_out/kernel_bench_PAL.json $O/ 2>/dev/null
python tools/gpu_demod_only.py f32 f64 mixed > $O/demod_only.log 2>&1
python tools/sweep.py > $O/sweep.log 2>&1; cp gpurun_out/sweep.json $O/ 2>/dev/null
python tools/gpu_e2e_profile.py > $O/e2e_profile.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_pal.csv python bench.py --steps 2 --warmup 3 --skip-cpu > $O/ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_ntsc_audio.csv python bench.py --system NTSC --audio --steps 2 --warmup 3 --skip-cpu > $O/ncu_launch_ntsc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_kernel -c 2 -o $O/demod_mixed python tools/gpu_demod_only.py mixed > $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tbc_kernel|pilot_|refine_hsync|peaks_phase1|peaks_to_host|copy_small' -c 9 -o $O/small python bench.py --steps 1 --warmup 0 --skip-cpu > $O/ncu_small.log 2>&1
ls -la $O
tail -3 $O/pytest_gpu.log; cat $O/bench_pal.json; echo; cut -c1-400 $O/bench_pal_f64.json; echo; cut -c1-400 $O/bench_ntsc_audio.json; echo; cat $O/bench_reference.json; tail -12 $O/kernel_bench_pal.log | cut -c1-250; cat $O/demod_only.log; tail -5 $O/e2e_profile.log

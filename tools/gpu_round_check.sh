#!/bin/bash
# One GPU-box visit: parity suite, bench, per-stage times, launch list, ncu captures.
# Outputs under gpurun_out/<tag>.  usage: tools/gpu_round_check.sh [tag] [quick]
TAG=${1:-chk}
MODE=${2:-full}
O=gpurun_out/$TAG
mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
python tools/gpu_demod_only.py f32 f64 mixed > $O/demod_only.log 2>&1
python tools/gpu_stage_times.py > $O/stage_times.log 2>&1
python tools/gpu_host_profile.py > $O/host_profile.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu > $O/ncu_launch.log 2>&1
if [ "$MODE" = "full" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_kernel -c 2 -o $O/demod_mixed python tools/gpu_demod_only.py mixed > $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tbc_kernel|pilot_|refine_hsync_kernel|peaks_phase1' -c 6 -o $O/small python bench.py --steps 1 --warmup 0 --skip-cpu > $O/ncu_small.log 2>&1
fi
ls -la $O
tail -3 $O/pytest_gpu.log; cat $O/bench.json; cat $O/demod_only.log; tail -5 $O/stage_times.log; head -8 $O/host_profile.log

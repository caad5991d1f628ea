#!/bin/bash
# tests + ncu full captures of the smaller kernels (one GPU)
O=gpurun_out/${1:-r2n}
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
( time python -m pytest tests -m gpu -q -x --durations=5 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
python bench.py --system NTSC --audio --fmt lds --steps 3 --warmup 3 --skip-cpu --no-extra > $O/bench_ntsc.json 2> $O/bench_ntsc.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tbc_f32_kernel|burst_lines_kernel|burst_vote_kernel|audio2_kernel|peaks_phase1|refine_hsync_kernel|vbi_kernel' -c 8 -o $O/small_ntsc python bench.py --system NTSC --audio --fmt lds --steps 1 --warmup 3 --skip-cpu --no-extra > $O/ncu_small.log 2>&1
python tools/ncu_summary.py $O/small_ntsc.ncu-rep > $O/ncu_small_ntsc.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_mixed_kernel -c 1 -o $O/demod_mixed python tools/gpu_demod_only.py mixed > $O/ncu_full.log 2>&1
python tools/ncu_summary.py $O/demod_mixed.ncu-rep > $O/ncu_demod_mixed.csv 2>/dev/null
tail -4 $O/pytest_gpu.log; cut -c1-200 $O/bench_ntsc.json; ls -la $O

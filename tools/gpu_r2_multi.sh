#!/bin/bash
# multi-GPU visit: NCCL gather + strong-scaling checks, then bench at N GPUs (weak headline + nested strong record) and
# the standalone strong-scaling record.   usage: gpu_r2_multi.sh <ngpus> <outdir>
N=${1:-2}; O=gpurun_out/${2:-r2m$N}
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
( time python -m pytest tests/test_multigpu.py -m gpu -q -x -s ) > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_multigpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --scaling strong --seconds 16 --steps 5 > $O/bench_strong16_n$N.json 2> $O/bench_strong16_n$N.err; echo "strong rc=$?"
tail -5 $O/pytest_multigpu.log; cut -c1-400 $O/bench_n$N.json; tail -3 $O/bench_n$N.err; cut -c1-500 $O/bench_strong16_n$N.json; tail -3 $O/bench_strong16_n$N.err

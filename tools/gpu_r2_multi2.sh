#!/bin/bash
# multi-GPU visit 2: P2P (peer-store) gather against the NCCL gather and against no gather at all.
# usage: gpu_r2_multi2.sh <ngpus> <outdir> [full]
N=${1:-2}; O=gpurun_out/${2:-r2p$N}
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time python -m pytest tests/test_multigpu.py -m gpu -q -x -s ) > $O/pytest_multigpu_p2p.log 2>&1; echo "pytest p2p rc=$?" | tee -a $O/pytest_multigpu_p2p.log
( time LDD_GATHER=nccl python -m pytest tests/test_multigpu.py -m gpu -q -x -s ) > $O/pytest_multigpu_nccl.log 2>&1; echo "pytest nccl rc=$?" | tee -a $O/pytest_multigpu_nccl.log
$TR --master-port 29571 bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu --no-extra > $O/weak_p2p.json 2> $O/weak_p2p.err; echo "p2p rc=$?"
LDD_GATHER=nccl $TR --master-port 29572 bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu --no-extra > $O/weak_nccl.json 2> $O/weak_nccl.err; echo "nccl rc=$?"
LDD_BENCH_NO_GATHER=1 $TR --master-port 29573 bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu --no-extra > $O/weak_nogather.json 2> $O/weak_nogather.err; echo "nogather rc=$?"
if [ "$3" = "full" ]; then
  $TR --master-port 29574 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"
  $TR --master-port 29575 bench.py --gpus $N --scaling strong --seconds 60 --steps 3 > $O/bench_strong60_n$N.json 2> $O/bench_strong60_n$N.err; echo "strong rc=$?"
fi
tail -3 $O/pytest_multigpu_p2p.log; tail -3 $O/pytest_multigpu_nccl.log
for f in weak_p2p weak_nccl weak_nogather bench_n$N bench_strong60_n$N; do [ -f $O/$f.json ] && (echo $f; grep '^{' $O/$f.json | cut -c1-230; tail -2 $O/$f.err | cut -c1-300); done

#!/bin/bash
# multi-GPU visit 3: DMA-push gather (one peer copy per step) against the NCCL gather and no gather at all.
# usage: gpu_r2_multi3.sh <ngpus> <outdir> [tests]
N=${1:-2}; O=gpurun_out/${2:-r2q$N}
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$3" = "tests" ]; then
  ( time LDD_GATHER=push python -m pytest tests/test_multigpu.py -m gpu -q -x -s ) > $O/pytest_multigpu_push.log 2>&1; echo "pytest push rc=$?" | tee -a $O/pytest_multigpu_push.log
  ( time python -m pytest tests/test_multigpu.py -m gpu -q -x -s ) > $O/pytest_multigpu_default.log 2>&1; echo "pytest default rc=$?" | tee -a $O/pytest_multigpu_default.log
fi
LDD_GATHER=push $TR --master-port 29571 bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu --no-extra > $O/weak_push.json 2> $O/weak_push.err; echo "push rc=$?"
LDD_GATHER=nccl $TR --master-port 29572 bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu --no-extra > $O/weak_nccl.json 2> $O/weak_nccl.err; echo "nccl rc=$?"
LDD_BENCH_NO_GATHER=1 $TR --master-port 29573 bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu --no-extra > $O/weak_nogather.json 2> $O/weak_nogather.err; echo "nogather rc=$?"
tail -3 $O/pytest_multigpu_push.log 2>/dev/null; tail -3 $O/pytest_multigpu_default.log 2>/dev/null
for f in weak_push weak_nccl weak_nogather; do [ -f $O/$f.json ] && (echo $f; grep '^{' $O/$f.json | cut -c1-230; tail -2 $O/$f.err | cut -c1-300); done

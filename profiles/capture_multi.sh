#!/bin/bash
# profiles/capture_multi.sh <tag> <n_gpus> — one `gpurun --gpus N` call (charged N x): the N-GPU bench line launched as
# the driver launches it, the single-GPU line of the same box, then ONLY the multi-GPU tests (the single-GPU suite
# belongs to capture_ab.sh on a 1-GPU box).  Every step has its own short timeout: in r01k one hanging test ate the
# round's whole budget.
set -u
TAG=${1:-r02m}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 100 --warmup 3 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err; echo "bench n$N rc=$?"
tail -3 $OUT/bench_${TAG}_n$N.err | cut -c1-300
cat $OUT/bench_${TAG}_n$N.json | cut -c1-1200
timeout 240 python bench.py --no-cpu-baseline --skip-e2e --steps 100 > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err; echo "bench n1 rc=$?"
cat $OUT/bench_${TAG}_n1.json | cut -c1-400
timeout 480 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $OUT/pytest_multi_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_multi_$TAG.log
tail -5 $OUT/pytest_multi_$TAG.log | cut -c1-300

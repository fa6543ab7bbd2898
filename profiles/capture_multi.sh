#!/bin/bash
# profiles/capture_multi.sh <tag> <n_gpus> — one gpurun --gpus N call: the GPU suite (multi-GPU cases included), the
# single-GPU bench line (device part only), the N-GPU bench line launched as the driver launches it.
set -u
TAG=${1:-r01m}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -6 $OUT/pytest_gpu_$TAG.log | cut -c1-400
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err; echo "bench n1 rc=$?"
cat $OUT/bench_${TAG}_n1.json | cut -c1-700
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 100 --warmup 3 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err; echo "bench n$N rc=$?"
tail -5 $OUT/bench_${TAG}_n$N.err | cut -c1-400
cat $OUT/bench_${TAG}_n$N.json

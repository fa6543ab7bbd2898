#!/bin/bash
# profiles/capture_n8.sh <tag> — one `gpurun --gpus 8` call (charged 8 x): the 8-GPU bench line as the driver launches it
# (8 x configs[2], parity-gated), the configs[3] line (3.1 Gbp / 30x, 251 M overlap records over 8 GPUs), the one-process-per-GPU
# parity tests at world 1/2/4/8, and the per-phase timeline of a graph-replayed step.  Every step has its own timeout.
set -u
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 420 29517 bench.py --gpus 8 --steps 100 --warmup 3 > $OUT/bench_${TAG}_n8.json 2> $OUT/bench_${TAG}_n8.err; echo "bench c3 n8 rc=$?"
tail -2 $OUT/bench_${TAG}_n8.err | cut -c1-300; cut -c1-1500 $OUT/bench_${TAG}_n8.json
run 600 29518 bench.py --gpus 8 --steps 50 --warmup 3 --workload c4s > $OUT/bench_${TAG}_c4s_n8.json 2> $OUT/bench_${TAG}_c4s_n8.err; echo "bench c4s n8 rc=$?"
tail -2 $OUT/bench_${TAG}_c4s_n8.err | cut -c1-300; cut -c1-1500 $OUT/bench_${TAG}_c4s_n8.json
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "one_process_per_gpu and (8 or 4)" > $OUT/pytest_multi_${TAG}.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_multi_${TAG}.log
tail -3 $OUT/pytest_multi_${TAG}.log | cut -c1-300
run 300 29519 profiles/fabric_timeline.py > $OUT/timeline_${TAG}_n8.json 2> $OUT/timeline_${TAG}_n8.err; echo "timeline rc=$?"
head -c 1800 $OUT/timeline_${TAG}_n8.json

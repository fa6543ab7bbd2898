#!/bin/bash
# profiles/capture_quick.sh <tag> <kernel-regex> — parity tests, the bench line, and one `ncu --set full` capture of the named kernels
set -u
TAG=${1:-q}
RE=${2:-k_transitive_group|k_classify_events|k_classify_survivors}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
python bench.py --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s ${3:-9} -c ${4:-3} \
    -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log | cut -c1-300

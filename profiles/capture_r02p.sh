#!/bin/bash
# profiles/capture_r02p.sh <tag> — one 1-GPU gpurun call: whole GPU suite, the default bench line (with cli_baseline), the CLI
# bench of the noisy workload (configs[1]) and of configs[0] on two ranks, the bench line + ncu evidence of configs[4]
set -u
TAG=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -6 $OUT/pytest_gpu_$TAG.log | cut -c1-300
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -2 $OUT/bench_$TAG.err | cut -c1-300
timeout 600 python bench.py --cli --workload c2 > $OUT/bench_cli_c2_$TAG.json 2> $OUT/bench_cli_c2_$TAG.err; echo "cli c2 rc=$?"
timeout 300 python bench.py --cli --workload c1 --devices 0,0 > $OUT/bench_cli_c1_2ranks_$TAG.json 2> $OUT/bench_cli_c1_2ranks_$TAG.err; echo "cli c1 x2 rc=$?"
timeout 300 python bench.py --workload c5 --steps 100 > $OUT/bench_c5_$TAG.json 2> $OUT/bench_c5_$TAG.err; echo "bench c5 rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_c5_$TAG.csv \
    python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e > $OUT/bench_c5_under_ncu_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_transitive_group|k_transitive_light|k_transitive_heavy' -s 9 -c 3 \
    -o $OUT/prof_c5_$TAG -f python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e > $OUT/ncu_c5_$TAG.log 2>&1
tail -2 $OUT/ncu_c5_$TAG.log | cut -c1-200

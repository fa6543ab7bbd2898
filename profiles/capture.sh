#!/bin/bash
# profiles/capture.sh — what produced the files under profiles/ (run under gpurun on one B200):
#   gpurun --timeout 1500 -- 'bash profiles/capture.sh r01'
# 1. GPU parity tests, 2. the bench line, 3. the ncu launch list of the same bench command (shares, not absolutes),
# 4. one `ncu --set full` capture of the dominant kernels.  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/clocks_$TAG.csv &
SMI=$!
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json
kill $SMI
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'k_classify_events|k_classify_survivors|k_transitive_light|k_resolve$|k_fill_csr|k_emit_edges' -s 21 -c 7 \
    -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT

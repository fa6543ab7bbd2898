#!/bin/bash
# profiles/capture_final.sh <tag> — one 1-GPU gpurun call at the end of a round: GPU suite, smoke, bench line, ncu launch list
# of the bench command, one full capture of the step's main kernels, launch list of the multi-GPU session's kernels (one rank)
set -u
TAG=${1:-r02v}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/clocks_$TAG.csv &
SMI=$!
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -4 $OUT/pytest_gpu_$TAG.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -2 $OUT/bench_$TAG.err | cut -c1-300
kill $SMI
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e > $OUT/bench_under_ncu_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'k_classify_events|k_classify_survivors|k_relocate_runs|k_transitive_group|k_transitive_light|k_resolve$|k_resolve_prepare|k_fill_csr|k_emit_edges' -s 18 -c 10 \
    -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e > $OUT/ncu_full_$TAG.log 2>&1
tail -1 $OUT/ncu_full_$TAG.log | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_${TAG}_fabric1.csv \
    python profiles/fabric_world1.py > $OUT/fabric_world1_$TAG.json 2> $OUT/fabric_world1_$TAG.err; echo "fabric1 rc=$?"

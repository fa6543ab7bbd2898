#!/bin/bash
# profiles/capture_skew.sh <tag> — ncu evidence for the skewed-degree path of the transitive pass (BASELINE configs[4]):
# launch durations of all K3 kernels, then one full-set capture (divergence = smsp__thread_inst_executed_per_inst_executed,
# occupancy = sm__warps_active, sector efficiency of the CSR gathers = l1tex sectors / requests) of group / light / heavy.
set -u
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python profiles/skew_k3.py > $OUT/skew_$TAG.json 2> $OUT/skew_$TAG.err; echo "skew rc=$?"; cat $OUT/skew_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_skew_$TAG.csv \
    python profiles/skew_k3.py > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_transitive_group|k_transitive_light|k_transitive_heavy' -s 3 -c 3 \
    -o $OUT/prof_skew_$TAG -f python profiles/skew_k3.py > $OUT/ncu_skew_$TAG.log 2>&1
tail -2 $OUT/ncu_skew_$TAG.log | cut -c1-300

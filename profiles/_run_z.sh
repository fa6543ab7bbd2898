timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin_cli.py -m gpu -x -q -k "dropin" 2>&1 | tail -4
timeout 600 python bench.py --cli --workload c3 > gpurun_out/bench_cli_c3_r02z2.json 2> gpurun_out/bench_cli_c3_r02z2.err; echo "cli c3 rc=$?"
python - <<'P'
import json
c=json.load(open('gpurun_out/bench_cli_c3_r02z2.json'))
cb=c['cli_baseline']; print({k:cb.get(k) for k in ('dropin_wall_s','reference_wall_s','same_node_edge_transitive_counts','fasta_paf_written_s')}); print('dropin', {k:v for k,v in cb['dropin_logger'].items() if 'number' not in k}); print('ref', {k:v for k,v in cb['reference_logger'].items() if 'number' not in k}); print(c.get('device_ms'))
P

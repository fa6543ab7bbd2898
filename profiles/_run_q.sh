timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_r02q.json 2> gpurun_out/bench_r02q.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r02q.err | cut -c1-300
python - <<'P'
import json
b=json.load(open('gpurun_out/bench_r02q.json'))
print('ms/step', b['ms_per_step'], 'e2e', b['e2e'], 'frac', b['roofline']['frac'], b['roofline']['whole_step']['frac'], b['roofline']['stage_ms'])
c=b['cli_baseline']; print({k:c[k] for k in ('dropin_wall_s','reference_wall_s')}, c['dropin_logger'].get('rala::Graph::construct loaded overlaps'))
P

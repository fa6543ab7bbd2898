timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --ab --steps 200 > gpurun_out/ab_r02x.json 2> gpurun_out/ab_r02x.err; echo "ab rc=$?"; tail -2 gpurun_out/ab_r02x.err | cut -c1-300
python - <<'P'
import json
ab=json.load(open('gpurun_out/ab_r02x.json'))['ab']
for k,v in ab.items(): print(k, v['ms_per_step'], v['delta_us_vs_product'], v['same_result_as_product'])
P
timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 200 | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('ms/step', b['ms_per_step'], b['roofline']['stage_ms'], 'frac', b['roofline']['frac'], b['roofline']['whole_step']['frac'])"

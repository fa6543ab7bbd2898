timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_r02s_n2.json 2> gpurun_out/bench_r02s_n2.err; echo "bench n2 rc=$?"; tail -4 gpurun_out/bench_r02s_n2.err | cut -c1-400
python - <<'P'
import json
b=json.loads(open('gpurun_out/bench_r02s_n2.json').read().strip().splitlines()[-1])
print('ms/step', b['ms_per_step'], 'value', b['value']/1e9, 'parity', b['parity'], 'e2e', b['e2e'], 'agg', b['roofline']['aggregate_whole_step'])
P
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-600

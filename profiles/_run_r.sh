timeout 600 python -m pytest tests/test_multi_fabric.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config5 or self_overlap" 2>&1 | tail -3
for N in 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 profiles/fabric_timeline.py > gpurun_out/timeline_r02r_n$N.json 2> gpurun_out/timeline_r02r_n$N.err; echo "rc=$?"; tail -2 gpurun_out/timeline_r02r_n$N.err | cut -c1-300
cat gpurun_out/timeline_r02r_n$N.json
done

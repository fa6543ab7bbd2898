timeout 600 python -m pytest tests/test_multi_fabric.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 profiles/fabric_timeline.py > gpurun_out/timeline_r02y_n2.json 2> gpurun_out/timeline_r02y_n2.err; echo "rc=$?"; tail -2 gpurun_out/timeline_r02y_n2.err | cut -c1-300
cat gpurun_out/timeline_r02y_n2.json

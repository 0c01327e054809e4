mkdir -p gpurun_out
timeout 300 python scripts/host_profile.py 2>&1 | head -3
timeout 600 python bench.py --no-cpu-baseline --no-torch-eager --no-other-configs > gpurun_out/b66_n1.json 2> gpurun_out/b66_n1.err
python scripts/show_bench.py gpurun_out/b66_n1.json 2>/dev/null | sed -n '1,2p'
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --no-cpu-baseline --no-torch-eager --no-other-configs > gpurun_out/b66_n2.json 2> gpurun_out/b66_n2.err; echo rc=$?
python scripts/show_bench.py gpurun_out/b66_n2.json 2>/dev/null | sed -n '1,2p'
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_gemm.py -q -m gpu -x 2>&1 | tail -2

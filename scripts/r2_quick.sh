mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded_kinds.py -q -m gpu 2>&1 | tail -12

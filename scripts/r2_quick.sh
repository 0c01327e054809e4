mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded_kinds.py tests/test_gpu_sharded.py tests/test_staging_and_adam.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/t57.log
tail -40 gpurun_out/t57.log

mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_sharded.py tests/test_gpu_reference_trainer.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-torch-eager --no-other-configs > gpurun_out/b65_n1.json 2> gpurun_out/b65_n1.err
tail -3 gpurun_out/b65_n1.err
python scripts/show_bench.py gpurun_out/b65_n1.json 2>/dev/null | sed -n '1,3p'

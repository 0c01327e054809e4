mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -2
timeout 300 python bench.py --workload deepfm_pep_kdd --steps 10 --no-cpu-baseline --no-torch-eager --no-other-configs --small-batch 0 > gpurun_out/b67_pep.json 2> gpurun_out/b67_pep.err
python scripts/show_bench.py gpurun_out/b67_pep.json 2>/dev/null | sed -n '1p;4,9p'

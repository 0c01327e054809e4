set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_models.py tests/test_gpu_reference_parity.py tests/test_gpu_reference_trainer.py -q > gpurun_out/t48.log 2>&1; tail -5 gpurun_out/t48.log
timeout 900 python bench.py --no-other-configs > gpurun_out/b48_n1.json 2> gpurun_out/b48_n1.err; echo rc=$?
python scripts/show_bench.py gpurun_out/b48_n1.json > gpurun_out/b48_show.txt 2>&1; head -32 gpurun_out/b48_show.txt

mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_models.py tests/test_gpu_reference_parity.py tests/test_gpu_reference_trainer.py -q > gpurun_out/t52.log 2>&1; tail -8 gpurun_out/t52.log
for shape in "65536 400 624 0 0" "65536 400 400 0 0"; do FMT=fp16 python scripts/gemm_one.py $shape 20 2>&1 | tail -1; done
timeout 600 python bench.py --no-other-configs --no-cpu-baseline --no-torch-eager > gpurun_out/b52_n1.json 2> gpurun_out/b52_n1.err
python scripts/show_bench.py gpurun_out/b52_n1.json 2>/dev/null | head -24

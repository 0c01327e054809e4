set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -x > gpurun_out/t45_gemm.log 2>&1; tail -15 gpurun_out/t45_gemm.log
timeout 600 python scripts/gemm_planes_check.py --fp16 > gpurun_out/gemm_check_fp16.log 2>&1; tail -12 gpurun_out/gemm_check_fp16.log
RSB_GEMM_DRAIN_FP16=1 timeout 600 python scripts/gemm_planes_check.py --fp16 > gpurun_out/gemm_check_fp16_d1.log 2>&1; tail -6 gpurun_out/gemm_check_fp16_d1.log
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_reference_parity.py tests/test_gpu_reference_trainer.py tests/test_gpu_oracle_shapes.py -q > gpurun_out/t45_models.log 2>&1; tail -15 gpurun_out/t45_models.log
timeout 600 python bench.py --no-other-configs > gpurun_out/b45_n1.json 2> gpurun_out/b45_n1.err; echo rc=$?
python scripts/show_bench.py gpurun_out/b45_n1.json | head -30

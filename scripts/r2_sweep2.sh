mkdir -p gpurun_out
{
for shape in "65536 400 624 0 0" "65536 400 400 0 0" "65536 624 400 0 0"; do
  FMT=fp16 python scripts/gemm_one.py $shape 20 2>&1 | tail -1
  FMT=fp16 RSB_GEMM_PAIRS=1 python scripts/gemm_one.py $shape 20 2>&1 | tail -1
  FMT=fp16 RSB_GEMM_TMA_STORE=0 python scripts/gemm_one.py $shape 20 2>&1 | tail -1
done
python scripts/gemm_one.py 65536 400 624 0 0 20 2>&1 | tail -1
} > gpurun_out/sweep51.log 2>&1
cat gpurun_out/sweep51.log
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --no-other-configs --no-cpu-baseline --no-torch-eager > gpurun_out/b51_n1.json 2> gpurun_out/b51_n1.err
python scripts/show_bench.py gpurun_out/b51_n1.json 2>/dev/null | head -24

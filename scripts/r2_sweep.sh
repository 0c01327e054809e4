mkdir -p gpurun_out
{
for pairs in 0 1; do
  for d in 1 2 3 4; do
    FMT=fp16 RSB_GEMM_PAIRS=$pairs RSB_GEMM_DRAIN_FP16=$d python scripts/gemm_one.py 65536 400 624 0 0 20 2>&1 | tail -1
    FMT=fp16 RSB_GEMM_PAIRS=$pairs RSB_GEMM_DRAIN_FP16=$d python scripts/gemm_one.py 65536 400 400 0 0 20 2>&1 | tail -1
  done
done
FMT=fp16 RSB_GEMM_DEBUG=1 python scripts/gemm_one.py 65536 400 624 0 0 3 2>&1 | tail -3
FMT=fp16 RSB_GEMM_DEBUG=1 RSB_GEMM_PAIRS=1 python scripts/gemm_one.py 65536 400 624 0 0 3 2>&1 | tail -3
FMT=fp16 RSB_GEMM_TMA_STORE=0 python scripts/gemm_one.py 65536 400 624 0 0 20 2>&1 | tail -1
RSB_GEMM_PAIRS=1 python scripts/gemm_one.py 65536 400 624 0 0 20 2>&1 | tail -1
python scripts/gemm_one.py 65536 400 624 0 0 20 2>&1 | tail -1
} > gpurun_out/sweep46.log 2>&1
cat gpurun_out/sweep46.log

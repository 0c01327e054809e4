# drain-interval experiment for the FP16X2 dense tail: step time and which parity tests move
mkdir -p gpurun_out
for d in 2 3 4; do
  RSB_GEMM_DRAIN_FP16=$d timeout 300 python bench.py --no-other-configs --no-cpu-baseline --no-torch-eager > gpurun_out/b55_d$d.json 2> gpurun_out/b55_d$d.err
  echo "drain $d"; python scripts/show_bench.py gpurun_out/b55_d$d.json 2>/dev/null | head -1
done
RSB_GEMM_PAIRS=1 RSB_GEMM_DRAIN_FP16=4 timeout 300 python bench.py --no-other-configs --no-cpu-baseline --no-torch-eager > gpurun_out/b55_d4p.json 2> gpurun_out/b55_d4p.err
echo "drain 4 pairs"; python scripts/show_bench.py gpurun_out/b55_d4p.json 2>/dev/null | head -1
for d in 3 4; do
  RSB_GEMM_DRAIN_FP16=$d timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_sharded.py 2>&1 | tail -8 > gpurun_out/t55_d$d.log
  tail -3 gpurun_out/t55_d$d.log
done

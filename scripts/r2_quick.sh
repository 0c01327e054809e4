mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/t_final2.log; cat gpurun_out/t_final2.log

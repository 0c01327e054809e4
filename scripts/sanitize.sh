#!/bin/bash
# compute-sanitizer memcheck over a small slice of the GPU tests (one tool per gpurun call).
#   gpurun -- 'bash scripts/sanitize.sh memcheck'
tool=${1:-memcheck}
timeout 1500 compute-sanitizer --tool $tool --error-exitcode 97 --launch-timeout 0 \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x \
  -k "chunk_boundaries or small_table or qr or cerp or optembed or pep or fused_sparse or sort_applies or dcn or empty or odd_widths or int32" \
  > gpurun_out/sanitize_$tool.log 2>&1
echo "exit code $?" >> gpurun_out/sanitize_$tool.log
tail -15 gpurun_out/sanitize_$tool.log

mkdir -p gpurun_out
timeout 300 python -m pytest "tests/test_gpu_sharded_kinds.py::test_sharded_variants_world2_match_single_gpu" -q -m gpu -k "pep_feature_dim or dcn" 2>&1 | grep -E "passed|failed|Error|assert" | head -12

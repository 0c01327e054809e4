mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_gpu_sharded_kinds.py::test_sharded_variants_world2_match_single_gpu" tests/test_staging_and_adam.py::test_deepfm_trains_the_same_with_the_one_launch_adam -q -m gpu -k "dcn or one_launch" 2>&1 | grep -v "^  \|Warning" | grep -E "Error|error|assert|^E " | head -40 > gpurun_out/t59.log
cat gpurun_out/t59.log

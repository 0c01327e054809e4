set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_kernels.py tests/test_gpu_models.py -q -x > gpurun_out/t44.log 2>&1; tail -5 gpurun_out/t44.log
timeout 600 python bench.py --no-other-configs > gpurun_out/b44_n1.json 2> gpurun_out/b44_n1.err; echo rc=$?
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 > gpurun_out/b44_n2.json 2> gpurun_out/b44_n2.err; echo rc=$?
RSB_HOT_FIELD_ROWS=0 timeout 600 $TR bench.py --gpus 2 --no-parity-check > gpurun_out/b44_n2_nohot.json 2> gpurun_out/b44_n2_nohot.err; echo rc=$?
tail -c 300 gpurun_out/b44_n2.err

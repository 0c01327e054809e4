mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --no-torch-eager > gpurun_out/final_n8.json 2> gpurun_out/final_n8.err; echo n=8 rc=$?
tail -2 gpurun_out/final_n8.err
python scripts/show_bench.py gpurun_out/final_n8.json 2>/dev/null | sed -n '1,9p'

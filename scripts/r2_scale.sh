set -x
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n > gpurun_out/b50_n$n.json 2> gpurun_out/b50_n$n.err; echo n=$n rc=$?
done
timeout 600 python bench.py --no-other-configs > gpurun_out/b50_n1.json 2> gpurun_out/b50_n1.err; echo n=1 rc=$?
for n in 1 2 4 8; do python scripts/show_bench.py gpurun_out/b50_n$n.json 2>/dev/null | head -3; done

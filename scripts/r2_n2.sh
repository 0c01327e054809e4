mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 > gpurun_out/b60_n2.json 2> gpurun_out/b60_n2.err; echo rc=$?
python scripts/show_bench.py gpurun_out/b60_n2.json 2>/dev/null | head -12
for w in deepfm_pep_kdd_sharded deepfm_qr_criteo_sharded dcnmix_full_avazu_sharded; do
  timeout 600 $TR bench.py --gpus 2 --steps 10 --workload $w --no-cpu-baseline --no-torch-eager --no-other-configs > gpurun_out/b60_n2_$w.json 2> gpurun_out/b60_n2_$w.err; echo $w rc=$?
  tail -3 gpurun_out/b60_n2_$w.err
  python scripts/show_bench.py gpurun_out/b60_n2_$w.json 2>/dev/null | sed -n '1p;4,12p'
  timeout 600 python bench.py --steps 10 --workload $w --no-cpu-baseline --no-torch-eager --no-other-configs > gpurun_out/b60_n1_$w.json 2> gpurun_out/b60_n1_$w.err; echo $w n1 rc=$?
  python scripts/show_bench.py gpurun_out/b60_n1_$w.json 2>/dev/null | sed -n '1p'
done

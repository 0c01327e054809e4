mkdir -p gpurun_out
python scripts/r2_l2fetch.py
for g in 0 64 32; do
  for w in deepfm_pep_kdd deepfm_full_roofline deepfm_full_criteo_sharded; do
    RSB_L2_FETCH=$g timeout 300 python bench.py --workload $w --steps 10 --no-parity-check --no-other-configs --no-cpu-baseline --no-torch-eager --small-batch 0 > gpurun_out/b62_${w}_g$g.json 2> gpurun_out/b62_${w}_g$g.err
    echo "== granularity $g $w"; python scripts/show_bench.py gpurun_out/b62_${w}_g$g.json 2>/dev/null | sed -n '1p;7,9p'
  done
done

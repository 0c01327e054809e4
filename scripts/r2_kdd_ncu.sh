# ncu --set full of the gather-side kernels on the KDD-shaped PEP workload (the kernels furthest below the HBM roofline)
mkdir -p gpurun_out
LEAN="--workload deepfm_pep_kdd --steps 2 --warmup 3 --no-parity-check --no-other-configs --no-cpu-baseline --no-torch-eager --small-batch 0"
timeout 300 python bench.py $LEAN > gpurun_out/r2_kdd_plain.json 2> gpurun_out/r2_kdd_plain.err; echo plain rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lookup_|seg_' --launch-skip 30 -c 12 -o /tmp/r2_kdd -f python bench.py $LEAN > gpurun_out/r2_kdd_ncu.log 2>&1; echo full rc=$?
ncu -i /tmp/r2_kdd.ncu-rep --page raw --csv > gpurun_out/r2_kdd_raw.csv 2>/dev/null
ncu -i /tmp/r2_kdd.ncu-rep --page details --csv > gpurun_out/r2_kdd_details.csv 2>/dev/null
ls -la /tmp/r2_kdd.ncu-rep gpurun_out/r2_kdd_raw.csv

# final ncu evidence of the headline step (1 GPU): un-profiled run, launch list, --set full of every own kernel
mkdir -p gpurun_out
LEAN="--steps 2 --warmup 3 --no-parity-check --no-other-configs --no-cpu-baseline --no-torch-eager --small-batch 0"
timeout 300 python bench.py $LEAN > gpurun_out/r2f_plain.json 2> gpurun_out/r2f_plain.err; echo plain rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2f_launches.csv python bench.py $LEAN > gpurun_out/r2f_ncu_l.log 2>&1; echo launches rc=$?
timeout 1500 ncu --set full --clock-control none -k regex:'planes_gemm|lookup_|seg_|bn_|sort_|splitk|split_planes|absmax|relu_dropout|rank1|colsum|fc_grad|partials|adam_dense|records_unpack' --launch-skip 250 -c 90 -o /tmp/r2f_full -f python bench.py $LEAN > gpurun_out/r2f_ncu_f.log 2>&1; echo full rc=$?
ncu -i /tmp/r2f_full.ncu-rep --page raw --csv > gpurun_out/r2f_full_raw.csv 2>/dev/null; ls -la /tmp/r2f_full.ncu-rep gpurun_out/r2f_full_raw.csv

# ncu evidence of the headline step (1 GPU): launch list + --set full of every own kernel; only CSV / small files are kept
set -x
mkdir -p gpurun_out
LEAN="--steps 2 --warmup 3 --no-parity-check --no-other-configs --no-cpu-baseline --no-torch-eager --small-batch 0"
timeout 300 python bench.py $LEAN > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err; echo plain rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches.csv python bench.py $LEAN > gpurun_out/r2_ncu_l.log 2>&1; echo launches rc=$?
timeout 1500 ncu --set full --clock-control none -k regex:'planes_gemm|lookup_|seg_|bn_|sort_|splitk|split_planes|absmax|relu_dropout|rank1|colsum|fc_grad|partials' --launch-skip 250 -c 80 -o /tmp/r2_full -f python bench.py $LEAN > gpurun_out/r2_ncu_f.log 2>&1; echo full rc=$?
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null; ls -la /tmp/r2_full.ncu-rep gpurun_out/r2_full_raw.csv
FMT=fp16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:planes_gemm --launch-skip 3 -c 1 -o gpurun_out/r2_gemm_fp16 -f python scripts/gemm_one.py 65536 400 624 0 0 3 > gpurun_out/r2_ncu_g.log 2>&1; echo gemm rc=$?
ncu -i gpurun_out/r2_gemm_fp16.ncu-rep --page raw --csv > gpurun_out/r2_gemm_fp16_raw.csv 2>/dev/null
du -sh gpurun_out

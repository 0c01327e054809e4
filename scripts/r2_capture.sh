set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t43.log 2>&1; tail -5 gpurun_out/t43.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke43.log 2>&1; tail -2 gpurun_out/smoke43.log
timeout 900 python bench.py > gpurun_out/b43_n1.json 2> gpurun_out/b43_n1.err; tail -c 600 gpurun_out/b43_n1.json
LEAN="--steps 2 --warmup 3 --no-parity-check --no-other-configs --no-cpu-baseline --no-torch-eager --small-batch 0"
timeout 300 python bench.py $LEAN > gpurun_out/plain43.json 2> gpurun_out/plain43.err; echo plain rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches43.csv python bench.py $LEAN > gpurun_out/ncu_l43.log 2>&1; echo launches rc=$?
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'planes_gemm|lookup_|segment_|bn_|scatter|splitk|split_planes|relu_dropout|rank1|colsum|fc_grad' --launch-skip 300 -c 70 -o gpurun_out/r2_full43 -f python bench.py $LEAN > gpurun_out/ncu_f43.log 2>&1; echo full rc=$?
ls -la gpurun_out | tail -12

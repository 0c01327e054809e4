# final record of the round on one GPU: the driver's own commands
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/t_final.log; cat gpurun_out/t_final.log
python -c "import __graft_entry__ as G; G.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err ) 2>&1 | grep real
python scripts/show_bench.py gpurun_out/final_n1.json 2>/dev/null | sed -n '1,9p'

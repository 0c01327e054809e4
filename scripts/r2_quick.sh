mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/b61_n1.json 2> gpurun_out/b61_n1.err ) 2>&1 | tail -3
python scripts/show_bench.py gpurun_out/b61_n1.json 2>/dev/null
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b61_ref.json 2> gpurun_out/b61_ref.err ) 2>&1 | tail -3
cat gpurun_out/b61_ref.json | head -c 600

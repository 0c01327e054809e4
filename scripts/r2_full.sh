set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t49.log 2>&1; tail -5 gpurun_out/t49.log
timeout 900 python bench.py --no-other-configs > gpurun_out/b49_n1.json 2> gpurun_out/b49_n1.err; echo rc=$?
python scripts/show_bench.py gpurun_out/b49_n1.json > gpurun_out/b49_show.txt 2>&1; head -30 gpurun_out/b49_show.txt

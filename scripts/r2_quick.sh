mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_staging_and_adam.py -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --no-other-configs --no-cpu-baseline --no-torch-eager > gpurun_out/b56_n1.json 2> gpurun_out/b56_n1.err
tail -5 gpurun_out/b56_n1.err
python scripts/show_bench.py gpurun_out/b56_n1.json 2>/dev/null | head -30
python -c "
import json; d=json.load(open('gpurun_out/b56_n1.json')); print(d['e2e'])"

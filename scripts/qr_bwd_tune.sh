#!/bin/bash
# sweep the TINY (QR emb1-in-registers) backward launch knobs on the headline workload
for ki in 1 2; do for cta in 3 4 6 8; do
  RSB_TINY_KI=$ki RSB_TINY_CTAS_PER_SM=$cta python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-eager --small-batch 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ki=$ki cta=$cta', d['value'], d['kernels']['lookup_bwd_rows']['ms_avg'])"
done; done

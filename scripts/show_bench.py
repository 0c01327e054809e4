"""Summarise a bench.py JSON line: python scripts/show_bench.py file.json"""
import json, sys
d = json.load(open(sys.argv[1]))
print("value", d['value'], d['ms_per_step'], "e2e", d['e2e']['value'], "pipelined", (d['e2e'].get('pipelined_loop') or {}).get('value'))
print("parity", d.get('parity_check'))
print("clocks", d.get('clocks'))
for r in d['roofline_top3']:
    print("top3", {k: r.get(k) for k in ('kernel', 'bound', 'achieved', 'peak', 'frac', 'share_of_step', 'ms_per_launch')})
for r in d['roofline_hot_path']:
    print("hot ", {k: r.get(k) for k in ('kernel', 'achieved', 'frac', 'share_of_step', 'ms_per_launch')}, r.get('nvlink') and r['nvlink']['frac'])
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['share_of_step'])[:16]:
    print(f"   {k:28s} calls {v['calls_per_step']:5.1f} ms_avg {v['ms_avg']:.4f} share {v['share_of_step']:.3f}")
for k, v in (d.get('other_configs') or {}).items():
    if 'error' in v:
        print(k, v)
        continue
    print(k, v['value'], v['ms_per_step'], 'ref-loop', v['reference_trainer_loop']['value'])
    print("    top:", [(r['kernel'], r['frac'], r['share_of_step']) for r in v['roofline_top3']])
    print("    hot:", [(r['kernel'], r['frac'], r['ms_per_launch']) for r in v['roofline_hot_path']])
print("small", d.get('small_batch'))

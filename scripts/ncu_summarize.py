"""Turn ncu CSV output into the summaries kept under profiles/.

  launches:  python scripts/ncu_summarize.py launches <launches.csv> <steps_captured> > profiles/rN_launch_summary.md
             (csv from `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`)
  full:      python scripts/ncu_summarize.py full <raw.csv> > profiles/rN_ncu_full.json
             (csv from `ncu -i x.ncu-rep --page raw --csv`)
"""
from __future__ import annotations

import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict


def _rows(path):
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    return list(csv.DictReader(lines))


def short(name: str, n: int = 110) -> str:
    name = re.sub(r"\s+", " ", name)
    return name if len(name) <= n else name[:n]


def launches(path: str, steps: int):
    rows = _rows(path)
    per = defaultdict(lambda: [0.0, 0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = r["Kernel Name"]
        per[k][0] += us
        per[k][1] += 1
    total = sum(v[0] for v in per.values())

    def own(k):     # kernels compiled from recsys-benchmark_b200/csrc (namespaces rsb / pg + two free functions)
        return "rsb::" in k or "pg::" in k or k.startswith("rank1_bound_kernel") or "rsb" in k

    if steps <= 0:  # one forward gather per step
        steps = max(1, sum(v[1] for k, v in per.items() if "lookup_fwd_kernel" in k))
    ours = sum(v[0] for k, v in per.items() if own(k))
    print("| us/step | launches/step | share | kernel |")
    print("|---:|---:|---:|---|")
    for k, (us, n) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        if us / total < 0.001:
            continue
        mark = "**" if own(k) else ""
        print(f"| {us / steps:.1f} | {n / steps:.1f} | {100 * us / total:.1f}% | {mark}`{short(k)}`{mark} |")
    print()
    print(f"Total {total / steps:.0f} us/step over {steps} captured steps; hand-written / own-instantiated kernels "
          f"{ours / steps:.0f} us/step ({100 * ours / total:.1f}%).")


KEYS = OrderedDict([
    ("gpu__time_duration.sum", "time_ns"),
    ("dram__bytes_read.sum", "dram_read_bytes"),
    ("dram__bytes_write.sum", "dram_write_bytes"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_insts"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_hmma_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
])


def full(path: str):
    rows = _rows(path)
    # --page raw --csv: one row per launch, metrics as columns (second line = units)
    out = []
    for r in rows:
        name = r.get("Kernel Name")
        if not name or not r.get("ID", "").strip().isdigit():
            continue
        rec = {"kernel": short(name, 160)}
        for m, key in KEYS.items():
            v = r.get(m)
            if v in (None, "", "n/a"):
                continue
            try:
                rec[key] = float(v.replace(",", ""))
            except ValueError:
                rec[key] = v
        if "dram_read_bytes" in rec and "dram_write_bytes" in rec:
            rec["dram_bytes"] = rec["dram_read_bytes"] + rec["dram_write_bytes"]
        out.append(rec)
    units = {}
    if rows and not rows[0].get("ID", "").strip().isdigit():
        units = {KEYS[m]: rows[0].get(m) for m in KEYS if rows[0].get(m)}
    json.dump({"units": units, "launches": out}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]))
    else:
        full(sys.argv[2])

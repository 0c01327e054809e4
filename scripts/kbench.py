"""Micro-benchmark of the backward-side kernels at the headline shape (not part of the product or the
bench contract): sort_rows / segment_reduce_apply / CSR gather, CUDA events, L2 flushed between calls.

  python scripts/kbench.py [--rows criteo|roofline] [--qr 5]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import __graft_entry__ as G  # noqa: E402

G.build()
import recsys_benchmark_b200 as R  # noqa: E402
import recsys_benchmark_b200.functional as RF  # noqa: E402
from recsys_benchmark_b200 import _lib as L  # noqa: E402


def timeit(fn, reps=20, flush=None):
    ms = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
    ms.sort()
    return ms[len(ms) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="criteo")
    ap.add_argument("--qr", type=int, default=5)
    ap.add_argument("--batch", type=int, default=65536)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dims = bench.CRITEO_DIMS if args.rows == "criteo" else bench.ROOFLINE_DIMS
    n = sum(dims)
    x, _ = bench.make_batches(dims, args.batch, 1, 2023, torch.int64)[0]
    offsets = torch.tensor([0] + dims[:-1]).cumsum(0)
    rows = (x + offsets).to(dev).reshape(-1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for label, div in (("full", 0), (f"qr{args.qr}", args.qr)):
        nr = n if div == 0 else (n - 1) // div + 1
        out[f"sort_{label}"] = timeit(lambda: RF.sort_rows(rows, nr, key_div=div), flush=flush)
        sk, pm = RF.sort_rows(rows, nr, key_div=div)
        ref = torch.sort(rows // div if div else rows, stable=True)
        assert torch.equal(sk.long(), ref.values) and torch.equal(pm.long(), ref.indices), "sort mismatch"
        rg = torch.randn(rows.numel(), 16, device=dev)
        dst = torch.zeros(nr, 16, device=dev)
        out[f"segment_dense_{label}"] = timeit(
            lambda: RF.segment_reduce_apply(L.APPLY_DENSE, sk, pm, rg, dst), flush=flush)
        m, v = torch.zeros_like(dst), torch.zeros_like(dst)
        out[f"segment_adam_{label}"] = timeit(
            lambda: RF.segment_reduce_apply(L.APPLY_SPARSE_ADAM, sk, pm, rg, dst, m, v, lr=1e-3, step=3), flush=flush)
        out[f"unique_rows_{label}"] = int(torch.unique(ref.values).numel())
    w = torch.randn(n, 16, device=dev) * (torch.rand(n, 16, device=dev) < 0.2)
    pe = R.PrunedEmbedding.from_weight(w, compact=True)
    xd = x.to(dev).int()
    od = offsets.to(dev)
    fc = torch.randn(n, 1, device=dev)
    b0 = torch.zeros(1, device=dev)
    out["csr_lookup_keep20"] = timeit(lambda: pe.lookup(xd, od, fc, b0), flush=flush)
    van = R.get_embedding({"name": "vanilla"}, dims, 16).to(dev)
    with torch.no_grad():
        out["dense_lookup"] = timeit(lambda: van.lookup(xd, od, fc, b0), flush=flush)
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()})


if __name__ == "__main__":
    main()

"""One GEMM shape a few times (for ncu): python scripts/gemm_one.py M N K a_mn b_mn [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as G
G.build()
from recsys_benchmark_b200 import planes as P
m, n, k, a_mn, b_mn = [int(v) for v in sys.argv[1:6]]
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
a = torch.randn((k, m) if a_mn else (m, k), device="cuda")
b = torch.randn((k, n) if b_mn else (n, k), device="cuda")
pa, pb = P.split(a), P.split(b)
out = torch.empty(m, n, device="cuda")
for _ in range(iters):
    P.gemm(pa, pb, m, n, k, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), out=out)
torch.cuda.synchronize()
print("done")

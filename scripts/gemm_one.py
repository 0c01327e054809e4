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
FMT = P.FP16X2 if os.environ.get("FMT", "") == "fp16" else P.BF16X3
pa, pb = P.split(a, fmt=FMT), P.split(b, fmt=FMT)
out = torch.empty(m, n, device="cuda")
for _ in range(3):
    P.gemm(pa, pb, m, n, k, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), out=out)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(iters):
    P.gemm(pa, pb, m, n, k, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), out=out)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / iters
A = (a.double().t() if a_mn else a.double())[:2048]
B = b.double() if b_mn else b.double().t()
ref = A @ B
err = float((out[:2048].double() - ref).abs().max() / ref.abs().max()) if not a_mn else -1.0
print(f"{m}x{n}x{k} a_mn={a_mn} b_mn={b_mn} fmt={os.environ.get('FMT', 'bf16')} pairs={os.environ.get('RSB_GEMM_PAIRS', '0')} "
      f"drain16={os.environ.get('RSB_GEMM_DRAIN_FP16', '2')}: {ms:.4f} ms = {2.0 * m * n * k / ms / 1e9:.1f} TFLOP/s fp32-eq, err {err:.2e}")

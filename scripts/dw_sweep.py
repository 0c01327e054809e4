"""Scratch: weight-gradient GEMM variants (split-K factor, operand layout) on the MLP shapes."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G
G.build()
import recsys_benchmark_b200.linalg as LA

dev = torch.device("cuda:0")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

B = 65536
for (m, n) in [(400, 624), (400, 400)]:
    gz = torch.randn(B, m, device=dev); x = torch.randn(B, n, device=dev)
    ref = (gz.double().t() @ x.double())
    fl = 2.0 * B * m * n
    print(f"dW [{m}x{n}], K={B}: auto split = {LA._split_for(B, m, n)}")
    for sk in (8, 16, 32, 64, 128, 256):
        ms = t(lambda: LA.gemm(gz, x, trans_a=True, split_k=sk))
        err = ((LA.gemm(gz, x, trans_a=True, split_k=sk).double() - ref).norm() / ref.norm()).item()
        print(f"  cr split {sk:4d}: {ms:.3f} ms {fl / ms / 1e9:.1f} TF/s err {err:.1e}")
    gzt = gz.t().contiguous()
    for sk in (16, 32, 64, 128):
        ms = t(lambda: LA.gemm(gzt, x, split_k=sk))
        print(f"  rr (A pre-transposed) split {sk:4d}: {ms:.3f} ms {fl / ms / 1e9:.1f} TF/s (+ transpose {t(lambda: gz.t().contiguous()):.3f} ms)")
    # swapped roles: dW^T = x^T @ gz
    for sk in (32, 64, 128):
        ms = t(lambda: LA.gemm(x, gz, trans_a=True, split_k=sk))
        print(f"  cr swapped (dW^T) split {sk:4d}: {ms:.3f} ms {fl / ms / 1e9:.1f} TF/s")
    ms = t(lambda: gz.t() @ x)
    print(f"  cuBLAS fp32: {ms:.3f} ms {fl / ms / 1e9:.1f} TF/s")

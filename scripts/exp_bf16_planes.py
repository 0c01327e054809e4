"""Experiment: fp32-accurate GEMM as ONE bf16 tensor-core GEMM over K-concatenated 3-term splits
(6 plane pairs), using the library bf16 GEMM with fp32 output.  Prints time and error vs fp64."""
import torch
dev = "cuda"
def split3(x):
    a0 = x.to(torch.bfloat16); r = x - a0.float()
    a1 = r.to(torch.bfloat16); r = r - a1.float()
    a2 = r.to(torch.bfloat16)
    return a0, a1, a2
def t(fn, n=10):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
for (M, N, K) in [(65536, 400, 624), (65536, 400, 400), (65536, 624, 400)]:
    torch.manual_seed(0)
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.05
    ref = x.double() @ w.double().t()
    a0, a1, a2 = split3(x); b0, b1, b2 = split3(w)
    A6 = torch.cat([a0, a0, a1, a0, a1, a2], 1).contiguous()     # [M, 6K]
    B6 = torch.cat([b0, b1, b0, b2, b1, b0], 1).contiguous()     # [N, 6K]
    try:
        f = lambda: torch.mm(A6, B6.t(), out_dtype=torch.float32)
        y = f()
    except Exception as ex:
        print("out_dtype unsupported:", ex); break
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    e32 = float(((x @ w.t()).double() - ref).abs().max() / ref.abs().max())
    ms = t(f)
    ms_split = t(lambda: torch.cat(split3(x), 1))
    print(f"M={M} N={N} K={K}: bf16x6 mm {ms:.3f} ms ({2*M*N*K/ms/1e9:.0f} TF/s fp32-eq), relerr {err:.2e} (cuBLAS fp32 {e32:.2e}); torch split of A ~{ms_split:.3f} ms; fp32 mm {t(lambda: x @ w.t()):.3f} ms")
    A3 = torch.cat([a0, a1, a2], 1).contiguous(); 
    ms3 = t(lambda: torch.mm(A3, torch.cat([b0, b0, b0], 1).t(), out_dtype=torch.float32))
    print(f"    3K-wide bf16 mm alone: {ms3:.3f} ms")

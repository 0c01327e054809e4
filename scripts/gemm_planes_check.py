"""Correctness + speed of rsb_gemm_planes on a B200: every operand majorness, edge shapes, batches, split-K; error
against fp64; CUDA-event timing of the MLP shapes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as G
G.build()
from recsys_benchmark_b200 import planes as P

DEV = "cuda:0"
torch.manual_seed(0)
FMT = P.FP16X2 if "--fp16" in sys.argv else P.BF16X3


def err(got, ref):
    return float((got.double() - ref).abs().max() / ref.abs().max())


def check(m, n, k, a_mn, b_mn, split_k=0, bias=True):
    a = torch.randn((k, m) if a_mn else (m, k), device=DEV)
    b = torch.randn((k, n) if b_mn else (n, k), device=DEV)
    bi = torch.randn(n, device=DEV) if bias else None
    pa, pb = P.split(a, fmt=FMT), P.split(b, fmt=FMT)
    ea = err(pa.float(), a.double())
    out = P.gemm(pa, pb, m, n, k, a_mn_major=a_mn, b_mn_major=b_mn, bias=bi, split_k=split_k)
    A = a.double().t() if a_mn else a.double()
    B = b.double() if b_mn else b.double().t()
    ref = A @ B + (bi.double() if bias else 0)
    cub = ((a.t() if a_mn else a) @ (b if b_mn else b.t())) + (bi if bias else 0)
    return err(out, ref), err(cub, ref), ea


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


results = {"checks": [], "timing": []}
quick = "--quick" in sys.argv
shapes = [(128, 16, 32), (128, 128, 64), (256, 400, 624), (2048, 400, 400), (1000, 64, 352), (4096, 256, 352),
          (520, 352, 256), (132, 12, 20), (8192, 400, 176), (8192, 176, 400), (4, 4, 4), (65536, 400, 624)]
if quick:
    shapes = shapes[:3]
for (m, n, k) in shapes:
    for a_mn, b_mn in ((False, False), (False, True), (True, True), (True, False)):
        if (a_mn and m % 8) or (b_mn and n % 8):
            continue
        try:
            e1, e2, ea = check(m, n, k, a_mn, b_mn)
            results["checks"].append(dict(m=m, n=n, k=k, a_mn=a_mn, b_mn=b_mn, err=e1, cublas=e2, split_err=ea))
            print(f"m={m} n={n} k={k} a_mn={int(a_mn)} b_mn={int(b_mn)}  err {e1:.2e}  (cuBLAS fp32 {e2:.2e}, planes {ea:.1e})",
                  "OK" if e1 < 1e-6 else "BAD", flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f"m={m} n={n} k={k} a_mn={int(a_mn)} b_mn={int(b_mn)}  EXC {ex}", flush=True)
            results["checks"].append(dict(m=m, n=n, k=k, a_mn=a_mn, b_mn=b_mn, exc=str(ex)))
if not quick:
    # weight-gradient shapes (K = batch), split-K
    for (m, n, k) in [(400, 624, 65536), (400, 400, 65536), (624, 400, 65536), (400, 176, 8192)]:
        for sk in (0, 1, 7):
            e1, e2, _ = check(m, n, k, True, True, split_k=sk, bias=False)
            results["checks"].append(dict(m=m, n=n, k=k, a_mn=True, b_mn=True, split_k=sk, err=e1, cublas=e2))
            print(f"dW m={m} n={n} k={k} split_k={sk}  err {e1:.2e} (cuBLAS {e2:.2e})", "OK" if e1 < 1e-6 else "BAD", flush=True)

    # timing: the MLP shapes of the headline step
    for name, (m, n, k, a_mn, b_mn) in {"fwd1 65536x400x624": (65536, 400, 624, False, False),
                                        "fwd2 65536x400x400": (65536, 400, 400, False, False),
                                        "dX1 65536x624x400": (65536, 624, 400, False, False),
                                        "dW1 400x624x65536": (400, 624, 65536, True, True),
                                        "dW2 400x400x65536": (400, 400, 65536, True, True)}.items():
        a = torch.randn((k, m) if a_mn else (m, k), device=DEV)
        b = torch.randn((k, n) if b_mn else (n, k), device=DEV)
        pa, pb = P.split(a, fmt=FMT), P.split(b, fmt=FMT)
        out = torch.empty(m, n, device=DEV)
        ms = timeit(lambda: P.gemm(pa, pb, m, n, k, a_mn_major=a_mn, b_mn_major=b_mn, out=out))
        ms_split = timeit(lambda: P.split(a, fmt=FMT))
        ms_cublas = timeit(lambda: torch.matmul(a.t() if a_mn else a, b if b_mn else b.t(), out=out))
        tf = 2.0 * m * n * k / ms / 1e9
        results["timing"].append(dict(name=name, ms=ms, tflops_fp32_equiv=tf, ms_split_a=ms_split, ms_cublas_fp32=ms_cublas))
        print(f"{name}: {ms:.4f} ms = {tf:.1f} TFLOP/s fp32-equivalent ({6 * tf:.0f} bf16 MMA TFLOP/s); split(A) {ms_split:.4f} ms; "
              f"cuBLAS fp32 {ms_cublas:.4f} ms", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open("gpurun_out/gemm_planes_check%s.json" % ("_fp16" if FMT == P.FP16X2 else ""), "w"), indent=1)

"""In-tree build of librsb.so (hand-written sm_100a kernels + the C ABI of include/rsb.h).

`python recsys-benchmark_b200/build.py` or `__graft_entry__.build()`.  Uses nvcc
directly (cross-compiles without a GPU); objects are cached per source by content hash;
the .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
GEMM = os.path.join(CSRC, "gemm")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "librsb.so")
STAMP = os.path.join(PKG, ".librsb.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr", "-diag-suppress", "20012",
]


def sources():
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    srcs += sorted(os.path.join(GEMM, f) for f in os.listdir(GEMM) if f.endswith(".cu"))
    return srcs


def _deps(src):
    deps = [src]
    d = os.path.dirname(src)
    deps += sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cuh"))
    text = "".join(open(f).read() for f in deps)
    if "rsb.h" in text:
        deps.append(os.path.join(ROOT, "include", "rsb.h"))
    return deps


def _hash(files, extra=""):
    h = hashlib.sha256()
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(extra.encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    jobs, objs, stamps = [], [], []
    for src in sources():
        is_gemm = os.path.dirname(src) == GEMM
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
        if is_gemm:
            inc += ["-I", GEMM]
        dig = _hash(_deps(src), " ".join(NVCC_FLAGS))
        stamp = obj + ".sha"
        stamps.append(dig)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
            continue
        cmd = [nvcc, *NVCC_FLAGS, *inc, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        jobs.append((src, stamp, dig, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                                       text=True)))
    failed = False
    for src, stamp, dig, p in jobs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {os.path.basename(src)} (rc={p.returncode})\n{out}\n")
        if p.returncode == 0:
            with open(stamp, "w") as fh:
                fh.write(dig)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building librsb.so")
    link_dig = hashlib.sha256("".join(stamps).encode()).hexdigest()
    if jobs or force or not os.path.exists(LIB) or not os.path.exists(STAMP) or open(STAMP).read().strip() != link_dig:
        subprocess.check_call([nvcc, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC", "-lcudart", "-lcuda",
                               "-Wno-deprecated-gpu-targets"])
        with open(STAMP, "w") as fh:
            fh.write(link_dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

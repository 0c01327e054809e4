"""Diagnose the e2e input-staging variants (scratch; not part of the bench contract)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import __graft_entry__ as G
G.build()
import recsys_benchmark_b200 as R
from recsys_benchmark_b200.data import DevicePrefetcher, DeferredScalar

dev = torch.device("cuda:0")
wl = bench.WORKLOADS["deepfm_qr_criteo"]
dims = wl["dims"]; b = 65536
torch.manual_seed(0)
model = R.get_ctr_model(dims, dict(num_factor=16, hidden_sizes=[400, 400, 400], p_dropout=0.5, use_batchnorm=False,
                                   embedding_config=dict(wl["emb"]))).to(dev).train()
opts = R.get_optimizers(model, dict(wl["opt"]))
crit = torch.nn.BCEWithLogitsLoss()
pool = bench.make_batches(dims, b, 8, 1, torch.int32)
host = [(x.pin_memory(), y.pin_memory()) for x, y in pool]
devp = [(x.to(dev), y.to(dev)) for x, y in pool]

def step(x, y):
    loss = crit(model(x), y)
    for o in opts: o.zero_grad()
    loss.backward()
    for o in opts: o.step()
    return loss

def timeit(fn, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(n); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def resident(n):
    for i in range(n): step(*devp[i % 8])
def resident_item(n):
    for i in range(n): step(*devp[i % 8]).item()
def blocking(n):
    for i in range(n):
        x, y = host[i % 8]
        step(x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)).item()
def prefetch(n):
    for x, y in DevicePrefetcher((host[i % 8] for i in range(n)), dev): step(x, y).item()
def prefetch_deferred(n):
    r = DeferredScalar(dev)
    for x, y in DevicePrefetcher((host[i % 8] for i in range(n)), dev): r.push(step(x, y))
    r.flush()
def blocking_deferred(n):
    r = DeferredScalar(dev)
    for i in range(n):
        x, y = host[i % 8]
        r.push(step(x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)))
    r.flush()

for f in (resident, resident_item, blocking, prefetch, prefetch_deferred, blocking_deferred):
    f(5)
    print(f.__name__, [round(timeit(f), 3) for _ in range(4)], flush=True)

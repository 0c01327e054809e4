"""Host-side cost of one training step of the headline workload: enqueue time per step (no sync inside) and a cProfile."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as G
G.build()
import recsys_benchmark_b200 as R
from recsys_benchmark_b200.sharded import ShardedDeepFM
import bench as B

dev = torch.device("cuda:0")
torch.manual_seed(2023)
wl = B.WORKLOADS["deepfm_full_criteo_sharded"]
dims = wl["dims"]
model = ShardedDeepFM(dims, 16, [400, 400, 400], p_dropout=0.5, use_batchnorm=True, embedding_config=dict(wl["emb"])).to(dev)
model.train()
opts = R.get_optimizers(model, dict(wl["opt"]))
crit = torch.nn.BCEWithLogitsLoss()
pool = B.make_batches(dims, 65536, 4, 2023, torch.int32, "uniform")
pool = [(x.to(dev), y.to(dev)) for x, y in pool]


def step(x, y):
    loss = crit(model(x), y)
    for o in opts:
        o.zero_grad()
    loss.backward()
    model.sync_gradients()
    for o in opts:
        o.step()
    model.finish_step()
    return loss


for i in range(10):
    step(*pool[i % 4])
torch.cuda.synchronize()
for trial in range(3):
    t0 = time.perf_counter()
    for i in range(8):
        step(*pool[i % 4])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3 * (t1 - t0) / 8:.3f} ms/step, with drain {1e3 * (t2 - t0) / 8:.3f} ms/step", flush=True)
pr = cProfile.Profile()
pr.enable()
for i in range(8):
    step(*pool[i % 4])
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print(s.getvalue()[:9000])

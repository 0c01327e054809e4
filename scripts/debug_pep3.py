import sys, os, pathlib, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import refimport
import tests.test_gpu_reference_parity as T
import __graft_entry__ as G
G.build()
import recsys_benchmark_b200 as R
refimport.activate()
import src.models as ref_models
DEV = "cuda:0"
dims, b, cfg, opt_cfg, tweak, _ = T.CASES["kdd_pep_feature_dim"]
td = pathlib.Path(tempfile.mkdtemp())
torch.manual_seed(2023)
ref = T._build(ref_models, dims, cfg, td / "ref").to(DEV)
T._tweak(ref, tweak)
ours = T._build(R, dims, cfg, td / "ours")
ours.load_state_dict(ref.state_dict(), strict=True)
ours.to(DEV)
x, y = T._batch(dims, b, 7)
crit = torch.nn.BCEWithLogitsLoss()
acts = {"ref": {}, "ours": {}}
def mk(name, idx):
    def hook(m, inp, out):
        acts[name][f"bn{idx}_in"] = inp[0].detach().clone()
        acts[name][f"bn{idx}_out"] = out.detach().clone()
        out.register_hook(lambda g: acts[name].__setitem__(f"bn{idx}_gout", g.clone()))
        inp[0].register_hook(lambda g: acts[name].__setitem__(f"bn{idx}_gin", g.clone()))
    return hook
for name, m in (("ref", ref), ("ours", ours)):
    for idx, mod in enumerate(m._deep_branch):
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.register_forward_hook(mk(name, idx))
for name, m in (("ref", ref), ("ours", ours)):
    m.train()
    crit(m(x), y.float()).backward()
def rel(a, b_):
    return float((a - b_).abs().max() / b_.abs().max())
for k in sorted(acts["ref"]):
    a, r = acts["ours"][k], acts["ref"][k]
    extra = ""
    if k.endswith("_out"):
        extra = f" relu sign mismatches {int(((a > 0) != (r > 0)).sum())}, |ref|<1e-5: {int((r.abs() < 1e-5).sum())}"
    print(k, "rel diff", rel(a, r), "rows differing >1e-5*max:", int(((a - r).abs() > 1e-5 * r.abs().max()).any(1).sum()), extra)

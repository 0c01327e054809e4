import sys, os, pathlib, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import refimport
import tests.test_gpu_reference_parity as T
import __graft_entry__ as G
G.build()
import recsys_benchmark_b200 as R
refimport.activate()
import src.models as ref_models, src.models.deepfm as ref_deepfm
DEV = "cuda:0"
print("tf32 matmul:", torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
torch.manual_seed(5)
cfg = dict(embedding_config={"name": "vanilla", "sparse": True})
td = pathlib.Path(tempfile.mkdtemp())
ref = T._build(ref_models, T.CRITEO_DIMS, cfg, td).to(DEV)
ours = T._build(R, T.CRITEO_DIMS, cfg, td)
ours.load_state_dict(ref.state_dict(), strict=True)
ours.to(DEV)
fused = len(sys.argv) > 1 and sys.argv[1] == "fused"
opt_cfg = dict(learning_rate=1e-3, weight_decay=1e-6, sparse=True)
o_ref = ref_deepfm.get_optimizers(ref, dict(opt_cfg))
o_ours = R.get_optimizers(ours, dict(opt_cfg, fused_sparse=fused))
crit = torch.nn.BCEWithLogitsLoss()
x, y = T._batch(T.CRITEO_DIMS, 2048, 100)
stash = {}
def ref_hook(mod, inp):
    inp[0].register_hook(lambda g: stash.__setitem__("ref_gdeep", g.clone()))
ref._deep_branch.register_forward_pre_hook(ref_hook)
orig = ours.embedding.lookup
def lookup(*a, **k):
    emb, yy = orig(*a, **k)
    emb.register_hook(lambda g: stash.__setitem__("ours_gdeep", g.clone()))
    yy.register_hook(lambda g: stash.__setitem__("ours_gy", g.clone()))
    return emb, yy
ours.embedding.lookup = lookup
outs = {}
for name, m, opts in (("ref", ref, o_ref), ("ours", ours, o_ours)):
    m.train()
    out = m(x)
    loss = crit(out, y.float())
    for o in opts:
        o.zero_grad()
    loss.backward()
    outs[name] = out.detach()
def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())
print("logits rel diff", rel(outs["ours"], outs["ref"]))
print("g_deep rel diff", rel(stash["ours_gdeep"].reshape(2048, -1), stash["ref_gdeep"]))
g_ref = ref.embedding.get_weight().grad
if fused:
    (table, rows, rg, pair), = o_ours[0]._pending
else:
    g = ours.embedding.get_weight().grad
    rg = g._values()
print("per-lookup grad rel diff", rel(rg.reshape(-1, 16), g_ref._values()), "max", float(g_ref._values().abs().max()))
d = (rg.reshape(-1, 16) - g_ref._values()).abs()
print("share of elements off by >1e-5*max:", float((d > 1e-5 * g_ref._values().abs().max()).float().mean()))
for k, p in ref.named_parameters():
    po = dict(ours.named_parameters())[k]
    if p.grad is not None and po.grad is not None and not p.grad.is_sparse:
        print("  grad", k, rel(po.grad, p.grad))

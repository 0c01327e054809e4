import sys, os, pathlib, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import refimport
import tests.test_gpu_reference_parity as T
import __graft_entry__ as G
G.build()
import recsys_benchmark_b200 as R
import recsys_benchmark_b200.functional as RF
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
stash = {}
def _fh(m, i, o):
    o.register_hook(lambda g: stash.__setitem__("ref_gemb", g.clone()))
ref.embedding.register_forward_hook(_fh)
orig = ours.embedding.lookup
def lookup(*a, **k):
    emb, yy = orig(*a, **k)
    stash["emb"] = emb.detach().clone()
    emb.register_hook(lambda g: stash.__setitem__("ours_gdeep", g.clone()))
    yy.register_hook(lambda g: stash.__setitem__("ours_gy", g.clone()))
    return emb, yy
ours.embedding.lookup = lookup
for name, m in (("ref", ref), ("ours", ours)):
    m.train()
    out = m(x)
    crit(out, y.float()).backward()
emb = stash["emb"]                       # [B,F,D]
S = emb.sum(1, keepdim=True)
exp_gemb = stash["ours_gdeep"].reshape(emb.shape) + stash["ours_gy"].reshape(-1, 1, 1) * (S - emb)
ref_gemb = stash["ref_gemb"]
print("expected-from-ours-upstream vs ref g_emb: max abs diff", float((exp_gemb - ref_gemb).abs().max()), "max", float(ref_gemb.abs().max()))
rows = (x + ref.offsets)
w, s = ref.embedding.emb.weight.detach(), ref.embedding.s.detach()
keep = (w[rows].abs() - torch.sigmoid(s[rows])) > 0
gw_ours = ours.embedding.emb.weight.grad[rows]          # per lookup (valid for rows that appear once)
uniq, cnt = torch.unique(rows, return_counts=True)
once = torch.isin(rows, uniq[cnt == 1])
sel = keep & once.unsqueeze(-1)
d = ((gw_ours - ref_gemb).abs() * sel)
print("kernel gw vs ref g_emb on kept, once-only elements: max abs diff", float(d.max()))
bad = d > 1e-5 * float(ref_gemb.abs().max())
print("bad elements", int(bad.sum()), "bad lookups", int(bad.any(-1).sum()), "bad samples", int(bad.any(-1).any(-1).sum()))
bs, fs, ds = torch.nonzero(bad, as_tuple=True)
from collections import Counter
print("fields of bad lookups", sorted(Counter(fs.tolist()).items()))
print("dims of bad", sorted(Counter(ds.tolist()).items()))
print("samples (first 20)", sorted(set(bs.tolist()))[:20])
for i in range(min(6, len(bs))):
    bb, ff, dd = int(bs[i]), int(fs[i]), int(ds[i])
    print(f"  b={bb} f={ff} d={dd} ours={float(gw_ours[bb,ff,dd]):.6e} ref={float(ref_gemb[bb,ff,dd]):.6e} exp={float(exp_gemb[bb,ff,dd]):.6e} "
          f"gdeep={float(stash['ours_gdeep'].reshape(emb.shape)[bb,ff,dd]):.6e} gy={float(stash['ours_gy'][bb]):.6e} S={float(S[bb,0,dd]):.6e} e={float(emb[bb,ff,dd]):.6e}")

# ---- isolate the dX GEMM of the first layer on the actual data ----
import recsys_benchmark_b200.linalg as LA
cap = {}
orig_gemm = LA.gemm
def spy(a, b_, **kw):
    out = orig_gemm(a, b_, **kw)
    if tuple(a.shape) == (8192, 400) and tuple(b_.shape) == (400, 176) and not kw.get("trans_a") and not kw.get("trans_b"):
        cap["a"], cap["b"], cap["out"] = a.detach().clone(), b_.detach().clone(), out.detach().clone()
    return out
LA.gemm = spy
ours.zero_grad()
crit(ours(x), y.float()).backward()
torch.cuda.synchronize()
a, b_, out = cap["a"], cap["b"], cap["out"]
ref64 = a.double() @ b_.double()
again = orig_gemm(a, b_)
cub = a @ b_
print("captured dX gemm: in-backward err", float((out.double() - ref64).abs().max() / ref64.abs().max()),
      "standalone err", float((again.double() - ref64).abs().max() / ref64.abs().max()),
      "cublas err", float((cub.double() - ref64).abs().max() / ref64.abs().max()))
print("a: finite", bool(torch.isfinite(a).all()), "max", float(a.abs().max()), "min nonzero", float(a[a != 0].abs().min()), "zeros share", float((a == 0).float().mean()),
      "stride", a.stride(), "ptr%256", a.data_ptr() % 256)
bad = ((out.double() - ref64).abs() > 1e-5 * ref64.abs().max())
print("bad elements in-backward", int(bad.sum()), "rows", int(bad.any(1).sum()), "cols", sorted(set(torch.nonzero(bad)[:, 1].tolist()))[:40])

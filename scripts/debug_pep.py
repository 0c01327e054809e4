"""Debug: PEP feature_dim s-gradient discrepancy at the KDD shape (model-level, vs the reference on the same GPU)."""
import copy, sys, os, tempfile
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
for early in (True,):
    RF.EARLY_SORT = early
    with tempfile.TemporaryDirectory() as td:
        import pathlib
        torch.manual_seed(2023)
        ref = T._build(ref_models, dims, cfg, pathlib.Path(td) / "ref").to(DEV)
        T._tweak(ref, tweak)
        ours = T._build(R, dims, cfg, pathlib.Path(td) / "ours")
        ours.load_state_dict(ref.state_dict(), strict=True)
        ours.to(DEV)
    x, y = T._batch(dims, b, 7)
    crit = torch.nn.BCEWithLogitsLoss()
    gs = {}
    for name, m in (("ref", ref), ("ours", ours)):
        m.train()
        torch.manual_seed(12)
        out = m(x)
        loss = crit(out, y.float())
        loss.backward()
        gs[name] = (m.embedding.s.grad.clone(), m.embedding.emb.weight.grad.clone())
    d = (gs["ours"][0] - gs["ref"][0]).abs()
    tol = 1e-5 * float(gs["ref"][0].abs().max()) + 1e-5 * gs["ref"][0].abs()
    bad = (d > tol)
    rows = (x + ref.offsets).reshape(-1)
    uniq, cnt = torch.unique(rows, return_counts=True)
    dup_rows = set(uniq[cnt > 1].tolist())
    bad_rows = torch.nonzero(bad.any(1)).reshape(-1).tolist()
    print(f"early_sort={early}: bad elements {int(bad.sum())}, bad rows {len(bad_rows)}, of which duplicated in batch "
          f"{sum(r in dup_rows for r in bad_rows)}; duplicated rows total {len(dup_rows)}")
    dw = (gs["ours"][1] - gs["ref"][1]).abs()
    print("   weight grad max diff", float(dw.max()), "max", float(gs["ref"][1].abs().max()))
    top = torch.topk(d.reshape(-1), 8).indices.tolist()
    for r in [t // 16 for t in top]:
        c = int(cnt[(uniq == r).nonzero()[0, 0]])
        j = int(bad[r].nonzero()[0, 0])
        print("   row", r, "count", c, "dim", j, "ours", float(gs["ours"][0][r, j]), "ref", float(gs["ref"][0][r, j]),
              "w", float(ref.embedding.emb.weight[r, j]), "s", float(ref.embedding.s[r, j]),
              "gw ours", float(gs["ours"][1][r, j]), "gw ref", float(gs["ref"][1][r, j]))

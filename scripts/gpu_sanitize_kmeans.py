"""compute-sanitizer target: one small run of each k-means method through tclip_kmeans_run in the forms the driver can take
(Cholesky / triangular + chained, features as coordinates, feature-space fallback), shapes with ragged class tiles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import torch
from tclip_b200 import ops, tasks

dev = torch.device("cuda:0")
for (K, D, n) in ((131, 300, 17), (260, 128, 40), (30, 24, 75), (20, 40, 100), (1000, 1024, 75)):
    T = 2
    td, _ = tasks.make_zero_shot_batch(T, K, n_query=n, seed=3, softmax_feature=False, embed_dim=D)
    x = td["x_q"].to(dev)
    g = torch.Generator().manual_seed(1)
    u0 = torch.softmax(4.0 * torch.randn(T, n, K, generator=g), dim=-1).to(dev)
    for method in (ops.KMEANS_SOFT, ops.KMEANS_GAUSS, ops.KMEANS_HARD):
        res = ops.kmeans_run(x, u0.clone(), method, 3, 30.0, lambd=float(int(K / 5) * n), want_w=True)
        torch.cuda.synchronize()
        assert torch.isfinite(res["u"]).all()
    print("ok", K, D, n, flush=True)

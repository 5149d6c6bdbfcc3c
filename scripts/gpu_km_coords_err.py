"""Measurement: error of the k-means loop in sample coordinates (csrc/kmeans_run.cu) and of the feature-space stage kernels
against a float64 evaluation of the same loop (torch, feature space), on the cases of
tests/test_gpu_kmeans.py::test_sample_coordinates_equal_feature_space.  TCLIP_KM_TRI=0/1 selects the dense / triangular
treatment of the Cholesky factor."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import torch
from tclip_b200 import ops, tasks

dev = torch.device("cuda:0")
EPS = 1e-15


def loop64(x, u, method, iters, temperature, lambd):
    x = x.double(); u = u.double(); n = x.shape[1]
    v = torch.zeros(u.shape[0], u.shape[2], dtype=torch.float64, device=x.device)
    w = None
    if method != ops.KMEANS_HARD:
        w = torch.einsum("tnk,tnd->tkd", u, x) / u.sum(1).clamp_min(EPS).unsqueeze(-1)
    for _ in range(iters):
        cs = u.sum(1)
        wn = torch.einsum("tnk,tnd->tkd", u, x) / cs.clamp_min(EPS).unsqueeze(-1)
        if method == ops.KMEANS_HARD:
            w = torch.where((cs > EPS).unsqueeze(-1), wn, torch.zeros_like(wn))
        else:
            w = torch.where((cs > EPS).unsqueeze(-1), wn, w)
        d2 = ((x.unsqueeze(2) - w.unsqueeze(1)) ** 2).sum(-1)
        if method == ops.KMEANS_HARD:
            u = torch.nn.functional.one_hot(torch.softmax(d2, -1).argmin(-1), u.shape[2]).double()
        else:
            l = temperature * (-0.5 * d2)
            if method == ops.KMEANS_GAUSS:
                l = l + (lambd * v).unsqueeze(1) / n
            u = torch.softmax(l, -1)
            if method == ops.KMEANS_GAUSS:
                v = torch.log(u.sum(1) / n + EPS) + 1
    return u


for (K, D, n, seed) in ((60, 256, 75, 5), (40, 90, 33, 6), (20, 200, 96, 9), (100, 512, 75, 11)):
    td, _ = tasks.make_zero_shot_batch(3, K, n_query=n, seed=seed, softmax_feature=False, embed_dim=D)
    x = td["x_q"].to(dev)
    x[1, 5] = x[1, 2]
    x[2, 1] = 0.5 * (x[2, 0] + x[2, 3])
    g = torch.Generator().manual_seed(seed)
    u0 = torch.softmax(4.0 * torch.randn(3, n, K, generator=g), dim=-1).to(dev)
    for name, method in (("soft", ops.KMEANS_SOFT), ("gauss", ops.KMEANS_GAUSS), ("hard", ops.KMEANS_HARD)):
        lam = float(int(K / 5) * n)
        res = ops.kmeans_run(x, u0.clone(), method, 5, 30.0, lambd=lam, want_w=True)
        u, v = u0.clone(), torch.zeros(3, K, device=dev)
        w = None if method == ops.KMEANS_HARD else ops.kmeans_centroids(u, x, None)
        for _ in range(5):
            w = ops.kmeans_centroids(u, x, w, keep_old=(method != ops.KMEANS_HARD))
            u, labels = ops.kmeans_assign(x, w, method, 30.0, v=v, lambd=lam)
            if method == ops.KMEANS_GAUSS:
                _, v, _ = ops.colsum_v(u, want_v=True, want_live=False)
        u64 = loop64(x, u0, method, 5, 30.0, lam)
        e_c = (res["u"].double() - u64).abs().max().item()
        e_f = (u.double() - u64).abs().max().item()
        e_x = (res["u"] - u).abs().max().item()
        print(f"K={K} D={D} n={n} {name}: |coords - f64| {e_c:.2e}  |feature - f64| {e_f:.2e}  |coords - feature| {e_x:.2e}")

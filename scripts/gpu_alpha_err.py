"""Scratch: alpha error of the GPU path vs the fp64 oracle for the library named by $TCLIP_LIB (error attribution)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import logging; logging.disable(logging.INFO)
import torch
from tclip_b200 import tasks
from tclip_b200.methods.dirichlet import EM_DIRICHLET, HARD_EM_DIRICHLET
from oracle import restated as R
from oracle.ref_loader import make_args
dev = torch.device("cuda:0")
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
cache = {}
for K, T, iters, hard, seed in [(20, 5, 6, False, 0), (100, 8, 8, False, 3), (100, 8, 10, True, 4)]:
    td, _ = tasks.make_zero_shot_batch(T, K, seed=seed)
    m = (HARD_EM_DIRICHLET if hard else EM_DIRICHLET)(model=None, device=dev, log_file=None, args=make_args(K, iters=iters, mm_mode="dense"))
    m.run_task({k: v.clone() for k, v in td.items()})
    r32 = R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=hard)
    r64 = R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=hard, dtype=torch.float64)
    a = m.alpha.cpu()
    print(os.environ.get("TCLIP_LIB", "product")[-12:], K, hard, "mm", m.mm_iters.cpu().tolist() == r32.mm_iters,
          "gpu-vs-64", ["%.1e" % rel(a[t], r64.alpha[t]) for t in range(T)], "ref32-vs-64", ["%.1e" % rel(r32.alpha[t], r64.alpha[t]) for t in range(T)], flush=True)

"""compute-sanitizer target: small EM-Dirichlet / Hard EM-Dirichlet batches through run_task in both M-step schedules, zero-
and few-shot, at shapes that take the few-rows kernel, the chunked kernel with row lists, the sparse E-step (estep_task_kernel)
and — K = 136, D = 136 — the tcgen05 contraction with ragged tiles.  iter_mm is cut so that a sanitizer run stays short."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import logging; logging.disable(logging.INFO)
import torch
from oracle.ref_loader import make_args
from tclip_b200 import tasks
from tclip_b200.methods.dirichlet import EM_DIRICHLET, HARD_EM_DIRICHLET

dev = torch.device("cuda:0")
for (K, T, iters) in ((20, 4, 3), (136, 3, 3), (100, 12, 3)):
    td, _ = tasks.make_zero_shot_batch(T, K, seed=7)
    for cls in (EM_DIRICHLET, HARD_EM_DIRICHLET):
        for mode in ("dense", "skip_dead"):
            a = make_args(K, iters=iters, mm_mode=mode)
            a.iter_mm = 120
            m = cls(model=None, device=dev, log_file=None, args=a)
            logs = m.run_task({k: v.clone() for k, v in td.items()})
            torch.cuda.synchronize()
            assert torch.isfinite(m.alpha).all()
    print("zero-shot ok", K, T, flush=True)
from tclip_b200.methods.dirichlet import FEW_SHOT_EM_DIRICHLET, FEW_SHOT_HARD_EM_DIRICHLET
for (K, T, shots) in ((20, 3, 2), (100, 4, 4)):
    td, _ = tasks.make_few_shot_batch(T, K, shots=shots, seed=1)
    for cls in (FEW_SHOT_EM_DIRICHLET, FEW_SHOT_HARD_EM_DIRICHLET):
        a = make_args(K, iters=3, k_eff=5)
        a.iter_mm = 120
        m = cls(model=None, device=dev, log_file=None, args=a)
        logs = m.run_task({k: v.clone() for k, v in td.items()}, shot=shots)
        torch.cuda.synchronize()
        assert torch.isfinite(m.alpha).all()
    print("few-shot ok", K, T, flush=True)

"""Scratch: where one skip_dead run_task step spends its time (K=D=1000, T=75)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import logging; logging.disable(logging.INFO)
import torch
from tclip_b200 import tasks, ops
from tclip_b200.methods.dirichlet import EM_DIRICHLET, HARD_EM_DIRICHLET
from oracle.ref_loader import make_args
dev = torch.device("cuda:0")
K, T = int(os.environ.get("PT_K", 1000)), int(os.environ.get("PT_T", 75))
ONLY_SKIP = "--skip-only" in sys.argv
for hard, iters in ((False, 20), (True, 10)):
    for mode in ("skip_dead", "dense"):
        if mode == "dense" and (hard or ONLY_SKIP): continue
        for rep in range(2):
            td, _ = tasks.make_zero_shot_batch(T, K, seed=2020, batch_index=rep)
            m = (HARD_EM_DIRICHLET if hard else EM_DIRICHLET)(model=None, device=dev, log_file=None, args=make_args(K, iters=iters, mm_mode=mode))
            torch.cuda.synchronize(); t0 = time.time()
            xq = td["x_q"].to(dev); yq = td["y_q"].long().squeeze(2).to(dev)
            torch.cuda.synchronize(); t1 = time.time()
            n_task, crit = m._run_em(xq)
            torch.cuda.synchronize(); t2 = time.time()
            m._log_iterations(n_task, crit)
            t3 = time.time()
            m.compute_acc_clustering(xq, yq)
            torch.cuda.synchronize(); t4 = time.time()
            start, ev = m._em_events; mmev = m._mm_events
            it_ms = [ (start if i == 0 else ev[i-1]).elapsed_time(ev[i]) for i in range(iters)]
            mm_ms = [mmev[2*i].elapsed_time(mmev[2*i+1]) for i in range(iters)]
            if rep == 1:
                print(f"hard={hard} mode={mode}: h2d {1e3*(t1-t0):.1f} ms | em {1e3*(t2-t1):.1f} ms (device {start.elapsed_time(ev[-1]):.1f}) | acc+matching {1e3*(t4-t3):.1f} ms")
                print("   per-iter total ms", [round(x, 1) for x in it_ms])
                print("   per-iter MM ms   ", [round(x, 1) for x in mm_ms])
                print("   mm_iters", m.mm_iters.cpu().tolist(), "n_live", m.n_live.cpu().tolist())

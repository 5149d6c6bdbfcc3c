"""Scratch: is the few-rows M-step launch-bound or latency-bound?  Same EM with different check_every (fewer, longer chunks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import torch
from tclip_b200 import tasks, ops
dev = torch.device("cuda:0")
K, T = 1000, 75
td, _ = tasks.make_zero_shot_batch(T, K, seed=2020)
xq = td["x_q"].to(dev)
for ce in (50, 250, 1000):
    for rep in range(2):
        res = ops.dirichlet_em(xq, K, 8, 1000, float(int(K / 5) * 75), False, check_every=ce, mm_mode=ops.TCLIP_MM_SKIP_DEAD, record_events=True)
        torch.cuda.synchronize()
    mm = res["mm_events"]
    print("check_every", ce, "MM ms per outer iteration", [round(mm[2*i].elapsed_time(mm[2*i+1]), 2) for i in range(8)], "n_live", res["n_live"].cpu().tolist(), flush=True)

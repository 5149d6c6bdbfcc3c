"""TEST INFRASTRUCTURE ONLY — ImageNet-shape (K = D = 1000) answers of the restated CPU oracle, float32 and float64,
frozen into tests/golden/oracle_k1000_*.npz so the GPU box can check the headline configuration without spending
minutes of CPU time per task.  The inputs are stored next to the answers (oracle_k1000_inputs.npz): the generator's
matmul/softmax is not bit-reproducible across host CPUs.  Run: python oracle/make_k1000_fixture.py (~45 min on 8 cores)"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))

from oracle import restated  # noqa: E402
from tclip_b200 import tasks  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
K, T, SEED = 1000, 3, 2020

if __name__ == "__main__":
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    td, _ = tasks.make_zero_shot_batch(T, K, seed=SEED, batch_index=777)
    np.savez_compressed(os.path.join(OUT, "oracle_k1000_inputs.npz"), x_q=td["x_q"].numpy(), y_q=td["y_q"].numpy())
    for name, hard, iters in (("em", False, 20), ("hard", True, 10)):
        t0 = time.time()
        r32 = restated.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=hard)
        r64 = restated.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=hard, dtype=torch.float64)
        a32, a64 = r32.alpha, r64.alpha
        live = (r64.u.sum(1) > 1e-15)                                   # [T,K] clusters alive at the end
        np.savez_compressed(
            os.path.join(OUT, f"oracle_k1000_{name}.npz"),
            K=K, T=T, seed=SEED, batch_index=777, iters=iters, hard=hard,
            x_q_checksum=float(td["x_q"].double().sum()), y_q=td["y_q"].numpy(),
            preds32=r32.preds.numpy(), preds64=r64.preds.numpy(), acc32=r32.acc, acc64=r64.acc,
            mm_iters32=np.asarray(r32.mm_iters), mm_iters64=np.asarray(r64.mm_iters),
            n_live32=np.asarray(r32.n_live), n_live64=np.asarray(r64.n_live),
            criterions32=r32.criterions, criterions64=r64.criterions,
            live=live.numpy(),
            # alpha: per-row norms of every row, and the full rows of the clusters alive at the end (float64 and float32)
            row_norm64=a64.norm(dim=2).numpy(), row_norm32=a32.norm(dim=2).numpy(),
            live_rows64=a64[live].numpy(), live_rows32=a32[live].numpy(),
            task_err32=np.asarray([((a32[t].double() - a64[t]).norm() / a64[t].norm()).item() for t in range(T)]),
            v32=r32.v.numpy())
        print(name, "done in %.0f s" % (time.time() - t0), "acc32", r32.acc.ravel(), "acc64", r64.acc.ravel(),
              "mm", r32.mm_iters, "n_live", r32.n_live, flush=True)

"""TEST INFRASTRUCTURE ONLY — ImageNet-shape (K = D = 1000) EM-Dirichlet answers of the restated CPU oracle for a batch of
T = 16 tasks, float32 and float64, frozen into tests/golden/oracle_k1000_t16.npz.

Why a second K = 1000 fixture: the MM exit test is global over the run_task batch (zero_shot/em_dirichlet.py:169-175), so the
iteration count of outer iteration 0 and which kernels the skip-dead schedule picks depend on the batch size; the first
fixture (make_k1000_fixture.py) has T = 3.  The inputs are regenerated on the test box (tclip_b200.tasks generates them in
float64, bit-reproducible across hosts) and checked against a weighted checksum.
Run: python oracle/make_k1000_t16_fixture.py   (~20 min on 8 cores)"""
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
K, T, ITERS, SEED, BATCH_INDEX = 1000, 16, 3, 2020, 778

if __name__ == "__main__":
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    td, _ = tasks.make_zero_shot_batch(T, K, seed=SEED, batch_index=BATCH_INDEX)
    t0 = time.time()
    r32 = restated.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=ITERS)
    print("float32 done in %.0f s" % (time.time() - t0), "mm", r32.mm_iters, "n_live", r32.n_live, flush=True)
    r64 = restated.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=ITERS, dtype=torch.float64)
    a32, a64 = r32.alpha, r64.alpha
    live = (r64.u.sum(1) > 1e-15)                                   # [T,K] clusters alive at the end
    path = os.path.join(OUT, "oracle_k1000_t16.npz")
    np.savez_compressed(
        path, K=K, T=T, seed=SEED, batch_index=BATCH_INDEX, iters=ITERS, hard=False,
        checksum_x_q=restated.weighted_checksum(td["x_q"]), y_q=td["y_q"].numpy(),
        preds32=r32.preds.numpy(), preds64=r64.preds.numpy(), acc32=r32.acc, acc64=r64.acc,
        mm_iters32=np.asarray(r32.mm_iters), mm_iters64=np.asarray(r64.mm_iters),
        n_live32=np.asarray(r32.n_live), n_live64=np.asarray(r64.n_live),
        criterions32=r32.criterions, criterions64=r64.criterions, live=live.numpy(),
        row_norm64=a64.norm(dim=2).numpy(), row_norm32=a32.norm(dim=2).numpy(),
        live_rows64=a64[live].numpy(), live_rows32=a32[live].numpy(),
        task_err32=np.asarray([((a32[t].double() - a64[t]).norm() / a64[t].norm()).item() for t in range(T)]),
        v32=r32.v.numpy())
    print("done in %.0f s" % (time.time() - t0), "acc32", r32.acc.ravel(), "mm", r32.mm_iters, r64.mm_iters, "n_live",
          r32.n_live, "saved %.2f MB" % (os.path.getsize(path) / 1e6), flush=True)

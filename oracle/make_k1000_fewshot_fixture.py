"""TEST INFRASTRUCTURE ONLY — few-shot EM-Dirichlet at ImageNet shape (K = D = 1000, 4 shots => S = 4000 support samples,
BASELINE config 3): answers of the restated CPU oracle (float32 and float64) frozen into
tests/golden/oracle_k1000_fewshot.npz.  The 32 MB support set is not stored: tclip_b200.tasks generates it in float64
(bit-reproducible across hosts) and the fixture keeps a weighted checksum of every input tensor.  The live reference cannot run this shape (16 GB per
task for its [S,K,D] one-hot product).  Run: python oracle/make_k1000_fewshot_fixture.py (~10 min)."""
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
K, T, SHOTS, ITERS, SEED = 1000, 2, 4, 4, 2020

if __name__ == "__main__":
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    td, _ = tasks.make_few_shot_batch(T, K, shots=SHOTS, k_eff=5, seed=SEED, batch_index=555)
    save = dict(K=K, T=T, shots=SHOTS, iters=ITERS, k_eff=5, seed=SEED, batch_index=555,
                y_q=td["y_q"].numpy(), y_s=td["y_s"].numpy(),
                checksum_x_q=restated.weighted_checksum(td["x_q"]), checksum_x_s=restated.weighted_checksum(td["x_s"]))
    for hard in (False, True):
        t0 = time.time()
        r32 = restated.dirichlet_few_shot(td["x_s"], td["y_s"], td["x_q"], td["y_q"], K, 5, iters=ITERS, hard=hard)
        r64 = restated.dirichlet_few_shot(td["x_s"], td["y_s"], td["x_q"], td["y_q"], K, 5, iters=ITERS, hard=hard,
                                          dtype=torch.float64)
        tag = "hard" if hard else "em"
        # every cluster is live in the few-shot setting: keep per-row norms and a sample of full rows
        rows = torch.arange(0, K, 37)
        save.update({f"preds32_{tag}": r32.preds.numpy(), f"acc32_{tag}": r32.acc, f"mm_iters32_{tag}": np.asarray(r32.mm_iters),
                     f"mm_iters64_{tag}": np.asarray(r64.mm_iters), f"criterions32_{tag}": r32.criterions,
                     f"row_norm64_{tag}": r64.alpha.norm(dim=2).numpy(), f"row_norm32_{tag}": r32.alpha.norm(dim=2).numpy(),
                     f"rows64_{tag}": r64.alpha[:, rows].numpy(), f"rows32_{tag}": r32.alpha[:, rows].numpy(),
                     f"task_err32_{tag}": np.asarray([((r32.alpha[t].double() - r64.alpha[t]).norm() / r64.alpha[t].norm()).item()
                                                      for t in range(T)])})
        print(tag, "done in %.0f s" % (time.time() - t0), "acc", r32.acc.ravel(), "mm", r32.mm_iters, r64.mm_iters, flush=True)
    np.savez_compressed(os.path.join(OUT, "oracle_k1000_fewshot.npz"), **save)
    print("saved", os.path.getsize(os.path.join(OUT, "oracle_k1000_fewshot.npz")) / 1e6, "MB")

"""Measurement: how the live rows of the skip-dead schedule behave inside an M-step (K=D=1000, T=75 by default).

For every outer iteration: live rows, how many reach a bit-exact fixed point (and when), how many enter a longer cycle
(Brent's search, statistics build of mm_spec_kernel), and the M-step time of the product build (rows stop at their fixed
point) next to the statistics build (every row runs all iter_mm iterations).  Also checks that both builds give the same
alpha / labels bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import logging; logging.disable(logging.INFO)
import numpy as np
import torch
from tclip_b200 import tasks, ops

dev = torch.device("cuda:0")
K, T = int(os.environ.get("PT_K", 1000)), int(os.environ.get("PT_T", 75))
NOISE = float(os.environ.get("PT_NOISE", tasks.NOISE_SCALE))
KEFF = tuple(int(v) for v in os.environ.get("PT_KEFF", "3,10").split(","))


def run(xq, iters, hard, probe):
    lambd = int(K / 5) * 75
    torch.cuda.synchronize()
    r = ops.dirichlet_em(xq, K, iters, 1000, float(lambd), hard, mm_mode=ops.TCLIP_MM_SKIP_DEAD, record_events=True,
                         spec_probe=probe)
    torch.cuda.synchronize()
    ev = r["mm_events"]
    r["mm_ms"] = [ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(iters)]
    return r


for hard, iters in ((False, 20), (True, 10)):
    td, _ = tasks.make_zero_shot_batch(T, K, seed=2020, batch_index=1, noise=NOISE, k_eff_range=KEFF)
    xq = td["x_q"].to(dev)
    run(xq, iters, hard, False)                     # warm-up
    prod = run(xq, iters, hard, False)
    stat = run(xq, iters, hard, True)
    same_alpha = torch.equal(prod["alpha"], stat["alpha"])
    same_lab = torch.equal(prod["labels"], stat["labels"])
    print(f"== hard={hard} K={K} T={T} noise={NOISE} classes per task {KEFF}: product == statistics build: alpha {same_alpha} labels {same_lab} "
          f"mm_iters equal {torch.equal(prod['mm_iters'], stat['mm_iters'])}")
    n_live = prod["n_live"].cpu().tolist()
    pr = stat["spec_probe"].cpu().numpy()
    mm_rows_p, mm_rows_s = prod["mm_rows"].cpu().tolist(), stat["mm_rows"].cpu().tolist()
    labels = prod["labels"].cpu().numpy()
    sizes = np.stack([np.bincount(labels[t], minlength=K) for t in range(T)])   # final cluster sizes
    for it in range(iters):
        p = pr[it]
        used = p[:, 0] >= 0
        n = int(used.sum())
        if n == 0:
            print(f" it {it:2d}: n_live {n_live[it]:6d}  (chunk kernels)  MM ms product {prod['mm_ms'][it]:.3f} stats {stat['mm_ms'][it]:.3f}"
                  f"  row-iters {mm_rows_p[it]:.3e}")
            continue
        q = p[used]
        fixed = q[:, 1] >= 0
        cyc = (q[:, 2] >= 0) & ~fixed
        never = ~fixed & ~cyc
        fx = np.sort(q[fixed, 1])
        msg = (f" it {it:2d}: n_live {n_live[it]:6d} spec rows {n:5d} | fixed point {int(fixed.sum()):5d} "
               f"(iter min/median/p90/max {fx[0] if fx.size else -1}/{fx[fx.size // 2] if fx.size else -1}/"
               f"{fx[int(fx.size * 0.9)] if fx.size else -1}/{fx[-1] if fx.size else -1}) | cycle only {int(cyc.sum()):4d} "
               f"(periods {np.unique(q[cyc, 3]).tolist()[:8]}, seen at median {int(np.median(q[cyc, 2])) if cyc.any() else -1}) "
               f"| neither {int(never.sum()):4d} | MM ms product {prod['mm_ms'][it]:.3f} stats {stat['mm_ms'][it]:.3f} "
               f"| row-iters product {mm_rows_p[it]:.3e} stats {mm_rows_s[it]:.3e}")
        print(msg)
    # the last iteration: who are the rows that never stop?  (live rows in ascending (task, class) order = probe order)
    it = iters - 1
    live_sizes = sizes[sizes > 0]
    q = pr[it][pr[it][:, 0] >= 0]
    if live_sizes.size == q.shape[0]:
        cyc = q[:, 2] >= 0
        for lo, hi in ((1, 1), (2, 2), (3, 4), (5, 8), (9, 1000)):
            sel = (live_sizes >= lo) & (live_sizes <= hi)
            if sel.any():
                per = q[sel & cyc, 3]
                print(f"   cluster size {lo}..{hi}: {int(sel.sum())} rows, periodic {int((sel & cyc).sum())}, periods "
                      f"{dict(zip(*[a.tolist() for a in np.unique(per, return_counts=True)]))}, detected at (median) "
                      f"{int(np.median(q[sel & cyc, 2])) if (sel & cyc).any() else -1}")
        never = q[:, 1] < 0
        print(f"   last iteration: cluster sizes of rows that never reach a fixed point: {np.bincount(live_sizes[never])[:8].tolist()} "
              f"(index = size), of rows that do: {np.bincount(live_sizes[~never])[:12].tolist()}")
    else:
        print(f"   (live rows {q.shape[0]} != non-empty final clusters {live_sizes.size}: size attribution skipped)")
    print("   total EM MM ms: product %.2f, statistics build %.2f" % (sum(prod["mm_ms"]), sum(stat["mm_ms"])))

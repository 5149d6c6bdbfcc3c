"""Scratch diagnostics run on the GPU box (not a test): parity numbers + timings printed as text."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import numpy as np, torch
from tclip_b200 import ops, tasks
from tclip_b200.methods.dirichlet import EM_DIRICHLET, HARD_EM_DIRICHLET
from oracle import restated as R
from oracle.ref_loader import make_args

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), flush=True)

def rel(a, b):
    return ((a - b).norm() / b.norm()).item()

def parity(K, T, iters, hard, seed, mode):
    td, _ = tasks.make_zero_shot_batch(T, K, seed=seed)
    args = make_args(K, iters=iters, mm_mode=mode)
    cls = HARD_EM_DIRICHLET if hard else EM_DIRICHLET
    m = cls(model=None, device=dev, log_file=None, args=args)
    torch.cuda.synchronize(); t0 = time.time()
    logs = m.run_task({k: v.clone() for k, v in td.items()})
    torch.cuda.synchronize(); t1 = time.time()
    r32 = R.dirichlet_zero_shot(td['x_q'], td['y_q'], K, iters=iters, hard=hard)
    r64 = R.dirichlet_zero_shot(td['x_q'], td['y_q'], K, iters=iters, hard=hard, dtype=torch.float64)
    a = m.alpha.cpu()
    print(f"--- K={K} T={T} iters={iters} hard={hard} seed={seed} mode={mode}: gpu {t1-t0:.3f}s oracle32 {r32.seconds:.1f}s")
    print("   mm_iters gpu", m.mm_iters.cpu().tolist(), "\n   mm_iters o32", r32.mm_iters, "\n   mm_iters o64", r64.mm_iters)
    print("   n_live gpu", m.n_live.cpu().tolist(), "o32", r32.n_live)
    print("   labels agree vs o32 %.5f vs o64 %.5f ; o32 vs o64 %.5f" % ((m.labels.cpu().long() == r32.preds).float().mean(), (m.labels.cpu().long() == r64.preds).float().mean(), (r32.preds == r64.preds).float().mean()))
    print("   acc gpu %.4f o32 %.4f o64 %.4f" % (logs['acc'].mean(), r32.acc.mean(), r64.acc.mean()))
    pt = lambda x, y: [round(rel(x[t], y[t]), 7) for t in range(min(T, 4))]
    print("   alpha frob rel per task: gpu-vs-o64", pt(a.double(), r64.alpha), " o32-vs-o64", pt(r32.alpha.double(), r64.alpha), " gpu-vs-o32", pt(a, r32.alpha))
    e = ((a.double() - r64.alpha).abs() / r64.alpha.abs()); e32 = ((r32.alpha.double() - r64.alpha).abs() / r64.alpha.abs())
    print("   alpha elementwise rel: gpu-vs-o64 max %.3e p99 %.3e frac>1e-4 %.4f | o32-vs-o64 max %.3e p99 %.3e frac>1e-4 %.4f" % (e.max(), e.flatten().quantile(0.99) if e.numel() < 1.6e7 else -1, (e > 1e-4).float().mean(), e32.max(), e32.flatten().quantile(0.99) if e.numel() < 1.6e7 else -1, (e32 > 1e-4).float().mean()))
    print("   crit gpu", np.round(logs['criterions'][:6], 6), "o32", np.round(r32.criterions[:6], 6), flush=True)

def timing(K, T, iters, hard, mode, reps=2):
    td, _ = tasks.make_zero_shot_batch(T, K, seed=2020)
    args = make_args(K, iters=iters, mm_mode=mode)
    cls = HARD_EM_DIRICHLET if hard else EM_DIRICHLET
    for r in range(reps):
        m = cls(model=None, device=dev, log_file=None, args=args)
        torch.cuda.synchronize(); t0 = time.time()
        logs = m.run_task({k: v.clone() for k, v in td.items()})
        torch.cuda.synchronize(); t1 = time.time()
        s, ev = m._em_events
        em_ms = s.elapsed_time(ev[-1])
        rows = m.mm_rows.cpu().numpy()
        upd = float(rows.sum()) * K
        print(f"TIMING K={K} T={T} iters={iters} hard={hard} mode={mode} rep{r}: wall {t1-t0:.3f}s em {em_ms/1e3:.3f}s tasks/s {T/(t1-t0):.2f} acc {logs['acc'].mean():.4f} mm_iters {m.mm_iters.cpu().tolist()} n_live {m.n_live.cpu().tolist()} elem-updates {upd:.3e} -> {upd/(em_ms/1e3):.3e}/s", flush=True)

def parity_fs(K, T, shots, iters, hard, seed):
    from tclip_b200.methods.dirichlet import FEW_SHOT_EM_DIRICHLET, FEW_SHOT_HARD_EM_DIRICHLET
    td, _ = tasks.make_few_shot_batch(T, K, shots=shots, seed=seed)
    args = make_args(K, iters=iters, k_eff=5)
    cls = FEW_SHOT_HARD_EM_DIRICHLET if hard else FEW_SHOT_EM_DIRICHLET
    m = cls(model=None, device=dev, log_file=None, args=args)
    logs = m.run_task({k: v.clone() for k, v in td.items()}, shot=shots)
    r32 = R.dirichlet_few_shot(td['x_s'], td['y_s'], td['x_q'], td['y_q'], K, 5, iters=iters, hard=hard)
    r64 = R.dirichlet_few_shot(td['x_s'], td['y_s'], td['x_q'], td['y_q'], K, 5, iters=iters, hard=hard, dtype=torch.float64)
    a = m.alpha.cpu()
    print(f"--- FEW-SHOT K={K} T={T} shots={shots} iters={iters} hard={hard}")
    print("   mm_iters gpu", m.mm_iters.cpu().tolist(), "o32", r32.mm_iters, "o64", r64.mm_iters)
    print("   labels agree vs o32 %.5f vs o64 %.5f" % ((m.labels.cpu().long() == r32.preds).float().mean(), (m.labels.cpu().long() == r64.preds).float().mean()))
    print("   acc gpu %.4f o32 %.4f" % (logs['acc'].mean(), r32.acc.mean()))
    print("   alpha frob rel gpu-vs-o64 %.3e o32-vs-o64 %.3e gpu-vs-o32 %.3e" % (rel(a.double(), r64.alpha), rel(r32.alpha.double(), r64.alpha), rel(a, r32.alpha)))
    print("   crit gpu", np.round(logs['criterions'][:6], 6), "o32", np.round(r32.criterions[:6], 6), flush=True)

which = sys.argv[1:] or ["parity", "timing"]
if "sanity" in which:
    parity(20, 3, 3, False, 1, "dense")
    parity(20, 3, 3, True, 1, "skip_dead")
    parity(37, 2, 3, False, 2, "skip_dead")
    parity_fs(20, 2, 2, 3, False, 3)
if "fewshot" in which:
    parity_fs(20, 2, 2, 3, False, 3)
    parity_fs(20, 2, 2, 3, True, 3)
    parity_fs(100, 4, 4, 6, False, 4)
if "parity" in which:
    parity(20, 3, 4, False, 1, "dense")
    parity(20, 3, 4, False, 1, "skip_dead")
    parity(100, 8, 20, False, 0, "dense")
    parity(100, 8, 20, False, 0, "skip_dead")
    parity(100, 8, 10, True, 1, "skip_dead")
if "timing" in which:
    timing(100, 100, 20, False, "dense")
    timing(100, 100, 20, False, "skip_dead")
    timing(1000, 75, 20, False, "skip_dead")
    timing(1000, 75, 10, True, "skip_dead")
    timing(1000, 75, 20, False, "dense", reps=1)

"""Measurement: BASELINE config 4 (soft k-means / EM-Gaussian / hard k-means on RN50-shape visual features, D = 1024,
K = 1000, 100 tasks per run_task batch): where one batch spends its time, strictly serial run_task calls."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import logging; logging.disable(logging.INFO)
import torch
from tclip_b200 import tasks, ops
from tclip_b200.config import make_args
from tclip_b200.methods import kmeans as M
from oracle import ref_loader

ref_loader._install_clip_stub()
dev = torch.device("cuda:0")
K, T, D = int(os.environ.get("PT_K", 1000)), int(os.environ.get("PT_T", 100)), int(os.environ.get("PT_D", 1024))
td, txt = tasks.make_zero_shot_batch(T, K, seed=2020, softmax_feature=False, embed_dim=D)
host = {k: v.pin_memory() for k, v in td.items()}
model = ref_loader.StubTextModel(txt)
for name, cls, iters in (("soft k-means", M.SOFT_KMEANS, 20), ("EM-Gaussian", M.EM_GAUSSIAN, 20), ("hard k-means", M.HARD_KMEANS, 10)):
    args = make_args(K, iters=iters, use_softmax_feature=False)
    for rep in range(3):
        m = cls(model=model, device=dev, log_file=None, args=args)
        torch.cuda.synchronize(); t0 = time.time()
        logs = m.run_task(dict(host))
        torch.cuda.synchronize(); t1 = time.time()
    # stages of the last call, re-run one by one with events
    q = host["x_q"].to(dev); y = host["y_q"].long().squeeze(2).to(dev)
    def timed(fn):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); e1.synchronize()
        return r, e0.elapsed_time(e1)
    m = cls(model=model, device=dev, log_file=None, args=args)
    u0, ms_init = timed(lambda: m._initial_u(q))
    res, ms_loop = timed(lambda: ops.kmeans_run(q, u0, m.mode, iters, 30.0, lambd=float(getattr(m, "lambd", 0.0)), record_events=True))
    ev = res["events"]
    it_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    m.u, m.labels, m.v = res["u"], res["labels"], res["v"]
    _, ms_acc = timed(lambda: m.compute_acc_clustering(q, y))
    _, ms_w = timed(lambda: ops.kmeans_expand_centroids(res["coef"], q))
    print(f"{name}: run_task {1e3 * (t1 - t0):.2f} ms per batch of {T} = {T / (t1 - t0):.0f} tasks/s | acc {logs['acc'].mean():.4f} | "
          f"initial u {ms_init:.2f} ms, loop {ms_loop:.2f} ms (per iteration {sum(it_ms) / iters:.3f}), accuracy {ms_acc:.2f} ms, "
          f"w on demand {ms_w:.2f} ms")

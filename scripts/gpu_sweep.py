"""Measurement: how EM-Dirichlet throughput at ImageNet shape (K = D = 1000, 75 tasks per batch, iter 20) depends on the
workload — how many clusters per task stay alive.  The canonical generator (3..10 true classes per task, noise 9) collapses
to ~3 live clusters per task; tasks drawn from more classes keep more clusters alive, up to and past the row counts at which
the skip-dead schedule switches kernels (1480 live rows: mm_spec_kernel -> mm_chunk_kernel over the row list; 4096: row-wise
-> dense E-step).  Prints one line per workload: serial ms per batch, tasks/s with 4 batches in flight, live clusters per
task, executed element-updates and their fraction of the FP32 peak, accuracy."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import logging; logging.disable(logging.INFO)
import torch
from tclip_b200 import tasks, ops
from tclip_b200.config import make_args
from tclip_b200.methods.dirichlet import EM_DIRICHLET
from tclip_b200.pipeline import BatchPipeline

dev = torch.device("cuda:0")
K, T, ITERS = 1000, 75, 20
n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
for _ in range(30):
    ops.probe_issue_rate("ffma", n_sm * 8, 4000)
flop, ms = ops.probe_issue_rate("ffma", n_sm * 8, 4000)
peak = flop / (ms * 1e-3) / 1e12
print(f"FP32 peak (register-only FFMA probe) {peak:.1f} TFLOP/s")
pipe = BatchPipeline(dev, streams=4)
WORKLOADS = [((3, 10), 9.0), ((10, 20), 9.0), ((20, 30), 9.0), ((20, 30), 6.0), ((40, 50), 6.0), ((60, 75), 5.0)]
print("| classes per task | noise | arg-max acc | EM acc | live clusters/task (it 1, 2, last) | live rows (last) | serial ms/batch | "
      "tasks/s serial | tasks/s 4 in flight | executed updates/task | frac serial | frac in flight | dense-mode ms/batch |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for ke, noise in WORKLOADS:
    batches = []
    for b in range(8):
        td, _ = tasks.make_zero_shot_batch(T, K, seed=2020, batch_index=b, k_eff_range=ke, noise=noise)
        batches.append((td["x_q"].to(dev), td["y_q"].long().squeeze(2).to(dev)))
    argmax_acc = float((batches[0][0].argmax(-1) == batches[0][1]).float().mean())
    args = make_args(K, iters=ITERS, mm_mode="skip_dead")

    def step(b):
        m = EM_DIRICHLET(model=None, device=dev, log_file=None, args=args)
        m.run_method(query=b[0], y_q=b[1])
        return (m.mm_rows.sum(), m.n_live.clone(), torch.cat(m.test_acc, dim=1).mean())

    step(batches[0]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs = [step(b) for b in batches[:4]]
    e1.record(); torch.cuda.synchronize()
    ms_serial = e0.elapsed_time(e1) / 4
    pipe.map(step, batches[:4]); torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    pipe.map(step, batches + batches)
    e3.record(); torch.cuda.synchronize()
    ms_flight = e2.elapsed_time(e3) / 16
    upd = float(sum(o[0] for o in outs).item()) * K / 4
    nl = outs[0][1].cpu().tolist()
    acc = float(sum(o[2] for o in outs).item()) / 4
    dense_ms = float("nan")
    if ke == (3, 10) or ke == (40, 50):
        m = EM_DIRICHLET(model=None, device=dev, log_file=None, args=make_args(K, iters=ITERS, mm_mode="dense"))
        torch.cuda.synchronize(); t0 = time.time()
        m.run_method(query=batches[0][0], y_q=batches[0][1])
        torch.cuda.synchronize(); dense_ms = 1e3 * (time.time() - t0)
    print(f"| {ke[0]}..{ke[1]} | {noise} | {argmax_acc:.3f} | {acc:.3f} | {nl[1] / T:.1f}, {nl[2] / T:.1f}, {nl[-1] / T:.1f} | {nl[-1]} | "
          f"{ms_serial:.1f} | {T / ms_serial * 1e3:.0f} | {T / ms_flight * 1e3:.0f} | {upd / T:.3e} | "
          f"{upd * 74 / (ms_serial * 1e-3) / 1e12 / peak:.2f} | {upd * 74 / (ms_flight * 1e-3) / 1e12 / peak:.2f} | {dense_ms:.0f} |", flush=True)
pipe.close()

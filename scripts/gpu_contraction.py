"""Scratch: the E-step contraction on tcgen05 (3 x TF32) against float64 — error (bias, rms, max) and time of
  tcgen05            3 x TF32, per-D-block partial products summed in round-to-nearest fp32 outside the tensor core
  tcgen05_tmem_sum   3 x TF32, the whole sum left in the TMEM accumulator
  simt               the CUDA-core fp32 kernel
on alpha / log z taken from a real EM run (heavy elements ~1e5) and on small / ragged shapes."""
import os, sys, time
os.environ["TCLIP_CONTRACTION"] = "simt"     # the EM run that produces alpha must not depend on the kernel under test
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import logging; logging.disable(logging.INFO)
import torch
from tclip_b200 import tasks, ops
from tclip_b200.config import make_args
from tclip_b200.methods.dirichlet import EM_DIRICHLET
dev = torch.device("cuda:0")

def stats(name, got, ref):
    err = got.double() - ref
    scale = ref.abs().max().item()
    print(f"  {name:18s} bias {err.mean().item():+.3e}  rms {err.pow(2).mean().sqrt().item():.3e}  max {err.abs().max().item():.3e}"
          f"   (|l3| max {scale:.3e}; rms/|l3|max {err.pow(2).mean().sqrt().item() / scale:.2e})", flush=True)

def bench(mode, logz, alpha, reps=10):
    ops.contraction(logz, alpha, mode); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.contraction(logz, alpha, mode)
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps

def case(T, K, iters, seed, label):
    td, _ = tasks.make_zero_shot_batch(T, K, seed=seed)
    m = EM_DIRICHLET(model=None, device=dev, log_file=None, args=make_args(K, iters=iters, mm_mode="skip_dead"))
    m.run_task({k: v.clone() for k, v in td.items()})
    alpha = m.alpha.contiguous()
    logz = ops.log_features(td["x_q"].to(dev))
    ref = torch.einsum("tnd,tkd->tnk", logz.double(), (alpha - 1.0).double())
    print(f"{label}: T={T} n={logz.shape[1]} K=D={K}, alpha after {iters} outer iterations (max {alpha.max().item():.3e})", flush=True)
    for mode in ("tcgen05", "tcgen05_tmem_sum", "simt"):
        got = ops.contraction(logz, alpha, mode)
        torch.cuda.synchronize()
        stats(mode, got, ref)
    for mode in ("tcgen05", "tcgen05_tmem_sum", "simt"):
        ms = bench(mode, logz, alpha)
        flop = 2.0 * T * logz.shape[1] * K * K
        print(f"  {mode:18s} {ms * 1e3:8.1f} us   {flop / ms / 1e9:8.2f} TFLOP/s (fp32-equivalent)   alpha read at {alpha.numel() * 4 / ms / 1e6:7.1f} GB/s", flush=True)

case(2, 20, 3, 3, "tiny")
case(4, 100, 5, 1, "caltech-shape")
case(3, 1000, 2, 2, "imagenet-shape, 3 tasks")
case(75, 1000, 6, 2020, "imagenet-shape, full batch")

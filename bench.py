#!/usr/bin/env python
"""bench.py — EM-Dirichlet tasks/sec at ImageNet shape (K = D = 1000, n_query = 75) on N B200s, next to the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--method em|hard|soft|gauss|hardkm]
                    [--mm-mode ...] [--tasks N]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A *step* is one ``run_task`` batch of 75 synthetic zero-shot tasks (``batch_size`` 75) through the whole
hot path: H2D, the EM loop (``iter`` outer iterations of moments -> MM M-step -> E-step), cluster prototypes, Hungarian
label matching, accuracy.  Every rank runs its own, different batches (whole batches are the sharding unit because the
MM early exit is batch-global); ``value`` = tasks of all ranks / max-over-ranks device time.  Rank 0 prints ONE JSON line.

  value        inputs already resident in HBM (``run_method``), timed with CUDA events;
  e2e          the reference-facing call ``run_task(task_dic)`` with pinned HOST tensors: H2D + EM + D2H of the
               accuracies inside the timed region;
               both legs keep ``--streams`` (default 8) whole batches in flight per GPU — every batch is one unchanged
               ``run_method`` / ``run_task`` call on its own CUDA stream and host thread (``tclip_b200.pipeline``), because
               half of a batch is a latency-bound tail that leaves the SMs idle; ``serial`` inside ``value``'s line and
               inside ``e2e`` is the same leg strictly one batch after the other (the reference's evaluator loop);
  roofline     ``frac`` describes the TIMED STEP: element-updates the step really executed (device-side counters of
               every M-step) x 74 flop / step time / FP32 FMA peak; ``serial_frac`` is the same for the
               strictly serial leg.  ``kernel_frac`` is the dominant kernel (mm_chunk_kernel, the MM M-step) run alone on
               a full batch of rows (T*K rows x D, two launches = the first 101 MM iterations of an M-step), CUDA events
               on the launching stream.  peak = FP32 FMA issue rate measured by the library's register-only
               microbenchmark in the same run (MEASURED_PEAKS.json has no FP32 line; the theoretical figure is quoted
               beside it).  ``dense_equivalent`` is what the reference's dense schedule would have executed for the same
               result (the skip-dead schedule proves most of it redundant), ``per_outer_iteration`` the CUDA-event times
               of every outer iteration and of its M-step;
  cpu_baseline the restated oracle (a port of the reference's CPU path; the Python reference itself cannot travel to
               the GPU box) on a bounded sample of the same workload: 2 tasks in one batch, outer iteration 0 plus two
               full 1000-iteration M-step iterations, extrapolated to ``iter`` iterations (every later outer iteration
               repeats the same 1000 MM iterations).

``--method soft|gauss|hardkm`` benches BASELINE config 4 instead (soft k-means / EM-Gaussian / hard k-means on RN50-shape
visual features, D = 1024, K = 1000, 100 tasks per run_task batch) with an HBM roofline (the reference's
w-space loop moves 8 MB per task and iteration); ``--tasks N`` benches BASELINE config 5 (N EM-Dirichlet tasks sharded over
the ranks by whole batches, task construction — sampler + device-side gather — inside the timed region, strong scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "transductive-clip_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

K_CLASSES = 1000
N_QUERY = 75
TASKS_PER_BATCH = 75
FLOP_PER_UPDATE = 74.0      # canonical FP32 flop (FMA = 2) of one MM element-update
MUFU_PER_UPDATE = 5.0       # this kernel: rcp X, rcp P, lg2 P, sqrt, rcp (tclip_math.cuh); ln X is a polynomial
SEED = 2020                 # the reference's default seed (config/datasets_config/*.yaml:10)
CPU_SAMPLE_TASKS = 2         # tasks in the bounded CPU sample (one run_task batch)
CPU_SAMPLE_STEADY = 2        # full 1000-iteration outer iterations timed after outer iteration 0
REF_SAMPLE_TASKS = 2         # tasks per step of the `--impl reference` arm (one batch; outer iteration 0 + one full one)
KM_CLASSES, KM_DIM, KM_TASKS_PER_BATCH = 1000, 1024, 100   # BASELINE config 4


KMEANS = {"soft": ("soft k-means", "SOFT_KMEANS", 20), "gauss": ("EM-Gaussian", "EM_GAUSSIAN", 20),
          "hardkm": ("hard k-means", "HARD_KMEANS", 10)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", default="em", choices=["em", "hard", "soft", "gauss", "hardkm"],
                    help="em: EM-Dirichlet, iter 20 (the metric); hard: Hard EM-Dirichlet, iter 10 (BASELINE configs[1]); "
                         "soft / gauss / hardkm: soft k-means, EM-Gaussian, hard k-means at BASELINE config 4")
    ap.add_argument("--tasks", type=int, default=0,
                    help="BASELINE config 5: this many tasks in total, sharded over the ranks by whole batches, task "
                         "construction inside the timed region (0 = the default per-step bench)")
    ap.add_argument("--mm-mode", default="skip_dead", choices=["skip_dead", "dense"])
    ap.add_argument("--classes", type=int, default=None)
    ap.add_argument("--tasks-per-batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--classes-per-task", default="3,10",
                    help="lo,hi: true classes per synthetic task, uniform (3,10 = the reference's sampler, src/sampler_zero_shot.py:54, "
                         "and the metric's workload; e.g. 20,30 keeps ~25 clusters per task alive: profiles/r2_workload_sweep.md)")
    ap.add_argument("--noise", type=float, default=None, help="noise scale of the synthetic embeddings (default 9, SURVEY.md §8(d))")
    ap.add_argument("--streams", type=int, default=8,
                    help="run_task batches in flight per GPU (own CUDA stream + host thread each); 1 = strictly serial")
    a = ap.parse_args()
    a.k_eff_range = tuple(int(v) for v in a.classes_per_task.split(","))
    km = a.method in KMEANS
    if a.classes is None:
        a.classes = KM_CLASSES if km else K_CLASSES
    if a.tasks_per_batch is None:
        a.tasks_per_batch = KM_TASKS_PER_BATCH if km else TASKS_PER_BATCH
    return a


def workload_name(a):
    if a.method in KMEANS:
        nm, _, it = KMEANS[a.method]
        return (f"{nm} zero-shot, synthetic RN50-shape visual features (D={KM_DIM}, K={a.classes}, n_query={N_QUERY}, "
                f"batch_size {a.tasks_per_batch} tasks per run_task, iter {it})"), it
    it = 20 if a.method == "em" else 10
    nm = "EM-Dirichlet" if a.method == "em" else "Hard EM-Dirichlet"
    variant = "" if (a.k_eff_range == (3, 10) and a.noise is None) else \
        f"; NOT the metric's generator: {a.k_eff_range[0]}..{a.k_eff_range[1]} classes per task, noise {a.noise if a.noise is not None else 9.0}"
    return (f"{nm} zero-shot, synthetic ImageNet-shape softmax features (K=D={a.classes}, n_query={N_QUERY}, "
            f"batch_size {a.tasks_per_batch} tasks per run_task, iter {it}, iter_mm 1000{variant})"), it


def gen_kwargs(a):
    kw = {"k_eff_range": a.k_eff_range}
    if a.noise is not None:
        kw["noise"] = a.noise
    return kw


def metric_name(a):
    if a.method in KMEANS:
        return f"{KMEANS[a.method][0]} tasks/sec (D={KM_DIM}, K={a.classes}, N={N_QUERY})"
    return "EM-Dirichlet tasks/sec (K=D=1000, N=75)"


# ----------------------------------------------------------------------------------------------------------------------
# clocks sampler (pynvml; same fields as the profiling recipe's nvidia-smi line)
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index: int, period: float = 0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_sample(a, iters_full: int, n_tasks: int = CPU_SAMPLE_TASKS, n_steady: int = CPU_SAMPLE_STEADY):
    """One bounded sample of the CPU path on the same workload.

    Dirichlet: the restated oracle on ONE run_task batch of ``n_tasks`` tasks, outer iteration 0 (its MM loop exits early)
    plus ``n_steady`` later outer iterations (each runs all 1000 MM iterations), timed per outer iteration:
    seconds/task = (t_iter0 + (iter - 1) * median(t_iter1..) + t_accuracy) / n_tasks.
    k-means family: ``n_tasks`` tasks through all ``iter`` iterations (no inner loop to extrapolate over)."""
    import statistics
    import torch
    from oracle import restated
    from tclip_b200 import tasks
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    if a.method in KMEANS:
        td, txt = tasks.make_zero_shot_batch(n_tasks, a.classes, n_query=N_QUERY, seed=SEED, batch_index=10_000,
                                             softmax_feature=False, embed_dim=KM_DIM)
        t0 = time.time()
        restated.kmeans_family(td["x_q"], td["y_q"], a.classes, method={"soft": "soft", "gauss": "gauss", "hardkm": "hard"}[a.method],
                               iters=iters_full, use_softmax_feature=False, text=txt, contraction="einsum")
        wall = time.time() - t0
        return {"value": n_tasks / wall, "unit": "tasks/s", "cores": cores, "kind": "port",
                "sample": (f"oracle/restated.py kmeans_family (torch CPU fp32, {cores} threads, feature-space loop as the reference) on "
                           f"one batch of {n_tasks} tasks D={KM_DIM} K={a.classes}, all {iters_full} iterations + accuracy: {wall:.1f} s"),
                "seconds_measured": wall}
    td, _ = tasks.make_zero_shot_batch(n_tasks, a.classes, n_query=N_QUERY, seed=SEED, batch_index=10_000, **gen_kwargs(a))
    t0 = time.time()
    n_it = min(1 + n_steady, iters_full)
    r = restated.dirichlet_zero_shot(td["x_q"], td["y_q"], a.classes, iters=n_it, hard=(a.method == "hard"))
    wall = time.time() - t0
    t_it0 = r.iter_seconds[0]
    t_it1 = statistics.median(r.iter_seconds[1:]) if n_it > 1 else t_it0
    t_acc = max(wall - sum(r.iter_seconds), 0.0)
    per_batch = t_it0 + (iters_full - 1) * t_it1 + t_acc
    updates = float(sum(r.mm_iters)) * n_tasks * a.classes * a.classes
    return {
        "value": n_tasks / per_batch, "unit": "tasks/s", "cores": cores, "kind": "port",
        "sample": (f"oracle/restated.py (torch CPU fp32, {cores} threads) on one run_task batch of T={n_tasks} tasks K=D={a.classes}: "
                   f"{n_it} of {iters_full} outer iterations measured ({'+'.join(str(i) for i in r.mm_iters)} MM iterations, "
                   f"{wall:.1f} s), extrapolated as t_iter0 + {iters_full - 1} x median(t_iter1..) + t_accuracy = "
                   f"{per_batch / n_tasks:.1f} s/task"),
        "element_updates_per_s": updates / sum(r.iter_seconds), "seconds_measured": wall,
        "tasks_in_sample": n_tasks,
    }


def run_reference(a):
    """``--impl reference``: every step is one bounded CPU sample (one batch of 2 tasks; outer iteration 0 + one full
    1000-iteration outer iteration for the Dirichlet methods, all iterations for the k-means family), all ``--steps`` of them
    are run, the line's value is their median."""
    import statistics
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, iters_full = workload_name(a)
    vals, last = [], None
    for i in range(min(a.warmup, 1) + a.steps):          # one real warm-up (thread pool, allocator) is enough on the CPU
        last = cpu_sample(a, iters_full, n_tasks=REF_SAMPLE_TASKS, n_steady=1)
        if i >= min(a.warmup, 1):
            vals.append(last["value"])
    v = statistics.median(vals)
    last["value"] = v
    last["per_step_values"] = [round(x, 6) for x in vals]
    print(json.dumps({
        "impl": "reference", "metric": metric_name(a), "value": v, "unit": "tasks/s",
        "n_gpus": a.gpus, "steps": len(vals), "warmup": a.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "note": "CPU arm: one step = one bounded sample (one batch of %d tasks; Dirichlet: outer iteration 0 + one "
                   "full outer iteration, extrapolated per task); value = median over the steps; host cores only, no GPU" % REF_SAMPLE_TASKS},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------------
class Harness:
    """What every B200 leg shares: process group, barrier + max-over-ranks timing, batches in flight, clock sampling."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        from tclip_b200 import ops
        self.a, self.torch, self.dist, self.ops = a, torch, dist, ops
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a B200: no CUDA device (there is no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        ops.device_check(self.local_rank)
        from tclip_b200.pipeline import BatchPipeline
        self.pipe = BatchPipeline(self.dev, streams=a.streams) if a.streams > 1 else None
        self.n_sm = torch.cuda.get_device_properties(self.dev).multi_processor_count

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def run_many(self, fn, items):
        """Whole batches, up to --streams of them in flight (each on its own CUDA stream from its own host thread); every
        call is complete (stream synchronised) when it returns."""
        return self.pipe.map(fn, list(items)) if self.pipe else [fn(i) for i in items]

    def timed(self, fn, items, serial=False):
        """barrier + synchronize, CUDA events on the (otherwise idle) default stream around the whole leg, max over ranks."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = [fn(i) for i in items] if serial else self.run_many(fn, items)
        e1.record()
        self.barrier()
        ms_rank = e0.elapsed_time(e1)
        return out, self.max_over_ranks(ms_rank), ms_rank

    def warm_clocks(self):
        # a fresh box idles at low clocks and W steps of a few ms do not always bring it up: ~0.5 s of register-only FMA work
        # (untimed) before the timed region, so both legs run at the clocks the sampler reports
        for _ in range(50):
            self.ops.probe_issue_rate("ffma", self.n_sm * 8, 4000)

    def peaks(self):
        ops = self.ops
        ops.probe_issue_rate("ffma", self.n_sm * 8, 200)
        flop, ms_f = ops.probe_issue_rate("ffma", self.n_sm * 8, 4000)
        ops.probe_issue_rate("mufu", self.n_sm * 8, 50)
        mops, ms_m = ops.probe_issue_rate("mufu", self.n_sm * 8, 1000)
        return flop / (ms_f * 1e-3) / 1e12, mops / (ms_m * 1e-3) / 1e12

    def gather_mean(self, value: float) -> float:
        """The one NCCL use of the path: gather the accuracies."""
        if self.world == 1:
            return value
        torch, dist = self.torch, self.dist
        t = torch.tensor([value], device=self.dev)
        gathered = [torch.zeros_like(t) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(t, gathered, dst=0)
        return float(torch.cat(gathered).mean().item()) if self.rank == 0 else value

    def gather_list(self, vals):
        if self.world == 1:
            return [vals]
        torch, dist = self.torch, self.dist
        tt = torch.tensor(vals, device=self.dev, dtype=torch.float64)
        gl = [torch.zeros_like(tt) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(tt, gl, dst=0)
        return [g.tolist() for g in gl] if self.rank == 0 else [vals]

    def h2d_ms(self, host_tensor):
        torch = self.torch
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            _x = host_tensor.to(self.dev, non_blocking=True)
        c1.record()
        c1.synchronize()
        return c0.elapsed_time(c1) / 3

    def close(self):
        if self.pipe:
            self.pipe.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def concurrency_note(a):
    return ("%d whole batches in flight per GPU, each one unchanged run_task / run_method call on its own CUDA stream and host "
            "thread (tclip_b200.pipeline); `serial` = one at a time, as the reference's evaluator loop" % a.streams)


def main():
    a = parse()
    if os.environ.get("TCLIP_SWITCH_INTERVAL"):   # measurement knob: the interpreter's thread switch interval (default 5 ms)
        sys.setswitchinterval(float(os.environ["TCLIP_SWITCH_INTERVAL"]))
    if a.impl == "reference":
        run_reference(a)
    elif a.tasks > 0:
        bench_config5(a)
    elif a.method in KMEANS:
        bench_kmeans(a)
    else:
        bench_dirichlet(a)


def bench_dirichlet(a):
    h = Harness(a)
    torch, ops, dev, rank, world = h.torch, h.ops, h.dev, h.rank, h.world
    from tclip_b200 import tasks
    from tclip_b200.config import make_args
    from tclip_b200.methods.dirichlet import EM_DIRICHLET, HARD_EM_DIRICHLET

    name, iters = workload_name(a)
    K, T = a.classes, a.tasks_per_batch
    cls = EM_DIRICHLET if a.method == "em" else HARD_EM_DIRICHLET
    args = make_args(K, n_query=N_QUERY, iters=iters, mm_mode=a.mm_mode)
    n_steps = a.warmup + a.steps

    # synthetic batches, one per step and rank, generated before anything is timed; pinned host memory
    host = []
    for s in range(n_steps):
        td, _ = tasks.make_zero_shot_batch(T, K, n_query=N_QUERY, seed=SEED, batch_index=s * world + rank, **gen_kwargs(a))
        host.append({k: v.pin_memory() for k, v in td.items()})
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def step_resident(s):
        m = cls(model=None, device=dev, log_file=None, args=args)
        m.run_method(query=resident[s][0], y_q=resident[s][1])
        # keep the CUDA events and the small per-iteration counters only (a retained 300 MB alpha would force a cudaMalloc
        # in a later step); they are read after the timed region
        return (m._mm_events, m._em_events, m.mm_rows, m.mm_iters, m.n_live, torch.cat(m.test_acc, dim=1).mean())

    def step_e2e(s):
        m = cls(model=None, device=dev, log_file=None, args=args)
        return m.run_task(task_dic=dict(host[s]))

    timed = list(range(a.warmup, n_steps))
    # ---- leg 1: inputs resident in HBM ------------------------------------------------------------------------------
    resident = [(hh["x_q"].to(dev), hh["y_q"].long().squeeze(2).to(dev)) for hh in host]
    torch.cuda.synchronize()
    if a.warmup > 0:
        h.run_many(step_resident, [s % a.warmup for s in range(max(a.warmup, a.streams))])   # every stream gets warm
    h.warm_clocks()
    sampler = ClockSampler(h.local_rank)
    h.barrier()
    sampler.start()
    launches0 = ops.launch_count()
    kept, ms_resident, _ = h.timed(step_resident, timed)
    launches = ops.launch_count() - launches0
    accs = [float(k[5].item()) for k in kept]
    del kept

    # ---- leg 2: end to end through run_task with pinned host inputs ---------------------------------------------------
    if a.warmup > 0:
        h.run_many(step_e2e, [s % a.warmup for s in range(max(1, a.streams))])
    all_logs, ms_e2e, ms_e2e_rank = h.timed(step_e2e, timed)
    d2h_bytes = all_logs[-1]["acc"].nbytes + all_logs[-1]["criterions"].nbytes
    del all_logs

    # ---- the same two legs strictly one batch after the other (what the reference's evaluator loop does); the per-M-step
    # device times and work counters come from here (undisturbed by other streams) -------------------------------------
    if a.warmup > 0:
        step_resident(0)        # the default stream has not been used yet (its scratch and tensors are not allocated)
        torch.cuda.synchronize()
    kept, ms_resident_serial, ms_serial_rank = h.timed(step_resident, timed, serial=True)
    mm_ms, updates, dense_updates = 0.0, 0.0, 0.0
    it_ms, it_mm_ms, n_live_sum = [0.0] * iters, [0.0] * iters, [0.0] * iters
    for ev, (start, iter_ev), mm_rows, mm_iters, n_live, _acc in kept:
        for i in range(iters):
            w = ev[2 * i].elapsed_time(ev[2 * i + 1])
            mm_ms += w
            it_mm_ms[i] += w / len(kept)
            it_ms[i] += (start if i == 0 else iter_ev[i - 1]).elapsed_time(iter_ev[i]) / len(kept)
        for i, v in enumerate(n_live.tolist()):
            n_live_sum[i] += v / len(kept)
        updates += float(mm_rows.sum().item()) * K
        dense_updates += float(mm_iters.sum().item()) * T * K * K
    del kept
    if a.warmup > 0:
        step_e2e(0)             # the allocation pattern of run_task on the default stream, once, untimed
        torch.cuda.synchronize()
    _, ms_e2e_serial, _ = h.timed(step_e2e, timed, serial=True)
    clocks = sampler.stop()
    # diagnostics of the e2e leg: this rank's host->device copy of one batch on its own (pinned memory, CUDA events),
    # and every rank's e2e time (a slow PCIe path or a starved host core shows up here, not in the device-resident leg)
    per_rank = h.gather_list([ms_e2e_rank / a.steps, h.h2d_ms(host[a.warmup]["x_q"])])

    # ---- the dominant kernel alone: mm_chunk_kernel on a full batch of rows (T*K rows x D), two launches (51 + 50 MM
    # iterations, exactly the first two chunks of an M-step), CUDA events on the launching stream -------------------------
    xq0 = resident[a.warmup][0]
    logz0 = ops.log_features(xq0)
    colsum0, _, _ = ops.colsum_v(xq0)
    y0 = ops.moments(xq0, logz0, colsum0)                       # the moments of outer iteration 0 (u = z)
    alpha0 = torch.ones(T, K, K, device=dev)
    ops.mm_update_alpha(alpha0, y0, iter_mm=101, tol=0.0)       # warm-up
    kernel_ms = []
    for _ in range(3):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        ops.mm_update_alpha(alpha0, y0, iter_mm=101, tol=0.0)   # tol 0: the exit test never fires, both chunks run
        k1.record()
        k1.synchronize()
        kernel_ms.append(k0.elapsed_time(k1))
    kernel_ms = sorted(kernel_ms)[1]
    kernel_updates = float(T * K) * K * 101
    del xq0, logz0, colsum0, y0, alpha0

    # ---- roofline denominators: register-only FFMA / MUFU microbenchmarks, GPU still warm ---------------------------
    fp32_peak, mufu_peak = h.peaks()
    acc_mean = h.gather_mean(sum(accs) / len(accs))

    if rank == 0:
        tasks_total = T * a.steps * world
        value = tasks_total / (ms_resident * 1e-3)
        e2e = tasks_total / (ms_e2e * 1e-3)
        kernel_achieved = kernel_updates * FLOP_PER_UPDATE / (kernel_ms * 1e-3) / 1e12
        upd_step = updates / a.steps                     # element-updates one step (one batch of this rank) executes
        dense_step = dense_updates / a.steps
        step_ms, step_serial_ms = ms_resident / a.steps, ms_resident_serial / a.steps
        # whole job: `world` ranks each execute upd_step per step; per-GPU rate = upd_step / step time
        step_tflops = upd_step * FLOP_PER_UPDATE / (step_ms * 1e-3) / 1e12
        serial_tflops = upd_step * FLOP_PER_UPDATE / (step_serial_ms * 1e-3) / 1e12
        sm_mhz = clocks.get("sm_max_mhz") or 1965
        out = {
            "metric": metric_name(a), "value": value, "unit": "tasks/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "serial": {"value": tasks_total / (ms_resident_serial * 1e-3), "ms_per_step": step_serial_ms},
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "mm_mode": a.mm_mode, "tasks_per_step_per_gpu": T, "seed": SEED,
                       "streams": a.streams, "concurrency": concurrency_note(a),
                       "l2": "per-step working set alpha/y/work = 3 x %.0f MB > 126 MB L2; a different batch every step"
                             % (T * K * K * 4 / 1e6),
                       "mean_accuracy": acc_mean,
                       "n_live_per_task": [round(v / T, 2) for v in n_live_sum]},
            "e2e": {"value": e2e, "unit": "tasks/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / a.steps,
                    "serial": {"value": tasks_total / (ms_e2e_serial * 1e-3), "ms_per_step": ms_e2e_serial / a.steps},
                    "per_rank_ms_per_step": [round(x[0], 3) for x in per_rank],
                    "per_rank_h2d_ms": [round(x[1], 3) for x in per_rank]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "fp32-issue",
                "scope": "the timed step: every kernel of a run_task batch, %d batches in flight (value leg); per GPU" % a.streams,
                "achieved": step_tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": step_tflops / fp32_peak,
                "serial_frac": serial_tflops / fp32_peak,
                "how": "element-updates the step executed (device-side counters of every M-step: live rows x iterations + "
                       "free-running dead rows until their fixed point) x 74 flop / step time (CUDA events) / peak",
                "element_updates_executed_per_step": upd_step,
                "element_updates_executed_per_task": upd_step / T,
                "flop_per_element_update": FLOP_PER_UPDATE,
                "peak_source": "measured in this run: libtclip_b200 register-only FFMA microbenchmark "
                               "(MEASURED_PEAKS.json has only HBM and bf16-tensor peaks; this path is bound by neither)",
                "peak_theoretical": self_theoretical_fp32(h.n_sm, sm_mhz),
                "dense_equivalent": {
                    "what": "element-updates the reference's dense schedule performs for the same result (mm_iters x T x K x D); "
                            "the skip-dead schedule proves most of them redundant (DESIGN.md §3.2)",
                    "element_updates_per_task": dense_step / T,
                    "effective_tflops": dense_step * FLOP_PER_UPDATE / (step_ms * 1e-3) / 1e12},
                "kernel": "mm_chunk_kernel (Dirichlet MM M-step)",
                "kernel_frac": kernel_achieved / fp32_peak, "kernel_achieved": kernel_achieved,
                "kernel_how": "mm_chunk_kernel alone on a full batch of rows (%d x %d), 2 launches = 101 MM iterations, median of "
                              "3, CUDA events on the launching stream; algorithmic flop = element-updates x 74" % (T * K, K),
                "kernel_launch_ms": kernel_ms / 2, "kernel_element_updates_per_launch": kernel_updates / 2,
                "kernel_element_updates_per_s": kernel_updates / (kernel_ms * 1e-3),
                # dram__bytes_read.sum + dram__bytes_write.sum of one such launch (ncu --set full, profiles/r1_mm_chunk_kernel.md),
                # valid for the default shape only
                "traffic": 865.1e6 if (K == 1000 and T == 75) else None, "traffic_unit": "bytes per mm_chunk_kernel launch",
                "algorithmic_bytes_per_launch": 12.0 * T * K * K,
                "mufu_per_element_update": MUFU_PER_UPDATE,
                "kernel_mufu_achieved_tops": kernel_updates * MUFU_PER_UPDATE / (kernel_ms * 1e-3) / 1e12,
                "mufu_peak_tops": mufu_peak,
                "mm_share_of_serial_step": mm_ms / ms_serial_rank,
                "mm_element_updates_per_s_in_step": updates / (mm_ms * 1e-3),
                "per_outer_iteration": {"ms": [round(x, 3) for x in it_ms], "mm_ms": [round(x, 3) for x in it_mm_ms],
                                        "what": "serial leg, mean over the timed steps: CUDA events after every outer iteration "
                                                "and around every M-step (recorded by the C driver)"},
                "hbm_peak_gbs_measured": _measured_peaks().get("hbm_gbs"),
            },
        }
        if world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_sample(a, iters)
        print(json.dumps(out))
    h.close()


def self_theoretical_fp32(n_sm, sm_mhz):
    """148 SM x 128 FP32 lanes x 2 flop x f."""
    return n_sm * 128 * 2 * sm_mhz * 1e6 / 1e12


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE config 4: soft k-means / EM-Gaussian / hard k-means on RN50-shape visual features
# ----------------------------------------------------------------------------------------------------------------------
def bench_kmeans(a):
    h = Harness(a)
    torch, ops, dev, rank, world = h.torch, h.ops, h.dev, h.rank, h.world
    from oracle.ref_loader import StubTextModel, _install_clip_stub   # the stub `clip` tokenizer + text model (no CLIP offline)
    from tclip_b200 import tasks
    from tclip_b200.config import make_args
    from tclip_b200.methods import kmeans as M

    _install_clip_stub()
    name, iters = workload_name(a)
    K, T, D = a.classes, a.tasks_per_batch, KM_DIM
    cls = getattr(M, KMEANS[a.method][1])
    args = make_args(K, n_query=N_QUERY, iters=iters, use_softmax_feature=False)
    n_steps = a.warmup + a.steps
    host, txt = [], None
    for s in range(n_steps):
        td, txt = tasks.make_zero_shot_batch(T, K, n_query=N_QUERY, seed=SEED, batch_index=s * world + rank,
                                             softmax_feature=False, embed_dim=D)
        host.append({k: v.pin_memory() for k, v in td.items()})
    model = StubTextModel(txt)
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def step_resident(s):
        m = cls(model=model, device=dev, log_file=None, args=args)
        m.run_method(query=resident[s][0], y_q=resident[s][1])
        return (m._events, torch.cat(m.test_acc, dim=1).mean())

    def step_e2e(s):
        m = cls(model=model, device=dev, log_file=None, args=args)
        return m.run_task(task_dic=dict(host[s]))

    timed = list(range(a.warmup, n_steps))
    resident = [(hh["x_q"].to(dev), hh["y_q"].long().squeeze(2).to(dev)) for hh in host]
    torch.cuda.synchronize()
    if a.warmup > 0:
        h.run_many(step_resident, [s % a.warmup for s in range(max(a.warmup, a.streams))])
    h.warm_clocks()
    sampler = ClockSampler(h.local_rank)
    h.barrier()
    sampler.start()
    launches0 = ops.launch_count()
    kept, ms_resident, _ = h.timed(step_resident, timed)
    launches = ops.launch_count() - launches0
    accs = [float(k[1].item()) for k in kept]
    del kept
    if a.warmup > 0:
        h.run_many(step_e2e, [s % a.warmup for s in range(max(1, a.streams))])
    all_logs, ms_e2e, ms_e2e_rank = h.timed(step_e2e, timed)
    d2h_bytes = all_logs[-1]["acc"].nbytes + all_logs[-1]["criterions"].nbytes
    del all_logs
    if a.warmup > 0:
        step_resident(0)
        torch.cuda.synchronize()
    kept, ms_resident_serial, ms_serial_rank = h.timed(step_resident, timed, serial=True)
    loop_ms = sum(ev[0].elapsed_time(ev[-1]) for ev, _ in kept) / len(kept)      # the EM loop alone (events of the C driver)
    del kept
    if a.warmup > 0:
        step_e2e(0)
        torch.cuda.synchronize()
    _, ms_e2e_serial, _ = h.timed(step_e2e, timed, serial=True)
    clocks = sampler.stop()
    per_rank = h.gather_list([ms_e2e_rank / a.steps, h.h2d_ms(host[a.warmup]["x_q"])])
    acc_mean = h.gather_mean(sum(accs) / len(accs))

    fp32_peak, _ = h.peaks()
    if rank == 0:
        tasks_total = T * a.steps * world
        step_ms, step_serial_ms = ms_resident / a.steps, ms_resident_serial / a.steps
        hbm_peak = _measured_peaks().get("hbm_gbs") or 6546.2
        unit_bytes = 2.0 * K * D * 4                      # w [K,D] written + read once per task and iteration
        alg_bytes_step = unit_bytes * T * iters
        achieved = alg_bytes_step / (step_ms * 1e-3) / 1e9
        out = {
            "metric": metric_name(a), "value": tasks_total / (ms_resident * 1e-3), "unit": "tasks/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "serial": {"value": tasks_total / (ms_resident_serial * 1e-3), "ms_per_step": step_serial_ms},
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "tasks_per_step_per_gpu": T, "seed": SEED, "streams": a.streams,
                       "concurrency": concurrency_note(a),
                       "l2": "a different batch every step (x_q 30.7 MB + u 30 MB per batch); the loop's own state fits the L2 by design",
                       "mean_accuracy": acc_mean},
            "e2e": {"value": tasks_total / (ms_e2e * 1e-3), "unit": "tasks/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / a.steps,
                    "serial": {"value": tasks_total / (ms_e2e_serial * 1e-3), "ms_per_step": ms_e2e_serial / a.steps},
                    "per_rank_ms_per_step": [round(x[0], 3) for x in per_rank],
                    "per_rank_h2d_ms": [round(x[1], 3) for x in per_rank]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "hbm",
                "scope": "the timed step (initial assignment, Gram/Cholesky, %d loop iterations, prototypes, matching), %d batches in flight"
                         % (iters, a.streams),
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "serial_frac": alg_bytes_step / (step_serial_ms * 1e-3) / 1e9 / hbm_peak,
                "loop_frac": alg_bytes_step / (loop_ms * 1e-3) / 1e9 / hbm_peak, "loop_ms_serial": loop_ms,
                "algorithmic_bytes_per_unit": unit_bytes, "unit_is": "one task through one loop iteration in the reference's "
                "w-space formulation (w [K,D] written + read, SURVEY.md §8(d))",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if _measured_peaks().get("hbm_gbs") else "B200_PROFILING.md fallback",
                # dram bytes of one kproj_iter_kernel launch (ncu --set full, profiles/r2_kmeans.md): the sample-coordinate
                # loop keeps its state (0.34 MB per task) in L2 and moves far less than the w-space formulation's 8 MB
                "traffic": KM_TRAFFIC.get((K, T, D)), "traffic_unit": "bytes per kproj_iter_kernel launch (one iteration of all tasks)",
                "note": "the loop runs in the coordinates of the task's own samples (csrc/kmeans_run.cu), so a fraction above 1 of the "
                        "w-space HBM bound is possible: it measures the reformulation, not DRAM efficiency",
                # what the loop is really bound by: the FP32 rate of the iteration kernel in sample coordinates
                "loop_fp32": km_loop_fp32(N_QUERY, K, D, T, iters, loop_ms, fp32_peak),
            },
        }
        if world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_sample(a, iters)
        print(json.dumps(out))
    h.close()


def km_loop_fp32(n, K, D, T, iters, loop_ms, fp32_peak):
    """FP32 work of the k-means loop in sample coordinates (csrc/kmeans_run.cu), per iteration of one task, r = min(n, D)
    coordinates: centroids 2 K r n flop (multiply-add), distances 3 n K r flop (subtract, multiply-add).  The triangular
    form (D > n: Cholesky coordinates) runs the 16 x 16 blocks on and below the diagonal of the r x n index square only."""
    r = min(n, D)
    dense = 5.0 * K * r * n
    executed = dense
    if D > n:
        nb = -(-n // 16)
        blocks_c = sum(min(16, n - 16 * b) * 16 * (b + 1) for b in range(nb))          # sample block b x coordinate blocks <= b
        blocks_d = sum(min(16, r - 16 * b) * 16 * (nb - b) for b in range(nb))         # coordinate block b x sample groups >= b
        executed = 2.0 * K * blocks_c + 3.0 * K * blocks_d + 2.0 * K * r               # + the running sum of w^2
    tf = executed * T * iters / (loop_ms * 1e-3) / 1e12
    return {"bound": "fp32", "flop_per_task_iteration": executed, "dense_flop_per_task_iteration": dense,
            "achieved": tf, "dense_equivalent": dense * T * iters / (loop_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": tf / fp32_peak, "scope": "the loop of a strictly serial batch (loop_ms_serial), padding to 16-blocks counted as work",
            "peak_source": "FFMA issue-rate probe of this run (no FP32 line in MEASURED_PEAKS.json)"}


KM_TRAFFIC = {(1000, 100, 1024): 69.26e6}   # (K, T, D) -> dram__bytes_read.sum + dram__bytes_write.sum of one kproj_iter_kernel
                                             # launch (ncu --set full, profiles/r2_kmeans.md)


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE config 5: N ImageNet-shape EM-Dirichlet tasks sharded over the ranks, task construction in the loop
# ----------------------------------------------------------------------------------------------------------------------
def bench_config5(a):
    """The evaluator's loop (src/eval_zero_shot.py:151-180) for ``--tasks`` tasks: for every batch the index sampler
    (tclip_b200.tasks.ZeroShotQuerySampler, the reference's random calls), the device-side gather of the cached features
    (DeviceTaskSource -> tclip_gather_tasks), run_method, accuracy.  Batch i goes to rank i mod W (whole batches: the MM exit
    test is batch-global); the timed region is the whole job of a rank, max over ranks; STRONG scaling."""
    h = Harness(a)
    torch, ops, dev, rank, world = h.torch, h.ops, h.dev, h.rank, h.world
    import random
    from tclip_b200 import tasks
    from tclip_b200.config import make_args
    from tclip_b200.methods.dirichlet import EM_DIRICHLET, HARD_EM_DIRICHLET

    name, iters = workload_name(a)
    K, T = a.classes, a.tasks_per_batch
    cls = EM_DIRICHLET if a.method == "em" else HARD_EM_DIRICHLET
    args = make_args(K, n_query=N_QUERY, iters=iters, mm_mode=a.mm_mode)
    n_batches = a.tasks // T                                            # int(number_tasks / batch_size), eval_zero_shot.py:151
    mine = tasks.shard_batches(n_batches, rank, world)
    # the cached feature matrix of a synthetic "test set": 50 images per class, softmax features (generated on the device:
    # plumbing, not the product; same recipe as tasks.make_zero_shot_batch)
    per_class = 50
    g = torch.Generator(device=dev).manual_seed(SEED)
    txt = tasks.text_prototypes(K, SEED).to(dev)
    labels = torch.arange(K, device=dev).repeat_interleave(per_class)
    img = txt[labels] + tasks.NOISE_SCALE * torch.randn(labels.numel(), txt.shape[1], device=dev, generator=g) / txt.shape[1] ** 0.5
    img = img / img.norm(dim=-1, keepdim=True)
    feats = torch.softmax(30.0 * img @ txt.T, dim=-1)
    del img
    source = tasks.DeviceTaskSource(feats, labels, dev)
    labels_host = labels.cpu()
    # every batch has its own sampler seeded by the batch index, so that the job is the same whatever the number of ranks
    labels_host_index = tasks.ZeroShotQuerySampler.index_lists(labels_host, K)
    # the samplers draw from the global Python / torch RNGs: batches in flight would interleave their draws, so the index
    # lists of a rank's batches are drawn by the submitting thread, in order, inside the timed region
    def draw(b):
        random.seed(SEED * 7919 + b)
        torch.manual_seed(SEED * 7919 + b)
        sampler = tasks.ZeroShotQuerySampler(T, K, N_QUERY, labels_host_index, force_query_size=True)
        return torch.stack([torch.as_tensor(i, dtype=torch.int64) for i in sampler])

    def run_batch(idx):
        td = source.generate_from_indices(idx)
        m = cls(model=None, device=dev, log_file=None, args=args)
        m.run_method(query=td["x_q"], y_q=td["y_q"].squeeze(2))
        return torch.cat(m.test_acc, dim=1).mean()

    warm = [draw(b) for b in mine[:max(a.streams, 2)]]
    h.run_many(run_batch, warm)
    h.warm_clocks()
    sampler_clock = ClockSampler(h.local_rank)
    h.barrier()
    sampler_clock.start()
    launches0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.time()
    e0.record()
    t_draw = 0.0
    if h.pipe:
        futures = []
        for b in mine:
            t0 = time.time()
            idx = draw(b)
            t_draw += time.time() - t0
            futures.append(h.pipe.submit(run_batch, idx))
        accs = [f.result() for f in futures]
    else:
        accs = []
        for b in mine:
            t0 = time.time()
            idx = draw(b)
            t_draw += time.time() - t0
            accs.append(run_batch(idx))
    e1.record()
    h.barrier()
    wall = time.time() - t_host0
    ms = h.max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launch_count() - launches0
    clocks = sampler_clock.stop()
    acc_mean = h.gather_mean(float(torch.stack(accs).mean().item()) if accs else 0.0)
    per_rank = h.gather_list([e0.elapsed_time(e1), t_draw * 1e3, float(len(mine))])
    if rank == 0:
        n_tasks = n_batches * T
        out = {
            "metric": metric_name(a), "value": n_tasks / (ms * 1e-3), "unit": "tasks/s", "n_gpus": world,
            "steps": n_batches, "warmup": a.warmup, "ms_per_step": ms / max(len(mine), 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 5: %d tasks = %d run_task batches of %d (%s), sharded over %d rank(s) by whole "
                                   "batches; cached features [%d, %d] resident, index sampler + device-side gather + EM + accuracy "
                                   "inside the timed region" % (n_tasks, n_batches, T, name, world, feats.shape[0], K),
                       "mm_mode": a.mm_mode, "seed": SEED, "streams": a.streams, "concurrency": concurrency_note(a),
                       "mean_accuracy": acc_mean,
                       "per_rank": {"ms": [round(x[0], 1) for x in per_rank], "sampler_ms": [round(x[1], 1) for x in per_rank],
                                    "batches": [int(x[2]) for x in per_rank]}},
            "e2e": {"value": n_tasks / (ms * 1e-3), "unit": "tasks/s", "h2d_bytes_per_step": T * N_QUERY * 8,
                    "d2h_bytes_per_step": 4 + 4, "note": "only the sampled indices cross PCIe (features are resident); wall %.2f s" % wall},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(out))
    h.close()


def _measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


if __name__ == "__main__":
    main()

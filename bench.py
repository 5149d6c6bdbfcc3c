#!/usr/bin/env python
"""bench.py — EM-Dirichlet tasks/sec at ImageNet shape (K = D = 1000, n_query = 75) on N B200s, next to the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--method em|hard] [--mm-mode ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A *step* is one ``run_task`` batch of 75 synthetic zero-shot tasks (``batch_size`` 75) through the whole
hot path: H2D, the EM loop (``iter`` outer iterations of moments -> MM M-step -> E-step), cluster prototypes, Hungarian
label matching, accuracy.  Every rank runs its own, different batches (whole batches are the sharding unit because the
MM early exit is batch-global); ``value`` = tasks of all ranks / max-over-ranks device time.  Rank 0 prints ONE JSON line.

  value        inputs already resident in HBM (``run_method``), timed with CUDA events;
  e2e          the reference-facing call ``run_task(task_dic)`` with pinned HOST tensors: H2D + EM + D2H of the
               accuracies inside the timed region;
               both legs keep ``--streams`` (default 4) whole batches in flight per GPU — every batch is one unchanged
               ``run_method`` / ``run_task`` call on its own CUDA stream and host thread (``tclip_b200.pipeline``), because
               half of a batch is a latency-bound tail that leaves the SMs idle; ``serial`` inside ``value``'s line and
               inside ``e2e`` is the same leg strictly one batch after the other (the reference's evaluator loop);
  roofline     the dominant kernel (mm_chunk_kernel, the MM M-step) run alone on a full batch of rows (T*K rows x D, two
               launches = the first 101 MM iterations of an M-step): algorithmic flop (74 per element-update, SURVEY.md
               §8(d)) / launch time from CUDA events on the launching stream; peak = FP32 FMA issue rate measured by the
               library's register-only microbenchmark in the same run (MEASURED_PEAKS.json has no FP32 line).
               ``in_step`` carries the same counters of the M-steps inside the timed steps (CUDA events recorded by the
               C driver around every M-step);
  cpu_baseline the restated oracle (a port of the reference's CPU path; the Python reference itself cannot travel to
               the GPU box) on a bounded sample: 1 task, 8 outer iterations, extrapolated to ``iter`` iterations.
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
MUFU_PER_UPDATE = 4.0       # this kernel: rcp(X P), lg2 P, sqrt, rcp (tclip_math.cuh); ln X is a polynomial
SEED = 2020                 # the reference's default seed (config/datasets_config/*.yaml:10)
CPU_SAMPLE_ITERS = 8         # outer iterations of the bounded CPU sample (~10-20 s of CPU work at K=D=1000)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", default="em", choices=["em", "hard"],
                    help="em: EM-Dirichlet, iter 20 (the metric); hard: Hard EM-Dirichlet, iter 10 (BASELINE configs[1])")
    ap.add_argument("--mm-mode", default="skip_dead", choices=["skip_dead", "dense"])
    ap.add_argument("--classes", type=int, default=K_CLASSES)
    ap.add_argument("--tasks-per-batch", type=int, default=TASKS_PER_BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=4,
                    help="run_task batches in flight per GPU (own CUDA stream + host thread each); 1 = strictly serial")
    return ap.parse_args()


def workload_name(a):
    it = 20 if a.method == "em" else 10
    nm = "EM-Dirichlet" if a.method == "em" else "Hard EM-Dirichlet"
    return (f"{nm} zero-shot, synthetic ImageNet-shape softmax features (K=D={a.classes}, n_query={N_QUERY}, "
            f"batch_size {a.tasks_per_batch} tasks per run_task, iter {it}, iter_mm 1000)"), it


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
def cpu_sample(a, iters_full: int, dense_updates_per_task: float | None = None):
    """One bounded sample of the CPU path: the restated oracle on 1 task of the same workload, CPU_SAMPLE_ITERS outer
    iterations (outer iteration 0 exits its MM loop early, every later one runs all 1000 MM iterations),
    timed per outer iteration and extrapolated:  seconds/task = t_iter0 + (iter - 1) * mean(t_iter1..) + t_accuracy."""
    import torch
    from oracle import restated
    from tclip_b200 import tasks
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    td, _ = tasks.make_zero_shot_batch(1, a.classes, n_query=N_QUERY, seed=SEED, batch_index=10_000)
    t0 = time.time()
    n_it = min(CPU_SAMPLE_ITERS, iters_full)
    r = restated.dirichlet_zero_shot(td["x_q"], td["y_q"], a.classes, iters=n_it, hard=(a.method == "hard"))
    wall = time.time() - t0
    t_it0 = r.iter_seconds[0]
    t_it1 = sum(r.iter_seconds[1:]) / max(n_it - 1, 1)
    t_acc = max(wall - sum(r.iter_seconds), 0.0)
    per_task = t_it0 + (iters_full - 1) * t_it1 + t_acc
    updates = float(sum(r.mm_iters)) * a.classes * a.classes
    return {
        "value": 1.0 / per_task, "unit": "tasks/s", "cores": cores, "kind": "port",
        "sample": (f"oracle/restated.py (torch CPU fp32, {cores} threads) on 1 task K=D={a.classes}, {n_it} of {iters_full} outer "
                   f"iterations measured ({'+'.join(str(i) for i in r.mm_iters)} MM iterations, {wall:.1f} s), extrapolated as "
                   f"t_iter0 + {iters_full - 1} x mean(t_iter1..) + t_accuracy = {per_task:.1f} s/task"),
        "element_updates_per_s": updates / sum(r.iter_seconds), "seconds_measured": wall,
    }


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, iters_full = workload_name(a)
    vals, t_all = [], time.time()
    last = None
    for i in range(a.warmup + a.steps):
        # bounded: at K = 1000 one sample is ~15 s of CPU work; cap the whole arm near 3 minutes
        if i >= a.warmup:
            last = cpu_sample(a, iters_full)
            vals.append(last["value"])
        elif i == 0:
            cpu_sample(a, iters_full)          # one real warm-up (thread pool, allocator); the others are skipped
        if time.time() - t_all > 170 and vals:
            break
    v = sum(vals) / len(vals)
    last["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": "EM-Dirichlet tasks/sec (K=D=1000, N=75)", "value": v, "unit": "tasks/s",
        "n_gpus": a.gpus, "steps": len(vals), "warmup": a.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "note": "CPU arm: one step = one bounded sample (1 task, %d outer iterations, "
                   "extrapolated per task); host cores only, no GPU" % CPU_SAMPLE_ITERS},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return

    import torch
    import torch.distributed as dist
    from tclip_b200 import ops, tasks
    from tclip_b200.config import make_args
    from tclip_b200.methods.dirichlet import EM_DIRICHLET, HARD_EM_DIRICHLET

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a B200: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.device_check(local_rank)

    name, iters = workload_name(a)
    K, T = a.classes, a.tasks_per_batch
    cls = EM_DIRICHLET if a.method == "em" else HARD_EM_DIRICHLET
    args = make_args(K, n_query=N_QUERY, iters=iters, mm_mode=a.mm_mode)
    n_steps = a.warmup + a.steps

    # synthetic batches, one per step and rank, generated before anything is timed; pinned host memory
    host = []
    for s in range(n_steps):
        td, _ = tasks.make_zero_shot_batch(T, K, n_query=N_QUERY, seed=SEED, batch_index=s * world + rank)
        host.append({k: v.pin_memory() for k, v in td.items()})
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from tclip_b200.pipeline import BatchPipeline
    pipe = BatchPipeline(dev, streams=a.streams) if a.streams > 1 else None

    def run_many(fn, items):
        """Whole batches, up to --streams of them in flight (each on its own CUDA stream from its own host thread); every
        call is complete (stream synchronised) when it returns."""
        return pipe.map(fn, list(items)) if pipe else [fn(i) for i in items]

    def step_resident(s):
        m = cls(model=None, device=dev, log_file=None, args=args)
        m.run_method(query=resident[s][0], y_q=resident[s][1])
        # keep the CUDA events and the small per-iteration counters only (a retained 300 MB alpha would force a cudaMalloc
        # in a later step); they are read after the timed region
        return (m._mm_events, m.mm_rows, m.mm_iters, torch.cat(m.test_acc, dim=1).mean())

    def step_e2e(s):
        m = cls(model=None, device=dev, log_file=None, args=args)
        return m.run_task(task_dic=dict(host[s]))

    timed = list(range(a.warmup, n_steps))
    # ---- leg 1: inputs resident in HBM ------------------------------------------------------------------------------
    resident = [(h["x_q"].to(dev), h["y_q"].long().squeeze(2).to(dev)) for h in host]
    torch.cuda.synchronize()
    if a.warmup > 0:
        run_many(step_resident, [s % a.warmup for s in range(max(a.warmup, a.streams))])   # every stream gets warm
    # a fresh box idles at low clocks and W steps of ~50 ms do not always bring it up: ~0.5 s of register-only FMA work
    # (untimed) before the timed region, so both legs run at the clocks the sampler reports
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    for _ in range(50):
        ops.probe_issue_rate("ffma", n_sm * 8, 4000)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = ops.launch_count()
    # the default stream is idle during the legs, so these two events are processed the moment they are recorded:
    # before the first batch is submitted, and after every batch's stream has been synchronised
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kept = run_many(step_resident, timed)
    e1.record()
    barrier()
    launches = ops.launch_count() - launches0
    ms_resident = max_over_ranks(e0.elapsed_time(e1))
    accs = [float(k[3].item()) for k in kept]
    del kept

    # ---- leg 2: end to end through run_task with pinned host inputs ---------------------------------------------------
    if a.warmup > 0:
        run_many(step_e2e, [s % a.warmup for s in range(max(1, a.streams))])
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    all_logs = run_many(step_e2e, timed)
    e3.record()
    barrier()
    d2h_bytes = all_logs[-1]["acc"].nbytes + all_logs[-1]["criterions"].nbytes
    del all_logs

    # ---- the same two legs strictly one batch after the other (what the reference's evaluator loop does), for reference;
    # the per-M-step device times and work counters of `roofline.in_step` come from here (undisturbed by other streams)
    if a.warmup > 0:
        step_resident(0)        # the default stream has not been used yet (its scratch and tensors are not allocated)
        torch.cuda.synchronize()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    kept = [step_resident(s) for s in timed]
    e5.record()
    barrier()
    ms_resident_serial = max_over_ranks(e4.elapsed_time(e5))
    mm_ms, updates, dense_updates = 0.0, 0.0, 0.0
    for ev, mm_rows, mm_iters, _acc in kept:
        mm_ms += sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(iters))
        updates += float(mm_rows.sum().item()) * K
        dense_updates += float(mm_iters.sum().item()) * T * K * K
    del kept
    if a.warmup > 0:
        step_e2e(0)             # the allocation pattern of run_task on the default stream, once, untimed
        torch.cuda.synchronize()
    e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e6.record()
    for s in timed:
        step_e2e(s)
    e7.record()
    barrier()
    ms_e2e_serial = max_over_ranks(e6.elapsed_time(e7))
    clocks = sampler.stop()
    if pipe:
        pipe.close()
    ms_e2e_rank = e2.elapsed_time(e3)
    ms_e2e = max_over_ranks(ms_e2e_rank)
    # diagnostics of the e2e leg: this rank's host->device copy of one batch on its own (pinned memory, CUDA events),
    # and every rank's e2e time (a slow PCIe path or a starved host core shows up here, not in the device-resident leg)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(3):
        _x = host[a.warmup]["x_q"].to(dev, non_blocking=True)
    c1.record()
    c1.synchronize()
    h2d_ms_rank = c0.elapsed_time(c1) / 3
    del _x
    per_rank = [[ms_e2e_rank / a.steps, h2d_ms_rank]]
    if world > 1:
        tt = torch.tensor(per_rank[0], device=dev, dtype=torch.float64)
        gl = [torch.zeros_like(tt) for _ in range(world)] if rank == 0 else None
        dist.gather(tt, gl, dst=0)
        if rank == 0:
            per_rank = [g.tolist() for g in gl]

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
    ops.probe_issue_rate("ffma", n_sm * 8, 200)
    flop, ms_f = ops.probe_issue_rate("ffma", n_sm * 8, 4000)
    ops.probe_issue_rate("mufu", n_sm * 8, 50)
    mops, ms_m = ops.probe_issue_rate("mufu", n_sm * 8, 1000)
    fp32_peak = flop / (ms_f * 1e-3) / 1e12
    mufu_peak = mops / (ms_m * 1e-3) / 1e12

    # the one NCCL use of the path: gather the accuracies
    acc_mean = sum(accs) / len(accs)
    if world > 1:
        t = torch.tensor([acc_mean], device=dev)
        gathered = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, gathered, dst=0)
        if rank == 0:
            acc_mean = float(torch.cat(gathered).mean().item())

    if rank == 0:
        tasks_total = T * a.steps * world
        value = tasks_total / (ms_resident * 1e-3)
        e2e = tasks_total / (ms_e2e * 1e-3)
        achieved = kernel_updates * FLOP_PER_UPDATE / (kernel_ms * 1e-3) / 1e12
        out = {
            "metric": "EM-Dirichlet tasks/sec (K=D=1000, N=75)", "value": value, "unit": "tasks/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_resident / a.steps, "higher_is_better": True,
            "serial": {"value": tasks_total / (ms_resident_serial * 1e-3), "ms_per_step": ms_resident_serial / a.steps},
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "mm_mode": a.mm_mode, "tasks_per_step_per_gpu": T, "seed": SEED,
                       "streams": a.streams,
                       "concurrency": ("%d whole batches in flight per GPU, each one unchanged run_task / run_method call on "
                                       "its own CUDA stream and host thread (tclip_b200.pipeline); `serial` = one at a time, "
                                       "as the reference's evaluator loop" % a.streams),
                       "l2": "per-step working set alpha/y/work = 3 x %.0f MB > 126 MB L2; a different batch every step"
                             % (T * K * K * 4 / 1e6),
                       "mean_accuracy": acc_mean},
            "e2e": {"value": e2e, "unit": "tasks/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e / a.steps,
                    "serial": {"value": tasks_total / (ms_e2e_serial * 1e-3), "ms_per_step": ms_e2e_serial / a.steps},
                    "per_rank_ms_per_step": [round(x[0], 3) for x in per_rank],
                    "per_rank_h2d_ms": [round(x[1], 3) for x in per_rank]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "fp32-issue", "kernel": "mm_chunk_kernel (Dirichlet MM M-step)",
                "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one such launch (ncu --set full, profiles/r1_mm_chunk_kernel.md),
                # valid for the default shape only
                "traffic": 865.1e6 if (K == 1000 and T == 75) else None, "traffic_unit": "bytes per launch",
                "peak_source": "measured in this run: libtclip_b200 register-only FFMA microbenchmark "
                               "(MEASURED_PEAKS.json has only HBM and bf16-tensor peaks; this kernel is bound by neither)",
                "how": "mm_chunk_kernel alone on a full batch of rows (%d x %d), 2 launches = 101 MM iterations, median of "
                       "3, CUDA events on the launching stream; algorithmic flop = element-updates x 74" % (T * K, K),
                "launch_ms": kernel_ms / 2, "element_updates_per_launch": kernel_updates / 2,
                "element_updates_per_s": kernel_updates / (kernel_ms * 1e-3),
                "flop_per_element_update": FLOP_PER_UPDATE,
                "mufu_achieved_tops": kernel_updates * MUFU_PER_UPDATE / (kernel_ms * 1e-3) / 1e12,
                "mufu_peak_tops": mufu_peak,
                "algorithmic_bytes_per_launch": 12.0 * T * K * K,
                "in_step": {"mm_share_of_step": mm_ms / (e4.elapsed_time(e5)),
                            "element_updates_per_s": updates / (mm_ms * 1e-3),
                            "element_updates_executed_per_task": updates / (T * a.steps),
                            "element_updates_dense_per_task": dense_updates / (T * a.steps)},
                "hbm_peak_gbs_measured": _measured_peaks().get("hbm_gbs"),
            },
        }
        if world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_sample(a, iters)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


if __name__ == "__main__":
    main()

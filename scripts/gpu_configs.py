"""Scratch: end-to-end tasks/s of the BASELINE.json configs through run_task (pinned host tensors, strictly serial calls,
3 timed batches after 1 warm-up each).  Synthetic inputs of the named shapes; parity of every path is in tests/."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import logging; logging.disable(logging.INFO)
import torch
from tclip_b200 import tasks
from tclip_b200.config import make_args
from tclip_b200.methods import dirichlet as D, kmeans as KM
dev = torch.device("cuda:0")

class StubText:
    def __init__(self, txt): self.txt = txt
    def encode_text(self, tokens): return self.txt[tokens.reshape(-1).long().cpu()].clone().to(tokens.device)

import types
sys.modules.setdefault("clip", types.ModuleType("clip"))
sys.modules["clip"].tokenize = lambda texts: torch.arange(len(texts)).unsqueeze(1)

def timed(name, make, run, n_tasks, reps=3):
    batches = [make(i) for i in range(reps + 1)]
    run(batches[0]); torch.cuda.synchronize()
    t0 = time.time()
    accs = []
    for b in batches[1:]:
        accs.append(float(run(b)["acc"].mean()))
    torch.cuda.synchronize()
    dt = (time.time() - t0) / reps
    print(f"{name:78s} {n_tasks / dt:9.1f} tasks/s  ({dt * 1e3:8.1f} ms per batch of {n_tasks}, mean acc {sum(accs) / len(accs):.3f})", flush=True)

def pin(td): return {k: v.pin_memory() for k, v in td.items()}

def timed_in_flight(name, batches, run, n_tasks, streams):
    """The same run_task calls with `streams` batches in flight (tclip_b200.pipeline), 2 rounds after 1 warm-up round."""
    from tclip_b200.pipeline import BatchPipeline
    with BatchPipeline(dev, streams=streams) as pipe:
        pipe.map(run, [batches[i % len(batches)] for i in range(streams)])
        torch.cuda.synchronize(); t0 = time.time()
        n = 2 * streams
        pipe.map(run, [batches[i % len(batches)] for i in range(n)])
        torch.cuda.synchronize(); dt = (time.time() - t0) / n
    print(f"{name:78s} {n_tasks / dt:9.1f} tasks/s  ({dt * 1e3:8.1f} ms per batch of {n_tasks}, {streams} batches in flight)", flush=True)

# config 1: EM-Dirichlet zero-shot, Caltech101 shape
a = make_args(100, iters=20)
timed("cfg1 EM-Dirichlet zero-shot K=D=100, batch 100", lambda i: pin(tasks.make_zero_shot_batch(100, 100, seed=2020, batch_index=i)[0]),
      lambda td: D.EM_DIRICHLET(model=None, device=dev, log_file=None, args=a).run_task(dict(td)), 100)
timed_in_flight("cfg1 in flight", [pin(tasks.make_zero_shot_batch(100, 100, seed=2020, batch_index=i)[0]) for i in range(4)],
                lambda td: D.EM_DIRICHLET(model=None, device=dev, log_file=None, args=a).run_task(dict(td)), 100, 8)
# config 2: Hard EM-Dirichlet zero-shot, ImageNet shape, iter 10
a2 = make_args(1000, iters=10)
timed("cfg2 Hard EM-Dirichlet zero-shot K=D=1000, batch 75, iter 10", lambda i: pin(tasks.make_zero_shot_batch(75, 1000, seed=2020, batch_index=i)[0]),
      lambda td: D.HARD_EM_DIRICHLET(model=None, device=dev, log_file=None, args=a2).run_task(dict(td)), 75)
# config 5 shape (the bench metric): EM-Dirichlet zero-shot, ImageNet shape, iter 20
a5 = make_args(1000, iters=20)
timed("cfg5 EM-Dirichlet zero-shot K=D=1000, batch 75, iter 20", lambda i: pin(tasks.make_zero_shot_batch(75, 1000, seed=2020, batch_index=i)[0]),
      lambda td: D.EM_DIRICHLET(model=None, device=dev, log_file=None, args=a5).run_task(dict(td)), 75)
# config 3: EM-Dirichlet 4-shot few-shot, ImageNet shape (every row live in every M-step), at the evaluator's batch size 75:
# 15 distinct tasks tiled 5 times (the float64 host generator needs ~1 s per 4000-sample support set)
a3 = make_args(1000, iters=20, k_eff=5)
_fs = {}
def mk_few(i):
    if i not in _fs:
        td = tasks.make_few_shot_batch(15, 1000, shots=4, seed=2020, batch_index=i)[0]
        _fs[i] = pin({k: v.repeat(5, *([1] * (v.dim() - 1))).contiguous() for k, v in td.items()})
    return _fs[i]
def run_few(td):
    m = D.FEW_SHOT_EM_DIRICHLET(model=None, device=dev, log_file=None, args=a3)
    logs = m.run_task(dict(td), shot=4)
    run_few.updates = float(m.mm_iters.sum().item()) * 75 * 1000 * 1000
    return logs
t_before = time.time()
timed("cfg3 EM-Dirichlet 4-shot few-shot K=D=1000 (S=4000), batch 75", mk_few, run_few, 75, reps=2)
from tclip_b200 import ops
n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
ops.probe_issue_rate("ffma", n_sm * 8, 4000); flop, ms = ops.probe_issue_rate("ffma", n_sm * 8, 4000)
td = mk_few(1); torch.cuda.synchronize(); t0 = time.time(); run_few(td); torch.cuda.synchronize(); dt = time.time() - t0
print(f"     cfg3 at T=75: {run_few.updates:.3e} element-updates per batch in {dt * 1e3:.1f} ms = "
      f"{run_few.updates * 74 / dt / 1e12 / (flop / (ms * 1e-3) / 1e12):.2f} of the FP32 peak (whole run_task incl. 1.2 GB H2D)", flush=True)
timed_in_flight("cfg3 in flight", [_fs[i] for i in sorted(_fs)], run_few, 75, 3)
# config 4: soft k-means / EM-Gaussian on visual features D=1024, K=1000, batch 100
for cls, nm in ((KM.SOFT_KMEANS, "soft k-means"), (KM.EM_GAUSSIAN, "EM-Gaussian"), (KM.HARD_KMEANS, "hard k-means (iter 10)")):
    a4 = make_args(1000, iters=10 if "hard" in nm else 20, use_softmax_feature=False)
    made = {}
    def mk(i):
        td, txt = tasks.make_zero_shot_batch(100, 1000, seed=2020, batch_index=i, softmax_feature=False, embed_dim=1024)
        made["txt"] = txt
        return pin(td)
    timed(f"cfg4 {nm} visual features D=1024, K=1000, batch 100", mk,
          lambda td: cls(model=StubText(made["txt"]), device=dev, log_file=None, args=a4).run_task(dict(td)), 100)

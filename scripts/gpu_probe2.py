"""Scratch: issue-rate probes + quick M-step timing (not a test)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import torch
from tclip_b200 import ops
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
nsm = torch.cuda.get_device_properties(0).multi_processor_count
for which, it in (("ffma", 4000), ("ffma2", 4000), ("mufu", 1000), ("mix", 2000)):
    ops.probe_issue_rate(which, nsm * 8, 100)
    n, ms = ops.probe_issue_rate(which, nsm * 8, it)
    print(f"probe {which}: {n:.3e} ops in {ms:.3f} ms -> {n / ms / 1e9:.3f} T/s", flush=True)
# M-step alone, dense, K = D = 1000, 75 tasks -> 75000 rows
g = torch.Generator().manual_seed(0)
for rows, D in ((75000, 1000), (3750, 1000), (20000, 100)):
    y = torch.log(torch.softmax(3 * torch.randn(rows, 4, D, generator=g), -1)).mean(1).to(dev)
    a0 = torch.ones(rows, D, device=dev)
    for iter_mm in (50, 200):
        ops.mm_update_alpha(a0, y, iter_mm=50, check_every=0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out, iters = ops.mm_update_alpha(a0, y, iter_mm=iter_mm, check_every=0); e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1); upd = rows * D * iter_mm
        print(f"mm rows={rows} D={D} iters={iter_mm}: {ms:.2f} ms -> {upd / (ms * 1e-3) / 1e9:.1f} G elem-updates/s", flush=True)

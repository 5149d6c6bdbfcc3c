"""Scratch: when does the fp32 trajectory of an empty cluster's row (y = -10, start = alpha after outer iteration 0)
become periodic on the GPU, and with which period?  States after t = 0..60 iterations via the dense kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import torch
from tclip_b200 import tasks, ops
dev = torch.device("cuda:0")
K, T = 1000, 8
td, _ = tasks.make_zero_shot_batch(T, K, seed=2020)
xq = td["x_q"].to(dev)
logz = ops.log_features(xq)
colsum, _, _ = ops.colsum_v(xq)
y0 = ops.moments(xq, logz, colsum)
a0, it0 = ops.mm_update_alpha(torch.ones(T, K, K, device=dev), y0, iter_mm=1000)   # exits at 51 like the real batch
print("iteration-0 M-step ran", int(it0.item()), "iterations")
rows = a0.reshape(T * K, K).contiguous()
y = torch.full_like(rows, -10.0)
states = [rows.clone()]
cur = rows
for t in range(1, 61):
    cur, _ = ops.mm_update_alpha(cur, y, iter_mm=1, check_every=0)
    states.append(cur.clone())
first = torch.full((T * K,), -1, dtype=torch.long, device=dev)
period = torch.zeros(T * K, dtype=torch.long, device=dev)
for t in range(1, 61):
    for p in (1, 2, 3, 4, 5, 6):
        if t - p < 0: continue
        same = (states[t] == states[t - p]).all(dim=1) & (first < 0)
        first[same] = t
        period[same] = p
import collections
f = first.cpu().tolist(); p = period.cpu().tolist()
print("rows never periodic within 60:", sum(1 for x in f if x < 0))
print("histogram of first t with state(t) == state(t-p):", sorted(collections.Counter(f).items()))
print("periods:", sorted(collections.Counter(p).items()))

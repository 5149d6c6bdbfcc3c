"""Scratch (ncu target): three launches of the tcgen05 contraction at ImageNet shape on a heavy-tailed synthetic alpha."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))
import torch
from tclip_b200 import ops
dev = torch.device("cuda:0")
T, n, K = 75, 75, 1000
g = torch.Generator(device=dev).manual_seed(0)
logz = torch.log(torch.softmax(5 * torch.randn(T, n, K, device=dev, generator=g), -1) + 1e-15)
alpha = (0.03 + torch.exp(2.5 * torch.randn(T, K, K, device=dev, generator=g))).clamp(max=3e5)
for _ in range(3):
    out = ops.contraction(logz, alpha, "tcgen05")
torch.cuda.synchronize()
print("ok", float(out.abs().max()))

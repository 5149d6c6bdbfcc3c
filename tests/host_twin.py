"""TEST INFRASTRUCTURE: the oracle's EM loop with its M-step replaced by the CPU twin of the CUDA kernel's arithmetic
(csrc/tclip_math.cuh compiled by tests/build_host_shim.sh).  Lets the CPU suite check that the kernel's series /
packed formulation reproduces the reference's MM iteration counts, labels and alpha before any GPU time is spent.
MUFU approximations are exact libm calls here; the -m gpu tests cover the hardware approximations."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

from oracle import restated as R

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "_build", "libtclip_host_math.so")
_lib = None


def shim():
    global _lib
    if _lib is None:
        if not os.path.isfile(SHIM):
            subprocess.run(["bash", os.path.join(HERE, "build_host_shim.sh")], check=True)
        _lib = ctypes.CDLL(SHIM)
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        _lib.tclip_host_mm_rows.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return _lib


def mm_update_alpha_twin(alpha0, y_cst, iter_mm, check_every=R.MM_CHECK_EVERY, tol=R.MM_TOL):
    """Same chunking as csrc/dirichlet_mm.cu: chunks end at l = 50, 100, ... where the batch-global criterion of the
    last iteration is tested on float32 sums."""
    lib = shim()
    shape = alpha0.shape
    D = shape[-1]
    Dp = D + (D & 1)                                        # the twin works on pairs
    a = np.ones((alpha0.numel() // D, Dp), dtype=np.float32)
    y = -np.ones_like(a)
    a[:, :D] = alpha0.reshape(-1, D).numpy()
    y[:, :D] = y_cst.reshape(-1, D).numpy()
    rows = a.shape[0]
    start = 0
    while start < iter_mm:
        cand = ((max(start, 1) + check_every - 1) // check_every) * check_every
        end, has_check = (cand, True) if cand <= iter_mm - 1 else (iter_mm - 1, False)
        n = end - start + 1
        if n > 1:
            lib.tclip_host_mm_rows(a, y, rows, Dp, D, n - 1)
        prev = a.copy()
        lib.tclip_host_mm_rows(a, y, rows, Dp, D, 1)
        start = end + 1
        if has_check:
            num = float(((a[:, :D].astype(np.float64) - prev[:, :D]) ** 2).sum())
            den = float((prev[:, :D].astype(np.float64) ** 2).sum())
            if np.float32(num) / np.float32(den) < tol:
                break
    return torch.from_numpy(a[:, :D].copy()).reshape(shape), start


class patched_oracle:
    """Context manager: oracle.restated with the twin M-step (float32 only)."""

    def __enter__(self):
        self._orig = R.mm_update_alpha
        R.mm_update_alpha = lambda a0, y, iter_mm, check_every=R.MM_CHECK_EVERY, tol=R.MM_TOL: \
            mm_update_alpha_twin(a0, y, iter_mm, check_every, tol)
        return self

    def __exit__(self, *exc):
        R.mm_update_alpha = self._orig
        return False


def few_shot_tasks_on_host(fs, ls, fq, lq, idx_s, idx_q):
    """``Tasks_Generator_few_shot.get_task`` + ``generate_tasks`` (src/task_generator_few_shot.py:27-99, softmax features)
    written out with torch indexing on the host: the checker of ``tasks.DeviceFewShotTaskSource``.  Validated against the
    reference's own class in tests/test_host_logic.py."""
    import torch
    out = {"x_s": [], "y_s": [], "x_q": [], "y_q": []}
    for i_s, i_q in zip(idx_s, idx_q):
        lab_s, lab_q = ls[i_s], lq[i_q]
        uniq = torch.flip(torch.unique(lab_s, sorted=False), dims=(0,))
        ns, nq = torch.zeros_like(lab_s), torch.zeros_like(lab_q)
        for j, y in enumerate(uniq):
            ns[lab_s == y] = j
            nq[lab_q == y] = j
        out["x_s"].append(fs[i_s][:, uniq]); out["y_s"].append(ns.long())
        out["x_q"].append(fq[i_q][:, uniq]); out["y_q"].append(nq.long())
    return {k: torch.stack(v) for k, v in out.items()}

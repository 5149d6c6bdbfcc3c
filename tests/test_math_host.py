"""CPU: the M-step's special-function arithmetic (csrc/tclip_math.cuh compiled as plain C++ by
tests/build_host_shim.sh) against SciPy in float64, and the host twin of the MM inner loop against the oracle."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
from scipy import special

from oracle import restated as R

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "_build", "libtclip_host_math.so")


@pytest.fixture(scope="module")
def shim():
    if not os.path.isfile(SHIM):
        subprocess.run(["bash", os.path.join(HERE, "build_host_shim.sh")], check=True)
    lib = ctypes.CDLL(SHIM)
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    lib.tclip_host_psi1_N.argtypes = [f32p, f32p, f32p, ctypes.c_int]
    lib.tclip_host_mm_update.argtypes = [f32p, f32p, f32p, ctypes.c_int, ctypes.c_double]
    lib.tclip_host_mm_update_pair.argtypes = [f32p, f32p, f32p, ctypes.c_int, ctypes.c_double]
    lib.tclip_host_mm_update_pair_split.argtypes = [f32p, f32p, f32p, ctypes.c_int, ctypes.c_double]
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    lib.tclip_host_row_psi_walk.argtypes = [f64p, f32p, f32p, f32p, ctypes.c_int]
    lib.tclip_host_row_psi_walk.restype = ctypes.c_int
    lib.tclip_host_digamma.argtypes = [ctypes.c_double]
    lib.tclip_host_digamma.restype = ctypes.c_double
    lib.tclip_host_mm_rows.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.tclip_host_mm_rows_anchored.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.tclip_host_mm_rows_anchored.restype = ctypes.c_int
    return lib


def _grid():
    return np.concatenate([np.logspace(-12, -1, 200), np.linspace(0.05, 0.08, 50), np.logspace(-1, 6, 600)]).astype(np.float32)


def test_psi_and_curvature_numerator(shim):
    a = _grid()
    psi1, N = np.empty_like(a), np.empty_like(a)
    shim.tclip_host_psi1_N(a, psi1, N, a.size)
    a64 = a.astype(np.float64)
    psi_ref = special.digamma(a64 + 1)
    n_ref = a64 * psi_ref - special.gammaln(a64 + 1)
    # below ~1e-3 the float64 difference above cancels catastrophically: use the Taylor series
    # N(a) = sum_{k>=2} (-1)^k zeta(k) (1 - 1/k) a^k as the yard-stick there
    small = a64 < 1e-3
    n_ref[small] = sum((-1) ** k * special.zeta(k) * (1 - 1 / k) * a64[small] ** k for k in range(2, 9))
    assert np.max(np.abs(psi1 - psi_ref) / np.maximum(np.abs(psi_ref), 1.0)) < 1e-6
    # N = a psi(a+1) - lnGamma(a+1) >= 0, ~ a^2 pi^2/12 for small a, ~ a for large a
    assert np.max(np.abs(N - n_ref) / n_ref) < 2e-4   # worst just above the Taylor/Stirling switch at a = 1/16
    assert np.max(np.abs(N - n_ref)[a > 1] / n_ref[a > 1]) < 2e-6


def test_digamma_f64(shim):
    """row_psi: psi(s) = k ln2 + dpsi with dpsi carried as ONE float32 (|dpsi| < 0.4 + 1/(2s) for s >= 10), so the
    reconstruction is good to a float32 ulp of dpsi, i.e. ~3e-8 absolute where it matters (large row totals)."""
    for s in (1e-3, 0.5, 1.0, 9.99, 10.0, 123.4, 1e5, 3e8):
        ref = special.digamma(s)
        got = shim.tclip_host_digamma(s)
        k = np.floor(np.log2(s) + 0.5)
        assert abs(got - ref) <= 1.2e-7 * max(abs(ref - k * np.log(2)), 0.25), (s, got, ref)


def test_one_mm_update_matches_the_reference_formula(shim):
    """a_new = (-b + sqrt(b^2 + 4c)) / (2c) in float64 with torch's special functions (em_dirichlet.py:153-167)."""
    g = np.random.default_rng(0)
    a = _grid()
    a = a[(a >= 1e-3) | (a <= 1e-11)]   # in between the float64 evaluation of c below cancels; covered by the N test
    y = -g.uniform(0.5, 12.0, a.size).astype(np.float32)
    s = 37.5
    out = np.empty_like(a)
    out2 = np.empty_like(a)
    a = a[: a.size // 2 * 2]
    y, out, out2 = y[: a.size], out[: a.size], out2[: a.size]
    shim.tclip_host_mm_update(a, y, out, a.size, float(special.digamma(s)))
    shim.tclip_host_mm_update_pair(a, y, out2, a.size, s)   # the kernel's packed shift-4 form (takes the row total)
    a64, y64 = torch.from_numpy(a).double(), torch.from_numpy(y).double()
    psi1 = torch.polygamma(0, a64 + 1)
    c = torch.where(a64 > 1e-11, (2 * (-torch.lgamma(a64 + 1) + psi1 * a64) / a64 ** 2).abs(),
                    torch.polygamma(1, torch.ones(1, dtype=torch.float64)))
    b = psi1 - special.digamma(s) - c * a64 - y64
    ref = ((-b + torch.sqrt(b * b + 4 * c)) / (2 * c)).numpy()
    # away from the fixed point one step inherits the relative error of the curvature c = 2N/a^2, which is ~1e-4 just
    # above the Taylor/Stirling switch (cancellation in N, see test_psi_and_curvature_numerator) and ~1e-6 elsewhere;
    # the fixed point itself does not depend on c
    for got in (out, out2):
        err = np.abs(got - ref) / ref
        assert err.max() < 3e-4
        assert err[(a < 0.05) | (a > 1.0)].max() < 5e-6


def test_two_phase_update_is_bit_identical(shim):
    """mm_update_pre + mm_update_post (the software-pipelined form of mm_spec_kernel) == mm_update_pair, bit for bit,
    on both sides of the small-a switch and for small and large row totals."""
    g = np.random.default_rng(1)
    a = np.concatenate([_grid(), g.uniform(0.01, 0.2, 400).astype(np.float32), g.uniform(0.5, 3e5, 400).astype(np.float32)])
    a = np.ascontiguousarray(a[: a.size // 2 * 2])
    y = -g.uniform(0.1, 35.0, a.size).astype(np.float32)
    for s in (0.7, 12.5, 1000.0, 2.5e5):
        out, out2 = np.empty_like(a), np.empty_like(a)
        shim.tclip_host_mm_update_pair(a, y, out, a.size, s)
        shim.tclip_host_mm_update_pair_split(a, y, out2, a.size, s)
        assert np.array_equal(out.view(np.uint32), out2.view(np.uint32)), s


def test_anchored_row_psi_follows_the_full_evaluation(shim):
    """row_psi_anchored (Taylor expansion around an anchor total, used by mm_spec_kernel) against the stateless float64
    evaluation along realistic walks of a row total: slow drift, growth of a diverging row, jumps, tiny totals."""
    g = np.random.default_rng(5)
    walks = {
        "drift": 1.7e5 * np.cumprod(1 + 1e-6 * g.standard_normal(4000)),
        "growth": 40.0 * np.cumprod(np.full(4000, 1 + 3e-3)),
        "slow_growth": 900.0 * np.cumprod(np.full(4000, 1 + 2e-5)),
        "jumps": np.abs(g.standard_normal(2000)) * 1e4 + 20.0,
        "small": np.linspace(0.3, 25.0, 3000),
    }
    for name, s in walks.items():
        s = np.ascontiguousarray(s, dtype=np.float64)
        a, f, dk = (np.empty(s.size, np.float32) for _ in range(3))
        n_full = shim.tclip_host_row_psi_walk(s, a, f, dk, s.size)
        # psi(s) = dpsi + k ln2: compare the reconstructed values (the anchored form keeps the anchor's k)
        rec_a = a.astype(np.float64) + (dk.astype(np.float64) / 8388608.0) * np.log(2.0)
        err = np.abs(rec_a - f.astype(np.float64))
        assert err.max() <= 6.5e-8, (name, err.max())                  # one float32 ulp of |dpsi| < 0.75
        same_k = dk == 0
        assert np.mean(a[same_k] == f[same_k]) > 0.97, (name, np.mean(a[same_k] == f[same_k]))
        ref = special.digamma(s)
        k = (dk.astype(np.float64) / 8388608.0)
        del ref, k
        if name in ("drift", "slow_growth"):
            assert n_full <= 8, (name, n_full)                          # the logarithm is almost never evaluated
        if name == "growth":
            assert n_full < s.size // 3, (name, n_full)


def test_mm_rows_track_the_oracle(shim):
    g = torch.Generator().manual_seed(3)
    rows, D, iters = 6, 40, 120
    y = torch.log(torch.softmax(2 * torch.randn(rows, 6, D, generator=g), -1)).mean(1).contiguous()
    a = np.ones((rows, D), dtype=np.float32)
    shim.tclip_host_mm_rows(a, y.numpy(), rows, D, D, iters)
    ref, done = R.mm_update_alpha(torch.ones(rows, D, dtype=torch.float64), y.double(), iters, check_every=0 or 10 ** 9)
    assert done == iters
    assert np.max(np.abs(a - ref.numpy()) / ref.numpy()) < 2e-5


def test_mm_rows_with_anchored_psi_track_the_stateless_form(shim):
    """1000 MM iterations with psi(s) from the anchored expansion vs the float64 evaluation every iteration: same alpha to
    float32 round-off (no drift), on converging rows and on a diverging singleton-cluster row, with few re-anchorings."""
    g = torch.Generator().manual_seed(4)
    rows, D, iters = 5, 200, 1000
    z = torch.softmax(3 * torch.randn(rows, 7, D, generator=g), -1)
    y = torch.log(z + 1e-15).mean(1)
    y[0] = torch.log(z[0, 0] + 1e-15)          # a single sample: the Dirichlet MLE diverges, alpha keeps growing
    y = y.contiguous().numpy()
    a_full = np.ones((rows, D), dtype=np.float32)
    a_anch = np.ones((rows, D), dtype=np.float32)
    shim.tclip_host_mm_rows(a_full, y, rows, D, D, iters)
    n_full = shim.tclip_host_mm_rows_anchored(a_anch, y, rows, D, D, iters)
    rel = np.abs(a_anch - a_full) / a_full
    assert rel[1:].max() < 5e-6, rel[1:].max()
    assert rel[0].max() < 2e-4, rel[0].max()                    # the diverging row amplifies every rounding difference
    assert np.isfinite(a_anch).all() and a_anch[0].sum() > 20 * a_anch[1].sum()
    assert n_full < rows * iters // 4, n_full

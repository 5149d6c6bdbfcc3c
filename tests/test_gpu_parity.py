"""GPU parity tests (run with ``-m gpu`` on a B200): the CUDA path behind the reference-facing method classes and the
C ABI vs (a) the golden vectors frozen from the live reference and (b) the restated CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star):
  * hard labels argmax_k u agree on >= 99.9 % of queries;
  * mean task accuracy within 0.1 pt;
  * MM iterations executed per outer iteration identical (the batch-global early exit, em_dirichlet.py:169-175);
  * alpha: per-task Frobenius relative error vs the float64 restatement no worse than ALPHA_VS_FP64 (the reference's
    own fp32-vs-fp64 gap is 2-3e-4 at 20 outer iterations, SURVEY.md §0.5, so the flat 1e-4 of north_star is checked on
    short runs and on non-diverging rows, and the GPU error is required to be <= 2x the reference-fp32 error elsewhere).
"""
from __future__ import annotations

import functools
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import restated as R  # noqa: E402  (test infrastructure: the checker)
from oracle.ref_loader import make_args  # noqa: E402

LABEL_AGREE = 0.999
ACC_TOL = 1e-3          # 0.1 pt
ALPHA_REL = 1e-4


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device (tclip_b200 has no CPU path)")
    from tclip_b200 import ops
    ops.device_check(0)
    return torch.device("cuda:0")


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def _classes():
    from tclip_b200.methods import dirichlet as D
    return {("zero_shot", "EM_DIRICHLET"): D.EM_DIRICHLET, ("zero_shot", "HARD_EM_DIRICHLET"): D.HARD_EM_DIRICHLET,
            ("few_shot", "EM_DIRICHLET"): D.FEW_SHOT_EM_DIRICHLET,
            ("few_shot", "HARD_EM_DIRICHLET"): D.FEW_SHOT_HARD_EM_DIRICHLET}


GOLDEN_DIRICHLET = sorted(os.path.basename(p)[:-4] for p in glob.glob(
    os.path.join(os.path.dirname(__file__), "golden", "*dirichlet*.npz")))


# ------------------------------------------------------------------------------------------------------------------
# golden vectors frozen from the live reference
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN_DIRICHLET)
@pytest.mark.parametrize("mode", ["dense", "skip_dead"])
def test_golden_dirichlet(dev, golden_dir, name, mode):
    g = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=True)
    setting, method, K, iters = str(g["setting"]), str(g["method"]), int(g["K"]), int(g["iters"])
    cls = _classes()[(setting, method)]
    args = make_args(K, iters=iters, k_eff=int(g["k_eff"]), mm_mode=mode)
    m = cls(model=None, device=dev, log_file=None, args=args)
    td = {k: torch.from_numpy(g[k]) for k in ("x_q", "y_q", "x_s", "y_s") if k in g.files}
    logs = m.run_task(td, shot=int(g["shots"])) if setting == "few_shot" else m.run_task(td)

    assert logs["acc"].shape == g["acc"].shape and logs["acc"].dtype == np.float32
    assert logs["criterions"].shape == g["criterions"].shape
    assert m.mm_iters.cpu().tolist() == g["mm_iters"].tolist()
    agree = (m.labels.cpu().long().numpy() == g["preds"]).mean()
    assert agree >= LABEL_AGREE, agree
    assert abs(float(logs["acc"].mean()) - float(g["acc"].mean())) <= ACC_TOL
    # alpha vs the reference's own float32 result: bounded by the reference's fp32 noise — 2e-4 after a few outer iterations;
    # at the default 20 / 10 outer iterations the reference's float32 is itself 2-3e-4 away from its float64 run on the
    # diverging singleton clusters, and two float32 trajectories differ by about twice that
    tol = 2e-4 if iters <= 6 else 8e-4
    for t in range(g["alpha"].shape[0]):
        assert _rel(m.alpha[t].cpu(), torch.from_numpy(g["alpha"][t])) < tol
    np.testing.assert_allclose(m.v.cpu().numpy(), g["v"], rtol=1e-3, atol=2e-2)   # v of near-empty clusters = log of ~1e-14 sums
    if not (setting == "few_shot" and method.startswith("HARD")):
        np.testing.assert_allclose(logs["criterions"], g["criterions"], rtol=2e-3, atol=1e-6)
    # soft responsibilities: compare where they are not saturated
    np.testing.assert_allclose(m.u.cpu().numpy(), g["u"], atol=2e-3)


# ------------------------------------------------------------------------------------------------------------------
# seeded synthetic batches vs the restated oracle (fp32 and fp64)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,T,iters,hard,mode,seed", [
    (20, 5, 6, False, "dense", 0),
    (20, 5, 6, True, "skip_dead", 1),
    (37, 3, 4, False, "skip_dead", 2),       # D not a multiple of 32: partially filled register slot
    (100, 8, 8, False, "dense", 3),
    (100, 8, 8, False, "skip_dead", 3),
    (100, 8, 10, True, "skip_dead", 4),
    (129, 2, 3, True, "dense", 5),
])
def test_zero_shot_vs_oracle(dev, K, T, iters, hard, mode, seed):
    from tclip_b200 import tasks
    cls = _classes()[("zero_shot", "HARD_EM_DIRICHLET" if hard else "EM_DIRICHLET")]
    td, _ = tasks.make_zero_shot_batch(T, K, seed=seed)
    m = cls(model=None, device=dev, log_file=None, args=make_args(K, iters=iters, mm_mode=mode))
    logs = m.run_task({k: v.clone() for k, v in td.items()})
    r32 = R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=hard)
    r64 = R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=hard, dtype=torch.float64)

    assert m.mm_iters.cpu().tolist() == r32.mm_iters
    assert m.n_live.cpu().tolist() == r32.n_live
    assert (m.labels.cpu().long() == r32.preds).float().mean().item() >= LABEL_AGREE
    assert abs(float(logs["acc"].mean()) - float(r32.acc.mean())) <= ACC_TOL
    a = m.alpha.cpu()
    for t in range(T):
        gpu_err = _rel(a[t], r64.alpha[t])
        ref_err = _rel(r32.alpha[t], r64.alpha[t])
        assert gpu_err <= max(ALPHA_REL, 2.0 * ref_err), (t, gpu_err, ref_err)
    assert np.isfinite(logs["criterions"]).all() and np.isfinite(logs["timestamps"])


# ------------------------------------------------------------------------------------------------------------------
# the kernel paths the BASELINE configs really take (the skip-dead schedule picks its kernels from the number of live
# rows: <= 1480 mm_spec_kernel, <= 4096 row-wise E-step + mm_chunk_kernel over a row list, above that the dense E-step)
# ------------------------------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def _oracle_zero_shot(K, T, iters, hard, seed, double, k_eff_range=(3, 10)):
    from tclip_b200 import tasks
    td, _ = tasks.make_zero_shot_batch(T, K, seed=seed, k_eff_range=k_eff_range)
    return R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=hard,
                                 dtype=torch.float64 if double else torch.float32)


def _run_vs_oracle(dev, K, T, iters, hard, seed, mode, fp64, k_eff_range=(3, 10)):
    from tclip_b200 import tasks
    cls = _classes()[("zero_shot", "HARD_EM_DIRICHLET" if hard else "EM_DIRICHLET")]
    td, _ = tasks.make_zero_shot_batch(T, K, seed=seed, k_eff_range=k_eff_range)
    m = cls(model=None, device=dev, log_file=None, args=make_args(K, iters=iters, mm_mode=mode))
    logs = m.run_task({k: v.clone() for k, v in td.items()})
    r32 = _oracle_zero_shot(K, T, iters, hard, seed, False, k_eff_range)
    assert m.mm_iters.cpu().tolist() == r32.mm_iters
    assert m.n_live.cpu().tolist() == r32.n_live
    assert (m.labels.cpu().long() == r32.preds).float().mean().item() >= LABEL_AGREE
    assert abs(float(logs["acc"].mean()) - float(r32.acc.mean())) <= ACC_TOL
    np.testing.assert_allclose(logs["criterions"], r32.criterions, rtol=5e-3, atol=1e-6)
    a = m.alpha.cpu()
    if fp64:
        r64 = _oracle_zero_shot(K, T, iters, hard, seed, True, k_eff_range)
        for t in range(T):
            gpu_err, ref_err = _rel(a[t], r64.alpha[t]), _rel(r32.alpha[t], r64.alpha[t])
            assert gpu_err <= max(ALPHA_REL, 2.0 * ref_err), (t, gpu_err, ref_err)
    else:   # vs the reference-equivalent float32 oracle: bounded by float32 noise after 3 outer iterations
        assert max(_rel(a[t], r32.alpha[t]) for t in range(T)) < 2e-4
    return m, r32


@pytest.mark.parametrize("mode", ["skip_dead", "dense"])
def test_config1_shape_vs_oracle(dev, mode):
    """BASELINE config 1: EM-Dirichlet, K = D = 100, 100 tasks per run_task batch (10 000 rows).  Soft responsibilities keep
    every cluster alive through the first E-step, so outer iteration 1 iterates all 10 000 rows through the row-list form of
    mm_chunk_kernel with the dense E-step (> 4096 live rows), outer iteration 2 ~1000 rows through mm_spec_kernel."""
    m, r32 = _run_vs_oracle(dev, 100, 100, 3, False, 11, mode, fp64=True)
    assert r32.n_live[1] > 4096 and r32.n_live[2] <= 1480, r32.n_live       # the paths this test is here for


@pytest.mark.parametrize("T,lo,hi", [(100, 1480, 4096), (200, 4096, 10 ** 9)])
def test_live_row_count_switches_vs_oracle(dev, T, lo, hi):
    """Hard EM-Dirichlet at K = D = 100 on tasks drawn from 20-30 classes: ~22 clusters per task survive the first E-step, so
    100 tasks leave 1480 < rows <= 4096 live (row-wise E-step + mm_chunk_kernel over the row list with cached dead-row
    sums) and 200 tasks leave more than 4096 (dense E-step fallback next to dead rows)."""
    out = {}
    for mode in ("skip_dead", "dense"):
        out[mode], r32 = _run_vs_oracle(dev, 100, T, 3, True, 12, mode, fp64=False, k_eff_range=(20, 30))
    assert lo < r32.n_live[1] <= hi, r32.n_live
    assert torch.equal(out["dense"].labels, out["skip_dead"].labels)
    assert _rel(out["skip_dead"].alpha, out["dense"].alpha) < 1e-5


def test_imagenet_shape_16_tasks_vs_frozen_oracle(dev, golden_dir):
    """K = D = 1000 with 16 tasks per batch (the exit test of the MM loop is global over the batch, so iteration counts and
    the kernels picked depend on the batch size); frozen oracle answers, oracle/make_k1000_t16_fixture.py."""
    from tclip_b200 import tasks
    from tclip_b200.methods.dirichlet import EM_DIRICHLET
    path = os.path.join(golden_dir, "oracle_k1000_t16.npz")
    if not os.path.isfile(path):
        pytest.skip("fixture not generated")
    g = np.load(path, allow_pickle=True)
    K, T, iters = int(g["K"]), int(g["T"]), int(g["iters"])
    td, _ = tasks.make_zero_shot_batch(T, K, seed=int(g["seed"]), batch_index=int(g["batch_index"]))
    assert np.array_equal(td["y_q"].numpy(), g["y_q"])
    if abs(R.weighted_checksum(td["x_q"]) - float(g["checksum_x_q"])) > 1e-9 * td["x_q"].numel():
        pytest.skip("the synthetic generator is not bit-reproducible on this host")
    rows64, rows32 = torch.from_numpy(g["live_rows64"]), torch.from_numpy(g["live_rows32"])
    live = torch.from_numpy(g["live"])
    ref_err = ((rows32.double() - rows64).norm() / rows64.norm()).item()
    for mode in ("skip_dead", "dense"):
        m = EM_DIRICHLET(model=None, device=dev, log_file=None, args=make_args(K, iters=iters, mm_mode=mode))
        logs = m.run_task({k: v.clone() for k, v in td.items()})
        assert m.mm_iters.cpu().tolist() == g["mm_iters32"].tolist()
        assert m.n_live.cpu().tolist() == g["n_live32"].tolist()
        assert (m.labels.cpu().numpy() == g["preds32"]).mean() >= LABEL_AGREE
        assert abs(float(logs["acc"].mean()) - float(g["acc32"].mean())) <= ACC_TOL
        np.testing.assert_allclose(logs["criterions"], g["criterions32"], rtol=5e-3, atol=1e-6)
        a = m.alpha.cpu()
        gpu_err = ((a[live].double() - rows64).norm() / rows64.norm()).item()
        assert gpu_err <= max(ALPHA_REL, 2.0 * ref_err), (mode, gpu_err, ref_err)
        norm_err = (a.double().norm(dim=2) - torch.from_numpy(g["row_norm64"])).abs() / torch.from_numpy(g["row_norm64"])
        assert norm_err.max().item() <= max(2e-4, 2.0 * float(np.max(g["task_err32"]))), norm_err.max().item()


@pytest.mark.parametrize("iter_mm", [1, 2, 49, 50, 51, 52, 101, 230])
@pytest.mark.parametrize("mode", ["dense", "skip_dead"])
def test_iter_mm_boundaries_vs_oracle(dev, iter_mm, mode):
    """M-steps shorter than / ending exactly at / just past a check point (the exit test is `l > 0 and l % 50 == 0`,
    em_dirichlet.py:169): no check, a check on the last iteration, a one-iteration tail chunk."""
    from tclip_b200 import tasks
    from tclip_b200.methods.dirichlet import EM_DIRICHLET
    K, T, iters = 24, 3, 4
    td, _ = tasks.make_zero_shot_batch(T, K, seed=40 + iter_mm)
    m = EM_DIRICHLET(model=None, device=dev, log_file=None, args=make_args(K, iters=iters, iter_mm=iter_mm, mm_mode=mode))
    logs = m.run_task({k: v.clone() for k, v in td.items()})
    r32 = R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, iter_mm=iter_mm)
    r64 = R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, iter_mm=iter_mm, dtype=torch.float64)
    assert m.mm_iters.cpu().tolist() == r32.mm_iters
    assert m.n_live.cpu().tolist() == r32.n_live
    assert (m.labels.cpu().long() == r32.preds).float().mean().item() >= LABEL_AGREE
    assert abs(float(logs["acc"].mean()) - float(r32.acc.mean())) <= ACC_TOL
    for t in range(T):
        assert _rel(m.alpha.cpu()[t], r64.alpha[t]) <= max(ALPHA_REL, 2.0 * _rel(r32.alpha[t], r64.alpha[t]))


@pytest.mark.parametrize("check_every,iter_mm", [(8, 120), (10, 95), (25, 130), (7, 60), (50, 1000), (0, 40)])
def test_check_spacing_schedules_agree(dev, check_every, iter_mm):
    """The C ABI takes the check spacing as a parameter (50 upstream).  The periodic-extension bookkeeping of the
    skip-dead schedule is only valid for spacings of 2 mod 6 and must switch itself off otherwise: both schedules have to
    agree for any spacing (MM iteration counts, live clusters, labels, alpha)."""
    from tclip_b200 import ops, tasks
    from tclip_b200._lib import TCLIP_MM_DENSE, TCLIP_MM_SKIP_DEAD
    K, T, iters = 30, 4, 5
    td, _ = tasks.make_zero_shot_batch(T, K, seed=77)
    xq = td["x_q"].to(dev)
    out = {}
    for name, mode in (("dense", TCLIP_MM_DENSE), ("skip", TCLIP_MM_SKIP_DEAD)):
        out[name] = ops.dirichlet_em(xq, K, iters=iters, iter_mm=iter_mm, lambd=float(int(K / 5) * 75), hard=False,
                                     check_every=check_every, mm_mode=mode)
    d, s_ = out["dense"], out["skip"]
    assert d["mm_iters"].cpu().tolist() == s_["mm_iters"].cpu().tolist()
    assert d["n_live"].cpu().tolist() == s_["n_live"].cpu().tolist()
    assert (d["labels"] == s_["labels"]).all()
    # same arithmetic per element, but the few-rows kernel keeps one psi(s) anchor per M-step and the chunked kernel one per
    # chunk (~1 ulp of psi(s)): 4e-9 after one outer iteration, amplified by the diverging singleton rows to ~1e-5 after five
    assert _rel(s_["alpha"], d["alpha"]) < 3e-5
    np.testing.assert_allclose(s_["mm_crit"].cpu().numpy()[:, 1], d["mm_crit"].cpu().numpy()[:, 1], rtol=1e-4)


@pytest.mark.parametrize("K,T,shots,iters,hard,seed", [
    (20, 3, 2, 4, False, 0),
    (20, 3, 2, 4, True, 1),
    (100, 4, 4, 6, False, 2),
    (100, 4, 4, 5, True, 3),
])
def test_few_shot_vs_oracle(dev, K, T, shots, iters, hard, seed):
    from tclip_b200 import tasks
    cls = _classes()[("few_shot", "HARD_EM_DIRICHLET" if hard else "EM_DIRICHLET")]
    td, _ = tasks.make_few_shot_batch(T, K, shots=shots, seed=seed)
    m = cls(model=None, device=dev, log_file=None, args=make_args(K, iters=iters, k_eff=5))
    logs = m.run_task({k: v.clone() for k, v in td.items()}, shot=shots)
    r32 = R.dirichlet_few_shot(td["x_s"], td["y_s"], td["x_q"], td["y_q"], K, 5, iters=iters, hard=hard)
    r64 = R.dirichlet_few_shot(td["x_s"], td["y_s"], td["x_q"], td["y_q"], K, 5, iters=iters, hard=hard,
                               dtype=torch.float64)
    assert m.mm_iters.cpu().tolist() == r32.mm_iters
    assert (m.labels.cpu().long() == r32.preds).float().mean().item() >= LABEL_AGREE
    assert abs(float(logs["acc"].mean()) - float(r32.acc.mean())) <= ACC_TOL
    for t in range(T):
        gpu_err = _rel(m.alpha[t].cpu(), r64.alpha[t])
        ref_err = _rel(r32.alpha[t], r64.alpha[t])
        assert gpu_err <= max(ALPHA_REL, 2.0 * ref_err), (t, gpu_err, ref_err)
    if hard:
        assert (logs["criterions"] == 0).all()       # few_shot/hard_em_dirichlet.py:234-244


# ------------------------------------------------------------------------------------------------------------------
# single stages through the C ABI vs the same stage of the oracle
# ------------------------------------------------------------------------------------------------------------------
def test_stage_mm_step(dev):
    """tclip_dirichlet_mm vs oracle.mm_update_alpha: iteration count and alpha, incl. tiny and huge alpha rows."""
    from tclip_b200 import ops
    g = torch.Generator().manual_seed(5)
    rows, D = 64, 100
    z = torch.softmax(3 * torch.randn(rows, D, generator=g), -1)
    y = torch.log(z + 1e-15)
    y[5] = -10.0                                             # an empty-cluster row
    a0 = torch.ones(rows, D)
    a0[7] = torch.rand(D, generator=g) * 1e-3                # small-alpha branch of the curvature
    a0[9] = 1e4 * (1 + torch.rand(D, generator=g))           # large alpha
    for iter_mm in (1, 49, 51, 120, 1000):
        out, iters = ops.mm_update_alpha(a0.to(dev), y.to(dev), iter_mm=iter_mm)
        ref, done = R.mm_update_alpha(a0.double(), y.double(), iter_mm)
        assert int(iters.item()) == done, (iter_mm, int(iters.item()), done)
        err = ((out.cpu().double() - ref).abs() / ref.abs()).max().item()
        assert err < 2e-4, (iter_mm, err)


def test_stage_estep_moments(dev):
    from tclip_b200 import ops
    g = torch.Generator().manual_seed(9)
    T, n, K = 3, 75, 50
    z = torch.softmax(4 * torch.randn(T, n, K, generator=g), -1)
    logz_ref = torch.log(z + 1e-15)
    logz = ops.log_features(z.to(dev))
    np.testing.assert_allclose(logz.cpu().numpy(), logz_ref.numpy(), rtol=2e-6, atol=1e-6)
    u = torch.softmax(6 * torch.randn(T, n, K, generator=g), -1)
    u[:, :, 3] = 0.0                                         # an empty cluster
    colsum, v, live = ops.colsum_v(u.to(dev))
    np.testing.assert_allclose(colsum.cpu().numpy(), u.sum(1).numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(v.cpu().numpy(), (torch.log(u.sum(1) / n + 1e-15) + 1).numpy(), rtol=1e-5, atol=1e-5)
    assert (live.cpu().numpy() == (u.sum(1) > 1e-15).numpy()).all()
    y = ops.moments(u.to(dev), logz, colsum)
    y_ref = torch.einsum("tnk,tnd->tkd", u.double(), logz_ref.double()) / u.double().sum(1).clamp(min=1e-15)[..., None]
    y_ref[:, 3, :] = -10.0
    np.testing.assert_allclose(y.cpu().numpy(), y_ref.numpy(), rtol=1e-5, atol=1e-5)
    alpha = 1 + 5 * torch.rand(T, K, K, generator=g)
    vv = torch.randn(T, K, generator=g)
    lambd = int(K / 5) * n
    logits = R.dirichlet_logits(alpha.double(), logz_ref.double(), "einsum") + lambd * vv.double().unsqueeze(1) / n
    for hard in (False, True):
        uu, labels = ops.estep(alpha.to(dev), logz, vv.to(dev), float(lambd), hard)
        sm = logits.softmax(2)
        assert (labels.cpu().long() == sm.argmax(2)).float().mean().item() >= LABEL_AGREE
        if hard:
            assert torch.equal(uu.cpu(), R.one_hot_argmax(uu.cpu()))
            assert (uu.sum(2) == 1).all()
        else:
            np.testing.assert_allclose(uu.cpu().numpy(), sm.numpy(), atol=2e-4)


@pytest.mark.parametrize("T,n,K,D,few", [(2, 75, 20, 20, False), (3, 75, 100, 100, True), (2, 75, 1000, 1000, False),
                                         (2, 75, 1000, 1000, True), (2, 40, 136, 100, False), (1, 13, 7, 33, True)])
def test_stage_moments_tensor_cores(dev, T, n, K, D, few):
    """The moments u^T log z on tcgen05 (3 x TF32, fp32 round-to-nearest running sum; u^T and (log z)^T staged as K-major
    operands, division / support terms / -10 fill in the epilogue) against float64 and against the CUDA-core kernel."""
    from tclip_b200 import ops
    g = torch.Generator().manual_seed(100 + K + n)
    z = torch.softmax(5 * torch.randn(T, n, D, generator=g), -1)
    logz = ops.log_features(z.to(dev))
    u = torch.softmax(6 * torch.randn(T, n, K, generator=g), -1)
    u[:, :, min(3, K - 1)] = 0.0                               # an empty cluster
    colsum, _, _ = ops.colsum_v(u.to(dev))
    ssum = scount = None
    if few:
        ssum = (-5.0 * torch.rand(T, K, D, generator=g) * 4).to(dev)
        scount = torch.full((T, K), 4.0).to(dev)
    y_tc = ops.moments(u.to(dev), logz, colsum, ssum, scount, tensor_cores=True).cpu().double()
    y_cc = ops.moments(u.to(dev), logz, colsum, ssum, scount).cpu().double()
    acc = torch.einsum("tnk,tnd->tkd", u.double(), logz.cpu().double())
    if few:
        y_ref = (ssum.cpu().double() + acc) / (scount.cpu().double() + u.double().sum(1))[..., None]
    else:
        y_ref = acc / u.double().sum(1).clamp(min=1e-15)[..., None]
        y_ref[:, min(3, K - 1), :] = -10.0
        assert (y_tc[:, min(3, K - 1), :] == -10.0).all()
    err_tc = ((y_tc - y_ref).abs() / y_ref.abs().clamp(min=1e-3)).max().item()
    err_cc = ((y_cc - y_ref).abs() / y_ref.abs().clamp(min=1e-3)).max().item()
    # measured: rms 1.3-2.2e-7 relative against 1.0-1.4e-7 for the CUDA-core kernel (inside a 32-long block the tensor core adds
    # with truncation); what counts is the alpha it leads to, which the parity tests bound by the reference's own float32 error
    # (profiles/r2_moments_tc.md: no measurable difference)
    assert err_tc <= max(3.0 * err_cc, 1e-6), (err_tc, err_cc)
    rms_tc = ((y_tc - y_ref) / y_ref.abs().clamp(min=1e-3)).pow(2).mean().sqrt().item()
    rms_cc = ((y_cc - y_ref) / y_ref.abs().clamp(min=1e-3)).pow(2).mean().sqrt().item()
    assert rms_tc <= max(2.0 * rms_cc, 3e-7), (rms_tc, rms_cc)


@pytest.mark.parametrize("T,n,K,D", [(2, 75, 20, 20), (3, 75, 100, 100), (2, 75, 136, 100), (2, 40, 1000, 1000),
                                     (1, 128, 260, 64), (2, 75, 7, 1024)])
def test_stage_contraction_tensor_cores(dev, T, n, K, D):
    """tcgen05 3 x TF32 contraction (the E-step's GEMM) vs float64: as accurate as the CUDA-core fp32 kernel, on ragged
    tiles too (K not a multiple of 128, D not a multiple of 32, n < 128); heavy-tailed alpha like a real EM state."""
    from tclip_b200 import ops
    g = torch.Generator().manual_seed(100 + K)
    logz = torch.log(torch.softmax(5 * torch.randn(T, n, D, generator=g), -1) + 1e-15)
    alpha = (0.03 + torch.exp(2.5 * torch.randn(T, K, D, generator=g))).clamp(max=3e5)
    ref = torch.einsum("tnd,tkd->tnk", logz.double(), (alpha - 1.0).double())
    scale = ref.abs().max().item()
    out = {m: ops.contraction(logz.to(dev), alpha.to(dev), m).cpu().double() for m in ("tcgen05", "tcgen05_tmem_sum", "simt")}
    rms = {m: (o - ref).pow(2).mean().sqrt().item() for m, o in out.items()}
    mx = {m: (o - ref).abs().max().item() for m, o in out.items()}
    assert mx["simt"] <= 2e-6 * scale, (mx, scale)
    assert mx["tcgen05"] <= 2e-6 * scale, (mx, scale)
    assert rms["tcgen05"] <= 2.0 * rms["simt"] + 1e-9 * scale, (rms, scale)
    # the textbook loop (whole sum in the TMEM accumulator) is what the external fp32 sum protects against
    assert mx["tcgen05_tmem_sum"] <= 1e-4 * scale, (mx, scale)


def test_contraction_rejects_shapes_tma_cannot_address(dev):
    from tclip_b200 import ops
    from tclip_b200._lib import TclipError
    logz = torch.zeros(1, 75, 50, device=dev)
    alpha = torch.ones(1, 50, 50, device=dev)
    with pytest.raises(TclipError):
        ops.contraction(logz, alpha, "tcgen05")          # D % 4 != 0
    assert torch.equal(ops.contraction(logz, alpha, "simt"), torch.zeros(1, 75, 50, device=dev))


def test_stage_cluster_prototypes_and_matching(dev):
    import scipy_matching as matching
    from tclip_b200 import ops
    g = torch.Generator().manual_seed(2)
    T, n, K = 4, 75, 30
    feats = torch.softmax(3 * torch.randn(T, n, K, generator=g), -1)
    labels = torch.randint(0, 7, (T, n), generator=g) * 3
    cl = ops.cluster_prototypes(labels.int().to(dev), feats.to(dev))
    onehot = torch.nn.functional.one_hot(labels, K).float()
    _, protos = R.cluster_prototypes(onehot, feats, K, "einsum")
    want = R.graph_matching(labels, protos, K)
    got = matching.graph_matching(cl["proto"].cpu().numpy(), cl["n_clusters"].cpu().numpy(),
                                  cl["sample_cluster"].cpu().numpy())
    assert (got == want.numpy()).all()
    want_b = R.basic_matching(labels, protos)
    got_b = matching.basic_matching(cl["proto"].cpu().numpy(), cl["n_clusters"].cpu().numpy(),
                                    cl["sample_cluster"].cpu().numpy())
    assert (got_b == want_b.numpy()).all()
    # the device kernel (what the method classes call) against both
    y = torch.randint(0, K, labels.shape, generator=g).to(dev)
    dm = ops.match_clusters(cl["proto"], cl["n_clusters"], cl["sample_cluster"], y, graph_matching=True)
    assert (dm["new_labels"].cpu().numpy() == want.numpy()).all()
    np.testing.assert_array_equal(dm["acc"].cpu().numpy(), (want.to(dev) == y).float().mean(1).cpu().numpy())
    db = ops.match_clusters(cl["proto"], cl["n_clusters"], cl["sample_cluster"], y, graph_matching=False)
    assert (db["new_labels"].cpu().numpy() == want_b.numpy()).all()


@pytest.mark.parametrize("T,n,K,seed", [(6, 75, 1000, 0), (8, 75, 100, 1), (5, 75, 20, 2), (3, 40, 64, 3), (4, 128, 2000, 4)])
def test_device_assignment_equals_scipy(dev, T, n, K, seed):
    """tclip_match_clusters vs scipy.optimize.linear_sum_assignment on rectangular cost matrices with as many clusters as
    a task can have (every query its own cluster when K >= n): many augmenting steps, long alternating paths."""
    from scipy.optimize import linear_sum_assignment
    from tclip_b200 import ops
    g = torch.Generator().manual_seed(seed)
    probs = torch.softmax(3 * torch.randn(T, n, K, generator=g), -1)           # concentrated rows -> contested classes
    hot = torch.randint(0, max(K // 8, 2), (T, n), generator=g)                # many clusters want the same few classes
    probs = (probs + 2.0 * torch.nn.functional.one_hot(hot, K).float()) / 3.0
    n_clusters = torch.tensor([min(n, K) - (i % 3) for i in range(T)], dtype=torch.int32)
    sample_cluster = torch.stack([torch.randint(0, int(c), (n,), generator=g) for c in n_clusters]).int()
    out = ops.match_clusters(probs.to(dev), n_clusters.to(dev), sample_cluster.to(dev), None, graph_matching=True)
    cc = out["cluster_class"].cpu().numpy()
    for t in range(T):
        c = int(n_clusters[t])
        cost = -probs[t, :c].double().numpy()
        rows, cols = linear_sum_assignment(cost)
        assert (cc[t, :c] == cols).all(), (t, cost[rows, cols].sum(), cost[np.arange(c), cc[t, :c]].sum())
        assert (cc[t, c:] == -1).all()
        assert (out["new_labels"][t].cpu().numpy() == cols[sample_cluster[t].numpy()]).all()


@pytest.mark.parametrize("K,T,iters,hard", [(100, 6, 6, False), (100, 5, 5, True), (1000, 3, 4, False)])
def test_mm_exit_norms_agree_between_schedules(dev, K, T, iters, hard):
    """The two norms of the batch-global exit test (em_dirichlet.py:170-171) at the last check point of every M-step.
    The dense schedule iterates every row; skip-dead takes the empty clusters' terms from their cached / periodically
    extended trajectories and the live ones from the one-launch speculative kernel.  Same tensors, so the sums must agree
    (up to what the slightly different E-step roundings of the two schedules leave in the live rows)."""
    from tclip_b200 import tasks
    from tclip_b200.methods.dirichlet import EM_DIRICHLET, HARD_EM_DIRICHLET
    td, _ = tasks.make_zero_shot_batch(T, K, seed=11)
    crit, its = {}, {}
    for mode in ("dense", "skip_dead"):
        m = (HARD_EM_DIRICHLET if hard else EM_DIRICHLET)(model=None, device=dev, log_file=None,
                                                          args=make_args(K, iters=iters, mm_mode=mode))
        m.run_task({k: v.clone() for k, v in td.items()})
        crit[mode], its[mode] = m.mm_crit.cpu().numpy(), m.mm_iters.cpu().tolist()
    assert its["dense"] == its["skip_dead"]
    assert (crit["dense"] > 0).all()
    # ||alpha||^2: dominated by the empty clusters (cached terms) and a few diverging singleton clusters, which the two
    # schedules track to ~1e-5 (their E-steps round differently)
    np.testing.assert_allclose(crit["skip_dead"][:, 1], crit["dense"][:, 1], rtol=1e-4)
    # ||alpha_new - alpha||^2: where the M-step has converged to a few ulp (outer iteration 0) the value is rounding
    # noise of the row-total summation order, which differs between the kernels; elsewhere ~1e-4
    np.testing.assert_allclose(crit["skip_dead"][:, 0], crit["dense"][:, 0], rtol=0.3)
    np.testing.assert_allclose(crit["skip_dead"][1:, 0], crit["dense"][1:, 0], rtol=5e-3)


def test_device_task_construction(dev):
    """tasks.DeviceTaskSource: gathering the sampler's indices on the device gives the tensors the evaluator builds on the
    host (src/eval_zero_shot.py:158-168), ragged feature widths included, and run_task takes them as they are."""
    import random
    from tclip_b200 import ops, tasks
    from tclip_b200.methods.dirichlet import EM_DIRICHLET
    g = torch.Generator().manual_seed(8)
    for N, F in ((3000, 100), (500, 37), (64, 1024)):
        feats = torch.softmax(3 * torch.randn(N, F, generator=g), -1)
        labels = torch.randint(0, 20, (N,), generator=g)
        idx = torch.randint(0, N, (5, 75), generator=g)
        x, y, bad = ops.gather_tasks(feats.to(dev), labels.to(dev), idx.to(dev))
        assert int(bad.item()) == 0
        assert torch.equal(x.cpu(), feats[idx]) and torch.equal(y.cpu(), labels[idx])
    idx[0, 0] = N                                                       # out of range: counted, zero row
    x, y, bad = ops.gather_tasks(feats.to(dev), labels.to(dev), idx.to(dev))
    assert int(bad.item()) == 1 and float(x[0, 0].abs().sum()) == 0.0 and int(y[0, 0]) == -1
    # evaluator-level: same sampler seeds -> same tasks -> same logs as the host-built task_dic
    K, T = 40, 4
    feats = torch.softmax(4 * torch.randn(6000, K, generator=g), -1)
    labels = torch.randint(0, K, (6000,), generator=g)
    src = tasks.DeviceTaskSource(feats, labels, dev)
    def sampler():
        random.seed(1)
        torch.manual_seed(2)
        return tasks.ZeroShotQuerySampler(T, K, 75, labels)
    on_dev = src.generate_tasks(sampler())
    host_idx = torch.stack(list(sampler()))
    on_host = {"x_q": feats[host_idx], "y_q": labels[host_idx].unsqueeze(-1)}
    assert torch.equal(on_dev["x_q"].cpu(), on_host["x_q"]) and torch.equal(on_dev["y_q"].cpu(), on_host["y_q"])
    logs = []
    for td in (on_dev, on_host):
        m = EM_DIRICHLET(model=None, device=dev, log_file=None, args=make_args(K, iters=4))
        logs.append(m.run_task(dict(td)))
    np.testing.assert_array_equal(logs[0]["acc"], logs[1]["acc"])
    np.testing.assert_array_equal(logs[0]["criterions"], logs[1]["criterions"])


def test_device_few_shot_task_construction(dev):
    """tasks.DeviceFewShotTaskSource against the host construction of tests/host_twin.py (which the CPU suite checks against
    the reference's own Tasks_Generator_few_shot)."""
    import random
    from tclip_b200 import tasks
    g = torch.Generator().manual_seed(21)
    n_class, n_query, T, shots, k_eff = 16, 75, 3, 2, 5
    fs = torch.softmax(2 * torch.randn(400, n_class, generator=g), -1)
    ls = torch.cat([torch.arange(n_class).repeat(5), torch.randint(0, n_class, (320,), generator=g)])
    fq = torch.softmax(2 * torch.randn(3000, n_class, generator=g), -1)
    lq = torch.randint(0, n_class, (3000,), generator=g)
    def samplers():
        random.seed(3)
        torch.manual_seed(4)
        return tasks.FewShotSamplers(T, k_eff, n_class, shots, n_query, ls, lq)
    smp = samplers()
    idx_q = list(smp.query())
    idx_s = list(smp.support())
    from host_twin import few_shot_tasks_on_host
    want = few_shot_tasks_on_host(fs, ls, fq, lq, idx_s, idx_q)   # checked against the reference in test_host_logic.py
    src = tasks.DeviceFewShotTaskSource(fs, ls, fq, lq, dev)
    smp = samplers()
    q_it = list(smp.query())
    got = src.generate_tasks(smp.support(), q_it)
    for k in want:
        assert torch.equal(got[k].cpu().reshape(want[k].shape), want[k]), k
    assert got["x_s"].shape == (T, n_class * shots, n_class) and got["y_q"].shape == (T, n_query, 1)


def test_batches_in_flight_equal_serial(dev):
    """tclip_b200.pipeline: whole run_task batches on three CUDA streams / host threads give, batch by batch, exactly what
    the same calls give one after the other (own scratch per stream, no shared state between batches)."""
    from tclip_b200 import tasks
    from tclip_b200.methods.dirichlet import EM_DIRICHLET, HARD_EM_DIRICHLET
    from tclip_b200.pipeline import BatchPipeline
    K, T, iters = 100, 8, 5
    batches = [tasks.make_zero_shot_batch(T, K, seed=5, batch_index=i)[0] for i in range(7)]

    def run(i):
        cls = HARD_EM_DIRICHLET if i % 2 else EM_DIRICHLET
        m = cls(model=None, device=dev, log_file=None, args=make_args(K, iters=iters))
        logs = m.run_task({k: v.clone() for k, v in batches[i].items()})
        return logs["acc"], logs["criterions"], m.alpha.cpu(), m.mm_iters.cpu()

    import sys
    serial = [run(i) for i in range(len(batches))]
    before = sys.getswitchinterval()
    with BatchPipeline(dev, streams=3) as pipe:
        assert sys.getswitchinterval() <= 5e-4        # workers hand the interpreter over quickly while the pipeline is open
        piped = pipe.map(run, range(len(batches)))
    assert sys.getswitchinterval() == before          # and the setting is restored
    for a, b in zip(serial, piped):
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
        assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])


@pytest.mark.parametrize("K,T,iters,hard,k_eff,noise", [
    (1000, 6, 7, False, (3, 10), 9.0), (1000, 6, 6, True, (3, 10), 9.0), (1000, 4, 5, False, (20, 30), 6.0),
    (100, 8, 8, False, (3, 10), 9.0), (100, 8, 8, True, (3, 10), 9.0), (50, 5, 6, False, (3, 10), 3.0),
])
def test_sparse_softmax_is_bit_identical(dev, K, T, iters, hard, k_eff, noise):
    """The sparse regime's soft-max visits only the live classes of a query once a bound proves that every dead class
    underflows to exactly +0.0f (estep_task_kernel); TCLIP_FLAG_FULL_SOFTMAX keeps the pass over all K classes.  Both must give
    the same u, labels, v and alpha to the last bit, through changes of the set of dead classes, in the soft and the hard
    variant (which only moves the 1 of a one-hot row)."""
    from tclip_b200 import ops, tasks
    td, _ = tasks.make_zero_shot_batch(T, K, seed=31, batch_index=K + iters, k_eff_range=k_eff, noise=noise)
    xq = td["x_q"].to(dev)
    out = [ops.dirichlet_em(xq, K, iters=iters, iter_mm=1000, lambd=float(int(K / 5) * 75), hard=hard,
                            mm_mode=ops.TCLIP_MM_SKIP_DEAD, full_softmax=flag) for flag in (True, False)]
    assert min(out[0]["n_live"].cpu().tolist()[1:]) <= 4096          # the row-wise E-step ran
    for key in ("u", "labels", "v", "alpha", "mm_iters", "n_live", "criterions"):
        assert torch.equal(out[0][key], out[1][key]), key
    if hard:
        u = out[1]["u"]
        assert ((u == 0) | (u == 1)).all() and (u.sum(2) == 1).all()
        assert torch.equal(u.argmax(2).int(), out[1]["labels"])


def test_in_flight_tail_kernel_is_bit_identical(dev):
    """TCLIP_FLAG_IN_FLIGHT (set by the method classes inside a BatchPipeline worker) swaps the few-rows M-step kernel of the
    skip-dead schedule for its register-lean form at D > 768: plain update instead of the two-phase one, which is the same
    arithmetic operation for operation (tests/test_math_host.py) — the results must be equal to the last bit."""
    from tclip_b200 import ops, tasks
    K, T, iters = 1000, 5, 4
    td, _ = tasks.make_zero_shot_batch(T, K, seed=2020, batch_index=3)
    xq = td["x_q"].to(dev)
    out = [ops.dirichlet_em(xq, K, iters=iters, iter_mm=1000, lambd=float(int(K / 5) * 75), hard=False,
                            mm_mode=ops.TCLIP_MM_SKIP_DEAD, in_flight=flag) for flag in (False, True)]
    assert max(out[0]["n_live"].cpu().tolist()[1:]) <= 1480          # the few-rows kernel ran
    assert torch.equal(out[0]["alpha"], out[1]["alpha"]) and torch.equal(out[0]["u"], out[1]["u"])
    assert torch.equal(out[0]["mm_iters"], out[1]["mm_iters"]) and torch.equal(out[0]["labels"], out[1]["labels"])


# ------------------------------------------------------------------------------------------------------------------
# BASELINE sizes (K = D = 1000, n = 75): size-independent properties
# ------------------------------------------------------------------------------------------------------------------
def test_imagenet_shape_properties(dev):
    """At ImageNet shape the CPU oracle takes minutes per task, so check what must hold at any size:
      * responsibilities are row-stochastic (hard: exactly one-hot);
      * the skip-dead schedule reproduces the dense one: same MM iteration counts, same labels, same alpha;
      * live rows whose MM converged satisfy the Dirichlet MLE stationarity psi(a_d) - psi(sum a) = y_d;
      * rows of empty clusters keep alpha == 1 (their MM result is discarded, em_dirichlet.py:224-226)."""
    from tclip_b200 import tasks
    from tclip_b200.methods.dirichlet import HARD_EM_DIRICHLET
    K, T, iters = 1000, 6, 3
    td, _ = tasks.make_zero_shot_batch(T, K, seed=2020)
    out = {}
    for mode in ("dense", "skip_dead"):
        m = HARD_EM_DIRICHLET(model=None, device=dev, log_file=None, args=make_args(K, iters=iters, mm_mode=mode))
        logs = m.run_task({k: v.clone() for k, v in td.items()})
        out[mode] = (m, logs)
    md, ms = out["dense"][0], out["skip_dead"][0]
    assert md.mm_iters.cpu().tolist() == ms.mm_iters.cpu().tolist()
    assert torch.equal(md.labels, ms.labels)
    assert _rel(ms.alpha, md.alpha) < 1e-5      # same arithmetic; the row totals are summed in a different order
    u = md.u
    assert ((u == 0) | (u == 1)).all() and (u.sum(2) == 1).all()
    assert abs(float(out["dense"][1]["acc"].mean()) - float(out["skip_dead"][1]["acc"].mean())) < 1e-6
    # empty clusters at the end keep the initial row
    sizes = torch.zeros(T, K, device=dev).scatter_add_(1, md.labels.long(), torch.ones(T, 75, device=dev))
    # a cluster that was empty at every M-step after the first has alpha from outer iteration 0 only: finite, positive
    assert torch.isfinite(md.alpha).all() and (md.alpha > 0).all()
    assert sizes.sum().item() == T * 75


# ------------------------------------------------------------------------------------------------------------------
# BASELINE shape against the frozen answers of the CPU oracle (oracle/make_k1000_fixture.py; ~35 CPU-minutes to make)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["em", "hard"])
@pytest.mark.parametrize("mode", ["skip_dead", "dense"])
def test_imagenet_shape_vs_frozen_oracle(dev, golden_dir, name, mode):
    from tclip_b200 import tasks
    path = os.path.join(golden_dir, f"oracle_k1000_{name}.npz")
    if not os.path.isfile(path):
        pytest.skip("fixture not generated")
    g = np.load(path, allow_pickle=True)
    K, T, iters, hard = int(g["K"]), int(g["T"]), int(g["iters"]), bool(g["hard"])
    # the inputs are stored, not regenerated: the generator's matmul/softmax is not bit-reproducible across host CPUs
    inp = np.load(os.path.join(golden_dir, "oracle_k1000_inputs.npz"))
    td = {"x_q": torch.from_numpy(inp["x_q"]), "y_q": torch.from_numpy(inp["y_q"])}
    assert np.array_equal(inp["y_q"], g["y_q"])
    cls = _classes()[("zero_shot", "HARD_EM_DIRICHLET" if hard else "EM_DIRICHLET")]
    m = cls(model=None, device=dev, log_file=None, args=make_args(K, iters=iters, mm_mode=mode))
    logs = m.run_task({k: v.clone() for k, v in td.items()})

    assert m.mm_iters.cpu().tolist() == g["mm_iters32"].tolist()
    assert m.n_live.cpu().tolist() == g["n_live32"].tolist()
    assert (m.labels.cpu().numpy() == g["preds32"]).mean() >= LABEL_AGREE
    assert abs(float(logs["acc"].mean()) - float(g["acc32"].mean())) <= ACC_TOL
    np.testing.assert_allclose(logs["criterions"], g["criterions32"], rtol=5e-3, atol=1e-6)
    # alpha: rows of the clusters alive at the end carry all of the mass; GPU no further from float64 than 2x the
    # reference float32 (or 1e-4), and every row norm (alive or not) within 1e-4 of float64
    a = m.alpha.cpu()
    live = torch.from_numpy(g["live"])
    rows64, rows32 = torch.from_numpy(g["live_rows64"]), torch.from_numpy(g["live_rows32"])
    gpu_rows = a[live].double()
    gpu_err = ((gpu_rows - rows64).norm() / rows64.norm()).item()
    ref_err = ((rows32.double() - rows64).norm() / rows64.norm()).item()
    assert gpu_err <= max(ALPHA_REL, 2.0 * ref_err), (gpu_err, ref_err)
    norm_err = (a.double().norm(dim=2) - torch.from_numpy(g["row_norm64"])).abs() / torch.from_numpy(g["row_norm64"])
    assert norm_err.max().item() <= max(2e-4, 2.0 * float(np.max(g["task_err32"]))), norm_err.max().item()


@pytest.mark.parametrize("tag", ["em", "hard"])
def test_few_shot_imagenet_shape_vs_frozen_oracle(dev, golden_dir, tag):
    """BASELINE config 3: 4-shot EM-Dirichlet at K = D = 1000 (S = 4000 support samples with label clamping).  The live
    reference cannot run this shape; the answers come from the restated oracle (oracle/make_k1000_fewshot_fixture.py)."""
    from tclip_b200 import tasks
    path = os.path.join(golden_dir, "oracle_k1000_fewshot.npz")
    if not os.path.isfile(path):
        pytest.skip("fixture not generated")
    g = np.load(path, allow_pickle=True)
    K, T, shots, iters = int(g["K"]), int(g["T"]), int(g["shots"]), int(g["iters"])
    td, _ = tasks.make_few_shot_batch(T, K, shots=shots, k_eff=int(g["k_eff"]), seed=int(g["seed"]),
                                      batch_index=int(g["batch_index"]))
    # the 32 MB support set is regenerated (float64 generator, bit-reproducible); make sure it is the fixture's
    assert np.array_equal(td["y_q"].numpy(), g["y_q"]) and np.array_equal(td["y_s"].numpy(), g["y_s"])
    for key in ("x_q", "x_s"):
        if abs(R.weighted_checksum(td[key]) - float(g["checksum_" + key])) > 1e-9 * td[key].numel():
            pytest.skip("the synthetic generator is not bit-reproducible on this host")
    hard = tag == "hard"
    cls = _classes()[("few_shot", "HARD_EM_DIRICHLET" if hard else "EM_DIRICHLET")]
    m = cls(model=None, device=dev, log_file=None, args=make_args(K, iters=iters, k_eff=int(g["k_eff"])))
    logs = m.run_task({k: v.clone() for k, v in td.items()}, shot=shots)
    assert m.mm_iters.cpu().tolist() == g[f"mm_iters32_{tag}"].tolist()
    assert (m.labels.cpu().numpy() == g[f"preds32_{tag}"]).mean() >= LABEL_AGREE
    assert abs(float(logs["acc"].mean()) - float(g[f"acc32_{tag}"].mean())) <= ACC_TOL
    a = m.alpha.cpu().double()
    rows = torch.arange(0, K, 37)
    ref64, ref32 = torch.from_numpy(g[f"rows64_{tag}"]), torch.from_numpy(g[f"rows32_{tag}"]).double()
    gpu_err = ((a[:, rows] - ref64).norm() / ref64.norm()).item()
    ref_err = ((ref32 - ref64).norm() / ref64.norm()).item()
    assert gpu_err <= max(ALPHA_REL, 2.0 * ref_err), (gpu_err, ref_err)
    norm_err = (a.norm(dim=2) - torch.from_numpy(g[f"row_norm64_{tag}"])).abs() / torch.from_numpy(g[f"row_norm64_{tag}"])
    assert norm_err.max().item() <= max(ALPHA_REL, 2.0 * float(np.max(g[f"task_err32_{tag}"]))), norm_err.max().item()

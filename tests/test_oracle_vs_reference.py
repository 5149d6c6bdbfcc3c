"""CPU, build container only: the restated oracle against the LIVE reference classes imported unchanged from
/root/reference (or $TCLIP_REF) on fresh seeds.  Skipped where the checkout does not exist (the GPU box)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import ref_loader, restated as R
from tclip_b200 import tasks

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")


@pytest.mark.parametrize("method,hard", [("EM_DIRICHLET", False), ("HARD_EM_DIRICHLET", True)])
def test_zero_shot_dirichlet_live(method, hard):
    K, T, iters = 16, 3, 3
    td, _ = tasks.make_zero_shot_batch(T, K, seed=123)
    logs, inst = ref_loader.run_reference(method, "zero_shot", td, ref_loader.make_args(K, iters=iters))
    r = R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=hard, contraction="broadcast")
    assert torch.equal(inst.u, r.u) and torch.equal(inst.alpha, r.alpha) and torch.equal(inst.v, r.v)
    assert np.array_equal(logs["acc"], r.acc) and np.array_equal(logs["criterions"], r.criterions)


@pytest.mark.parametrize("method,hard", [("EM_DIRICHLET", False), ("HARD_EM_DIRICHLET", True)])
def test_few_shot_dirichlet_live(method, hard):
    K, T, iters, shots = 12, 2, 3, 2
    td, _ = tasks.make_few_shot_batch(T, K, shots=shots, seed=321)
    logs, inst = ref_loader.run_reference(method, "few_shot", td, ref_loader.make_args(K, iters=iters, k_eff=5), shot=shots)
    r = R.dirichlet_few_shot(td["x_s"], td["y_s"], td["x_q"], td["y_q"], K, 5, iters=iters, hard=hard,
                             contraction="broadcast")
    assert torch.equal(inst.u, r.u) and torch.equal(inst.alpha, r.alpha)
    assert np.array_equal(logs["acc"], r.acc) and np.array_equal(logs["criterions"], r.criterions)


@pytest.mark.parametrize("method,km", [("SOFT_KMEANS", "soft"), ("HARD_KMEANS", "hard"), ("EM_GAUSSIAN", "gauss"),
                                       ("EM_GAUSSIAN_COV", "gauss_cov"), ("KL_KMEANS", "kl")])
@pytest.mark.parametrize("softmax", [True, False])
def test_kmeans_family_live(method, km, softmax):
    K, T, iters = 12, 2, 3
    td, txt = tasks.make_zero_shot_batch(T, K, seed=77, softmax_feature=softmax, embed_dim=32)
    args = ref_loader.make_args(K, iters=iters, use_softmax_feature=softmax)
    logs, inst = ref_loader.run_reference(method, "zero_shot", td, args, model=ref_loader.StubTextModel(txt))
    r = R.kmeans_family(td["x_q"], td["y_q"], K, method=km, iters=iters, use_softmax_feature=softmax, text=txt,
                        contraction="broadcast")
    assert torch.equal(inst.u, r.u) and torch.equal(inst.w, r.w)
    assert np.array_equal(logs["acc"], r.acc) and np.array_equal(logs["criterions"], r.criterions, equal_nan=True)


def test_einsum_mode_at_imagenet_shape_live():
    """The ``einsum`` contraction mode of the restatement (the one every K = 1000 fixture and the CPU baseline use) against
    the LIVE reference at K = D = 1000: one task, two outer iterations (51 + 1000 MM iterations; the second M-step and the
    second E-step see moments / logits that went through the einsum path).  The reference forms the [T,n,K,D] broadcast
    (300 MB per temporary at T = 1); einsum sums in a different order, so equality is to float32 rounding, not bitwise."""
    K, T, iters = 1000, 1, 2
    td, _ = tasks.make_zero_shot_batch(T, K, seed=2020, batch_index=31)
    logs, inst = ref_loader.run_reference("EM_DIRICHLET", "zero_shot", td, ref_loader.make_args(K, iters=iters))
    r = R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, contraction="einsum")
    assert r.mm_iters == [51, 1000]
    assert torch.equal(inst.u.argmax(2), r.preds)
    assert np.array_equal(logs["acc"], r.acc)
    rel = ((inst.alpha.double() - r.alpha.double()).norm() / inst.alpha.double().norm()).item()
    assert rel < 2e-5, rel
    live = inst.u.sum(1) > 1e-15
    assert torch.equal(live, r.u.sum(1) > 1e-15)
    np.testing.assert_allclose(logs["criterions"], r.criterions, rtol=1e-4)
    np.testing.assert_allclose(inst.u.numpy(), r.u.numpy(), atol=1e-4)

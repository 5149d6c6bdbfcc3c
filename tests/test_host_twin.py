"""CPU: the kernel's arithmetic (packed shift-4 Stirling + software ln X + row_psi, compiled as plain C++) dropped into
the oracle's EM loop must reproduce the oracle: MM iteration counts, labels, accuracy, and alpha no further from the
float64 restatement than twice the reference-float32 error."""
from __future__ import annotations

import pytest
import torch

from oracle import restated as R
from tclip_b200 import tasks

import host_twin


@pytest.mark.parametrize("K,T,iters,hard,seed", [(20, 4, 5, False, 0), (33, 2, 4, True, 1), (64, 3, 4, False, 2)])
def test_twin_em_matches_oracle(K, T, iters, hard, seed):
    td, _ = tasks.make_zero_shot_batch(T, K, seed=seed)
    r32 = R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=hard)
    r64 = R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=hard, dtype=torch.float64)
    with host_twin.patched_oracle():
        rt = R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=hard)
    assert rt.mm_iters == r32.mm_iters
    assert (rt.preds == r32.preds).float().mean().item() >= 0.999
    assert abs(float(rt.acc.mean()) - float(r32.acc.mean())) <= 1e-3
    for t in range(T):
        e_twin = ((rt.alpha[t].double() - r64.alpha[t]).norm() / r64.alpha[t].norm()).item()
        e_ref = ((r32.alpha[t].double() - r64.alpha[t]).norm() / r64.alpha[t].norm()).item()
        assert e_twin <= max(1e-4, 2 * e_ref), (t, e_twin, e_ref)

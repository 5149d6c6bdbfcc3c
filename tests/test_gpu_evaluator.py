"""GPU: the evaluator-level boundary: cached features written by ``tclip_b200.features.save_features``,
read back, and pushed through the reference evaluator's call sequence (``tests/evaluator_harness.py``, pinned to the live
``Evaluator_zero_shot`` in the build container) into the DROP-IN modules ``src.methods.zero_shot.*`` picked by
``args.name_method`` — what ``main.py`` does after feature extraction.  Accuracies are checked against the restated
oracle run on the very same task dictionaries."""
from __future__ import annotations

import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import evaluator_harness as H  # noqa: E402
from oracle import ref_loader, restated as R  # noqa: E402  (test infrastructure: the checker)
from oracle.ref_loader import make_args  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device (tclip_b200 has no CPU path)")
    ref_loader._install_clip_stub()
    return torch.device("cuda:0")


def _cached_features(tmp_path, K, per_class, softmax, seed):
    """A synthetic test set ([K * per_class] images) through the extraction epilogue and the .plk cache."""
    from tclip_b200 import features, tasks
    g = torch.Generator().manual_seed(seed)
    txt = tasks.text_prototypes(K, seed, 256)
    labels = torch.arange(K).repeat_interleave(per_class)
    emb = txt[labels] + 7.0 * torch.randn(labels.numel(), 256, generator=g) / 16.0     # un-normalised "encode_image" output
    dev = torch.device("cuda:0")
    feats = (features.softmax_features(emb.to(dev), txt.to(dev), 30.0) if softmax else features.visual_features(emb.to(dev))).cpu()
    path = (features.softmax_features_path("synthetic", "test", "RN50", 30, root=str(tmp_path)) if softmax
            else features.visual_features_path("synthetic", "test", "RN50", root=str(tmp_path)))
    features.save_features(path, feats, labels)
    loaded, lab = features.load_features(path)
    assert torch.equal(loaded, feats) and torch.equal(lab, labels)
    return loaded, lab, txt


@pytest.mark.parametrize("name_method,softmax,iters", [
    ("EM_DIRICHLET", True, 4), ("HARD_EM_DIRICHLET", True, 4), ("SOFT_KMEANS", True, 5), ("EM_GAUSSIAN", False, 5),
    ("HARD_KMEANS", False, 4),
])
def test_evaluator_call_sequence_into_drop_in_modules(dev, tmp_path, name_method, softmax, iters):
    K, per_class, batch_size, n_batches = 40, 30, 6, 3
    feats, labels, txt = _cached_features(tmp_path, K, per_class, softmax, seed=5)
    args = make_args(K, iters=iters, use_softmax_feature=softmax, name_method=name_method, number_tasks=batch_size * n_batches,
                     batch_size=batch_size, used_test_set="test", dataset="synthetic")
    model = ref_loader.StubTextModel(txt)
    seen = []

    def build(**kw):
        m = H.method_builder(name_method)(**kw)          # src.methods.zero_shot.<module>.<CLASS>: the drop-in shim
        assert type(m).__module__.startswith("tclip_b200.methods")
        run = m.run_task

        def recording_run_task(task_dic):
            seen.append({k: v.clone() for k, v in task_dic.items()})
            return run(task_dic=task_dic)
        m.run_task = recording_run_task
        return m

    random.seed(3), torch.manual_seed(3)
    acc, t, per_batch = H.evaluate_tasks(args, dev, feats, labels, model=model, log_file=os.path.join(str(tmp_path), "log.txt"), build=build)
    assert len(seen) == n_batches and seen[0]["x_q"].shape == (batch_size, 75, feats.shape[1]) and seen[0]["y_q"].shape == (batch_size, 75, 1)
    assert np.isfinite(acc) and np.isfinite(t) and t > 0
    # the same task dictionaries through the restated oracle
    want = []
    for td in seen:
        if name_method.endswith("DIRICHLET"):
            r = R.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=name_method.startswith("HARD"))
        else:
            km = {"SOFT_KMEANS": "soft", "EM_GAUSSIAN": "gauss", "HARD_KMEANS": "hard"}[name_method]
            r = R.kmeans_family(td["x_q"], td["y_q"], K, method=km, iters=iters, use_softmax_feature=softmax, text=txt)
        want.append(H.confidence_interval(r.acc[:, -1])[0])
    assert abs(acc - float(np.mean(want))) <= 1e-3, (acc, want, per_batch)

"""TEST INFRASTRUCTURE ONLY — freeze golden vectors from the LIVE reference into ``tests/golden/*.npz``.

Run in the build container (needs the reference checkout, see ``oracle/ref_loader.py``):

    python oracle/make_golden.py

Every fixture holds the exact inputs fed to the unchanged reference classes on CPU/float32 and what they
returned (``logs`` of ``run_task`` plus the instance's ``u``/``alpha``/``v``/``w``), and — from the restated
oracle in reference-exact ``broadcast`` mode — the MM iteration counts per outer iteration, which the reference
does not expose.  The script refuses to write a fixture when the restatement is not bit-identical to the live
reference on it, so the committed files pin both.  torch 2.11.0 / scipy 1.18.1 / numpy 2.3.5 generated them.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "transductive-clip_b200"))

from oracle import ref_loader, restated  # noqa: E402
from tclip_b200 import tasks  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name, family, method, setting, K, T, iters, extra
CASES = [
    ("zs_em_dirichlet_k20", "dirichlet", "EM_DIRICHLET", "zero_shot", 20, 3, 4, {}),
    ("zs_hard_em_dirichlet_k20", "dirichlet", "HARD_EM_DIRICHLET", "zero_shot", 20, 3, 3, {}),
    ("zs_em_dirichlet_k100", "dirichlet", "EM_DIRICHLET", "zero_shot", 100, 2, 3, {}),
    ("zs_hard_em_dirichlet_k100", "dirichlet", "HARD_EM_DIRICHLET", "zero_shot", 100, 2, 3, {}),
    # the reference's default outer iteration counts (config/methods_config/{em,hard_em}_dirichlet.yaml: iter 20 / 10)
    ("zs_em_dirichlet_k100_iter20", "dirichlet", "EM_DIRICHLET", "zero_shot", 100, 8, 20, {}),
    ("zs_hard_em_dirichlet_k100_iter10", "dirichlet", "HARD_EM_DIRICHLET", "zero_shot", 100, 8, 10, {}),
    ("fs_em_dirichlet_k20", "dirichlet", "EM_DIRICHLET", "few_shot", 20, 2, 3, {"shots": 2}),
    ("fs_hard_em_dirichlet_k20", "dirichlet", "HARD_EM_DIRICHLET", "few_shot", 20, 2, 3, {"shots": 2}),
    ("zs_soft_kmeans_k20", "kmeans", "SOFT_KMEANS", "zero_shot", 20, 3, 4, {"softmax": True}),
    ("zs_hard_kmeans_k20", "kmeans", "HARD_KMEANS", "zero_shot", 20, 3, 3, {"softmax": True}),
    ("zs_em_gaussian_k20", "kmeans", "EM_GAUSSIAN", "zero_shot", 20, 3, 4, {"softmax": True}),
    ("zs_soft_kmeans_visual_k20", "kmeans", "SOFT_KMEANS", "zero_shot", 20, 3, 4, {"softmax": False, "embed": 64}),
    ("zs_hard_kmeans_visual_k20", "kmeans", "HARD_KMEANS", "zero_shot", 20, 3, 3, {"softmax": False, "embed": 64}),
    ("zs_em_gaussian_visual_k20", "kmeans", "EM_GAUSSIAN", "zero_shot", 20, 3, 4, {"softmax": False, "embed": 64}),
    ("zs_em_gaussian_cov_k20", "kmeans", "EM_GAUSSIAN_COV", "zero_shot", 20, 3, 4, {"softmax": True}),
    ("zs_em_gaussian_cov_visual_k20", "kmeans", "EM_GAUSSIAN_COV", "zero_shot", 20, 3, 4, {"softmax": False, "embed": 64}),
    ("zs_kl_kmeans_k20", "kmeans", "KL_KMEANS", "zero_shot", 20, 3, 3, {"softmax": True}),
    ("zs_kl_kmeans_visual_k20", "kmeans", "KL_KMEANS", "zero_shot", 20, 3, 3, {"softmax": False, "embed": 64}),
]
KM = {"SOFT_KMEANS": "soft", "HARD_KMEANS": "hard", "EM_GAUSSIAN": "gauss", "EM_GAUSSIAN_COV": "gauss_cov", "KL_KMEANS": "kl"}


def build(case):
    name, family, method, setting, K, T, iters, extra = case
    seed = 11 + len(name)
    save = {"method": method, "setting": setting, "K": K, "iters": iters, "n_query": 75, "iter_mm": 1000,
            "k_eff": 5, "temperature": 30.0}
    model = None
    if family == "dirichlet" and setting == "zero_shot":
        td, txt = tasks.make_zero_shot_batch(T, K, seed=seed)
    elif family == "dirichlet":
        td, txt = tasks.make_few_shot_batch(T, K, shots=extra["shots"], k_eff=5, seed=seed)
        save["shots"] = extra["shots"]
    else:
        td, txt = tasks.make_zero_shot_batch(T, K, seed=seed, softmax_feature=extra["softmax"],
                                             embed_dim=extra.get("embed", tasks.EMBED_DIM))
        model = ref_loader.StubTextModel(txt)
        save["use_softmax_feature"] = extra["softmax"]
        save["text"] = txt.numpy()
    args = ref_loader.make_args(K, iters=iters, k_eff=5, use_softmax_feature=extra.get("softmax", True))
    logs, inst = ref_loader.run_reference(method, setting, td, args, model=model, shot=extra.get("shots"))

    if family == "dirichlet" and setting == "zero_shot":
        r = restated.dirichlet_zero_shot(td["x_q"], td["y_q"], K, iters=iters, hard=method.startswith("HARD"),
                                         contraction="broadcast")
    elif family == "dirichlet":
        r = restated.dirichlet_few_shot(td["x_s"], td["y_s"], td["x_q"], td["y_q"], K, 5, iters=iters,
                                        hard=method.startswith("HARD"), contraction="broadcast")
    else:
        r = restated.kmeans_family(td["x_q"], td["y_q"], K, method=KM[method], iters=iters,
                                   use_softmax_feature=extra["softmax"], text=txt, contraction="broadcast")
    # the restatement must reproduce the live reference bit for bit on this fixture
    assert torch.equal(inst.u, r.u), name
    assert np.array_equal(logs["acc"], r.acc), name
    assert np.array_equal(logs["criterions"], r.criterions, equal_nan=True), name
    if family == "dirichlet":
        assert torch.equal(inst.alpha, r.alpha) and torch.equal(inst.v, r.v), name
        save.update(alpha=inst.alpha.numpy(), v=inst.v.numpy(), mm_iters=np.asarray(r.mm_iters))
    else:
        assert torch.equal(inst.w, r.w), name
        save.update(w=inst.w.numpy())
        if method in ("EM_GAUSSIAN", "EM_GAUSSIAN_COV"):
            save.update(v=inst.v.numpy())
        if method == "EM_GAUSSIAN_COV":
            assert torch.equal(inst.s, r.s), name
            save.update(s=inst.s.numpy())
    for k, t in td.items():
        save[k] = t.numpy()
    save.update(u=inst.u.numpy(), acc=logs["acc"], criterions=logs["criterions"], preds=r.preds.numpy())
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(f"{name}: acc={logs['acc'].ravel()} mm_iters={getattr(r, 'mm_iters', None)}")


if __name__ == "__main__":
    if not ref_loader.available():
        sys.exit("reference checkout not found; golden vectors can only be regenerated in the build container")
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    only = set(sys.argv[1:])          # optional: fixture names to (re)build; the committed ones are otherwise left alone
    for c in CASES:
        if not only or c[0] in only:
            build(c)

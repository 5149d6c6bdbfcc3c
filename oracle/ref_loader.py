"""TEST INFRASTRUCTURE ONLY — loader for the *live* reference (SegoleneMartin/transductive-CLIP).

Imports the reference's own method classes, unchanged, from a checkout that is NOT part of this
repository (``$TCLIP_REF`` or ``/root/reference``).  It exists so that
  * ``oracle/make_golden.py`` can freeze golden vectors from the real reference, and
  * ``tests/test_oracle_vs_reference.py`` can prove ``oracle/restated.py`` == reference
in the build container.  The checkout does not travel to the GPU box, so nothing under ``-m gpu``,
``smoke()`` or ``bench.py`` may call this module; they use the committed fixtures instead.

Only blocker for importing the reference here is ``import clip`` at ``src/utils.py:15`` (openai/CLIP is
not installed and its weights are unobtainable offline), so a stub module is registered first.  For the
visual-feature methods the stub's ``tokenize`` returns class indices and ``StubTextModel.encode_text``
returns rows of a fixed text matrix, which ``clip_weights`` (``src/utils.py:363-377``) then L2-normalises.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

import torch

_REF_ENV = "TCLIP_REF"
_DEFAULT_REF = "/root/reference"


def reference_root() -> str | None:
    """Path of the reference checkout, or None when it is not present (e.g. on the GPU box)."""
    for cand in (os.environ.get(_REF_ENV), _DEFAULT_REF):
        if cand and os.path.isfile(os.path.join(cand, "src", "methods", "zero_shot", "em_dirichlet.py")):
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


def _install_clip_stub() -> None:
    if "clip" in sys.modules and not getattr(sys.modules["clip"], "_tclip_stub", False):
        return
    stub = types.ModuleType("clip")
    stub._tclip_stub = True
    # one "token" per prompt = its position in the class list
    stub.tokenize = lambda texts: torch.arange(len(texts)).unsqueeze(1)
    sys.modules["clip"] = stub


class StubTextModel:
    """Stands in for the CLIP model: ``encode_text(tokens)`` returns rows of a fixed [K, D] matrix."""

    def __init__(self, text_matrix: torch.Tensor):
        self.text_matrix = text_matrix

    def encode_text(self, tokens: torch.Tensor) -> torch.Tensor:
        idx = tokens.reshape(-1).long()
        return self.text_matrix[idx.cpu()].clone().to(idx.device)   # like a model living on the caller's device


_LOADED: dict[str, types.ModuleType] = {}


def load(module: str) -> types.ModuleType:
    """Import ``src.<module>`` from the reference checkout (e.g. ``methods.zero_shot.em_dirichlet``)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference checkout not found (set $TCLIP_REF); the live oracle cannot run here")
    if module in _LOADED:
        return _LOADED[module]
    _install_clip_stub()
    # The product mirror also exposes a top-level package called ``src``; make sure the reference's wins here.
    clash = sys.modules.get("src")
    if clash is not None and not str(getattr(clash, "__file__", "") or getattr(clash, "__path__", [""])[0]).startswith(root):
        for name in [n for n in sys.modules if n == "src" or n.startswith("src.")]:
            del sys.modules[name]
    # The reference's ``src`` has no __init__.py (a namespace package), so ANY regular package called ``src`` on sys.path
    # (the product mirror has one) would win regardless of order: hide those path entries during the import.
    saved = list(sys.path)
    sys.path[:] = [root] + [p for p in saved if p != root and not os.path.isdir(os.path.join(p or ".", "src"))]
    try:
        mod = importlib.import_module("src." + module)
    finally:
        sys.path[:] = saved
    _LOADED[module] = mod
    return mod


def _product_config():
    """``Cfg`` / ``make_args`` live in the product package (bench.py needs them without touching oracle/)."""
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "transductive-clip_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from tclip_b200 import config
    return config


Cfg = _product_config().Cfg
make_args = _product_config().make_args


def run_reference(method: str, setting: str, task_dic: dict, args: Cfg, model=None, shot: int | None = None):
    """Run the unchanged reference class on CPU.  Returns (logs, instance)."""
    modname = {"EM_DIRICHLET": "em_dirichlet", "HARD_EM_DIRICHLET": "hard_em_dirichlet", "EM_GAUSSIAN": "em_gaussian",
               "SOFT_KMEANS": "soft_kmeans", "HARD_KMEANS": "hard_kmeans", "EM_GAUSSIAN_COV": "em_gaussian_cov",
               "KL_KMEANS": "kl_kmeans"}[method]
    mod = load(f"methods.{setting}.{modname}")
    cls = getattr(mod, method)
    log_file = os.path.join(tempfile.mkdtemp(prefix="tclip_ref_"), "ref.log")
    inst = cls(model=model, device=torch.device("cpu"), log_file=log_file, args=args)
    td = {k: v.clone() for k, v in task_dic.items()}  # few-shot reference mutates its inputs in place
    logs = inst.run_task(task_dic=td, shot=shot) if setting == "few_shot" else inst.run_task(task_dic=td)
    return logs, inst

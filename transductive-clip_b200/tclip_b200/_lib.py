"""ctypes binding of ``libtclip_b200.so`` (C ABI declared in ``include/tclip_b200.h``).

The shared library is the product: there is no Python/PyTorch fallback.  If it has not been built
(``python __graft_entry__.py`` or ``transductive-clip_b200/csrc/build.sh``) importing the ops raises
``TclipLibraryError``; on a device that is not an sm_100 part every call fails with ``TclipError``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TCLIP_LIB") or os.path.join(_HERE, "libtclip_b200.so")   # TCLIP_LIB: diagnostic builds only

TCLIP_OK = 0
TCLIP_MM_DENSE = 0
TCLIP_MM_SKIP_DEAD = 1
TCLIP_FLAG_IN_FLIGHT = 1
TCLIP_FLAG_FULL_SOFTMAX = 2


class TclipLibraryError(RuntimeError):
    """libtclip_b200.so is missing or cannot be loaded."""


class TclipError(RuntimeError):
    """A C-ABI call returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libtclip_b200 error {code}: {message}")
        self.code = code


class DirichletProblem(ctypes.Structure):
    """Mirror of ``tclip_dirichlet_problem`` (include/tclip_b200.h)."""
    _fields_ = [
        ("n_task", c_int), ("n_query", c_int), ("n_class", c_int), ("dim", c_int),
        ("n_support", c_int),
        ("iters", c_int), ("iter_mm", c_int), ("check_every", c_int),
        ("tol", c_float), ("lambd", c_float),
        ("hard", c_int), ("mm_mode", c_int),
        ("x_q", c_void_p), ("x_s", c_void_p), ("y_s", c_void_p),
        ("u", c_void_p), ("alpha", c_void_p), ("v", c_void_p), ("labels", c_void_p),
        ("criterions", c_void_p), ("mm_iters", c_void_p), ("n_live", c_void_p), ("mm_rows", c_void_p),
        ("iter_events", POINTER(c_void_p)),
        ("mm_events", POINTER(c_void_p)),
        ("mm_crit", c_void_p),
        ("spec_probe", c_void_p),
        ("flags", c_int),
    ]


class KMeansProblem(ctypes.Structure):
    """Mirror of ``tclip_kmeans_problem`` (include/tclip_b200.h)."""
    _fields_ = [
        ("n_task", c_int), ("n_query", c_int), ("n_class", c_int), ("dim", c_int),
        ("iters", c_int), ("method", c_int),
        ("temperature", c_float), ("lambd", c_float),
        ("x", c_void_p), ("u", c_void_p), ("v", c_void_p), ("labels", c_void_p), ("coef", c_void_p), ("w", c_void_p),
        ("criterions", c_void_p),
        ("iter_events", POINTER(c_void_p)),
    ]


# name -> (restype, argtypes); every entry must be declared in include/tclip_b200.h (tests/test_cabi.py checks it)
SIGNATURES = {
    "tclip_version": (c_int, []),
    "tclip_last_error": (c_char_p, []),
    "tclip_device_check": (c_int, [c_int]),
    "tclip_mm_max_dim": (c_int, []),
    "tclip_spec_rows_cap": (c_int, []),
    "tclip_launch_count": (c_longlong, []),
    "tclip_probe_issue_rate": (c_int, [c_int, c_void_p, c_int, c_int, POINTER(ctypes.c_double), c_void_p]),
    "tclip_log_features": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "tclip_dirichlet_colsum_v": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tclip_dirichlet_moments": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                        c_int, c_int, c_void_p]),
    "tclip_dirichlet_moments_tc_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "tclip_dirichlet_moments_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                           c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "tclip_dirichlet_support_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                              c_void_p]),
    "tclip_dirichlet_mm_workspace_bytes": (c_size_t, [c_int]),
    "tclip_dirichlet_mm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                   c_void_p, c_size_t, c_void_p]),
    "tclip_dirichlet_commit": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                       c_int, c_void_p]),
    "tclip_dirichlet_estep": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int,
                                      c_int, c_int, c_int, c_int, c_void_p]),
    "tclip_dirichlet_contraction": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "tclip_cluster_prototypes": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int, c_int, c_int, c_void_p]),
    "tclip_match_clusters": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                     c_int, c_int, c_int, c_void_p]),
    "tclip_gather_tasks": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_longlong, c_int,
                                   c_void_p, c_void_p]),
    "tclip_gather_tasks_remap": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong,
                                         c_longlong, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tclip_normalize_rows": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "tclip_kmeans_similarity": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_longlong, c_int, c_int, c_void_p]),
    "tclip_kmeans_centroids": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "tclip_kmeans_precisions": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p]),
    "tclip_kmeans_assign_cov": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                        c_int, c_int, c_int, c_int, c_void_p]),
    "tclip_kmeans_assign_kl": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "tclip_kmeans_assign": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_void_p, c_void_p, c_int,
                                    c_int, c_int, c_int, c_void_p]),
    "tclip_kmeans_udiff": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_void_p]),
    "tclip_kmeans_sample_coordinates": (c_int, [c_int, c_int]),
    "tclip_kmeans_workspace_bytes": (c_size_t, [POINTER(KMeansProblem)]),
    "tclip_kmeans_run": (c_int, [POINTER(KMeansProblem), c_void_p, c_size_t, c_void_p]),
    "tclip_kmeans_expand_centroids": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "tclip_dirichlet_em_workspace_bytes": (c_size_t, [POINTER(DirichletProblem)]),
    "tclip_dirichlet_em_run": (c_int, [POINTER(DirichletProblem), c_void_p, c_size_t, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the library once per process (method instances are rebuilt for every batch by the reference's
    evaluators, ``src/eval_zero_shot.py:171``, so the handle is cached at module level)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise TclipLibraryError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (or csrc/build.sh). "
            "tclip_b200 has no CPU or PyTorch fallback.")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover - depends on the box
        raise TclipLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != TCLIP_OK:
        msg = load().tclip_last_error()
        raise TclipError(status, msg.decode("utf-8", "replace") if msg else "")

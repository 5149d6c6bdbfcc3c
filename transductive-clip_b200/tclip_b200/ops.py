"""Tensor-level wrappers of the C ABI.  PyTorch is used for device memory and streams only; every operation below
is a kernel of ``libtclip_b200.so`` launched on the current CUDA stream.  CPU tensors are rejected — there is no
fallback path."""
from __future__ import annotations

import ctypes
import threading

import torch

from . import _lib
from ._lib import DirichletProblem, TCLIP_MM_DENSE, TCLIP_MM_SKIP_DEAD, check

__all__ = ["log_features", "colsum_v", "moments", "support_stats", "mm_update_alpha", "commit", "estep",
           "cluster_prototypes", "dirichlet_em", "device_check", "launch_count", "probe_issue_rate", "normalize_rows", "kmeans_similarity",
           "kmeans_centroids", "kmeans_run", "kmeans_expand_centroids", "kmeans_assign", "kmeans_udiff", "kmeans_precisions", "kmeans_assign_cov", "kmeans_assign_kl", "KMEANS_SOFT", "KMEANS_GAUSS", "KMEANS_HARD", "TCLIP_MM_DENSE", "TCLIP_MM_SKIP_DEAD"]


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor; tclip_b200 has no CPU path")
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")
    return t


def device_check(device: int | None = None) -> None:
    lib = _lib.load()
    dev = torch.cuda.current_device() if device is None else int(device)
    check(lib.tclip_device_check(dev))


def log_features(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    lib = _lib.load()
    _need(x, torch.float32, "x")
    out = torch.empty_like(x) if out is None else _need(out, torch.float32, "out")
    check(lib.tclip_log_features(_ptr(x), _ptr(out), x.numel(), _stream()))
    return out


def colsum_v(u: torch.Tensor, want_v: bool = True, want_live: bool = True):
    """(colsum [T,K], v [T,K] | None, live int32 [T,K] | None)."""
    lib = _lib.load()
    _need(u, torch.float32, "u")
    T, n, K = u.shape
    colsum = torch.empty(T, K, device=u.device, dtype=torch.float32)
    v = torch.empty(T, K, device=u.device, dtype=torch.float32) if want_v else None
    live = torch.empty(T, K, device=u.device, dtype=torch.int32) if want_live else None
    check(lib.tclip_dirichlet_colsum_v(_ptr(u), _ptr(colsum), _ptr(v), _ptr(live), T, n, K, _stream()))
    return colsum, v, live


def moments(u, logz, colsum, support_sum=None, support_count=None, tensor_cores: bool = False) -> torch.Tensor:
    """y_cst [T,K,D]; ``tensor_cores`` selects the tcgen05 form (measured slower than the CUDA-core kernel at n = 75; the EM
    driver uses it only under TCLIP_MOMENTS=tc)."""
    lib = _lib.load()
    _need(u, torch.float32, "u"), _need(logz, torch.float32, "logz"), _need(colsum, torch.float32, "colsum")
    T, n, K = u.shape
    D = logz.shape[2]
    if support_sum is not None:
        _need(support_sum, torch.float32, "support_sum"), _need(support_count, torch.float32, "support_count")
    y = torch.empty(T, K, D, device=u.device, dtype=torch.float32)
    if tensor_cores:
        nbytes = lib.tclip_dirichlet_moments_tc_workspace_bytes(T, n, K, D)
        ws = torch.empty(nbytes, device=u.device, dtype=torch.uint8)
        check(lib.tclip_dirichlet_moments_tc(_ptr(u), _ptr(logz), _ptr(colsum), _ptr(support_sum), _ptr(support_count),
                                             _ptr(y), T, n, K, D, _ptr(ws), nbytes, _stream()))
        return y
    check(lib.tclip_dirichlet_moments(_ptr(u), _ptr(logz), _ptr(colsum), _ptr(support_sum), _ptr(support_count),
                                      _ptr(y), T, n, K, D, _stream()))
    return y


def support_stats(log_support: torch.Tensor, y_s: torch.Tensor, K: int):
    lib = _lib.load()
    _need(log_support, torch.float32, "log_support"), _need(y_s, torch.int64, "y_s")
    T, S, D = log_support.shape
    ssum = torch.empty(T, K, D, device=log_support.device, dtype=torch.float32)
    scount = torch.empty(T, K, device=log_support.device, dtype=torch.float32)
    check(lib.tclip_dirichlet_support_stats(_ptr(log_support), _ptr(y_s), _ptr(ssum), _ptr(scount), T, S, K, D,
                                            _stream()))
    return ssum, scount


def mm_update_alpha(alpha0: torch.Tensor, y: torch.Tensor, iter_mm: int = 1000, check_every: int = 50,
                    tol: float = 1e-11):
    """The M-step on [..., D] rows.  Returns (alpha, iters_done int32 device tensor)."""
    lib = _lib.load()
    _need(alpha0, torch.float32, "alpha0"), _need(y, torch.float32, "y")
    if alpha0.shape != y.shape:
        raise ValueError("alpha0 and y must have the same shape")
    D = alpha0.shape[-1]
    rows = alpha0.numel() // D
    out = torch.empty_like(alpha0)
    nbytes = lib.tclip_dirichlet_mm_workspace_bytes(rows)
    ws = torch.empty(nbytes, device=alpha0.device, dtype=torch.uint8)
    iters = torch.zeros(1, device=alpha0.device, dtype=torch.int32)
    check(lib.tclip_dirichlet_mm(_ptr(alpha0), _ptr(out), _ptr(y), rows, D, int(iter_mm), int(check_every), float(tol),
                                 _ptr(iters), _ptr(ws), nbytes, _stream()))
    return out, iters


def commit(alpha: torch.Tensor, work: torch.Tensor, live: torch.Tensor | None):
    """In place: alpha[row] <- work[row] on live rows.  Returns (criterion [1], task_criterion [T])."""
    lib = _lib.load()
    _need(alpha, torch.float32, "alpha"), _need(work, torch.float32, "work")
    if live is not None:
        _need(live, torch.int32, "live")
    T, K, D = alpha.shape
    rowstat = torch.empty(T * K * 2, device=alpha.device, dtype=torch.float64)
    task_crit = torch.empty(T, device=alpha.device, dtype=torch.float32)
    crit = torch.empty(1, device=alpha.device, dtype=torch.float32)
    check(lib.tclip_dirichlet_commit(_ptr(alpha), _ptr(work), _ptr(live), _ptr(rowstat), _ptr(task_crit), _ptr(crit),
                                     T, K, D, _stream()))
    return crit, task_crit


def estep(alpha: torch.Tensor, logz: torch.Tensor, v: torch.Tensor, lambd: float, hard: bool):
    """(u [T,n,K], labels int32 [T,n])."""
    lib = _lib.load()
    _need(alpha, torch.float32, "alpha"), _need(logz, torch.float32, "logz"), _need(v, torch.float32, "v")
    T, K, D = alpha.shape
    n = logz.shape[1]
    norm = torch.empty(T, K, device=alpha.device, dtype=torch.float64)
    u = torch.empty(T, n, K, device=alpha.device, dtype=torch.float32)
    labels = torch.empty(T, n, device=alpha.device, dtype=torch.int32)
    check(lib.tclip_dirichlet_estep(_ptr(alpha), _ptr(logz), _ptr(v), float(lambd), _ptr(norm), _ptr(u), _ptr(labels),
                                    T, n, K, D, int(bool(hard)), _stream()))
    return u, labels


CONTRACTION_MODES = {"tcgen05": 0, "tcgen05_tmem_sum": 1, "simt": 2}


def contraction(logz: torch.Tensor, alpha: torch.Tensor, mode: str = "tcgen05") -> torch.Tensor:
    """l3 [T,n,K] = sum_d logz[t,i,d] (alpha[t,k,d] - 1)  (get_logits, zero_shot/em_dirichlet.py:37-38)."""
    lib = _lib.load()
    _need(alpha, torch.float32, "alpha"), _need(logz, torch.float32, "logz")
    T, K, D = alpha.shape
    n = logz.shape[1]
    l3 = torch.empty(T, n, K, device=alpha.device, dtype=torch.float32)
    check(lib.tclip_dirichlet_contraction(_ptr(logz), _ptr(alpha), _ptr(l3), T, n, K, D, CONTRACTION_MODES[mode],
                                          _stream()))
    return l3


def cluster_prototypes(labels: torch.Tensor, feats: torch.Tensor):
    """Inputs of the label matching: dict(cluster_label, cluster_size, sample_cluster [T,n] int32,
    n_clusters [T] int32, proto [T,n,D])."""
    lib = _lib.load()
    _need(labels, torch.int32, "labels"), _need(feats, torch.float32, "feats")
    T, n, D = feats.shape
    dev = feats.device
    out = {
        "cluster_label": torch.empty(T, n, device=dev, dtype=torch.int32),
        "cluster_size": torch.empty(T, n, device=dev, dtype=torch.int32),
        "sample_cluster": torch.empty(T, n, device=dev, dtype=torch.int32),
        "n_clusters": torch.empty(T, device=dev, dtype=torch.int32),
        "proto": torch.empty(T, n, D, device=dev, dtype=torch.float32),
    }
    check(lib.tclip_cluster_prototypes(_ptr(labels), _ptr(feats), _ptr(out["cluster_label"]), _ptr(out["cluster_size"]),
                                       _ptr(out["sample_cluster"]), _ptr(out["n_clusters"]), _ptr(out["proto"]),
                                       T, n, D, _stream()))
    return out


def match_clusters(probs: torch.Tensor, n_clusters: torch.Tensor, sample_cluster: torch.Tensor,
                   y_q: torch.Tensor | None = None, graph_matching: bool = True) -> dict:
    """Cluster -> class matching on the device (``compute_graph_matching`` / ``compute_basic_matching``,
    src/utils.py:380-417): dict(cluster_class [T,n] int32, new_labels [T,n] int64, acc [T] float32 | None)."""
    lib = _lib.load()
    _need(probs, torch.float32, "probs"), _need(n_clusters, torch.int32, "n_clusters")
    _need(sample_cluster, torch.int32, "sample_cluster")
    T, rows, K = probs.shape
    n = sample_cluster.shape[1]
    dev = probs.device
    if y_q is not None:
        _need(y_q, torch.int64, "y_q")
    out = {
        "cluster_class": torch.empty(T, n, device=dev, dtype=torch.int32),
        "new_labels": torch.empty(T, n, device=dev, dtype=torch.int64),
        "acc": torch.empty(T, device=dev, dtype=torch.float32) if y_q is not None else None,
    }
    check(lib.tclip_match_clusters(_ptr(probs), _ptr(n_clusters), _ptr(sample_cluster), _ptr(y_q),
                                   int(bool(graph_matching)), _ptr(out["cluster_class"]), _ptr(out["new_labels"]),
                                   _ptr(out["acc"]), T, n, K, rows, _stream()))
    return out


def gather_tasks(features: torch.Tensor, labels: torch.Tensor | None, idx: torch.Tensor):
    """x_q [*, F] = features[idx], y_q [*] = labels[idx] on the device (``idx`` int64 CUDA tensor of any shape).
    Raises ``IndexError`` if an index lies outside the feature matrix."""
    lib = _lib.load()
    _need(features, torch.float32, "features"), _need(idx, torch.int64, "idx")
    if labels is not None:
        _need(labels, torch.int64, "labels")
    N, F = features.shape
    x_q = torch.empty(*idx.shape, F, device=features.device, dtype=torch.float32)
    y_q = torch.empty(idx.shape, device=features.device, dtype=torch.int64) if labels is not None else None
    bad = torch.zeros(1, device=features.device, dtype=torch.int32)
    check(lib.tclip_gather_tasks(_ptr(features), _ptr(labels), _ptr(idx), _ptr(x_q), _ptr(y_q), N, idx.numel(), F,
                                 _ptr(bad), _stream()))
    return x_q, y_q, bad


def gather_tasks_remap(features: torch.Tensor, labels: torch.Tensor, idx: torch.Tensor, col_perm: torch.Tensor,
                       label_map: torch.Tensor):
    """Few-shot task construction: idx [T, m], col_perm [T, U], label_map [T, n_labels] (int64 CUDA tensors) ->
    x [T, m, U] = features[idx][:, :, col_perm], y [T, m] = label_map[t, labels[idx]]; ``bad`` counts out-of-range rows."""
    lib = _lib.load()
    _need(features, torch.float32, "features"), _need(labels, torch.int64, "labels"), _need(idx, torch.int64, "idx")
    _need(col_perm, torch.int64, "col_perm"), _need(label_map, torch.int64, "label_map")
    N, F = features.shape
    T, m = idx.shape
    U, n_labels = col_perm.shape[1], label_map.shape[1]
    x = torch.empty(T, m, U, device=features.device, dtype=torch.float32)
    y = torch.empty(T, m, device=features.device, dtype=torch.int64)
    bad = torch.zeros(1, device=features.device, dtype=torch.int32)
    check(lib.tclip_gather_tasks_remap(_ptr(features), _ptr(labels), _ptr(idx), _ptr(col_perm), _ptr(label_map), _ptr(x),
                                       _ptr(y), N, T * m, m, F, U, n_labels, _ptr(bad), _stream()))
    return x, y, bad


# ---- k-means family ------------------------------------------------------------------------------------------------
KMEANS_SOFT, KMEANS_GAUSS, KMEANS_HARD = 0, 1, 2


def normalize_rows(x: torch.Tensor) -> torch.Tensor:
    """x / ||x|| along the last dimension."""
    lib = _lib.load()
    _need(x, torch.float32, "x")
    out = torch.empty_like(x)
    D = x.shape[-1]
    check(lib.tclip_normalize_rows(_ptr(x), _ptr(out), x.numel() // D, D, _stream()))
    return out


def kmeans_similarity(a: torch.Tensor, text: torch.Tensor, scale: float) -> torch.Tensor:
    """softmax_k(scale * a @ text.T) for a [..., D], text [K, D] -> [..., K]."""
    lib = _lib.load()
    _need(a, torch.float32, "a"), _need(text, torch.float32, "text")
    K, D = text.shape
    M = a.numel() // D
    u = torch.empty(*a.shape[:-1], K, device=a.device, dtype=torch.float32)
    check(lib.tclip_kmeans_similarity(_ptr(a), _ptr(text), float(scale), _ptr(u), M, K, D, _stream()))
    return u


CENTROIDS_ZERO_EMPTY, CENTROIDS_KEEP_EMPTY, CENTROIDS_KL = 0, 1, 2


def kmeans_centroids(u: torch.Tensor, x: torch.Tensor, w: torch.Tensor | None, keep_old: bool = False,
                     mode: int | None = None) -> torch.Tensor:
    """Centroid update; ``w`` is updated in place (allocated when None: then empty clusters are zero)."""
    lib = _lib.load()
    _need(u, torch.float32, "u"), _need(x, torch.float32, "x")
    T, n, K = u.shape
    D = x.shape[2]
    if mode is None:
        mode = CENTROIDS_KEEP_EMPTY if (keep_old and w is not None) else CENTROIDS_ZERO_EMPTY
    if w is None:
        w = torch.empty(T, K, D, device=u.device, dtype=torch.float32)
    else:
        _need(w, torch.float32, "w")
    check(lib.tclip_kmeans_centroids(_ptr(u), _ptr(x), _ptr(w), T, n, K, D, int(mode), _stream()))
    return w


def kmeans_precisions(u: torch.Tensor, x: torch.Tensor, w: torch.Tensor, s: torch.Tensor | None) -> torch.Tensor:
    """Diagonal precisions; ``s`` is updated in place keeping the rows of empty clusters (allocated when None: s_init)."""
    lib = _lib.load()
    _need(u, torch.float32, "u"), _need(x, torch.float32, "x"), _need(w, torch.float32, "w")
    T, n, K = u.shape
    D = x.shape[2]
    keep = s is not None
    if s is None:
        s = torch.empty(T, K, D, device=u.device, dtype=torch.float32)
    else:
        _need(s, torch.float32, "s")
    check(lib.tclip_kmeans_precisions(_ptr(u), _ptr(x), _ptr(w), _ptr(s), T, n, K, D, int(keep), _stream()))
    return s


def kmeans_assign_cov(x: torch.Tensor, w: torch.Tensor, s: torch.Tensor, v: torch.Tensor, lambd: float):
    """(u, labels) of EM-Gaussian with diagonal covariance."""
    lib = _lib.load()
    for t_, nm in ((x, "x"), (w, "w"), (s, "s"), (v, "v")):
        _need(t_, torch.float32, nm)
    T, n, D = x.shape
    K = w.shape[1]
    u = torch.empty(T, n, K, device=x.device, dtype=torch.float32)
    det = torch.empty(T, K, device=x.device, dtype=torch.float32)
    labels = torch.empty(T, n, device=x.device, dtype=torch.int32)
    check(lib.tclip_kmeans_assign_cov(_ptr(x), _ptr(w), _ptr(s), _ptr(v), float(lambd), _ptr(det), _ptr(u), _ptr(labels),
                                      T, n, K, D, _stream()))
    return u, labels


def kmeans_assign_kl(x: torch.Tensor, w: torch.Tensor):
    """(one-hot u, labels) of KL k-means."""
    lib = _lib.load()
    _need(x, torch.float32, "x"), _need(w, torch.float32, "w")
    T, n, D = x.shape
    K = w.shape[1]
    u = torch.empty(T, n, K, device=x.device, dtype=torch.float32)
    labels = torch.empty(T, n, device=x.device, dtype=torch.int32)
    check(lib.tclip_kmeans_assign_kl(_ptr(x), _ptr(w), _ptr(u), _ptr(labels), T, n, K, D, _stream()))
    return u, labels


def kmeans_assign(x: torch.Tensor, w: torch.Tensor, mode: int, temperature: float, v: torch.Tensor | None = None,
                  lambd: float = 0.0):
    """(u [T,n,K], labels int32 [T,n]) from the squared distances to the centroids."""
    lib = _lib.load()
    _need(x, torch.float32, "x"), _need(w, torch.float32, "w")
    if v is not None:
        _need(v, torch.float32, "v")
    T, n, D = x.shape
    K = w.shape[1]
    u = torch.empty(T, n, K, device=x.device, dtype=torch.float32)
    labels = torch.empty(T, n, device=x.device, dtype=torch.int32)
    check(lib.tclip_kmeans_assign(_ptr(x), _ptr(w), _ptr(v), float(temperature), float(lambd), int(mode), _ptr(u),
                                  _ptr(labels), T, n, K, D, _stream()))
    return u, labels


def kmeans_udiff(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """mean over tasks of ||a[t] - b[t]||_F  -> tensor [1]."""
    lib = _lib.load()
    _need(a, torch.float32, "a"), _need(b, torch.float32, "b")
    T = a.shape[0]
    task = torch.empty(T, device=a.device, dtype=torch.float32)
    out = torch.empty(1, device=a.device, dtype=torch.float32)
    check(lib.tclip_kmeans_udiff(_ptr(a), _ptr(b), _ptr(task), _ptr(out), T, a.numel() // T, _stream()))
    return out


def kmeans_run(x: torch.Tensor, u0: torch.Tensor, method: int, iters: int, temperature: float, lambd: float = 0.0,
               want_w: bool = False, record_events: bool = False) -> dict:
    """The fused driver ``tclip_kmeans_run``: the whole soft k-means / EM-Gaussian / hard k-means loop enqueued on the
    current stream.  ``u0`` [T,n,K] is the initial assignment (updated in place and returned as ``u``).  Returns device
    tensors u, labels, v (EM-Gaussian), criterions, and either ``coef`` [T,n,K] (sample-coordinate form: the centroids
    are ``kmeans_expand_centroids(coef, x)``) or ``w`` [T,K,D]."""
    lib = _lib.load()
    _need(x, torch.float32, "x"), _need(u0, torch.float32, "u0")
    T, n, D = x.shape
    K = u0.shape[2]
    dev = x.device
    coords = bool(lib.tclip_kmeans_sample_coordinates(n, D))
    n_crit = (2 * iters) if method == KMEANS_HARD else iters
    crit = torch.zeros(max(n_crit, 1), device=dev, dtype=torch.float32)   # (iters = 0: the library still wants a buffer)
    out = {
        "u": u0,
        "labels": torch.zeros(T, n, device=dev, dtype=torch.int32),
        "v": torch.zeros(T, K, device=dev, dtype=torch.float32) if method == KMEANS_GAUSS else None,
        "criterions": crit[:n_crit],
        "coef": torch.empty(T, n, K, device=dev, dtype=torch.float32) if coords else None,
        "w": torch.empty(T, K, D, device=dev, dtype=torch.float32) if (want_w or not coords) else None,
    }
    events = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)] if record_events else []
    ev_arr = None
    if events:
        for e in events:
            e.record()
        ev_arr = (ctypes.c_void_p * (iters + 1))(*[e.cuda_event for e in events])
    p = _lib.KMeansProblem(
        n_task=T, n_query=n, n_class=K, dim=D, iters=int(iters), method=int(method), temperature=float(temperature),
        lambd=float(lambd), x=_ptr(x), u=_ptr(u0), v=_ptr(out["v"]), labels=_ptr(out["labels"]), coef=_ptr(out["coef"]),
        w=_ptr(out["w"]), criterions=_ptr(crit),
        iter_events=ctypes.cast(ev_arr, ctypes.POINTER(ctypes.c_void_p)) if ev_arr is not None else None)
    nbytes = lib.tclip_kmeans_workspace_bytes(ctypes.byref(p))
    if nbytes == 0:
        check(-1)
    ws = _workspace(nbytes, dev)
    check(lib.tclip_kmeans_run(ctypes.byref(p), _ptr(ws), ws.numel(), _stream()))
    out["events"] = events
    return out


def kmeans_expand_centroids(coef: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """w [T,K,D] = coef^T x per task (the centroids behind the coefficients ``kmeans_run`` returns)."""
    lib = _lib.load()
    _need(coef, torch.float32, "coef"), _need(x, torch.float32, "x")
    T, n, K = coef.shape
    D = x.shape[2]
    w = torch.empty(T, K, D, device=x.device, dtype=torch.float32)
    check(lib.tclip_kmeans_expand_centroids(_ptr(coef), _ptr(x), _ptr(w), T, n, K, D, _stream()))
    return w


_WORKSPACES: dict = {}
_WORKSPACES_LOCK = threading.Lock()


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """One cached scratch buffer per (device, CUDA stream), grown on demand (a new method object is built per batch;
    batches in flight on different streams — ``tclip_b200.pipeline`` — must not share scratch)."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    with _WORKSPACES_LOCK:
        ws = _WORKSPACES.get(key)
        if ws is None or ws.numel() < nbytes:
            _WORKSPACES.pop(key, None)
            ws = torch.empty(nbytes, device=device, dtype=torch.uint8)
            _WORKSPACES[key] = ws
    return ws


def release_workspaces(stream_ids=None) -> None:
    """Drop cached scratch buffers: those of the given CUDA stream handles, or all of them."""
    with _WORKSPACES_LOCK:
        for key in [k for k in _WORKSPACES if stream_ids is None or k[2] in stream_ids]:
            del _WORKSPACES[key]


def dirichlet_em(x_q: torch.Tensor, n_class: int, iters: int, iter_mm: int, lambd: float, hard: bool,
                 x_s: torch.Tensor | None = None, y_s: torch.Tensor | None = None, check_every: int = 50,
                 tol: float = 1e-11, mm_mode: int = TCLIP_MM_DENSE, record_events: bool = False,
                 spec_probe: bool = False, in_flight: bool = False, full_softmax: bool = False) -> dict:
    """The fused driver ``tclip_dirichlet_em_run``: the whole EM loop enqueued on the current stream.
    Returns device tensors u, alpha, v, labels, criterions, mm_iters, n_live, mm_rows (+ ``events``).
    ``spec_probe`` (measurement only) adds ``spec_probe`` int32 [iters, cap, 4]: per live row of the few-rows M-step kernel
    {iterations executed, fixed-point iteration | -1, cycle-detection iteration | -1, period}.  ``in_flight``: the caller
    keeps several batches in flight (``TCLIP_FLAG_IN_FLIGHT``: register-lean tail kernel, same results)."""
    lib = _lib.load()
    _need(x_q, torch.float32, "x_q")
    T, n, D = x_q.shape
    K = int(n_class)
    dev = x_q.device
    S = 0
    if x_s is not None:
        _need(x_s, torch.float32, "x_s"), _need(y_s, torch.int64, "y_s")
        S = x_s.shape[1]
    if D != K:
        raise ValueError("the Dirichlet methods need softmax features: feature dim must equal n_class")
    out = {
        "u": torch.empty(T, n, K, device=dev, dtype=torch.float32),
        "alpha": torch.empty(T, K, D, device=dev, dtype=torch.float32),
        "v": torch.empty(T, K, device=dev, dtype=torch.float32),
        "labels": torch.zeros(T, n, device=dev, dtype=torch.int32),
        "criterions": torch.zeros(max(iters, 1), device=dev, dtype=torch.float32)[:iters],
        "mm_iters": torch.zeros(max(iters, 1), device=dev, dtype=torch.int32)[:iters],
        "n_live": torch.zeros(max(iters, 1), device=dev, dtype=torch.int32)[:iters],
        "mm_rows": torch.zeros(max(iters, 1), device=dev, dtype=torch.int64)[:iters],
        "mm_crit": torch.zeros(max(iters, 1), 2, device=dev, dtype=torch.float64)[:iters],
    }
    if spec_probe:
        out["spec_probe"] = torch.full((max(iters, 1), lib.tclip_spec_rows_cap(), 4), -2, device=dev, dtype=torch.int32)
    events = [torch.cuda.Event(enable_timing=True) for _ in range(iters)] if record_events else []
    mm_events = [torch.cuda.Event(enable_timing=True) for _ in range(2 * iters)] if record_events else []
    ev_arr = mm_arr = None
    if events:
        for e in events + mm_events:  # torch creates the underlying cudaEvent_t lazily
            e.record()
        ev_arr = (ctypes.c_void_p * iters)(*[e.cuda_event for e in events])
        mm_arr = (ctypes.c_void_p * (2 * iters))(*[e.cuda_event for e in mm_events])
    p = DirichletProblem(
        n_task=T, n_query=n, n_class=K, dim=D, n_support=S, iters=int(iters), iter_mm=int(iter_mm),
        check_every=int(check_every), tol=float(tol), lambd=float(lambd), hard=int(bool(hard)), mm_mode=int(mm_mode),
        x_q=_ptr(x_q), x_s=_ptr(x_s), y_s=_ptr(y_s), u=_ptr(out["u"]), alpha=_ptr(out["alpha"]), v=_ptr(out["v"]),
        labels=_ptr(out["labels"]), criterions=_ptr(out["criterions"]), mm_iters=_ptr(out["mm_iters"]),
        n_live=_ptr(out["n_live"]), mm_rows=_ptr(out["mm_rows"]), mm_crit=_ptr(out["mm_crit"]),
        spec_probe=_ptr(out.get("spec_probe")), flags=(_lib.TCLIP_FLAG_IN_FLIGHT if in_flight else 0) | (_lib.TCLIP_FLAG_FULL_SOFTMAX if full_softmax else 0),
        iter_events=ctypes.cast(ev_arr, ctypes.POINTER(ctypes.c_void_p)) if ev_arr is not None else None,
        mm_events=ctypes.cast(mm_arr, ctypes.POINTER(ctypes.c_void_p)) if mm_arr is not None else None)
    nbytes = lib.tclip_dirichlet_em_workspace_bytes(ctypes.byref(p))
    if nbytes == 0:
        check(-1)
    ws = _workspace(nbytes, dev)
    check(lib.tclip_dirichlet_em_run(ctypes.byref(p), _ptr(ws), ws.numel(), _stream()))
    out["events"] = events
    out["mm_events"] = mm_events
    return out


def launch_count() -> int:
    """Kernels launched by libtclip_b200 in this process so far."""
    return int(_lib.load().tclip_launch_count())


def probe_issue_rate(which: str, n_blocks: int, iters: int, device=None) -> tuple[float, float]:
    """Run the register-only FFMA (``"ffma"``) or MUFU (``"mufu"``) microbenchmark once on the current stream.
    Returns (operations executed, milliseconds)."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    sink = torch.zeros(4, device=dev, dtype=torch.float32)
    ops_out = ctypes.c_double(0.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib.tclip_probe_issue_rate({"ffma": 0, "mufu": 1, "ffma2": 2, "mix": 3}[which], _ptr(sink), int(n_blocks), int(iters),
                                     ctypes.byref(ops_out), _stream()))
    e1.record()
    e1.synchronize()
    return ops_out.value, e0.elapsed_time(e1)

"""Cluster -> class label matching, host form: TEST INFRASTRUCTURE, the checker of ``ops.match_clusters`` (the method classes use the
device kernel; this file is not part of the product package).

Mirrors ``compute_graph_matching`` / ``compute_basic_matching`` of the reference (``src/utils.py:380-417``): per task,
clusters are taken in order of first appearance among the predictions, the cost row of cluster c is
``-probs[task, c, :]`` in float64, and SciPy's ``linear_sum_assignment`` (the reference's own third-party solver,
``src/utils.py:14,401``) gives the cluster -> class map.  The per-query Python loops of the reference
(7.2 ms/task) are replaced by the device-side ``tclip_cluster_prototypes`` kernel, which already
delivers the cost rows in first-appearance order, plus one vectorised gather here.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linear_sum_assignment


def graph_matching(proto: np.ndarray, n_clusters: np.ndarray, sample_cluster: np.ndarray) -> np.ndarray:
    """proto [T, n, K] float32 (rows beyond n_clusters[t] unused), sample_cluster [T, n] -> new labels [T, n]."""
    T, n = sample_cluster.shape
    out = np.empty((T, n), dtype=np.int64)
    for t in range(T):
        c = int(n_clusters[t])
        cost = -proto[t, :c].astype(np.float64)
        _, cols = linear_sum_assignment(cost, maximize=False)
        out[t] = cols[sample_cluster[t]]
    return out


def basic_matching(proto: np.ndarray, n_clusters: np.ndarray, sample_cluster: np.ndarray) -> np.ndarray:
    """Each cluster takes the arg-max class of its prototype (``compute_basic_matching``)."""
    T, n = sample_cluster.shape
    out = np.empty((T, n), dtype=np.int64)
    for t in range(T):
        c = int(n_clusters[t])
        best = proto[t, :c].argmax(axis=-1)
        out[t] = best[sample_cluster[t]]
    return out

"""EM-Dirichlet / Hard EM-Dirichlet method classes, zero-shot and few-shot, on libtclip_b200.

Drop-in for the reference classes (same constructor, same ``run_task`` contract, same logs):
  zero-shot  ``src/methods/zero_shot/em_dirichlet.py`` (BASE :9-121, EM_DIRICHLET :124-246),
             ``src/methods/zero_shot/hard_em_dirichlet.py`` (HARD_EM_DIRICHLET :124-271)
  few-shot   ``src/methods/few_shot/em_dirichlet.py`` (BASE :9-91, EM_DIRICHLET :94-220),
             ``src/methods/few_shot/hard_em_dirichlet.py`` (HARD_EM_DIRICHLET :94-251)
The callers are ``src/eval_zero_shot.py:113-138,171-177`` and ``src/eval_few_shot.py:189-211,250-259``.

All arithmetic of ``run_method`` happens in CUDA kernels behind ``ops.dirichlet_em`` (one C-ABI call that enqueues the
whole EM loop), the cluster -> class assignment included (``ops.match_clusters``: SciPy's shortest-augmenting-path
algorithm restated for one warp per task; ``tests/scipy_matching.py`` keeps the SciPy form the tests compare it with).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import ops
from ..logger import Logger
from ..pipeline import batches_in_flight

_MM_MODES = {"dense": ops.TCLIP_MM_DENSE, "skip_dead": ops.TCLIP_MM_SKIP_DEAD}


def _cfg(args, key, default):
    try:
        return args[key] if key in args else default
    except TypeError:
        return getattr(args, key, default)


class _DirichletBase(object):
    """Shared body of the four classes; the zero-/few-shot ``BASE`` classes below only differ in lambda, the inputs
    they move to the device and how accuracy is computed."""

    hard = False
    few_shot = False

    def __init__(self, model, device, log_file, args):
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.iter = args.iter
        self.model = model
        self.log_file = log_file
        self.logger = Logger(type(self).__module__, self.log_file)
        self.init_info_lists()
        self.args = args
        self.eps = 1e-15
        self.iter_mm = args.iter_mm
        mode = _cfg(args, "mm_mode", os.environ.get("TCLIP_MM_MODE", "skip_dead"))
        if mode not in _MM_MODES:
            raise ValueError(f"mm_mode must be one of {sorted(_MM_MODES)}, got {mode!r}")
        self.mm_mode = mode
        self.u = self.alpha = self.v = None
        self.mm_iters = self.n_live = self.mm_rows = self.mm_crit = None

    def __del__(self):
        try:
            self.logger.del_logger()
        except Exception:
            pass

    def init_info_lists(self):
        self.timestamps = []
        self.criterions = []
        self.test_acc = []

    def record_convergence(self, new_time, criterions):
        self.criterions.append(criterions)
        self.timestamps.append(new_time)

    def get_logs(self):
        self.criterions = torch.stack(self.criterions, dim=0).cpu().numpy()
        self.test_acc = torch.cat(self.test_acc, dim=1).cpu().numpy()
        return {'timestamps': np.array(self.timestamps).mean(), 'criterions': self.criterions,
                'acc': self.test_acc}

    # ------------------------------------------------------------------------------------------------------------
    def _require_cuda(self):
        if self.device.type != "cuda":
            raise RuntimeError("tclip_b200 runs on B200 GPUs only: device must be a CUDA device (no CPU fallback)")
        if self.device.index is not None:
            torch.cuda.set_device(self.device)

    def _to_device(self, t, dtype=None):
        t = t.to(self.device, non_blocking=True)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()

    def compute_acc(self, y_q):
        """Plain arg-max accuracy (few_shot/em_dirichlet.py:50-58)."""
        preds_q = self.labels.long()
        accuracy = (preds_q == y_q).float().mean(1, keepdim=True)
        self.test_acc.append(accuracy)

    def compute_acc_clustering(self, query, y_q):
        """Cluster prototypes, minimum-cost cluster -> class assignment and accuracy on the device
        (zero_shot/em_dirichlet.py:61-92, src/utils.py:380-417)."""
        if not self.args.use_softmax_feature:
            raise ValueError("The selected method is unable to handle query features that are not in the unit simplex")
        cl = ops.cluster_prototypes(self.labels, query)
        # softmax features: the prototypes are the class probabilities (zero_shot/em_dirichlet.py:72-74); assignment and
        # accuracy stay on the device, nothing but the accuracies crosses PCIe
        res = ops.match_clusters(cl["proto"], cl["n_clusters"], cl["sample_cluster"], y_q.contiguous(),
                                 graph_matching=(self.args.graph_matching == True))  # noqa: E712 (reference's test)
        self.new_labels = res["new_labels"]
        accuracy = res["acc"].unsqueeze(1)
        self.test_acc.append(accuracy)

    def _run_em(self, query, support=None, y_s=None):
        n_task = query.shape[0]
        if not self.args.use_softmax_feature:
            raise ValueError("The selected method is unable to handle query features that are not in the unit simplex")
        start = torch.cuda.Event(enable_timing=True)
        start.record()
        res = ops.dirichlet_em(query, self.args.num_classes_test, self.iter, self.iter_mm, float(self.lambd), self.hard,
                               x_s=support, y_s=y_s, mm_mode=_MM_MODES[self.mm_mode], record_events=True,
                               in_flight=batches_in_flight() > 1)
        self.u, self.alpha, self.v, self.labels = res["u"], res["alpha"], res["v"], res["labels"]
        self.mm_iters, self.n_live, self.mm_rows = res["mm_iters"], res["n_live"], res["mm_rows"]
        self.mm_crit = res["mm_crit"]
        self._em_events = (start, res["events"])
        self._mm_events = res["mm_events"]
        crit = res["criterions"]
        if self.hard and self.few_shot:
            crit = torch.zeros_like(crit)  # few_shot/hard_em_dirichlet.py:234-244 logs a criterion that is always 0
        return n_task, crit

    def _log_iterations(self, n_task, crit):
        # cumulative device time per task after every outer iteration (the reference logs un-synchronised wall time)
        start, events = self._em_events
        if events:
            events[-1].synchronize()
        for i in range(self.iter):
            t = start.elapsed_time(events[i]) / 1000.0
            self.record_convergence(new_time=t / n_task, criterions=crit[i])


# ----------------------------------------------------------------------------------------------------------------
# zero-shot
# ----------------------------------------------------------------------------------------------------------------
class BASE(_DirichletBase):
    """Zero-shot base: lambda = int(K / 5) * n_query (zero_shot/em_dirichlet.py:14)."""

    def __init__(self, model, device, log_file, args):
        super().__init__(model, device, log_file, args)
        self.lambd = int(args.num_classes_test / 5) * args.n_query

    def run_task(self, task_dic):
        """task_dic: {'x_q': float [T, n, K], 'y_q': int64 [T, n, 1]} CPU tensors -> logs dict."""
        self._require_cuda()
        y_q = task_dic['y_q']
        query = task_dic['x_q']
        query = self._to_device(query, torch.float32)
        y_q = y_q.long().squeeze(2).to(self.device, non_blocking=True)
        del task_dic
        self.run_method(query=query, y_q=y_q)
        return self.get_logs()

    def run_method(self, query, y_q):
        self.logger.info(" ==> Executing {} with LAMBDA = {} and T = {}".format(self._title, self.lambd, self.args.T))
        n_task, crit = self._run_em(query)
        self._log_iterations(n_task, crit)
        self.compute_acc_clustering(query, y_q)


class EM_DIRICHLET(BASE):
    hard = False
    _title = "EM-DIRICHLET"


class HARD_EM_DIRICHLET(BASE):
    hard = True
    _title = "HARD EM-DIRICHLET"


# ----------------------------------------------------------------------------------------------------------------
# few-shot
# ----------------------------------------------------------------------------------------------------------------
class FEW_SHOT_BASE(_DirichletBase):
    """Few-shot base: lambda = int(K / k_eff) * n_query (few_shot/em_dirichlet.py:14); support labels enter the
    M-step moments as fixed one-hot responsibilities (:196-200); plain arg-max accuracy (:220)."""

    few_shot = True

    def __init__(self, model, device, log_file, args):
        super().__init__(model, device, log_file, args)
        self.lambd = int(args.num_classes_test / args.k_eff) * args.n_query

    def run_task(self, task_dic, shot=10):
        self._require_cuda()
        y_s = task_dic['y_s']
        y_q = task_dic['y_q']
        support = task_dic['x_s']
        query = task_dic['x_q']
        support = self._to_device(support, torch.float32)
        query = self._to_device(query, torch.float32)
        y_s = y_s.long().squeeze(2).to(self.device, non_blocking=True).contiguous()
        y_q = y_q.long().squeeze(2).to(self.device, non_blocking=True)
        del task_dic
        self.run_method(support=support, query=query, y_s=y_s, y_q=y_q)
        return self.get_logs()

    def run_method(self, support, query, y_s, y_q):
        self.logger.info(" ==> Executing {} with LAMBDA = {} and T = {}".format(self._title, self.lambd, self.args.T))
        n_task, crit = self._run_em(query, support=support, y_s=y_s)
        self._log_iterations(n_task, crit)
        self.compute_acc(y_q=y_q)


class FEW_SHOT_EM_DIRICHLET(FEW_SHOT_BASE):
    hard = False
    _title = "EM-DIRICHLET"


class FEW_SHOT_HARD_EM_DIRICHLET(FEW_SHOT_BASE):
    hard = True
    _title = "HARD EM-DIRICHLET"

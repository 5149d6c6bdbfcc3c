"""Soft k-means, hard k-means and EM-Gaussian (identity covariance) method classes on libtclip_b200.

Drop-in for the reference classes (same constructor, ``run_task`` contract and logs):
  ``src/methods/zero_shot/soft_kmeans.py`` (SOFT_KMEANS :96-220), ``src/methods/zero_shot/hard_kmeans.py``
  (HARD_KMEANS :118-211), ``src/methods/zero_shot/em_gaussian.py`` (EM_GAUSSIAN :97-229); the caller is
  ``src/eval_zero_shot.py:119-137,171-177``.

The loop (w -> u [-> v]) is the reference's, every numeric step is a kernel of ``csrc/kmeans.cu`` reached through the C
ABI, the cluster -> class assignment included (``ops.match_clusters``).  Both feature kinds are
supported: softmax features (u starts from the features) and visual features (u starts from
``softmax(T * normalize(x) @ text.T)``; ``text`` comes from ``model.encode_text`` exactly like ``clip_weights``,
``src/utils.py:363-377``).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..logger import Logger


def clip_weights(model, classnames, template, device):
    """Unit-norm text embeddings of the class prompts (``src/utils.py:363-377``).  Needs the ``clip`` package for its
    tokenizer, like the reference."""
    import clip  # noqa: PLC0415  (openai/CLIP; tests register a stub)
    names = [c.replace('_', ' ') for c in classnames]
    tokens = torch.cat([clip.tokenize([template.format(c) for c in names])]).to(device)
    with torch.no_grad():
        text = model.encode_text(tokens).float()
    return ops.normalize_rows(text.to(device).contiguous())


class _KMeansBase(object):
    mode = ops.KMEANS_SOFT
    _title = "SOFT K-MEANS"

    def __init__(self, model, device, log_file, args):
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.iter = args.iter
        self.model = model
        self.log_file = log_file
        self.logger = Logger(type(self).__module__, self.log_file)
        self.init_info_lists()
        self.args = args
        self.eps = 1e-15
        self.u = self.v = self.labels = None
        self._w = self._coef = self._query = None

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

    def run_task(self, task_dic):
        """task_dic: {'x_q': float [T, n, F], 'y_q': int64 [T, n, 1]} CPU tensors -> logs dict."""
        self._require_cuda()
        y_q = task_dic['y_q']
        query = task_dic['x_q']
        query = query.to(self.device, non_blocking=True).float().contiguous()
        y_q = y_q.long().squeeze(2).to(self.device, non_blocking=True)
        del task_dic
        self.run_method(query=query, y_q=y_q)
        return self.get_logs()

    def _text(self):
        return clip_weights(self.model, self.args.classnames, self.args.template, self.device)

    def _initial_u(self, query):
        if self.args.use_softmax_feature:
            return query.clone()
        self._text_features = self._text()
        return ops.kmeans_similarity(ops.normalize_rows(query), self._text_features, float(self.args.T))

    def compute_acc_clustering(self, query, y_q):
        """Prototypes of the arg-max clusters, cluster -> class assignment and accuracy on the device
        (soft_kmeans.py:33-66, src/utils.py:380-417)."""
        cl = ops.cluster_prototypes(self.labels, query)
        probs = cl["proto"]
        if not self.args.use_softmax_feature:
            # probs = softmax(T * normalize(prototype) @ text.T) for all n prototype rows (rows beyond a task's cluster count
            # are zero prototypes -> NaN after the normalisation, as upstream, and are never read by the matching): one
            # tensor-core product, no host round trip for the cluster count
            probs = ops.kmeans_similarity(ops.normalize_rows(probs), self._text_features, float(self.args.T))
        res = ops.match_clusters(probs, cl["n_clusters"], cl["sample_cluster"], y_q.contiguous(),
                                 graph_matching=(self.args.graph_matching == True))  # noqa: E712 (reference's test)
        self.new_labels = res["new_labels"]
        self.test_acc.append(res["acc"].unsqueeze(1))

    # ------------------------------------------------------------------------------------------------------------
    @property
    def w(self):
        """Centroids [T,K,D].  The fused loop works in the coordinates of the task's own samples and never forms them
        (csrc/kmeans_run.cu); they are produced from the coefficients the first time somebody asks."""
        if self._w is None and self._coef is not None:
            self._w = ops.kmeans_expand_centroids(self._coef, self._query)
        return self._w

    @w.setter
    def w(self, value):
        self._w = value

    def run_method(self, query, y_q):
        """soft_kmeans.py:168-220 / hard_kmeans.py:153-211 / em_gaussian.py:171-229: one C-ABI call enqueues the whole loop
        (w_init, ``iter`` x {w_update, u_update[, v_update]}), no host synchronisation inside."""
        self.logger.info(" ==> Executing {} with T = {}".format(self._title, self.args.T))
        n_task = query.shape[0]
        hard = self.mode == ops.KMEANS_HARD
        u0 = self._initial_u(query)
        res = ops.kmeans_run(query, u0, self.mode, self.iter, float(self.args.T), lambd=float(getattr(self, "lambd", 0.0)),
                             record_events=True)
        self.u, self.labels, self.v = res["u"], res["labels"], res["v"]
        self._coef, self._w, self._query = res["coef"], res["w"], query
        crit, events = res["criterions"], res["events"]
        self._events = events
        self.compute_acc_clustering(query, y_q)
        # per-iteration device time (the reference logs un-synchronised wall time per iteration, soft_kmeans.py:203,216-218)
        events[-1].synchronize()
        for i in range(self.iter):
            dt = events[i].elapsed_time(events[i + 1]) / 1000.0
            if hard:
                self.record_convergence(new_time=dt, criterions=crit[2 * i])           # hard_kmeans.py:203 (un-normalised) ...
                self.record_convergence(new_time=dt / n_task, criterions=crit[2 * i + 1])   # ... and :208-209
            else:
                self.record_convergence(new_time=dt / n_task, criterions=crit[i])


class SOFT_KMEANS(_KMeansBase):
    mode = ops.KMEANS_SOFT
    _title = "SOFT K-MEANS"


class HARD_KMEANS(_KMeansBase):
    mode = ops.KMEANS_HARD
    _title = "HARD_KMEANS"


class EM_GAUSSIAN(_KMeansBase):
    """lambda = int(K / 5) * n_query (em_gaussian.py:20)."""
    mode = ops.KMEANS_GAUSS
    _title = "EM-GAUSSIAN"

    def __init__(self, model, device, log_file, args):
        super().__init__(model, device, log_file, args)
        self.lambd = int(args.num_classes_test / 5) * args.n_query


class EM_GAUSSIAN_COV(_KMeansBase):
    """EM-Gaussian with diagonal covariance (``src/methods/zero_shot/em_gaussian_cov.py:97-257``): centroids w, diagonal
    precisions s, responsibilities u = softmax(-1/2 sum_d s (w - x)^2 + 1/2 sum_d log s + lambda v / n) — no temperature in
    the E-step — and the class-proportion dual v."""
    _title = "EM_GAUSSIAN_COV"

    def __init__(self, model, device, log_file, args):
        super().__init__(model, device, log_file, args)
        self.lambd = int(args.num_classes_test / 5) * args.n_query
        self.s = None

    def run_method(self, query, y_q):
        self.logger.info(" ==> Executing {} with T = {}".format(self._title, self.args.T))
        n_task, n_class = query.shape[0], self.args.num_classes_test
        self.v = torch.zeros(n_task, n_class, device=self.device)
        self.u = self._initial_u(query)
        self.w = ops.kmeans_centroids(self.u, query, None)                  # w_init
        self.s = ops.kmeans_precisions(self.u, query, self.w, None)        # s_init
        zero = torch.zeros((), device=self.device)
        for _ in range(self.iter):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            self.w = ops.kmeans_centroids(self.u, query, self.w, keep_old=True)
            self.s = ops.kmeans_precisions(self.u, query, self.w, self.s)
            self.u, self.labels = ops.kmeans_assign_cov(query, self.w, self.s, self.v, float(self.lambd))
            _, self.v, _ = ops.colsum_v(self.u, want_v=True, want_live=False)
            t1.record()
            t1.synchronize()
            self.record_convergence(new_time=t0.elapsed_time(t1) / 1000.0 / n_task, criterions=zero)
        self.compute_acc_clustering(query, y_q)


class KL_KMEANS(_KMeansBase):
    """KL k-means (``src/methods/zero_shot/kl_kmeans.py:114-189``): centroids = cluster means (size clamped at 1), hard
    assignment to the centroid of minimum KL(x || w); the criterion is recorded twice per iteration as upstream."""
    _title = "KL KMEANS"

    def run_method(self, query, y_q):
        self.logger.info(" ==> Executing {} with T = {}".format(self._title, self.args.T))
        n_task = query.shape[0]
        self.u = self._initial_u(query)
        u_old = self.u.clone()
        self.w = None
        for _ in range(self.iter):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            self.w = ops.kmeans_centroids(self.u, query, self.w, mode=ops.CENTROIDS_KL)
            self.u, self.labels = ops.kmeans_assign_kl(query, self.w)
            crit = ops.kmeans_udiff(u_old, self.u)[0]
            u_old = self.u.clone()
            t1.record()
            t1.synchronize()
            dt = t0.elapsed_time(t1) / 1000.0
            self.record_convergence(new_time=dt, criterions=crit)            # kl_kmeans.py:181 (un-normalised) ...
            self.record_convergence(new_time=dt / n_task, criterions=crit)   # ... and :186-187
        self.compute_acc_clustering(query, y_q)


BASE = _KMeansBase

"""Synthetic ``task_dic`` batches with the layout the reference's task generators produce.

The reference builds a batch from cached CLIP features (``src/task_generator_zero_shot.py:49-65``,
``src/task_generator_few_shot.py:83-99``): ``x_q`` float32 [T, n_query, F], ``y_q`` int64 [T, n_query, 1]
(+ ``x_s`` [T, S, F], ``y_s`` [T, S, 1] for few-shot).  Neither CLIP weights nor the datasets exist offline,
so benchmarks and parity tests use the calibrated generator of SURVEY.md §8(d):

  text prototypes  txt = normalize(randn(K, E)),  E = 1024
  per task         k_eff ~ U{3..10} (``src/sampler_zero_shot.py:54``); classes = randperm(K)[:k_eff]
                   y = classes[randint(k_eff, n)];  img = normalize(txt[y] + s * randn(n, E) / sqrt(E)),  s = 9
  softmax feature  z = softmax(T * img @ txt.T),  T = 30 (``src/utils.py:287-290``)

Pure host code (torch CPU); everything is driven by one seeded ``torch.Generator``.  The random draws are float32,
all arithmetic after them (normalisation, the E x K similarity product, the soft-max) runs in float64 and is rounded to
float32 once at the end, so a batch is a bit-reproducible function of the seed across host CPUs (a float32 matmul is
not: its summation order depends on the SIMD width and thread count).
"""
from __future__ import annotations

import math

import torch

EMBED_DIM = 1024
NOISE_SCALE = 9.0


def text_prototypes(K: int, seed: int, embed_dim: int = EMBED_DIM) -> torch.Tensor:
    """Unit-norm class prototypes [K, E]; a fixed function of (K, seed, E)."""
    g = torch.Generator().manual_seed(int(seed) * 7919 + 13)
    t = torch.randn(K, embed_dim, generator=g).double()
    return (t / t.norm(dim=-1, keepdim=True)).float()


def _embed(txt: torch.Tensor, labels: torch.Tensor, g: torch.Generator, noise: float) -> torch.Tensor:
    e = txt.shape[1]
    img = txt[labels].double() + noise * torch.randn(labels.shape[0], e, generator=g).double() / math.sqrt(e)
    return img / img.norm(dim=-1, keepdim=True)      # float64


def _features(img: torch.Tensor, txt: torch.Tensor, temperature: float, softmax_feature: bool) -> torch.Tensor:
    if softmax_feature:
        return (temperature * img @ txt.double().T).softmax(dim=-1).float()
    return img.float()


def make_zero_shot_batch(n_task: int, K: int, n_query: int = 75, seed: int = 0, temperature: float = 30.0,
                         softmax_feature: bool = True, noise: float = NOISE_SCALE, embed_dim: int = EMBED_DIM,
                         k_eff_range: tuple[int, int] = (3, 10), batch_index: int = 0):
    """One ``run_task`` batch.  Returns (task_dic, txt) — txt is what the stub text model returns."""
    txt = text_prototypes(K, seed, embed_dim)
    g = torch.Generator().manual_seed(int(seed) * 1000003 + int(batch_index) * 101 + 1)
    xs, ys = [], []
    lo, hi = k_eff_range
    hi = min(hi, K)
    lo = min(lo, hi)
    for _ in range(n_task):
        k_eff = int(torch.randint(lo, hi + 1, (1,), generator=g))
        classes = torch.randperm(K, generator=g)[:k_eff]
        y = classes[torch.randint(0, k_eff, (n_query,), generator=g)]
        img = _embed(txt, y, g, noise)
        xs.append(_features(img, txt, temperature, softmax_feature))
        ys.append(y)
    task_dic = {"x_q": torch.stack(xs).float().contiguous(),
                "y_q": torch.stack(ys).long().unsqueeze(-1).contiguous()}
    return task_dic, txt


def make_few_shot_batch(n_task: int, K: int, shots: int, n_query: int = 75, k_eff: int = 5, seed: int = 0,
                        temperature: float = 30.0, softmax_feature: bool = True, noise: float = NOISE_SCALE,
                        embed_dim: int = EMBED_DIM, batch_index: int = 0):
    """Few-shot batch: ``shots`` support samples of *every* class (``src/sampler_few_shot.py:64-76``), the
    query drawn from ``k_eff`` classes (``config/main_config.yaml:7``).  S = K * shots."""
    txt = text_prototypes(K, seed, embed_dim)
    g = torch.Generator().manual_seed(int(seed) * 1000003 + int(batch_index) * 101 + 2)
    xq, yq, xsup, ysup = [], [], [], []
    for _ in range(n_task):
        classes = torch.randperm(K, generator=g)[:min(k_eff, K)]
        y = classes[torch.randint(0, classes.numel(), (n_query,), generator=g)]
        xq.append(_features(_embed(txt, y, g, noise), txt, temperature, softmax_feature))
        yq.append(y)
        ys = torch.arange(K).repeat_interleave(shots)
        ys = ys[torch.randperm(ys.numel(), generator=g)]
        xsup.append(_features(_embed(txt, ys, g, noise), txt, temperature, softmax_feature))
        ysup.append(ys)
    task_dic = {"x_q": torch.stack(xq).float().contiguous(), "y_q": torch.stack(yq).long().unsqueeze(-1).contiguous(),
                "x_s": torch.stack(xsup).float().contiguous(), "y_s": torch.stack(ysup).long().unsqueeze(-1).contiguous()}
    return task_dic, txt


def shard_batches(n_batches: int, rank: int, world_size: int) -> list[int]:
    """Whole ``run_task`` batches are the unit of sharding (the MM break test is batch-global,
    ``src/methods/zero_shot/em_dirichlet.py:169-175``): batch i goes to rank i mod W."""
    return [i for i in range(n_batches) if i % world_size == rank]


# ----------------------------------------------------------------------------------------------------------------------
# Device-side task construction from cached features
# ----------------------------------------------------------------------------------------------------------------------
class ZeroShotQuerySampler:
    """The index sampler of the reference's zero-shot evaluator: ``CategoriesSampler_zero_shot`` +
    ``SamplerQuery_zero_shot`` (``src/sampler_zero_shot.py:6-72``), same random calls in the same order
    (``random.randint(3, 10)`` for k_eff, ``torch.randperm`` for the classes and for the samples), so that with the same
    seeds it yields the same index lists.  Iterating yields ``n_batch`` int64 tensors of ``n_query`` indices into the
    cached feature matrix.  ``force_query_size`` re-draws the classes until they hold at least ``n_query`` samples, as the
    evaluator asks (``src/eval_zero_shot.py:153-154``)."""

    def __init__(self, n_batch: int, n_class: int, n_query: int, labels, force_query_size: bool = True):
        self.n_batch, self.n_class, self.n_query = int(n_batch), int(n_class), int(n_query)
        self.force_query_size = force_query_size
        # ``labels``: the label vector of the cached features, or the per-class index lists ``index_lists`` made of it (the
        # evaluator builds a new sampler for every batch; the lists of a fixed label vector need to be made only once)
        self.m_ind_query = labels if isinstance(labels, list) else self.index_lists(labels, self.n_class)

    @staticmethod
    def index_lists(labels, n_class: int) -> list:
        """Per class the indices of its samples, ascending (``src/sampler_zero_shot.py:35-41``)."""
        import numpy as np
        lab = np.asarray(labels.cpu() if isinstance(labels, torch.Tensor) else labels)
        return [torch.from_numpy(np.argwhere(lab == i).reshape(-1)) for i in range(int(n_class))]

    def __len__(self):
        return self.n_batch

    def __iter__(self):
        import random
        for _ in range(self.n_batch):
            k_eff = random.randint(3, 10)
            query_size, n_trials = 0, 0
            while query_size < self.n_query and n_trials < 1:
                classes = torch.randperm(self.n_class)[:k_eff].tolist()
                pool = torch.cat([self.m_ind_query[c] for c in classes])
                query = pool[torch.randperm(len(pool))[:self.n_query]]
                if not self.force_query_size:
                    n_trials += 1
                query_size = len(query)
            yield query


class DeviceTaskSource:
    """Cached features and labels resident on the GPU; ``generate_tasks(sampler)`` turns the sampler's index lists into a
    ``task_dic`` of CUDA tensors with one gather kernel (``tclip_gather_tasks``).  Stands in for the per-task indexing of
    the evaluator plus ``Tasks_Generator_zero_shot.generate_tasks`` (``src/eval_zero_shot.py:158-168``,
    ``src/task_generator_zero_shot.py:36-65``): same tensors, but only T*n indices cross PCIe instead of T*n*F floats.
    ``sampler`` is any iterable of index tensors — this module's ``ZeroShotQuerySampler`` or the reference's own
    ``SamplerQuery_zero_shot``."""

    def __init__(self, all_features, all_labels, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("tclip_b200 runs on B200 GPUs only: device must be a CUDA device (no CPU fallback)")
        self.features = all_features.to(self.device, torch.float32).contiguous()
        self.labels = all_labels.to(self.device).long().contiguous()

    def generate_tasks(self, sampler) -> dict:
        idx = torch.stack([torch.as_tensor(i, dtype=torch.int64) for i in sampler])          # [T, n] on the host
        return self.generate_from_indices(idx)

    def generate_from_indices(self, idx: torch.Tensor) -> dict:
        """``idx`` int64 [T, n] (host): the index lists a sampler produced for one batch."""
        from . import ops
        idx = idx.pin_memory().to(self.device, non_blocking=True)
        x_q, y_q, bad = ops.gather_tasks(self.features, self.labels, idx)
        if int(bad.item()):
            raise IndexError("sampler index outside the cached feature matrix")
        return {"x_q": x_q, "y_q": y_q.unsqueeze(-1)}


class FewShotSamplers:
    """The two index samplers of the reference's few-shot evaluator (``src/sampler_few_shot.py``): ``support()`` mirrors
    ``SamplerSupport_few_shot`` (``s_shot`` samples of every class, one ``torch.randperm`` per class) and ``query()``
    mirrors ``SamplerQuery_few_shot`` (``k_eff`` classes by ``torch.randperm``, then ``n_query`` of their samples).  The
    evaluator iterates the query sampler completely before the support sampler (``src/eval_few_shot.py:233-243``); keep
    that order to reproduce its random stream."""

    def __init__(self, n_batch: int, k_eff: int, n_class: int, s_shot: int, n_query: int, label_support, label_query,
                 force_query_size: bool = True):
        import numpy as np
        self.n_batch, self.k_eff, self.n_class = int(n_batch), int(k_eff), int(n_class)
        self.s_shot, self.n_query, self.force_query_size = int(s_shot), int(n_query), force_query_size
        ls = np.asarray(label_support.cpu() if isinstance(label_support, torch.Tensor) else label_support)
        lq = np.asarray(label_query.cpu() if isinstance(label_query, torch.Tensor) else label_query)
        n = int(ls.max()) + 1
        self.m_ind_support = [torch.from_numpy(np.argwhere(ls == i).reshape(-1)) for i in range(n)]
        self.m_ind_query = [torch.from_numpy(np.argwhere(lq == i).reshape(-1)) for i in range(n)]

    def support(self):
        for _ in range(self.n_batch):
            yield torch.cat([self.m_ind_support[c][torch.randperm(len(self.m_ind_support[c]))[:self.s_shot]]
                             for c in range(self.n_class)])

    def query(self):
        for _ in range(self.n_batch):
            query_size, n_trials = 0, 0
            while query_size < self.n_query and n_trials < 1:
                classes = torch.randperm(self.n_class)[:self.k_eff].tolist()
                pool = torch.cat([self.m_ind_query[c] for c in classes])
                query = pool[torch.randperm(len(pool))[:self.n_query]]
                if not self.force_query_size:
                    n_trials += 1
                query_size = len(query)
            yield query


class DeviceFewShotTaskSource:
    """Few-shot counterpart of ``DeviceTaskSource``: support and query features resident on the GPU; ``generate_tasks``
    applies the sampler's index lists and the per-task relabelling / column re-ordering of
    ``Tasks_Generator_few_shot.get_task`` (``src/task_generator_few_shot.py:27-58``) in two gather kernels.  The task's
    ``unique_labels`` come from the very torch call the reference makes (``torch.flip(torch.unique(labels_support,
    sorted=False), dims=(0,))`` on the host copy of the S support labels), so the class order is the reference's."""

    def __init__(self, features_support, labels_support, features_query, labels_query, device,
                 use_softmax_feature: bool = True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("tclip_b200 runs on B200 GPUs only: device must be a CUDA device (no CPU fallback)")
        self.use_softmax_feature = bool(use_softmax_feature)
        self.labels_support_host = labels_support.cpu().long()
        self.fs = features_support.to(self.device, torch.float32).contiguous()
        self.ls = self.labels_support_host.to(self.device).contiguous()
        self.fq = features_query.to(self.device, torch.float32).contiguous()
        self.lq = labels_query.to(self.device).long().contiguous()
        # label_map must cover every label a gathered sample can carry: a query label that no support sample has maps to
        # 0, as ``zeros_like`` does in the reference's get_task
        self._max_label = max(int(self.labels_support_host.max()), int(labels_query.max()))

    def generate_tasks(self, sampler_support, sampler_query) -> dict:
        from . import ops
        idx_q = torch.stack([torch.as_tensor(i, dtype=torch.int64) for i in sampler_query])      # query sampler first
        idx_s = torch.stack([torch.as_tensor(i, dtype=torch.int64) for i in sampler_support])
        T, F = idx_s.shape[0], self.fs.shape[1]
        if self.use_softmax_feature:
            uniq = [torch.flip(torch.unique(self.labels_support_host[idx_s[t]], sorted=False), dims=(0,)) for t in range(T)]
            if len({int(u.numel()) for u in uniq}) != 1:
                raise ValueError("tasks of one batch must see the same number of support classes")
            col_perm = torch.stack(uniq)                                                         # [T, U]
            n_labels = max(self._max_label, int(col_perm.max())) + 1
            label_map = torch.zeros(T, n_labels, dtype=torch.int64)                              # zeros_like in get_task
            label_map.scatter_(1, col_perm, torch.arange(col_perm.shape[1]).expand(T, -1))
        else:  # visual features: data and labels pass through unchanged
            col_perm = torch.arange(F).expand(T, -1).contiguous()
            n_labels = self._max_label + 1
            label_map = torch.arange(n_labels).expand(T, -1).contiguous()
        dev = self.device
        put = lambda x: x.contiguous().pin_memory().to(dev, non_blocking=True)
        idx_q, idx_s, col_perm, label_map = put(idx_q), put(idx_s), put(col_perm), put(label_map)
        x_s, y_s, bad_s = ops.gather_tasks_remap(self.fs, self.ls, idx_s, col_perm, label_map)
        x_q, y_q, bad_q = ops.gather_tasks_remap(self.fq, self.lq, idx_q, col_perm, label_map)
        if int((bad_s + bad_q).item()):
            raise IndexError("sampler index, class column or label outside the cached features")
        return {"x_s": x_s, "y_s": y_s.unsqueeze(-1), "x_q": x_q, "y_q": y_q.unsqueeze(-1)}

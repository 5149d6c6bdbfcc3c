"""TEST INFRASTRUCTURE — the call sequence of the reference's zero-shot evaluator around ``run_task``, restated
(``src/eval_zero_shot.py:140-187``): for every batch a fresh index sampler, the per-task feature/label slices, the merged
``task_dic``, a NEW method object picked by ``args.name_method``, ``run_task(task_dic=tasks)`` and the 95 % confidence
interval of ``logs['acc'][:, -1]``.  ``tests/test_host_logic.py`` pins it to the live evaluator in the build container (same
seeds -> the same ``task_dic`` tensors and the same results); ``tests/test_gpu_evaluator.py`` drives the drop-in method
classes through it on the GPU box, where the reference checkout does not exist."""
from __future__ import annotations

import importlib

import numpy as np
import torch

from tclip_b200 import tasks

# args.name_method -> (module under src.methods.zero_shot, class name): the dotted names eval_zero_shot.py:12-19 imports
ZERO_SHOT_METHODS = {
    "KL_KMEANS": "kl_kmeans", "EM_DIRICHLET": "em_dirichlet", "HARD_EM_DIRICHLET": "hard_em_dirichlet",
    "EM_GAUSSIAN": "em_gaussian", "EM_GAUSSIAN_COV": "em_gaussian_cov", "SOFT_KMEANS": "soft_kmeans",
    "HARD_KMEANS": "hard_kmeans",
}


def method_builder(name_method: str, package: str = "src.methods.zero_shot"):
    """What ``get_method_builder`` resolves (``src/eval_zero_shot.py:113-138``), through the drop-in module paths."""
    if name_method not in ZERO_SHOT_METHODS:
        raise ValueError("The method your entered does not exist or is not a zero-shot method. Please check the spelling")
    return getattr(importlib.import_module(f"{package}.{ZERO_SHOT_METHODS[name_method]}"), name_method)


def confidence_interval(data, axis=0):
    a = 1.0 * np.array(data)
    m, std = np.mean(a, axis=axis), np.std(a, axis=axis)
    return m, 1.96 * (std / np.sqrt(a.shape[axis]))


def merge_tasks(loader_query):
    """``Tasks_Generator_zero_shot.generate_tasks`` (``src/task_generator_zero_shot.py:36-65``)."""
    n_task = len(loader_query)
    x = torch.cat([d for d, _ in loader_query], dim=0)
    y = torch.cat([lab.long() for _, lab in loader_query], dim=0)
    n_samples = loader_query[0][0].size(0)
    return {"x_q": x.view(n_task, n_samples, x.size(-1)), "y_q": y.view(n_task, n_samples, -1)}


def evaluate_tasks(args, device, all_features_query, all_labels_query, model=None, log_file=None, build=None):
    """Returns (mean accuracy, mean time, per-batch accuracies)."""
    build = build or (lambda **kw: method_builder(args.name_method)(**kw))
    results_task, results_task_time = [], []
    for _ in range(int(args.number_tasks / args.batch_size)):
        sampler_query = tasks.ZeroShotQuerySampler(args.batch_size, args.n_class, args.n_query, all_labels_query,
                                                   force_query_size=True)
        test_loader_query = [(all_features_query[indices, :], all_labels_query[indices]) for indices in sampler_query]
        task_dic = merge_tasks(test_loader_query)
        method = build(model=model, device=device, log_file=log_file, args=args)
        logs = method.run_task(task_dic=task_dic)
        acc_mean, _ = confidence_interval(logs["acc"][:, -1])
        results_task.append(acc_mean)
        results_task_time.append(logs["timestamps"])
        del task_dic
    return np.asarray(results_task).mean(), np.asarray(results_task_time).mean(), results_task

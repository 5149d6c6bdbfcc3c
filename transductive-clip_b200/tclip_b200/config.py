"""The configuration object the method classes read: an attribute-access dict with the same surface as the reference's
``CfgNode`` (``src/utils.py:40-63``), and a helper that fills the keys of SURVEY.md §8(b) for synthetic runs (bench, tests).
With the reference's evaluators the real ``CfgNode`` is passed instead; nothing here is required by the classes."""
from __future__ import annotations


class Cfg(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def make_args(K: int, n_query: int = 75, iters: int = 20, iter_mm: int = 1000, k_eff: int = 5, T: float = 30,
              use_softmax_feature: bool = True, graph_matching: bool = True, **extra) -> Cfg:
    """iter, iter_mm, num_classes_test, n_class, n_query, k_eff, T, use_softmax_feature, graph_matching, classnames,
    template (``config/main_config.yaml``, ``config/methods_config/*.yaml``; ``n_class`` is set by ``main.py:33``)."""
    cfg = Cfg(iter=iters, iter_mm=iter_mm, num_classes_test=K, n_class=K, n_query=n_query, k_eff=k_eff, T=T,
              use_softmax_feature=use_softmax_feature, graph_matching=graph_matching,
              classnames=[f"c{i}" for i in range(K)], template="a photo of a {}.")
    cfg.update(extra)
    return cfg

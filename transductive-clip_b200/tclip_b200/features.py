"""The step right before the hot path: CLIP embeddings -> the cached features the evaluators read.

The reference extracts features once per dataset (``src/utils.py:251-360``): for every image batch
``image_features = normalize(model.encode_image(images))`` and, for softmax features,
``similarity = softmax(T * image_features @ text_features.T)`` with ``text_features = clip_weights(...)`` (unit rows,
``src/utils.py:363-377``); the results are concatenated and pickled as ``{'concat_features', 'concat_labels'}`` under
``data/<dataset>/saved_features/<set>_softmax_<backbone>_T<T>.plk`` / ``<set>_visual_<backbone>.plk``
(``save_pickle`` / ``load_pickle``, ``src/utils.py:241-249``); the evaluators load them back (``src/eval_zero_shot.py:
47-66``).  The CLIP backbone itself is out of scope here; this module is the epilogue after ``encode_image`` (one fused
normalise + similarity + soft-max kernel pair of libtclip_b200) and the cache format, byte-compatible with the reference's
pickles in both directions.
"""
from __future__ import annotations

import os
import pickle

import torch

from . import ops


def softmax_features_path(dataset: str, set_name: str, backbone: str, T, root: str = "data") -> str:
    return os.path.join(root, dataset, "saved_features", "{}_softmax_{}_T{}.plk".format(set_name, backbone, T))


def visual_features_path(dataset: str, set_name: str, backbone: str, root: str = "data") -> str:
    return os.path.join(root, dataset, "saved_features", "{}_visual_{}.plk".format(set_name, backbone))


def visual_features(image_embeddings: torch.Tensor) -> torch.Tensor:
    """``image_features / image_features.norm(dim=-1, keepdim=True)`` (``src/utils.py:343-344``) on the device."""
    return ops.normalize_rows(image_embeddings.float().contiguous())


def softmax_features(image_embeddings: torch.Tensor, text_features: torch.Tensor, T: float) -> torch.Tensor:
    """``softmax(T * normalize(image_embeddings) @ text_features.T)`` (``src/utils.py:286-290``); ``text_features`` are the
    unit-norm rows ``clip_weights`` returns.  CUDA tensors in, CUDA tensor [N, K] out."""
    return ops.kmeans_similarity(visual_features(image_embeddings), text_features.float().contiguous(), float(T))


def extract(encode_image, loader, device, text_features: torch.Tensor | None = None, T: float | None = None) -> dict:
    """The extraction loop of ``extract_features_softmax`` / ``extract_features_visual`` for one temperature:
    ``encode_image(images) -> [B, E]`` is the caller's backbone; returns the dict the reference pickles
    (softmax features on the CPU, visual features where they were computed, labels on the CPU — as upstream)."""
    feats, labels = [], []
    with torch.no_grad():
        for images, lab in loader:
            emb = encode_image(images.to(device)).float()
            if text_features is not None:
                feats.append(softmax_features(emb, text_features.to(device), T).cpu())
            else:
                feats.append(visual_features(emb))
            labels.append(lab.cpu())
    return {"concat_features": torch.cat(feats, dim=0), "concat_labels": torch.cat(labels, dim=0)}


def save_features(path: str, features: torch.Tensor, labels: torch.Tensor) -> None:
    """``save_pickle(path, {'concat_features': ..., 'concat_labels': ...})`` (``src/utils.py:241-243,298-306``)."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump({"concat_features": features, "concat_labels": labels}, f)


def load_features(path: str, device=None):
    """``load_pickle`` + the evaluator's unpacking (``src/eval_zero_shot.py:58-66``): (features float32, labels int64),
    moved to ``device`` when given."""
    with open(path, "rb") as f:
        d = pickle.load(f)
    feats, labels = d["concat_features"], d["concat_labels"].long()
    if device is not None:
        feats, labels = feats.to(device), labels.to(device)
    return feats, labels

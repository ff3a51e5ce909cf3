"""Caller-side glue around the matcher: the helpers every user of the reference calls
right after ``matcher(...)`` -- ``rbd``, ``batch_to_device``, ``match_pair`` (reference utils.py:41-69,
150-165) -- plus ``match_pairs``, the batched form the reference lacks (it matches one pair per call).

No image IO, resizing or extractor lives here (out of scope): ``match_pair`` takes
any extractor object with the reference's ``extract(image, **conf) -> feats`` method.
"""
from __future__ import annotations

from collections.abc import Mapping, Sequence
from typing import Callable

import numpy as np
import torch

from .ragged import match_ragged


def map_tensor(obj, func: Callable):
    """Apply ``func`` to every tensor inside nested dicts / lists; strings and other leaves pass
    through (reference utils.py:41-53; ``collections.abc`` instead of the removed aliases)."""
    if isinstance(obj, (str, bytes)):
        return obj
    if isinstance(obj, torch.Tensor):
        return func(obj)
    if isinstance(obj, Mapping):
        return {k: map_tensor(v, func) for k, v in obj.items()}
    if isinstance(obj, Sequence):
        return [map_tensor(v, func) for v in obj]
    return obj


def batch_to_device(batch: dict, device="cpu", non_blocking: bool = True):
    """Move every tensor of a (nested) batch to ``device`` and detach it (reference utils.py:56-62)."""
    return map_tensor(batch, lambda t: t.to(device=device, non_blocking=non_blocking).detach())


def rbd(data: dict) -> dict:
    """Remove the batch dimension: first element of every tensor / array / list value, other values
    unchanged (reference utils.py:65-70) -- e.g. ``stop`` stays a python int."""
    return {k: v[0] if isinstance(v, (torch.Tensor, np.ndarray, list)) else v for k, v in data.items()}


def match_pair(extractor, matcher, image0: torch.Tensor, image1: torch.Tensor, device="cpu", **preprocess):
    """Extract + match one image pair and return ``(feats0, feats1, matches01)`` without batch dimension
    on ``device`` (reference utils.py:150-165)."""
    feats0 = extractor.extract(image0, **preprocess)
    feats1 = extractor.extract(image1, **preprocess)
    matches01 = matcher({"image0": feats0, "image1": feats1})
    feats0, feats1, matches01 = [batch_to_device(rbd(x), device) for x in (feats0, feats1, matches01)]
    return feats0, feats1, matches01


def match_pairs(extractor, matcher, images0, images1, device="cpu", **preprocess):
    """Batched ``match_pair``: extract every image, then match all pairs -- whatever their keypoint
    counts -- in ONE matcher call (ragged batch, lightglue_b200.ragged).  Returns a list of
    ``(feats0, feats1, matches01)`` triples shaped like ``match_pair``'s."""
    feats0 = [extractor.extract(im, **preprocess) for im in images0]
    feats1 = [extractor.extract(im, **preprocess) for im in images1]
    if len(feats0) != len(feats1):
        raise ValueError("images0 and images1 must have the same length")
    results = match_ragged(matcher, [{"image0": a, "image1": b} for a, b in zip(feats0, feats1)])
    out = []
    for f0, f1, m01 in zip(feats0, feats1, results):
        out.append(tuple(batch_to_device(rbd(x), device) for x in (f0, f1, m01)))
    return out

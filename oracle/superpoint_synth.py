"""Seeded synthetic SuperPoint weights and images (test infrastructure, SURVEY.md 8f1 groundwork).

The official ``superpoint_v1.pth`` cannot be downloaded here (superpoint.py:155-156 fetches it from GitHub), so the
extractor oracle is pinned on synthetic weights with the reference's parameter names: uniform in
+-sqrt(6 / fan_in) (variance preserving through the ReLU stack -- ``nn.Conv2d``'s default bounds are 2.4x smaller
and make a random network's output almost independent of the image), small biases, and the detector's last
layer scaled so that the 65-way soft-max is peaked and the NMS maxima are well separated.  Only uniform draws and exact elementwise fp32 ops are
used, so the tensors are bit-identical on every CPU (same rule as lightglue_b200/synth.py)."""
from __future__ import annotations

import math
from typing import Dict

import torch

LAYERS = (  # name, out channels, in channels, kernel   (superpoint.py:137-153)
    ("conv1a", 64, 1, 3), ("conv1b", 64, 64, 3), ("conv2a", 64, 64, 3), ("conv2b", 64, 64, 3),
    ("conv3a", 128, 64, 3), ("conv3b", 128, 128, 3), ("conv4a", 128, 128, 3), ("conv4b", 128, 128, 3),
    ("convPa", 256, 128, 3), ("convPb", 65, 256, 1), ("convDa", 256, 128, 3), ("convDb", 256, 256, 1),
)


def make_superpoint_state_dict(seed: int = 0, detector_gain: float = 4.0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, co, ci, k in LAYERS:
        bound = math.sqrt(6.0 / (ci * k * k))
        w = (torch.rand(co, ci, k, k, generator=g) * 2.0 - 1.0) * bound
        b = (torch.rand(co, generator=g) * 2.0 - 1.0) * 0.05
        if name == "convPb":
            w = w * detector_gain
        sd[f"{name}.weight"], sd[f"{name}.bias"] = w, b
    return sd


def make_image(h: int, w: int, b: int = 1, seed: int = 0) -> torch.Tensor:
    """[b, 1, h, w] in [0, 1]: random 8x8 blocks plus per-pixel noise (exact fp32 arithmetic only)."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand(b, 1, (h + 7) // 8, (w + 7) // 8, generator=g)  # (extents that are not multiples of 8: cropped blocks)
    fine = torch.rand(b, 1, h, w, generator=g)
    return coarse.repeat_interleave(8, 2).repeat_interleave(8, 3)[:, :, :h, :w] * 0.7 + fine * 0.3

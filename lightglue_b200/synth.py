"""Seeded synthetic keypoints / descriptors / weights for parity tests and benchmarks.

There is no network in this environment, so the official ``*_lightglue.pth`` checkpoints
(reference lightglue/lightglue.py:348-349, 416-421) cannot be fetched.  Everything is therefore
measured on synthetic data of the reference's shapes:

* inputs  : SuperPoint/DISK-shaped pairs -- pixel keypoints, unit-norm descriptors, image1 a
  permuted + perturbed copy of image0 so that a ground-truth assignment exists;
* weights : a state_dict with the *reference key names* (SURVEY.md §8a "Parameter inventory"),
  drawn from the same distributions ``nn.Linear``'s default init uses, then re-scaled so that
  matching is non-degenerate (plain random init yields 0 matches and never stops/prunes).

Everything here is generated from an explicit ``torch.Generator`` on the CPU using only operations
that are bit-reproducible across CPUs (uniform draws, exact fp64 sums, correctly-rounded
elementwise IEEE ops): no ``randn`` (its vectorised Box-Muller differs in the last ulp between
AVX2/AVX-512 hosts) and no fp32 reductions (``normalize``).  The same seed therefore gives
bit-identical tensors in the build container, in the golden-fixture generator and on the GPU box.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

N_LAYERS = 9
DIM = 256


def _gauss(g: torch.Generator, *shape: int) -> torch.Tensor:
    """~N(0,1) as an Irwin-Hall sum of 12 uniforms, fp64.  Each uniform is a 24-bit fp32 value, so the
    fp64 sum is exact (order independent); the result is snapped to a 2^-12 grid."""
    acc = torch.zeros(*shape, dtype=torch.float64)
    for _ in range(12):
        acc += torch.rand(*shape, generator=g).double()
    return torch.round((acc - 6.0) * 4096.0) / 4096.0


def _unit_rows(v: torch.Tensor) -> torch.Tensor:
    """Row-normalise fp64 values that lie on a 2^-20 grid: the sum of squares is exact in fp64
    (<= 48 significant bits), sqrt and the division are correctly rounded -> reproducible fp32 rows."""
    v = torch.round(v * 1048576.0) / 1048576.0
    nrm = (v * v).sum(-1, keepdim=True).sqrt()
    return (v / nrm).float()


def make_pair(
    n: int,
    d: int = 256,
    b: int = 1,
    seed: int = 1000,
    m: int | None = None,
    noise: float = 0.05,
    w: float = 1024.0,
    h: float = 768.0,
    scale_ori: bool = False,
) -> Tuple[dict, torch.Tensor]:
    """One batch of ``b`` synthetic pairs.  image0 has ``m`` (default ``n``) keypoints, image1 has ``n``.

    Returns ``(data, perm)`` where ``data`` is the dict the matcher takes
    (lightglue.py:456-468) and ``perm[b, j]`` is the image0 index image1's keypoint j was copied
    from (ground truth: ``matches1 == perm`` where matched).
    """
    m = n if m is None else m
    big = max(m, n)
    g = torch.Generator().manual_seed(int(seed))
    k_all = torch.rand(b, big, 2, generator=g) * torch.tensor([w, h])
    d_all = _unit_rows(_gauss(g, b, big, d))
    perm = torch.stack([torch.randperm(big, generator=g) for _ in range(b)])
    k1 = torch.gather(k_all, 1, perm[..., None].expand(-1, -1, 2))
    k1 = (k1.double() + 2.0 * _gauss(g, b, big, 2)).float()
    d1 = torch.gather(d_all, 1, perm[..., None].expand(-1, -1, d))
    d1 = _unit_rows(d1.double() + noise * _gauss(g, b, big, d))
    size = torch.tensor([[w, h]]).expand(b, 2).contiguous()
    f0 = {"keypoints": k_all[:, :m].contiguous(), "descriptors": d_all[:, :m].contiguous(), "image_size": size}
    f1 = {"keypoints": k1[:, :n].contiguous(), "descriptors": d1[:, :n].contiguous(), "image_size": size.clone()}
    if scale_ori:
        for f, cnt in ((f0, m), (f1, n)):
            f["scales"] = 1.0 + 4.0 * torch.rand(b, cnt, generator=g)
            f["oris"] = (torch.rand(b, cnt, generator=g) * 2.0 - 1.0) * math.pi
    return {"image0": f0, "image1": f1}, perm[:, :n]


def _linear(g: torch.Generator, out_f: int, in_f: int) -> Tuple[torch.Tensor, torch.Tensor]:
    # nn.Linear default init == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias
    bound = 1.0 / math.sqrt(in_f)
    wt = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
    bs = (torch.rand(out_f, generator=g) * 2 - 1) * bound
    return wt, bs


def make_state_dict(
    seed: int = 0,
    input_dim: int = 256,
    adaptive: bool = False,
    add_scale_ori: bool = False,
    n_layers: int = N_LAYERS,
) -> Dict[str, torch.Tensor]:
    """A synthetic LightGlue state_dict under the reference's key names.

    Recipe: default-init distributions, then
    ``final_proj.weight *= 2``, ``matchability.bias = 3`` (non-adaptive) so that a few hundred
    matches straddle ``filter_threshold``; the adaptive variant additionally sharpens the
    token-confidence and matchability heads so early-exit and point pruning both fire.
    """
    g = torch.Generator().manual_seed(int(seed))
    sd: Dict[str, torch.Tensor] = {}
    d = DIM
    if input_dim != d:
        sd["input_proj.weight"], sd["input_proj.bias"] = _linear(g, d, input_dim)
    pos_dim = 4 if add_scale_ori else 2
    sd["posenc.Wr.weight"] = _gauss(g, 32, pos_dim).float()  # normal(0, gamma**-2), gamma = 1
    for i in range(n_layers):
        p = f"transformers.{i}.self_attn."
        sd[p + "Wqkv.weight"], sd[p + "Wqkv.bias"] = _linear(g, 3 * d, d)
        sd[p + "out_proj.weight"], sd[p + "out_proj.bias"] = _linear(g, d, d)
        sd[p + "ffn.0.weight"], sd[p + "ffn.0.bias"] = _linear(g, 2 * d, 2 * d)
        sd[p + "ffn.1.weight"] = (1.0 + 0.1 * _gauss(g, 2 * d)).float()
        sd[p + "ffn.1.bias"] = (0.05 * _gauss(g, 2 * d)).float()
        sd[p + "ffn.3.weight"], sd[p + "ffn.3.bias"] = _linear(g, d, 2 * d)
        p = f"transformers.{i}.cross_attn."
        sd[p + "to_qk.weight"], sd[p + "to_qk.bias"] = _linear(g, d, d)
        sd[p + "to_v.weight"], sd[p + "to_v.bias"] = _linear(g, d, d)
        sd[p + "to_out.weight"], sd[p + "to_out.bias"] = _linear(g, d, d)
        sd[p + "ffn.0.weight"], sd[p + "ffn.0.bias"] = _linear(g, 2 * d, 2 * d)
        sd[p + "ffn.1.weight"] = (1.0 + 0.1 * _gauss(g, 2 * d)).float()
        sd[p + "ffn.1.bias"] = (0.05 * _gauss(g, 2 * d)).float()
        sd[p + "ffn.3.weight"], sd[p + "ffn.3.bias"] = _linear(g, d, 2 * d)
    for i in range(n_layers):
        p = f"log_assignment.{i}."
        sd[p + "matchability.weight"], sd[p + "matchability.bias"] = _linear(g, 1, d)
        sd[p + "final_proj.weight"], sd[p + "final_proj.bias"] = _linear(g, d, d)
        sd[p + "final_proj.weight"] *= 2.0
        if adaptive:
            sd[p + "matchability.weight"] *= 16.0
            sd[p + "matchability.bias"].fill_(2.0)
        else:
            sd[p + "matchability.bias"].fill_(3.0)
    for i in range(n_layers - 1):
        p = f"token_confidence.{i}.token.0."
        sd[p + "weight"], sd[p + "bias"] = _linear(g, 1, d)
        if adaptive:
            sd[p + "weight"] *= 8.0
            sd[p + "bias"].fill_(0.8 * i)
    return sd


def checksum(t: torch.Tensor) -> int:
    """Order-sensitive, exactly reproducible checksum (int64 arithmetic on the fp32 bit patterns) used
    to pin regenerated tensors to the golden fixtures."""
    bits = t.detach().to(torch.float32).contiguous().flatten().view(torch.int32).to(torch.int64)
    wts = torch.arange(1, bits.numel() + 1, dtype=torch.int64) % 977 + 1
    return int((bits * wts).sum())

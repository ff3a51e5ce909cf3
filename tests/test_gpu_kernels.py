"""Kernel-level numerics: the attention kernel of every precision mode (through the C-ABI entry point
``lg_attention``) against a plain PyTorch fp32 reference of the same op -- softmax(q k^T / sqrt(64)) v per
(pair, head), self and cross (reference lightglue.py:97-137, 170, 210-214).

The tensor-core modes compute on fp16-rounded q, k, v (as the reference's flash path does, 116-121), so their
reference uses the same rounded operands in fp32 arithmetic; what remains is the kernel's own error: P rounded
to fp16 before P.V, and the bf16 (bf16: hi only, bf16x3: hi + lo) output."""
import pytest
import torch

from lightglue_b200 import LightGlue

pytestmark = pytest.mark.gpu


def torch_reference(q, k, v):
    """[B, H, Lq, 64] x [B, H, Lk, 64] -> [B, Lq, H*64], fp32 on the GPU, no fused kernels."""
    s = torch.einsum("bhid,bhjd->bhij", q.double(), k.double()) * (64 ** -0.5)
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("bhij,bhjd->bhid", p, v.double())
    return o.transpose(1, 2).flatten(-2).float()


@pytest.mark.parametrize("precision,tol", [("fp32", 5e-5), ("bf16x3", 3e-3), ("bf16", 2e-2)])
@pytest.mark.parametrize("cross", [False, True])
@pytest.mark.parametrize("m,n", [(300, 517), (2048, 2048), (64, 1)])
def test_attention_kernel_vs_torch_fp32(precision, tol, cross, m, n):
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n)
    b = 2
    mk = lambda l, scale: torch.randn(b, 4, l, 64, device="cuda", generator=g) * scale  # noqa: E731
    # logits with a spread of ~ +-12 so that the soft-max is peaked like the trained model's, values O(1)
    q0, k0, v0, q1, k1, v1 = mk(m, 2.0), mk(m, 2.0), mk(m, 1.0), mk(n, 2.0), mk(n, 2.0), mk(n, 1.0)
    mod = LightGlue(features=None, precision=precision, depth_confidence=-1, width_confidence=-1).eval().cuda()
    c0, c1 = mod.attention(q0, k0, v0, q1, k1, v1, cross=cross)
    if precision != "fp32":  # the operands the kernel actually multiplies
        q0, k0, v0, q1, k1, v1 = [t.half().float() for t in (q0, k0, v0, q1, k1, v1)]
    r0 = torch_reference(q0, k1 if cross else k0, v1 if cross else v0)
    r1 = torch_reference(q1, k0 if cross else k1, v0 if cross else v1)
    e0, e1 = float((c0 - r0).abs().max()), float((c1 - r1).abs().max())
    print(f"[{precision}] cross={cross} m={m} n={n}: max|err| {e0:.2e} {e1:.2e}")
    assert c0.shape == (b, m, 256) and c1.shape == (b, n, 256)
    assert torch.isfinite(c0).all() and torch.isfinite(c1).all()
    assert e0 <= tol and e1 <= tol

"""The persistent one-CTA-per-SM attention kernel (double-buffered S; picked automatically for large batches) forced onto
small problems with LG_ATTN_V=3: kernel-level numerics against torch, and the reference-generated fixtures end to end
(dense, ragged, adaptive: work items that are skipped, single-tile items, sequences pruned between layers).
Reference: lightglue.py:113-137 (Attention.forward), 170, 210-214."""
import pytest
import torch

from lightglue_b200 import LightGlue
from tests.helpers import compare_outputs, load_case
from tests.test_gpu_kernels import torch_reference
from tests.test_gpu_parity import TC_ADAPTIVE_CASES, TC_CASES, build, to_cuda

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def force_persistent(monkeypatch):
    monkeypatch.setenv("LG_ATTN_V", "3")


@pytest.mark.parametrize("cross", [False, True])
@pytest.mark.parametrize("b,m,n", [(2, 300, 517), (2, 2048, 2048), (2, 64, 1), (3, 1000, 130), (40, 256, 384)])
def test_persistent_attention_kernel_vs_torch(cross, b, m, n):
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n)
    mk = lambda l, scale: torch.randn(b, 4, l, 64, device="cuda", generator=g) * scale  # noqa: E731
    q0, k0, v0, q1, k1, v1 = mk(m, 2.0), mk(m, 2.0), mk(m, 1.0), mk(n, 2.0), mk(n, 2.0), mk(n, 1.0)
    mod = LightGlue(features=None, precision="bf16x3", depth_confidence=-1, width_confidence=-1).eval().cuda()
    c0, c1 = mod.attention(q0, k0, v0, q1, k1, v1, cross=cross)
    q0, k0, v0, q1, k1, v1 = [t.half().float() for t in (q0, k0, v0, q1, k1, v1)]
    r0 = torch_reference(q0, k1 if cross else k0, v1 if cross else v0)
    r1 = torch_reference(q1, k0 if cross else k1, v0 if cross else v1)
    e0, e1 = float((c0 - r0).abs().max()), float((c1 - r1).abs().max())
    print(f"persistent cross={cross} b={b} m={m} n={n}: max|err| {e0:.2e} {e1:.2e}")
    assert mod.debug_timeout_code() == 0
    assert torch.isfinite(c0).all() and torch.isfinite(c1).all()
    assert e0 <= 3e-3 and e1 <= 3e-3  # tolerance of the bf16x3 mode in test_gpu_kernels.py


@pytest.mark.parametrize("name", TC_CASES)
def test_persistent_attention_fixtures_index_exact(name):
    fix, data, sd = load_case(name)
    m = build(fix, sd, "bf16x3")
    out = m(to_cuda(data))
    flips, dmax = compare_outputs(out, fix["out"], score_tol=1e-3)
    assert m.debug_timeout_code() == 0
    print(f"[persistent attention] {name}: flips={flips} max|dscore|={dmax:.2e}")


@pytest.mark.parametrize("name", TC_ADAPTIVE_CASES)
def test_persistent_attention_adaptive_identical_decisions(name):
    fix, data, sd = load_case(name)
    gold = fix["out"]
    m = build(fix, sd, "bf16x3")
    out = m(to_cuda(data))
    assert m.debug_timeout_code() == 0
    assert int(out["stop"]) == int(gold["stop"])
    assert torch.equal(out["prune0"].cpu().double(), gold["prune0"].double())
    assert torch.equal(out["prune1"].cpu().double(), gold["prune1"].double())
    compare_outputs(out, gold, score_tol=1e-3)

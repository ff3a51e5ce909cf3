"""Time the attention kernel alone through the kernel-level entry point (lg_attention) at the bench shape.

    python tools/attn_bench.py [--batch 32] [--n 2048] [--precision bf16] [--cross]
    ncu --set full --import-source on --clock-control none -k regex:tc_attention --launch-skip 2 --launch-count 1 \
        -o gpurun_out/attn python tools/attn_bench.py --iters 3

Prints microseconds per launch (device time of the attention class, CUDA events around the kernel inside the
library) and the fraction of the measured sustained bf16 peak (MEASURED_PEAKS.json) for 4*Nq*Nk*64 FLOP per
(sequence, head)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightglue_b200 import LightGlue  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--cross", action="store_true")
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()

m = LightGlue(features=None, precision=a.precision, depth_confidence=-1, width_confidence=-1).eval().cuda()
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda s: torch.randn(a.batch, 4, a.n, 64, device="cuda", generator=g) * s  # noqa: E731
q0, k0, v0, q1, k1, v1 = mk(2.0), mk(2.0), mk(1.0), mk(2.0), mk(2.0), mk(1.0)
for _ in range(2):
    m.attention(q0, k0, v0, q1, k1, v1, cross=a.cross)
torch.cuda.synchronize()
m.timing = True
for _ in range(a.iters):
    m.attention(q0, k0, v0, q1, k1, v1, cross=a.cross)
torch.cuda.synchronize()
ms, cnt = m.kernel_times()["attention"]
us = ms / cnt * 1e3
flops = 4.0 * a.n * a.n * 64 * 4 * 2 * a.batch
peak = 1400.0
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    d = json.load(open(p))
    peak = d.get("bf16_tflops_sustained", d.get("bf16_tflops", peak))
print(f"attention {a.precision} B={a.batch} N={a.n} cross={a.cross}: {us:.1f} us/launch, "
      f"{flops / (us * 1e-6) / 1e12:.0f} TFLOP/s = {flops / (us * 1e-6) / 1e12 / peak:.3f} of {peak:.0f}")

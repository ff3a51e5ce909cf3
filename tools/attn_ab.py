"""A/B of attention-kernel switches through lg_attention on one box: correctness against torch (fp64 softmax on the
fp16-rounded operands) on a few shapes, then device time per launch at the bench shape, once per environment set.

    python tools/attn_ab.py --env-sets "|LG_ATTN_NO_PINGPONG=1"      # default build vs no exponential-phase token

(The logs under profiles/r2_y_attention_*.log were produced by an earlier form of this tool that also selected the
experimental persistent kernel of commit 8fb04ec: `v2` there is the shipped kernel, `v3:<variant>:<no token>` the others.)"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightglue_b200 import LightGlue  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--env-sets", default="", help="'|'-separated sets of ';'-separated NAME=VALUE assignments; '' = none")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--skip-check", action="store_true")
a = ap.parse_args()


def ref(q, k, v):
    s = torch.einsum("bhid,bhjd->bhij", q.double(), k.double()) * (64 ** -0.5)
    return torch.einsum("bhij,bhjd->bhid", torch.softmax(s, -1), v.double()).transpose(1, 2).flatten(-2).float()


m = LightGlue(features=None, precision=a.precision, depth_confidence=-1, width_confidence=-1).eval().cuda()
touched = set()
for env_set in a.env_sets.split("|"):
    for name in touched:
        os.environ.pop(name, None)
    for kv in filter(None, env_set.split(";")):
        name, val = kv.split("=", 1)
        os.environ[name] = val
        touched.add(name)
    tag = env_set or "default"
    if not a.skip_check:
        for (b, mm, nn) in [(2, 300, 517), (2, 2048, 2048), (2, 64, 1), (3, 1000, 130)]:
            g = torch.Generator(device="cuda").manual_seed(mm * 7 + nn)
            mk = lambda l, sc: torch.randn(b, 4, l, 64, device="cuda", generator=g) * sc  # noqa: E731
            t = [mk(mm, 2.0), mk(mm, 2.0), mk(mm, 1.0), mk(nn, 2.0), mk(nn, 2.0), mk(nn, 1.0)]
            for cross in (False, True):
                c0, c1 = m.attention(*t, cross=cross)
                torch.cuda.synchronize()
                q0, k0, v0, q1, k1, v1 = [x.half().float() for x in t]
                r0 = ref(q0, k1 if cross else k0, v1 if cross else v0)
                r1 = ref(q1, k0 if cross else k1, v0 if cross else v1)
                e0, e1 = float((c0 - r0).abs().max()), float((c1 - r1).abs().max())
                ok = e0 < 3e-3 and e1 < 3e-3 and bool(torch.isfinite(c0).all()) and bool(torch.isfinite(c1).all())
                print(f"[{tag}] check b={b} m={mm} n={nn} cross={cross}: {e0:.2e} {e1:.2e} {'ok' if ok else 'FAIL'} "
                      f"timeout={m.debug_timeout_code()}", flush=True)
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda s: torch.randn(a.batch, 4, a.n, 64, device="cuda", generator=g) * s  # noqa: E731
    t = [mk(2.0), mk(2.0), mk(1.0), mk(2.0), mk(2.0), mk(1.0)]
    for cross in (False, True):
        for _ in range(3):
            m.attention(*t, cross=cross)
        torch.cuda.synchronize()
        m.timing = True
        for _ in range(a.iters):
            m.attention(*t, cross=cross)
        torch.cuda.synchronize()
        ms, cnt = m.kernel_times()["attention"]
        m.timing = False
        us = ms / cnt * 1e3
        flops = 4.0 * a.n * a.n * 64 * 4 * 2 * a.batch
        print(f"[{tag}] B={a.batch} N={a.n} cross={cross}: {us:.1f} us/launch, {flops / us / 1e6:.0f} TFLOP/s "
              f"= {flops / us / 1e6 / 1449.7:.3f} of the sustained bf16 peak", flush=True)

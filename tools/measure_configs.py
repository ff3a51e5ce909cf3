#!/usr/bin/env python
"""Throughput / latency of the BASELINE.json configs other than the bench headline (configs[1]).
Prints one JSON object; numbers go into DESIGN.md / profiles.  CUDA-event timed, 5 warm-up + 20 timed forwards.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightglue_b200 import LightGlue, synth  # noqa: E402

torch.set_grad_enabled(False)


def timed(m, data, reps=20, warm=5):
    for _ in range(warm):
        out = m(data)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = m(data)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def cuda(d):
    return {k: {kk: vv.cuda() for kk, vv in v.items()} for k, v in d.items()}


res = {}
# config 0/1-like: N=512 single pair (reference parity shape), N=2048 batches
for prec in ("bf16x3", "bf16"):
    sd = synth.make_state_dict()
    m = LightGlue(features=None, depth_confidence=-1, width_confidence=-1, precision=prec)
    m.load_state_dict(sd, strict=False)
    m = m.cuda()
    for (b, n) in ((1, 512), (1, 2048), (8, 2048), (32, 2048)):
        if prec == "fp32" and b > 8:
            continue
        base, _ = synth.make_pair(n, b=min(b, 4), seed=7)
        data = {k: {kk: vv.repeat((b + 3) // 4 if b > 4 else 1, *([1] * (vv.dim() - 1)))[:b] for kk, vv in v.items()}
                for k, v in base.items()}
        ms, _ = timed(m, cuda(data), reps=10 if prec == "fp32" else 20)
        res[f"{prec}_B{b}_N{n}"] = {"ms_per_forward": ms, "pairs_per_s": b * 1000.0 / ms}
# config 2: adaptive (depth 0.95 / width 0.99), N=2048, one pair per forward and batched per-pair adaptivity
sd = synth.make_state_dict(adaptive=True)
for prec in ("bf16x3", "bf16"):
    for b in (1, 32):
        m = LightGlue(features=None, precision=prec)
        m.load_state_dict(sd, strict=False)
        m = m.cuda()
        base, _ = synth.make_pair(2048, b=min(b, 4), seed=9)
        data = {k: {kk: vv.repeat((b + 3) // 4 if b > 4 else 1, *([1] * (vv.dim() - 1)))[:b] for kk, vv in v.items()} for k, v in base.items()}
        ms, out = timed(m, cuda(data))
        kept = float((out["prune0"].float() >= out["stop"]).float().mean())
        res[f"adaptive_{prec}_B{b}_N2048"] = {"ms_per_forward": ms, "pairs_per_s": b * 1000.0 / ms, "stop": int(out["stop"]),
                                              "fraction_never_pruned": kept}
# config 3: DISK d=128, N=4096, B=16
sd = synth.make_state_dict(input_dim=128)
base, _ = synth.make_pair(4096, d=128, b=4, seed=11)
data = {k: {kk: vv.repeat(4, *([1] * (vv.dim() - 1))) for kk, vv in v.items()} for k, v in base.items()}
for prec in ("bf16x3", "bf16"):
    m = LightGlue(features=None, input_dim=128, depth_confidence=-1, width_confidence=-1, precision=prec)
    m.load_state_dict(sd, strict=False)
    m = m.cuda()
    ms, out = timed(m, cuda(data), reps=10)
    res[f"disk_{prec}_B16_N4096"] = {"ms_per_forward": ms, "pairs_per_s": 16 * 1000.0 / ms,
                                     "matches_per_pair": float((out["matches0"] > -1).float().sum(1).mean())}
print(json.dumps(res, indent=1))

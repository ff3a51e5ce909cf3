#!/usr/bin/env python
"""Quick on-GPU check of the tensor-core path against the golden fixtures: index flips, score error, the
in-kernel wait-timeout code and per-kernel-class times.  `python tools/tc_check.py [precision ...]`."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lightglue_b200 import LightGlue, synth  # noqa: E402
from tests.helpers import load_case  # noqa: E402

torch.set_grad_enabled(False)
precs = sys.argv[1:] or ["bf16x3", "bf16"]
bad = 0
for prec in precs:
    for name in ("c1_n512", "n2048", "ragged_b2", "disk_d128", "adaptive_n512"):
        fix, data, sd = load_case(name)
        rc, conf = fix["recipe"], fix["conf"]
        m = LightGlue(features=None, input_dim=rc["d"], add_scale_ori=rc.get("scale_ori", False), precision=prec, **conf)
        m.load_state_dict(sd, strict=False)
        m = m.eval().cuda()
        m.pruning_keypoint_thresholds = dict(LightGlue.pruning_keypoint_thresholds, flash=rc.get("pruning_threshold", -1))
        cd = {k: {kk: vv.cuda() for kk, vv in v.items()} for k, v in data.items()}
        t0 = time.time()
        out = m(cd)
        torch.cuda.synchronize()
        code = m.debug_timeout_code()
        gold = fix["out"]
        flips = int((out["matches0"].cpu() != gold["matches0"]).sum()) + int((out["matches1"].cpu() != gold["matches1"]).sum())
        ds = float((out["matching_scores0"].cpu() - gold["matching_scores0"]).abs().max())
        nan = bool(torch.isnan(out["matching_scores0"]).any())
        ok = code == 0 and not nan and (flips == 0 if prec != "bf16" else flips < 40) and int(out["stop"]) == int(gold["stop"])
        bad += not ok
        print(f"[{prec}] {name:16s} flips={flips:4d} max|dscore|={ds:.2e} stop={out['stop']}/{gold['stop']} timeout_code={code:#x} "
              f"nan={nan} {'ok' if ok else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
        if code:
            print("   debug words:", [hex(w) for w in m.debug_words if w])
print("tc_check:", "all ok" if bad == 0 else f"{bad} FAILED")
sys.exit(1 if bad else 0)

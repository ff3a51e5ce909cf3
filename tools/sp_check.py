#!/usr/bin/env python
"""On-GPU check of the SuperPoint extractor against the reference-generated fixtures, both arithmetic modes, plus timing.
`python tools/sp_check.py`"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lightglue_b200 import synth  # noqa: E402
from lightglue_b200.superpoint import SuperPoint  # noqa: E402
from oracle import superpoint_synth as sps  # noqa: E402

torch.set_grad_enabled(False)
GOLDEN = os.path.join(ROOT, "tests", "golden")
bad = 0
for prec in ("fp32", "bf16x3"):
    for name in ("sp_240x320", "sp_b2_top256", "sp_nms2_thr01", "sp_480x640_top512"):
        fix = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
        rc, conf, gold = fix["recipe"], fix["conf"], fix["out"]
        image = sps.make_image(rc["h"], rc["w"], rc["b"], rc["seed"])
        m = SuperPoint(weights=None, precision=prec, **conf)
        m.load_state_dict(sps.make_superpoint_state_dict(0))
        m = m.eval().cuda()
        try:
            out = m({"image": image.cuda()})
        except Exception as exc:  # noqa: BLE001
            print(f"[{prec}] {name}: FAILED {exc!r}"[:300])
            bad += 1
            continue
        torch.cuda.synchronize()
        same, ds, dd, nk, order = True, 0.0, 0.0, 0, True
        for b in range(rc["b"]):
            k, g = out["keypoints"][b].cpu(), gold["keypoints"][b]
            nk += g.shape[0]
            if k.shape != g.shape or not torch.equal(k, g):
                order = False
                ks = {tuple(x) for x in k.tolist()}
                gs = {tuple(x) for x in g.tolist()}
                same = same and ks == gs
                print(f"   image {b}: {len(ks ^ gs)} keypoints differ as sets; shape {tuple(k.shape)} vs {tuple(g.shape)}")
                continue
            ds = max(ds, float((out["keypoint_scores"][b].cpu() - gold["keypoint_scores"][b]).abs().max()))
            st = gold["desc_stride"][b]
            dd = max(dd, float((out["descriptors"][b].cpu()[::st] - gold["descriptors"][b]).abs().max()))
        ok = order and ds <= 2e-5 and dd <= 2e-5
        bad += not ok
        print(f"[{prec}] {name:20s} kpts={nk} identical_order={order} same_set={same} max|dscore|={ds:.2e} max|ddesc|={dd:.2e} "
              f"{'ok' if ok else 'DIFF'}", flush=True)
    # timing at the reference's default extraction size (resize 1024): 768 x 1024
    image = sps.make_image(768, 1024, 1, 7).cuda()
    m = SuperPoint(weights=None, precision=prec, max_num_keypoints=2048)
    m.load_state_dict(sps.make_superpoint_state_dict(0))
    m = m.eval().cuda()
    for _ in range(2):
        m({"image": image})
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(5):
        m({"image": image})
    torch.cuda.synchronize()
    print(f"[{prec}] 768x1024, top-2048: {(time.time() - t0) / 5 * 1e3:.2f} ms per image", flush=True)
print("sp_check:", "all ok" if bad == 0 else f"{bad} with differences")

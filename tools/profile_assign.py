"""Drive the materialising assignment stage alone (lg_assign -> final_proj, LSE sweep, arg-max sweep that
also writes the [B, M+1, N+1] log-assignment matrix) so ncu can capture it:

  ncu --set full --import-source on --clock-control none -k regex:tc_linear --launch-skip 3 --launch-count 3 \
      -o gpurun_out/assign python tools/profile_assign.py
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightglue_b200 import LightGlue, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--time", action="store_true")
ap.add_argument("--precision", default="bf16x3")
a = ap.parse_args()

m = LightGlue(features=None, precision=a.precision, depth_confidence=-1, width_confidence=-1)
m.load_state_dict(synth.make_state_dict(), strict=False)
m = m.eval().cuda()
g = torch.Generator(device="cuda").manual_seed(0)
x0 = torch.randn(a.batch, a.n, 256, device="cuda", generator=g)
x1 = torch.randn(a.batch, a.n, 256, device="cuda", generator=g)
for _ in range(a.iters):
    full, *_ = m.log_assignment_matrix(8, x0, x1)
torch.cuda.synchronize()
if a.time:
    m.timing = True
    for _ in range(5):
        m.log_assignment_matrix(8, x0, x1)
    torch.cuda.synchronize()
    t = m.kernel_times()
    ms, cnt = t["assign_matrix"]
    byt = a.batch * ((2 * a.n) * 256 * 2 + 256 * 256 * 2 + (a.n + 1) ** 2 * 4 + 2 * a.n * 12)
    print(f"assign_matrix: {ms / cnt * 1e3:.1f} us/launch, {byt / (ms / cnt * 1e-3) / 1e9:.0f} GB/s algorithmic")
print("ok", tuple(full.shape))

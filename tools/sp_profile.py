"""Two SuperPoint forwards at 768 x 1024 (for `ncu --metrics gpu__time_duration.sum`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lightglue_b200.superpoint import SuperPoint  # noqa: E402

torch.set_grad_enabled(False)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
g = torch.Generator().manual_seed(11)
im = torch.rand(1, 1, 768, 1024, generator=g).cuda()
sp = SuperPoint(weights=None, max_num_keypoints=2048, precision=prec).eval().cuda()
for _ in range(2):
    sp({"image": im})
torch.cuda.synchronize()
print("done")

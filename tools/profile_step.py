"""Two denoising steps at the bench configuration (200 f x 64x64) for ncu captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
from dawn_pytorch_b200 import DynamicNfUnet3D  # noqa: E402

torch.manual_seed(0)
net = DynamicNfUnet3D(**bench.CTOR).eval().cuda()
x_t, fea, cond = bench.synth_clip(1)
net.update_num_frames(bench.F_CLIP)
net.set_clip_invariants(fea.cuda(), cond.cuda())
t = torch.full((1,), 500, dtype=torch.long, device="cuda")
out = torch.empty((3, bench.F_CLIP, bench.H_LAT, bench.W_LAT), device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    net.forward_x3(x_t.cuda(), t, out)
torch.cuda.synchronize()
print("done", float(out.abs().max()), "launches/step", net.last_launch_count())

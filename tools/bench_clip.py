"""Whole-clip pipeline timing (round-2 tool, not the driver's bench): source encoder + 20-step DDIM over the CUDA UNet (eager or
one CUDA graph) + batched LFG decode for ONE clip at the BASELINE configs[2] shape (200 frames, 256x256 video, 64x64 latent).

  python tools/bench_clip.py [--frames 200] [--steps 20] [--graph] [--chunk 50] [--clips 3]
Prints stage times (CUDA events) and frames/s.  Synthetic weights and inputs; the decode runs in chunks of --chunk frames."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dawn_pytorch_b200 import FlowDiffusion          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=50)
    ap.add_argument("--clips", type=int, default=3)
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    torch.manual_seed(0)
    m = FlowDiffusion(sampling_timesteps=a.steps, pose_dim=6, win_width=40).cuda()
    m.update_num_frames(a.frames)
    F, S = a.frames, a.size
    img = torch.rand(1, 3, S, S, device="cuda")
    hubert = torch.randn(1, F, 1024, device="cuda")
    pose = torch.randn(1, 7, F, device="cuda") * 0.2
    eye = torch.rand(1, 2, F, device="cuda")
    bbox = torch.tensor([[0.3 * S, 0.7 * S, 0.25 * S, 0.8 * S, S, S]], device="cuda").unsqueeze(-1).repeat(1, 1, F)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for clip in range(a.clips):
        ev[0].record()
        fea = m.generator.compute_fea(img)
        mask = m.face_loc_emb(m.generate_bbox_mask(bbox, size=S))
        cond = torch.cat([hubert, pose[:, :6].permute(0, 2, 1) - pose[:, :6, :1].permute(0, 2, 1), eye.permute(0, 2, 1) - eye[:, :, :1].permute(0, 2, 1)], dim=-1)
        ev[1].record()
        pred = m.diffusion.ddim_sample(torch.cat([fea, mask], dim=1), (1, 3, F, S // 4, S // 4), cond=cond, use_graph=a.graph)
        ev[2].record()
        frames = [m.generator.decode_sample(img, pred[0][:, i:i + a.chunk].contiguous()) for i in range(0, F, a.chunk)]
        ev[3].record()
        torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        tot = sum(t)
        print(f"clip {clip}: prep {t[0]:.1f} ms | {a.steps} DDIM steps {t[1]:.1f} ms ({t[1] / a.steps:.2f} ms/step{' graph' if a.graph else ''}) | "
              f"decode {t[2]:.1f} ms | total {tot:.1f} ms = {F / tot * 1e3:.0f} frames/s ({F / tot * 1e3 / 25:.1f} x real time at 25 fps); "
              f"finite {bool(torch.isfinite(frames[0]).all())}", flush=True)


if __name__ == "__main__":
    main()

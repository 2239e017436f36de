"""One-shot diagnosis of the LFG decoder on a GPU: every stage against the CPU oracle, no assertions (prints a table)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lfg_oracle as L        # noqa: E402
from oracle import weights as W           # noqa: E402
from dawn_pytorch_b200 import LfgGenerator  # noqa: E402


def ot(a, ref):
    a, ref = a.detach().float().cpu(), ref.detach().float().cpu()
    d = (a - ref).abs()
    return f"{(d / (1e-4 + 1e-3 * ref.abs())).max().item():10.3f} x tol   max|d| {d.max().item():.3e}   |ref| max {ref.abs().max().item():.3f}   nan {int(torch.isnan(a).sum())}"


def main():
    with open(os.path.join(ROOT, "tests", "golden", "lfg_state_dict_schema.json")) as f:
        sch = json.load(f)
    sd = W.lfg_synth_state_dict([(n, tuple(s)) for n, s in sch["entries"]])
    g = LfgGenerator(num_channels=3, num_regions=10, block_expansion=64, max_features=512, num_down_blocks=2, num_bottleneck_blocks=6,
                     skips=True)
    g.load_state_dict(sd, strict=True)
    g = g.cuda()
    for case, (nf, H, Wd, h, w) in {'lfg_small': (3, 64, 64, 16, 16), 'lfg_rect': (2, 64, 96, 16, 24), 'lfg_128': (2, 128, 128, 32, 32)}.items():
        src, flow, occ = W.lfg_synth_inputs(case, nf, H, Wd, h, w)
        taps = {}
        with torch.no_grad():
            ref = L.forward_with_flow(sd, L.LfgCfg(), src, flow, occ, taps=taps)
            fea = L.compute_fea(sd, L.LfgCfg(), src)
        try:
            print(f"[{case}] fea        {ot(g.compute_fea(src.cuda()), fea)}")
            out = g.forward_with_flow(src.cuda(), flow.cuda(), occ.cuda())
            torch.cuda.synchronize()
            print(f"[{case}] deformed   {ot(out['deformed'], ref['deformed'])}")
            for name in ("bottleneck", "up0", "up1"):
                print(f"[{case}] {name:10s} {ot(g.read_tap(name), taps[name])}")
            print(f"[{case}] prediction {ot(out['prediction'], ref['prediction'])}   launches {g.last_launch_count()}")
        except Exception as e:  # noqa: BLE001
            print(f"[{case}] FAILED: {type(e).__name__}: {e}")
    # timing at the bench shape: 200 frames of 256x256 from a 64x64 flow (BASELINE configs[4] decode stage), 50-frame chunks
    try:
        nf, H, Wd, h, w = 50, 256, 256, 64, 64
        src, flow, occ = W.lfg_synth_inputs("lfg_bench", nf, H, Wd, h, w)
        s, fl, oc = src.cuda(), flow.cuda(), occ.cuda()
        g.forward_with_flow(s, fl, oc, need_deformed=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            out = g.forward_with_flow(s, fl, oc, need_deformed=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        print(f"[bench] {nf} frames 256x256: {ms:.2f} ms per call = {ms * 4:.1f} ms per 200-frame clip ({nf * 78.7 / ms:.1f} TFLOP/s algorithmic at 78.7 GFLOP/frame); "
              f"workspace {g.workspace_bytes() / 2**30:.2f} GiB; finite {bool(torch.isfinite(out['prediction']).all())}")
    except Exception as e:  # noqa: BLE001
        print(f"[bench] FAILED: {type(e).__name__}: {e}")


if __name__ == "__main__":
    main()

"""GPU diagnostic: per-tap parity of the CUDA path vs the CPU oracle + first timings.  Writes gpurun_out/diag.json."""
import json
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O          # noqa: E402
from tests import gpu_common as G            # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
res = {"device": torch.cuda.get_device_name(0)}


def tapdiff(case):
    net = G.cuda_net()
    x, t, cond, _, _ = G.clip(case)
    taps_o = {}
    with torch.no_grad():
        ref = O.unet_forward(G.synth_sd(), O.UnetCfg(), x, t, cond, taps=taps_o)
    bufs = net.request_taps(list(taps_o), x.shape[2], x.shape[3], x.shape[4], torch.device("cuda"))
    net.update_num_frames(x.shape[2])
    with torch.no_grad():
        out = net.forward_with_cond_scale(x.cuda(), t.cuda(), cond=cond.cuda(), cond_scale=1.0)
    torch.cuda.synchronize()
    rows = []
    for name, r in taps_o.items():
        g = bufs[name].cpu()
        rows.append((name, G.over_tol(g, r), float((g - r).abs().max()), float(r.abs().max()), bool(torch.isfinite(g).all())))
    net.clear_taps()
    rows.append(("OUT", G.over_tol(out.cpu(), ref), float((out.cpu() - ref).abs().max()), float(ref.abs().max()), True))
    return rows


for case in ("band", "odd", "cfg1"):
    try:
        rows = tapdiff(case)
        res[case] = rows
        print(f"== {case}")
        for r in rows:
            print(f"  {r[0]:24s} over_tol {r[1]:10.3f}  max|d| {r[2]:.3e}  |ref|max {r[3]:.3f} finite {r[4]}")
    except Exception as e:
        traceback.print_exc()
        res[case] = "ERROR: " + repr(e)
        break


def timeit(F, h, w, reps=3):
    from oracle import weights as W
    net = G.cuda_net()
    x_t, fea, cond = W.synth_inputs("bench", F, h, w)
    net.update_num_frames(F)
    net.set_clip_invariants(fea[0].cuda(), cond[0].cuda())
    xt = x_t[0].cuda()
    t = torch.full((1,), 500, dtype=torch.long).cuda()
    out = torch.empty((3, F, h, w), device="cuda")
    for _ in range(2):
        net.forward_x3(xt, t, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        net.forward_x3(xt, t, out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, net.last_launch_count(), net.workspace_bytes()


if "--time" in sys.argv:
    for (F, h, w) in ((16, 32, 32), (100, 32, 32), (200, 64, 64)):
        try:
            ms, nl, wsb = timeit(F, h, w)
            res[f"time_{F}x{h}x{w}"] = dict(ms=ms, launches=nl, workspace_gb=wsb / 2**30)
            print(f"forward_x3 F={F} {h}x{w}: {ms:.2f} ms/step, {nl} launches, workspace {wsb/2**30:.2f} GiB")
        except Exception as e:
            traceback.print_exc()
            res[f"time_{F}x{h}x{w}"] = "ERROR: " + repr(e)
            break

with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
    json.dump(res, f, indent=1)

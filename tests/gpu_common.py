"""Helpers shared by the -m gpu tests: synthetic clips, the CUDA module with synthetic weights, tolerances."""
import json
import os

import numpy as np
import torch

from oracle import unet_oracle as O
from oracle import weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CTOR = dict(dim=64, cond_dim=1032, cond_aud=1024, cond_pose=6, cond_eye=2, num_frames=40, channels=275,
            out_grid_dim=2, out_conf_dim=1, dim_mults=(1, 2, 4, 8), use_hubert_audio_cond=True,
            learn_null_cond=False, use_final_activation=False, use_deconv=True, padding_mode="zeros", win_width=40)
CASES = {'cfg1': (16, 32, 32, 500), 'band': (96, 8, 8, 952), 'odd': (23, 16, 16, 47)}
RTOL, ATOL = 1e-3, 1e-4            # BASELINE.json north_star: rtol=1e-3 / atol=1e-4 fp32


def over_tol(a, ref):
    a, ref = a.detach().float().cpu(), ref.detach().float().cpu()
    return ((a - ref).abs() / (ATOL + RTOL * ref.abs())).max().item()


def schema():
    with open(os.path.join(ROOT, "tests", "golden", "state_dict_schema.json")) as f:
        return json.load(f)


_SD = None


def synth_sd():
    global _SD
    if _SD is None:
        _SD = W.synth_state_dict([(n, tuple(s)) for n, s in schema()["entries"]])
    return _SD


_NET = None


def cuda_net():
    """One module instance for the whole test session (weights are uploaded once)."""
    global _NET
    if _NET is None:
        from dawn_pytorch_b200 import DynamicNfUnet3D
        net = DynamicNfUnet3D(**CTOR).eval()
        net.load_state_dict(synth_sd(), strict=True)
        _NET = net.cuda()
    return _NET


def clip(case, F=None, h=None, w=None, t=None):
    if case in CASES and F is None:
        F, h, w, t = CASES[case]
    x_t, fea, cond = W.synth_inputs(case, F, h, w)
    x = torch.cat([x_t, fea.unsqueeze(2).expand(-1, -1, F, -1, -1)], dim=1).contiguous()
    return x, torch.full((1,), t, dtype=torch.long), cond, x_t, fea


def golden(case):
    return np.load(os.path.join(ROOT, "tests", "golden", f"{case}.npz"))

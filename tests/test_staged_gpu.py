"""GPU parity tests written without GPU budget left in their round: NOT part of `-m gpu` until they have passed once on a B200
(run with `-m staged_gpu`; skipped without a CUDA device).  Move a test into its permanent file when it is validated."""
import os

import numpy as np
import pytest
import torch

from oracle import weights as W
from tests import gpu_common as G

pytestmark = pytest.mark.staged_gpu


def test_classifier_free_guidance_sampling_matches_reference_golden():
    """N4: `ddim_sample(cond_scale=2)` — two hoisted UNet forwards per step (conditioning / all-zero null conditioning, U:879-890,
    920) — against the REAL reference's ddim_sample on the 'odd' clip (oracle/make_golden_cfg.py, 3 steps, injected noise)."""
    from dawn_pytorch_b200 import DynamicNfGaussianDiffusion
    g = np.load(os.path.join(G.ROOT, "tests", "golden", "ddim_cfg2_odd.npz"))
    steps, scale = int(g["steps"]), float(g["cond_scale"])
    net = G.cuda_net()
    D = DynamicNfGaussianDiffusion(denoise_fn=net, num_frames=40, image_size=32, sampling_timesteps=steps, timesteps=1000,
                                   loss_type='l2', use_dynamic_thres=True, null_cond_prob=0.1, ddim_sampling_eta=1.0).cuda()
    F, h, w, _ = G.CASES["odd"]
    _, fea, cond = W.synth_inputs("odd", F, h, w)
    D.update_num_frames(F)

    def noise_fn(k, shape):
        return torch.from_numpy(W.pseudo_normal(f"cfg2/noise{k}", tuple(shape)))

    img = D.ddim_sample(fea.cuda(), (1, 3, F, h, w), cond=cond.cuda(), cond_scale=scale, noise_fn=noise_fn)
    torch.cuda.synchronize()
    d = (img.cpu() - torch.from_numpy(g["sample"])).abs().max().item()
    print(f"cfg sampling (cond_scale {scale}, {steps} steps): max|d| = {d:.3e}")
    assert d < 2e-4

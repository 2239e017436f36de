"""TEST INFRASTRUCTURE — classifier-free-guidance sampling golden from the REAL reference:
`GaussianDiffusion.ddim_sample(..., cond_scale=2)` (U:1156-1208, two UNet forwards per step through forward_with_cond_scale,
U:879-890) on the 'odd' clip with 3 sampling steps and injected noise (torch.randn / randn_like patched, as make_golden_e2e.py).

Run in the build container only:    python oracle/make_golden_cfg.py"""
import importlib
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import weights as W          # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
U_MOD = 'DM_3.modules.video_flow_diffusion_multiGPU_v0_crema_plus_faceemb_ca_multi_test'
TAG, STEPS, SCALE = 'cfg2', 3, 2.0
CASE = (23, 16, 16)                       # the 'odd' geometry


def main():
    sys.path.insert(0, os.path.join(HERE, 'shims'))
    sys.path.insert(0, '/root/reference')
    warnings.filterwarnings("ignore")
    U = importlib.import_module(U_MOD)
    with open(os.path.join(GOLD, 'state_dict_schema.json')) as f:
        schema = [(n, tuple(s)) for n, s in json.load(f)['entries']]
    net = U.DynamicNfUnet3D(dim=64, cond_dim=1032, cond_aud=1024, cond_pose=6, cond_eye=2, num_frames=40, channels=275, out_grid_dim=2,
                            out_conf_dim=1, dim_mults=(1, 2, 4, 8), use_hubert_audio_cond=True, learn_null_cond=False,
                            use_final_activation=False, use_deconv=True, padding_mode="zeros", win_width=40).eval()
    net.load_state_dict(W.synth_state_dict(schema), strict=True)
    D = U.DynamicNfGaussianDiffusion(denoise_fn=net, num_frames=40, image_size=32, sampling_timesteps=STEPS, timesteps=1000,
                                     loss_type='l2', use_dynamic_thres=True, null_cond_prob=0.1, ddim_sampling_eta=1.0).eval()
    Fr, h, w = CASE
    _, fea, cond = W.synth_inputs('odd', Fr, h, w)
    net.update_num_frames(Fr); D.update_num_frames(Fr)
    counter = {"k": -1}

    def injected(shape):
        k = counter["k"]
        counter["k"] += 1
        return torch.from_numpy(W.pseudo_normal(f"{TAG}/noise{k}", tuple(shape)))

    real_randn, real_randn_like = torch.randn, torch.randn_like
    torch.randn = lambda *size, **kw: injected(size[0] if len(size) == 1 and not isinstance(size[0], int) else size)
    torch.randn_like = lambda t, **kw: injected(t.shape)
    try:
        with torch.no_grad():
            img = D.ddim_sample(fea, (1, 3, Fr, h, w), cond=cond, cond_scale=SCALE)
    finally:
        torch.randn, torch.randn_like = real_randn, real_randn_like
    assert counter["k"] == STEPS - 1
    print("cfg sample", tuple(img.shape), float(img.abs().max()), float(img.abs().mean()))
    np.savez_compressed(os.path.join(GOLD, "ddim_cfg2_odd.npz"), sample=img.numpy(), steps=np.int64(STEPS), cond_scale=np.float32(SCALE))


if __name__ == "__main__":
    main()

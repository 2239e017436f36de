"""TEST INFRASTRUCTURE — end-to-end golden from the REAL reference `FlowDiffusion.sample_one_video` (FD:325-406).

Run in the build container only (needs /root/reference):    python oracle/make_golden_e2e.py
The reference wrapper cannot be constructed offline (its __init__ loads checkpoints and calls .cuda(), FD:112-135), so the
instance is assembled by hand from the reference's own classes — `Generator`, `DynamicNfUnet3D`, `DynamicNfGaussianDiffusion`,
`Face_loc_Encoder` — with deterministic synthetic weights, and the UNMODIFIED `sample_one_video` / `generate_bbox_mask` /
`GaussianDiffusion.sample` / `ddim_sample` / `Generator.forward_with_flow` code runs on the CPU.  Two process-level patches make
that possible: `Tensor.cuda` is the identity, and `torch.randn` / `torch.randn_like` return the injected noise tensors
(oracle.weights.pseudo_normal) in call order, which is what `noise_fn` feeds the CUDA sampler in the test.
"""
import importlib
import os
import sys
import warnings

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import lfg_oracle as L       # noqa: E402
from oracle import weights as W          # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
FD_MOD = 'DM_3.modules.video_flow_diffusion_model_multiGPU_v0_crema_vgg_floss_plus_faceemb_flow_fast_init_cond_test'
U_MOD = 'DM_3.modules.video_flow_diffusion_multiGPU_v0_crema_plus_faceemb_ca_multi_test'
TAG, NF, IMG, STEPS = 'e2e', 8, 64, 3
PROBE_N = 4096


def probe_idx(name, numel):
    u = W.uniform01('probe/' + name, PROBE_N)
    return np.minimum((u.astype(np.float64) * numel).astype(np.int64), numel - 1)


def e2e_inputs():
    """What unified_video_generator.py:371-380 passes: image in [0, 1], HuBERT features, pose (7 values per frame, 6 used), blink,
    face box [x0, x1, y0, y1, W, H] per frame, first-frame pose / blink."""
    img = torch.from_numpy(W.uniform01(f"{TAG}/img", 3 * IMG * IMG).reshape(1, 3, IMG, IMG))
    hubert = torch.from_numpy(W.pseudo_normal(f"{TAG}/hubert", (1, NF, 1024)))
    pose = torch.from_numpy(W.symmetric(f"{TAG}/pose", (1, 7, NF), 0.3))
    eye = torch.from_numpy(W.uniform01(f"{TAG}/eye", 2 * NF).reshape(1, 2, NF))
    bbox = torch.tensor([[20., 44., 16., 50., 64., 64.]]).unsqueeze(-1).repeat(1, 1, NF)
    init_pose = pose[:, :6, 0].clone() + 0.05
    init_eye = eye[:, :, 0].clone()
    return img, hubert, pose, eye, bbox, init_pose, init_eye


def face_sd():
    shapes = {"conv1.weight": (8, 1, 3, 3), "conv1.bias": (8,), "conv2.weight": (16, 8, 3, 3), "conv2.bias": (16,)}
    return {k: torch.from_numpy(np.ascontiguousarray(W.synth_value("face_loc_emb." + k, s))).float() for k, s in shapes.items()}


def main():
    import json
    # the shims and the reference tree are only put on the path when goldens are generated: tests import this module for
    # e2e_inputs() / face_sd() and must not see the shim packages
    sys.path.insert(0, os.path.join(HERE, 'shims'))
    sys.path.insert(0, '/root/reference')
    warnings.filterwarnings("ignore")
    FD = importlib.import_module(FD_MOD)
    U = importlib.import_module(U_MOD)
    from LFG.modules.generator import Generator
    import yaml
    with open('/root/reference/config/hdtf128.yaml') as f:
        mp = yaml.safe_load(f)['model_params']
    with open(os.path.join(GOLD, 'state_dict_schema.json')) as f:
        unet_schema = [(n, tuple(s)) for n, s in json.load(f)['entries']]

    fd = FD.FlowDiffusion.__new__(FD.FlowDiffusion)                   # FD:112-135 needs checkpoints + a GPU: assemble by hand
    nn.Module.__init__(fd)
    fd.use_residual_flow, fd.pose_dim, fd.is_train = False, 6, False
    fd.generator = Generator(num_regions=mp['num_regions'], num_channels=mp['num_channels'], revert_axis_swap=mp['revert_axis_swap'],
                             **mp['generator_params']).eval()        # FD:116-121
    fd.generator.load_state_dict(W.lfg_synth_state_dict(L.state_dict_schema()), strict=False)
    fd.unet = U.DynamicNfUnet3D(dim=64, cond_dim=1024 + 6 + 2, cond_aud=1024, cond_pose=6, cond_eye=2, num_frames=40, channels=3 + 256 + 16,
                                out_grid_dim=2, out_conf_dim=1, dim_mults=(1, 2, 4, 8), use_hubert_audio_cond=True, learn_null_cond=False,
                                use_final_activation=False, use_deconv=True, padding_mode="zeros", win_width=40)      # FD:140-155
    fd.unet.load_state_dict(W.synth_state_dict(unet_schema), strict=True)
    fd.diffusion = U.DynamicNfGaussianDiffusion(denoise_fn=fd.unet, num_frames=40, image_size=32, sampling_timesteps=STEPS, timesteps=1000,
                                                loss_type='l2', use_dynamic_thres=True, null_cond_prob=0.1, ddim_sampling_eta=1.0)  # FD:157-167
    fd.face_loc_emb = FD.Face_loc_Encoder()
    fd.face_loc_emb.load_state_dict(face_sd(), strict=True)
    fd.eval()
    fd.update_num_frames(NF)

    img, hubert, pose, eye, bbox, init_pose, init_eye = e2e_inputs()
    counter = {"k": -1}

    def injected(shape):
        k = counter["k"]
        counter["k"] += 1
        return torch.from_numpy(W.pseudo_normal(f"{TAG}/noise{k}", tuple(shape)))

    real_randn, real_randn_like, real_cuda = torch.randn, torch.randn_like, torch.Tensor.cuda
    torch.randn = lambda *size, **kw: injected(size[0] if len(size) == 1 and not isinstance(size[0], int) else size)
    torch.randn_like = lambda t, **kw: injected(t.shape)
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with torch.no_grad():
            out = fd.sample_one_video(sample_img=img, sample_audio_hubert=hubert, sample_pose=pose, sample_eye=eye, sample_bbox=bbox,
                                      init_pose=init_pose, init_eye=init_eye, cond_scale=1.0)
            mask = fd.generate_bbox_mask(bbox.clone(), size=IMG)
            face = fd.face_loc_emb(mask)
    finally:
        torch.randn, torch.randn_like, torch.Tensor.cuda = real_randn, real_randn_like, real_cuda
    assert counter["k"] == STEPS - 1, counter            # start image + one noise per step except the last (U:1166, 1201)
    vid, warped = out["sample_out_vid"], out["sample_warped_vid"]
    print("grid", tuple(out["sample_vid_grid"].shape), float(out["sample_vid_grid"].abs().max()), "conf", float(out["sample_vid_conf"].min()),
          float(out["sample_vid_conf"].max()), "vid", tuple(vid.shape), float(vid.min()), float(vid.max()), "mask px", float(mask.sum()))
    ip = probe_idx(TAG + '/vid', vid.numel())
    np.savez_compressed(os.path.join(GOLD, "e2e_sample_one_video.npz"),
                        sample_vid_grid=out["sample_vid_grid"].numpy(), sample_vid_conf=out["sample_vid_conf"].numpy(),
                        out_vid_probe=vid.flatten()[ip].numpy(), warped_vid_probe=warped.flatten()[ip].numpy(),
                        out_vid_absmean=np.float32(vid.abs().mean()), bbox_mask_sum=np.float32(mask.sum()), face_emb=face.numpy(),
                        steps=np.int64(STEPS), frames=np.int64(NF), image=np.int64(IMG))
    print("wrote", os.path.join(GOLD, "e2e_sample_one_video.npz"))


if __name__ == "__main__":
    main()

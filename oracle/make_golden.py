"""TEST INFRASTRUCTURE — generate tests/golden/* by running the REAL reference.

Run in the build container only (needs /root/reference; the GPU box never runs this):
    python oracle/make_golden.py
It imports the unmodified reference UNet (U) and its windowed twin (UL) with the two shims in
oracle/shims, loads the deterministic synthetic weights of oracle/weights.py, and
  1. dumps the reference state_dict schema               -> tests/golden/state_dict_schema.json
  2. runs the reference on seeded synthetic clips         -> tests/golden/<case>.npz
     (inputs that are cheap to regenerate are NOT stored: oracle.weights.synth_inputs(tag) is
      exact on every platform; only outputs and probe samples are stored)
  3. checks oracle/unet_oracle.py against the reference at every sub-module boundary
     (forward hooks) and records per-tap probes (fixed sample of elements + abs-mean) so the
     restatement stays pinned where the reference cannot run.
"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, 'shims'))
sys.path.insert(0, '/root/reference')

from oracle import weights as W          # noqa: E402
from oracle import unet_oracle as O      # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
U_MOD = 'DM_3.modules.video_flow_diffusion_multiGPU_v0_crema_plus_faceemb_ca_multi_test'
UL_MOD = U_MOD + '_local_opt'

CTOR = dict(dim=64, cond_dim=1032, cond_aud=1024, cond_pose=6, cond_eye=2, num_frames=40, channels=275,
            out_grid_dim=2, out_conf_dim=1, dim_mults=(1, 2, 4, 8), use_hubert_audio_cond=True,
            learn_null_cond=False, use_final_activation=False, use_deconv=True, padding_mode="zeros",
            win_width=40)                                        # FD:140-155

# name -> (F, h, w, t)
CASES = {
    'cfg1': (16, 32, 32, 500),        # BASELINE configs[0]: 128x128 video, 16 frames
    'band': (96, 8, 8, 952),          # F > 81: the +-40 window is active
    'odd':  (23, 16, 16, 47),         # ragged: F not a multiple of anything
}
PROBE_N = 64


def probe_idx(name, numel):
    u = W.uniform01('probe/' + name, PROBE_N)
    return np.minimum((u.astype(np.float64) * numel).astype(np.int64), numel - 1)


def build_x(x_t, fea):
    Fr = x_t.shape[2]
    return torch.cat([x_t, fea.unsqueeze(2).expand(-1, -1, Fr, -1, -1)], dim=1).contiguous()


def hook_taps(net, taps):
    """Register forward hooks at the boundaries the oracle's `taps` uses."""
    hs = []

    def add(mod, name):
        hs.append(mod.register_forward_hook(lambda m, i, o, name=name: taps.__setitem__(name, o.detach().clone())))

    add(net.init_conv, 'init_conv')
    add(net.init_temporal_attn, 'init_temporal_attn')
    for L, blk in enumerate(net.downs):
        for j in range(5):
            if not isinstance(blk[j], torch.nn.Identity):
                add(blk[j], f'downs.{L}.{j}')
    add(net.mid_block1, 'mid_block1')
    add(net.mid_spatial_attn, 'mid_spatial_attn')
    add(net.mid_temporal_attn, 'mid_temporal_attn')
    add(net.mid_block2, 'mid_block2')
    for K, blk in enumerate(net.ups):
        for j in range(5):
            if not isinstance(blk[j], torch.nn.Identity):
                add(blk[j], f'ups.{K}.{j}')
    add(net.final_conv[0], 'final_conv.0')
    add(net.occlusion_map[0], 'occlusion_map.0')
    return hs


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    U = importlib.import_module(U_MOD)
    UL = importlib.import_module(UL_MOD)
    net = U.DynamicNfUnet3D(**CTOR).eval()
    schema = [(k, list(v.shape)) for k, v in net.state_dict().items()]
    with open(os.path.join(GOLD, 'state_dict_schema.json'), 'w') as f:
        json.dump({'ctor': {k: (list(v) if isinstance(v, tuple) else v) for k, v in CTOR.items()},
                   'entries': schema}, f)
    sd = W.synth_state_dict(schema)
    net.load_state_dict(sd, strict=True)
    net_l = UL.DynamicNfUnet3D(**CTOR).eval()
    net_l.load_state_dict(sd, strict=True)                        # same names
    cfg = O.UnetCfg()
    report = {}

    for case, (Fr, h, w, t) in CASES.items():
        x_t, fea, cond = W.synth_inputs(case, Fr, h, w)
        x = build_x(x_t, fea)
        tt = torch.full((1,), t, dtype=torch.long)
        net.update_num_frames(Fr)
        taps_ref, taps_or = {}, {}
        hs = hook_taps(net, taps_ref)
        t0 = time.time()
        with torch.no_grad():
            ref = net.forward_with_cond_scale(x, tt, cond=cond, cond_scale=1.0)
        t_ref = time.time() - t0
        for hdl in hs:
            hdl.remove()
        t0 = time.time()
        with torch.no_grad():
            ora = O.unet_forward(sd, cfg, x, tt, cond, band=None, taps=taps_or)
            ora_band = O.unet_forward(sd, cfg, x, tt, cond, band=cfg.win)
        t_or = time.time() - t0
        tol = 1e-4 + 1e-3 * ref.abs()
        r_glob = ((ora - ref).abs() / tol).max().item()
        r_band = ((ora_band - ref).abs() / tol).max().item()
        worst_tap = 0.0
        probes = {}
        for name, tr in taps_ref.items():
            to = taps_or[name]
            assert to.shape == tr.shape, (name, to.shape, tr.shape)
            rr = ((to - tr).abs() / (1e-4 + 1e-3 * tr.abs())).max().item()
            worst_tap = max(worst_tap, rr)
            flat = tr.reshape(-1)
            idx = probe_idx(f'{case}/{name}', flat.numel())
            probes[name] = dict(shape=list(tr.shape), absmean=float(flat.abs().mean()),
                                vals=flat[idx].tolist())
        assert set(taps_or) == set(taps_ref), set(taps_or) ^ set(taps_ref)
        print(f'[{case}] F={Fr} {h}x{w} t={t}: ref {t_ref:.2f}s oracle(2x) {t_or:.2f}s  |ref|max {ref.abs().max():.3f} '
              f'oracle/ref x tol: global {r_glob:.4f} banded {r_band:.4f} worst tap {worst_tap:.4f}')
        assert r_glob < 0.2 and r_band < 0.2 and worst_tap < 0.2, 'oracle restatement disagrees with the reference'
        out = dict(eps=ref.numpy())
        extra = {}
        if case == 'band':                                        # direct UL golden: proves U == UL
            net_l.update_num_frames(Fr)
            with torch.no_grad():
                ref_l = net_l.forward_with_cond_scale(x, tt, cond=cond, cond_scale=1.0)
            extra['ul_vs_u_maxabs'] = float((ref_l - ref).abs().max())
            out['eps_local_opt'] = ref_l.numpy()
            print(f'   UL (local_opt) vs U max|d| = {extra["ul_vs_u_maxabs"]:.3e}')
            assert extra['ul_vs_u_maxabs'] < 2e-5
        if case == 'odd':                                         # CFG path: cond_scale != 1 -> two forwards (U:886-890)
            with torch.no_grad():
                ref_cfg = net.forward_with_cond_scale(x, tt, cond=cond, cond_scale=2.0)
                ora_cfg = O.forward_with_cond_scale(sd, cfg, x, tt, cond, cond_scale=2.0)
            rc = ((ora_cfg - ref_cfg).abs() / (1e-4 + 1e-3 * ref_cfg.abs())).max().item()
            print(f'   cond_scale=2: oracle/ref x tol {rc:.4f}')
            assert rc < 0.2
            out['eps_cond_scale2'] = ref_cfg.numpy()
        np.savez_compressed(os.path.join(GOLD, f'{case}.npz'), **out)
        report[case] = dict(F=Fr, h=h, w=w, t=t, ref_absmax=float(ref.abs().max()),
                            oracle_over_tol=r_glob, oracle_band_over_tol=r_band, worst_tap_over_tol=worst_tap,
                            ref_seconds=t_ref, probes=probes, **extra)

    # sampler golden (row a16): 3 DDIM steps with injected noise on the 'band' clip
    D = U.DynamicNfGaussianDiffusion(denoise_fn=net, num_frames=40, image_size=32, sampling_timesteps=20,
                                     timesteps=1000, loss_type='l2', use_dynamic_thres=True,
                                     null_cond_prob=0.1, ddim_sampling_eta=1.0)          # FD:156-166
    pairs = O.ddim_time_pairs()
    exp_pairs = [(int(a), int(b)) for a, b in pairs]
    acp_o, prev_o = O.cosine_alphas_cumprod()
    assert torch.equal(acp_o, D.alphas_cumprod) and torch.equal(prev_o, D.alphas_cumprod_prev)
    Fr, h, w, _ = CASES['odd']
    x_t, fea, cond = W.synth_inputs('odd', Fr, h, w)
    net.update_num_frames(Fr)
    img = x_t.clone()
    img_o = x_t.clone()
    fea_rep = fea.unsqueeze(2).repeat(1, 1, Fr, 1, 1)
    steps = [pairs[0], pairs[9], pairs[-1]]                    # first, middle, last (t_next = 0 -> no noise)
    traj = []
    for k, (t, tn) in enumerate(steps):
        noise = torch.from_numpy(W.pseudo_normal(f'odd/noise{k}', tuple(img.shape)))
        tc = torch.full((1,), t, dtype=torch.long)
        with torch.no_grad():
            # reference arithmetic, U:1170-1205, with the injected noise instead of randn_like
            alpha, alpha_next = D.alphas_cumprod_prev[t], D.alphas_cumprod_prev[tn]
            eps = net.forward_with_cond_scale(torch.cat([img, fea_rep], dim=1), tc, cond=cond, cond_scale=1.0)
            x0 = D.predict_start_from_noise(img, t=tc, noise=eps)
            s = torch.quantile(x0.reshape(1, -1).abs(), 0.9, dim=-1).clamp_(min=1.).view(-1, 1, 1, 1, 1)
            x0 = x0.clamp(-s, s) / s
            sigma = 1.0 * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            c = ((1 - alpha_next) - sigma ** 2).sqrt()
            nz = noise if tn > 0 else 0.
            img = x0 * alpha_next.sqrt() + c * eps + sigma * nz
            eps_o = O.unet_forward(sd, cfg, torch.cat([img_o, fea_rep], dim=1), tc, cond)
            img_o = O.ddim_step(eps_o, img_o, t, tn, noise)
        traj.append(img.numpy().copy())
        d = (img_o - img).abs().max().item()
        print(f'   ddim step {k} (t={t}->{tn}): oracle vs ref max|d| {d:.3e}, s={float(s):.4f}')
        assert d < 2e-4
    np.savez_compressed(os.path.join(GOLD, 'ddim_odd.npz'), x_after=np.stack(traj),
                        steps=np.array(steps, dtype=np.int64))
    report['ddim_pairs'] = exp_pairs
    with open(os.path.join(GOLD, 'report.json'), 'w') as f:
        json.dump(report, f)
    print('golden vectors written to', GOLD)


if __name__ == '__main__':
    main()

"""TEST INFRASTRUCTURE — generate tests/golden/lfg_*.npz by running the REAL reference LFG Generator.

Run in the build container only (needs /root/reference):    python oracle/make_golden_lfg.py
Imports the unmodified `LFG.modules.generator.Generator` (shims for the un-installed, decode-irrelevant imports
matplotlib / skimage under oracle/shims), loads the deterministic synthetic weights of oracle/weights.py, runs
`forward_with_flow` frame by frame exactly as `sample_one_video` does (FD:375-383) and `compute_fea`, checks
oracle/lfg_oracle.py against it and stores the reference outputs as golden vectors.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, 'shims'))
sys.path.insert(0, '/root/reference')
warnings.filterwarnings("ignore")

from oracle import lfg_oracle as L       # noqa: E402
from oracle import weights as W          # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
# name -> (frames, H, W, h, w): full outputs stored
CASES = {'lfg_small': (3, 64, 64, 16, 16), 'lfg_rect': (2, 64, 96, 16, 24)}
# larger cases: only a fixed sample of output elements is stored
PROBED = {'lfg_128': (2, 128, 128, 32, 32)}
PROBE_N = 4096


def probe_idx(name, numel):
    u = W.uniform01('probe/' + name, PROBE_N)
    return np.minimum((u.astype(np.float64) * numel).astype(np.int64), numel - 1)


def main():
    import yaml
    from LFG.modules.generator import Generator
    with open('/root/reference/config/hdtf128.yaml') as f:
        cfg = yaml.safe_load(f)
    mp = cfg['model_params']
    gen = Generator(num_regions=mp['num_regions'], num_channels=mp['num_channels'], revert_axis_swap=mp['revert_axis_swap'],
                    **mp['generator_params']).eval()                                            # FD:116-121
    ref_sd = gen.state_dict()
    schema = L.state_dict_schema()
    decode_keys = [k for k in ref_sd if not k.startswith('pixelwise_flow_predictor.')]
    assert [n for n, _ in schema] == decode_keys, "oracle schema must list the reference's decode-path keys in order"
    for n, s in schema:
        assert tuple(ref_sd[n].shape) == tuple(s), n
    sd = W.lfg_synth_state_dict(schema)
    missing, unexpected = gen.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith('pixelwise_flow_predictor.') for k in missing)
    with open(os.path.join(GOLD, 'lfg_state_dict_schema.json'), 'w') as f:
        json.dump({"entries": [[n, list(s)] for n, s in schema],
                   "ignored_prefix": "pixelwise_flow_predictor.",
                   "n_reference_keys": len(ref_sd)}, f, indent=0)
    ocfg = L.LfgCfg()
    worst = 0.0
    for name, (nf, H, Wd, h, w) in {**CASES, **PROBED}.items():
        src, flow, occ = W.lfg_synth_inputs(name, nf, H, Wd, h, w)
        preds, defs = [], []
        with torch.no_grad():
            fea_ref = gen.compute_fea(src)
            for i in range(nf):                                                                 # FD:375-383: batch 1 per frame
                o = gen.forward_with_flow(source_image=src, optical_flow=flow[i:i + 1], occlusion_map=occ[i:i + 1])
                preds.append(o["prediction"]); defs.append(o["deformed"])
            pred_ref, def_ref = torch.cat(preds), torch.cat(defs)
            taps = {}
            mine = L.forward_with_flow(sd, ocfg, src, flow, occ, taps=taps)
            fea = L.compute_fea(sd, ocfg, src)
        d_pred = (mine["prediction"] - pred_ref).abs().max().item()
        d_def = (mine["deformed"] - def_ref).abs().max().item()
        d_fea = (fea - fea_ref).abs().max().item()
        rel = ((mine["prediction"] - pred_ref).abs() / (1e-4 + 1e-3 * pred_ref.abs())).max().item()
        worst = max(worst, rel)
        print(f"{name}: oracle vs reference  prediction max|d| {d_pred:.2e} ({rel:.3f} x tol)  deformed {d_def:.2e}  fea {d_fea:.2e};"
              f"  |prediction| in [{pred_ref.min():.3f}, {pred_ref.max():.3f}], |bottleneck| max {taps['bottleneck'].abs().max():.2f},"
              f" up1 max {taps['up1'].abs().max():.2f}")
        assert rel < 0.2 and d_def < 1e-5 and d_fea < 1e-4
        if name in CASES:
            np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), prediction=pred_ref.numpy(), deformed=def_ref.numpy(),
                                fea_absmean=np.float32(fea_ref.abs().mean()), fea_probe=fea_ref.flatten()[probe_idx(name + '/fea', fea_ref.numel())].numpy())
        else:
            ip = probe_idx(name + '/pred', pred_ref.numel())
            np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), prediction_probe=pred_ref.flatten()[ip].numpy(),
                                deformed_probe=def_ref.flatten()[ip].numpy(), prediction_absmean=np.float32(pred_ref.abs().mean()),
                                fea_probe=fea_ref.flatten()[probe_idx(name + '/fea', fea_ref.numel())].numpy())
    print("worst oracle-vs-reference:", worst, "x tol")


if __name__ == "__main__":
    main()

"""CPU: the LFG flow-decoder oracle (oracle/lfg_oracle.py) against golden vectors produced by the REAL reference
`Generator.forward_with_flow` / `compute_fea` (oracle/make_golden_lfg.py), on the same seeded inputs and synthetic weights."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import lfg_oracle as L
from oracle import weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {'lfg_small': (3, 64, 64, 16, 16), 'lfg_rect': (2, 64, 96, 16, 24)}
PROBED = {'lfg_128': (2, 128, 128, 32, 32)}
PROBE_N = 4096


def probe_idx(name, numel):
    u = W.uniform01('probe/' + name, PROBE_N)
    return np.minimum((u.astype(np.float64) * numel).astype(np.int64), numel - 1)


def lfg_sd():
    with open(os.path.join(GOLD, "lfg_state_dict_schema.json")) as f:
        sch = json.load(f)
    schema = [(n, tuple(s)) for n, s in sch["entries"]]
    assert schema == [(n, tuple(s)) for n, s in L.state_dict_schema()]
    return W.lfg_synth_state_dict(schema), sch


def test_schema_is_the_reference_decode_path():
    sd, sch = lfg_sd()
    assert len(sd) == 121 and sch["n_reference_keys"] == 196          # 75 pixelwise_flow_predictor.* entries are never read by the decode path
    assert sd["first.conv.weight"].shape == (64, 3, 7, 7) and sd["final.weight"].shape == (3, 64, 7, 7)
    assert all((v > 0).all() for k, v in sd.items() if k.endswith("running_var"))


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_matches_reference_golden(case):
    sd, _ = lfg_sd()
    nf, H, Wd, h, w = CASES[case]
    src, flow, occ = W.lfg_synth_inputs(case, nf, H, Wd, h, w)
    g = np.load(os.path.join(GOLD, f"{case}.npz"))
    with torch.no_grad():
        out = L.forward_with_flow(sd, L.LfgCfg(), src, flow, occ)
        fea = L.compute_fea(sd, L.LfgCfg(), src)
    ref = torch.from_numpy(g["prediction"])
    assert out["prediction"].shape == ref.shape == (nf, 3, H, Wd)
    assert ((out["prediction"] - ref).abs() / (1e-4 + 1e-3 * ref.abs())).max().item() <= 0.2
    assert (out["deformed"] - torch.from_numpy(g["deformed"])).abs().max().item() <= 1e-5
    assert fea.shape == (1, 256, H // 4, Wd // 4)
    assert np.abs(fea.flatten()[probe_idx(case + '/fea', fea.numel())].numpy() - g["fea_probe"]).max() <= 1e-4
    assert 0.0 < ref.min() and ref.max() < 1.0 and ref.std() > 0.05   # a non-degenerate image


def test_oracle_matches_reference_probes_128():
    sd, _ = lfg_sd()
    nf, H, Wd, h, w = PROBED['lfg_128']
    src, flow, occ = W.lfg_synth_inputs('lfg_128', nf, H, Wd, h, w)
    g = np.load(os.path.join(GOLD, "lfg_128.npz"))
    with torch.no_grad():
        out = L.forward_with_flow(sd, L.LfgCfg(), src, flow, occ)
    ip = probe_idx('lfg_128/pred', out["prediction"].numel())
    ref = g["prediction_probe"]
    assert (np.abs(out["prediction"].flatten()[ip].numpy() - ref) / (1e-4 + 1e-3 * np.abs(ref))).max() <= 0.2
    assert np.abs(out["deformed"].flatten()[ip].numpy() - g["deformed_probe"]).max() <= 1e-5

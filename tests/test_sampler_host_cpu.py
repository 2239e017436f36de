"""CPU: host-side control flow of the sampler (no GPU): which native calls one `ddim_sample` makes, in which order, for the
shipped cond_scale = 1 loop and for classifier-free guidance; the DDIM schedule and coefficients against the oracle's restatement
of the reference arithmetic (U:1156-1205)."""
import unittest.mock as um

import torch

from oracle import unet_oracle as O
from tests import gpu_common as G


class _FakeLib:
    def __init__(self, calls):
        self.calls = calls

    def __getattr__(self, name):
        def f(*a):
            self.calls.append(name)
            return 0
        return f


def _sampler(steps):
    from dawn_pytorch_b200 import DynamicNfGaussianDiffusion, DynamicNfUnet3D
    net = DynamicNfUnet3D(**G.CTOR).eval()
    D = DynamicNfGaussianDiffusion(denoise_fn=net, num_frames=40, image_size=32, sampling_timesteps=steps, timesteps=1000, loss_type='l2',
                                   use_dynamic_thres=True, null_cond_prob=0.1, ddim_sampling_eta=1.0)
    return D, net


def test_native_call_sequence_of_one_clip():
    import dawn_pytorch_b200.diffusion as dd
    D, net = _sampler(3)
    D.update_num_frames(4)
    calls = []
    net.set_clip_invariants = lambda f, c: calls.append(("invariants", bool(c.abs().sum() > 0)))
    net.forward_x3 = lambda x, t, e: calls.append(("forward_x3", int(t)))
    net._handle = None
    stream = type("S", (), {"cuda_stream": 0})()
    with um.patch.object(dd, "lib", _FakeLib(calls)), um.patch("torch.cuda.current_stream", lambda: stream):
        noise = lambda k, shp: torch.zeros(shp)                         # noqa: E731
        D.ddim_sample(torch.rand(1, 272, 8, 8), (1, 3, 4, 8, 8), cond=torch.randn(1, 4, 1032), cond_scale=1.0, noise_fn=noise)
        plain = list(calls)
        calls.clear()
        D.ddim_sample(torch.rand(1, 272, 8, 8), (1, 3, 4, 8, 8), cond=torch.randn(1, 4, 1032), cond_scale=2.0, noise_fn=noise)
        guided = list(calls)
    ts = [t for t, _ in D.ddim_schedule()]
    assert ts == [750, 500, 250]                                        # linspace(0, 1000, 5)[:-1] reversed, U:1160-1162
    assert plain == [("invariants", True)] + [c for t in ts for c in (("forward_x3", t), "dawn_unet_ddim_step")]
    per_step = lambda t: [("invariants", True), ("forward_x3", t), ("invariants", False), ("forward_x3", t), "dawn_unet_ddim_step"]   # noqa: E731
    assert guided == [c for t in ts for c in per_step(t)]              # cond then all-zero null cond, U:879-890, 920


def test_schedule_and_coefficients_match_the_oracle():
    D, _ = _sampler(20)
    pairs = D.ddim_schedule()
    assert pairs == [(int(a), int(b)) for a, b in O.ddim_time_pairs()]
    assert pairs[0] == (952, 904) and pairs[-1] == (47, 0) and len(pairs) == 20      # SURVEY a16 [verified]
    acp, prev = O.cosine_alphas_cumprod()
    assert torch.equal(acp, D.alphas_cumprod) and torch.equal(prev, D.alphas_cumprod_prev)
    # one update with the host coefficients == the oracle's ddim_step (reference arithmetic, incl. the alphas_cumprod_prev indexing quirk)
    g = torch.Generator().manual_seed(0)
    img, eps, noise = [torch.randn(1, 3, 5, 8, 8, generator=g) for _ in range(3)]
    for (t, tn) in (pairs[0], pairs[9], pairs[-1]):
        ca, cb, san, c, sigma = D.ddim_coefficients(t, tn)
        x0 = ca * img - cb * eps
        s = torch.quantile(x0.reshape(1, -1).abs(), 0.9, dim=-1).clamp(min=1.0)
        mine = (x0.clamp(-s, s) / s) * san + c * eps + (sigma * noise if tn > 0 else 0)
        ref = O.ddim_step(eps, img, t, tn, noise)
        assert (mine - ref).abs().max().item() < 1e-5

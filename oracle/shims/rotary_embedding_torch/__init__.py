"""Oracle scaffolding (test infrastructure, never shipped): stand-in for the un-vendored
third-party package `rotary-embedding-torch==0.3.5` (reference requirements.txt:133) as the
reference uses it: `RotaryEmbedding(dim)` (…ca_multi_test.py:761) /
`RotaryEmbedding(dim, seq_before_head_dim=True)` (…_local_opt.py:763) and
`.rotate_queries_or_keys(t)` (…ca_multi_test.py:692-693, local_attention.py:331-332).

Published algorithm restated: theta=10000, freqs = theta^(-2i/dim),
i = 0..dim/2-1, stored as a non-trainable nn.Parameter named `freqs`; position = arange(seq);
each angle repeated pair-wise; interleaved rotate_half (x0,x1)->(-x1,x0).
PARITY UNPINNED at this boundary: no reference test pins the rotary arithmetic."""
import torch
from torch import nn
from einops import rearrange, repeat


def rotate_half(x):
    x = rearrange(x, '... (d r) -> ... d r', r=2)
    x1, x2 = x.unbind(dim=-1)
    x = torch.stack((-x2, x1), dim=-1)
    return rearrange(x, '... d r -> ... (d r)')


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000, seq_before_head_dim=False, learned_freq=False):
        super().__init__()
        freqs = 1. / (theta ** (torch.arange(0, dim, 2)[:(dim // 2)].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad=learned_freq)
        self.default_seq_dim = -3 if seq_before_head_dim else -2

    def rotate_queries_or_keys(self, t, seq_dim=None):
        seq_dim = self.default_seq_dim if seq_dim is None else seq_dim
        n = t.shape[seq_dim]
        pos = torch.arange(n, device=t.device, dtype=t.dtype)
        freqs = torch.einsum('..., f -> ... f', pos.type(self.freqs.dtype), self.freqs)
        freqs = repeat(freqs, '... n -> ... (n r)', r=2)
        if seq_dim == -3:
            freqs = rearrange(freqs, 'n d -> n 1 d')
        rot = freqs.shape[-1]
        tm, tr = t[..., :rot], t[..., rot:]
        tm = tm * freqs.cos() + rotate_half(tm) * freqs.sin()
        return torch.cat((tm, tr), dim=-1)

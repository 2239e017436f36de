"""TEST INFRASTRUCTURE — CPU oracle for the DAWN denoising UNet (one "denoising step").

This file is a from-scratch fp32 restatement (torch CPU ops, functional style, frames-as-batch
layout) of the reference's hot path so that parity can be checked on machines where
/root/reference does not exist (the GPU box).  It is NOT product code: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Reference files restated (all under /root/reference/DM_3/modules/):
  U  = video_flow_diffusion_multiGPU_v0_crema_plus_faceemb_ca_multi_test.py
  UL = ..._test_local_opt.py,  LA = local_attention.py
Each function cites the U/LA lines it follows.

Pinning: oracle/make_golden.py imports the real reference (with oracle/shims for the two
un-vendored third-party imports) in the build container, checks this restatement against it
at every sub-module boundary and on the full forward, and commits the golden vectors under
tests/golden/.  tests/test_oracle_golden.py re-checks the restatement against those vectors
on every run.  The rotary-embedding arithmetic is third-party (rotary-embedding-torch 0.3.5,
not vendored): PARITY UNPINNED at that one boundary (see oracle/shims/rotary_embedding_torch).
"""
import math
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- config


class UnetCfg:
    """Hyper-parameters as the consumer constructs them (FD:140-155, config/DAWN_128.yaml)."""

    def __init__(self, dim=64, dim_mults=(1, 2, 4, 8), channels=275, cond_aud=1024, cond_pose=6,
                 cond_eye=2, out_grid_dim=2, out_conf_dim=1, attn_heads=8, attn_dim_head=32,
                 resnet_groups=8, init_kernel_size=7, win_width=40):
        self.dim = dim
        self.dim_mults = tuple(dim_mults)
        self.channels = channels
        self.cond_aud, self.cond_pose, self.cond_eye = cond_aud, cond_pose, cond_eye
        self.cond_dim = cond_aud + cond_pose + cond_eye
        self.out_grid_dim, self.out_conf_dim = out_grid_dim, out_conf_dim
        self.heads, self.dim_head = attn_heads, attn_dim_head
        self.groups = resnet_groups
        self.init_k = init_kernel_size
        self.win = win_width
        dims = [dim] + [dim * m for m in dim_mults]
        self.in_out = list(zip(dims[:-1], dims[1:]))           # U:783-784


# ----------------------------------------------------------------------------- primitives
# activations are (F, C, H, W): frames as the batch dimension, one clip (b = 1) at a time.

def chan_layernorm(x, gamma, eps=1e-5):
    """U:179-188 LayerNorm over channels, biased variance, gain only."""
    mean = x.mean(dim=1, keepdim=True)
    var = x.var(dim=1, unbiased=False, keepdim=True)
    return (x - mean) / (var + eps).sqrt() * gamma.reshape(1, -1, 1, 1)


def token_layernorm(x, g, eps=1e-5):
    """U:190-203 LayerNorm_img over the last dim (fp32 -> eps 1e-5), rsqrt form, gain only."""
    mean = x.mean(dim=-1, keepdim=True)
    var = x.var(dim=-1, unbiased=False, keepdim=True)
    return (x - mean) * (var + eps).rsqrt() * g


def clip_groupnorm(x, groups, w, b, eps=1e-5):
    """U:230 nn.GroupNorm on the 5-D (b,C,F,H,W) tensor: statistics span ALL frames of the clip."""
    Fr, C, H, W = x.shape
    y = x.permute(1, 0, 2, 3).reshape(1, C, Fr * H * W)
    y = F.group_norm(y, groups, w, b, eps)
    return y.reshape(C, Fr, H, W).permute(1, 0, 2, 3)


def rel_pos_bucket(rel, num_buckets=32, max_distance=32):
    """U:91-109 T5 bidirectional bucket of rel = j - i (int64 tensor)."""
    n = -rel
    nb = num_buckets // 2
    ret = (n < 0).long() * nb
    n = n.abs()
    max_exact = nb // 2
    is_small = n < max_exact
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return ret + torch.where(is_small, n, large)


def rel_pos_bias(emb_w, n, window):
    """U:111-119: (heads, n, n) bias = Embedding[bucket] - 1e8 outside |j-i| <= window."""
    pos = torch.arange(n)
    rel = pos[None, :] - pos[:, None]
    bucket = rel_pos_bucket(rel, 32, 32)                        # U:767-768 max_distance=32
    vals = emb_w[bucket]                                        # (n, n, heads)
    mask = -((rel > window) | (rel < -window)).float() * 1e8
    return vals.permute(2, 0, 1) + mask


def rotary(t, freqs):
    """rotary-embedding-torch 0.3.5 rotate_queries_or_keys (third-party).
    t: (..., n, d) with the sequence on dim -2; interleaved pairs; position = arange(n)."""
    n = t.shape[-2]
    ang = torch.arange(n, dtype=t.dtype)[:, None] * freqs[None, :]      # (n, d/2), fp32 product
    ang = ang.repeat_interleave(2, dim=-1)                              # (n, d)
    t2 = t.reshape(*t.shape[:-1], -1, 2)
    rot = torch.stack((-t2[..., 1], t2[..., 0]), dim=-1).reshape(t.shape)
    return t * ang.cos() + rot * ang.sin()


def sinusoidal(t, dim):
    """U:150-162."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half) * -e)
    e = t[:, None].float() * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


# ----------------------------------------------------------------------------- sub-modules

def block(sd, p, x, groups, film=None):
    """U:226-248 Block: conv(1,3,3) -> GroupNorm over the clip -> FiLM -> SiLU."""
    w = sd[p + '.proj.weight'][:, :, 0]
    y = F.conv2d(x, w, sd[p + '.proj.bias'], padding=1)
    y = clip_groupnorm(y, groups, sd[p + '.norm.weight'], sd[p + '.norm.bias'])
    if film is not None:
        scale, shift = film
        y = y * (scale.reshape(1, -1, 1, 1) + 1) + shift.reshape(1, -1, 1, 1)
    return F.silu(y)


def cross_attention(sd, p, tok, ctx, heads=8, dim_head=8, scale=8.0):
    """U:481-559 CrossAttention with exactly two keys per query (null + the frame's own condition).
    tok: (F, n, ci) tokens of each frame; ctx: (F, 2*co) conditioning vector of each frame."""
    Fr, n, _ = tok.shape
    x = token_layernorm(tok, sd[p + '.norm.g'])
    q = x @ sd[p + '.to_q.weight'].t()                                   # (F, n, 64)
    kv = ctx @ sd[p + '.to_kv.weight'].t()                               # (F, 128)
    k, v = kv.chunk(2, dim=-1)
    q = q.reshape(Fr, n, heads, dim_head).permute(0, 2, 1, 3)            # (F, h, n, d)
    k = k.reshape(Fr, heads, 1, dim_head)
    v = v.reshape(Fr, heads, 1, dim_head)
    nk = sd[p + '.null_kv'][0].reshape(1, 1, 1, dim_head).expand(Fr, heads, 1, dim_head)
    nv = sd[p + '.null_kv'][1].reshape(1, 1, 1, dim_head).expand(Fr, heads, 1, dim_head)
    k = torch.cat((nk, k), dim=-2)
    v = torch.cat((nv, v), dim=-2)
    q = F.normalize(q, dim=-1) * sd[p + '.q_scale']
    k = F.normalize(k, dim=-1) * sd[p + '.k_scale']
    sim = torch.einsum('fhid,fhjd->fhij', q, k) * scale
    attn = sim.softmax(dim=-1, dtype=torch.float32)
    out = torch.einsum('fhij,fhjd->fhid', attn, v)
    out = out.permute(0, 2, 1, 3).reshape(Fr, n, heads * dim_head)
    out = out @ sd[p + '.to_out.0.weight'].t()
    return token_layernorm(out, sd[p + '.to_out.1.g'])


def resnet_block(sd, p, x, cfg, t_emb=None, cond=None):
    """U:363-479 ResnetBlock_ca_mul.  x (F,ci,H,W); t_emb (256,) or None; cond (F,1032) or None."""
    Fr, ci, H, W = x.shape
    film = None
    has_cond = (p + '.audio_mlp.1.weight') in sd and cond is not None
    if (p + '.time_mlp.1.weight') in sd:
        te = F.silu(t_emb) @ sd[p + '.time_mlp.1.weight'].t() + sd[p + '.time_mlp.1.bias']
        film = te.chunk(2, dim=0)                                        # scale first (U:435)
    if has_cond:
        a_in = cond[:, :cfg.cond_aud]
        p_in = cond[:, cfg.cond_aud:cfg.cond_aud + cfg.cond_pose]
        e_in = cond[:, cfg.cond_aud + cfg.cond_pose:]
        a = F.silu(a_in) @ sd[p + '.audio_mlp.1.weight'].t() + sd[p + '.audio_mlp.1.bias']
        po = F.silu(p_in) @ sd[p + '.pose_mlp.1.weight'].t() + sd[p + '.pose_mlp.1.bias']
        e = F.silu(e_in) @ sd[p + '.eye_mlp.1.weight'].t() + sd[p + '.eye_mlp.1.bias']
        tok = x.permute(0, 2, 3, 1).reshape(Fr, H * W, ci)               # raw block input (U:454)
        hc = (cross_attention(sd, p + '.cross_attn_pose', tok, po)
              + cross_attention(sd, p + '.cross_attn_aud', tok, a)
              + cross_attention(sd, p + '.cross_attn_eye', tok, e))
        hc = hc.reshape(Fr, H, W, -1).permute(0, 3, 1, 2)
    h = block(sd, p + '.block1', x, cfg.groups, film)
    if has_cond:
        h = hc + h
    h = block(sd, p + '.block2', h, cfg.groups)
    if (p + '.res_conv.weight') in sd:
        res = F.conv2d(x, sd[p + '.res_conv.weight'][:, :, 0], sd[p + '.res_conv.bias'])
    else:
        res = x
    return h + res


def spatial_linear_attention(sd, p, x, heads=8, dim_head=32):
    """Residual(PreNorm(SpatialLinearAttention)) U:602-627, 832-833.  p = '<...>.fn'."""
    Fr, C, H, W = x.shape
    xn = chan_layernorm(x, sd[p + '.norm.gamma'])
    qkv = F.conv2d(xn, sd[p + '.fn.to_qkv.weight'])
    q, k, v = qkv.reshape(Fr, 3, heads, dim_head, H * W).unbind(dim=1)   # (F, h, d, n)
    q = q.softmax(dim=-2) * dim_head ** -0.5
    k = k.softmax(dim=-1)
    ctx = torch.einsum('fhdn,fhen->fhde', k, v)
    out = torch.einsum('fhde,fhdn->fhen', ctx, q).reshape(Fr, heads * dim_head, H, W)
    out = F.conv2d(out, sd[p + '.fn.to_out.weight'], sd[p + '.fn.to_out.bias'])
    return out + x


def temporal_attention(sd, p, x, bias, freqs, heads=8, dim_head=32, band=None):
    """Residual(PreNorm(EinopsToAndFrom(Attention))) over frames, U:648-725 (global form) ==
    LA:275-342 + LA:71-99 (banded form; UL).  p = '<...>.fn';  bias (heads, F, F).
    band=None: materialise the (F,F) scores like U.  band=w: only keys |i-j|<=w like UL."""
    Fr, C, H, W = x.shape
    xn = chan_layernorm(x, sd[p + '.norm.gamma'])
    seq = xn.permute(2, 3, 0, 1).reshape(H * W, Fr, C)                    # (hw, F, C)
    qkv = seq @ sd[p + '.fn.fn.to_qkv.weight'].t()
    q, k, v = qkv.reshape(H * W, Fr, 3, heads, dim_head).permute(2, 0, 3, 1, 4)  # (hw,h,F,d)
    q = q * dim_head ** -0.5
    q, k = rotary(q, freqs), rotary(k, freqs)
    if band is None:
        sim = torch.einsum('phid,phjd->phij', q, k) + bias
        sim = sim - sim.amax(dim=-1, keepdim=True)
        attn = sim.softmax(dim=-1)
        out = torch.einsum('phij,phjd->phid', attn, v)
    else:
        out = torch.empty_like(q)
        for i in range(Fr):
            lo, hi = max(0, i - band), min(Fr, i + band + 1)
            s = torch.einsum('phd,phjd->phj', q[:, :, i], k[:, :, lo:hi]) + bias[:, i, lo:hi]
            s = s - s.amax(dim=-1, keepdim=True)
            out[:, :, i] = torch.einsum('phj,phjd->phd', s.softmax(dim=-1), v[:, :, lo:hi])
    out = out.permute(0, 2, 1, 3).reshape(H * W, Fr, heads * dim_head)
    out = out @ sd[p + '.fn.fn.to_out.weight'].t()
    return out.reshape(H, W, Fr, C).permute(2, 3, 0, 1) + x


def mid_spatial_attention(sd, p, x, heads=8, dim_head=32):
    """U:841-843: full softmax attention over the h*w tokens of each frame, no rotary/bias."""
    Fr, C, H, W = x.shape
    xn = chan_layernorm(x, sd[p + '.norm.gamma'])
    tok = xn.permute(0, 2, 3, 1).reshape(Fr, H * W, C)
    qkv = tok @ sd[p + '.fn.fn.to_qkv.weight'].t()
    q, k, v = qkv.reshape(Fr, H * W, 3, heads, dim_head).permute(2, 0, 3, 1, 4)
    q = q * dim_head ** -0.5
    sim = torch.einsum('fhid,fhjd->fhij', q, k)
    sim = sim - sim.amax(dim=-1, keepdim=True)
    out = torch.einsum('fhij,fhjd->fhid', sim.softmax(dim=-1), v)
    out = out.permute(0, 2, 1, 3).reshape(Fr, H * W, heads * dim_head)
    out = out @ sd[p + '.fn.fn.to_out.weight'].t()
    return out.reshape(Fr, H, W, C).permute(0, 3, 1, 2) + x


# ----------------------------------------------------------------------------- the UNet

def unet_forward(sd, cfg, x, time, cond, null_cond_prob=0.0, band=None, taps=None):
    """U:892-956 Unet3D.forward for ONE clip.
    x (1,275,F,h,w) fp32; time (1,) int64; cond (1,F,1032).  Returns (1,3,F,h,w).
    band=None follows U (global scores + additive window mask); band=cfg.win follows UL.
    taps: optional dict filled with intermediate activations (NCFHW) for sub-module pinning."""
    assert x.shape[0] == 1, "oracle handles one clip at a time (batch elements are independent)"
    Fr = x.shape[2]
    if null_cond_prob == 1:                                   # U:920-926, null emb = zeros
        cond = torch.zeros_like(cond)
    elif null_cond_prob != 0:
        raise NotImplementedError("stochastic cond dropout is training-only")
    c = cond[0]
    xf = x[0].permute(1, 0, 2, 3)                             # (F, 275, h, w)

    def tap(name, t):
        if taps is not None:
            taps[name] = t.permute(1, 0, 2, 3).unsqueeze(0).clone()

    bias = rel_pos_bias(sd['time_rel_pos_bias.relative_attention_bias.weight'], Fr, cfg.win)  # U:908
    freqs = sd['init_temporal_attn.fn.fn.fn.rotary_emb.freqs']
    pad = cfg.init_k // 2
    xf = F.conv2d(xf, sd['init_conv.weight'][:, :, 0], sd['init_conv.bias'], padding=pad)      # U:910
    r = xf
    tap('init_conv', xf)
    xf = temporal_attention(sd, 'init_temporal_attn.fn', xf, bias, freqs, band=band)           # U:913
    tap('init_temporal_attn', xf)
    te = sinusoidal(time, cfg.dim)[0]
    te = F.gelu(te @ sd['time_mlp.1.weight'].t() + sd['time_mlp.1.bias'])
    te = te @ sd['time_mlp.3.weight'].t() + sd['time_mlp.3.bias']                              # U:915

    skips = []
    nres = len(cfg.in_out)
    for L in range(nres):                                                                      # U:934-940
        xf = resnet_block(sd, f'downs.{L}.0', xf, cfg, te, c)
        tap(f'downs.{L}.0', xf)
        xf = resnet_block(sd, f'downs.{L}.1', xf, cfg, te, c)
        tap(f'downs.{L}.1', xf)
        xf = spatial_linear_attention(sd, f'downs.{L}.2.fn', xf)
        tap(f'downs.{L}.2', xf)
        xf = temporal_attention(sd, f'downs.{L}.3.fn', xf, bias, freqs, band=band)
        tap(f'downs.{L}.3', xf)
        skips.append(xf)
        if L < nres - 1:
            xf = F.conv2d(xf, sd[f'downs.{L}.4.weight'][:, :, 0], sd[f'downs.{L}.4.bias'],
                          stride=2, padding=1)                                                 # U:175-176
            tap(f'downs.{L}.4', xf)
    xf = resnet_block(sd, 'mid_block1', xf, cfg, te, c)                                        # U:942-945
    tap('mid_block1', xf)
    xf = mid_spatial_attention(sd, 'mid_spatial_attn.fn', xf)
    tap('mid_spatial_attn', xf)
    xf = temporal_attention(sd, 'mid_temporal_attn.fn', xf, bias, freqs, band=band)
    tap('mid_temporal_attn', xf)
    xf = resnet_block(sd, 'mid_block2', xf, cfg, te, c)
    tap('mid_block2', xf)
    for K in range(nres):                                                                      # U:947-953
        xf = torch.cat((xf, skips.pop()), dim=1)
        xf = resnet_block(sd, f'ups.{K}.0', xf, cfg, te, c)
        tap(f'ups.{K}.0', xf)
        xf = resnet_block(sd, f'ups.{K}.1', xf, cfg, te, c)
        tap(f'ups.{K}.1', xf)
        xf = spatial_linear_attention(sd, f'ups.{K}.2.fn', xf)
        tap(f'ups.{K}.2', xf)
        xf = temporal_attention(sd, f'ups.{K}.3.fn', xf, bias, freqs, band=band)
        tap(f'ups.{K}.3', xf)
        if K < nres - 1:
            xf = F.conv_transpose2d(xf, sd[f'ups.{K}.4.weight'][:, :, 0], sd[f'ups.{K}.4.bias'],
                                    stride=2, padding=1)                                       # U:165-167
            tap(f'ups.{K}.4', xf)
    xf = torch.cat((xf, r), dim=1)                                                             # U:955
    outs = []
    for head in ('final_conv', 'occlusion_map'):                                               # U:956
        hd = resnet_block(sd, head + '.0', xf, cfg, None, None)      # heads: no time/cond MLPs (U:862,875)
        tap(head + '.0', hd)
        outs.append(F.conv2d(hd, sd[head + '.1.weight'][:, :, 0], sd[head + '.1.bias']))
    out = torch.cat(outs, dim=1)                                     # (F, 3, h, w)
    return out.permute(1, 0, 2, 3).unsqueeze(0).contiguous()


def forward_with_cond_scale(sd, cfg, x, time, cond, cond_scale=1.0, band=None):
    """U:879-890.  Batch elements are independent, so the batch is looped."""
    outs = []
    for b in range(x.shape[0]):
        xb, tb, cb = x[b:b + 1], time[b:b + 1], cond[b:b + 1]
        logits = unet_forward(sd, cfg, xb, tb, cb, 0.0, band)
        if cond_scale != 1:
            null = unet_forward(sd, cfg, xb, tb, cb, 1.0, band)
            logits = null + (logits - null) * cond_scale
        outs.append(logits)
    return torch.cat(outs, dim=0)


# ----------------------------------------------------------------------------- sampler (row a16)

def cosine_alphas_cumprod(timesteps=1000, s=0.008):
    """U:975-985 + U:1012-1016: returns (alphas_cumprod, alphas_cumprod_prev) as fp32."""
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.9999)
    acp = torch.cumprod(1. - betas, dim=0)
    prev = F.pad(acp[:-1], (1, 0), value=1.)
    return acp.float(), prev.float()


def ddim_time_pairs(total=1000, sampling=20):
    """U:1162-1164."""
    times = torch.linspace(0., total, steps=sampling + 2)[:-1]
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


def ddim_step(eps, img, t, t_next, noise, eta=1.0, dynamic_thres=True, pct=0.9):
    """U:1170-1205: one DDIM update given the UNet's eps prediction. noise: tensor or None."""
    acp, prev = cosine_alphas_cumprod()
    alpha, alpha_next = prev[t], prev[t_next]
    x0 = torch.sqrt(1. / acp)[t] * img - torch.sqrt(1. / acp - 1)[t] * eps      # U:1072-1076
    s = torch.ones(x0.shape[0])
    if dynamic_thres:
        s = torch.quantile(x0.reshape(x0.shape[0], -1).abs(), pct, dim=-1).clamp(min=1.)
    s = s.view(-1, *((1,) * (x0.ndim - 1)))
    x0 = x0.clamp(-s, s) / s
    sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
    c = ((1 - alpha_next) - sigma ** 2).sqrt()
    nz = noise if (t_next > 0 and noise is not None) else 0.
    return x0 * alpha_next.sqrt() + c * eps + sigma * nz

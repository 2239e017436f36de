"""TEST INFRASTRUCTURE — frame-sharded evaluation of the oracle over torch.distributed (gloo on CPU).

Checks the HOST-SIDE LOGIC of the exact multi-GPU partition without a GPU: rank r owns the contiguous
frames [r*F/N, (r+1)*F/N); before every temporal attention the +-win_width boundary frames of the layer input are
exchanged with the adjacent ranks (send/recv), every GroupNorm all-reduces its partial sums, positions (rotary, relative
bias) are global frame indices.  With those two exchanges the sharded forward equals the unsharded one — the same plan
the CUDA library implements over NCCL (csrc/unet.cu: temporal_attn, gn_allreduce).
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import unet_oracle as O


class Shard:
    def __init__(self, F_global, win):
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        assert F_global % self.world == 0
        self.Fl = F_global // self.world
        assert self.world == 1 or self.Fl >= win, "each rank must own at least win_width frames"
        self.F_global, self.win = F_global, win
        self.lo = self.rank * self.Fl
        self.halo_l = win if self.rank > 0 else 0
        self.halo_r = win if self.rank < self.world - 1 else 0

    def exchange(self, x):
        """x (Fl, C, H, W) -> (halo_l + Fl + halo_r, C, H, W) with the neighbours' boundary frames."""
        parts, reqs = [], []
        w = self.win
        left = torch.empty((w,) + tuple(x.shape[1:])) if self.halo_l else None
        right = torch.empty((w,) + tuple(x.shape[1:])) if self.halo_r else None
        if self.halo_l:
            reqs.append(dist.isend(x[:w].contiguous(), self.rank - 1))
            reqs.append(dist.irecv(left, self.rank - 1))
        if self.halo_r:
            reqs.append(dist.isend(x[-w:].contiguous(), self.rank + 1))
            reqs.append(dist.irecv(right, self.rank + 1))
        for r in reqs:
            r.wait()
        if left is not None:
            parts.append(left)
        parts.append(x)
        if right is not None:
            parts.append(right)
        return torch.cat(parts, dim=0)


def sharded_unet_forward(sd, cfg, x_local, time, cond_local, F_global):
    """x_local (1, 275, Fl, h, w), cond_local (1, Fl, 1032): this rank's frames.  Returns the local eps (1, 3, Fl, h, w)."""
    sh = Shard(F_global, cfg.win)
    orig_gn, orig_ta = O.clip_groupnorm, O.temporal_attention

    def gn(x, groups, w, b, eps=1e-5):
        Fr, C, H, W = x.shape
        xg = x.permute(1, 0, 2, 3).reshape(groups, -1).double()
        st = torch.stack([xg.sum(dim=1), (xg * xg).sum(dim=1)])
        dist.all_reduce(st)                                            # 16 doubles per norm, like the CUDA path
        n = float(sh.F_global * H * W * (C // groups))
        mean = st[0] / n
        var = st[1] / n - mean * mean
        rstd = 1.0 / torch.sqrt(var + eps)
        m = mean.float().repeat_interleave(C // groups).reshape(1, C, 1, 1)
        r = rstd.float().repeat_interleave(C // groups).reshape(1, C, 1, 1)
        return (x - m) * r * w.reshape(1, C, 1, 1) + b.reshape(1, C, 1, 1)

    def ta(sd_, p, x, bias, freqs, heads=8, dim_head=32, band=None):
        Fr, C, H, W = x.shape
        xe = sh.exchange(x)
        Fe = xe.shape[0]
        xn = O.chan_layernorm(xe, sd_[p + '.norm.gamma'])
        seq = xn.permute(2, 3, 0, 1).reshape(H * W, Fe, C)
        qkv = seq @ sd_[p + '.fn.fn.to_qkv.weight'].t()
        q, k, v = qkv.reshape(H * W, Fe, 3, heads, dim_head).permute(2, 0, 3, 1, 4)
        q = q * dim_head ** -0.5
        pos0 = sh.lo - sh.halo_l                                       # global frame index of xe[0]
        ang = (torch.arange(Fe, dtype=q.dtype) + pos0)[:, None] * freqs[None, :]
        ang = ang.repeat_interleave(2, dim=-1)

        def rot(t):
            t2 = t.reshape(*t.shape[:-1], -1, 2)
            r = torch.stack((-t2[..., 1], t2[..., 0]), dim=-1).reshape(t.shape)
            return t * ang.cos() + r * ang.sin()

        q, k = rot(q), rot(k)
        gb = bias[:, pos0 + sh.halo_l:pos0 + sh.halo_l + Fr, pos0:pos0 + Fe]   # (heads, own queries, local keys), global indices
        sim = torch.einsum('phid,phjd->phij', q[:, :, sh.halo_l:sh.halo_l + Fr], k) + gb
        sim = sim - sim.amax(dim=-1, keepdim=True)
        out = torch.einsum('phij,phjd->phid', sim.softmax(dim=-1), v)
        out = out.permute(0, 2, 1, 3).reshape(H * W, Fr, heads * dim_head)
        out = out @ sd_[p + '.fn.fn.to_out.weight'].t()
        return out.reshape(H, W, Fr, C).permute(2, 3, 0, 1) + x

    O.clip_groupnorm, O.temporal_attention = gn, ta
    try:
        # the relative-position bias table is indexed with GLOBAL frame numbers inside `ta`
        orig_bias = O.rel_pos_bias
        O.rel_pos_bias = lambda emb_w, n, window: orig_bias(emb_w, sh.F_global, window)
        try:
            return O.unet_forward(sd, cfg, x_local, time, cond_local)
        finally:
            O.rel_pos_bias = orig_bias
    finally:
        O.clip_groupnorm, O.temporal_attention = orig_gn, orig_ta


def sharded_dynamic_threshold(x0_local, q):
    """TEST INFRASTRUCTURE — host-side logic of `dawn_unet_ddim_step` on a frame-sharded clip (csrc/sampler.cu,
    SURVEY 8e-iii): s = max(1, torch.quantile(|x0|.flatten(), q)) over the WHOLE clip (reference U:1186-1193) without
    gathering x0.  Non-negative fp32 values order like their bit patterns, so the order statistic `lo = floor(q*(n-1))` is
    found by a 4-pass radix select over 8-bit digits whose 256-bin histograms are all-reduced; one more pass gives the
    count of keys <= v[lo] and the smallest key above it (all-reduced sum / min), enough for torch's linear interpolation."""
    import numpy as np
    keys = x0_local.detach().abs().float().contiguous().view(-1).numpy().view(np.uint32)
    n_local = keys.size
    n_global = n_local * dist.get_world_size()
    rank_f = np.float32(q) * np.float32(n_global - 1)                 # ATen evaluates the rank in the input dtype
    lo, hi = int(np.floor(rank_f)), int(np.ceil(rank_f))
    w = np.float32(rank_f - np.floor(rank_f))
    prefix, mask, rem = np.uint32(0), np.uint32(0), lo
    for shift in (24, 16, 8, 0):
        sel = keys[(keys & mask) == prefix]
        hist = torch.from_numpy(np.bincount((sel >> np.uint32(shift)) & np.uint32(255), minlength=256).astype(np.int64))
        dist.all_reduce(hist)
        cum = 0
        for b in range(256):
            if cum + int(hist[b]) > rem:
                break
            cum += int(hist[b])
        rem -= cum
        prefix = np.uint32(prefix | np.uint32(b << shift))
        mask = np.uint32(mask | np.uint32(255 << shift))
    count_le = torch.tensor([int((keys <= prefix).sum())], dtype=torch.int64)
    above = keys[keys > prefix]
    min_gt = torch.tensor([int(above.min()) if above.size else 0xFFFFFFFF], dtype=torch.int64)
    dist.all_reduce(count_le)
    dist.all_reduce(min_gt, op=dist.ReduceOp.MIN)
    vlo = np.array([prefix], dtype=np.uint32).view(np.float32)[0]
    vhi = vlo
    if hi > lo and int(count_le) < lo + 2:
        vhi = np.array([int(min_gt)], dtype=np.uint32).view(np.float32)[0]
    d = np.float32(vhi - vlo)
    qv = np.float32(vlo + w * d) if w < 0.5 else np.float32(vhi - d * (np.float32(1) - w))      # at::lerp
    return float(max(qv, np.float32(1.0))), float(qv)

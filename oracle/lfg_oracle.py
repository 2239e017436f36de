"""TEST INFRASTRUCTURE — CPU restatement of DAWN's LFG flow decoder, the stage that turns the
denoised latent flow / occlusion maps into video frames.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline may import this; the product path never does.

Restates, batched over frames (the reference loops over frames with batch 1, FD:375-383):
  LFG/modules/generator.py:132-136  Generator.compute_fea          (source-image features fed to the diffusion UNet)
  LFG/modules/generator.py:138-171  Generator.forward_with_flow    (decode one frame from flow + occlusion)
  LFG/modules/generator.py:59-69    deform_input                   (bilinear flow resize + grid_sample)
  LFG/modules/generator.py:71-90    apply_optical                  (warp skip, bilinear occlusion resize, blend)
  LFG/modules/util.py:70-150        ResBlock2d / UpBlock2d / DownBlock2d / SameBlock2d
  LFG/sync_batchnorm/batchnorm.py:50-53  eval-mode BatchNorm == F.batch_norm with running statistics
Pinned against the real reference by oracle/make_golden_lfg.py (golden vectors under tests/golden/lfg_*.npz).
"""
import torch
import torch.nn.functional as F


class LfgCfg:
    """generator_params of config/hdtf128.yaml:82-93 == config/hdtf256.yaml:82-93."""

    def __init__(self, num_channels=3, block_expansion=64, max_features=512, num_down_blocks=2, num_bottleneck_blocks=6,
                 skips=True):
        self.num_channels, self.block_expansion, self.max_features = num_channels, block_expansion, max_features
        self.num_down_blocks, self.num_bottleneck_blocks, self.skips = num_down_blocks, num_bottleneck_blocks, skips

    def down_features(self, i):                                   # generator.py:40-43
        return (min(self.max_features, self.block_expansion * 2 ** i), min(self.max_features, self.block_expansion * 2 ** (i + 1)))

    def up_features(self, i):                                     # generator.py:47-50
        n = self.num_down_blocks
        return (min(self.max_features, self.block_expansion * 2 ** (n - i)), min(self.max_features, self.block_expansion * 2 ** (n - i - 1)))


def state_dict_schema(cfg=None):
    """(name, shape) of every Generator parameter/buffer the decode path reads (the reference's `generator` checkpoint entry
    also holds `pixelwise_flow_predictor.*`, which forward_with_flow never touches: generator.py:138-171)."""
    cfg = cfg or LfgCfg()
    out = []

    def conv_bn(prefix, ci, co, k):
        out.append((f"{prefix}.conv.weight", (co, ci, k, k)))
        out.append((f"{prefix}.conv.bias", (co,)))
        bn(f"{prefix}.norm", co)

    def bn(prefix, c):
        out.extend([(f"{prefix}.weight", (c,)), (f"{prefix}.bias", (c,)), (f"{prefix}.running_mean", (c,)),
                    (f"{prefix}.running_var", (c,)), (f"{prefix}.num_batches_tracked", ())])

    conv_bn("first", cfg.num_channels, cfg.block_expansion, 7)
    for i in range(cfg.num_down_blocks):
        ci, co = cfg.down_features(i)
        conv_bn(f"down_blocks.{i}", ci, co, 3)
    for i in range(cfg.num_down_blocks):
        ci, co = cfg.up_features(i)
        conv_bn(f"up_blocks.{i}", ci, co, 3)
    cb = cfg.down_features(cfg.num_down_blocks - 1)[1]
    for i in range(cfg.num_bottleneck_blocks):
        p = f"bottleneck.r{i}"
        out.extend([(f"{p}.conv1.weight", (cb, cb, 3, 3)), (f"{p}.conv1.bias", (cb,)),
                    (f"{p}.conv2.weight", (cb, cb, 3, 3)), (f"{p}.conv2.bias", (cb,))])
        bn(f"{p}.norm1", cb)
        bn(f"{p}.norm2", cb)
    out.extend([("final.weight", (cfg.num_channels, cfg.block_expansion, 7, 7)), ("final.bias", (cfg.num_channels,))])
    return out


# ----------------------------------------------------------------------------- primitives
def bn_eval(sd, p, x):
    """LFG/sync_batchnorm/batchnorm.py:50-53 in eval mode: running statistics, eps 1e-5."""
    return F.batch_norm(x, sd[f"{p}.running_mean"], sd[f"{p}.running_var"], sd[f"{p}.weight"], sd[f"{p}.bias"], False, 0.1, 1e-5)


def conv_bn_relu(sd, p, x, pad):
    """SameBlock2d / the conv part of Down/UpBlock2d (util.py:107-110, 127-130, 147-150)."""
    return F.relu(bn_eval(sd, f"{p}.norm", F.conv2d(x, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"], padding=pad)))


def res_block(sd, p, x):
    """ResBlock2d.forward util.py:85-93: pre-activation (norm -> relu -> conv) twice, then += x."""
    out = F.conv2d(F.relu(bn_eval(sd, f"{p}.norm1", x)), sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)
    out = F.conv2d(F.relu(bn_eval(sd, f"{p}.norm2", out)), sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    return out + x


def deform_input(inp, flow):
    """generator.py:59-69.  inp (F, C, H, W); flow (F, h, w, 2) in [-1, 1] (x, y).  The flow is resized bilinearly
    (align_corners=False, torch default) when its resolution differs; grid_sample defaults: bilinear, zeros, align_corners=False."""
    _, h_old, w_old, _ = flow.shape
    _, _, h, w = inp.shape
    if h_old != h or w_old != w:
        flow = F.interpolate(flow.permute(0, 3, 1, 2), size=(h, w), mode='bilinear', align_corners=False).permute(0, 2, 3, 1)
    return F.grid_sample(inp, flow, mode='bilinear', padding_mode='zeros', align_corners=False)


def apply_optical(prev, skip, flow, occ):
    """generator.py:71-90 with motion_params = {optical_flow, occlusion_map}."""
    skip = deform_input(skip, flow)
    if skip.shape[2] != occ.shape[2] or skip.shape[3] != occ.shape[3]:
        occ = F.interpolate(occ, size=skip.shape[2:], mode='bilinear', align_corners=False)
    if prev is not None:
        return skip * occ + prev * (1 - occ)
    return skip * occ


# ----------------------------------------------------------------------------- the two entry points
def encode_source(sd, cfg, source):
    """first + down blocks on the source image (generator.py:140-146): per clip, not per frame.  Returns [skip0, ..., skipN]."""
    out = conv_bn_relu(sd, "first", source, 3)
    skips = [out]
    for i in range(cfg.num_down_blocks):
        out = F.avg_pool2d(conv_bn_relu(sd, f"down_blocks.{i}", out, 1), 2)          # util.py:126-131
        skips.append(out)
    return skips


def compute_fea(sd, cfg, source):
    """generator.py:132-136."""
    return encode_source(sd, cfg, source)[-1]


def forward_with_flow(sd, cfg, source, flow, occ, taps=None):
    """generator.py:138-171 for all frames at once.  source (1, 3, H, W) in [0, 1]; flow (F, h, w, 2); occ (F, 1, h, w).
    Returns dict(prediction (F, 3, H, W), deformed (F, 3, H, W))."""
    nf = flow.shape[0]
    skips = [s.expand(nf, -1, -1, -1) for s in encode_source(sd, cfg, source)]
    src = source.expand(nf, -1, -1, -1)
    deformed = deform_input(src, flow)                                               # generator.py:152
    out = apply_optical(None, skips[-1], flow, occ)                                  # :154
    if taps is not None:
        taps["warp0"] = out
    for i in range(cfg.num_bottleneck_blocks):                                       # :156
        out = res_block(sd, f"bottleneck.r{i}", out)
    if taps is not None:
        taps["bottleneck"] = out
    for i in range(cfg.num_down_blocks):                                             # :157-160
        if cfg.skips:
            out = apply_optical(out, skips[-(i + 1)], flow, occ)
        out = conv_bn_relu(sd, f"up_blocks.{i}", F.interpolate(out, scale_factor=2), 1)   # util.py:106-111 (nearest)
        if taps is not None:
            taps[f"up{i}"] = out
    if cfg.skips:
        out = apply_optical(out, skips[0], flow, occ)                                # :161-162
    out = torch.sigmoid(F.conv2d(out, sd["final.weight"], sd["final.bias"], padding=3))   # :163-164
    if cfg.skips:
        out = apply_optical(out, src, flow, occ)                                     # :166-167
    return {"prediction": out, "deformed": deformed}

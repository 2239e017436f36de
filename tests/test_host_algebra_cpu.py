"""CPU: the algebraic rewrites the CUDA paths rely on, checked with plain torch ops (host-side logic, no GPU):
 * nearest-2x upsample + 3x3 conv == four output-parity 2x2 convs over the low-resolution grid with summed taps
   (csrc/lfg.cu::pack_up, reference UpBlock2d LFG/modules/util.py:106-111);
 * a k x k conv == the sum of k row-convolutions (1 x k) over vertically shifted copies of the input
   (csrc/unet.cu::init_map, the per-clip part of the 7x7 init conv, reference U:776-777);
 * the init conv is linear in its input channels: conv(cat[x_t, fea]) == conv3(x_t) + map(fea)  (SURVEY a2, U:1167, 1177);
 * eval-mode BatchNorm after a conv folds into the conv's weights and bias (csrc/lfg.cu::pack_conv).
"""
import torch
import torch.nn.functional as F

K_UP_OFF = [[-1, 0], [0, 1]]                       # csrc/lfg.cu: kUpOff[parity][tap] -> low-resolution offset


def up_in_set(parity, tap, k):                      # csrc/lfg.cu: up_in_set
    return (k == 0 if tap == 0 else k >= 1) if parity == 0 else (k <= 1 if tap == 0 else k == 2)


def test_upsample_conv_equals_parity_class_convs():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 6, 7, generator=g)
    w = torch.randn(4, 5, 3, 3, generator=g)
    b = torch.randn(4, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2), w, b, padding=1)
    out = torch.empty_like(ref)
    H, W = x.shape[-2:]
    xp = F.pad(x, (1, 1, 1, 1))
    for py in range(2):
        for px in range(2):
            acc = b.view(1, -1, 1, 1).expand(2, 4, H, W).clone()
            for ty in range(2):
                for tx in range(2):
                    wsum = sum(w[:, :, ky, kx] for ky in range(3) for kx in range(3) if up_in_set(py, ty, ky) and up_in_set(px, tx, kx))
                    dy, dx = K_UP_OFF[py][ty], K_UP_OFF[px][tx]
                    patch = xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
                    acc = acc + torch.einsum('oc,bchw->bohw', wsum, patch)
            out[:, :, py::2, px::2] = acc
    assert (out - ref).abs().max().item() < 1e-4


def test_square_conv_equals_sum_of_row_convs_over_shifted_copies():
    g = torch.Generator().manual_seed(1)
    k, pad = 7, 3
    x = torch.randn(1, 6, 12, 10, generator=g)
    w = torch.randn(8, 6, k, k, generator=g)
    b = torch.randn(8, generator=g)
    ref = F.conv2d(x, w, b, padding=pad)
    H = x.shape[2]
    total = b.view(1, -1, 1, 1)
    for ky in range(k):
        shifted = torch.zeros_like(x)                                  # copy ky: the frame shifted by ky - pad rows, zero outside
        lo, hi = max(0, pad - ky), min(H, H + pad - ky)
        shifted[:, :, lo:hi] = x[:, :, lo + ky - pad:hi + ky - pad]
        total = total + F.conv2d(shifted, w[:, :, ky:ky + 1, :], None, padding=(0, pad))
    assert (total - ref).abs().max().item() < 1e-4


def test_init_conv_is_linear_in_the_frame_invariant_channels():
    g = torch.Generator().manual_seed(2)
    Fr, h, w_ = 5, 8, 8
    x_t = torch.randn(1, 3, Fr, h, w_, generator=g)
    fea = torch.relu(torch.randn(1, 9, h, w_, generator=g))
    wt = torch.randn(4, 12, 1, 7, 7, generator=g)
    b = torch.randn(4, generator=g)
    x = torch.cat([x_t, fea.unsqueeze(2).expand(-1, -1, Fr, -1, -1)], dim=1)
    ref = F.conv3d(x, wt, b, padding=(0, 3, 3))
    fmap = F.conv2d(fea, wt[:, 3:, 0], b, padding=3)                   # per clip: (1, 4, h, w), bias included
    live = F.conv3d(x_t, wt[:, :3], None, padding=(0, 3, 3))           # per step: 3 live channels
    assert (live + fmap.unsqueeze(2) - ref).abs().max().item() < 1e-4


def test_batchnorm_after_conv_folds_into_the_conv():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 9, 9, generator=g)
    w, b = torch.randn(5, 6, 3, 3, generator=g), torch.randn(5, generator=g)
    gamma, beta = 1 + 0.2 * torch.randn(5, generator=g), 0.1 * torch.randn(5, generator=g)
    rm, rv = 0.2 * torch.randn(5, generator=g), 1 + 0.3 * torch.rand(5, generator=g)
    ref = F.batch_norm(F.conv2d(x, w, b, padding=1), rm, rv, gamma, beta, False, 0.1, 1e-5)
    s = gamma / torch.sqrt(rv + 1e-5)
    t = beta - rm * s
    out = F.conv2d(x, w * s.view(-1, 1, 1, 1), b * s + t, padding=1)
    assert (out - ref).abs().max().item() < 1e-4

"""CPU: the work decomposition of the tcgen05 temporal-attention kernel (csrc/temporal_tc.cuh: segments of a frame sequence, two
row tiles per segment, key-column windows) evaluated on the host through the C-ABI and checked by brute force against what the
attention needs (reference LA:71-99: query i sees keys [i - w, i + w] inside the sequence): every query frame is produced exactly
once, and all of its band keys lie inside the window, inside the tile's 160-column score tile and inside the 128 columns the owning
softmax warp reads; operand offsets respect the 8-row / 16-key alignment of the shared-memory descriptors."""
import ctypes

import pytest

from dawn_pytorch_b200 import _lib

WMAX, SN, CW, TB, TZ = 240, 160, 128, 512, 240


def plan(F, band, qlo, qhi):
    out = (ctypes.c_int * (14 * 16))()
    n = _lib.lib.dawn_temporal_tc_plan(F, band, qlo, qhi, out)
    return [list(out[14 * s:14 * s + 14]) for s in range(n)]


def check(F, band, qlo, qhi):
    segs = plan(F, band, qlo, qhi)
    assert segs, (F, band, qlo, qhi)
    seen = set()
    for w0, wn, qa, qb, *tiles in segs:
        assert 1 <= wn <= WMAX and w0 >= 0 and w0 + wn <= F
        for j in range(2):
            r0, r1, q0, q1, kb = tiles[5 * j:5 * j + 5]
            assert 0 <= r1 - r0 <= 128 and r0 % 8 == 0 and kb % 16 == 0 and 0 <= kb and kb + SN <= WMAX
            for r in range(q0, q1):
                f = w0 + r
                assert f not in seen
                seen.add(f)
                klo, khi = max(0, f - band), min(F - 1, f + band)
                assert w0 + kb <= klo and khi < w0 + kb + SN and khi < w0 + wn
            if q1 > q0:
                for wq in range(4):
                    xw = max(r0 + 32 * wq - band - kb, 0)
                    cstart = min(xw & ~15, SN - CW)
                    for lane in range(32):
                        row = r0 + 32 * wq + lane
                        ib = kb + cstart - min(row, WMAX - 1) + TZ
                        assert 0 <= ib and ib + CW - 1 < TB
                        if q0 <= row < q1:
                            assert kb + cstart <= max(0, row - band) and min(wn - 1, row + band) < kb + cstart + CW
    assert seen == set(range(qlo, qhi))


@pytest.mark.parametrize("band", [1, 8, 39, 40])
def test_every_query_is_covered_once_with_all_its_keys_on_chip(band):
    for F in list(range(1, 300)) + [399, 400, 401, 512, 800, 1000]:
        qs = [(0, F)]
        if F > 2 * band + 1:
            qs.append((band, F - band))            # interior shard: halos on both sides
        if F > band + 1:
            qs += [(band, F), (0, F - band)]       # first / last shard
        if F > 5:
            qs.append((3, F - 2))
        for a, b in qs:
            if b > a:
                check(F, band, a, b)


def test_unsupported_shapes_are_refused():
    assert plan(100, 41, 0, 100) == [] and plan(100, 0, 0, 100) == [] and plan(10, 40, 5, 5) == [] and plan(10, 40, 0, 11) == []
    assert plan(10 ** 5, 40, 0, 10 ** 5) == []          # more than 16 segments per pixel: the caller uses the unfused path

"""CPU: static check of the built library's SASS (cuobjdump, no GPU needed).  The tcgen05 kernels must contain tensor-core MMAs
(UTCHMMA) and TMEM loads (LDTM), and their MMA issue paths must stay free of the per-lane uniform-register loops (`BRA.U.ANY`) that
ptxas emits when an operand is not provably warp-uniform: issuing from inside `if (lane == 0)` cost ~75 cycles per MMA against 16
cycles of tensor-pipe time (profiles/r2_h_ttc_trace.md).  The loaders (cp.async.bulk / TMA, one elected thread) may keep
theirs: at most one loop per bulk copy / tensor load."""
import collections
import os
import re
import shutil
import subprocess

import pytest

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dawn_pytorch_b200", "libdawn_unet.so")


@pytest.fixture(scope="module")
def sass_counts():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    cur, cnt = None, {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln) if cur else None
        if m:
            for key in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "BRA.U.ANY"):
                if m.group(1).startswith(key):
                    cnt[cur][key] += 1
    return cnt


def kernels(cnt, name):
    return {k: v for k, v in cnt.items() if name in k}


def test_temporal_attention_kernel_issues_tcgen05_from_uniform_registers(sass_counts):
    ks = kernels(sass_counts, "temporal_tc_kernel")
    assert len(ks) == 2                                   # traced and product instantiations
    for k, c in ks.items():
        assert c["UTCHMMA"] >= 39 and c["LDTM"] > 0 and c["STTM"] > 0 and c["UBLKCP"] > 0, (k, dict(c))
        assert c["BRA.U.ANY"] == 0, (k, dict(c))


@pytest.mark.parametrize("name", ["tc_gemm_kernel", "tc_conv3_kernel"])
def test_gemm_and_halo_conv_kernels_keep_their_mma_issue_loop_free(sass_counts, name):
    ks = kernels(sass_counts, name)
    assert ks
    for k, c in ks.items():
        assert c["UTCHMMA"] >= 12 and c["LDTM"] > 0 and c["UBLKCP"] > 0, (k, dict(c))
        assert c["BRA.U.ANY"] <= c["UBLKCP"] + c["UTMALDG"], (k, dict(c))      # only the loaders' copies; none per MMA (r2-g: ~2 per MMA)


def test_tma_fed_halo_conv_uses_tensor_loads(sass_counts):
    assert any(c["UTMALDG"] > 0 for c in kernels(sass_counts, "tc_conv3_kernel").values())

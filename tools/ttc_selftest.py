"""GPU: tcgen05 temporal attention (temporal_tc.cu) stage by stage against a host computation and against the mma.sync kernel;
each case in its own time-boxed subprocess (a hung kernel must not take the whole call down)."""
import ctypes
import subprocess
import sys

CASES = [  # F, P, band, q_lo, q_hi
    (16, 4, 40, 0, 16),
    (96, 8, 40, 0, 96),
    (200, 16, 40, 0, 200),
    (23, 3, 40, 0, 23),
    (1, 2, 40, 0, 1),
    (81, 5, 40, 0, 81),
    (64, 9, 8, 0, 64),
    (180, 6, 40, 40, 140),       # shard with halos on both sides (100 owned frames)
    (112, 300, 40, 0, 112),      # more units than SMs
    (305, 300, 40, 40, 265),     # two segments of 112 / 113 queries
    (224, 6, 40, 0, 224),
    (240, 6, 40, 0, 240),        # the largest single window
    (240, 5, 40, 40, 240),       # edge shard: 200 owned frames + left halo
    (240, 5, 40, 0, 200),        # edge shard: right halo
    (241, 5, 40, 0, 241),        # one frame more: two segments
    (280, 300, 40, 40, 240),     # two segments, more units than SMs
    (400, 7, 40, 0, 400),        # long single-GPU clip: three segments
    (200, 4096, 40, 0, 200),     # the bench shape of one level-0 layer
    (280, 4096, 40, 40, 240),    # interior shard of a long clip: 2 x 100-query segments per pixel
    (40, 4096, 40, 0, 40),       # short clip
]


def one(args):
    sys.path.insert(0, ".")
    from dawn_pytorch_b200 import _lib
    err = (ctypes.c_float * 6)()
    mr = ctypes.c_float()
    tr = (ctypes.c_uint64 * 48)()
    ms = ctypes.c_float()
    rc = _lib.lib.dawn_selftest_temporal_tc(*args, err, ctypes.byref(mr), tr, ctypes.byref(ms))
    print(f"ttc {args}: rc={rc} proj_rel={err[0]:.2e} S={err[1]:.2e} O={err[2]:.2e} out_pix0={err[3]:.2e} vs_mma_sync={err[4]:.2e} "
          f"nan={int(err[5])} max|ref|={mr.value:.2f} {ms.value:.3f} ms" + ("" if rc == 0 else " ERR " + _lib.lib.dawn_last_error().decode()), flush=True)
    if args[1] >= 1024:
        t = list(tr)
        names_wg = ["prologue+x_wait", "proj_wait", "E1 ld + kv_free wait", "E1 compute+arrive", "s_wait(+table bar)", "E2", "o_wait", "E3", "y_wait", "epilogue"]
        names_mma = ["x_wait", "wq_wait", "proj issue", "kv_wait", "S issue", "p_wait0", "p_wait1", "PV issue", "wo_wait", "oh_wait0", "oh_wait1", "Y issue"]
        its = max(t[45], 1)
        print(f"   CTA0: {its} head iterations, MMA thread total {t[44]} cycles = {t[44] / its:.0f} per head")
        for j in (0, 1):
            print(f"   WG{j} cycles/head: " + ", ".join(f"{n} {t[16 * j + i] / its:.0f}" for i, n in enumerate(names_wg)))
        print("   MMA cycles/head: " + ", ".join(f"{n} {t[32 + i] / its:.0f}" for i, n in enumerate(names_mma)))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(tuple(int(a) for a in sys.argv[1:]))
    else:
        # page the image in first (the first import on a fresh box can take a minute): the per-case time boxes below are for kernels
        subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, '.'); import torch, dawn_pytorch_b200; torch.zeros(1, device='cuda')"],
                       timeout=400)
        for c in CASES:
            try:
                r = subprocess.run([sys.executable, __file__, *map(str, c)], timeout=60, capture_output=True, text=True)
                print(r.stdout.strip() or ("NO OUTPUT rc=%d %s" % (r.returncode, r.stderr[-400:])), flush=True)
            except subprocess.TimeoutExpired:
                print(f"ttc {c}: TIMEOUT (hang) - stopping", flush=True)
                break

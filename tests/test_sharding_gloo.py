"""CPU, world_size 2 over gloo: the exact frame-sharding plan (halo exchange per temporal attention + GroupNorm
all-reduce, global positions) reproduces the unsharded oracle.  Host-side logic of the N > 1 path (no GPU needed)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import json
        from oracle import sharded as S
        from oracle import unet_oracle as O
        from oracle import weights as W
        with open(os.path.join(ROOT, "tests", "golden", "state_dict_schema.json")) as f:
            schema = json.load(f)
        sd = W.synth_state_dict([(n, tuple(s)) for n, s in schema["entries"]])
        Fg, h, w, t = 96, 8, 8, 952                       # the 'band' golden clip: 48 frames per rank >= window 40
        x_t, fea, cond = W.synth_inputs("band", Fg, h, w)
        x = torch.cat([x_t, fea.unsqueeze(2).expand(-1, -1, Fg, -1, -1)], dim=1)
        Fl = Fg // world
        lo = rank * Fl
        with torch.no_grad():
            out = S.sharded_unet_forward(sd, O.UnetCfg(), x[:, :, lo:lo + Fl].contiguous(), torch.full((1,), t),
                                         cond[:, lo:lo + Fl].contiguous(), Fg)
        parts = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(parts, out)
        if rank == 0:
            full = torch.cat(parts, dim=2)
            ref = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "band.npz"))["eps"])
            q.put(float(((full - ref).abs() / (1e-4 + 1e-3 * ref.abs())).max()))
    finally:
        dist.destroy_process_group()


def test_frame_sharded_oracle_equals_reference_golden_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    over_tol = q.get(timeout=5)
    print("sharded (world 2) vs reference golden: x tol =", over_tol)
    assert over_tol < 0.5


def test_partition_plan_arithmetic():
    """frame ranges, halo sizes and global rotary offsets used by dawn_unet_init_shard (csrc/unet.cu)"""
    win = 40
    for world, Fl in ((2, 48), (4, 100), (8, 100)):
        covered = []
        for r in range(world):
            lo, hi = r * Fl, (r + 1) * Fl
            hl = win if r > 0 else 0
            hr = win if r < world - 1 else 0
            pos0 = lo - hl
            assert pos0 >= 0 and hi + hr <= world * Fl
            # every key a local query may attend (|i-j| <= win) is inside the halo-extended range
            assert max(0, lo - win) >= pos0 and min(world * Fl, hi + win) <= hi + hr
            covered += list(range(lo, hi))
        assert covered == list(range(world * Fl))


def _quantile_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sharded as S
        res = []
        for seed, n_local, scale, qq in ((0, 3 * 48 * 8 * 8, 1.7, 0.9), (1, 1000, 0.2, 0.9), (2, 7, 3.0, 0.5), (3, 4096, 1.0, 0.999)):
            g = torch.Generator().manual_seed(seed)
            full = torch.randn(world * n_local, generator=g) * scale
            if seed == 2:
                full[:] = full[0]                          # all keys equal: ties on every digit
            s, qv = S.sharded_dynamic_threshold(full[rank * n_local:(rank + 1) * n_local], qq)
            ref = torch.quantile(full.abs(), qq)
            res.append((s, qv, float(ref), float(ref.clamp(min=1.0))))
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


def test_sharded_dynamic_threshold_equals_torch_quantile_world2():
    """Row a16 under sharding: the all-reduced radix select gives torch.quantile's value bit for bit."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_quantile_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for s, qv, ref, ref_s in q.get(timeout=5):
        assert qv == ref, (qv, ref)
        assert s == ref_s


def _worker3(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import json
        from oracle import sharded as S
        from oracle import unet_oracle as O
        from oracle import weights as W
        with open(os.path.join(ROOT, "tests", "golden", "state_dict_schema.json")) as f:
            schema = json.load(f)
        sd = W.synth_state_dict([(n, tuple(s)) for n, s in schema["entries"]])
        Fg, h, w, t = 120, 8, 8, 476                      # 40 frames per rank == the window: the middle rank has BOTH halos
        x_t, fea, cond = W.synth_inputs("band3", Fg, h, w)
        x = torch.cat([x_t, fea.unsqueeze(2).expand(-1, -1, Fg, -1, -1)], dim=1)
        Fl = Fg // world
        lo = rank * Fl
        with torch.no_grad():
            out = S.sharded_unet_forward(sd, O.UnetCfg(), x[:, :, lo:lo + Fl].contiguous(), torch.full((1,), t),
                                         cond[:, lo:lo + Fl].contiguous(), Fg)
        parts = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(parts, out)
        if rank == 0:
            with torch.no_grad():
                ref = O.unet_forward(sd, O.UnetCfg(), x, torch.full((1,), t), cond)
            full = torch.cat(parts, dim=2)
            q.put(float(((full - ref).abs() / (1e-4 + 1e-3 * ref.abs())).max()))
    finally:
        dist.destroy_process_group()


def test_frame_sharded_oracle_interior_rank_world3():
    """world_size 3, 40 frames per rank: rank 1 exchanges halos on both sides (the N >= 3 case of SURVEY 8e)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker3, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    over_tol = q.get(timeout=5)
    print("sharded (world 3) vs unsharded oracle: x tol =", over_tol)
    assert over_tol < 0.5

"""2+ GPUs (torchrun): exact frame sharding of one clip vs the reference golden / the unsharded CUDA path."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import weights as W            # noqa: E402
from tests import gpu_common as G          # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    from dawn_pytorch_b200 import DynamicNfUnet3D
    net = DynamicNfUnet3D(**G.CTOR).eval()
    net.load_state_dict(G.synth_sd(), strict=True)
    net = net.to(dev)
    cases = [("band", 96, 8, 8, 952, True), ("shardbig", 80 * world, 32, 32, 500, False)]
    for name, Fg, h, w, t, has_golden in cases:
        if Fg % world or Fg // world < 40:
            continue
        Fl, lo = Fg // world, rank * (Fg // world)
        x_t, fea, cond = W.synth_inputs(name, Fg, h, w)
        tt = torch.full((1,), t, dtype=torch.long, device=dev)
        net.update_num_frames(Fl)
        net.init_shard(Fl, h, w, dev)
        net.set_clip_invariants(fea[0].to(dev), cond[0, lo:lo + Fl].contiguous().to(dev))
        out = net.forward_x3(x_t[0, :, lo:lo + Fl].contiguous().to(dev), tt).clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); dist.barrier(); e0.record()
        for _ in range(3):
            net.forward_x3(x_t[0, :, lo:lo + Fl].contiguous().to(dev), tt)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        parts = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(parts, out)
        full = torch.cat(parts, dim=1).cpu()[None]
        if rank == 0:
            if has_golden:
                ref = torch.from_numpy(G.golden(name)["eps"])
                print(f"[{name}] sharded x{world} vs reference golden: over_tol {G.over_tol(full, ref):.3f}", flush=True)
            # unsharded run of the whole clip on rank 0 with a second module instance
            net1 = DynamicNfUnet3D(**G.CTOR).eval()
            net1.load_state_dict(G.synth_sd(), strict=True)
            net1 = net1.to(dev)
            net1.update_num_frames(Fg)
            net1.set_clip_invariants(fea[0].to(dev), cond[0].to(dev))
            one = net1.forward_x3(x_t[0].to(dev), tt)
            torch.cuda.synchronize()
            print(f"[{name}] F={Fg} {h}x{w}: sharded x{world} vs single-GPU CUDA: over_tol {G.over_tol(full, one.cpu()[None]):.4f}"
                  f"  max|d| {(full[0] - one.cpu()).abs().max():.2e};  sharded step {ms:.2f} ms", flush=True)
            del net1
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

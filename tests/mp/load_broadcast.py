"""world-size-N check of sdxl_unet_load_broadcast (run under torchrun, one rank per GPU): rank 0 owns the weight pack, the
library broadcasts it over its own NCCL call, every rank runs a forward on its own input and compares with the CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "stable-diffusion-xl-burn_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

import sdxl_b200
from sdxl_b200 import sharding
from oracle import unet_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")   # only to share the NCCL unique id: the data path is the library's own ncclBroadcast
dev = torch.device("cuda", local)
ctx = sdxl_b200.Context(local)
cfg = sdxl_b200.TINY
w = sdxl_b200.synth_weights(cfg, seed=0)                      # every rank can rebuild the weights for the oracle ...
print(f"rank {rank}: gloo up, creating the NCCL communicator", flush=True)
comm = sharding.nccl_comm_init(rank, world, dev)
print(f"rank {rank}: NCCL communicator ready, loading", flush=True)
d = sdxl_b200.Diffuser(ctx, cfg, w if rank == 0 else None, nccl_comm=comm, rank=rank, root=0)   # ... but only rank 0 loads them
print(f"rank {rank}: model loaded through sdxl_unet_load_broadcast", flush=True)
g = torch.Generator().manual_seed(100 + rank)
x = torch.randn(1, 4, 16, 16, generator=g)
c = torch.randn(1, 7, cfg.context_dim, generator=g).half().float()
y = torch.randn(1, cfg.adm_in_channels, generator=g).half().float()
out = d.unet_forward(x, [500], c, y).cpu()
ref = O.unet_forward(cfg, O.to_f32(w), x, torch.tensor([500]), c, y)
err = float((out - ref).norm() / ref.norm())
print(f"rank {rank}/{world}: load_broadcast forward rel err {err:.3e}", flush=True)
errs = [None] * world
dist.all_gather_object(errs, err)
d.close()
sharding.nccl_comm_destroy(comm)
ctx.close()
dist.destroy_process_group()
assert all(e < 2e-3 for e in errs), errs
if rank == 0:
    print("LOAD_BROADCAST_OK", errs)

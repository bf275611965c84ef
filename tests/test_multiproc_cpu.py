"""world_size-2 gloo test (CPU) of the N>1 host logic: prompt sharding, single-message weight broadcast,
max-over-ranks timing, latent gather. The GPU path uses the same functions over NCCL."""
import hashlib
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-xl-burn_b200"))
    from sdxl_b200 import TINY, build_pack, synth_weights
    from sdxl_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    pack = build_pack(synth_weights(TINY, seed=0)) if rank == 0 else None
    pack = sharding.broadcast_pack(pack, 0, dev)
    digest = hashlib.sha256(pack.numpy().tobytes()).hexdigest()
    shard = sharding.shard_indices(8, rank, world)
    tmax = sharding.max_over_ranks(10.0 + rank, dev)
    lat = sharding.gather_latents(torch.full((1, 4, 2, 2), float(rank)))
    q.put((rank, digest, shard, tmax, [float(t.mean()) for t in lat]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_broadcast():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, d0, s0, t0, l0), (r1, d1, s1, t1, l1) = res
    assert d0 == d1                                  # identical weights on every rank after one broadcast
    assert sorted(s0 + s1) == list(range(8)) and not set(s0) & set(s1)   # prompts partition, no overlap
    assert s0 == [0, 2, 4, 6] and s1 == [1, 3, 5, 7]
    assert t0 == t1 == 11.0                          # max over ranks
    assert l0 == l1 == [0.0, 1.0]

"""Multi-GPU plumbing for the sampling path: independent prompts shard across ranks (one process per
GPU), the packed weights are broadcast once from rank 0 (NCCL over NVLink on GPUs; gloo in CPU tests),
and there is NO collective inside the sampling loop. The reference itself is single-GPU
(src/bin/sample/main.rs:131); this is the new build's only distributed logic.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Global batch index i -> rank i mod world (configs 3/4: one image per GPU when n_items == world)."""
    return list(range(rank, n_items, world))


def broadcast_pack(pack: Optional[torch.Tensor], src: int, device: torch.device) -> torch.Tensor:
    """One flat-buffer broadcast of the weight pack (a single message); non-src ranks pass None."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert pack is not None
        return pack
    n = torch.tensor([pack.numel() if pack is not None else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    if pack is None:
        pack = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(pack, src)
    return pack


def max_over_ranks(value: float, device: torch.device) -> float:
    """Timing rule: every multi-GPU number is the max over ranks of the device-measured time."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_latents(latent: torch.Tensor) -> List[torch.Tensor]:
    """Optional: collect each rank's final latent [1,4,h,w] (reporting only; 256 KB per rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [latent]
    out = [torch.empty_like(latent) for _ in range(dist.get_world_size())]
    dist.all_gather(out, latent)
    return out


# ---------------------------------------------------------------------------------------------------
# NCCL communicator for the C ABI's own multi-GPU load (sdxl_unet_load_broadcast)
# ---------------------------------------------------------------------------------------------------
class _NcclUniqueId(__import__("ctypes").Structure):
    _fields_ = [("internal", __import__("ctypes").c_char * 128)]


def nccl_comm_init(rank: int, world: int, device: torch.device):
    """Creates an ncclComm_t (returned as a ctypes void pointer) the way a non-Python host would: ncclGetUniqueId on rank 0,
    the 128-byte id shared through the already-initialised torch.distributed group (any backend), ncclCommInitRank on every
    rank. The handle is what `sdxl_unet_load_broadcast` takes; call nccl_comm_destroy when done."""
    import ctypes as C
    lib = C.CDLL("libnccl.so.2")   # the copy torch already loaded
    uid = _NcclUniqueId()
    if rank == 0:
        rc = lib.ncclGetUniqueId(C.byref(uid))
        if rc != 0:
            raise RuntimeError(f"ncclGetUniqueId failed with {rc}")
    box = [C.string_at(C.byref(uid), 128) if rank == 0 else None]   # all 128 bytes (a c_char array read stops at the first NUL)
    dist.broadcast_object_list(box, src=0)
    C.memmove(C.byref(uid), box[0], 128)
    torch.cuda.set_device(device)
    comm = C.c_void_p()
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _NcclUniqueId, C.c_int]
    rc = lib.ncclCommInitRank(C.byref(comm), world, uid, rank)
    if rc != 0:
        raise RuntimeError(f"ncclCommInitRank failed with {rc}")
    return comm


def nccl_comm_destroy(comm) -> None:
    import ctypes as C
    lib = C.CDLL("libnccl.so.2")
    lib.ncclCommDestroy.argtypes = [C.c_void_p]
    lib.ncclCommDestroy(comm)

"""Host side of the multi-GPU paths.  One process per GPU, launched by torchrun; torch.distributed is the
control plane only (rendezvous, handle exchange, result reduction) — the data path is either nothing at all (independent
views) or the NVLink peer-memory min-merge inside libvkv (vkv_merge, csrc/merge.cu).

  * independent views (BASELINE config 4, the default bench): scene replicated, view i -> rank i mod n, no collective.
  * one huge view (BASELINE config 5): MeshletDraw[] split by contiguous range, ids stay global (visbuffer.h.glsl:15-16:
    drawIndex is the index into the draw list), each rank rasterises into its own full-resolution 64-bit visbuffer and the
    buffers are min-merged before every pyramid build.

The reference is single-GPU; there is no reference interface to mirror here.
"""
import ctypes as C

import numpy as np


def shard_range(n_draws: int, rank: int, world: int):
    """contiguous range [first, first+count) of the draw list owned by `rank`: [r*N/n, (r+1)*N/n)"""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    lo = n_draws * rank // world
    hi = n_draws * (rank + 1) // world
    return lo, hi - lo


def interleaved_owner(draw_ids, world: int, block_log2: int = 11):
    """rank owning each MeshletDraw under vkv_set_shard_interleaved: block b = id >> block_log2 belongs to rank b mod world"""
    return (np.asarray(draw_ids, dtype=np.uint64) >> np.uint64(block_log2)) % np.uint64(world)


def view_shard(n_views: int, rank: int, world: int):
    """round-robin view assignment: view i -> rank i mod world (independent views: cubemap faces, shadow cascades)"""
    return list(range(rank, n_views, world))


def sweep_start(n_views: int, rank: int, world: int) -> int:
    """camera sweeps: every rank walks the whole ring with the same step, starting at rank * n_views / world, so the
    frame-to-frame coherence the previous-frame HiZ test relies on does not depend on the number of GPUs"""
    return rank * n_views // world


def exchange_handles(handle: bytes, dist) -> list:
    """all-gather the 128-byte IPC handles in rank order (any backend: the payload is host bytes)"""
    if len(handle) != 128:
        raise ValueError("IPC handle must be 128 bytes")
    world = dist.get_world_size()
    out = [None] * world
    dist.all_gather_object(out, handle)
    if any(h is None or len(h) != 128 for h in out):
        raise RuntimeError("handle exchange failed")
    return out


def attach_peers(renderer, dist):
    """export this rank's visbuffer, gather everyone's handle, map the peers (vkv_ipc_export / vkv_ipc_attach)"""
    handles = exchange_handles(renderer.ipc_export(), dist)
    renderer.ipc_attach(dist.get_rank(), handles)
    dist.barrier()
    return handles


class _CudaView:
    """expose a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy"""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


def nccl_min_merge(renderer, dist):
    """Cross-check of vkv_merge with the library collective: ncclAllReduce(min) on the keys viewed as int64.
    Every valid key has its top bit set (depth in [0,1] => ~floatBits(depth) >= 0xC07FFFFF), so signed and unsigned
    order agree.  Not the product path — tests and bench --merge nccl only."""
    import torch
    n = renderer.W * renderer.H
    renderer.sync()
    t = torch.as_tensor(_CudaView(renderer.visbuffer64_ptr(), n, "<i8"), device=f"cuda:{torch.cuda.current_device()}")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()


def merge_host(images):
    """what the merge computes, on host arrays (tests): element-wise unsigned min"""
    out = np.array(images[0], dtype=np.uint64, copy=True)
    for im in images[1:]:
        np.minimum(out, np.asarray(im, dtype=np.uint64), out=out)
    return out

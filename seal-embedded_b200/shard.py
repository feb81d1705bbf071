"""Batch sharding across GPUs: one process per GPU, no collective on the data path.

Ciphertexts are independent (the reference encrypts one message per call, seal_embedded.c:98-215), so a
batch of B items splits into contiguous index ranges, rank r of G owning items [B*r/G, B*(r+1)/G)
(SURVEY.md 8e).  Read-only state (keys, twiddle tables) is replicated by constructing the same
``Context`` on every rank.  The only optional communication is the collation of finished
ciphertexts (``all_gather_ciphertexts``), kept off the default path because at the large configs a
full gather moves tens of GB per GPU.

Everything here is index arithmetic and torch.distributed plumbing; it runs unchanged over the
``gloo`` backend on CPU tensors, which is how tests/test_sharding_gloo.py covers the N>1 logic
without GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    first: int  # first item index owned by this rank
    count: int  # number of items owned by this rank

    @property
    def stop(self) -> int:
        return self.first + self.count


def shard_range(batch: int, rank: int, world: int) -> Shard:
    """Contiguous partition: rank r owns [floor(B*r/G), floor(B*(r+1)/G)).  Sizes differ by at most 1,
    every item is owned exactly once, and empty shards are legal (B < G)."""
    if world < 1 or not (0 <= rank < world) or batch < 0:
        raise ValueError(f"bad shard request batch={batch} rank={rank} world={world}")
    first = batch * rank // world
    stop = batch * (rank + 1) // world
    return Shard(rank, world, first, stop - first)


def owner_of(item: int, batch: int, world: int) -> int:
    """Rank that owns item index `item` under shard_range."""
    if not (0 <= item < batch):
        raise ValueError("item out of range")
    # smallest r with floor(B*(r+1)/G) > item
    r = (item * world) // batch
    while shard_range(batch, r, world).stop <= item:
        r += 1
    while shard_range(batch, r, world).first > item:
        r -= 1
    return r


def all_gather_ciphertexts(local, batch: int, group=None):
    """Collate per-rank ciphertext shards [count_r][np][2][n] into the full [batch][np][2][n] tensor on
    every rank (one all_gather; NCCL over NVLink on GPUs, gloo in the CPU tests).

    Shards may differ by one item, so each rank pads to the largest shard for the collective and the
    padding is dropped when the pieces are concatenated in rank order."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = shard_range(batch, rank, world)
    if local.shape[0] != mine.count:
        raise ValueError(f"rank {rank} holds {local.shape[0]} items, shard_range says {mine.count}")
    counts = [shard_range(batch, r, world).count for r in range(world)]
    cap = max(counts)
    if cap == 0:
        return local.new_empty((0,) + tuple(local.shape[1:]))
    if min(counts) == cap:
        # equal shards (the usual case): one collective straight into the result, no padding, no concatenation
        out = local.new_empty((batch,) + tuple(local.shape[1:]))
        try:
            dist.all_gather_into_tensor(out, local.contiguous(), group=group)
            return out
        except (RuntimeError, NotImplementedError):  # a backend without the flat variant
            del out
    padded = local
    if mine.count < cap:
        padded = local.new_zeros((cap,) + tuple(local.shape[1:]))
        padded[: mine.count] = local
    pieces = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(pieces, padded.contiguous(), group=group)
    return torch.cat([pieces[r][: counts[r]] for r in range(world)], dim=0)


def bind_to_gpu_numa(gpu_index: int) -> str:
    """Pin the calling process to the CPUs next to GPU `gpu_index` (NVML's ideal CPU affinity), so that the
    pinned staging buffers it allocates afterwards are first-touched on that GPU's NUMA node and the
    host<->device copies of eight ranks do not all cross one socket.  Returns a description; a no-op (with
    the reason) when NVML or the affinity call is unavailable."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return "no usable CPUs in the GPU's affinity mask"
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} CPUs ({min(cpus)}-{max(cpus)})"
    except Exception as e:  # noqa: BLE001 - best effort
        return f"unavailable: {type(e).__name__}"


def encrypt_asym_sharded(ctx, values, seeds, batch: int, rank: int, world: int, out=None):
    """Encrypt this rank's shard of a global batch held in host numpy arrays `values` [batch][vlen] and
    `seeds` [batch][64] through the host-pointer C ABI.  Returns (Shard, ciphertexts of the shard)."""
    sh = shard_range(batch, rank, world)
    ct = ctx.encrypt_asym_host(values[sh.first:sh.stop], seeds[sh.first:sh.stop], out)
    return sh, ct

"""Multi-GPU: one process per GPU, prompts sharded across ranks, ONE collective per batch.

The path shards by independent units (prompts never interact: SURVEY.md §8e), so there is no data-path
collective inside the LRP pass; the only exchange is the gather of the final `[B_local, S]` fp32 token-relevance
vectors (<= 256 KiB in total at the headline configuration: latency-bound, nothing to fuse with a GEMM).  It is
issued on the compute stream right after the Gradient x Input reduction kernel.
Backends: NCCL over NVLink 5 / NVSwitch on GPUs; gloo for the CPU tests of the host logic.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun). -> (rank, world, local)"""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous, balanced [lo, hi) slice of `n_items` prompts owned by `rank` (first n%world ranks get one more)"""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_relevance(rel_local: torch.Tensor, n_total: int, world: int) -> torch.Tensor:
    """The single collective of the path: all ranks contribute `[B_local, S]` fp32, every rank receives
    `[n_total, S]` in prompt order.  Ragged shards are padded to the largest shard for the collective."""
    if world == 1 or not dist.is_initialized():
        return rel_local
    S = rel_local.shape[1]
    b_max = -(-n_total // world)
    buf = rel_local
    if rel_local.shape[0] != b_max:
        buf = torch.zeros((b_max, S), dtype=rel_local.dtype, device=rel_local.device)
        buf[: rel_local.shape[0]] = rel_local
    out = torch.empty((world * b_max, S), dtype=rel_local.dtype, device=rel_local.device)
    dist.all_gather_into_tensor(out, buf.contiguous())
    if n_total == world * b_max:
        return out
    pieces = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        pieces.append(out[r * b_max: r * b_max + (hi - lo)])
    return torch.cat(pieces, dim=0)


def attribute_sharded(engine, input_ids: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Shard `[N,S]` prompt ids over the ranks, run the engine on the local shard, gather `[N,S]` relevance."""
    N = input_ids.shape[0]
    lo, hi = shard_range(N, rank, world)
    local = input_ids[lo:hi].to(engine.device, non_blocking=True)
    rels = [engine.attribute_device(local[i:i + engine.micro_batch]) for i in range(0, hi - lo, engine.micro_batch)]
    rel_local = torch.cat(rels, dim=0) if len(rels) != 1 else rels[0]
    return gather_relevance(rel_local, N, world)

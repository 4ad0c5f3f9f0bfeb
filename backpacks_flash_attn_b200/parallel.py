"""Batch-sharded (data-parallel) inference over the GPUs of one node.

Every operator on the Backpack forward path is independent across the batch dimension (attention and the
sense-mix only mix along the sequence), so the path shards into independent units with NO data-path
collective: weights are replicated, `input_ids` is split contiguously, one process per GPU.
NCCL (over NVLink / NVSwitch) is used only for control-plane collectives: proving the replicas hold the
same weights, the barrier + max-over-ranks timing of the benchmark, and optionally returning the
last-position argmax ids.  The same code runs on CPU with the gloo backend (tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .utils.weights import parameter_checksum


def init_distributed(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment.  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_range(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [start, end) slice of the batch owned by `rank`; the first (global_batch % world) ranks
    get one extra sequence."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(input_ids: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    s, e = shard_range(input_ids.shape[0], rank, world)
    return input_ids[s:e]


def assert_replicas_match(model: torch.nn.Module) -> float:
    """all_reduce(MIN) and all_reduce(MAX) of a parameter checksum must agree on every rank."""
    cs = parameter_checksum(model)
    if dist.is_initialized() and dist.get_world_size() > 1:
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if lo.item() != hi.item():
            raise RuntimeError(f"data-parallel replicas diverge: checksum min {lo.item()} max {hi.item()}")
    return cs.item()


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def sum_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.item()


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def gather_next_tokens(logits_last: torch.Tensor) -> torch.Tensor:
    """All-gather of the greedy next-token ids (local_batch,) -> (global_batch,), equal shard sizes.
    The full logits (global_batch, seqlen, vocab) are never moved."""
    ids = logits_last.argmax(dim=-1).to(torch.int64).contiguous()
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return ids
    out = torch.empty(ids.numel() * dist.get_world_size(), dtype=ids.dtype, device=ids.device)
    dist.all_gather_into_tensor(out, ids)
    return out

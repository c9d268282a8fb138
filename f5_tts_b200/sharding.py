"""Utterance sharding across the GPUs of one box.

The ODE-sampling path is embarrassingly parallel over utterances: the reference splits prompts between processes with
Accelerate ``split_between_processes`` (eval/eval_infer_batch.py:181) or a ``DistributedSampler``
(runtime/triton_trtllm/benchmark.py:340-341) and never communicates inside the NFE loop.  Here: one process per GPU,
contiguous shards of the duration-sorted batch (limits padding, like eval/utils_eval.py:155-205), weights replicated,
and ONE all-gather of the finished (padded) mel + audio at the end.
"""
from __future__ import annotations

import torch


def shard_bounds(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) — the first n_items % world ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def plan_shards(durations: list[int], world: int) -> list[list[int]]:
    """Indices per rank: sort by duration (descending) so every shard pads to a similar length."""
    order = sorted(range(len(durations)), key=lambda i: -durations[i])
    return [order[slice(*shard_bounds(len(order), r, world))] for r in range(world)]


def gather_padded(local: torch.Tensor, lengths: torch.Tensor, group=None):
    """All-gather variable-length rows.  local [b, L, ...] (rank-local max length L), lengths int64 [b].
    Returns (list over ranks of [b_r, L_r, ...] tensors, list of length tensors).  Works on NCCL and gloo."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    meta = torch.tensor([local.shape[0], local.shape[1]], dtype=torch.int64, device=local.device)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    bmax, lmax = max(int(m[0]) for m in metas), max(int(m[1]) for m in metas)
    pad = torch.zeros((bmax, lmax) + tuple(local.shape[2:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0], : local.shape[1]] = local
    lpad = torch.zeros(bmax, dtype=torch.int64, device=local.device)
    lpad[: lengths.shape[0]] = lengths
    # flat [world * bmax, ...] outputs: the layout both NCCL and gloo accept for all_gather_into_tensor
    out = torch.empty((world * bmax,) + tuple(pad.shape[1:]), dtype=pad.dtype, device=pad.device)
    lout = torch.empty(world * bmax, dtype=torch.int64, device=pad.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    dist.all_gather_into_tensor(lout, lpad, group=group)
    out = out.view((world, bmax) + tuple(pad.shape[1:]))
    lout = lout.view(world, bmax)
    res, lens = [], []
    for r in range(world):
        b, L = int(metas[r][0]), int(metas[r][1])
        res.append(out[r, :b, :L])
        lens.append(lout[r, :b])
    return res, lens

"""One process per GPU, batch-sharded sampling, ONE all-gather of the results at the end.

The reference's multi-GPU path (test_flow_latent_ddp.py:22-34, 116-143; ddp_utils.py:17-30) runs independent
per-rank batches with replicated weights and exchanges nothing but barriers (results go to JPEG files on a
shared disk).  The samples are independent ODEs, so the path shards over the batch axis with no data-path
collective; BASELINE.json's north_star adds a single all-gather of the final latents, done here with
torch.distributed (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """torchrun-style bootstrap (replaces ddp_utils.init_processes).  Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def rank_seed(seed: int, rank: int) -> int:
    """test_flow_latent_ddp.py:30 / test_flow_latent.py:106: seed = args.seed + rank."""
    return seed + rank


def total_samples(n_sample: int, batch_size: int, world_size: int) -> int:
    """test_flow_latent_ddp.py:116-123: round n_sample up to a multiple of the global batch."""
    g = batch_size * world_size
    return int(math.ceil(n_sample / g) * g)


def file_index(j: int, world_size: int, rank: int, total: int) -> int:
    """test_flow_latent_ddp.py:138: global index of the j-th image of this rank's current batch."""
    return j * world_size + rank + total


def all_gather_batch(x: torch.Tensor) -> torch.Tensor:
    """Gather per-rank result batches [B_r, ...] into [world * B_r, ...] ordered by file_index, i.e. rank
    interleaved (image j of rank r lands at j * world + r), on every rank.  One collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    world = dist.get_world_size()
    x = x.contiguous()
    flat = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(flat, x)          # rank-major: flat[r * B + j]
    out = flat.reshape((world, x.shape[0]) + tuple(x.shape[1:]))
    # out[r, j] -> position j * world + r
    return out.transpose(0, 1).reshape((world * x.shape[0],) + tuple(x.shape[1:]))

"""Multi-GPU plumbing for the spectral path: utterance sharding and the metric all-reduce.

The path itself is embarrassingly parallel over utterances (every STFT / mask / iSTFT
touches one mixture only; the reference's own loops are per file, process.py:89-110), so
ranks own contiguous blocks of mixtures and NO collective sits on the data path.  The only
exchange is one all-reduce (sum) per batch of a 4-float vector
``[sum max-cross-SNR, sum ae-loss * count, sum SDR, count]`` that reproduces the batch means
of main.py:353-361 / :446-457 across ranks - NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block ``[lo, hi)`` of ``n_items`` utterances owned by ``rank``; the first
    ``n_items % world`` ranks get one extra item, empty blocks are allowed."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError(f"shard_range: bad (n_items={n_items}, rank={rank}, world={world})")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def metric_vector(snr_sum=0.0, ae_sum=0.0, sdr_sum=0.0, count=0.0, device=None):
    return torch.tensor([snr_sum, ae_sum, sdr_sum, count], dtype=torch.float32, device=device)


def allreduce_metrics(vec: torch.Tensor):
    """Sum the per-rank ``[sum_snr, sum_ae, sum_sdr, count]`` vector over all ranks (no-op
    without an initialised process group) and return the global means
    ``(snr, ae_loss, sdr, count)``."""
    if vec.numel() != 4:
        raise ValueError("allreduce_metrics: expected a 4-vector [sum_snr, sum_ae, sum_sdr, count]")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    c = float(vec[3])
    if c <= 0:
        return 0.0, 0.0, 0.0, 0.0
    return float(vec[0]) / c, float(vec[1]) / c, float(vec[2]) / c, c

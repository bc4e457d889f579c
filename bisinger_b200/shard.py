"""Batch sharding of independent utterances over replica processes (one process per GPU).

The hot path has no cross-utterance coupling (no BatchNorm, no cross-batch op in DiffNet, the sampler or HiFi-GAN), so
multi-GPU inference is N independent replicas: phrases are grouped into batches, batches go round-robin to ranks, and the
only exchange is a final reduction of (audio seconds, elapsed time) for reporting -- no collective on the data path.
The reference has no multi-GPU inference at all (inference scripts use one device, inference/m4singer/base_svs_infer.py:20-23);
its DDP training splits batches the same way (x[rank::world], tasks/tts/tts.py:84-87).
"""
from __future__ import annotations

from typing import List, Tuple


def make_batches(n_items: int, batch_size: int) -> List[Tuple[int, int]]:
    """[(start, stop)) item ranges of consecutive batches; the last one may be ragged."""
    if n_items < 0 or batch_size <= 0:
        raise ValueError("n_items must be >= 0 and batch_size > 0")
    return [(s, min(s + batch_size, n_items)) for s in range(0, n_items, batch_size)]


def shard_batches(n_items: int, batch_size: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """Batches owned by `rank`: round-robin over the batch index."""
    if not 0 <= rank < world:
        raise ValueError("rank must be in [0, world)")
    return make_batches(n_items, batch_size)[rank::world]


def reduce_throughput(audio_seconds: float, elapsed_seconds: float, group=None):
    """(total audio seconds over all ranks, max elapsed over ranks) via torch.distributed when initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return audio_seconds, elapsed_seconds
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    a = torch.tensor([audio_seconds], dtype=torch.float64, device=dev)
    t = torch.tensor([elapsed_seconds], dtype=torch.float64, device=dev)
    dist.all_reduce(a, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(a.item()), float(t.item())

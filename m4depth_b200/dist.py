"""Multi-GPU plumbing for the path: sequences shard by batch, one process per GPU, no data-path
collective; the single collective is an all-gather of each rank's 14 metric partials at the end of a run.
(The reference has no distributed code in its current tree; its legacy TF1 tower pipeline is out of scope.)
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (NCCL on GPU boxes, gloo for CPU tests)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_sequences(n_sequences, rank, world):
    """Contiguous block of sequence indices owned by ``rank`` (sizes differ by at most one; ragged tails allowed)."""
    base, rem = divmod(n_sequences, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def all_gather_partials(partials):
    """partials: float64 tensor [14] (MetricsAccumulator.partials()) -> [world,14] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return partials.reshape(1, -1)
    out = [torch.empty_like(partials) for _ in range(dist.get_world_size())]
    dist.all_gather(out, partials.contiguous())
    return torch.stack(out)


def max_over_ranks(value, device):
    """Max of a python float over ranks (timing protocol: the slowest rank defines the step time)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

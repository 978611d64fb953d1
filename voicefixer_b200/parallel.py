"""Multi-GPU plumbing: one process per GPU, utterances sharded by batch.

Exactly two collectives, as the north star prescribes: one broadcast of the packed weight arena
from rank 0 at start-up and a gather of the output waveforms to rank 0.  The functions take the
process group's backend as given, so the same code runs under NCCL on the GPUs and under gloo on
CPU tensors in the tests (world_size 2)."""
import os
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialises torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device(f"cuda:{local}")
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank`; the first n_items % world ranks get one extra."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_arena(table, arena, device, src=0):
    """Rank `src` holds (table, arena uint8 tensor); everyone returns (table, arena on `device`).
    One broadcast of the layout (python object) and one of the bytes (NCCL over NVLink on GPUs)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return table, arena
    meta = [table, int(arena.numel()) if arena is not None else 0]
    dist.broadcast_object_list(meta, src=src)
    table, nbytes = meta
    if dist.get_rank() != src:
        arena = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dist.broadcast(arena, src=src)
    return table, arena


def gather_waveforms(local_out, dst=0):
    """local_out (b_r, L) on each rank -> (sum b_r, L) on rank dst (None elsewhere).  Equal shards
    use one dist.gather; ragged shards are padded to the largest shard first."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_out
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [torch.zeros(1, dtype=torch.int64, device=local_out.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local_out.shape[0]], dtype=torch.int64, device=local_out.device))
    sizes = [int(s.item()) for s in sizes]
    bmax = max(sizes)
    send = local_out
    if local_out.shape[0] != bmax:
        send = torch.zeros(bmax, local_out.shape[1], dtype=local_out.dtype, device=local_out.device)
        send[: local_out.shape[0]] = local_out
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send.contiguous(), bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)

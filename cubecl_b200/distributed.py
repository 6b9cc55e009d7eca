"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only for rendezvous / id exchange / barriers.

The data path uses the C ABI's own NCCL communicator (b200_comm_init / b200_all_reduce), keyed by the sorted device set
like the reference's CommunicationId (crates/cubecl-runtime/src/server/base.rs:605-620).  The reference shares the
ncclUniqueId through a process-global map because it is single-process (cubecl-cuda/src/compute/communication.rs:11-25);
with one process per GPU the id travels through the torch.distributed store instead.

Sharding rules:
  * batched matmul: contiguous B/world batches per rank, no collective on the data path
  * reduce-sum:     contiguous outer-axis slabs per rank, local reduce, one all-reduce of the (tiny) partial result
"""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass
class Env:
    rank: int
    local_rank: int
    world_size: int


def env() -> Env:
    return Env(int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [start, stop) of `total` units for `rank`; the first `total % world` ranks get one extra unit."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def init_process_group(backend: str | None = None):
    """Rendezvous over 127.0.0.1 (the container hostname may not resolve). Returns torch.distributed."""
    import torch.distributed as dist

    if dist.is_initialized():
        return dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    e = env()
    if backend is None:
        import torch
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kwargs = {}
    if backend == "nccl":
        import torch
        torch.cuda.set_device(e.local_rank)
        kwargs["device_id"] = torch.device("cuda", e.local_rank)
    dist.init_process_group(backend=backend, rank=e.rank, world_size=e.world_size, **kwargs)
    return dist


def exchange_unique_id(make_id, dist=None) -> bytes:
    """Rank 0 creates the 128-byte NCCL id with `make_id()`; everyone receives it (object broadcast = CPU store path)."""
    if dist is None:
        import torch.distributed as dist  # type: ignore[no-redef]
    payload = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    uid = payload[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("unique id exchange failed")
    return bytes(uid)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Timing rule: multi-GPU numbers are the max over ranks."""
    import torch
    if dist is None:
        import torch.distributed as dist  # type: ignore[no-redef]
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def connect_p2p(client, dist=None) -> list[int]:
    """Exchange every rank's mailbox export (all_gather_object) and connect; returns the device set."""
    if dist is None:
        import torch.distributed as dist  # type: ignore[no-redef]
    mine = client.p2p_export()
    everyone = [None] * dist.get_world_size()
    dist.all_gather_object(everyone, mine)
    client.p2p_connect(everyone)
    return sorted(e[0] for e in everyone)

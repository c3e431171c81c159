"""Multi-GPU plumbing: one process per GPU, environments block-sharded with no data-path collective; the only
exchange is the all-reduce(sum) of the PPO gradient buffer (6 149 fp32 + 3 scalars) per minibatch, NCCL over
NVLink through torch.distributed (gloo on CPU for the tests).  SURVEY.md section 8e."""
from __future__ import annotations

import os


def world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n_total: int, rank: int, world_size: int):
    """Contiguous block of environments / samples owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n_total, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class DeviceBuffer:
    """Zero-copy view of a raw device pointer for torch (`torch.as_tensor(DeviceBuffer(...), device='cuda')`)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def grad_tensor(agent):
    """The policy handle's gradient buffer as a torch CUDA tensor (no copy)."""
    import torch
    ptr, n = agent.grad_buffer()
    return torch.as_tensor(DeviceBuffer(ptr, n), device=f"cuda:{torch.cuda.current_device()}")


def norm_stats_tensor(agent):
    """The partial-sum buffer of PPOAgent.Normalize (double[2 * 128]) as a torch CUDA tensor (no copy)."""
    import ctypes as C

    import torch

    from ._lib import check, lib
    p, n = C.c_void_p(), C.c_int32(0)
    check(lib().wb_normalize_stats_buffer(agent._h, C.byref(p), C.byref(n)))
    return torch.as_tensor(DeviceBuffer(p.value, n.value, "<f8"), device=f"cuda:{torch.cuda.current_device()}")


def allreduce_sum_(tensor):
    """In-place sum over ranks; a no-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def connect_peers(agent) -> bool:
    """Wire `agent` (one per rank) for the fused gradient reduction + all-reduce over NVLink peer memory: every rank exports its
    exchange buffer as a CUDA IPC handle, the 64-byte handles are all-gathered through the existing process group, every rank
    maps its peers' buffers.  Returns False (NCCL path stays in use) without an initialised multi-rank group."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from ._lib import check, lib
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2 or dist.get_world_size() > 8:
        return False
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_ubyte * 64)()
    if lib().wb_comm_local_handle(agent._h, buf) != 0:
        return False  # not the default networks: the fused exchange does not apply, the NCCL all-reduce does
    cuda = dist.get_backend() == "nccl"
    mine = torch.tensor(list(buf), dtype=torch.uint8, device="cuda" if cuda else "cpu")
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    handles = bytes(torch.cat(gathered).cpu().numpy().tobytes())
    check(lib().wb_comm_connect(agent._h, rank, world, handles))
    dist.barrier()  # nobody pushes before everybody has mapped everybody
    return True


def train_minibatch_sharded(agent, states, actions, logp, advantages, returns, grad_view=None):
    """One data-parallel PPO minibatch: local gradient over this rank's shard (already divided by the GLOBAL
    batch size hp.batch_size), all-reduce(sum), identical Adam on every rank (weights stay bit-identical)."""
    out = agent.Gradients(states, actions, logp, advantages, returns)
    if grad_view is None:
        grad_view = grad_tensor(agent)
    allreduce_sum_(grad_view)
    agent.Optimise()
    return out

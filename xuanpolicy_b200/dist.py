"""Env-sharded data parallelism: one process per GPU; exchanges over NVLink peer memory (PeerComm, csrc/peer_comm.cu),
NCCL all-reduces as the fallback.

The reference is single-process; its only collective is the (disabled) mpi4py Allreduce of [sums..., count]
inside RunningMeanStd (xuance/common/statistic_tools.py:6-32).  The PPO path shards by environment:

  * rank r of W owns envs [r*N/W, (r+1)*N/W): its own env state / RNG, rollout-buffer shard, local index
    permutation and local minibatch of B/W samples; the policy and the Adam state are replicated;
  * per update there are exactly two exchanges, both latency-bound:
      1. all-reduce(sum) of (sum adv, sum adv^2)  -> every rank normalises with the GLOBAL-minibatch mean/std,
         the sharded equivalent of memory_tools.py:241-242;
      2. all-reduce(sum) of the flat gradient, each rank having scaled its loss by 1/(B_local*W), so that the
         sum is the gradient of the global-minibatch mean; clip + Adam then run identically on every rank.
  Nothing is exchanged during the rollout or the GAE scan.

The default exchange path is `PeerComm` below (kernels of ours over CUDA-IPC-mapped peer memory, inside the epoch
graph).  `allreduce_adv_stats` / `allreduce_flat_grad` / `broadcast_parameters` are the torch.distributed form of the
same exchanges: the agent and learner call them on the NCCL fallback path (XB_PEER_COMM=0 or no IPC) and the CPU
tests drive them over gloo.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist


def init_from_env(backend="nccl"):
    """Initialise torch.distributed from torchrun's RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* variables."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 0, 1
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return rank, local, world


def shard_envs(total_envs, world, rank):
    """[lo, hi) of the envs owned by `rank`; shards must be equal for mean-of-means == global mean."""
    if total_envs % world != 0:
        raise ValueError("env count %d is not divisible by world size %d" % (total_envs, world))
    per = total_envs // world
    return rank * per, (rank + 1) * per


def allreduce_adv_stats(stats, group=None):
    """stats = fp64 [2] (sum, sumsq) of the local minibatch's advantages -> global sums, in place."""
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def mean_std_from_stats(stats, count):
    """(mean, population std) exactly as the kernels derive them (csrc/ppo_loss.cu load_adv_norm)."""
    mean = stats[0] / count
    var = torch.clamp(stats[1] / count - mean * mean, min=0.0)
    return mean, torch.sqrt(var)


def allreduce_flat_grad(flat_grad, group=None):
    """Sum of per-rank gradients of loss/(B_local*W) == gradient of the global-minibatch mean loss."""
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def broadcast_parameters(flat_param, src=0, group=None):
    """Ranks are seeded identically, so this is a safety net rather than a requirement."""
    dist.broadcast(flat_param, src=src, group=group)
    return flat_param


class _DeviceMemory:
    """Zero-copy torch view of a raw device allocation (``__cuda_array_interface__``)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerComm:
    """One rank's side of the NVLink peer-memory exchange (csrc/peer_comm.cu, include/xb200.h section 5b).

    Allocates this rank's comm block [barrier flags | minibatch statistics | gradient inbox [2][8][n]], exchanges CUDA
    IPC handles through the process group and maps every peer's block.  `stats` (fp64) is a torch view INTO the block;
    the gradient inbox is only touched by xb_peer_allreduce_grad_norm (peers push their slices into it with P2P
    stores).  Single node only (IPC), world size <= 8."""

    def __init__(self, n_grad_floats, device, group=None):
        from . import _lib
        lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("peer-memory exchange supports up to 8 ranks of one node")
        self.n = (int(n_grad_floats) + 3) // 4 * 4
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            ptr = C.c_void_p()
            _lib.call("xb_peer_alloc", C.byref(ptr), lib.xb_peer_block_bytes(self.n))
            self._own = ptr.value
            handle = C.create_string_buffer(64)
            _lib.call("xb_peer_export", ptr, handle)
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self._imported = []
            self.bases = (C.c_void_p * self.world)()
            for r in range(self.world):
                if r == self.rank:
                    self.bases[r] = self._own
                else:
                    p = C.c_void_p()
                    _lib.call("xb_peer_import", C.create_string_buffer(handles[r], 64), C.byref(p))
                    self._imported.append(p.value)
                    self.bases[r] = p.value
            self._keep = _DeviceMemory(self._own + lib.xb_peer_stats_offset(), lib.xb_peer_stats_max(), "<f8")
            self.stats = torch.as_tensor(self._keep, device=self.device)
            self.tickets = torch.zeros(64, dtype=torch.int32, device=self.device)
            torch.cuda.synchronize(self.device)
        dist.barrier(group)          # every rank has mapped every block before the first kernel touches one

    def check(self):
        """Raises if a cross-GPU barrier timed out on this rank (a peer process died or fell > 60 s behind).
        Costs one 4-byte D2H read: call it where the host synchronises anyway."""
        if int(self.tickets[62].item()) != 0:
            raise RuntimeError("xb200 peer-memory barrier timed out on rank %d: a peer rank is gone or stalled" % self.rank)

    def close(self):
        from . import _lib
        if self._own is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
        for p in self._imported:
            _lib.call("xb_peer_close", C.c_void_p(p))
        _lib.call("xb_peer_free", C.c_void_p(self._own))
        self._own, self._imported = None, []

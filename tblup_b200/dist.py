"""One process per GPU: shard a generation over ``torch.distributed`` ranks and all-gather the fitness.

The path shards with no data-path collective (every (individual, fold) evaluation is independent; genotypes are
replicated), so the only exchange is the P x n_slots float64 fitness vector -- ``all_gather_into_tensor`` over
NCCL on the GPUs (gloo on CPU in the tests).  Takes the place of the result queue drain at
tblup/evaluator.py:396-398."""
import numpy as np
import torch
import torch.distributed as dist

from .evaluator import shard_bounds


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def evaluate_sharded(eval_fn, flat, off, n_slots, device=None):
    """``eval_fn(flat_shard, off_shard) -> array (p_local, n_slots)`` runs on this rank's slice of the batch;
    returns the full (P, n_slots) array on every rank.  Shards are contiguous and balanced by genome length."""
    rank, world = rank_world()
    P = off.size - 1
    cuts = shard_bounds(np.diff(off), world)
    lo, hi = cuts[rank], cuts[rank + 1]
    local = np.empty((0, n_slots))
    if hi > lo:
        local = np.asarray(eval_fn(flat[off[lo]:off[hi]], off[lo:hi + 1] - off[lo]), dtype=np.float64)
        local = local.reshape(hi - lo, n_slots)
    if world == 1:
        return local
    widest = max(cuts[r + 1] - cuts[r] for r in range(world))
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    send = torch.full((widest * n_slots,), float("nan"), dtype=torch.float64, device=dev)
    send[:(hi - lo) * n_slots] = torch.from_numpy(local.ravel()).to(dev)
    recv = torch.empty(world * widest * n_slots, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy().reshape(world, widest, n_slots)
    out = np.empty((P, n_slots))
    for r in range(world):
        out[cuts[r]:cuts[r + 1]] = recv[r, :cuts[r + 1] - cuts[r]]
    return out

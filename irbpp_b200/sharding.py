"""Multi-GPU layout of the packing environments: bins are independent, so they are partitioned
contiguously by env index, one process per GPU, with NO collective inside a step.  The only exchange
is the end-of-rollout gather of observations / masks to every learner rank; the
reference has no counterpart (single host, pipes: wrapper/shmem_vec_env.py:47-57).

Works with any ``torch.distributed`` backend (``nccl`` on GPUs, ``gloo`` in the CPU tests)."""
import numpy as np


def shard_range(num_envs_total, rank, world_size):
    """Contiguous block [lo, hi) of env indices owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(num_envs_total, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sequences(sequences, rank, world_size):
    lo, hi = shard_range(len(sequences), rank, world_size)
    return np.ascontiguousarray(sequences[lo:hi])


def candidate_mask(obs, selected_action):
    """The action mask is column 4 of the candidate rows (reference tools.py:298-299)."""
    n = obs.shape[0]
    return obs[:, :selected_action * 5].reshape(n, selected_action, 5)[:, :, 4]


def gather_rollout(local, world_size=None, group=None):
    """All-gather a per-rank tensor ``[n_local, ...]`` (equal n_local on every rank) into
    ``[world * n_local, ...]`` ordered by rank, i.e. by global env index."""
    import torch
    import torch.distributed as dist
    if world_size is None:
        world_size = dist.get_world_size(group)
    if world_size == 1:
        return local
    local = local.contiguous()
    out = torch.empty((world_size * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out


def scatter_actions(global_actions, rank, world_size):
    """Slice of a global action vector (ordered by global env index) that belongs to ``rank``."""
    lo, hi = shard_range(global_actions.shape[0], rank, world_size)
    return global_actions[lo:hi]

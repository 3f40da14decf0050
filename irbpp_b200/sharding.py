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


class AsyncRolloutGather(object):
    """The rollout-end gather taken off the critical path ("overlap with the next step's
    scan"): ``start(local)`` enqueues the all-gather of a finished rollout's tensor on a side stream
    (after everything already enqueued on the caller's current stream), the caller goes on stepping the
    next rollout, ``finish()`` makes the current stream wait for the gather and returns the gathered
    tensor.  Two output buffers alternate, so a gathered tensor stays valid while the next gather is in
    flight.  On CPU tensors (gloo, the tests) the collective simply runs asynchronously."""

    def __init__(self, world_size=None, group=None):
        import torch.distributed as dist
        self.world = dist.get_world_size(group) if world_size is None else world_size
        self.group = group
        self._bufs = [None, None]
        self._turn = 0
        self._work = None
        self._out = None
        self._side = None
        self.events = None                      # (start, end) CUDA events of the last gather on the side stream

    def start(self, local):
        import torch
        import torch.distributed as dist
        if self._work is not None:
            raise RuntimeError("a gather is already in flight")
        local = local.contiguous()
        if self.world == 1:
            self._out, self._work = local, False
            return
        shape = (self.world * local.shape[0],) + tuple(local.shape[1:])
        buf = self._bufs[self._turn]
        if buf is None or buf.shape != shape or buf.dtype != local.dtype or buf.device != local.device:
            buf = torch.empty(shape, dtype=local.dtype, device=local.device)
            self._bufs[self._turn] = buf
        self._turn ^= 1
        if local.is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(device=local.device)
            cur = torch.cuda.current_stream(local.device)
            self._side.wait_stream(cur)                       # `local` is complete in stream order
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(self._side):
                e0.record()
                dist.all_gather_into_tensor(buf, local, group=self.group)
                e1.record()
            local.record_stream(self._side)
            self.events = (e0, e1)
            self._work = True
        else:
            self._work = dist.all_gather_into_tensor(buf, local, group=self.group, async_op=True)
        self._out = buf

    def finish(self):
        import torch
        if self._work is None:
            raise RuntimeError("no gather in flight")
        if self._work is True:
            torch.cuda.current_stream(self._out.device).wait_stream(self._side)
        elif self._work is not False:
            self._work.wait()
        out, self._out, self._work = self._out, None, None
        return out


def scatter_actions(global_actions, rank, world_size):
    """Slice of a global action vector (ordered by global env index) that belongs to ``rank``."""
    lo, hi = shard_range(global_actions.shape[0], rank, world_size)
    return global_actions[lo:hi]

"""Two-GPU test (``pytest -m gpu`` on a box with >= 2 devices; skipped otherwise): SURVEY.md 8(e).

Bins sharded by index over two ranks (one process per GPU, NCCL) must produce exactly the observations of ONE
unsharded environment over all bins, and every rollout-gather implementation of ``irbpp_b200.sharding`` must
deliver exactly the concatenation of the shards.  The world-size-2 host logic is covered on CPU (gloo) in
``tests/test_sharding.py``; this one runs the CUDA kernels and the CUDA gathers."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_TOTAL, STEPS, SEL = 256, 6, 500


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _choose(valid_np, first_bin, step):
    """Deterministic and independent of the sharding: a feasible candidate row picked from the GLOBAL bin index."""
    n = valid_np.shape[0]
    acts = np.zeros(n, dtype=np.int64)
    for i in range(n):
        idx = np.nonzero(valid_np[i])[0]
        acts[i] = idx[(3 * (first_bin + i) + step) % len(idx)] if len(idx) else 0
    return acts


def _rollout(env, torch, first_bin):
    """STEPS steps with the deterministic policy; returns the observation after every step (device tensors)."""
    obs = env.reset()
    out = [obs.clone()]
    for t in range(STEPS):
        valid = (obs[:, :SEL * 5].view(obs.shape[0], SEL, 5)[:, :, 4] == 1).cpu().numpy()
        obs, rew, done, infos = env.step(_choose(valid, first_bin, t))
        out.append(obs.clone())
    return out


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from irbpp_b200 import shapes, sharding
    from irbpp_b200.vec_env import GpuVecEnv
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok, note = True, ""
    try:
        lib = shapes.make_irregular_library(12, seed=7, num_rotations=8)
        seqs = shapes.make_sequences(N_TOTAL, 64, lib.num_shapes, seed=5)
        lo, hi = sharding.shard_range(N_TOTAL, rank, world)
        env = GpuVecEnv(lib, sharding.shard_sequences(seqs, rank, world), device=str(dev))
        mine = _rollout(env, torch, lo)
        gathers = [("nccl", sharding.AsyncRolloutGather(world)), ("compact", sharding.CompactRolloutGather(SEL, world)),
                   ("sendrecv", sharding.AsyncRolloutGather(world, point_to_point=True)), ("symm", sharding.SymmMemGather(world))]
        full = None
        if rank == 0:                                       # the unsharded run of all bins on this rank's GPU
            whole = GpuVecEnv(lib, seqs, device=str(dev))
            full = _rollout(whole, torch, 0)
            whole.close()
        for name, g in gathers:
            for t in (0, STEPS):                            # two gathers per implementation: buffers alternate
                g.start(mine[t])
                got = g.finish()
                torch.cuda.synchronize(dev)
                if got.shape != (N_TOTAL, mine[t].shape[1]) or not torch.equal(got[lo:hi], mine[t]):
                    ok, note = False, "%s: own shard differs (step %d)" % (name, t)
                if rank == 0 and not torch.equal(got, full[t]):
                    ok, note = False, "%s: gathered shards differ from the unsharded environment (step %d)" % (name, t)
            note += " %s=%s;" % (name, getattr(g, "kind", "nccl all-gather")[:40])
        env.close()
    except Exception as exc:                                 # reported through the queue, the parent asserts
        ok, note = False, "%s: %r" % (type(exc).__name__, exc)
    q.put((rank, ok, note))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_bins_and_cuda_gathers_equal_the_unsharded_environment():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted(q.get(timeout=240) for _ in range(world))
        for p in procs:
            p.join(timeout=60)
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
    assert all(ok for _, ok, _ in res), res

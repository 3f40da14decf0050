"""Host-side multi-GPU logic on CPU: index sharding and the rollout-end gather with world_size = 2 over
gloo (the N > 1 path of bench.py uses the same helpers over NCCL)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from irbpp_b200 import sharding


def test_shard_range_partitions_exactly():
    for n in (1, 7, 4096, 32768, 10):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_candidate_mask_is_column_four():
    obs = np.arange(2 * 3533, dtype=np.float32).reshape(2, 3533)
    m = sharding.candidate_mask(obs, 500)
    assert m.shape == (2, 500) and m[1, 3] == obs[1, 3 * 5 + 4]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_total, width, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seqs = np.arange(n_total * 4, dtype=np.int32).reshape(n_total, 4)
    mine = sharding.shard_sequences(seqs, rank, world)
    lo, hi = sharding.shard_range(n_total, rank, world)
    assert np.array_equal(mine, seqs[lo:hi])
    # each rank's "observations": row i holds its global env index
    local = torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1).repeat(1, width)
    full = sharding.gather_rollout(local, world)
    ok = full.shape == (n_total, width) and bool((full[:, 0] == torch.arange(n_total, dtype=torch.float32)).all())
    acts = torch.arange(n_total, dtype=torch.int64) * 3
    mine_acts = sharding.scatter_actions(acts, rank, world)
    ok = ok and bool((mine_acts == acts[lo:hi]).all())
    # the pipelined form bench.py uses: start the gather, keep "stepping", finish; two rollouts back to back
    g = sharding.AsyncRolloutGather(world)
    for rollout in range(3):
        g.start(local + rollout)
        busy = torch.ones(4).sum()                     # work that overlaps the collective
        got = g.finish()
        ok = ok and got.shape == (n_total, width) and bool((got[:, 0] == torch.arange(n_total, dtype=torch.float32) + rollout).all())
    g2 = sharding.AsyncRolloutGather(world, point_to_point=True)      # the same gather as send/recv pairs
    for rollout in range(2):
        g2.start(local + 7 * rollout)
        got = g2.finish()
        ok = ok and got.shape == (n_total, width) and bool((got[:, 0] == torch.arange(n_total, dtype=torch.float32) + 7 * rollout).all())
    try:
        g.finish(); ok = False                         # nothing in flight
    except RuntimeError:
        pass
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the bench's max-over-ranks timing reduction
    ok = ok and float(t.item()) == float(world)
    out_q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_rollout_world2_gloo():
    world, n_total, width = 2, 16, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, width, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]

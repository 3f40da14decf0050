#!/usr/bin/env python
"""Diagnostics for the copy-engine gather: peer access, bandwidth of a cross-device copy inside one process, and of a
push into another process's buffer opened through CUDA IPC.  Run under torchrun with 2 ranks (one GPU each)."""
import os, sys, time
import torch
import torch.distributed as dist
from torch.multiprocessing.reductions import rebuild_cuda_tensor, reduce_tensor

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = 58 * 1024 * 1024 // 4
src = torch.ones(n, dtype=torch.float32, device=dev)

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

if rank == 0:
    other = (lr + 1) % torch.cuda.device_count()
    print("devices", torch.cuda.device_count(), "can_access_peer", lr, "->", other, torch.cuda.can_device_access_peer(lr, other), flush=True)
    dst_local = torch.empty(n, dtype=torch.float32, device="cuda:%d" % other)          # same process, other device
    ms = timed(lambda: dst_local.copy_(src, non_blocking=True))
    print("same-process cross-device copy: %.3f ms = %.0f GB/s" % (ms, n * 4 / ms / 1e6), flush=True)
if world > 1:
    buf = torch.empty(n, dtype=torch.float32, device=dev)
    mine = reduce_tensor(buf)[1]
    every = [None] * world
    dist.all_gather_object(every, mine)
    peer = rebuild_cuda_tensor(*every[(rank + 1) % world])
    print("rank", rank, "peer tensor device", peer.device, flush=True)
    ms = timed(lambda: peer.copy_(src, non_blocking=True))
    print("rank %d IPC push via tensor.copy_: %.3f ms = %.0f GB/s" % (rank, ms, n * 4 / ms / 1e6), flush=True)
    try:
        from cuda import cudart
        st = torch.cuda.current_stream().cuda_stream
        def raw():
            err, = cudart.cudaMemcpyPeerAsync(peer.data_ptr(), peer.device.index, src.data_ptr(), lr, n * 4, st)
            assert int(err) == 0, err
        ms = timed(raw)
        print("rank %d IPC push via cudaMemcpyPeerAsync: %.3f ms = %.0f GB/s" % (rank, ms, n * 4 / ms / 1e6), flush=True)
        def raw2():
            err, = cudart.cudaMemcpyAsync(peer.data_ptr(), src.data_ptr(), n * 4, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice, st)
            assert int(err) == 0, err
        ms = timed(raw2)
        print("rank %d IPC push via cudaMemcpyAsync(D2D): %.3f ms = %.0f GB/s" % (rank, ms, n * 4 / ms / 1e6), flush=True)
    except Exception as exc:
        print("cuda-python path failed:", repr(exc), flush=True)
    try:
        import torch.distributed._symmetric_memory as symm_mem
        sbuf = symm_mem.empty(n, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(sbuf, dist.group.WORLD)
        sbuf.copy_(src); hdl.barrier(); torch.cuda.synchronize()
        remote = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
        dst = torch.empty(n, dtype=torch.float32, device=dev)
        ms = timed(lambda: dst.copy_(remote, non_blocking=True))
        print("rank %d symmetric-memory pull via tensor.copy_: %.3f ms = %.0f GB/s (value check %s)" % (rank, ms, n * 4 / ms / 1e6, bool((dst == 1).all())), flush=True)
        ms = timed(lambda: remote.copy_(src, non_blocking=True))
        print("rank %d symmetric-memory push via tensor.copy_: %.3f ms = %.0f GB/s" % (rank, ms, n * 4 / ms / 1e6), flush=True)
        hdl.barrier(); torch.cuda.synchronize()
    except Exception as exc:
        print("symmetric memory path failed:", repr(exc), flush=True)
    out = torch.empty(world * n, dtype=torch.float32, device=dev)
    ms = timed(lambda: dist.all_gather_into_tensor(out, src))
    print("rank %d nccl all_gather_into_tensor: %.3f ms" % (rank, ms), flush=True)
    dist.barrier()
    dist.destroy_process_group()

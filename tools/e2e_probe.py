"""Where the end-to-end step time goes (host actions -> GpuVecEnv.step -> host results): wall-clock
stamps between the stages of step_async / step_wait, mean microseconds over the timed steps.
    python tools/e2e_probe.py [steps]
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from irbpp_b200 import shapes  # noqa: E402
from irbpp_b200.vec_env import GpuVecEnv, LazyInfos  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    dev = torch.device("cuda:0")
    lib = shapes.make_blockout_library(32, seed=1)
    N_ENVS = 4096
    seqs = shapes.make_sequences(N_ENVS, 128, lib.num_shapes, seed=0)
    env = GpuVecEnv(lib, seqs, device="cuda:0")
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    obs = env.reset()
    for _ in range(bench.BURN_IN):
        obs, _ = env.step_device(bench.device_policy(torch, obs, gen))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    names = ["actions_arg", "take_obs + stream", "step_async (memcpy + 2 launches)", "spare obs alloc (overlapped)",
             "step_wait (sync + error scan)", "reward / done copies + views", "total"]
    acc = np.zeros(len(names))
    n = env.num_envs
    for k in range(steps + 5):
        acts = bench.device_policy(torch, obs, gen).cpu().numpy()
        flush.fill_(k & 255)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        keep, ptr, on_dev = env._actions_arg(acts, "actions")
        t1 = time.perf_counter()
        o = env._take_obs()
        stream = env._stream()
        t2 = time.perf_counter()
        env._check(env._lib.irbpp_step_async(env._h, ptr, on_dev, o.data_ptr(), stream))
        t3 = time.perf_counter()
        env._spare_obs = env._new_obs(env.obs_len); env._spare_stream = env._stream()
        t_rel = round(time.time() - env._tstart, 6)
        t3b = time.perf_counter()
        res = env._result
        env._check(env._lib.irbpp_step_wait(env._h, ctypes.byref(res)))
        t4 = time.perf_counter()
        src, offs = env._result_block(res)
        reward = src[offs["reward"]:offs["reward"] + 4 * n].view(np.float32).copy()
        done = src[offs["done"]:offs["done"] + n].view(np.bool_).copy()
        infos = LazyInfos(n, src, offs, done, t_rel, borrowed=True)
        rew = torch.from_numpy(reward).unsqueeze(dim=1)
        t5 = time.perf_counter()
        obs = o
        if k >= 5:
            acc += [t1 - t0, t2 - t1, t3 - t2, t3b - t3, t4 - t3b, t5 - t4, t5 - t0]
    # alternative host path: actions through an explicit pinned H2D copy, results through one D2H copy
    pin = torch.empty(N_ENVS, dtype=torch.int64).pin_memory()
    a_dev = torch.empty(N_ENVS, dtype=torch.int64, device=dev)
    pin_np = pin.numpy()
    staged = np.zeros(3)
    for k in range(steps + 5):
        acts = bench.device_policy(torch, obs, gen).cpu().numpy()
        flush.fill_(k & 255)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        pin_np[:] = acts
        a_dev.copy_(pin, non_blocking=True)
        o = env._new_obs(env.obs_len)
        env._check(env._lib.irbpp_step_async(env._h, a_dev.data_ptr(), 1, o.data_ptr(), env._stream()))
        t3 = time.perf_counter()
        res = env._result
        env._check(env._lib.irbpp_step_wait(env._h, ctypes.byref(res)))
        t4 = time.perf_counter()
        src, offs = env._result_block(res)
        block = src.copy()
        t5 = time.perf_counter()
        obs = o
        if k >= 5:
            staged += [t3 - t0, t4 - t3, t5 - t0]
    # the same step with everything on the device, timed by events, for comparison
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_ms = 0.0
    for k in range(steps):
        a = bench.device_policy(torch, obs, gen)
        flush.fill_(k & 255)
        e0.record(); obs, _ = env.step_device(a); e1.record()
        torch.cuda.synchronize(dev)
        dev_ms += e0.elapsed_time(e1)
    for n, a in zip(names, acc):
        print("%-36s %8.1f us" % (n, 1e6 * a / steps))
    print("%-36s %8.1f us" % ("device-resident step (events)", 1e3 * dev_ms / steps))
    print("staged-copy host path: submit %.1f us, wait %.1f us, total %.1f us" % tuple(1e6 * staged / steps))


if __name__ == "__main__":
    main()

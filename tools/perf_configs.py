#!/usr/bin/env python
"""Device time per batched step for the BASELINE.json configurations other than the bench line
(parity cases, timed here only to know where they stand).  GPU only."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irbpp_b200 import shapes
from irbpp_b200.vec_env import GpuVecEnv

def policy(obs, sel, gen):
    n = obs.shape[0]
    mask = obs[:, :sel * 5].view(n, sel, 5)[:, :, 4] == 1
    return torch.argmax(torch.rand((n, sel), device=obs.device, generator=gen) + mask.float(), dim=1)

def run(name, lib, n, k=1, steps=30, burn=120):
    seqs = shapes.make_sequences(n, 128, lib.num_shapes, seed=0)
    env = GpuVecEnv(lib, seqs, device="cuda:0", buffer_size=k)
    gen = torch.Generator(device="cuda:0"); gen.manual_seed(1)
    obs = env.reset()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda:0")
    def one():
        nonlocal obs
        if k > 1:
            order = torch.randint(0, k, (n,), device="cuda:0", generator=gen)
            loc = env.get_action_candidates(order, as_tensor=True)
            acts = policy(loc, 500, gen)
        else:
            acts = policy(obs, 500, gen)
        obs, _ = env.step_device(acts)
    for _ in range(burn): one()
    torch.cuda.synchronize()
    t = 0.0
    for s in range(steps):
        flush.fill_(float(s))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); one(); b.record(); torch.cuda.synchronize()
        t += a.elapsed_time(b)
    ms = t / steps
    print(json.dumps({"config": name, "bins": n, "R": lib.num_rotations, "k": k, "ms_per_step_incl_policy": round(ms, 4),
                      "env_steps_per_s": round(n / ms * 1e3)}))
    env.close()

run("BlockOut online (bench line, incl. torch policy)", shapes.make_blockout_library(32, seed=1), 4096)
run("General-like irregular, R=8", shapes.make_irregular_library(32, seed=2), 4096)
run("Cube, R=2", shapes.make_cube_library(seed=3), 4096)
run("BlockOut buffered k=10 (get_action_candidates + step)", shapes.make_blockout_library(32, seed=1), 4096, k=10)
run("irregular R=8, 32768 bins on one GPU", shapes.make_irregular_library(32, seed=2), 32768, steps=10, burn=60)

#!/usr/bin/env python
"""Per-phase SM cycles of irbpp_env_kernel on the bench workload (steady state).  GPU only."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from irbpp_b200 import shapes
from irbpp_b200.vec_env import GpuVecEnv
lib = bench.workload()
seqs = shapes.make_sequences(bench.N_ENVS, bench.SEQ_LEN, lib.num_shapes, seed=0)
env = GpuVecEnv(lib, seqs, device="cuda:0")
gen = torch.Generator(device="cuda:0"); gen.manual_seed(1)
obs = env.reset()
for _ in range(150):
    obs, _ = env.step_device(bench.device_policy(torch, obs, gen))
env.debug_phase_cycles(True)
n = 20
for _ in range(n):
    obs, _ = env.step_device(bench.device_policy(torch, obs, gen))
c = env.debug_phase_cycles(False).astype(np.float64) / (n * bench.N_ENVS)
names = ["scan:load+apply", "scan:obs+scan+bitmaps", "cand:contours(per 4 bins)", "cand:select/pad(per 4 bins)", "warp0:find-start(per 4 bins)", "warp0:follow(per 4 bins)", "warp0:approx(per 4 bins)"]
print(json.dumps({"lib": os.environ.get("IRBPP_LIB", "default"), "cycles_per_cta": dict(zip(names, [round(float(v)) for v in c[:7]])), "note": "candidates-kernel counters are summed over N/4 CTAs but divided by N: multiply by 4 for cycles per CTA"}))

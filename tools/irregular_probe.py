#!/usr/bin/env python
"""Kernel time split for the irregular R=8 configuration (per-cell scan path).  GPU only; run under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irbpp_b200 import shapes
from irbpp_b200.vec_env import GpuVecEnv
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lib = shapes.make_irregular_library(32, seed=2, num_rotations=R)
n = 4096
seqs = shapes.make_sequences(n, 128, lib.num_shapes, seed=0)
env = GpuVecEnv(lib, seqs, device="cuda:0")
gen = torch.Generator(device="cuda:0"); gen.manual_seed(1)
obs = env.reset()
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 60):
    mask = obs[:, :2500].view(n, 500, 5)[:, :, 4] == 1
    acts = torch.argmax(torch.rand((n, 500), device="cuda:0", generator=gen) + mask.float(), dim=1)
    obs, _ = env.step_device(acts)
torch.cuda.synchronize()
env.debug_phase_cycles(True)
for _ in range(5):
    mask = obs[:, :2500].view(n, 500, 5)[:, :, 4] == 1
    acts = torch.argmax(torch.rand((n, 500), device="cuda:0", generator=gen) + mask.float(), dim=1)
    obs, _ = env.step_device(acts)
c = env.debug_phase_cycles(False)
print("per step: images %.0f  micro-tasks %.0f  rounds %.0f  overflow redos %.1f" % (c[4]/5, c[5]/5, c[6]/5, c[7]/5))
k = (obs[:, :2500].view(n, 500, 5)[:, :, 4] == 1).sum(dim=1).float()
print("valid candidates per bin: mean %.1f max %d" % (k.mean().item(), int(k.max().item())))

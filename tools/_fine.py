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
c = env.debug_phase_cycles(False).astype(np.float64) / n
print(os.environ.get("IRBPP_LIB", "default"))
print("scan per CTA: A %.0f  rest %.0f" % (c[0] / 4096, c[1] / 4096))
print("cand per CTA: C %.0f  D %.0f | slots4-7 per CTA: %s | raw per step %s" % (c[2] / 1024, c[3] / 1024, [round(v / 1024, 1) for v in c[4:8]], [round(v) for v in c[4:8]]))

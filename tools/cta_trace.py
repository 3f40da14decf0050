#!/usr/bin/env python
"""Per-CTA timeline of the candidates kernel (profiling build -DIRBPP_PROBE_TRACE; GPU only):
    IRBPP_LIB=.../libirbpp_trace.so python tools/cta_trace.py [config]
One steady-state step is traced: for every CTA the globaltimer at kernel entry, after the dependency wait, after the
prologue, after the first follow phase, after the first approximation phase, at the end of phase C and at the end,
plus the task count and the longest contour of its first batch."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from irbpp_b200 import shapes
from irbpp_b200.vec_env import GpuVecEnv

config = sys.argv[1] if len(sys.argv) > 1 else "blockout"
n = bench.CONFIGS[config]["bins"]
lib = bench.make_library(config)
env = GpuVecEnv(lib, shapes.make_sequences(n, 128, lib.num_shapes, seed=0), device="cuda:0")
gen = torch.Generator(device="cuda:0"); gen.manual_seed(1)
obs = env.reset()
for _ in range(150):
    obs, _ = env.step_device(bench.device_policy(torch, obs, gen))
path = os.path.join(tempfile.mkdtemp(), "trace.bin")
os.environ["IRBPP_TRACE_FILE"] = path
res = []
for rep in range(3):
    env.debug_phase_cycles(True)                     # clears the buffer, enables
    acts = bench.device_policy(torch, obs, gen)
    torch.cuda.synchronize()
    obs, _ = env.step_device(acts)
    torch.cuda.synchronize()
    env.debug_phase_cycles(False)                    # dumps the trace
    tr = np.fromfile(path, dtype=np.uint64)[8:].reshape(-1, 8).astype(np.int64)
    tr = tr[tr[:, 6] > 0]                            # the CTAs that ran (the grid depends on the bins per CTA)
    t0 = tr[:, 0].min()
    rel = (tr[:, :7] - t0) / 1e3                     # microseconds since the first CTA started
    dur = rel[:, 6] - rel[:, 0]
    mx = tr[:, 7] & 255; ntask = tr[:, 7] >> 8
    out = {"rep": rep, "kernel_us": float(rel[:, 6].max()), "cta_start_us_max": float(rel[:, 0].max()),
           "wait_end_us": [float(np.percentile(rel[:, 1], q)) for q in (50, 100)],
           "cta_dur_us": {q: float(np.percentile(dur, q)) for q in (10, 50, 90, 99, 100)},
           "phase_us_mean": {"wait": float((rel[:, 1] - rel[:, 0]).mean()), "prologue": float((rel[:, 2] - rel[:, 1]).mean()),
                             "follow": float((rel[:, 3] - rel[:, 2]).mean()), "approx(+sort)": float((rel[:, 4] - rel[:, 3]).mean()),
                             "rest of C": float((rel[:, 5] - rel[:, 4]).mean()), "D": float((rel[:, 6] - rel[:, 5]).mean())},
           "longest_contour": {q: int(np.percentile(mx, q)) for q in (50, 90, 99, 100)},
           "tasks_per_cta": {q: int(np.percentile(ntask, q)) for q in (50, 100)}}
    # the slowest CTAs: what made them slow
    worst = np.argsort(-rel[:, 6])[:5]
    out["slowest"] = [{"cta": int(i), "end_us": float(rel[i, 6]), "follow": float(rel[i, 3] - rel[i, 2]), "approx": float(rel[i, 4] - rel[i, 3]),
                       "restC": float(rel[i, 5] - rel[i, 4]), "D": float(rel[i, 6] - rel[i, 5]), "maxn": int(mx[i]), "ntask": int(ntask[i])} for i in worst]
    # mean follow / approx time by longest contour
    by = {}
    for lo, hi in ((0, 8), (9, 12), (13, 16), (17, 24), (25, 64)):
        sel = (mx >= lo) & (mx <= hi)
        if sel.any():
            by["%d-%d" % (lo, hi)] = {"ctas": int(sel.sum()), "follow": float((rel[sel, 3] - rel[sel, 2]).mean()),
                                      "approx": float((rel[sel, 4] - rel[sel, 3]).mean()), "end": float(rel[sel, 6].mean())}
    out["by_longest_contour"] = by
    res.append(out)
print(json.dumps(res, indent=1))

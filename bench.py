#!/usr/bin/env python
"""bench.py -- env-steps/s of the packing-environment hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # B200 arm, one JSON line on rank 0
    python bench.py --impl reference [--gpus N] --steps K --warmup W  # reference CPU arm (oracle port)

A "step" is one batched ``step()`` of 4096 BlockOut bins per GPU (BASELINE.json configs[1]: online,
selectedAction=500, R=4; synthetic polycube shapes and sequences).  For N > 1 the
driver launches one rank per GPU with torch.distributed.run; bins are sharded by index (weak scaling:
4096 per GPU), no collective inside a step, one NCCL all-gather of the observations at rollout end
(inside the timed region).

Timing: every timed step is bracketed by CUDA events on the launching stream; between timed steps
the L2 is flushed by writing a 256 MiB buffer (outside the event pairs).  ``value`` uses actions that
are already on the device; ``e2e`` calls the public ``GpuVecEnv.step`` with HOST actions (pinned H2D
copy, kernel, D2H of reward/done/info arrays) and is timed with the host clock around the call.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ENVS = 4096            # bins per GPU (BASELINE.json metric)
SEL = 500
SEQ_LEN = 128
CPU_BURN_IN = 40         # untimed batched steps of the CPU arm before its timed sample
BURN_IN = 150            # untimed steps before the W warm-up steps so episodes are in steady state
METRIC = "env steps/sec (4096 bins, BlockOut)"
UNIT = "env-steps/s"


def workload():
    from irbpp_b200 import shapes
    lib = shapes.make_blockout_library(32, seed=1, num_rotations=4)
    return lib


def algorithmic_bytes_per_env_step(lib, sel=SEL):
    """2*Hx*Hy*8 [hm r+w] + R*Ax*Ay*16 [posZ+mask] + sum_r 2*w_r*h_r*8 [B, maskB of the
    next item] + 2*w*h*8 [T, maskT of the placed rotation] + obs_len*4 [float32 observation], averaged
    over the shape library (ids are uniform)."""
    R = lib.num_rotations
    wh = (lib.dims[:, :, 0].astype(np.float64) * lib.dims[:, :, 1])
    obs_len = sel * 5 + 9 + 1024
    return 2 * 32 * 32 * 8 + R * 16 * 16 * 16 + float(wh.sum(axis=1).mean()) * 16 + float(wh.mean()) * 16 + obs_len * 4


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); smax = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def device_policy(torch, obs, gen):
    """Stand-in for the agent (trainer.py:161-162): a uniformly random candidate with V == 1 (row 0 if
    none), computed on the device from the observation's mask column."""
    n = obs.shape[0]
    mask = obs[:, :SEL * 5].view(n, SEL, 5)[:, :, 4] == 1
    score = torch.rand((n, SEL), device=obs.device, generator=gen) + mask.float()
    return torch.argmax(score, dim=1)


def host_policy(rng, obs):
    n = obs.shape[0]
    mask = obs[:, :SEL * 5].reshape(n, SEL, 5)[:, :, 4] == 1
    score = rng.random((n, SEL)) + mask
    return np.argmax(score, axis=1).astype(np.int64)


# ---------------------------------------------------------------------------------------------------
# CPU reference arm / baseline (oracle port; the reference itself is Python that cannot be installed:
# it imports trimesh / gym / pybullet which are absent)
# ---------------------------------------------------------------------------------------------------

def run_cpu_arm(steps, warmup, envs_per_core=2, max_seconds=None):
    from irbpp_b200 import shapes
    from oracle.cpu_vec_env import SubprocOracleVecEnv
    lib = workload()
    cores = os.cpu_count() or 1
    n = cores * envs_per_core
    seqs = shapes.make_sequences(n, SEQ_LEN, lib.num_shapes, seed=0)
    vec = SubprocOracleVecEnv(dict(ZRotNum=4, selectedAction=SEL), lib, seqs, num_procs=cores)
    rng = np.random.default_rng(0)
    obs = vec.reset()
    for _ in range(CPU_BURN_IN + warmup):            # untimed: reach a steady mix of episode phases
        obs, _, _, _ = vec.step(host_policy(rng, obs))
    t_total, done_steps = 0.0, 0
    for _ in range(steps):
        acts = host_policy(rng, obs)
        t0 = time.perf_counter()
        obs, _, _, _ = vec.step(acts)
        t_total += time.perf_counter() - t0
        done_steps += 1
        if max_seconds is not None and t_total > max_seconds:
            break
    vec.close()
    value = n * done_steps / t_total
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d bins (%d per core) x %d batched steps of the same BlockOut workload, one worker process "
                      "per core over pipes (ShmemVecEnv shape), oracle port with reference-style loops + cv2, "
                      "no PyBullet" % (n, envs_per_core, done_steps),
            "steps": done_steps, "ms_per_step": 1e3 * t_total / done_steps, "n_envs": n}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    res = run_cpu_arm(args.steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": res["steps"], "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BlockOut-like polycubes, online, selectedAction=500, R=4 (bounded sample: "
                                   "%d bins on %d host cores)" % (res["n_envs"], res["cores"])},
            "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
                             "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------

def main_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    cpu_base = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process (workers are forked)
        cpu_base = run_cpu_arm(steps=10 ** 9, warmup=2, max_seconds=args.cpu_seconds)
        for k in ("steps", "ms_per_step", "n_envs"):
            cpu_base.pop(k, None)

    import torch
    import torch.distributed as dist
    from irbpp_b200 import shapes, sharding
    from irbpp_b200.vec_env import GpuVecEnv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lib = workload()
    n_total = N_ENVS * world
    seqs_all = shapes.make_sequences(n_total, SEQ_LEN, lib.num_shapes, seed=0)
    seqs = sharding.shard_sequences(seqs_all, rank, world)
    env = GpuVecEnv(lib, seqs, device=dev, selected_action=SEL)
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident loop ("value") ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                             # sampled from burn-in through the timed steps (same load)
    obs = env.reset()
    for _ in range(BURN_IN + args.warmup):          # burn-in: bins spread over all episode phases
        acts = device_policy(torch, obs, gen)
        obs, _ = env.step_device(acts)
    if world > 1:                                   # warm-up of the rollout-end collective (NCCL channel setup)
        sharding.gather_rollout(obs, world)
    torch.cuda.synchronize(dev)
    launches0 = env.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        acts = device_policy(torch, obs, gen)
        flush.fill_(float(k))                       # L2 flush, outside the event pair
        ev[k][0].record()
        obs, _ = env.step_device(acts)
        ev[k][1].record()
    gather_ms = 0.0
    if world > 1:                                  # end-of-rollout gather of observations (NCCL all-gather)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        gathered = sharding.gather_rollout(obs, world)
        g1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches_timed = env.launch_count() - launches0
    # the timed region lasts ~10-30 ms: keep the same loop running (untimed) until nvidia-smi has
    # delivered enough clock samples under this load
    extra_steps = 0
    if rank == 0:
        t_end = time.perf_counter() + 1.5
        while len(sampler.rows) < 25 and time.perf_counter() < t_end:
            for _ in range(20):
                acts = device_policy(torch, obs, gen)
                obs, _ = env.step_device(acts)
            torch.cuda.synchronize(dev)
            extra_steps += 20
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "burn-in + timed steps + %d further untimed steps of the same loop" % extra_steps
    launches = launches_timed
    step_ms = [a.elapsed_time(b) for a, b in ev]
    if world > 1:
        gather_ms = g0.elapsed_time(g1)
        assert gathered.shape[0] == n_total
    t_dev_ms = float(sum(step_ms)) + gather_ms
    t = torch.tensor([t_dev_ms, float(np.mean(step_ms)), gather_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_dev_ms, kern_ms, gather_ms = [float(v) for v in t.tolist()]
    value = n_total * args.steps / (t_dev_ms * 1e-3)

    # ---- end-to-end loop through the public API with host actions ----
    e2e_steps = args.steps
    t_e2e = 0.0
    for k in range(min(args.warmup, 3)):
        acts = device_policy(torch, obs, gen).cpu().numpy()
        obs, rew, done, infos = env.step(acts)
    for k in range(e2e_steps):
        acts = device_policy(torch, obs, gen).cpu().numpy()        # agent -> action.cpu().numpy() (trainer.py:165)
        flush.fill_(float(k))
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        obs, rew, done, infos = env.step(acts)                      # H2D actions, kernel, D2H results
        torch.cuda.synchronize(dev)
        t_e2e += time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    e2e_value = n_total * e2e_steps / t_e2e
    h2d = N_ENVS * 8                                                # int64 actions
    d2h = N_ENVS * (8 + 8 + 4 + 4 + 4 + 1 + 1 + 1)                  # ratio, ep_reward, reward, counter, ep_len, done, valid, error

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        bytes_step = algorithmic_bytes_per_env_step(lib) * N_ENVS
        achieved = bytes_step / (kern_ms * 1e-3) / 1e9
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")))
            traffic = prof.get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "BlockOut-like polycubes (32 shapes, 2-5 cells of 0.04 m), 4096 bins per GPU, "
                                       "online (bufferSize=1), selectedAction=500, R=4, bin 0.32x0.32x0.30, "
                                       "random-valid policy", "bins_per_gpu": N_ENVS, "burn_in_steps": BURN_IN, "parallelism": "env-shard x%d" % world,
                           "l2": "flushed between timed steps by a 256 MiB write outside the event pairs",
                           "timing": "sum of per-step CUDA-event intervals on the launching stream"
                                     + (" + end-of-rollout all-gather" if world > 1 else "") + ", max over ranks"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src,
                             "kernel": "irbpp_scan_kernel + irbpp_candidates_kernel (one step = both launches)",
                             "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": bytes_step},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * t_e2e / e2e_steps,
                        "note": "GpuVecEnv.step(host int64 actions): actions staged in pinned memory and read by the kernel over "
                                "PCIe, both kernels, reward/done/info arrays written back over PCIe, stream sync; observations "
                                "stay on the device as in the reference (envs.py:163)"},
                "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall}
        if world > 1:
            line["gather_ms"] = gather_ms
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line))
    env.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="bound of the cpu_baseline sample (wall seconds)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())

"""CPU vector environment used ONLY as the timed baseline (``bench.py`` cpu_baseline /
``--impl reference``) -- TEST / MEASUREMENT INFRASTRUCTURE, never imported by the product.

It has the shape of the reference's ``ShmemVecEnv`` (``wrapper/shmem_vec_env.py:20-157``): one OS
process per group of bins, commands and pickled float64 observations over ``multiprocessing.Pipe``,
auto-reset on done in the worker (``:140-144``).  Each worker steps oracle envs whose geometry mirrors
the reference's own implementation style: the rot x X x Y Python loop with NumPy window maxima
(``space.py:98-129`` -> ``scan_loops``) and cv2 contour calls (``cvTools.py:77-103`` ->
``convex_hulls_cv2``; the cv2-free port when cv2 is missing).  The PyBullet settle is not part of it,
so this baseline is an UPPER bound on the real reference's speed."""
import multiprocessing as mp

import numpy as np

from .oracle_env import HAVE_CV2, OracleConfig, OracleEnv, PortGeometry


def _worker(pipe, cfg_kwargs, lib, sequences):
    cfg = OracleConfig(**cfg_kwargs)
    contours = "cv2" if HAVE_CV2 else "port"
    envs = [OracleEnv(cfg, lib, s, PortGeometry(cfg, lib, scan="loops", contours=contours)) for s in sequences]
    try:
        while True:
            cmd, data = pipe.recv()
            if cmd == "reset":
                pipe.send([e.reset() for e in envs])
            elif cmd == "step":
                out = []
                for e, a in zip(envs, data):
                    obs, rew, done, info = e.step(int(a))
                    if done:
                        obs = e.reset()
                    out.append((obs, rew, done, info))
                pipe.send(out)
            elif cmd == "get_action_candidates":          # shmem_vec_env.py:99-102,149-150
                pipe.send([e.get_action_candidates(int(a)) for e, a in zip(envs, data)])
            elif cmd == "close":
                pipe.send(None)
                break
            else:
                raise RuntimeError("unknown cmd %s" % cmd)
    except KeyboardInterrupt:
        pass


class SubprocOracleVecEnv(object):
    def __init__(self, cfg_kwargs, lib, sequences, num_procs, context="fork"):
        ctx = mp.get_context(context)
        n = len(sequences)
        self.num_envs = n
        self.num_procs = max(1, min(num_procs, n))
        bounds = np.linspace(0, n, self.num_procs + 1).astype(int)
        self.slices = [(int(bounds[i]), int(bounds[i + 1])) for i in range(self.num_procs)]
        self.pipes, self.procs = [], []
        for lo, hi in self.slices:
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(child, cfg_kwargs, lib, [sequences[i] for i in range(lo, hi)]))
            p.daemon = True
            p.start()
            child.close()
            self.pipes.append(parent); self.procs.append(p)

    def reset(self):
        for p in self.pipes:
            p.send(("reset", None))
        obs = []
        for p in self.pipes:
            obs.extend(p.recv())
        return np.stack(obs)

    def step(self, actions):
        for p, (lo, hi) in zip(self.pipes, self.slices):
            p.send(("step", list(actions[lo:hi])))
        outs = []
        for p in self.pipes:
            outs.extend(p.recv())
        obs, rews, dones, infos = zip(*outs)
        return np.stack(obs), np.array(rews), np.array(dones), infos

    def get_action_candidates(self, order_actions):
        for p, (lo, hi) in zip(self.pipes, self.slices):
            p.send(("get_action_candidates", list(order_actions[lo:hi])))
        out = []
        for p in self.pipes:
            out.extend(p.recv())
        return np.stack(out)

    def close(self):
        for p in self.pipes:
            try:
                p.send(("close", None)); p.recv(); p.close()
            except Exception:
                pass
        for pr in self.procs:
            pr.join(timeout=5)

#!/usr/bin/env python
"""Offline fuzz of the device contour routines (compiled for the host by tests/host_harness/contour_host.cpp)
against the cv2-pinned oracle port: python tools/fuzz_contours.py [images] [seed].  CPU only, test tooling."""
import ctypes, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import contours_port as cp

def build():
    out = os.path.join(tempfile.mkdtemp(), "contour_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-o", out,
                    os.path.join(ROOT, "tests", "host_harness", "contour_host.cpp")], check=True)
    return ctypes.CDLL(out)

def image(rng, kind):
    if kind == 0:
        return (rng.random((16, 16)) < rng.uniform(0.05, 0.95)).astype(np.uint8)
    if kind == 1:
        img = np.zeros((16, 16), np.uint8)
        for _ in range(int(rng.integers(1, 9))):
            x0, y0 = rng.integers(0, 14, 2); w, h = rng.integers(1, 9, 2)
            img[x0:x0 + w, y0:y0 + h] = rng.integers(0, 2)
        return img
    if kind == 2:
        img = np.ones((16, 16), np.uint8)
        for _ in range(int(rng.integers(1, 9))):
            x0, y0 = rng.integers(0, 15, 2); w, h = rng.integers(1, 5, 2)
            img[x0:x0 + w, y0:y0 + h] = 0
        return img
    if kind == 3:     # 2x2-blocky plateaus (BlockOut-like level sets)
        return np.kron((rng.random((8, 8)) < rng.uniform(0.3, 0.9)).astype(np.uint8), np.ones((2, 2), np.uint8))
    xs, ys = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    f = np.zeros((16, 16))
    for _ in range(int(rng.integers(1, 5))):
        cx, cy = rng.uniform(0, 16, 2)
        f += rng.uniform(0.5, 1.5) * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / rng.uniform(4, 60))
    return (f > rng.uniform(0.2, 0.9)).astype(np.uint8)

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lib = build()
    rng = np.random.default_rng(seed)
    out = np.zeros(8, np.uint32)
    stats = {"images": 0, "contours": 0, "long": 0}
    for it in range(n):
        img = image(rng, it % 5)
        rows = np.zeros(16, np.uint16)
        for y in range(16):
            rows[y] = sum(1 << x for x in range(16) if img[y, x])
        conts = cp.find_outer_contours(img)
        for legacy in (0, 1):
            want = set()
            for c in conts:
                want |= {(int(p[0]), int(p[1])) for p in cp.convex_vertices(cp.approx_poly_dp_closed(c, 1.0, bool(legacy)))}
            for mode in (5, 6):
                rc = lib.hull_bits(rows.ctypes.data_as(ctypes.c_void_p), legacy, mode, out.ctypes.data_as(ctypes.c_void_p))
                if rc != 0:
                    assert max(len(c) for c in conts) > 64, (it, legacy, mode)
                    continue
                got = {(b >> 4, b & 15) for b in range(256) if (int(out[b >> 5]) >> (b & 31)) & 1}
                assert got == want, (it, legacy, mode, sorted(got ^ want))
        stats["images"] += 1; stats["contours"] += len(conts); stats["long"] += sum(1 for c in conts if len(c) > 16)
    print("ok", stats)

if __name__ == "__main__":
    main()

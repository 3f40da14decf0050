"""Host-side checks that need no GPU: the C-ABI library builds, loads and exports every symbol that
include/irbpp.h declares; the device contour routines (compiled as host C++ by a test-only harness)
agree with the oracle."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built_lib():
    from irbpp_b200 import build
    return build.build()


def test_header_symbols_exported(built_lib):
    from irbpp_b200 import _lib
    header = open(os.path.join(ROOT, "include", "irbpp.h")).read()
    declared = set(re.findall(r"\b(irbpp_[a-z_]+)\s*\(", header))
    declared.discard("irbpp_env")
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().irbpp_abi_version() == _lib.ABI_VERSION


def test_create_rejects_bad_config_without_gpu(built_lib):
    """Argument validation happens before any CUDA call, so it is checkable here."""
    from irbpp_b200 import _lib
    lib = _lib.load()
    cfg = _lib.IrbppConfig()
    cfg.num_envs, cfg.num_rotations, cfg.selected_action, cfg.buffer_size = 4, 4, 500, 1
    cfg.bin_dimension[0], cfg.bin_dimension[1], cfg.bin_dimension[2] = 0.32, 0.32, 0.30
    cfg.resolution_act, cfg.resolution_h, cfg.resolution_z = 0.02, 0.01, 0.01
    h = ctypes.c_void_p()
    cfg.num_rotations = 0
    assert lib.irbpp_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.IRBPP_EINVAL
    cfg.num_rotations = 4
    cfg.resolution_h = 0.02      # heightmap would be 16x16: unsupported grid
    assert lib.irbpp_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.IRBPP_EINVAL
    assert b"unsupported grid" in lib.irbpp_last_error(None)
    cfg.resolution_h = 0.01
    cfg.selected_action = 2000   # more rows than poses
    assert lib.irbpp_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.IRBPP_EINVAL


def test_no_cpu_fallback_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from irbpp_b200 import _lib
    lib = _lib.load()
    cfg = _lib.IrbppConfig()
    cfg.num_envs, cfg.num_rotations, cfg.selected_action, cfg.buffer_size = 4, 4, 500, 1
    cfg.bin_dimension[0], cfg.bin_dimension[1], cfg.bin_dimension[2] = 0.32, 0.32, 0.30
    cfg.resolution_act, cfg.resolution_h, cfg.resolution_z = 0.02, 0.01, 0.01
    h = ctypes.c_void_p()
    assert lib.irbpp_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.IRBPP_ECUDA
    assert b"no CPU path" in lib.irbpp_last_error(None)


@pytest.fixture(scope="module")
def contour_harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("harness") / "contour_host.so")
    src = os.path.join(ROOT, "tests", "host_harness", "contour_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-o", out, src], check=True)
    return ctypes.CDLL(out)


def _dev_bits(lib, img, legacy, big):
    rows = np.zeros(16, np.uint16)
    for y in range(16):
        rows[y] = sum(1 << x for x in range(16) if img[y, x])
    out = np.zeros(8, np.uint32)
    rc = lib.hull_bits(rows.ctypes.data_as(ctypes.c_void_p), legacy, big, out.ctypes.data_as(ctypes.c_void_p))
    return rc, {(b >> 4, b & 15) for b in range(256) if (int(out[b >> 5]) >> (b & 31)) & 1}


def test_device_contour_routines_match_oracle(contour_harness):
    from oracle import contours_port as cp
    rng = np.random.default_rng(0)
    n_ovf = 0
    for it in range(1200):
        kind = it % 4
        if kind == 0:
            img = (rng.random((16, 16)) < rng.uniform(0.05, 0.95)).astype(np.uint8)
        elif kind == 1:
            img = np.zeros((16, 16), np.uint8)
            for _ in range(int(rng.integers(1, 7))):
                x0, y0 = rng.integers(0, 14, 2); w, h = rng.integers(1, 9, 2)
                img[x0:x0 + w, y0:y0 + h] = rng.integers(0, 2)
        elif kind == 2:
            img = np.ones((16, 16), np.uint8)
            for _ in range(int(rng.integers(1, 9))):
                x0, y0 = rng.integers(0, 15, 2); w, h = rng.integers(1, 5, 2)
                img[x0:x0 + w, y0:y0 + h] = 0
        elif it % 8 == 7:   # nested rings and islands, noise
            img = np.zeros((16, 16), np.uint8)
            a = int(rng.integers(0, 3)); img[a:16 - a, a:16 - a] = 1
            b = a + int(rng.integers(1, 3)); img[b:16 - b, b:16 - b] = 0
            c = b + int(rng.integers(1, 3)); img[c:16 - c, c:16 - c] = 1
            d = c + 1; img[d:16 - d, d:16 - d] = 0
            img ^= (rng.random((16, 16)) < 0.06).astype(np.uint8)
        else:
            img = ((np.add.outer(np.arange(16), np.arange(16)) % 2) == 0).astype(np.uint8)
            img &= (rng.random((16, 16)) < 0.9).astype(np.uint8)
        for legacy in (0, 1):
            want, maxlen = set(), 0
            for c in cp.find_outer_contours(img):
                maxlen = max(maxlen, len(c))
                want |= {(int(p[0]), int(p[1])) for p in cp.convex_vertices(cp.approx_poly_dp_closed(c, 1.0, bool(legacy)))}
            # one task per start pixel (what the kernels run): long-buffer and 64-point lane-scratch variants
            rc_big, got_big = _dev_bits(contour_harness, img, legacy, 5)
            assert rc_big == 0 and got_big == want
            rc_fast, got_fast = _dev_bits(contour_harness, img, legacy, 6)
            if rc_fast == 0:
                assert got_fast == want
            else:
                assert maxlen > 64                        # overflow of the 64-point buffer only
            n_ovf += rc_fast
    assert n_ovf > 0                                         # the overflow path was exercised


def test_device_pairwise_sum_equals_numpy(contour_harness):
    """np.sum(heightmapC_Prime) (space.py:217) is a pairwise reduction; the device restatement must
    associate identically for every window size up to 32 x 32."""
    contour_harness.pairwise_sum_host.restype = ctypes.c_double
    rng = np.random.default_rng(5)
    for n in list(range(0, 300)) + [511, 512, 513, 767, 1000, 1016, 1023, 1024]:
        for rep in range(3):
            a = np.ascontiguousarray(rng.random(n) * rng.choice([1.0, 1e-3, 37.0]))
            got = contour_harness.pairwise_sum_host(a.ctypes.data_as(ctypes.c_void_p), n)
            assert got == float(np.sum(a)), n


def test_device_heuristic_scores_match_oracle(contour_harness):
    """The per-pose arithmetic of irbpp_heuristic_kernel, compiled for the host, against the oracle's
    restatement of Space.get_heuristic_action (itself pinned to the reference by tests/golden)."""
    from irbpp_b200 import shapes
    from oracle.oracle_env import OracleConfig, OracleVecEnv, HEURISTICS
    lib = shapes.make_irregular_library(8, seed=5, num_rotations=4)
    R = lib.num_rotations
    cfg = OracleConfig(ZRotNum=R)
    env = OracleVecEnv(cfg, lib, shapes.make_sequences(2, 30, lib.num_shapes, seed=2))
    env.reset()
    for t in range(12):
        acts = []
        for e in env.envs:
            item = e.next_item_ID
            pool, offs, ws, hs = [], [], [], []
            for r in range(R):
                T, _, mT, _ = lib.tables[item][r]
                offs.append(sum(len(p) for p in pool)); ws.append(T.shape[0]); hs.append(T.shape[1])
                pool.append(np.where(mT != 0, T, -np.inf).reshape(-1))
            Ts = np.ascontiguousarray(np.concatenate(pool))
            offs = np.array(offs, np.int64); ws = np.array(ws, np.int32); hs = np.array(hs, np.int32)
            hm = np.ascontiguousarray(e.heightmap); posz = np.ascontiguousarray(e.posZmap)
            mask = np.ascontiguousarray(e.naiveMask.astype(np.uint8))
            for mi, m in enumerate(HEURISTICS):
                for d in range(4):
                    got = contour_harness.heuristic_pose_host(
                        mi, d, R, hm.ctypes.data_as(ctypes.c_void_p), posz.ctypes.data_as(ctypes.c_void_p),
                        mask.ctypes.data_as(ctypes.c_void_p), Ts.ctypes.data_as(ctypes.c_void_p),
                        offs.ctypes.data_as(ctypes.c_void_p), ws.ctypes.data_as(ctypes.c_void_p),
                        hs.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(cfg.resolutionAct))
                    want = e.heuristic_action(m, d)
                    assert (got >> 8, (got >> 4) & 15, got & 15) == want, (t, m, d)
            r_, x_, y_ = e.heuristic_action(HEURISTICS[t % 4], t % 4)
            acts.append((r_ * 16 + x_) * 16 + y_)
        env.step(acts, poses=True)


def test_device_border_follower_matches_oracle_contours(contour_harness):
    """cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE) + find_out_contour (cvTools.py:7-38,86): the point
    LISTS (order within a contour matters for approxPolyDP) of the device follower, compiled for the host,
    against the oracle's restatement (itself fuzzed against cv2 in test_oracle_contours.py)."""
    from oracle import contours_port as cp
    rng = np.random.default_rng(3)
    out = np.zeros(4096, np.int32)
    for it in range(800):
        kind = it % 4
        if kind == 0:
            img = (rng.random((16, 16)) < rng.uniform(0.05, 0.95)).astype(np.uint8)
        elif kind == 1:
            img = np.kron((rng.random((8, 8)) < rng.uniform(0.2, 0.9)).astype(np.uint8), np.ones((2, 2), np.uint8))
        elif kind == 2:
            img = np.ones((16, 16), np.uint8)
            for _ in range(int(rng.integers(1, 9))):
                x0, y0 = rng.integers(0, 15, 2); w, h = rng.integers(1, 5, 2)
                img[x0:x0 + w, y0:y0 + h] = 0
        else:
            img = np.zeros((16, 16), np.uint8)
            a = int(rng.integers(0, 3)); img[a:16 - a, a:16 - a] = 1
            b = a + int(rng.integers(1, 3)); img[b:16 - b, b:16 - b] = 0
            c = b + int(rng.integers(1, 3)); img[c:16 - c, c:16 - c] = 1
            img ^= (rng.random((16, 16)) < 0.05).astype(np.uint8)
        rows = np.zeros(16, np.uint16)
        for y in range(16):
            rows[y] = sum(1 << x for x in range(16) if img[y, x])
        nw = contour_harness.outer_contours_host(rows.ctypes.data_as(ctypes.c_void_p),
                                                 out.ctypes.data_as(ctypes.c_void_p), len(out))
        assert nw >= 0
        got, i = [], 0
        while i < nw:
            n = int(out[i]); pts = out[i + 1:i + 1 + n]; i += 1 + n
            got.append(tuple((int(p) >> 4, int(p) & 15) for p in pts))
        want = [tuple((int(p[0]), int(p[1])) for p in c_) for c_ in cp.find_outer_contours(img)]
        assert sorted(got) == sorted(want), it



"""Fuzz the cv2-free restatement (oracle/contours_port.py) against the cv2 build in the image --
the library the reference calls at cvTools.py:86,91.  Skipped where cv2 is not importable."""
import numpy as np
import pytest

from oracle import contours_port

cv2 = pytest.importorskip("cv2")


def _outer(img):
    from oracle.oracle_env import _outer_contours_cv2
    return [[(int(p[0]), int(p[1])) for p in c.reshape(-1, 2)] for c in _outer_contours_cv2((img * 255).astype(np.uint8))]


def _rand_img(rng):
    kind = int(rng.integers(0, 3))
    if kind == 0:
        return (rng.random((16, 16)) < rng.uniform(0.1, 0.9)).astype(np.uint8)
    if kind == 1:
        img = np.zeros((16, 16), np.uint8)
        for _ in range(int(rng.integers(1, 6))):
            x0, y0 = rng.integers(0, 14, 2); w, h = rng.integers(1, 9, 2)
            img[x0:x0 + w, y0:y0 + h] = rng.integers(0, 2)
        return img
    img = np.ones((16, 16), np.uint8)
    for _ in range(int(rng.integers(1, 8))):
        x0, y0 = rng.integers(0, 15, 2); w, h = rng.integers(1, 5, 2)
        img[x0:x0 + w, y0:y0 + h] = 0
    return img


def test_outer_contours_match_cv2():
    rng = np.random.default_rng(0)
    for _ in range(600):
        img = _rand_img(rng)
        assert sorted(map(tuple, contours_port.find_outer_contours(img))) == sorted(map(tuple, _outer(img)))


def test_approx_poly_matches_cv2_on_contours():
    rng = np.random.default_rng(1)
    n = 0
    for _ in range(400):
        img = _rand_img(rng)
        cs, _ = cv2.findContours((img * 255).astype(np.uint8), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        for c in cs:
            want = [tuple(p) for p in cv2.approxPolyDP(c, 1, True).reshape(-1, 2).tolist()]
            got = contours_port.approx_poly_dp_closed([tuple(p) for p in c.reshape(-1, 2).tolist()], 1.0)
            assert got == want
            n += 1
    assert n > 1000


def test_approx_poly_matches_cv2_on_arbitrary_rings():
    rng = np.random.default_rng(2)
    for _ in range(20000):
        k = int(rng.integers(1, 14))
        pts = rng.integers(0, 16, size=(k, 2)).astype(np.int32)
        want = [tuple(p) for p in cv2.approxPolyDP(pts.reshape(-1, 1, 2), 1, True).reshape(-1, 2).tolist()]
        got = contours_port.approx_poly_dp_closed([tuple(int(v) for v in p) for p in pts], 1.0)
        assert got == want

"""CPU restatement of the third-party arithmetic on the candidate-extraction path.

TEST INFRASTRUCTURE ONLY (oracle): the product path never imports this.

The reference only *calls* OpenCV here -- ``cv2.findContours(img, RETR_TREE,
CHAIN_APPROX_SIMPLE)`` at reference ``cvTools.py:86`` and
``cv2.approxPolyDP(c, 1, True)`` at ``cvTools.py:91``; the pinned dependency is
``opencv-contrib-python==4.4.0.46`` (reference ``requirements.txt:13``), absent
from ``/root/reference``.  What is restated below is the published algorithm
(Suzuki-Abe border following; Douglas-Peucker with OpenCV's closed-curve seed
and clean-up pass) as pinned against the cv2 build in this image (4.13.0) by
``tests/test_oracle_contours.py`` (random fuzz, run where cv2 is importable) and
by the committed fixtures under ``tests/golden/`` (generated from the unmodified
reference + cv2 by ``tests/golden/make_golden.py``).  Parity with the pinned
4.4.0.46 build itself is UNPINNED (cannot be executed offline).

Also here: NumPy float64 ``floor_divide`` semantics used at ``cvTools.py:78``.
"""
import math

import numpy as np

# direction codes, y grows downward: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE
_DX = (1, 1, 0, -1, -1, -1, 0, 1)
_DY = (0, -1, -1, -1, 0, 1, 1, 1)


def npy_floor_divide(a, b):
    """float64 ``a // b`` exactly as NumPy computes it (npy_divmod)."""
    mod = math.fmod(a, b)
    if b == 0.0:
        return a / b
    div = (a - mod) / b
    if mod != 0.0:
        if (b < 0) != (mod < 0):
            div -= 1.0
    if div != 0.0:
        fl = math.floor(div)
        if div - fl > 0.5:
            fl += 1.0
    else:
        fl = math.copysign(0.0, a / b)
    return fl


def find_outer_contours(img):
    """Outer borders of a 0/1 image, each as the point list (x=col, y=row) that
    ``cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE)`` returns for the contours kept by the
    reference's ``find_out_contour`` (``cvTools.py:7-38``: even hierarchy depth == outer borders).
    Hole borders are followed too (their marks steer later start detection) but not returned."""
    rows, cols = img.shape
    f = np.zeros((rows + 2, cols + 2), dtype=np.int32)
    f[1:-1, 1:-1] = (np.asarray(img) != 0)
    nbd = 1
    out = []
    for i in range(1, rows + 1):
        for j in range(1, cols + 1):
            v = f[i, j]
            if v == 0:
                continue
            if v == 1 and f[i, j - 1] == 0:
                hole = False
            elif v >= 1 and f[i, j + 1] == 0:
                hole = True
            else:
                continue
            nbd += 1
            pts = _follow_border(f, i, j, hole, nbd)
            if not hole:
                out.append([(x - 1, y - 1) for (x, y) in pts])
    return out


def _follow_border(f, i0, j0, hole, nbd):
    s_end = s = 0 if hole else 4
    while True:
        s = (s - 1) & 7
        if f[i0 + _DY[s], j0 + _DX[s]] != 0:
            break
        if s == s_end:
            break
    if f[i0 + _DY[s], j0 + _DX[s]] == 0:  # isolated pixel
        f[i0, j0] = -nbd
        return [(j0, i0)]
    pts = []
    i1, j1 = i0 + _DY[s], j0 + _DX[s]
    i3, j3 = i0, j0
    prev_s = s ^ 4
    while True:
        s_end = s
        while True:
            s += 1
            i4, j4 = i3 + _DY[s & 7], j3 + _DX[s & 7]
            if f[i4, j4] != 0:
                break
        s &= 7
        if ((s - 1) & 0xFFFFFFFF) < (s_end & 0xFFFFFFFF):
            f[i3, j3] = -nbd
        elif f[i3, j3] == 1:
            f[i3, j3] = nbd
        if s != prev_s:
            pts.append((j3, i3))
        prev_s = s
        if i4 == i0 and j4 == j0 and i3 == i1 and j3 == j1:
            break
        i3, j3 = i4, j4
        s = (s + 4) & 7
    return pts


def approx_poly_dp_closed(pts, eps=1.0, legacy_line_distance=False):
    """``cv2.approxPolyDP(contour, eps, closed=True)`` for integer points (cv2 4.13.0 behaviour:
    point-to-SEGMENT distance in the recursion).  ``legacy_line_distance=True`` selects the older
    point-to-infinite-line rule believed to be what 4.4.0.46 does (unverifiable offline)."""
    n = len(pts)
    if n == 0:
        return []
    P = pts
    eps2 = eps * eps
    if n == 1:
        return [P[0]]

    # 1. seed: find two far-apart points
    pos = 0
    far = 0
    le_eps = False
    for _ in range(3):
        pos = (pos + far) % n
        sx, sy = P[pos]
        maxd = 0
        far = 0
        for j in range(1, n):
            x, y = P[(pos + j) % n]
            d = (x - sx) * (x - sx) + (y - sy) * (y - sy)
            if d > maxd:
                maxd = d
                far = j
        le_eps = maxd <= eps2
    if le_eps:
        Q = [P[pos]]
        return Q
    stack = []
    right = (pos, (far + pos) % n)
    left = ((far + pos) % n, pos)
    stack.append(left)
    stack.append(right)

    # 2. Douglas-Peucker on the two arcs
    Q = []
    while stack:
        s, e = stack.pop()
        sx, sy = P[s]
        ex, ey = P[e]
        if (s + 1) % n == e:
            Q.append(P[s])
            continue
        dx = ex - sx
        dy = ey - sy
        maxd = 0.0
        mi = -1
        k = (s + 1) % n
        first = True
        while k != e:
            px, py = P[k]
            if legacy_line_distance:
                d = abs((py - sy) * dx - (px - sx) * dy)
                dcmp = float(d)
            else:
                # squared distance from (px,py) to the segment s-e.  All inputs are integers, so
                # every product below is exact in float64; the only rounding is the one division.
                vx, vy = px - sx, py - sy
                seg2 = dx * dx + dy * dy
                dot = vx * dx + vy * dy
                if seg2 == 0 or dot <= 0:
                    dcmp = float(vx * vx + vy * vy)
                elif dot >= seg2:
                    wx, wy = px - ex, py - ey
                    dcmp = float(wx * wx + wy * wy)
                else:
                    cross = vy * dx - vx * dy
                    dcmp = float(cross * cross) / float(seg2)
            if first or dcmp > maxd:
                if first:
                    maxd = dcmp
                    mi = k
                    first = False
                elif dcmp > maxd:
                    maxd = dcmp
                    mi = k
            k = (k + 1) % n
        if legacy_line_distance:
            le = maxd * maxd <= eps2 * float(dx * dx + dy * dy)
        else:
            le = maxd <= eps2
        if le:
            Q.append(P[s])
        else:
            stack.append((mi, e))
            stack.append((s, mi))

    # 3. clean-up of almost-collinear points on the closed ring
    c = len(Q)
    new_count = c
    if c <= 2:
        return Q
    src = list(Q)
    dst = list(Q)
    rpos = c - 1
    start = src[rpos]; rpos = (rpos + 1) % c
    wpos = rpos
    pt = src[rpos]; rpos = (rpos + 1) % c
    i = 0
    while i < c and new_count > 2:
        end = src[rpos]; rpos = (rpos + 1) % c
        dx = end[0] - start[0]
        dy = end[1] - start[1]
        dist = abs((pt[0] - start[0]) * dy - (pt[1] - start[1]) * dx)
        ip = (pt[0] - start[0]) * (end[0] - pt[0]) + (pt[1] - start[1]) * (end[1] - pt[1])
        if float(dist) * float(dist) <= 0.5 * eps2 * float(dx * dx + dy * dy) and dx != 0 and dy != 0 and ip >= 0:
            new_count -= 1
            start = end
            dst[wpos] = end; wpos = (wpos + 1) % c
            pt = src[rpos]; rpos = (rpos + 1) % c
            i += 2
            continue
        start = pt
        dst[wpos] = pt; wpos = (wpos + 1) % c
        pt = end
        i += 1
    return dst[:new_count]


def convex_vertices(poly):
    """reference ``find_convex_vetex`` (``cvTools.py:40-59``): all points if <= 3, else those with
    cross(B - A, C - A) < 0 for consecutive A, B, C on the ring."""
    n = len(poly)
    if n <= 3:
        return list(poly)
    keep = []
    for k in range(n):
        ax, ay = poly[k - 1]
        bx, by = poly[k]
        cx, cy = poly[(k + 1) % n]
        cross = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)
        if cross < 0:
            keep.append(poly[k])
    return keep


def convex_hulls_port(posZMap, mask, heightResolution=0.01, legacy_line_distance=False):
    """Restatement of reference ``cvTools.convexHulls`` (``cvTools.py:77-103``) without cv2.
    Returns (K x 2 int array of (col, row) sorted by col then row, V[K]) or ([], None)."""
    rows, cols = posZMap.shape
    mapInt = np.zeros((rows, cols), dtype=np.int64)
    for r in range(rows):
        for c in range(cols):
            mapInt[r, c] = int(np.int32(npy_floor_divide(float(posZMap[r, c]), float(heightResolution))))
    mapInt[mask == 0] = -1
    found = set()
    for h in np.unique(mapInt):
        if h == -1:
            continue
        img = (mapInt == h)
        for contour in find_outer_contours(img):
            approx = approx_poly_dp_closed(contour, 1.0, legacy_line_distance)
            for p in convex_vertices(approx):
                found.add((int(p[0]), int(p[1])))
    if not found:
        return [], None
    allc = np.array(sorted(found), dtype=np.int64)
    V = mask[(allc[:, 1], allc[:, 0])]
    return allc, V

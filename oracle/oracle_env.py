"""CPU oracle of the IR-BPP packing-environment hot path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module; the product (``irbpp_b200/``) never does and has no CPU fallback.

What is restated, and where it comes from in the reference
(paths relative to ``/root/reference``):

* drop-height / feasibility scan  ``environment/physics0/space.py:98-129``
* candidate extraction            ``environment/physics0/cvTools.py:61-103`` (calls cv2; see
  ``contours_port.py`` for the cv2-free restatement of findContours / approxPolyDP)
* candidate select / pad, observation assembly   ``environment/physics0/binPhy.py:183-232``
* action decode + prejudge        ``binPhy.py:234-245``
* placement outcome with ``simulation=False`` semantics   ``binPhy.py:266-289`` +
  ``environment/physics0/Interface.py:365-369`` (AABB top <= bin top)
* analytic heightmap update       ``space.py:75-94`` in the closed form the reference itself uses
  at ``space.py:213``:  ``hm[win] = max(hm[win], (T_r + z) * mT_r)``
* reward / done / info            ``binPhy.py:299-327``, ``:149-156``
* item queue                      ``environment/physics0/IRcreator.py:6-24``
* auto-reset                      ``wrapper/shmem_vec_env.py:140-144``
* episode info                    ``wrapper/monitor.py:58-75``

Pinning: ``tests/golden/make_golden.py`` drives the UNMODIFIED reference
``Space`` / ``cvTools`` (via ``ref_loader.py``) on seeded inputs and commits the
outputs; ``tests/test_oracle_*.py`` check this restatement against them.  The
PyBullet settle (``Interface.simulateToQuasistatic``) is out of scope
(parity unpinned for it).

Two intentional, documented choices where the reference is not a function of its
inputs: (1) ``np.argsort`` ties in the >selectedAction truncation and in the
no-candidate fallback (``binPhy.py:211,219``) are broken by lowest index
(``kind='stable'``); (2) item ids come from an explicit per-env sequence (cursor
advances on every ``generate_item``; wraps modulo its length) instead of the
global NumPy RNG.
"""
import numpy as np

from . import contours_port

try:  # cv2 is what the reference itself calls; present in this image
    import cv2  # noqa: F401
    HAVE_CV2 = True
except Exception:  # pragma: no cover
    HAVE_CV2 = False


class OracleConfig(object):
    """Hot-path relevant settings and their reference defaults (``arguments.py:11-96,115``)."""

    def __init__(self, bin_dimension=(0.32, 0.32, 0.30), resolutionAct=0.02, resolutionH=0.01,
                 resolutionZ=0.01, ZRotNum=4, selectedAction=500, bufferSize=1):
        self.bin_dimension = np.array(bin_dimension, dtype=np.float64)
        self.resolutionAct = resolutionAct
        self.resolutionH = resolutionH
        self.resolutionZ = resolutionZ
        self.ZRotNum = int(ZRotNum)
        self.selectedAction = int(selectedAction)
        self.bufferSize = int(bufferSize)
        self.stepSize = int(resolutionAct / resolutionH)
        assert self.stepSize == resolutionAct / resolutionH  # space.py:19-20
        self.rangeX_C, self.rangeY_C = [int(v) for v in np.ceil(self.bin_dimension[0:2] / resolutionH).astype(np.int32)]
        self.rangeX_A, self.rangeY_A = [int(v) for v in np.ceil(self.bin_dimension[0:2] / resolutionAct).astype(np.int32)]

    @property
    def loc_obs_len(self):  # binPhy.py:87-98 with heightMapPre
        return self.selectedAction * 5 + 9 + self.rangeX_C * self.rangeY_C

    @property
    def order_obs_len(self):
        return self.bufferSize + self.rangeX_C * self.rangeY_C

    @property
    def obs_len(self):
        return self.order_obs_len if self.bufferSize > 1 else self.loc_obs_len

    @property
    def act_len(self):
        return self.bufferSize if self.bufferSize > 1 else self.selectedAction


# ---------------------------------------------------------------------------
# a3: drop-height / feasibility scan
# ---------------------------------------------------------------------------

def scan_loops(cfg, heightmap, extents_r, tables_r):
    """``Space.get_possible_position`` restated with the reference's own loop structure
    (rot x X x Y, NumPy window max) -- this is the variant timed as the CPU baseline.
    Returns (posZmap, posZValid, naiveMask), each float64 [R, Ax, Ay]."""
    R = len(tables_r)
    Ax, Ay = cfg.rangeX_A, cfg.rangeY_A
    naiveMask = np.zeros((R, Ax, Ay))
    posZmap = np.full((R, Ax, Ay), 1e3)
    binz = cfg.bin_dimension[2]
    for r in range(R):
        boundingSize = np.round(extents_r[r], decimals=6)
        wH, hH = np.ceil(boundingSize[0:2] / cfg.resolutionH).astype(np.int32)
        wA, hA = np.ceil(boundingSize[0:2] / cfg.resolutionAct).astype(np.int32)
        _, B, _, mB = tables_r[r]
        for X in range(Ax - wA + 1):
            for Y in range(Ay - hA + 1):
                cx, cy = X * cfg.stepSize, Y * cfg.stepSize
                posZ = np.max((heightmap[cx:cx + wH, cy:cy + hH] - B) * mB)
                if np.round(posZ + boundingSize[2] - binz, decimals=6) <= 0:
                    naiveMask[r, X, Y] = 1
                posZmap[r, X, Y] = posZ
    posZValid = posZmap.copy()
    posZValid[naiveMask == 0] = 1e3
    return posZmap, posZValid, naiveMask


def scan_vectorized(cfg, heightmap, extents_r, tables_r):
    """Same result as ``scan_loops`` (value-equal; checked by tests), one strided view per rotation."""
    R = len(tables_r)
    Ax, Ay = cfg.rangeX_A, cfg.rangeY_A
    naiveMask = np.zeros((R, Ax, Ay))
    posZmap = np.full((R, Ax, Ay), 1e3)
    binz = cfg.bin_dimension[2]
    st = cfg.stepSize
    for r in range(R):
        boundingSize = np.round(extents_r[r], decimals=6)
        wH, hH = [int(v) for v in np.ceil(boundingSize[0:2] / cfg.resolutionH).astype(np.int32)]
        wA, hA = [int(v) for v in np.ceil(boundingSize[0:2] / cfg.resolutionAct).astype(np.int32)]
        nX, nY = Ax - wA + 1, Ay - hA + 1
        if nX <= 0 or nY <= 0:
            continue
        _, B, _, mB = tables_r[r]
        win = np.lib.stride_tricks.sliding_window_view(heightmap, (wH, hH))[::st, ::st][:nX, :nY]
        posZ = ((win - B) * mB).max(axis=(2, 3))
        posZmap[r, :nX, :nY] = posZ
        ok = np.round(posZ + boundingSize[2] - binz, decimals=6) <= 0
        naiveMask[r, :nX, :nY] = ok
    posZValid = posZmap.copy()
    posZValid[naiveMask == 0] = 1e3
    return posZmap, posZValid, naiveMask


# ---------------------------------------------------------------------------
# a5 / a6: candidate extraction
# ---------------------------------------------------------------------------

def _outer_contours_cv2(check):
    """Contours of even hierarchy depth (outer borders), what reference ``find_out_contour``
    (``cvTools.py:7-38``) keeps from ``cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE)``."""
    import cv2
    contours, hierarchy = cv2.findContours(image=check, mode=cv2.RETR_TREE, method=cv2.CHAIN_APPROX_SIMPLE)
    if len(contours) == 0:
        return []
    parent = hierarchy[0][:, 3]
    keep = []
    for i in range(len(contours)):
        depth = 0
        p = parent[i]
        while p != -1:
            depth += 1
            p = parent[p]
        if depth % 2 == 0:
            keep.append(contours[i])
    return keep


def convex_hulls_cv2(posZMap, mask, heightResolution):
    """``cvTools.convexHulls`` (``cvTools.py:77-103``) restated on top of the same cv2 calls."""
    import cv2
    mapInt = (posZMap // heightResolution).astype(np.int32)
    mapInt[mask == 0] = -1
    pts = []
    for h in np.unique(mapInt):
        if h == -1:
            continue
        check = np.where(mapInt == h, 255, 0).astype(np.uint8)
        for c in _outer_contours_cv2(check):
            approx = cv2.approxPolyDP(c, 1, True).reshape(-1, 2)
            poly = [(int(p[0]), int(p[1])) for p in approx]
            pts.extend(contours_port.convex_vertices(poly))
    if not pts:
        return [], None
    allc = np.unique(np.array(pts, dtype=np.int64), axis=0)
    V = mask[(allc[:, 1], allc[:, 0])]
    return allc, V


def convex_hull_actions(posZValid, mask, heightResolution, backend="port"):
    """``cvTools.getConvexHullActions`` (``cvTools.py:61-75``): rows ``[rot, row, col, H, V]``
    concatenated in rotation order, or None."""
    fn = convex_hulls_cv2 if backend == "cv2" else contours_port.convex_hulls_port
    rows = []
    for r in range(len(posZValid)):
        hulls, V = fn(posZValid[r], mask[r], heightResolution)
        if len(hulls) != 0:
            H = posZValid[r][hulls[:, 1], hulls[:, 0]]
            k = len(hulls)
            rows.append(np.stack([np.full(k, float(r)), hulls[:, 1].astype(np.float64),
                                  hulls[:, 0].astype(np.float64), H, V.astype(np.float64)], axis=1))
    if not rows:
        return None
    return np.concatenate(rows, axis=0)


def select_candidates(cfg, candidates, posZValid, naiveMask):
    """``binPhy.py:205-225``: truncate by height (stable), zero-pad, or the no-candidate fallback."""
    sel = cfg.selectedAction
    if candidates is not None:
        if len(candidates) > sel:
            idx = np.argsort(candidates[:, 3], kind="stable")[0:sel]
            candidates = candidates[idx]
        elif len(candidates) < sel:
            candidates = np.concatenate((candidates, np.zeros((sel - len(candidates), 5))), axis=0)
        return candidates
    flat = posZValid.reshape(-1)
    idx = np.argsort(flat, kind="stable")[0:sel]
    ROT, X, Y = np.unravel_index(idx, posZValid.shape)
    H = np.full(len(idx), cfg.bin_dimension[-1])
    V = naiveMask.reshape(-1)[idx]
    return np.stack([ROT.astype(np.float64), X.astype(np.float64), Y.astype(np.float64), H, V], axis=1)


# ---------------------------------------------------------------------------
# placement heuristics (space.py:162-227)
# ---------------------------------------------------------------------------

HEURISTICS = ("MINZ", "DBLF", "FIRSTFIT", "HM")


def heuristic_action_port(cfg, lib, heightmap, item_id, posZmap, naiveMask, method, dirIdx):
    """Restatement of ``Space.get_heuristic_action`` (space.py:162-227): score every pose, 1e6 where
    ``naiveMask == 0``, ``np.round(score, 6)``, ``np.argmin`` (first minimum in flat (rot, lx, ly)
    order).  The RANDOM branch of the reference (space.py:220-226) raises on every call
    (``np.random.choice`` of a tuple) and is not restated.  ``np.sum`` below is NumPy's own pairwise
    reduction, i.e. the very arithmetic the reference executes."""
    assert 0 <= dirIdx <= 3                                           # space.py:167
    Xflip, Yflip = dirIdx >= 2, dirIdx % 2 == 1                       # space.py:163-166
    R, AX, AY = naiveMask.shape
    X = np.arange(AX, dtype=np.float64)[:, None].repeat(AY, axis=1)   # coors[:, :, 0]  (space.py:43-47)
    Y = np.arange(AY, dtype=np.float64)[None, :].repeat(AX, axis=0)
    coorsX = AX - X if Xflip else X
    coorsY = AY - Y if Yflip else Y
    invalid = naiveMask == 0
    if method == "MINZ":
        score = posZmap.copy()
    elif method == "DBLF":
        score = np.broadcast_to(coorsX + coorsY, naiveMask.shape) * cfg.resolutionAct + 100 * posZmap
    elif method == "FIRSTFIT":
        score = np.broadcast_to(coorsX + coorsY, naiveMask.shape).copy()
    elif method == "HM":
        score = np.broadcast_to((coorsX + coorsY) * cfg.resolutionAct, naiveMask.shape).copy()
        score[invalid] = 1e6
        for r in range(R):
            T, _, mT, _ = lib.tables[item_id][r]
            w, h = T.shape
            for cx in range(AX):
                for cy in range(AY):
                    if naiveMask[r, cx, cy] == 0:
                        continue
                    z = posZmap[r, cx, cy]
                    x0, y0 = cx * cfg.stepSize, cy * cfg.stepSize
                    prime = np.max(((T + z) * mT, heightmap[x0:x0 + w, y0:y0 + h]), axis=0)
                    score[r, cx, cy] += np.sum(prime) * 100
    else:
        raise ValueError(method)
    score = np.array(score, dtype=np.float64)
    score[invalid] = 1e6
    score = np.round(score, decimals=6)
    rot, lx, ly = np.unravel_index(int(np.argmin(score)), score.shape)
    return int(rot), int(lx), int(ly)


# ---------------------------------------------------------------------------
# the environment (binPhy.PackingGame with simulation=False semantics)
# ---------------------------------------------------------------------------

class RefGeometry(object):
    """Geometry backend running the UNMODIFIED reference ``Space`` / ``cvTools`` (build container
    only).  Used by ``tests/golden/make_golden.py``."""

    def __init__(self, cfg, lib):
        from . import ref_loader
        _, space_mod, cv_mod = ref_loader.load_reference()
        self._cv = cv_mod
        self._mesh = ref_loader.MeshStandIn
        self.space = space_mod.Space(cfg.bin_dimension, cfg.resolutionAct, cfg.resolutionH, False,
                                     cfg.ZRotNum, lib.shot_info(), [100, 100, 100])
        self.lib = lib
        self.cfg = cfg

    def scan(self, heightmap, item_id):
        self.space.heightmapC = heightmap  # same array object the env updates
        meshes = [self._mesh(self.lib.extents[item_id, r]) for r in range(self.cfg.ZRotNum)]
        mask = self.space.get_possible_position(item_id, meshes, self.cfg.selectedAction)
        return self.space.posZmap.copy(), self.space.posZValid.copy(), mask.copy()

    def hull_actions(self, posZValid, naiveMask):
        return self._cv.getConvexHullActions(posZValid, naiveMask, self.cfg.resolutionZ)

    def heuristic_action(self, heightmap, item_id, posZmap, naiveMask, method, dirIdx):
        """Verbatim ``Space.get_heuristic_action`` (space.py:162-227) on the state of the last scan."""
        sp = self.space
        sp.heightmapC = heightmap
        sp.posZmap[:] = posZmap
        sp.naiveMask = naiveMask.copy()
        meshes = [self._mesh(self.lib.extents[item_id, r]) for r in range(self.cfg.ZRotNum)]
        rot, lx, ly = sp.get_heuristic_action(dirIdx, method, item_id, meshes)
        return int(rot), int(lx), int(ly)


class PortGeometry(object):
    def __init__(self, cfg, lib, scan="vectorized", contours="port"):
        self.cfg = cfg
        self.lib = lib
        self._scan = scan_loops if scan == "loops" else scan_vectorized
        self._contours = contours

    def scan(self, heightmap, item_id):
        return self._scan(self.cfg, heightmap, self.lib.extents[item_id], self.lib.tables[item_id])

    def hull_actions(self, posZValid, naiveMask):
        return convex_hull_actions(posZValid, naiveMask, self.cfg.resolutionZ, self._contours)

    def heuristic_action(self, heightmap, item_id, posZmap, naiveMask, method, dirIdx):
        return heuristic_action_port(self.cfg, self.lib, heightmap, item_id, posZmap, naiveMask, method, dirIdx)


class OracleEnv(object):
    """One bin.  Mirrors the public surface the reference's callers use: ``reset``, ``step``,
    ``get_action_candidates``, ``get_all_possible_observation``, ``get_ratio``."""

    def __init__(self, cfg, lib, sequence, geometry=None):
        self.cfg = cfg
        self.lib = lib
        self.sequence = np.asarray(sequence, dtype=np.int64)
        self.cursor = 0
        self.geo = geometry if geometry is not None else PortGeometry(cfg, lib)
        self.heightmap = np.zeros((cfg.rangeX_C, cfg.rangeY_C))
        self.item_list = []
        self.chooseItem = cfg.bufferSize > 1
        self.orderAction = 0
        self.next_item_vec = np.zeros(9)
        self.packed_ids = []
        self.candidates = None
        self.next_item_ID = None
        self.next_k_item_ID = None
        self.posZmap = self.posZValid = self.naiveMask = None
        self.binvol = np.prod(cfg.bin_dimension)

    # --- item queue (IRcreator.py:6-24) ---
    def _generate_item(self):
        self.item_list.append(int(self.sequence[self.cursor % len(self.sequence)]))
        self.cursor += 1

    def _preview(self, n):
        while len(self.item_list) < n:
            self._generate_item()
        return list(self.item_list[:n])

    # --- binPhy.py:128-147 ---
    def reset(self):
        self.heightmap[:] = 0
        self.item_list.clear()
        self.packed_ids = []
        self.next_item_vec[:] = 0
        return self.cur_observation()

    def get_ratio(self):  # binPhy.py:149-153
        total = 0
        for i in self.packed_ids:
            total += self.lib.volume[i]
        return total / self.binvol

    def _scan(self, item_id):
        self.posZmap, self.posZValid, self.naiveMask = self.geo.scan(self.heightmap, item_id)

    def get_action_candidates(self, orderAction):  # binPhy.py:161-169
        self.next_item_ID = self.next_k_item_ID[orderAction]
        self.chooseItem = False
        obs = self.cur_observation(genItem=False)
        self.chooseItem = True
        self.orderAction = orderAction
        return obs

    def get_all_possible_observation(self):  # binPhy.py:171-180
        self.chooseItem = False
        out = []
        for item in self.next_k_item_ID:
            self.next_item_ID = item
            out.append(self.cur_observation(genItem=False))
        return np.concatenate(out, axis=0)

    def cur_observation(self, genItem=True):  # binPhy.py:183-232
        cfg = self.cfg
        if not self.chooseItem:
            if genItem:
                self.next_item_ID = self._preview(1)[0]
            self.next_item_vec[0] = self.next_item_ID
            self._scan(self.next_item_ID)
            cand = self.geo.hull_actions(self.posZValid, self.naiveMask)
            self.candidates = select_candidates(cfg, cand, self.posZValid, self.naiveMask)
            return np.concatenate((self.candidates.reshape(-1), self.next_item_vec.reshape(-1),
                                   self.heightmap.reshape(-1)))
        self.next_k_item_ID = self._preview(cfg.bufferSize)
        return np.concatenate((np.array(self.next_k_item_ID, dtype=np.float64), self.heightmap.reshape(-1)))

    def heuristic_action(self, method, dirIdx=0):  # space.py:162-227 on the current scan
        return self.geo.heuristic_action(self.heightmap, self.next_item_ID, self.posZmap, self.naiveMask,
                                         method, dirIdx)

    def candidate_index(self, pose):
        """First row of the candidate table holding ``pose`` (rot, lx, ly), -1 if absent."""
        hit = np.nonzero((self.candidates[:, 0:3] == np.asarray(pose, dtype=np.float64)).all(axis=1))[0]
        return int(hit[0]) if len(hit) else -1

    def step(self, action):  # binPhy.py:248-337 with simulation=False
        rotIdx, lx, ly = [int(v) for v in self.candidates[int(action)][0:3]]   # action_to_position, :234-236
        return self.step_pose(rotIdx, lx, ly)

    def step_pose(self, rotIdx, lx, ly):  # PackingGame.step after action_to_position
        cfg = self.cfg
        targetFLB = np.round((lx * cfg.resolutionAct, ly * cfg.resolutionAct, cfg.bin_dimension[2]), decimals=6)
        item = self.next_item_ID
        extents = self.lib.extents[item, rotIdx]
        # prejudge (binPhy.py:238-245)
        success = True
        if np.round(targetFLB[0] + extents[0] - cfg.bin_dimension[0], decimals=6) > 0 \
                or np.round(targetFLB[1] + extents[1] - cfg.bin_dimension[1], decimals=6) > 0:
            success = False
        if np.sum(self.naiveMask) == 0:
            success = False
        height = self.posZmap[rotIdx, lx, ly]
        if success:
            # Interface.simulateHeight (Interface.py:365-369): AABB top = height + extent_z
            ez = np.round(extents, decimals=6)[2]
            if np.round(height + ez - cfg.bin_dimension[2], decimals=6) > 0:
                success = False
        if not success:
            info = {"counter": len(self.packed_ids), "ratio": self.get_ratio(), "Valid": True}
            obs = self.cur_observation()
            return obs, 0.0, True, info
        # analytic heightmap update (space.py:213 closed form)
        T, _, mT, _ = self.lib.tables[item][rotIdx]
        w, h = T.shape
        cx, cy = lx * cfg.stepSize, ly * cfg.stepSize
        win = self.heightmap[cx:cx + w, cy:cy + h]
        self.heightmap[cx:cx + w, cy:cy + h] = np.maximum(win, (T + height) * mT)
        self.packed_ids.append(item)
        item_ratio = self.lib.volume[item] / self.binvol
        reward = item_ratio * 10
        self.item_list.pop(self.orderAction)
        self._generate_item()
        obs = self.cur_observation()
        return obs, reward, False, {"Valid": True}


class OracleVecEnv(object):
    """In-process vector of oracle envs with the auto-reset and episode-info semantics of the
    reference's workers (``wrapper/shmem_vec_env.py:140-144``, ``wrapper/dummy_vec_env.py:45-55``,
    ``wrapper/monitor.py:58-75``).  Observations are float64 as the reference's envs return them;
    ``VecPyTorch`` (``envs.py:149-165``) is what casts to float32."""

    def __init__(self, cfg, lib, sequences, geometry_factory=None):
        self.cfg = cfg
        self.num_envs = len(sequences)
        self.envs = [OracleEnv(cfg, lib, sequences[i],
                               geometry_factory(cfg, lib) if geometry_factory else None)
                     for i in range(self.num_envs)]
        self._ep_rewards = [[] for _ in range(self.num_envs)]

    def reset(self):
        self._ep_rewards = [[] for _ in range(self.num_envs)]
        return np.stack([e.reset() for e in self.envs])

    def get_action_candidates(self, order_actions):
        return [e.get_action_candidates(int(a)) for e, a in zip(self.envs, order_actions)]

    def heuristic_actions(self, method, dirIdx=0):
        poses = np.array([e.heuristic_action(method, dirIdx) for e in self.envs], dtype=np.int32)
        index = np.array([e.candidate_index(p) for e, p in zip(self.envs, poses)], dtype=np.int64)
        return poses, index

    def step(self, actions, poses=False):
        obs, rews, dones, infos = [], [], [], []
        for i, e in enumerate(self.envs):
            if poses:
                a = int(actions[i])
                o, r, d, info = e.step_pose(a // (self.cfg.rangeX_A * self.cfg.rangeY_A),
                                            (a // self.cfg.rangeY_A) % self.cfg.rangeX_A, a % self.cfg.rangeY_A)
            else:
                o, r, d, info = e.step(int(actions[i]))
            self._ep_rewards[i].append(r)
            if d:
                info["episode"] = {"r": round(sum(self._ep_rewards[i]), 6), "l": len(self._ep_rewards[i])}
                self._ep_rewards[i] = []
                o = e.reset()
            obs.append(o); rews.append(r); dones.append(d); infos.append(info)
        return np.stack(obs), np.array(rews), np.array(dones), infos


def lowest_valid_action(obs_row, selectedAction):
    """Deterministic test policy: the lowest-H candidate with V == 1 (first index on ties); action 0
    if none is valid.  Works on float32 or float64 observation rows."""
    cand = np.asarray(obs_row[:selectedAction * 5]).reshape(selectedAction, 5)
    valid = cand[:, 4] == 1
    if not valid.any():
        return 0
    h = np.where(valid, cand[:, 3], np.inf)
    return int(np.argmin(h))

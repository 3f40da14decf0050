"""Load the UNMODIFIED reference geometry modules for golden-vector generation.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``irbpp_b200/``) may
import this module.  It only works where ``/root/reference`` is mounted (the
build container); the GPU box has no reference tree, which is why the vectors
it produces are committed under ``tests/golden/``.

The reference cannot be imported as a whole (``tools.py:5-15`` pulls in
trimesh / gym / matplotlib / transforms3d / pybullet, none installed).  The hot
functions do run verbatim once empty stand-ins for those imports are registered
(SURVEY.md section 8c):

* ``environment/physics0/space.py``   -> ``Space`` (``space.py:15-129``)
* ``environment/physics0/cvTools.py`` -> ``getConvexHullActions`` (``cvTools.py:61-103``)
* ``tools.py``                        -> ``gen_ray_origin_direction`` etc.

No reference source is copied: the files are executed from where they lie.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("IRBPP_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "environment", "physics0", "space.py"))


def _euler2mat(ai, aj, ak, axes="sxyz"):
    # transforms3d.euler.euler2mat(ai, aj, ak, 'sxyz') == Rz(ak) @ Ry(aj) @ Rx(ai)
    assert axes == "sxyz"
    ci, si = np.cos(ai), np.sin(ai)
    cj, sj = np.cos(aj), np.sin(aj)
    ck, sk = np.cos(ak), np.sin(ak)
    rx = np.array([[1, 0, 0], [0, ci, -si], [0, si, ci]])
    ry = np.array([[cj, 0, sj], [0, 1, 0], [-sj, 0, cj]])
    rz = np.array([[ck, -sk, 0], [sk, ck, 0], [0, 0, 1]])
    return rz @ ry @ rx


def _install_stubs():
    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    stub("trimesh")
    stub("pybullet")
    mpl = stub("matplotlib")
    plt = stub("matplotlib.pyplot")
    mpl.pyplot = plt
    gym = stub("gym", Env=object)
    envs = stub("gym.envs")
    reg = stub("gym.envs.registration", register=lambda *a, **k: None)
    gym.envs = envs
    envs.registration = reg
    t3d = stub("transforms3d")
    eul = stub("transforms3d.euler", euler2mat=_euler2mat)
    t3d.euler = eul


_CACHE = {}


def _load(modname, relpath):
    if modname in _CACHE:
        return _CACHE[modname]
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    _CACHE[modname] = mod
    return mod


def load_reference():
    """Return (tools, space, cvTools) modules of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    _install_stubs()
    tools = _load("tools", "tools.py")  # space.py does `from tools import ...`
    space = _load("irbpp_ref_space", os.path.join("environment", "physics0", "space.py"))
    cvtools = _load("irbpp_ref_cvTools", os.path.join("environment", "physics0", "cvTools.py"))
    return tools, space, cvtools


class MeshStandIn(object):
    """What ``Space.get_possible_position`` needs from a trimesh object: ``.extents``
    (``space.py:104``)."""

    def __init__(self, extents):
        self.extents = np.asarray(extents, dtype=np.float64)

"""``make_vec_envs`` mirror (reference ``envs.py:67-99``).

The reference builds ``VecPyTorch(ShmemVecEnv([make_env(...)] * num_processes))`` from an argparse
namespace that carries the datasets.  Here the same call returns a ``GpuVecEnv`` with the same triple
``(envs, [observation_space, action_space], obs_len)``.  ``args`` needs only the attributes the hot
path reads (``binPhy.py:25-49``):

    num_processes, device, bin_dimension, resolutionA, resolutionH, resolutionZ, ZRotNum,
    selectedAction, bufferSize, seed

plus the shape data: ``args.shapeLibrary`` (a ``shapes.ShapeLibrary``) or the reference's own
``args.shotInfo`` / ``args.shapeDict`` (extents) / ``args.infoDict`` (volumes), and optionally
``args.itemSequences`` (int ids ``[num_processes, L]``, replayed modulo L).  Without it the ids are drawn
i.i.d. uniform ON THE DEVICE from a counter-based generator seeded with ``args.seed`` -- the role of
``RandomItemCreator`` (``IRcreator.py:26-33``); there is no period.  ``args.approxLegacy`` (optional, default
False) selects the older point-to-line ``approxPolyDP`` rule (the reference pins OpenCV 4.4.0.46)."""
import numpy as np

from . import shapes
from .vec_env import GpuVecEnv


def library_from_reference_args(args):
    """Build a ``ShapeLibrary`` from the reference's ``shotInfo`` / ``shapeDict`` / ``infoDict``."""
    shot = args.shotInfo
    ids = sorted(shot.keys())
    R = len(shot[ids[0]])
    ext = np.zeros((len(ids), R, 3)); vol = np.zeros(len(ids)); tables = []
    for s, k in enumerate(ids):
        for r in range(R):
            ext[s, r] = np.asarray(args.shapeDict[k][r].extents, dtype=np.float64)   # space.py:104
        vol[s] = float(args.infoDict[k][0]["volume"])                                   # binPhy.py:151
        tables.append([tuple(np.asarray(m, dtype=np.float64) for m in shot[k][r]) for r in range(R)])
    return shapes.ShapeLibrary(args.resolutionH, args.resolutionA, ext, vol, tables, name="reference_args")


def make_vec_envs(args, log_dir=None, allow_early_resets=True):
    """-> ``(envs, [observation_space, action_space], obs_len)`` as reference ``envs.py:67-99``.
    ``log_dir`` / ``allow_early_resets`` are accepted for signature compatibility (the Monitor CSV of
    ``envs.py:48-52`` is not written; episode info is delivered in ``infos`` as the trainer reads it)."""
    lib = getattr(args, "shapeLibrary", None)
    if lib is None:
        lib = library_from_reference_args(args)
    n = int(args.num_processes)
    seqs = getattr(args, "itemSequences", None)        # None: i.i.d. ids generated on the device
    envs = GpuVecEnv(lib, seqs, num_envs=n, device=getattr(args, "device", "cuda:0"),
                     item_seed=int(getattr(args, "seed", 0)), approx_legacy=bool(getattr(args, "approxLegacy", False)),
                     selected_action=int(getattr(args, "selectedAction", 500)),
                     buffer_size=int(getattr(args, "bufferSize", 1)),
                     bin_dimension=tuple(getattr(args, "bin_dimension", (0.32, 0.32, 0.30))),
                     resolution_act=float(getattr(args, "resolutionA", 0.02)),
                     resolution_h=float(getattr(args, "resolutionH", 0.01)),
                     resolution_z=float(getattr(args, "resolutionZ", 0.01)))
    return envs, [envs.observation_space, envs.action_space], envs.obs_len

"""The reference's on-disk table cache (tools.shotInfoPre, tools.py:248-279): a library written in that
layout is read back unchanged, and the files are what the reference's own loader line expects."""
import os

import numpy as np
import pytest
import torch

from irbpp_b200 import shapes


@pytest.mark.parametrize("kind", ["blockout", "irregular"])
def test_cache_round_trip(tmp_path, kind):
    lib = shapes.make_blockout_library(6, seed=3) if kind == "blockout" else \
        shapes.make_irregular_library(5, seed=4, num_rotations=8)
    d = tmp_path / shapes.shotinfo_dir_name("blockout", "id2shape", lib.resolutionH)
    assert d.name == "blockout_id2shape_0.01"                       # tools.py:260
    shapes.save_shotinfo_dir(lib, str(d))
    assert sorted(os.listdir(d))[0] == "0_0.pt"
    # the reference's loader line (tools.py:271): a 4-sequence of arrays
    T, B, mT, mB = torch.load(str(d / "1_2.pt"), weights_only=False)
    assert np.array_equal(T, lib.tables[1][2][0]) and np.array_equal(mB, lib.tables[1][2][3])
    back = shapes.load_shotinfo_dir(str(d))
    assert back.num_shapes == lib.num_shapes and back.num_rotations == lib.num_rotations
    assert np.array_equal(back.extents, lib.extents) and np.array_equal(back.volume, lib.volume)
    assert np.array_equal(back.dims, lib.dims)
    for a, b in zip(back.flat(), lib.flat()):
        assert np.array_equal(a, b)
    # a cache written by the reference carries no extents: they come from the caller (args.infoDict)
    os.remove(str(d / shapes.SHOTINFO_META))
    with pytest.raises(ValueError):
        shapes.load_shotinfo_dir(str(d))
    again = shapes.load_shotinfo_dir(str(d), extents=lib.extents, volume=lib.volume, resolutionH=lib.resolutionH,
                                     resolutionAct=lib.resolutionAct)
    assert np.array_equal(again.flat()[3], lib.flat()[3])
    os.remove(str(d / "0_1.pt"))
    with pytest.raises(ValueError):
        shapes.load_shotinfo_dir(str(d), extents=lib.extents, volume=lib.volume, resolutionH=0.01, resolutionAct=0.02)

"""The C oracle must reproduce every committed reference fixture bit for bit (CPU only)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from golden_util import layout_fixtures, load, PLANES, MAPS


@pytest.mark.parametrize("path", layout_fixtures(), ids=lambda p: os.path.basename(p)[7:-4])
@pytest.mark.parametrize("fast_reach", [0, 1])
def test_oracle_reproduces_golden(path, fast_reach):
    g = load(path)
    cfgd = dict(g["meta"]["cfg"])
    carve = cfgd.pop("carve_subblock_roads")
    oc = O.OracleCity(O.make_cfg(fast_reach=fast_reach, **cfgd), g["hbands"], g["vbands"])
    oc.run_all(g["tape_zone"], g["tape_carve"], g["tape_entrance"], carve=carve)
    assert oc.n_blocks == g["meta"]["n_blocks"]
    for f in PLANES:
        assert np.array_equal(oc.planes()[f], g[f]), f
    for k in ("lights", "ctrl", "incoming", "outgoing"):
        assert np.array_equal(oc.links[k], g["links_" + k]), k
    maps = oc.simple_maps()
    for k in MAPS:
        assert np.array_equal(maps[k], g[k]), k


def test_fixtures_exist():
    assert len(layout_fixtures()) >= 12

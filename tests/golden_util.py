"""Loading of the committed reference fixtures (tests/golden/*.npz)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def layout_fixtures():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "layout_*.npz")))


def load(path):
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(bytes(d["meta"]).decode())
    d["name"] = os.path.basename(path)[len("layout_"):-4]
    return d


PLANES = ("cell_type", "dirs", "aux", "block_id")
MAPS = ("is_road_map", "road_type_map", "intersection_map", "allowed_dirs_map")

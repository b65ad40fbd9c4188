"""Shared by the tests and the developer scripts: golden-case table and loaders."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# (scan_id, shape, n_points, n_traversals, frames_per_traversal) -- must match oracle/make_golden.py
GOLDEN_CASES = {
    "small": (7, "lyft", 9000, 3, 1),
    "lyft60k_t2": (0, "lyft", 60000, 2, 1),
    "nusc_small": (11, "nusc", 8000, 3, 2),
}


def build_case(name):
    from modest_b200 import synth
    scan_id, shape_name, n_pts, n_trav, fpt = GOLDEN_CASES[name]
    shape = synth.NUSC if shape_name == "nusc" else synth.LYFT
    case = synth.make_scan_case(scan_id, shape, n_traversals=n_trav, frames_per_traversal=fpt, n_points=n_pts)
    return case, shape


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False)



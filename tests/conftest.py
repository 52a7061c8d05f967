import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robotics-path-planning_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(g["meta"]))
    return g, meta


def golden_names(prefix):
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def scenario_args(meta):
    """positional arguments shared by oracle.make_params / pyport.RRTStarPort"""
    return (meta["start"], meta["goal"], meta["obstacle_list"], meta["expand_dis"],
            meta["path_resolution"], meta["max_iter"], meta["play_area"], meta["robot_radius"],
            meta["connect_circle_dist"], meta["search_until_max_iter"])


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle as O
    O.build()
    return O

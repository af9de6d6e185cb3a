import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")
MESH_FILE = os.path.join(GOLD, "unstructured_square.txt")
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_mesh_file(path=MESH_FILE):
    tok = open(path).read().split()
    nv, nel = int(tok[0]), int(tok[1])
    xy = np.array([float(t) for t in tok[2:2 + 2 * nv]]).reshape(nv, 2)
    el = np.array([int(t) for t in tok[2 + 2 * nv:]]).reshape(nel, 4)
    return xy, el


@pytest.fixture(scope="session")
def unstructured():
    return load_mesh_file()


@pytest.fixture(scope="session")
def gold():
    class G:
        tables = np.load(os.path.join(GOLD, "tables.npz"))
        h1 = np.load(os.path.join(GOLD, "h1.npz"))
        ens = np.load(os.path.join(GOLD, "ensemble.npz"))
        import json
        hashes = json.load(open(os.path.join(GOLD, "hashes.json")))
    return G

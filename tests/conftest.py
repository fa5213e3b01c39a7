import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    from __graft_entry__ import load_package
    return load_package()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    from __graft_entry__ import load_package
    p = load_package()
    cache = {}

    def get(name):
        if name not in cache:
            npz = np.load(os.path.join(GOLDEN, name + ".npz"))
            scn = os.path.join(GOLDEN, name + ".scn")
            cache[name] = (dict(npz), p.sceneio.read_scene(scn) if os.path.exists(scn) else None)
        return cache[name]
    return get


@pytest.fixture(scope="module")
def pv_factory(pkg):
    made = []

    def make(**kw):
        pv = pkg.PhotonVolume(device=0, **kw)
        made.append(pv)
        return pv
    yield make
    for pv in made:
        pv.close()

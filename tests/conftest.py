import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def rb():
    import renderer_b200
    renderer_b200.lib()   # fails loudly if the library was not built
    return renderer_b200


@pytest.fixture(scope="session")
def pyport():
    from oracle import pyport as pp
    pp.port()
    return pp


_scene_cache = {}


@pytest.fixture(scope="session")
def load_scene(rb, pyport):
    """load_scene('torus.ply') -> renderer_b200.Scene with its BVH (cached beside the staged model,
    in the reference's own .bvh format)."""
    def _load(name, bvh=True):
        key = (name, bvh)
        if key not in _scene_cache:
            path = pyport.model_path(name)
            if not os.path.exists(path):
                pytest.skip(f"model {name} not staged (run __graft_entry__.build() where /root/reference exists)")
            s = rb.Scene(path)
            if bvh:
                s.UpdateBoundingVolumeHierarchy(path + ".bvh")
            _scene_cache[key] = s
        return _scene_cache[key]
    return _load


@pytest.fixture(scope="session")
def gpu(rb):
    r = rb.Renderer(0)
    yield r
    r.close()

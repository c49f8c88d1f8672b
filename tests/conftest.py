import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def obj_path():
    from rayito_b200 import build
    build.stage_assets()
    return build.model_path("bumpy.obj")


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref).  Tests that need it are skipped when
    the library was not built (it is built by __graft_entry__.build())."""
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref/libref_s7.so not built")
    return refapi


@pytest.fixture(scope="session")
def capi():
    from rayito_b200 import capi as _capi
    _capi.host()
    return _capi


@pytest.fixture(scope="session")
def scene1_host(capi, obj_path):
    return capi.HostScene(capi.RECIPE_STAGE7_SCENE1, obj_path)


@pytest.fixture(scope="session")
def scene1_ref(ref, obj_path):
    return ref.RefScene(1, obj_path)


@pytest.fixture(scope="session")
def scene2_host(capi):
    return capi.HostScene(capi.RECIPE_STAGE7_SCENE2)


@pytest.fixture(scope="session")
def scene2_ref(ref):
    return ref.RefScene(2)


SYNTH_GRID = (192, 160)


@pytest.fixture(scope="session")
def scene5_host(capi):
    return capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, SYNTH_GRID)


@pytest.fixture(scope="session")
def scene5_ref(ref):
    return ref.RefScene(5, None, SYNTH_GRID)


@pytest.fixture(scope="session")
def ref6():
    """The compiled Stage 6 reference (oracle/_ref/libref_s6.so)."""
    from oracle import refapi
    if not refapi.available(6):
        pytest.skip("oracle/_ref/libref_s6.so not built")
    return refapi


@pytest.fixture(scope="session")
def scene6_host(capi, obj_path):
    return capi.HostScene(capi.RECIPE_STAGE6_SCENE, obj_path)


@pytest.fixture(scope="session")
def scene6_ref(ref6, obj_path):
    return ref6.RefScene(6, obj_path, stage=6)


@pytest.fixture(scope="session")
def scene7_host(capi):
    return capi.HostScene(capi.RECIPE_EDGE_LINEAR_LIST)


@pytest.fixture(scope="session")
def scene7_ref(ref):
    return ref.RefScene(7)


@pytest.fixture(scope="session")
def scene8_host(capi):
    return capi.HostScene(capi.RECIPE_EDGE_NO_LIGHTS)


@pytest.fixture(scope="session")
def scene8_ref(ref):
    return ref.RefScene(8)


# Deep-tree edge scenes (fixtures/scene_recipes.h buildDeepScene): face BVH 42 / 46 deep, and with
# the sphere chain a top-level BVH 18 deep as well
DEEP_GRID = (40, 8)
DEEPER_GRID = (44, 8)
BIG_GRID = (1000, 1000)     # the displaced sphere at 1 M quads: face BVH depth 32 -> 33 stack entries


@pytest.fixture(scope="session")
def deep_host(capi):
    return capi.HostScene(capi.RECIPE_EDGE_DEEP_MESH, None, DEEP_GRID)


@pytest.fixture(scope="session")
def deep_ref(ref):
    return ref.RefScene(10, None, DEEP_GRID)


@pytest.fixture(scope="session")
def deepboth_host(capi):
    return capi.HostScene(capi.RECIPE_EDGE_DEEP_BOTH, None, DEEPER_GRID)


@pytest.fixture(scope="session")
def deepboth_ref(ref):
    return ref.RefScene(11, None, DEEPER_GRID)

import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the in-tree native pieces are git-ignored build products: build them if a fresh checkout lacks them
    from hijiki_b200 import _abi
    if not os.path.exists(_abi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def oracle():
    import _libs
    return _libs.oracle()


@pytest.fixture(scope="session")
def hosttest():
    import _libs
    return _libs.hosttest()


@pytest.fixture(scope="session")
def cbox(hosttest):
    import _libs
    return _libs.HostScene.from_obj(hosttest, _libs.CBOX_OBJ, put_spheres=False, with_bvh2=True)


@pytest.fixture(scope="session")
def cbox_spheres(hosttest):
    import _libs
    return _libs.HostScene.from_obj(hosttest, _libs.CBOX_OBJ, put_spheres=True, with_bvh2=True)


@pytest.fixture(scope="session")
def gpu_ctx():
    """One CUDA context for the whole GPU session (fails loudly if the library or GPU is missing)."""
    import hijiki_b200 as hj
    ctx = hj.Context(0)
    yield ctx
    ctx.close()

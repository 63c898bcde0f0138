import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def zmodel():
    import ksim_gym_zbot_b200  # noqa: F401
    from ksim_gym_zbot_b200 import model as M

    return M.build_zbot_syn()


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    import zb_oracle

    zb_oracle.build()

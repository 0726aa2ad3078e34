import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def restated():
    import oracle
    oracle.build()
    return oracle.Restated()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference sources as a shared library; skipped when it cannot be built."""
    import oracle
    try:
        oracle.build()
        return oracle.Reference()
    except Exception as e:  # pragma: no cover - only on boxes without /root/reference and no prebuilt .so
        pytest.skip("reference library unavailable: %s" % e)


@pytest.fixture(scope="session", params=["port", "reference"])
def any_oracle(request, restated):
    if request.param == "port":
        return restated
    return request.getfixturevalue("reference")

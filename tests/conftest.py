import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


_PROBLEMS = {}


@pytest.fixture(scope="session")
def problem_factory():
    import dycore_b200  # noqa: F401
    from dycore_b200 import harness

    def make(**spec):
        key = tuple(sorted(spec.items()))
        if key not in _PROBLEMS:
            _PROBLEMS[key] = harness.Problem(**spec)
        return _PROBLEMS[key]

    return make

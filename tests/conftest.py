import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


@pytest.fixture(autouse=True)
def _fresh_weight_cache():
    try:
        from dynavsr_b200 import ops
        ops.invalidate_weight_cache()
        ops._wcache.clear()
        ops._default_scope.registry.clear()
        ops._default_scope.table.update(dev=None, n=0, blocks=0, dirty=True)
    except Exception:
        pass
    yield

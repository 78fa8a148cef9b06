import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'beta-cores_b200'), os.path.join(ROOT, 'beta-cores_b200', 'examples', 'common'),
          os.path.join(ROOT, 'tests', 'golden')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


# The samplers of the test problems do D x D linear algebra (D <= 100) thousands of times: a multi-threaded BLAS pool only
# thrashes on such sizes (111.7 s vs 2.9 s per build iteration of the Gaussian example).
try:
    from threadpoolctl import threadpool_limits
    _BLAS_LIMIT = threadpool_limits(limits=1, user_api='blas')
except Exception:      # pragma: no cover
    _BLAS_LIMIT = None

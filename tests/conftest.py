import os
import sys

import pytest

# the licensed SMPL .pkl files are not in the repo: tests opt in to the synthetic SMPL-shaped model
os.environ.setdefault('PRK_SYNTHETIC_SMPL', '1')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


GOLDEN = os.path.join(ROOT, 'tests', 'golden')


@pytest.fixture(scope='session')
def golden():
    import numpy as np
    return {k: np.load(os.path.join(GOLDEN, k + '.npz')) for k in ('smpl_forward', 'euler', 'scores')}

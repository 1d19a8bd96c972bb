import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
CKPT_PLAIN = os.path.join(ROOT, 'oracle', '_ref', 'BMCNet_plain_nfs_x4.pth')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def plain_ckpt():
    """The one checkpoint the reference ships; staged by oracle/make_golden.py (git-ignored)."""
    if not os.path.exists(CKPT_PLAIN):
        pytest.skip('oracle/_ref/BMCNet_plain_nfs_x4.pth not staged (run oracle/make_golden.py)')
    import torch
    return torch.load(CKPT_PLAIN, map_location='cpu')

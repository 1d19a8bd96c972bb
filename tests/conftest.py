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


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on the CPU build host must not fail on the GPU tests: they are skipped (not
    silently passed) when there is no CUDA device or the library has not been built."""
    reason = None
    try:
        import torch
        if not torch.cuda.is_available():
            reason = 'no CUDA device (GPU tests run on the B200 box with -m gpu)'
    except Exception as e:          # pragma: no cover
        reason = 'torch unavailable: %s' % e
    if reason is None and not os.path.exists(os.path.join(ROOT, 'bmcnet_esr_b200', 'libbmc_b200.so')):
        reason = 'bmcnet_esr_b200/libbmc_b200.so not built (python -m bmcnet_esr_b200.build)'
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


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

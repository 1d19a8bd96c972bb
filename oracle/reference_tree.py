"""Access to the UNMODIFIED reference staged under baseline/_ref -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.

`__graft_entry__.build()` copies the reference's Python tree there in the build container (git-ignored: it never
enters history; not gpurun-ignored: it travels to the GPU box, where /root/reference does not exist).  Only
`tests/`, `bench.py --impl reference` and bench.py's `cpu_baseline` / `eager_b200` legs use this module; nothing
under bmcnet_esr_b200/ may import it.

The reference's scripts import h5py, matplotlib, skimage, cv2 and open3d at module scope (infer_BMCNet.py:11-16 via
dataloader/h5dataset.py, loss/restore.py, myutils/vis_events/*); none of them is installed here and none is
touched by `load_model` / `infer_body`.  `stub_missing_third_party()` registers inert placeholder modules for
exactly those names when (and only when) the real package is absent.
"""
import importlib
import importlib.util
import os
import sys
import types
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.path.join(ROOT, 'baseline', '_ref')
_ABSENT_OK = ['h5py', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.animation', 'mpl_toolkits',
              'mpl_toolkits.axes_grid1', 'cv2', 'skimage', 'skimage.metrics', 'open3d', 'lpips', 'IPython', 'tensorboardX', 'tensorboard']


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'models', 'BMCNet.py'))


def stub_missing_third_party():
    made = []
    for name in _ABSENT_OK:
        if name in sys.modules:
            continue
        top = name.split('.')[0]
        if top not in made and importlib.util.find_spec(top) is not None:
            continue                                   # the real package exists: use it
        m = types.ModuleType(name)
        m.__path__ = []                                # a package, so that submodule imports resolve
        m.__all__ = []
        m.__getattr__ = lambda attr, _n=name: mock.MagicMock(name='%s.%s' % (_n, attr))
        sys.modules[name] = m
        made.append(top)
    return made


def _on_path():
    if not available():
        raise RuntimeError('baseline/_ref is not staged: run `python __graft_entry__.py` where /root/reference exists')
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def models():
    """(BMCNet, BMCNet_plain) classes of the unmodified reference (models/BMCNet.py:87, models/BMCNet_plain.py:36)."""
    _on_path()
    return importlib.import_module('models.BMCNet').BMCNet, importlib.import_module('models.BMCNet_plain').BMCNet_plain


def encodings():
    """The unmodified reference dataloader/encodings.py module."""
    _on_path()
    return importlib.import_module('dataloader.encodings')


def load_script(name, swap=None):
    """Execute a reference top-level script (e.g. 'infer_BMCNet.py') as a module object without running its
    `__main__` block.  `swap` = (old import line, new import line): the ONE edit INTEGRATION.md section 2 asks a
    maintainer to make; the line must occur exactly once."""
    _on_path()
    stub_missing_third_party()
    path = os.path.join(REF_ROOT, name)
    src = open(path).read()
    if swap is not None:
        old, new = swap
        if src.count(old) != 1:
            raise RuntimeError('%s: expected exactly one occurrence of %r' % (name, old))
        src = src.replace(old, new)
    mod = types.ModuleType('ref_' + name.replace('.py', '') + ('_swapped' if swap else ''))
    mod.__file__ = path
    exec(compile(src, path, 'exec'), mod.__dict__)
    return mod

"""The reference's own hot-path files for runs on the GPU box (TEST / BASELINE INFRASTRUCTURE ONLY).

``/root/reference`` exists only in the authoring container.  ``ensure()`` (called by
``__graft_entry__.build()``) copies the few files the reference needs to build ``VQGANFCM`` and to
evaluate its losses -- unmodified -- into the git-ignored ``baseline/_ref/``, which
travels to the GPU box with the snapshot.  Nothing in ``favae_b200/`` imports this module; users are
``bench.py --impl reference`` (the timed CPU baseline) and ``tests/test_gpu_reference_model.py``.
The pip package ``focal-frequency-loss`` is absent everywhere, so the spectrum-loss callable handed to
the reference's wrappers is the restatement in ``oracle/ffl_oracle.py``.
"""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref')
FILES = ['models/__init__.py', 'models/codec.py', 'models/discriminator.py', 'models/l2_quantize.py',
         'models/vqgan_fcm.py', 'losses/vqgan_losses.py', 'losses/hinge.py', 'utils.py', 'LICENSE.md']


def ensure() -> str | None:
    """Copy the files when the reference checkout is present; return the tree's path or None."""
    if os.path.isdir(SRC):
        for f in FILES:
            dst = os.path.join(DST, f)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(os.path.join(SRC, f)):
                shutil.copyfile(os.path.join(SRC, f), dst)
    return path()


def path() -> str | None:
    ok = all(os.path.exists(os.path.join(DST, f)) for f in FILES)
    return DST if ok else None


def import_reference():
    """-> (models.l2_quantize, losses.vqgan_losses, models.vqgan_fcm) imported from the tree, or None.
    The tree is put FIRST on sys.path for the import and any previously imported ``models`` /
    ``losses`` packages of another origin are an error."""
    import importlib
    import sys
    import warnings
    p = path()
    if p is None:
        return None
    warnings.filterwarnings('ignore', category=FutureWarning)
    for name in ('models', 'losses'):
        m = sys.modules.get(name)
        if m is not None and not any(str(x).startswith(p) for x in getattr(m, '__path__', [])):
            raise RuntimeError(f'another package named {name} is already imported')
    sys.path.insert(0, p)
    try:
        l2q = importlib.import_module('models.l2_quantize')
        vl = importlib.import_module('losses.vqgan_losses')
        fcm = importlib.import_module('models.vqgan_fcm')
    finally:
        sys.path.remove(p)
    return l2q, vl, fcm

"""favae_b200 -- B200-native hot path of FA-VAE: vector-quantizer search and spectrum losses.

Public surface (mirrors the reference):

    from favae_b200 import VectorQuantize                  # models/l2_quantize.py
    from favae_b200 import FocalFrequencyLoss              # pip focal_frequency_loss
    from favae_b200.vqgan_losses import *                  # losses/vqgan_losses.py
    from favae_b200 import gaussian_blur_reflect           # the five _gaussian_blur copies
    favae_b200.patch_reference()                           # install all of the above into an
                                                           # importable reference checkout
"""
from __future__ import annotations

import sys

from .focal_frequency_loss import FocalFrequencyLoss
from .gaussian_blur import LazyBlur, gaussian_blur_reflect, install_reference_blur, lazy_gaussian_blur
from .l2_quantize import CosineSimCodebook, EuclideanCodebook, VectorQuantize
from .vqgan_losses import recon_ffl_features_loss, recon_ffl_loss, recon_sl_gaussian_features_loss

__all__ = ['VectorQuantize', 'CosineSimCodebook', 'EuclideanCodebook', 'FocalFrequencyLoss',
           'gaussian_blur_reflect', 'lazy_gaussian_blur', 'LazyBlur', 'recon_ffl_loss', 'recon_ffl_features_loss',
           'recon_sl_gaussian_features_loss', 'patch_reference', 'build']

__version__ = '0.1.0'


def build(force=False):
    """Compile the C-ABI library in-tree (nvcc, sm_100a)."""
    from . import _build
    return _build.build(force=force)


def patch_reference():
    """Make an importable FA-VAE checkout use this package (call before building the model).

    * ``focal_frequency_loss`` resolves to :mod:`favae_b200.focal_frequency_loss`
      (``favae_scripts/train_favae.py:27``);
    * ``models.l2_quantize.VectorQuantize`` (+ codebook classes) are replaced
      (``models/vqgan_fcm.py:102``);
    * ``losses.vqgan_losses`` functions are replaced (``favae_scripts/train_favae.py:24``);
    * every ``_gaussian_blur`` method in ``models.vqgan_fcm`` / ``models.codec`` is replaced by
      the deferred blur (``LazyBlur``), which ``recon_ffl_features_loss`` fuses with the spectrum
      loss (``models/codec.py:284-309, 655-686, 978-999, 1105-1123``, ``models/vqgan_fcm.py:131-134``).
    """
    from . import focal_frequency_loss as ffl_mod
    from . import l2_quantize as q_mod
    from . import vqgan_losses as l_mod
    prev = sys.modules.get('focal_frequency_loss')
    if prev is not None and prev is not ffl_mod and \
            getattr(prev, 'FocalFrequencyLoss', None) is not ffl_mod.FocalFrequencyLoss:
        import warnings
        warnings.warn('favae_b200.patch_reference(): the pip package focal_frequency_loss was already imported; '
                      'it is replaced in sys.modules, but names bound by earlier `from focal_frequency_loss '
                      'import FocalFrequencyLoss` statements still point at the pip class', stacklevel=2)
    sys.modules['focal_frequency_loss'] = ffl_mod
    done = ['focal_frequency_loss']
    try:
        import models.l2_quantize as ref_q
        for name in ('VectorQuantize', 'CosineSimCodebook', 'EuclideanCodebook'):
            setattr(ref_q, name, getattr(q_mod, name))
        done.append('models.l2_quantize')
    except ImportError:
        pass
    try:
        import losses.vqgan_losses as ref_l
        for name in l_mod.__all__:
            setattr(ref_l, name, getattr(l_mod, name))
        done.append('losses.vqgan_losses')
    except ImportError:
        pass
    try:
        import models.codec as codec
        import models.vqgan_fcm as fcm
        classes = [c for m in (codec, fcm) for c in vars(m).values()
                   if isinstance(c, type) and '_gaussian_blur' in vars(c)]
        install_reference_blur(*classes)
        done.append('_gaussian_blur x%d' % len(classes))
    except ImportError:
        pass
    return done

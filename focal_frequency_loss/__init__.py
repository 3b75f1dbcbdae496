"""Alias so that ``from focal_frequency_loss import FocalFrequencyLoss``
(/root/reference/favae_scripts/train_favae.py:27) resolves to the B200 implementation when the
repo root is on ``sys.path``."""
from favae_b200.focal_frequency_loss import FocalFrequencyLoss  # noqa: F401

__all__ = ['FocalFrequencyLoss']

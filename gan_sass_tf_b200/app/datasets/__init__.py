"""Datasets of the spectral path (reference: ``app/datasets/``).  Importing the package
registers ``'toy'`` (dataset.py:45-76) and ``'wave'`` (device-resident waveforms)."""
from . import dataset, wave  # noqa: F401

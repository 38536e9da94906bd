"""openmeters_b200 — B200-native (sm_100a CUDA) implementation of the OpenMeters DSP hot path.

Public surface mirrors the reference's processors (see processors.py) plus the
batched plans (batch.py).  Importing the package does not load the CUDA
extension; the first processor / plan construction does, and fails loudly if it
is missing.
"""
from . import _capi as capi  # noqa: F401
from .processors import (  # noqa: F401
    AudioBlock, LoudnessConfig, LoudnessProcessor, LoudnessSnapshot, OmbError, SpectrogramConfig,
    SpectrogramProcessor, SpectrogramUpdate, SpectrumConfig, SpectrumProcessor, SpectrumSnapshot,
)

__version__ = "0.1.0"

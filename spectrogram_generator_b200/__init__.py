"""B200-native STFT / spectrogram engine: a drop-in for the spectrogram path of
Karmotr1ne/Spectrogram-Generator (``scipy.signal.spectrogram`` as called at
PlotEngine.py:113 / :232).  See DESIGN.md and INTEGRATION.md.

    from spectrogram_generator_b200 import spectrogram      # instead of scipy.signal

The compute runs in hand-written sm_100a CUDA kernels (``libb200stft.so``, C ABI
in ``include/b2s.h``); importing the package does not need a GPU, calling it does
-- there is no CPU fallback.
"""
from .spectrogram import (Engine, Plan, engine, mean_spectrogram, pinned_empty, spectrogram,
                          spectrogram_batch, spectrogram_chunked, split_frames, triage)
from .plot_engine import SpectrogramPath
from . import distributed, plot_engine, synth, windows

__all__ = ["spectrogram", "spectrogram_batch", "mean_spectrogram", "spectrogram_chunked",
           "split_frames", "pinned_empty", "engine", "Engine", "Plan", "triage",
           "SpectrogramPath", "distributed", "synth", "windows"]

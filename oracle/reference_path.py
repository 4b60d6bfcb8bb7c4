"""The reference's own spectrogram path, run as the reference runs it -- TEST
INFRASTRUCTURE ONLY (see oracle/__init__.py).

/root/reference/PlotEngine.py cannot be imported in this image (PyQt5,
matplotlib, hmmlearn are absent), but the one call that *is* its spectrogram
path can: ``from scipy.signal import spectrogram`` (PlotEngine.py:8) invoked as
``spectrogram(data, fs=fs, nperseg=nperseg, scaling="density", mode="psd")``
(PlotEngine.py:113, :232).  SciPy ships in this image (1.18.1) on the build box
and the GPU box alike, so this module is both the pin for ``stft_oracle`` and
the "reference" CPU arm ``bench.py`` times on the host cores.
"""
from __future__ import annotations

import os

import numpy as np
from scipy.signal import spectrogram as _scipy_spectrogram   # == PlotEngine.py:8

from . import stft_oracle


def reference_call(data, fs, nperseg):
    """Exactly PlotEngine.py:113 / :232 (SciPy defaults for everything else:
    periodic Tukey(0.25), noverlap = nperseg//8, detrend='constant')."""
    return _scipy_spectrogram(data, fs=fs, nperseg=nperseg, scaling="density", mode="psd")


def reference_call_kw(data, fs, **kw):
    """The same library entry with explicit window / noverlap (how the
    BASELINE configs reach Hann / 75 % overlap through the reference's path)."""
    kw.setdefault("scaling", "density")
    kw.setdefault("mode", "psd")
    return _scipy_spectrogram(data, fs=fs, **kw)


def plot_spectrogram_compute(data, fs, settings, global_max=None):
    """PlotEngine._plot_spectrogram up to (not including) the matplotlib
    calls: PlotEngine.py:110-131."""
    f, t, Sxx = reference_call(data, fs, settings["nperseg"])
    return stft_oracle.plot_postprocess(f, t, Sxx, settings["fmin"], settings["fmax"],
                                        settings["log_scale"], global_max)


def calculate_features(signal, fs, settings):
    """PlotEngine._calculate_features: PlotEngine.py:229-242."""
    f, t, Sxx = reference_call(signal, fs, settings["nperseg"])
    return stft_oracle.band_features(np.asarray(f), t, np.asarray(Sxx),
                                     settings["fmin"], settings["fmax"])


# ---- CPU timing arm -------------------------------------------------------

def _worker(args):
    x, fs, kw, want_db, floor_rel = args
    f, t, S = reference_call_kw(x, fs, **kw)
    if want_db:
        S = 10.0 * np.log10(np.maximum(S, floor_rel * S.max()))
    # each worker returns its partial sum over sweeps (the cross-sweep mean's numerator);
    # the per-sweep spectrograms stay in the worker, as they would stay on a GPU
    return S.shape, S.reshape((-1,) + S.shape[-2:]).sum(axis=0)


def run_sharded(x2d, fs, kw, n_procs, want_db=False, floor_rel=1e-6):
    """All-cores CPU arm: rows of ``x2d`` (sweeps / channels / halo'd chunks)
    sharded over ``n_procs`` processes, each running the as-is SciPy call
    (BASELINE.md section 3)."""
    import multiprocessing as mp
    n_procs = max(1, int(n_procs))
    shards = [s for s in np.array_split(x2d, min(n_procs, len(x2d)), axis=0) if len(s)]
    if n_procs == 1:
        return [_worker((s, fs, kw, want_db, floor_rel)) for s in shards]
    ctx = mp.get_context("fork")
    with ctx.Pool(n_procs) as pool:
        return pool.map(_worker, [(s, fs, kw, want_db, floor_rel) for s in shards])


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:      # pragma: no cover
        return os.cpu_count() or 1

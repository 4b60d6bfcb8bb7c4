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
    """All-cores CPU arm, one shot: rows of ``x2d`` (sweeps / channels / halo'd chunks)
    sharded over ``n_procs`` processes, each running the as-is SciPy call
    (BASELINE.md section 3)."""
    with ShardedRunner(x2d, fs, kw, n_procs, want_db, floor_rel) as r:
        return r.step()


_SHARED = {}


def _worker_range(args):
    key, lo, hi = args
    x2d, fs, kw, want_db, floor_rel = _SHARED[key]
    return _worker((x2d[lo:hi], fs, kw, want_db, floor_rel))


class ShardedRunner:
    """The all-cores CPU arm with its set-up outside the timed region: the worker
    processes are forked once and inherit the input array (no pickling of samples);
    every ``step()`` is one pass of the as-is SciPy call over all rows, sharded over
    the processes, each returning its partial cross-sweep sum."""

    def __init__(self, x2d, fs, kw, n_procs, want_db=False, floor_rel=1e-6):
        import multiprocessing as mp
        self.n_procs = max(1, int(n_procs))
        self.key = id(self)
        _SHARED[self.key] = (x2d, fs, dict(kw), want_db, floor_rel)
        n = len(x2d)
        parts = min(self.n_procs, n) if n else 1
        edges = np.linspace(0, n, parts + 1).astype(int)
        self.ranges = [(self.key, int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        self.pool = mp.get_context("fork").Pool(self.n_procs) if self.n_procs > 1 else None

    def step(self):
        if self.pool is None:
            return [_worker_range(r) for r in self.ranges]
        return self.pool.map(_worker_range, self.ranges, chunksize=1)

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None
        _SHARED.pop(self.key, None)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:      # pragma: no cover
        return os.cpu_count() or 1

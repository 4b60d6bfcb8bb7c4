"""The reference's OWN ``PlotEngine`` methods on the spectrogram path, executed -- TEST
INFRASTRUCTURE ONLY (see oracle/__init__.py).

``/root/reference/PlotEngine.py`` imports PyQt5, matplotlib and hmmlearn at module level,
none of which is installed here; none of them is touched by the arithmetic of the three
methods on the hot path.  So the module is loaded with inert stand-ins for exactly those
imports, and the UNBOUND reference methods

  * ``PlotEngine._plot_spectrogram``      (PlotEngine.py:110-145)
  * ``PlotEngine._calculate_features``    (PlotEngine.py:229-242)
  * ``PlotEngine.calculate_absolute_power`` / ``calculate_band_powers``   (PlotEngine.py:686-719)

are called on a recording stand-in for ``self`` -- the reference's own lines run, SciPy and
NumPy are the real ones.  The normalised image is what the reference hands to
``ax_spec.pcolormesh`` (PlotEngine.py:134).  This pins ``stft_oracle.plot_postprocess /
band_features / band_powers`` (the restatement the GPU tests use where the reference tree is
absent) to the reference itself; ``tests/golden/make_plot_engine_golden.py`` stores its
outputs as fixtures that travel to the GPU box.

Nothing here is imported by the product.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_FILE = "/root/reference/PlotEngine.py"

_STUBS = ["PyQt5", "PyQt5.QtWidgets", "PyQt5.QtCore", "PyQt5.QtGui", "matplotlib", "matplotlib.backends",
          "matplotlib.backends.backend_qt5agg", "matplotlib.figure", "matplotlib.colors", "hmmlearn",
          "hmmlearn.hmm"]


def available(path: str = REFERENCE_FILE) -> bool:
    return os.path.exists(path)


def load_plot_engine(path: str = REFERENCE_FILE):
    """Import the reference's PlotEngine.py with stand-ins for the GUI / plotting / HMM packages it
    imports at module level; returns the reference's ``PlotEngine`` class (never instantiated)."""
    injected = []
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            if importlib.util.find_spec(name) is not None:
                continue
        except (ImportError, ValueError, AttributeError):
            pass
        m = types.ModuleType(name)
        m.__path__ = []                                     # packages: allow "from a.b import c"
        sys.modules[name] = m
        injected.append(name)
    try:
        for name, attrs in {"PyQt5": ["QtWidgets", "QtCore", "QtGui"], "PyQt5.QtGui": ["QCursor"],
                            "matplotlib.backends.backend_qt5agg": ["FigureCanvasQTAgg"],
                            "matplotlib.figure": ["Figure"], "matplotlib.colors": ["LinearSegmentedColormap"],
                            "hmmlearn": ["hmm"]}.items():
            mod = sys.modules[name]
            for a in attrs:
                if not hasattr(mod, a):
                    sub = sys.modules.get(f"{name}.{a}")
                    setattr(mod, a, sub if sub is not None else type(a, (), {}))
        spec = importlib.util.spec_from_file_location("_reference_PlotEngine", path)
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
    finally:
        for name in injected:
            sys.modules.pop(name, None)
    return module.PlotEngine


class _Recorder:
    """Stand-in for a matplotlib Axes / Figure: accepts any call, remembers pcolormesh's arguments."""

    def __init__(self):
        self.pcolormesh_args = None

    def pcolormesh(self, t, f, image, **kw):
        self.pcolormesh_args = (np.array(t), np.array(f), np.array(image))
        return object()

    def __getattr__(self, name):
        return lambda *a, **k: None


class _Self:
    """What the three reference methods read and write on ``self``."""

    def __init__(self):
        self.ax_spec = _Recorder()
        self.fig = _Recorder()
        self.last_fs = None
        self.last_settings = None
        self.last_t = np.array([])
        self.last_f = None
        self.last_Sxx = None


def plot_spectrogram(PlotEngine, data, fs, settings, global_max=None):
    """Runs the reference's ``_plot_spectrogram`` (PlotEngine.py:110-145); returns
    ``dict(last_f, last_t, last_Sxx, image)`` -- ``image`` is None when the band mask is empty."""
    s = _Self()
    PlotEngine._plot_spectrogram(s, data, fs, settings, global_max)
    rec = s.ax_spec.pcolormesh_args
    return dict(last_f=s.last_f, last_t=s.last_t, last_Sxx=s.last_Sxx, image=None if rec is None else rec[2],
                state=s)


def calculate_features(PlotEngine, signal, fs, settings):
    """Runs the reference's ``_calculate_features`` (PlotEngine.py:229-242): ``(t, features[F, 2])``."""
    return PlotEngine._calculate_features(_Self(), signal, fs, settings)


def power_summaries(PlotEngine, state, bands=None):
    """Runs ``calculate_absolute_power`` and ``calculate_band_powers`` (PlotEngine.py:686-719) on the
    state a ``plot_spectrogram`` call left behind."""
    return PlotEngine.calculate_absolute_power(state), PlotEngine.calculate_band_powers(state, bands)

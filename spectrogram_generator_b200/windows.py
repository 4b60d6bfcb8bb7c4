"""Window tables and axis arrays, built on the host in float64.

The kernels multiply by an fp32 table; the table is produced here with the
same arithmetic SciPy uses (``scipy/signal/windows/_windows.py``) and rounded
once.  The recipes matter: e.g. SciPy's Hann is ``general_cosine`` over
``linspace(-pi, pi, M+1)`` and differs from ``0.5 - 0.5 cos(2 pi n / N)`` in the
last bit.  Windows outside this table are delegated to
``scipy.signal.get_window`` (host-side set-up only, SciPy is a dependency of the
reference anyway); array windows are used as given.
"""
from __future__ import annotations

import math

import numpy as np

_COSINE_SUMS = {
    "hann": (0.5, 0.5), "han": (0.5, 0.5),
    "hamming": (0.54, 1.0 - 0.54), "hamm": (0.54, 1.0 - 0.54), "ham": (0.54, 1.0 - 0.54),
    "blackman": (0.42, 0.50, 0.08), "black": (0.42, 0.50, 0.08), "blk": (0.42, 0.50, 0.08),
    "nuttall": (0.3635819, 0.4891775, 0.1365995, 0.0106411),
    "nutl": (0.3635819, 0.4891775, 0.1365995, 0.0106411),
    "nut": (0.3635819, 0.4891775, 0.1365995, 0.0106411),
    "blackmanharris": (0.35875, 0.48829, 0.14128, 0.01168),
    "blackharr": (0.35875, 0.48829, 0.14128, 0.01168),
    "bkh": (0.35875, 0.48829, 0.14128, 0.01168),
    "flattop": (0.21557895, 0.41663158, 0.277263158, 0.083578947, 0.006947368),
    "flat": (0.21557895, 0.41663158, 0.277263158, 0.083578947, 0.006947368),
    "flt": (0.21557895, 0.41663158, 0.277263158, 0.083578947, 0.006947368),
}
_BOXCAR = ("boxcar", "box", "ones", "rect", "rectangular")


def _cosine_sum(M, coeffs, sym):
    # _windows.py:55-65
    if M <= 1:
        return np.ones(M, dtype=np.float64)
    Mx = M if sym else M + 1
    fac = np.linspace(-np.pi, np.pi, Mx, dtype=np.float64)
    w = np.zeros(Mx, dtype=np.float64)
    for k, a in enumerate(coeffs):
        w += a * np.cos(k * fac)
    return w if sym else w[:-1]


def _tukey(M, alpha, sym):
    # _windows.py:880-966
    if M <= 1 or alpha <= 0:
        return np.ones(M, dtype=np.float64)
    if alpha >= 1.0:
        return _cosine_sum(M, _COSINE_SUMS["hann"], sym)
    Mx = M if sym else M + 1
    n = np.arange(0, Mx, dtype=np.float64)
    width = int(math.floor(alpha * (Mx - 1) / 2.0))
    n1, n2, n3 = n[0:width + 1], n[width + 1:Mx - width - 1], n[Mx - width - 1:]
    w1 = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * n1 / alpha / (Mx - 1))))
    w2 = np.ones(n2.shape)
    w3 = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * n3 / alpha / (Mx - 1))))
    w = np.concatenate((w1, w2, w3))
    return w if sym else w[:-1]


_WINDOW_CACHE = {}


def get_window(window, nperseg: int) -> np.ndarray:
    """float64 window of length ``nperseg`` == ``scipy.signal.get_window(window,
    nperseg)`` (fftbins=True; ``_periodic``/``_symmetric`` suffixes honoured,
    _windows.py:2556-2561).  Named windows are cached (read-only arrays): the reference
    calls the path with the same (window, nperseg) for every sweep it plots."""
    if not (isinstance(nperseg, (int, np.integer)) and nperseg > 0):
        raise ValueError(f"Parameter Nx={nperseg} is not a positive integer")
    if isinstance(window, (str, tuple)):
        try:
            key = (window, int(nperseg))
            hit = _WINDOW_CACHE.get(key)
        except TypeError:                    # unhashable tuple entries: SciPy will reject them below
            key, hit = None, None
        if hit is not None:
            return hit
        w = _get_window_uncached(window, nperseg)
        if key is not None:
            if len(_WINDOW_CACHE) > 256:
                _WINDOW_CACHE.clear()
            w.setflags(write=False)
            _WINDOW_CACHE[key] = w
        return w
    return _get_window_uncached(window, nperseg)


def _get_window_uncached(window, nperseg: int) -> np.ndarray:
    if not isinstance(window, (str, tuple)):
        from scipy.signal import get_window as _gw      # float -> kaiser(beta), as SciPy does
        return np.asarray(_gw(window, int(nperseg)), dtype=np.float64)
    name = window if isinstance(window, str) else window[0]
    args = () if isinstance(window, str) else tuple(window[1:])
    if not isinstance(name, str):
        raise ValueError(f"First tuple entry of parameter window={window!r} is not a str!")
    sym = False
    base = name
    if base.endswith("_symmetric"):
        sym, base = True, base[:-10]
    elif base.endswith("_periodic"):
        sym, base = False, base[:-9]
    M = int(nperseg)
    if base in ("tukey", "tuk") and len(args) <= 1:
        return _tukey(M, float(args[0]) if args else 0.5, sym)
    if base in _COSINE_SUMS and not args:
        return _cosine_sum(M, _COSINE_SUMS[base], sym)
    if base in _BOXCAR and not args:
        return np.ones(M, dtype=np.float64)
    from scipy.signal import get_window as _gw
    return np.asarray(_gw(window, M), dtype=np.float64)


def rfftfreq(n: int, fs: float) -> np.ndarray:
    """Bit-exact ``scipy.fft.rfftfreq(n, 1/fs)`` (numpy recipe, reached from
    _spectral_py.py:2303): ``val = 1/(n*d)``; integer arange times val."""
    d = 1 / fs
    val = 1.0 / (n * d)
    return np.arange(0, n // 2 + 1, dtype=int) * val


def time_axis(n_samples: int, nperseg: int, noverlap: int, fs: float) -> np.ndarray:
    """Bit-exact segment times, _spectral_py.py:2324-2325."""
    return np.arange(nperseg / 2, n_samples - nperseg / 2 + 1, nperseg - noverlap) / float(fs)


def frame_count(n_samples: int, nperseg: int, hop: int) -> int:
    """(n - nperseg)//hop + 1 frames (sliding_window_view[..., ::step, :])."""
    return 0 if n_samples < nperseg else (n_samples - nperseg) // hop + 1

"""NumPy restatement of the reference's spectrogram path -- TEST INFRASTRUCTURE ONLY.

The reference computes its spectrogram with one library call,
``spectrogram(data, fs=fs, nperseg=nperseg, scaling="density", mode="psd")``
(/root/reference/PlotEngine.py:113 and :232, imported at :8).  The arithmetic
therefore lives in SciPy (1.18.1 here); this module restates it step by step in
plain NumPy (float64), each function citing the SciPy / reference lines it
follows.  ``SCIPY/`` below means ``site-packages/scipy/``.

It is the checker for the CUDA path and nothing else; see oracle/__init__.py.
"""
from __future__ import annotations

import math
import warnings

import numpy as np

# --------------------------------------------------------------------------
# windows  (SCIPY/signal/windows/_windows.py)
# --------------------------------------------------------------------------


def _extend(M, sym):
    # _windows.py:30-35 -- periodic ("DFT-even") windows are built one sample
    # longer and truncated.
    return (M, False) if sym else (M + 1, True)


def _truncate(w, needed):
    # _windows.py:38-43
    return w[:-1] if needed else w


def _general_cosine(M, a, sym):
    # _windows.py:55-65: linspace(-pi, pi, M) and sum_k a_k cos(k fac)
    if M <= 1:
        return np.ones(M, dtype=np.float64)
    M, trunc = _extend(M, sym)
    fac = np.linspace(-np.pi, np.pi, M, dtype=np.float64)
    w = np.zeros(M, dtype=np.float64)
    for k in range(len(a)):
        w += a[k] * np.cos(k * fac)
    return _truncate(w, trunc)


def _tukey(M, alpha, sym):
    # _windows.py:880-966
    if M <= 1:
        return np.ones(M, dtype=np.float64)
    if alpha <= 0:
        return np.ones(M, dtype=np.float64)
    if alpha >= 1.0:
        return _general_cosine(M, [0.5, 0.5], sym)
    M, trunc = _extend(M, sym)
    n = np.arange(0, M, dtype=np.float64)
    width = int(math.floor(alpha * (M - 1) / 2.0))
    n1 = n[0:width + 1]
    n2 = n[width + 1:M - width - 1]
    n3 = n[M - width - 1:]
    w1 = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * n1 / alpha / (M - 1))))
    w2 = np.ones(n2.shape)
    w3 = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * n3 / alpha / (M - 1))))
    return _truncate(np.concatenate((w1, w2, w3)), trunc)


_COSINE_SUMS = {
    # name -> coefficients (general_hamming(alpha) == [alpha, 1-alpha], :1112-1118)
    "hann": [0.5, 0.5],                                       # :795-876
    "hamming": [0.54, 1.0 - 0.54],                            # :1122-1199
    "blackman": [0.42, 0.50, 0.08],                           # :484-494
    "nuttall": [0.3635819, 0.4891775, 0.1365995, 0.0106411],  # :555-561
    "blackmanharris": [0.35875, 0.48829, 0.14128, 0.01168],   # :610-616
    "flattop": [0.21557895, 0.41663158, 0.277263158, 0.083578947, 0.006947368],  # :680-686
}


def get_window(window, nperseg):
    """Restates ``scipy.signal.get_window(window, Nx, fftbins=True)`` for the
    windows the path uses (_windows.py:2388-2590): a ``_periodic`` /
    ``_symmetric`` suffix overrides ``fftbins``; ``tukey`` takes an optional
    alpha (default 0.5)."""
    name = window if isinstance(window, str) else window[0]
    args = () if isinstance(window, str) else tuple(window[1:])
    sym = False                                   # fftbins=True -> periodic
    if name.endswith("_symmetric"):
        sym, name = True, name[:-10]
    elif name.endswith("_periodic"):
        sym, name = False, name[:-9]
    M = int(nperseg)
    if name in ("tukey", "tuk"):
        return _tukey(M, float(args[0]) if args else 0.5, sym)
    if name in ("hann", "han"):
        return _general_cosine(M, _COSINE_SUMS["hann"], sym)
    if name in ("hamming", "hamm", "ham"):
        return _general_cosine(M, _COSINE_SUMS["hamming"], sym)
    if name in _COSINE_SUMS:
        return _general_cosine(M, _COSINE_SUMS[name], sym)
    if name in ("boxcar", "box", "ones", "rect", "rectangular"):
        return np.ones(M, dtype=np.float64)       # :200-206
    if name in ("bartlett", "bart", "brt"):       # :778-791
        if M <= 1:
            return np.ones(M)
        Mx, trunc = _extend(M, sym)
        n = np.arange(0, Mx, dtype=np.float64)
        w = np.where(n <= (Mx - 1) / 2.0, 2.0 * n / (Mx - 1), 2.0 - 2.0 * n / (Mx - 1))
        return _truncate(w, trunc)
    if name in ("cosine", "halfcosine"):          # :1745-1754
        if M <= 1:
            return np.ones(M)
        Mx, trunc = _extend(M, sym)
        w = np.sin(np.pi / Mx * (np.arange(Mx, dtype=np.float64) + .5))
        return _truncate(w, trunc)
    raise ValueError(f"oracle: window {window!r} not restated")


# --------------------------------------------------------------------------
# the spectrogram itself  (SCIPY/signal/_spectral_py.py)
# --------------------------------------------------------------------------


def triage_segments(window, nperseg, input_length):
    """_spectral_py.py:2400-2461."""
    if isinstance(window, (str, tuple)):
        if nperseg is None:
            nperseg = 256
        if nperseg > input_length:
            warnings.warn(f"nperseg = {nperseg:d} is greater than input length "
                          f" = {input_length:d}, using nperseg = {input_length:d}",
                          stacklevel=3)
            nperseg = input_length
        win = get_window(window, nperseg)
    else:
        win = np.asarray(window)
        if win.ndim != 1:
            raise ValueError("window must be 1-D")
        if input_length < win.shape[-1]:
            raise ValueError("window is longer than input signal")
        if nperseg is None:
            nperseg = win.shape[0]
        elif nperseg != win.shape[0]:
            raise ValueError("value specified for nperseg is different from length of window")
    return win, nperseg


def rfftfreq(n, d):
    """numpy.fft.rfftfreq recipe reached from _spectral_py.py:2303 through
    SCIPY/fft/_helper.py:251-259: ``val = 1/(n*d)``, integer arange times val."""
    val = 1.0 / (n * d)
    return np.arange(0, n // 2 + 1, dtype=int) * val


def time_axis(n_samples, nperseg, noverlap, fs):
    """_spectral_py.py:2324-2325."""
    return np.arange(nperseg / 2, n_samples - nperseg / 2 + 1, nperseg - noverlap) / float(fs)


def frame_count(n_samples, nperseg, hop):
    """sliding_window_view(...)[..., ::step, :] -- _spectral_py.py:2377-2381."""
    if n_samples < nperseg:
        return 0
    return (n_samples - nperseg) // hop + 1


def spectrogram(x, fs=1.0, window=("tukey", .25), nperseg=None, noverlap=None,
                detrend="constant", scaling="density"):
    """float64 restatement of ``scipy.signal.spectrogram(..., mode='psd',
    return_onesided=True, nfft=None, axis=-1)`` for real input.

    Follows _spectral_py.py:1119-1132 (mode check, triage, noverlap default
    nperseg//8), :2209-2230 (validation, nstep), :2255-2279 (detrend, scale),
    :2346-2395 (_fft_helper: frames, detrend, window, rfft), :2313-2341
    (conj*., scale, one-sided doubling, time axis, moveaxis).
    Returns (f, t, Sxx) with Sxx of shape (..., nperseg//2+1, n_frames).
    """
    x = np.asarray(x, dtype=np.float64)
    if nperseg is not None:
        nperseg = int(nperseg)
        if nperseg < 1:
            raise ValueError("nperseg must be a positive integer")
    win, nperseg = triage_segments(window, nperseg, x.shape[-1])
    win = np.asarray(win, dtype=np.float64)
    if noverlap is None:
        noverlap = nperseg // 8                      # :1128-1129
    noverlap = int(noverlap)
    if noverlap >= nperseg:
        raise ValueError("noverlap must be less than nperseg.")
    step = nperseg - noverlap
    nfft = nperseg
    if scaling == "density":
        scale = 1.0 / (fs * (win * win).sum())       # :2274-2275
    elif scaling == "spectrum":
        scale = 1.0 / win.sum() ** 2                 # :2276-2277
    else:
        raise ValueError(f"Unknown scaling: {scaling!r}")
    f = rfftfreq(nfft, 1 / fs)
    # frames: x[j*step : j*step + nperseg]
    frames = np.lib.stride_tricks.sliding_window_view(x, nperseg, axis=-1)[..., ::step, :]
    if detrend == "constant":                        # _signaltools.py:4288-4290
        frames = frames - np.mean(frames, axis=-1, keepdims=True)
    elif detrend == "linear":
        from scipy.signal import detrend as _dt      # pragma: no cover - not on the reference's path
        frames = _dt(frames, type="linear", axis=-1)
    elif detrend not in (False, None):
        raise ValueError("oracle: detrend must be 'constant', 'linear' or False")
    frames = win * frames
    X = np.fft.rfft(frames, n=nfft, axis=-1)
    P = (np.conjugate(X) * X).real * scale
    if nfft % 2:
        P[..., 1:] *= 2
    else:
        P[..., 1:-1] *= 2
    t = time_axis(x.shape[-1], nperseg, noverlap, fs)
    return f, t, np.moveaxis(P, -1, -2)


def mean_spectrogram(x, **kw):
    """Cross-sweep mean.  The reference has no code for it;
    the oracle is the float64 mean of the per-sweep spectrograms."""
    f, t, S = spectrogram(x, **kw)
    return f, t, S.reshape((-1,) + S.shape[-2:]).mean(axis=0)


def to_db(S, floor):
    """10*log10(max(S, floor)) -- the engine's dB epilogue."""
    return 10.0 * np.log10(np.maximum(S, floor))


# --------------------------------------------------------------------------
# the reference's own post-processing around the call (restated; PlotEngine.py
# itself cannot be imported here: PyQt5 / matplotlib / hmmlearn are absent)
# --------------------------------------------------------------------------


def plot_postprocess(f, t, Sxx, fmin, fmax, log_scale, global_max=None):
    """PlotEngine._plot_spectrogram, /root/reference/PlotEngine.py:114-131.

    Returns dict(last_f, last_t, last_Sxx, image) where ``image`` is the
    normalised array handed to pcolormesh (None when the masked Sxx is empty).
    """
    mask = (f >= fmin) & (f <= fmax)
    f, Sxx = f[mask], Sxx[mask, :]
    out = dict(last_f=f.copy(), last_t=t.copy(), last_Sxx=Sxx.copy(), image=None)
    if Sxx.size == 0:
        out["last_t"] = np.array([])
        return out
    base = np.max(Sxx) if global_max is None or global_max <= 0 else global_max
    Sxx_norm = np.clip(Sxx / (base + 1e-20), 0.0, 1.0)
    if log_scale:
        eps = 1e-12
        Sxx_db = 10.0 * np.log10(Sxx_norm + eps)
        Sxx_db = np.nan_to_num(Sxx_db)
        min_db, max_db = np.min(Sxx_db), np.max(Sxx_db)
        Sxx_norm = (Sxx_db - min_db) / (max_db - min_db) if (max_db - min_db) > 1e-6 \
            else np.zeros_like(Sxx_db)
    out["image"] = Sxx_norm
    return out


def band_features(f, t, Sxx, fmin, fmax):
    """PlotEngine._calculate_features, /root/reference/PlotEngine.py:236-242."""
    if Sxx.size == 0:
        return None, None
    freq_mask = (f >= fmin) & (f <= fmax)
    power_feature = np.sum(Sxx[freq_mask, :], axis=0)
    log_power = np.log10(power_feature + 1e-20)
    delta_log_power = np.diff(log_power, prepend=log_power[0])
    return t, np.column_stack([log_power, delta_log_power])


DEFAULT_BANDS = {
    # /root/reference/PlotEngine.py:698-706
    "Delta (δ)": (0, 4), "Theta (θ)": (4, 8),
    "Alpha (α)": (8, 13), "Beta (β)": (13, 30),
    "Gamma (γ)": (30, 80), "HFO (ripples)": (80, 250),
}


def absolute_power(last_Sxx):
    """PlotEngine.calculate_absolute_power, PlotEngine.py:686-690."""
    return None if last_Sxx is None else np.sum(last_Sxx)


def band_powers(last_f, last_Sxx, bands=None):
    """PlotEngine.calculate_band_powers, PlotEngine.py:692-719."""
    if last_Sxx is None or last_f is None:
        return None
    S = np.maximum(0, last_Sxx)
    bands = DEFAULT_BANDS if bands is None else bands
    total = np.sum(S)
    if total < 1e-18:
        return {name: 0.0 for name in bands}
    out = {}
    for name, (low, high) in bands.items():
        m = (last_f >= low) & (last_f < high)
        out[name] = np.clip(np.sum(S[m, :]) / total, 0.0, None)
    return out


def combine_sweeps(signals, fs_list):
    """'Combine all sweeps' -- time concatenation with a segment map,
    /root/reference/PlotEngine.py:162-200 (no averaging)."""
    offset, seg, parts = 0.0, [], []
    for s, fs in zip(signals, fs_list):
        if s is None:
            continue
        dur = len(s) / fs
        seg.append((offset, offset + dur))
        parts.append(s)
        offset += dur
    return (np.concatenate(parts) if parts else None), seg

"""Host logic: SciPy-compatible triage (errors, warnings, defaults), window tables,
axis arrays, frame partitioning -- all without a GPU."""
import re
import warnings

import numpy as np
import pytest
import scipy.signal

import spectrogram_generator_b200 as sg
from spectrogram_generator_b200 import windows
from spectrogram_generator_b200.plot_engine import band_to_bins
from spectrogram_generator_b200.spectrogram import _prepare_input, _result_dtype


def tri(n, **kw):
    a = dict(fs=1.0, window=("tukey", .25), nperseg=None, noverlap=None, nfft=None, detrend="constant",
             return_onesided=True, scaling="density", mode="psd")
    a.update(kw)
    return sg.triage(n, a["fs"], a["window"], a["nperseg"], a["noverlap"], a["nfft"], a["detrend"],
                     a["return_onesided"], a["scaling"], a["mode"])


@pytest.mark.parametrize("window", ["hann", "hamming", "blackman", "boxcar", "flattop", "nuttall",
                                    "blackmanharris", ("tukey", 0.25), "tukey", ("tukey_periodic", .25),
                                    "hann_symmetric", ("kaiser", 8.0), "bartlett", ("gaussian", 30)])
@pytest.mark.parametrize("M", [1, 2, 31, 256, 1000, 4096])
def test_windows_bit_exact_vs_scipy(window, M):
    assert np.array_equal(windows.get_window(window, M), scipy.signal.get_window(window, M))


def test_window_bad_length_raises_like_scipy():
    with pytest.raises(ValueError, match="not a positive integer"):
        windows.get_window("hann", 0)


def test_defaults_follow_reference_call():
    p = tri(40000, fs=20000.0, nperseg=1024)                    # PlotEngine.py:113
    assert (p.nperseg, p.noverlap, p.hop, p.detrend) == (1024, 128, 896, 1)
    assert p.nframes == (40000 - 1024) // 896 + 1 and p.nbins == 513
    w = scipy.signal.get_window(("tukey", .25), 1024)
    assert np.array_equal(p.win64, w) and p.scale == 1.0 / (20000.0 * (w * w).sum())
    assert tri(1000).nperseg == 256                            # SciPy default nperseg
    assert tri(1000, scaling="spectrum", window="hann", nperseg=128).scale == \
        1.0 / scipy.signal.get_window("hann", 128).sum() ** 2


@pytest.mark.parametrize("kw,exc,msg", [
    (dict(mode="foo"), ValueError, "unknown value for mode"),
    (dict(nperseg=0), ValueError, "Parameter Nx=0 is not a positive integer"),
    (dict(nperseg=-4), ValueError, "Parameter Nx=-4 is not a positive integer"),
    (dict(nperseg=64, noverlap=64), ValueError, "noverlap must be less than nperseg"),
    (dict(nperseg=64, nfft=32), ValueError, "nfft must be greater than or equal to nperseg"),
    (dict(scaling="power"), ValueError, "Unknown scaling"),
    (dict(window=np.ones(8), nperseg=16), ValueError, "different from length of window"),
    (dict(window=np.ones((2, 8))), ValueError, "window must be 1-D"),
    (dict(window=np.ones(2000)), ValueError, "window is longer than input signal"),
])
def test_errors_match_scipy(kw, exc, msg):
    x = np.zeros(1000)
    with pytest.raises(exc, match=re.escape(msg)):
        tri(1000, **kw)
    with pytest.raises(exc, match=re.escape(msg)):             # SciPy raises the same
        scipy.signal.spectrogram(x, **kw)


@pytest.mark.parametrize("kw", [dict(mode="complex"), dict(mode="magnitude"), dict(return_onesided=False),
                                dict(detrend="linear"), dict(nperseg=64, nfft=128)])
def test_unimplemented_combinations_raise(kw):
    with pytest.raises(NotImplementedError):
        tri(1000, **kw)


def test_nperseg_clamp_warns_like_scipy():
    with pytest.warns(UserWarning, match="greater than input length"):
        p = tri(100, nperseg=256)
    assert p.nperseg == 100 and p.nframes == 1
    with pytest.raises(ValueError):                            # empty input: get_window(…, 0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tri(0, nperseg=256)


@pytest.mark.parametrize("fs,n,N", [(12345.678, 1024, 30000), (44100.0, 1000, 44100), (1.0, 256, 999),
                                    (20000.0, 512, 40000), (0.37, 96, 5000)])
def test_axes_bit_exact(fs, n, N):
    x = np.zeros(N)
    f, t, _ = scipy.signal.spectrogram(x, fs=fs, nperseg=n)
    assert np.array_equal(f, windows.rfftfreq(n, fs))
    assert np.array_equal(t, windows.time_axis(N, n, n // 8, fs))
    assert len(t) == windows.frame_count(N, n, n - n // 8)


def test_dtype_rule():
    for dt, want in [(np.float64, np.float64), (np.float32, np.float32), (np.int16, np.float32),
                     (np.int32, np.float64), (np.uint8, np.float32), (np.float16, np.float32)]:
        assert _result_dtype(np.zeros(4, dt)) == want
        S = scipy.signal.spectrogram(np.arange(64).astype(dt), nperseg=16)[2]
        assert S.dtype == want
    x, odt, cplx = _prepare_input(np.zeros((5, 3), np.int16), 0)
    assert x.shape == (3, 5) and x.dtype == np.float32 and odt == np.float32 and not cplx


def test_split_frames_and_spans():
    from spectrogram_generator_b200.distributed import sample_span, shard_frames, shard_rows
    parts = sg.split_frames(337497, 8)
    assert sum(c for _, c in parts) == 337497 and parts[0] == (0, 42188) and parts[-1][0] + parts[-1][1] == 337497
    assert [shard_rows(1000, 8, r) for r in (0, 7)] == [(0, 125), (875, 1000)]
    assert [shard_rows(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    f0, c = shard_frames(100, 3, 1)
    lo, hi = sample_span(f0, c, 512, 2048)
    assert (lo, hi) == (f0 * 512, (f0 + c - 1) * 512 + 2048)
    assert sg.split_frames(3, 8)[3:] == [(3, 0)] * 5


def test_band_to_bins_is_the_reference_mask():
    f = windows.rfftfreq(1024, 20000.0)
    for fmin, fmax in [(0.0, 30.0), (10.0, 5000.0), (19.6, 19.6), (0.0, 1e9), (5.0, 4.0)]:
        mask = (f >= fmin) & (f <= fmax)
        b = band_to_bins(f, fmin, fmax)
        if not mask.any():
            assert b is None
        else:
            assert np.array_equal(np.nonzero(mask)[0], np.arange(b[0], b[1] + 1))


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sg.spectrogram(np.zeros(1000, np.float32), nperseg=256)

"""The oracle pinned against the reference's own path (SciPy as the reference
calls it), the committed golden fixtures and SciPy's literal known answers."""
import warnings

import numpy as np
import pytest
import scipy.signal

from oracle import reference_path, stft_oracle
from util import GOLDEN_NAMES, load_golden

RTOL = 1e-12   # float64 restatement vs SciPy: only FFT-backend rounding (pocketfft vs DUCC) differs


def _mk(n, seed=0, dc=0.3):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(n) + dc + np.sin(0.05 * np.arange(n))


@pytest.mark.parametrize("nperseg", [32, 100, 256, 1000, 1024])
def test_oracle_matches_reference_call(nperseg):
    """PlotEngine.py:113 verbatim: spectrogram(data, fs=fs, nperseg=nperseg, scaling='density', mode='psd')."""
    x = _mk(5000)
    f, t, S = reference_path.reference_call(x, 20000.0, nperseg)
    fo, to, So = stft_oracle.spectrogram(x, fs=20000.0, nperseg=nperseg)
    assert np.array_equal(f, fo) and np.array_equal(t, to)
    assert S.shape == So.shape
    assert np.max(np.abs(S - So)) <= RTOL * S.max()


@pytest.mark.parametrize("window", ["hann", "hamming", "blackman", "boxcar", "flattop", "nuttall",
                                    "blackmanharris", "bartlett", "cosine", ("tukey", 0.25),
                                    ("tukey", 0.7), "tukey", ("tukey_periodic", .25), "hann_symmetric"])
@pytest.mark.parametrize("M", [1, 2, 7, 64, 255, 1024])
def test_oracle_windows_bit_exact(window, M):
    assert np.array_equal(stft_oracle.get_window(window, M), scipy.signal.get_window(window, M))


@pytest.mark.parametrize("kw", [
    dict(window="hann", nperseg=1024, noverlap=768),
    dict(window="hann", nperseg=512, noverlap=384, scaling="spectrum"),
    dict(window="hann", nperseg=64, noverlap=32, detrend=False),
    dict(window=("tukey", .25), nperseg=200, noverlap=13),
    dict(window="boxcar", nperseg=33, noverlap=0),
])
def test_oracle_matches_scipy_kwargs(kw):
    x = _mk(4321, seed=3).reshape(1, -1).repeat(3, 0) * np.array([[1.0], [2.0], [-0.5]])
    f, t, S = scipy.signal.spectrogram(x, fs=12345.678, **kw)
    fo, to, So = stft_oracle.spectrogram(x, fs=12345.678, **kw)
    assert np.array_equal(f, fo) and np.array_equal(t, to)
    assert np.max(np.abs(S - So)) <= RTOL * S.max()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_matches_golden(name):
    g = load_golden(name)
    x = g["x"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fo, to, So = stft_oracle.spectrogram(x.astype(np.float64), fs=g["fs"], **g["kw"])
    assert np.array_equal(g["f"], fo) and np.array_equal(g["t"], to)
    tol = 2e-5 if g["Sxx"].dtype == np.float32 else RTOL   # int16 fixture: SciPy ran in float32
    assert np.max(np.abs(g["Sxx"] - So)) <= tol * So.max()
    if "mean" in g:
        assert np.max(np.abs(g["mean"] - stft_oracle.mean_spectrogram(
            x.astype(np.float64), fs=g["fs"], **g["kw"])[2])) <= RTOL * So.max()


def test_scipy_known_answer_welch_vector():
    """scipy/signal/tests/test_spectral.py:246-256 (TestWelch.test_real_onesided_even):
    x = delta[0] + delta[8], N = 16, nperseg = 8 (Hann, 50 % overlap, detrend constant,
    density); the mean over frames of the spectrogram is Welch's estimate."""
    x = np.zeros(16)
    x[0] = 1
    x[8] = 1
    f, t, S = stft_oracle.spectrogram(x, fs=1.0, window="hann", nperseg=8, noverlap=4)
    q = np.array([0.08333333, 0.15277778, 0.22222222, 0.22222222, 0.11111111])
    np.testing.assert_allclose(f, np.linspace(0, 0.5, 5))
    np.testing.assert_allclose(S.mean(axis=-1), q, atol=1e-7, rtol=1e-7)


def test_scipy_shapes_and_clamp():
    """TestSpectrogram (test_spectral.py:967-1021): window_external shapes and the
    nperseg > len(x) clamp with its UserWarning."""
    x = np.random.default_rng(1).standard_normal(1024)
    f, t, S = stft_oracle.spectrogram(x, 10, ("tukey", 0.25), 16, 2)
    assert f.shape == (9,) and S.shape == (9, 73)
    with pytest.warns(UserWarning, match="greater than input length"):
        f2, t2, S2 = stft_oracle.spectrogram(x[:8], 10, "hann", 1024)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f3, t3, S3 = scipy.signal.spectrogram(x[:8], 10, "hann", 1024)
    assert np.array_equal(f2, f3) and np.allclose(S2, S3, rtol=1e-12, atol=1e-18)


def test_axis_recipes_bit_exact_odd_rates():
    for fs, n in [(12345.678, 1024), (44100.0, 1000), (1.0, 256), (20000.0, 513), (0.1, 96)]:
        assert np.array_equal(stft_oracle.rfftfreq(n, 1 / fs), np.fft.rfftfreq(n, 1 / fs))
        x = np.zeros(3 * n + 7)
        _, t, _ = scipy.signal.spectrogram(x, fs=fs, nperseg=n)
        assert np.array_equal(t, stft_oracle.time_axis(len(x), n, n // 8, fs))


def test_postprocess_and_features_follow_reference_lines():
    g = load_golden("ref_call_256")
    settings = dict(nperseg=256, fmin=0.0, fmax=3000.0, log_scale=True)
    out = reference_path.plot_spectrogram_compute(g["x"].astype(np.float64), g["fs"], settings)
    assert out["last_Sxx"].shape[0] == int(np.sum((g["f"] >= 0) & (g["f"] <= 3000.0)))
    assert out["image"].min() == 0.0 and out["image"].max() == 1.0
    t, feat = reference_path.calculate_features(g["x"].astype(np.float64), g["fs"], settings)
    assert feat.shape == (len(g["t"]), 2) and feat[0, 1] == 0.0
    bp = stft_oracle.band_powers(out["last_f"], out["last_Sxx"])
    assert set(bp) == set(stft_oracle.DEFAULT_BANDS) and all(v >= 0 for v in bp.values())
    empty = reference_path.plot_spectrogram_compute(g["x"].astype(np.float64), g["fs"],
                                                    dict(nperseg=256, fmin=1e6, fmax=2e6, log_scale=False))
    assert empty["image"] is None and empty["last_t"].size == 0

"""Pins the restated post-processing (oracle/stft_oracle.py: plot_postprocess, band_features,
band_powers, absolute_power) to the reference's OWN code: PlotEngine._plot_spectrogram,
_calculate_features, calculate_absolute_power and calculate_band_powers are executed from
/root/reference/PlotEngine.py (oracle/plot_engine_ref.py loads the file with inert stand-ins for
PyQt5 / matplotlib / hmmlearn and calls the unbound methods on a recording `self`) -- live where
the reference tree exists (the build container), and through the fixture those same calls
produced (tests/golden/plot_engine_ref.npz, tests/golden/make_plot_engine_golden.py) everywhere."""
import ast
import os

import numpy as np
import pytest
import scipy.signal

from oracle import plot_engine_ref as R
from oracle import stft_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "plot_engine_ref.npz")
CASES = ["gui_default", "gui_default_log", "wide_band_log", "empty_band"]


def load_case(name):
    z = np.load(GOLD)
    d = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + ".")}
    d["settings"] = ast.literal_eval(str(d["settings"]))
    d["fs"] = float(d["fs"])
    return d


def restated(x64, fs, settings):
    """The same three computations through the restatement, from the same SciPy call (PlotEngine.py:113)."""
    f, t, Sxx = scipy.signal.spectrogram(x64, fs=fs, nperseg=settings["nperseg"], scaling="density", mode="psd")
    post = stft_oracle.plot_postprocess(f, t, Sxx, settings["fmin"], settings["fmax"], settings["log_scale"])
    feat = stft_oracle.band_features(np.asarray(f), t, np.asarray(Sxx), settings["fmin"], settings["fmax"])
    return post, feat


@pytest.mark.parametrize("name", CASES)
def test_restatement_equals_the_fixture_made_by_the_reference(name):
    g = load_case(name)
    post, (ft, feat) = restated(g["x"].astype(np.float64), g["fs"], g["settings"])
    # same NumPy operations in the same order: identical bits
    assert np.array_equal(post["last_f"], g["last_f"]) and np.array_equal(post["last_t"], g["last_t"])
    assert np.array_equal(post["last_Sxx"], g["last_Sxx"])
    assert (post["image"] is not None) == bool(g["has_image"])
    if post["image"] is not None:
        assert np.array_equal(post["image"], g["image"])
    assert np.array_equal(ft, g["feat_t"]) and np.array_equal(feat, g["features"])
    assert stft_oracle.absolute_power(post["last_Sxx"]) == g["absolute_power"]
    bp = stft_oracle.band_powers(post["last_f"], post["last_Sxx"])
    assert list(bp.keys()) == [str(s) for s in g["band_names"]]
    assert np.array_equal(np.array([float(v) for v in bp.values()]), g["band_powers"])


@pytest.mark.skipif(not R.available(), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("seed,nperseg,fmin,fmax,log_scale,global_max", [
    (0, 1024, 0.0, 30.0, False, None), (1, 256, 5.0, 200.0, True, None), (2, 512, 0.0, 500.0, True, 1e-3),
    (3, 64, 0.0, 1e9, False, -1.0), (4, 4096, 0.0, 30.0, True, None), (5, 128, 450.0, 460.0, False, None),
    (6, 300, 0.0, 100.0, True, None)])
def test_restatement_equals_the_reference_run_live(seed, nperseg, fmin, fmax, log_scale, global_max):
    PE = R.load_plot_engine()
    rng = np.random.default_rng(seed)
    fs = 1000.0
    n = 9000
    x = rng.standard_normal(n) * 0.3 + np.sin(2 * np.pi * 12.0 * np.arange(n) / fs) - 65.0
    settings = dict(nperseg=nperseg, fmin=fmin, fmax=fmax, log_scale=log_scale)
    ref = R.plot_spectrogram(PE, x, fs, settings, global_max)
    f, t, Sxx = scipy.signal.spectrogram(x, fs=fs, nperseg=nperseg, scaling="density", mode="psd")
    post = stft_oracle.plot_postprocess(f, t, Sxx, fmin, fmax, log_scale, global_max)
    for k in ("last_f", "last_t", "last_Sxx"):
        assert np.array_equal(post[k], ref[k]), k
    assert (post["image"] is None) == (ref["image"] is None)
    if post["image"] is not None:
        assert np.array_equal(post["image"], ref["image"])
    rt, rfeat = R.calculate_features(PE, x, fs, settings)
    ot, ofeat = stft_oracle.band_features(np.asarray(f), t, np.asarray(Sxx), fmin, fmax)
    assert np.array_equal(rt, ot) and np.array_equal(rfeat, ofeat)
    total, bands = R.power_summaries(PE, ref["state"])
    assert total == stft_oracle.absolute_power(post["last_Sxx"])
    ob = stft_oracle.band_powers(post["last_f"], post["last_Sxx"])
    assert list(bands.keys()) == list(ob.keys())
    assert all(float(bands[k]) == float(ob[k]) for k in bands)


@pytest.mark.skipif(not R.available(), reason="the reference tree only exists in the build container")
def test_reference_call_site_is_the_scipy_call_the_oracle_follows():
    """The loaded module's `spectrogram` IS scipy.signal.spectrogram (PlotEngine.py:8), called with
    fs, nperseg, scaling='density', mode='psd' and nothing else (PlotEngine.py:113)."""
    PE = R.load_plot_engine()
    assert PE._plot_spectrogram.__globals__["spectrogram"] is scipy.signal.spectrogram
    src = open(R.REFERENCE_FILE).read()
    assert 'spectrogram(data, fs=fs, nperseg=nperseg, scaling="density", mode="psd")' in src
    assert "spectrogram(signal, fs=fs, nperseg=settings['nperseg'], scaling=\"density\", mode=\"psd\")" in src

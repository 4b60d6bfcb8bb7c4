"""Drop-in check: the reference's PlotEngine post-processing (restated from
PlotEngine.py:110-131, :229-242, :686-719 in oracle/) run once over SciPy's result
and once over the engine's, with the same settings dict the GUI builds
(GUI.py:421-431)."""
import numpy as np
import pytest

import spectrogram_generator_b200 as sg
from oracle import reference_path, stft_oracle
from util import assert_parity

pytestmark = pytest.mark.gpu


def sweep(n=40000, fs=20000.0, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    burst = (np.sin(2 * np.pi * 18.0 * t) * (np.abs(t - 1.0) < 0.3)).astype(np.float64)
    return (0.05 * rng.standard_normal(n) + 0.4 * np.sin(2 * np.pi * 7.0 * t) + burst - 0.065).astype(np.float32), fs


@pytest.mark.parametrize("settings", [
    dict(nperseg=1024, fmin=0.0, fmax=30.0, log_scale=False),       # GUI defaults (GUI.py:218-221)
    dict(nperseg=1024, fmin=0.0, fmax=30.0, log_scale=True),
    dict(nperseg=4096, fmin=1.0, fmax=250.0, log_scale=True),
    dict(nperseg=256, fmin=100.0, fmax=5000.0, log_scale=False),
])
def test_plot_spectrogram_matches_reference_postprocessing(settings):
    x, fs = sweep()
    ref = reference_path.plot_spectrogram_compute(x.astype(np.float64), fs, settings)
    path = sg.SpectrogramPath()
    img = path._plot_spectrogram(x, fs, settings)
    assert np.array_equal(path.last_f, ref["last_f"]) and np.array_equal(path.last_t, ref["last_t"])
    assert_parity(path.last_Sxx, ref["last_Sxx"], what="last_Sxx")
    assert img.shape == ref["image"].shape
    if settings["log_scale"]:
        # min-max normalised dB image over a ~120 dB range: 1e-3 dB <-> ~1e-5 of full scale
        assert np.max(np.abs(img - ref["image"])[ref["last_Sxx"] >= 1e-6 * ref["last_Sxx"].max()]) <= 2e-5
    else:
        assert np.max(np.abs(img - ref["image"])) <= 1e-4
    np.testing.assert_allclose(path.calculate_absolute_power(), stft_oracle.absolute_power(ref["last_Sxx"]), rtol=1e-5)
    bp, bpr = path.calculate_band_powers(), stft_oracle.band_powers(ref["last_f"], ref["last_Sxx"])
    for k in bpr:
        assert abs(bp[k] - bpr[k]) <= 1e-5


@pytest.mark.parametrize("name", ["gui_default", "gui_default_log", "wide_band_log", "empty_band"])
def test_against_the_fixture_the_reference_itself_produced(name):
    """tests/golden/plot_engine_ref.npz holds what the reference's OWN PlotEngine methods returned
    (executed from /root/reference/PlotEngine.py by tests/golden/make_plot_engine_golden.py) for float32
    sweeps on a -70 baseline: last_f / last_t / last_Sxx, the image handed to pcolormesh, the HMM features
    and the power summaries.  The engine's SpectrogramPath is held to the same bar as everywhere."""
    from test_reference_plot_engine import load_case
    g = load_case(name)
    path = sg.SpectrogramPath()
    img = path._plot_spectrogram(g["x"], g["fs"], g["settings"])
    assert np.array_equal(path.last_f, g["last_f"]) and np.array_equal(path.last_t, g["last_t"])
    assert path.last_Sxx.shape == g["last_Sxx"].shape
    if not bool(g["has_image"]):
        assert img is None and path.calculate_absolute_power() == 0
        return
    assert_parity(path.last_Sxx, g["last_Sxx"], what=f"{name} last_Sxx")
    big = g["last_Sxx"] >= 1e-6 * g["last_Sxx"].max()
    tol = 2e-5 if g["settings"]["log_scale"] else 1e-4
    assert np.max(np.abs(img - g["image"])[big]) <= tol
    # power summaries: reduced on the device (b2s_band_sums_f32)
    np.testing.assert_allclose(path.calculate_absolute_power(), g["absolute_power"], rtol=1e-5)
    bp = path.calculate_band_powers()
    assert list(bp.keys()) == [str(s) for s in g["band_names"]]
    np.testing.assert_allclose([float(v) for v in bp.values()], g["band_powers"], rtol=2e-5, atol=1e-12)
    # and they equal the host reduction of the copied-back array
    host = np.maximum(0, path.last_Sxx).astype(np.float64)
    for (low, high), v in zip(sg.plot_engine.DEFAULT_BANDS.values(), bp.values()):
        m = (path.last_f >= low) & (path.last_f < high)
        np.testing.assert_allclose(float(v), host[m].sum() / host.sum(), rtol=1e-12, atol=1e-15)
    # features (PlotEngine.py:229-242) through the fused band-power epilogue
    path.last_fs, path.last_settings = g["fs"], g["settings"]
    t, feat = path._calculate_features(g["x"])
    assert np.array_equal(t, g["feat_t"]) and feat.shape == g["features"].shape
    assert np.max(np.abs(feat - g["features"])) <= 1e-4 / np.log(10) * 2


def test_empty_band_mask_follows_reference_early_return():
    x, fs = sweep()
    path = sg.SpectrogramPath()
    assert path._plot_spectrogram(x, fs, dict(nperseg=1024, fmin=1e6, fmax=2e6, log_scale=False)) is None
    assert path.last_t.size == 0 and path.last_Sxx.size == 0


def test_calculate_features_matches_reference():
    x, fs = sweep(seed=2)
    settings = dict(nperseg=1024, fmin=0.0, fmax=30.0, log_scale=False)
    tr, fr = reference_path.calculate_features(x.astype(np.float64), fs, settings)
    path = sg.SpectrogramPath()
    path.last_fs, path.last_settings = fs, settings
    t, feat = path._calculate_features(x)
    assert np.array_equal(t, tr) and feat.shape == fr.shape
    assert np.max(np.abs(feat - fr)) <= 1e-4 / np.log(10) * 2      # log10 of a sum within 1e-4 relative


def test_plot_extra_source_selection_and_combine():
    a, fs = sweep(seed=3)
    b, _ = sweep(seed=4)
    path = sg.SpectrogramPath()
    settings = dict(nperseg=512, fmin=0.0, fmax=100.0, log_scale=False, mode_proc="None", mode_raw="Both",
                    draw_proc=False, draw_raw=True)
    infos = [dict(signal_raw=a, signal_proc=None, fs=fs, item="s0"), dict(signal_raw=b, signal_proc=None, fs=fs, item="s1")]
    cat = path.combine(infos, settings)
    want, seg = stft_oracle.combine_sweeps([a, b], [fs, fs])
    assert np.array_equal(cat, want)
    assert [(s["start_time_combined"], s["end_time_combined"]) for s in path.segment_map] == seg
    img = path.plot_extra(cat, None, fs, settings)
    ref = reference_path.plot_spectrogram_compute(want.astype(np.float64), fs, settings)
    assert path.spec_data_source is cat and np.array_equal(path.last_t, ref["last_t"])
    assert_parity(path.last_Sxx, ref["last_Sxx"])
    assert np.max(np.abs(img - ref["image"])) <= 1e-4


def test_combine_assembles_the_sweeps_on_the_device():
    """combine() copies the sweeps one by one to their offsets in one device buffer; the
    spectrogram / features of the array it returns use that buffer (no second upload of the concatenated
    signal) and are bit-identical to the plain upload.  Sweeps of unequal length, an odd total, mixed
    float32 / float64 (np.concatenate promotes), and a stale buffer is never used for another array."""
    import torch
    a, fs = sweep(seed=5)
    b, _ = sweep(seed=6)
    c, _ = sweep(seed=7)
    settings = dict(nperseg=256, fmin=1.0, fmax=300.0, log_scale=True, mode_proc="None", mode_raw="Both",
                    draw_proc=False, draw_raw=True)
    for parts in ([a, b[:-1777], c[:12345]], [a.astype(np.float64), b, c[:999]]):
        infos = [dict(signal_raw=q, signal_proc=None, fs=fs, item=i) for i, q in enumerate(parts)]
        path = sg.SpectrogramPath()
        cat = path.combine(infos, settings)
        want, seg = stft_oracle.combine_sweeps(parts, [fs] * len(parts))
        assert np.array_equal(cat, want) and cat.dtype == want.dtype
        assert path._combined is not None and path._combined[0] is cat
        assert torch.equal(path._combined[1].cpu(), torch.from_numpy(want))
        img = path.plot_extra(cat, None, fs, settings)
        t1, feat1 = path._calculate_features(cat, fs, settings)
        plain = sg.SpectrogramPath()                         # the same array without the device copy
        img0 = plain.plot_extra(want.copy(), None, fs, settings)
        t0, feat0 = plain._calculate_features(want.copy(), fs, settings)
        assert np.array_equal(path.last_Sxx, plain.last_Sxx) and np.array_equal(img, img0)
        assert np.array_equal(t1, t0) and np.array_equal(feat1, feat0)
        ref = reference_path.plot_spectrogram_compute(want.astype(np.float64), fs, settings)
        assert_parity(path.last_Sxx, ref["last_Sxx"])
        # another array of the same length does not pick the staged buffer up
        other = want[::-1].copy()
        path.plot_extra(other, None, fs, settings)
        plain.plot_extra(other.copy(), None, fs, settings)
        assert np.array_equal(path.last_Sxx, plain.last_Sxx)


def test_display_scale_kernel_matches_reference_lines():
    import torch
    rng = np.random.default_rng(11)
    S = (rng.random((200, 37)) ** 8).astype(np.float32) * 3e-4
    S[5, 5] = 0.0
    eng = sg.engine()
    Sd = torch.from_numpy(S).cuda()
    for log_scale in (False, True):
        for gmax in (None, 1e-4, 5.0):
            want = stft_oracle.plot_postprocess(np.array([1.0]), np.array([0.0]), S.reshape(1, -1), 0.0, 2.0,
                                                log_scale, gmax)["image"].reshape(S.shape)
            got = eng.display_scale(Sd, log_scale, gmax).cpu().numpy()
            assert np.max(np.abs(got - want)) <= 2e-6, (log_scale, gmax)
    const = torch.full((10, 10), 0.25, device="cuda")
    assert torch.equal(eng.display_scale(const, True), torch.zeros_like(const))      # dB range <= 1e-6 -> zeros
    assert torch.equal(eng.display_scale(const, False), torch.ones_like(const))


def test_band_power_entry_equals_cropped_sum():
    import torch
    x, fs = sweep(seed=5)
    plan = sg.triage(len(x), fs, ("tukey", .25), 1024, None, None, "constant", True, "density", "psd")
    eng = sg.engine()
    xd = torch.from_numpy(x).cuda().view(1, -1)
    full = eng.stft_psd(xd, plan)
    for kmin, kmax in [(0, 1), (0, 512), (3, 40), (512, 512)]:
        band = eng.band_power(xd, plan, kmin, kmax)
        torch.testing.assert_close(band, full[:, :, kmin:kmax + 1].sum(dim=-1), rtol=3e-6, atol=0)
    part = eng.band_power(xd, plan, 3, 40, frame0=4, nframes=9)
    assert torch.equal(part, eng.band_power(xd, plan, 3, 40)[:, 4:13])

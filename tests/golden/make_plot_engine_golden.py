"""Generates tests/golden/plot_engine_ref.npz by EXECUTING the reference's own PlotEngine methods
(/root/reference/PlotEngine.py:110-145, 229-242, 686-719) through oracle/plot_engine_ref.py.

Run in the build container, where /root/reference exists:  python tests/golden/make_plot_engine_golden.py
The fixture travels to the GPU box (which has no /root/reference): it pins the restatement in
oracle/stft_oracle.py (tests/test_reference_plot_engine.py) and is what the GPU drop-in tests compare the
engine's SpectrogramPath against (tests/test_gpu_dropin.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import plot_engine_ref as R  # noqa: E402

CASES = {
    # name: (n, fs, settings) -- float32 sweeps as pyabf delivers them (SweepManager.py:31), an
    # electrophysiology-like baseline, the GUI's default settings (GUI.py:213-221) and a wide band
    "gui_default": (30000, 1000.0, dict(nperseg=1024, fmin=0.0, fmax=30.0, log_scale=False)),
    "gui_default_log": (30000, 1000.0, dict(nperseg=1024, fmin=0.0, fmax=30.0, log_scale=True)),
    "wide_band_log": (40000, 20000.0, dict(nperseg=512, fmin=100.0, fmax=5000.0, log_scale=True)),
    "empty_band": (8000, 1000.0, dict(nperseg=256, fmin=600.0, fmax=700.0, log_scale=False)),
}


def signal(name, n, fs):
    rng = np.random.default_rng(sum(map(ord, name)))
    t = np.arange(n) / fs
    x = 0.3 * rng.standard_normal(n) + np.sin(2 * np.pi * 9.5 * t) + 0.4 * np.sin(2 * np.pi * 0.061 * fs * t) - 70.0
    return x.astype(np.float32)


def main():
    PE = R.load_plot_engine()
    out = {}
    for name, (n, fs, settings) in CASES.items():
        x = signal(name, n, fs)
        r = R.plot_spectrogram(PE, x.astype(np.float64), fs, settings)
        out[f"{name}.x"] = x
        out[f"{name}.fs"] = fs
        out[f"{name}.settings"] = np.array(repr(settings))
        out[f"{name}.last_f"] = r["last_f"]
        out[f"{name}.last_t"] = r["last_t"]
        out[f"{name}.last_Sxx"] = r["last_Sxx"]
        out[f"{name}.has_image"] = r["image"] is not None
        if r["image"] is not None:
            out[f"{name}.image"] = r["image"]
        t, feat = R.calculate_features(PE, x.astype(np.float64), fs, settings)
        out[f"{name}.feat_t"] = t
        out[f"{name}.features"] = feat
        total, bands = R.power_summaries(PE, r["state"])
        out[f"{name}.absolute_power"] = total
        out[f"{name}.band_names"] = np.array(list(bands.keys()))
        out[f"{name}.band_powers"] = np.array([float(v) for v in bands.values()])
        print(name, r["last_Sxx"].shape, None if r["image"] is None else r["image"].shape, feat.shape)
    np.savez_compressed(os.path.join(HERE, "plot_engine_ref.npz"), **out)


if __name__ == "__main__":
    main()

"""Generates tests/golden/*.npz from the reference's own spectrogram path.

Run in the build container (SciPy 1.18.1):  python tests/golden/make_golden.py
Each fixture stores the seeded input (float32 / int16) and the outputs of
``scipy.signal.spectrogram`` called the way the reference calls it
(PlotEngine.py:113: ``spectrogram(data, fs=fs, nperseg=nperseg, scaling="density",
mode="psd")``) or, for the BASELINE configs, the same entry with explicit
window / noverlap.  The float64 outputs are computed from the float32 samples
upcast to float64 -- the values the GPU path is compared against.
"""
import os
import sys

import numpy as np
import scipy
from scipy.signal import spectrogram

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from spectrogram_generator_b200 import synth  # noqa: E402


def save(name, x, fs, kw, **extra):
    f, t, S = spectrogram(x.astype(np.float64) if x.dtype != np.int16 else x, fs=fs,
                          scaling="density", mode="psd", **kw)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, fs=fs, f=f, t=t, Sxx=S,
                        kw=np.array(repr(kw)), scipy=np.array(scipy.__version__), **extra)
    print(name, x.shape, x.dtype, "->", S.shape, S.dtype)


def main():
    rng = np.random.default_rng(20261018)
    # 1. the reference's literal call: SciPy defaults (Tukey .25, noverlap nperseg//8, detrend)
    fs = 20000.0
    n = 6000
    x = (0.2 * rng.standard_normal(n) + np.sin(2 * np.pi * 1234.5 * np.arange(n) / fs) - 0.07).astype(np.float32)
    save("ref_call_256", x, fs, dict(nperseg=256))
    save("ref_call_1024", x, fs, dict(nperseg=1024))
    # 2. config 1 (shortened to 0.25 s): chirp, Hann 1024/256
    x1, kw1 = synth.config1(seconds=0.25)
    fs1 = kw1.pop("fs")
    save("c1_chirp_025s", x1, fs1, kw1)
    # 3. config 2 (4 sweeps x 0.2 s): Hann 512/128, plus the float64 cross-sweep mean
    x2, kw2 = synth.config2(batch=4, seconds=0.2)
    fs2 = kw2.pop("fs")
    S2 = spectrogram(x2.astype(np.float64), fs=fs2, scaling="density", mode="psd", **kw2)[2]
    save("c2_sweeps_4x02s", x2, fs2, kw2, mean=S2.mean(axis=0))
    # 4. int16 samples (SciPy's float32 pipeline: result_type(int16, complex64) == complex64)
    xi = np.round(3000 * np.sin(2 * np.pi * 50.0 * np.arange(4096) / 1000.0)
                  + 200 * rng.standard_normal(4096)).astype(np.int16)
    save("int16_512", xi, 1000.0, dict(nperseg=512))
    # 5. nperseg > len(x): SciPy clamps nperseg to len(x) (here 64 -> power of two)
    xs = rng.standard_normal(64).astype(np.float32)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        save("clamp_64", xs, 1.0, dict(nperseg=1024))


if __name__ == "__main__":
    main()

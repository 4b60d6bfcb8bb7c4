#!/usr/bin/env python
"""Single-call latency of the drop-in spectrogram() on BASELINE config 1 (one 10 s mono 44.1 kHz
chirp, nperseg 1024, hop 256, Hann) and on the reference's own call form, against SciPy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.signal
import spectrogram_generator_b200 as sg
from spectrogram_generator_b200 import synth

x, kw = synth.config1()
fs = kw.pop("fs")
x32 = x.astype(np.float32)
x64 = x.astype(np.float64)

def med(fn, n=30):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))

print("C1 explicit kwargs  b200 (float32 in): %.3f ms" % med(lambda: sg.spectrogram(x32, fs=fs, **kw)))
print("C1 explicit kwargs  b200 (float64 in): %.3f ms" % med(lambda: sg.spectrogram(x64, fs=fs, **kw)))
print("C1 explicit kwargs  scipy (float64)  : %.3f ms" % med(lambda: scipy.signal.spectrogram(x64, fs=fs, **kw), 5))
print("reference call form b200 (float32 in): %.3f ms" % med(lambda: sg.spectrogram(x32, fs=fs, nperseg=1024, scaling="density", mode="psd")))
print("reference call form scipy (float64)  : %.3f ms" % med(lambda: scipy.signal.spectrogram(x64, fs=fs, nperseg=1024, scaling="density", mode="psd"), 5))
xs = x32[:40000]
print("2 s sweep @20k, reference call form b200: %.3f ms" % med(lambda: sg.spectrogram(xs, fs=20000.0, nperseg=1024, scaling="density", mode="psd")))
print("2 s sweep @20k, reference call form scipy: %.3f ms" % med(lambda: scipy.signal.spectrogram(xs.astype(np.float64), fs=20000.0, nperseg=1024, scaling="density", mode="psd"), 10))

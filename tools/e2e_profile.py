#!/usr/bin/env python
"""cProfile of the end-to-end call of bench.py's e2e leg (host-side overhead hunting)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectrogram_generator_b200 as sg
from spectrogram_generator_b200 import synth

x, kw = synth.config2(batch=1000, seed=1234)
fs = kw.pop("fs")
xp = sg.pinned_empty(x.shape, np.float32)
xp[...] = x
out = sg.pinned_empty((1000, 309, 257), np.float32)
call = lambda: sg.mean_spectrogram(xp, fs=fs, return_per_sweep=True, out=out, **kw)
for _ in range(5):
    call()
torch.cuda.synchronize()
# time to the first enqueue and after the last sync: wall-clock stamps around the pieces
t0 = time.perf_counter()
for _ in range(20):
    call()
torch.cuda.synchronize()
print("ms per call", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    call()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)

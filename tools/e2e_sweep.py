#!/usr/bin/env python
"""End-to-end (pinned host in, NumPy out) time of mean_spectrogram on the C2 batch for a few
pipeline chunk sizes (tuning tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectrogram_generator_b200 as sg
from spectrogram_generator_b200 import spectrogram as sp, synth

x, kw = synth.config2(batch=1000, seed=1234)
fs = kw.pop("fs")
xp = sg.pinned_empty(x.shape, np.float32)
xp[...] = x
for mb in (4, 8, 16, 32, 64):
    sp._PIPE_CHUNK_BYTES = mb << 20
    res = None
    for _ in range(3):
        res = sg.mean_spectrogram(xp, fs=fs, return_per_sweep=True, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 8
    for _ in range(n):
        res = sg.mean_spectrogram(xp, fs=fs, return_per_sweep=True, **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"chunk {mb:3d} MB: {dt * 1e3:7.3f} ms/step  {x.size / dt / 1e9:6.2f} Gsamples/s")

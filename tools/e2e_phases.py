#!/usr/bin/env python
"""Where the end-to-end time of the public API goes (H2D, kernels, D2H, host)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrogram_generator_b200 as sg
from spectrogram_generator_b200 import synth

x, kw = synth.config2(batch=1000)
fs = kw.pop("fs")
xp = sg.pinned_empty(x.shape, np.float32); xp[...] = x
def T(f, n=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
h = torch.from_numpy(xp); print("pinned:", h.is_pinned())
d = h.cuda(); out_h = torch.empty((1000, 309, 257), pin_memory=True); out_d = torch.empty((1000, 309, 257), device="cuda")
print("H2D 160MB pinned  ms", T(lambda: d.copy_(h, non_blocking=True)))
print("H2D 160MB pageable ms", T(lambda: d.copy_(torch.from_numpy(x))))
print("D2H 318MB pinned  ms", T(lambda: out_h.copy_(out_d, non_blocking=True)))
print("alloc pinned 318MB ms", T(lambda: torch.empty((1000, 309, 257), pin_memory=True)))
print("api mean+per_sweep ms", T(lambda: sg.mean_spectrogram(xp, fs=fs, return_per_sweep=True, **kw)))
print("api spectrogram    ms", T(lambda: sg.spectrogram(xp, fs=fs, **kw)))
print("api spectrogram pageable ms", T(lambda: sg.spectrogram(x, fs=fs, **kw)))

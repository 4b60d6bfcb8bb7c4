#!/usr/bin/env python
"""Where the C2 step goes: STFT kernel, cross-sweep sum, both back to back (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectrogram_generator_b200 as sg
from spectrogram_generator_b200 import synth

x_host, kw = synth.config2(batch=1000, seed=1234)
fs = kw.pop("fs")
plan = sg.triage(x_host.shape[1], fs, kw["window"], kw["nperseg"], kw["noverlap"], None, "constant", True, "density", "psd")
eng = sg.engine()
x = torch.from_numpy(x_host).cuda()
S = torch.empty((1000, plan.nframes, plan.nbins), dtype=torch.float32, device="cuda")

def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n

print("stft only      %.4f ms" % timeit(lambda: eng.stft_psd(x, plan, out=S)))
print("sum only       %.4f ms" % timeit(lambda: eng.batch_sum(S, 1e-3)))
print("stft + sum     %.4f ms" % timeit(lambda: (eng.stft_psd(x, plan, out=S), eng.batch_sum(S, 1e-3))))

import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrogram_generator_b200 as sg
from spectrogram_generator_b200 import synth
x, kw = synth.config2(batch=1000); fs = kw.pop("fs")
xp = sg.pinned_empty(x.shape, np.float32); xp[...] = x
def one():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = sg.mean_spectrogram(xp, fs=fs, return_per_sweep=True, **kw)
    torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
print("drop results:", [round(one()[1], 1) for _ in range(6)])
held = None; ts = []
for _ in range(8):
    held, t = one(); ts.append(round(t, 1))
print("hold last result:", ts)

#!/usr/bin/env python
"""End-to-end leg of bench.py (pinned host in, caller-owned pinned result out) on the C2 batch: the API call for
a few pipeline chunk sizes, beside copy-only pipelines with the same chunking (no kernels) and the plain
both-direction copy -- where the 10 % between the call and the PCIe bound go (tuning tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectrogram_generator_b200 as sg
from spectrogram_generator_b200 import spectrogram as sp, synth

x, kw = synth.config2(batch=1000, seed=1234)
fs = kw.pop("fs")
xp = sg.pinned_empty(x.shape, np.float32)
xp[...] = x
B, n = x.shape
F, K = 309, 257
out = sg.pinned_empty((B, K, F) if False else (B, F, K), np.float32)


def timeit(fn, reps=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


h_in, h_out = torch.from_numpy(xp), torch.from_numpy(out)
d_in = torch.empty((B, n), device="cuda")
d_out = torch.empty((B, F, K), device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def raw_both():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def chunks(step, ramp=True):
    b, cur = 0, max(1, step // 16) if ramp else step
    while b < B:
        yield b, min(B, b + cur)
        b += cur
        cur = min(step, cur * 2)


def copy_pipeline(step, ramp=True):
    for b0, b1 in chunks(step, ramp):
        with torch.cuda.stream(s1):
            d_in[b0:b1].copy_(h_in[b0:b1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s1)
        with torch.cuda.stream(s2):
            s2.wait_event(ev)
            h_out[b0:b1].copy_(d_out[b0:b1], non_blocking=True)


print("H2D alone ms", round(timeit(lambda: d_in.copy_(h_in, non_blocking=True)), 3))
print("D2H alone ms", round(timeit(lambda: h_out.copy_(d_out, non_blocking=True)), 3))
print("both directions, one copy each ms", round(timeit(raw_both), 3))
row_bytes = F * K * 4
for mb in (8, 16, 32, 64, 128):
    step = max(1, (mb << 20) // row_bytes)
    t_copy = timeit(lambda: copy_pipeline(step))
    sp._PIPE_CHUNK_BYTES = mb << 20
    t_api = timeit(lambda: sg.mean_spectrogram(xp, fs=fs, return_per_sweep=True, out=out, **kw))
    print(f"chunk {mb:3d} MB: copy-only pipeline {t_copy:6.3f} ms   api {t_api:6.3f} ms   {x.size / t_api / 1e6:6.2f} Gsamples/s")

sp._PIPE_CHUNK_BYTES = 32 << 20
call = lambda: sg.mean_spectrogram(xp, fs=fs, return_per_sweep=True, out=out, **kw)
for rnd in range(2):
    for div, growth in [(16, 2.0), (16, 1.5), (16, 1.25), (32, 1.5), (8, 1.5), (16, 1.5), (16, 2.0)]:
        sp._PIPE_RAMP_DIV, sp._PIPE_RAMP_GROWTH = div, growth
        print(f"round {rnd} ramp 1/{div:2d} x {growth}: api {timeit(call, reps=12):6.3f} ms")
for mb, div, growth in [(16, 16, 1.5), (64, 32, 1.5), (32, 16, 2.0)]:
    sp._PIPE_CHUNK_BYTES, sp._PIPE_RAMP_DIV, sp._PIPE_RAMP_GROWTH = mb << 20, div, growth
    print(f"chunk {mb} MB ramp 1/{div} x {growth}: api {timeit(call, reps=12):6.3f} ms")

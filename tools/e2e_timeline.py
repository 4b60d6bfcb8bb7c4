#!/usr/bin/env python
"""GPU-side timeline of one end-to-end call of the C2 batch (the loop of spectrogram._host_pipeline, re-enacted with
timing events on every stream): when each chunk's H2D, kernel and D2H start and end, where the copy engines idle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectrogram_generator_b200 as sg
from spectrogram_generator_b200 import synth
import importlib
sp = importlib.import_module("spectrogram_generator_b200.spectrogram")

x, kw = synth.config2(batch=1000, seed=1234)
fs = kw.pop("fs")
xp = sg.pinned_empty(x.shape, np.float32); xp[...] = x
out = sg.pinned_empty((1000, 309, 257), np.float32)
plan = sp.triage(x.shape[-1], fs, kw["window"], kw["nperseg"], kw["noverlap"], None, "constant", True, "density", "psd")
eng = sg.engine()
dev = torch.device("cuda", 0)
B, n = x.shape
F, K = plan.nframes, plan.nbins
h_in, h_out = torch.from_numpy(xp), torch.from_numpy(out)
x_d = torch.empty((B, n), device=dev)
S_d = torch.empty((B, F, K), device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
cur = torch.cuda.current_stream()
row_bytes = F * K * 4
step = max(1, sp._PIPE_CHUNK_BYTES // row_bytes)
items, b, c = [], 0, max(1, step // sp._PIPE_RAMP_DIV)
while b < B:
    items.append((b, min(B, b + c))); b += c; c = min(step, c * 2)
parts = torch.empty((len(items), F, K), device=dev)


def E():
    return torch.cuda.Event(enable_timing=True)


def run(record):
    t_host0 = time.perf_counter()
    ev0 = E(); ev0.record(cur)
    s_in.wait_stream(cur); s_out.wait_stream(cur)
    rows = []
    for i, (b0, b1) in enumerate(items):
        with torch.cuda.stream(s_in):
            a = E(); a.record(s_in)
            x_d[b0:b1].copy_(h_in[b0:b1], non_blocking=True)
            ev_in = E(); ev_in.record(s_in)
        cur.wait_event(ev_in)
        k0 = E(); k0.record(cur)
        eng.stft_psd_sum(x_d[b0:b1], plan, out=S_d[b0:b1], sum_out=parts[i])
        ev_k = E(); ev_k.record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_k)
            d0 = E(); d0.record(s_out)
            h_out[b0:b1].copy_(S_d[b0:b1], non_blocking=True)
            d1 = E(); d1.record(s_out)
        rows.append((b1 - b0, a, ev_in, k0, ev_k, d0, d1, time.perf_counter() - t_host0))
    total = eng.batch_sum(parts, 1.0 / B)
    t_enq = time.perf_counter() - t_host0
    s_out.synchronize(); cur.synchronize()
    t_all = time.perf_counter() - t_host0
    if record:
        print(f"host: enqueue done {t_enq * 1e3:.3f} ms, all done {t_all * 1e3:.3f} ms")
        print("chunk sweeps | H2D start..end | kernel start..end | D2H start..end | host enqueue time (ms from call start)")
        last_d1 = 0.0
        for (nb, a, e_in, k0, e_k, d0, d1, th) in rows:
            f = lambda e: ev0.elapsed_time(e)
            gap = f(d0) - last_d1
            last_d1 = f(d1)
            print(f"{nb:5d} | {f(a):6.3f} {f(e_in):6.3f} | {f(k0):6.3f} {f(e_k):6.3f} | {f(d0):6.3f} {f(d1):6.3f} (D2H idle before: {gap:6.3f}) | {th * 1e3:6.3f}")


for _ in range(3):
    run(False)
run(True)
t0 = time.perf_counter()
for _ in range(8):
    sg.mean_spectrogram(xp, fs=fs, return_per_sweep=True, out=out, **kw)
torch.cuda.synchronize()
print("api ms per call", (time.perf_counter() - t0) / 8 * 1e3)

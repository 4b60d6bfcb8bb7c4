"""Seeded synthetic inputs for the BASELINE configs.

Everything is generated as float32 so that the CPU reference (which computes in
float64) can be fed the *same* fp32-representable samples.
"""
from __future__ import annotations

import numpy as np


def _chirp_phase(t, f0, f1, t1):
    # linear chirp phase: 2 pi (f0 t + (f1-f0) t^2 / (2 t1))
    return 2.0 * np.pi * (f0 * t + 0.5 * (f1 - f0) / t1 * t * t)


def config1(fs=44100.0, seconds=10.0):
    """C1: single mono linear chirp 20 Hz -> 20 kHz over 10 s, amplitude 1."""
    n = int(round(fs * seconds))
    t = np.arange(n, dtype=np.float64) / fs
    x = np.cos(_chirp_phase(t, 20.0, 20000.0, seconds))
    return x.astype(np.float32), dict(fs=fs, nperseg=1024, noverlap=768, window="hann")


def config2(batch=1000, fs=20000.0, seconds=2.0, seed=1234):
    """C2: `batch` sweeps, linear chirp 100 -> 5000 Hz with random phase + N(0, 0.1^2) noise."""
    rng = np.random.default_rng(seed)
    n = int(round(fs * seconds))
    t = np.arange(n, dtype=np.float64) / fs
    base = _chirp_phase(t, 100.0, 5000.0, seconds).astype(np.float32)
    phase = rng.uniform(0.0, 2.0 * np.pi, size=(batch, 1)).astype(np.float32)
    x = np.cos(base[None, :] + phase)
    x += 0.1 * rng.standard_normal((batch, n), dtype=np.float32)
    return x.astype(np.float32), dict(fs=fs, nperseg=512, noverlap=384, window="hann")


def config3(n=172_800_000, fs=48000.0, seed=2025):
    """C3: white noise N(0, 0.1^2) + tones at 1 kHz (1.0), 7 kHz (0.5), 15 kHz (0.25)."""
    rng = np.random.default_rng(seed)
    x = 0.1 * rng.standard_normal(n, dtype=np.float32)
    step = 1 << 22
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        t = np.arange(lo, hi, dtype=np.float64) / fs
        x[lo:hi] += (np.sin(2 * np.pi * 1000.0 * t) + 0.5 * np.sin(2 * np.pi * 7000.0 * t)
                     + 0.25 * np.sin(2 * np.pi * 15000.0 * t)).astype(np.float32)
    return x, dict(fs=fs, nperseg=2048, noverlap=1536, window="hann")


def config4(channels=16, fs=96000.0, seconds=60.0, seed=77):
    """C4: `channels` x noise + a channel-specific tone at 1 kHz * (c + 1)."""
    rng = np.random.default_rng(seed)
    n = int(round(fs * seconds))
    x = 0.1 * rng.standard_normal((channels, n), dtype=np.float32)
    t = np.arange(n, dtype=np.float64) / fs
    for c in range(channels):
        x[c] += np.sin(2 * np.pi * 1000.0 * (c + 1) * t).astype(np.float32)
    return x, dict(fs=fs, nperseg=4096, noverlap=3072, window="hann", scaling="density")


def config5(nperseg, overlap, batch=1, n=100_000, fs=10000.0, seed=5):
    """C5: noise + chirp, `n` samples, nperseg in {256..16384}, overlap in {.5,.75,.875}."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / fs
    chirp = np.cos(_chirp_phase(t, 50.0, 4000.0, n / fs)).astype(np.float32)
    x = chirp[None, :] + 0.2 * rng.standard_normal((batch, n), dtype=np.float32)
    hop = int(round(nperseg * (1.0 - overlap)))
    return x.astype(np.float32), dict(fs=fs, nperseg=nperseg, noverlap=nperseg - hop, window="hann")

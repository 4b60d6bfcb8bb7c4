"""Scratch diagnostic: staged-sample kernel (pairq) vs the round-1 kernel on the same input, repeated."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrogram_generator_b200 as sg
from spectrogram_generator_b200 import _lib

def run(nperseg, hop, B, nfr, detrend, reps=10):
    rng = np.random.default_rng(nperseg + hop)
    n = (nperseg + hop * (nfr - 1) + 8) // 4 * 4
    x = torch.from_numpy((rng.standard_normal((B, n)) + 1.0).astype(np.float32)).cuda()
    plan = sg.triage(n, 1000.0, "hann", nperseg, nperseg - hop, None, detrend, True, "density", "psd")
    eng = sg.engine()
    _lib.set_option("no_pairq", 1)
    ref = eng.stft_psd(x, plan)
    _lib.set_option("no_pairq", 0)          # opt in
    bad_total = 0
    for r in range(reps):
        got = eng.stft_psd(x, plan)
        err = (got - ref).abs() / ref.abs().max()
        bad = err > 1e-5
        nb = int(bad.sum())
        bad_total += nb
        if nb:
            idx = bad.nonzero()
            fr = sorted(set((int(a), int(b)) for a, b, _ in idx.tolist()))
            bins = idx[:, 2]
            print(f"  rep {r}: {nb} bad values in {len(fr)} frames {fr[:6]} bins {int(bins.min())}..{int(bins.max())} max err {float(err.max()):.2e}")
    _lib.set_option("no_pairq", 1)
    print(f"{nperseg}/{hop} B={B} nfr={plan.nframes} detrend={detrend}: {_lib.last_kernel().split(' ')[0]} bad values total {bad_total}")

for args in [(8192, 2048, 1, 9, False), (8192, 2048, 1, 9, "constant"), (8192, 2048, 64, 40, False), (16384, 4096, 1, 9, False),
             (16384, 4096, 32, 30, "constant"), (4096, 1024, 1, 9, False), (4096, 1024, 200, 60, False), (2048, 512, 1, 9, False), (2048, 512, 500, 80, "constant")]:
    run(*args)

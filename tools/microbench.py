#!/usr/bin/env python
"""Device-resident timing of the STFT kernel over a grid of shapes (CUDA events,
L2 defeated by working sets > 126 MB or an explicit flush).  Prints one JSON line
per shape: samples/s, algorithmic GB/s, fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrogram_generator_b200 as sg  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_shape(batch, n, nperseg, hop, iters=10, window="hann", detrend="constant", dtype=torch.float32, flush=None):
    plan = sg.triage(n, 1.0, window, nperseg, nperseg - hop, None, detrend, True, "density", "psd")
    eng = sg.engine()
    x = torch.randn((batch, n), device="cuda", dtype=dtype)
    out = torch.empty((batch, plan.nframes, plan.nbins), device="cuda", dtype=torch.float32)
    for _ in range(3):
        eng.stft_psd(x, plan, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.stft_psd(x, plan, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    bytes_alg = x.element_size() * batch * n + 4 * batch * plan.nframes * plan.nbins
    return dict(batch=batch, n=n, nperseg=nperseg, hop=hop, frames=plan.nframes, ms=round(ms, 4),
                ms_min=round(min(ts), 4), gsamples_s=round(batch * n / ms / 1e6, 2),
                gbs=round(bytes_alg / ms / 1e6, 1), frac=round(bytes_alg / ms / 1e6 / peak(), 4),
                mb=round(bytes_alg / 1e6, 1))


def time_mean(batch, n, nperseg, hop, iters=10, flush=None, fused=True):
    """Engine.stft_psd_sum (rows + cross-sweep sum in one call); fused=False: b2s_set_option("no_fused_sum")."""
    from spectrogram_generator_b200 import _lib
    plan = sg.triage(n, 1.0, "hann", nperseg, nperseg - hop, None, "constant", True, "density", "psd")
    eng = sg.engine()
    x = torch.randn((batch, n), device="cuda", dtype=torch.float32)
    out = torch.empty((batch, plan.nframes, plan.nbins), device="cuda", dtype=torch.float32)
    tot = torch.empty((plan.nframes, plan.nbins), device="cuda", dtype=torch.float32)
    _lib.set_option("no_fused_sum", 0 if fused else 1)
    try:
        for _ in range(3):
            eng.stft_psd_sum(x, plan, out=out, sum_out=tot)
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            if flush is not None:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.stft_psd_sum(x, plan, out=out, sum_out=tot)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        kernel = _lib.last_kernel()
    finally:
        _lib.set_option("no_fused_sum", 0)
    ms = float(np.median(ts))
    bytes_alg = 4 * batch * n + 4 * batch * plan.nframes * plan.nbins
    return dict(mean=True, fused=fused, batch=batch, n=n, nperseg=nperseg, hop=hop, frames=plan.nframes, ms=round(ms, 4),
                ms_min=round(min(ts), 4), gsamples_s=round(batch * n / ms / 1e6, 2),
                frac=round(bytes_alg / ms / 1e6 / peak(), 4), kernel=kernel.split(" (")[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="core")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"], help="sample type of the device-resident input")
    args = ap.parse_args()
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    flush = None if os.environ.get("B2S_MB_NOFLUSH") == "1" else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shapes = []
    if args.set in ("core", "all"):
        shapes += [(1000, 40000, 512, 128), (1000, 40000, 1024, 256), (1000, 40000, 1024, 896),
                   (1000, 40000, 1024, 512), (1000, 40000, 256, 64), (16, 5_760_000, 4096, 1024),
                   (1, 172_800_000, 2048, 512), (1, 441_000, 1024, 256)]
    if args.set == "c2":
        shapes += [(1000, 40000, 512, 128)]
    if args.set == "c4":
        shapes += [(16, 5_760_000, 4096, 1024)]
    if args.set == "n1024":
        shapes += [(1000, 40000, 1024, 256)]
    if args.set == "refcall":       # the reference's own call form: Tukey(0.25), noverlap = nperseg // 8, and no overlap
        for nperseg in (256, 512, 1024, 2048, 4096, 8192):
            shapes.append((1000, 200_000, nperseg, nperseg - nperseg // 8, ("tukey", .25)))
        for nperseg in (256, 512, 1024):
            shapes.append((1000, 200_000 // nperseg * nperseg, nperseg, nperseg, "hann"))
    if args.set == "small":         # nperseg <= 128 (warp kernel)
        shapes += [(1000, 200_000, 128, 112, ("tukey", .25)), (1000, 200_000, 128, 32), (1000, 200_000, 64, 56, ("tukey", .25)),
                   (1000, 200_000, 32, 28, ("tukey", .25))]
    if args.set == "anyhop":        # hops outside the sliding-window sets (B2S_NO_DUO=1 gives the warp kernel for comparison)
        shapes += [(1000, 200_000, 512, 300), (1000, 200_000, 512, 384), (1000, 200_000, 512, 100),
                   (1000, 200_000, 256, 192), (1000, 200_000, 256, 100)]
    if args.set == "n1024ref":      # nperseg 1024 at the reference's default overlap and without overlap
        shapes += [(1000, 200_000, 1024, 896), (1000, 200_704, 1024, 1024), (1000, 40000, 1024, 896),
                   (1000, 40960, 1024, 1024)]
    if args.set == "nonpow2":       # GUI-typed lengths (reference call form); B2S_NO_MIXED=1: the O(N^2) direct-DFT kernel
        for nperseg in (1000, 2000, 4800, 8000, 96, 352):
            shapes.append((1000, 40_000 if nperseg <= 2000 else 100_000, nperseg, nperseg - nperseg // 8, ("tukey", .25)))
        shapes += [(1000, 40_000, 1000, 250), (1000, 40_000, 2000, 500)]
    if args.set == "n2048":
        shapes += [(1024, 100_000, 2048, 512)]
    if args.set == "n8192":
        shapes += [(1024, 100_000, 8192, 2048)]
    if args.set == "n1024x":        # nperseg 1024 over hops and batch shapes (B2S_NO_PAIR=1: the four-step duo / CTA kernels)
        for hop in (128, 256, 512, 896, 1024):
            shapes += [(1000, 40000, 1024, hop), (1024, 100_000, 1024, hop)]
        shapes += [(1000, 200_000, 1024, 896, ("tukey", .25)), (1000, 200_704, 1024, 1024), (1, 441_000, 1024, 256),
                   (16, 5_760_000, 1024, 256)]
    if args.set in ("c5", "all"):
        for nperseg in (256, 512, 1024, 2048, 4096, 8192, 16384):
            for ov in (0.5, 0.75, 0.875):
                shapes.append((1024, 100_000, nperseg, int(nperseg * (1 - ov))))
    if args.set == "mean":          # rows + cross-sweep sum in one call: fused kernels against per-sweep kernel + two-pass sum
        for sh in [(1000, 40000, 512, 128), (1000, 40000, 1024, 256), (1000, 40000, 1024, 128), (1000, 40000, 1024, 512),
                   (1000, 200_000, 1024, 896), (1000, 200_704, 1024, 1024), (1000, 40000, 256, 64),
                   (1000, 40000, 2048, 512), (1000, 100_000, 2048, 512), (1000, 100_000, 4096, 1024), (1000, 100_000, 4096, 2048)]:
            for fused in (True, False):
                print(json.dumps(time_mean(*sh, flush=flush, fused=fused)), flush=True)
    if args.set == "meandyn":       # the sum-fused kernels: sweep blocks and static / dynamic unit schedule
        from spectrogram_generator_b200 import _lib
        for sh in [(1000, 40000, 512, 128), (1000, 40000, 1024, 256), (1000, 40000, 256, 64)]:
            for blocks, dyn in [(0, 0), (0, 1), (33, 1), (44, 1), (64, 1), (44, 0), (64, 0)]:
                _lib.set_option("sum_blocks", blocks)
                _lib.set_option("sum_dynamic", dyn)
                r = time_mean(*sh, flush=flush)
                r.update(sum_blocks=blocks, sum_dynamic=dyn)
                print(json.dumps(r), flush=True)
        _lib.set_option("sum_blocks", 0)
        _lib.set_option("sum_dynamic", 0)
    if args.set == "mean512":       # one shape, for ncu
        print(json.dumps(time_mean(1000, 40000, 512, 128, iters=3, flush=flush)), flush=True)
    if args.set == "mean1024":      # one shape, for ncu
        print(json.dumps(time_mean(1000, 40000, 1024, 256, iters=3, flush=flush)), flush=True)
    for s in shapes:
        kw = {"window": s[4]} if len(s) > 4 else {}
        r = time_shape(*s[:4], flush=flush, dtype=dtype, **kw)
        from spectrogram_generator_b200 import _lib
        r["kernel"] = _lib.last_kernel().split(" (")[0]
        if args.dtype == "f64":
            r["dtype"] = "f64"
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()

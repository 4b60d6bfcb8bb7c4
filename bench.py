#!/usr/bin/env python
"""Benchmark of the spectrogram hot path (BASELINE.json metric: STFT input
samples/sec and HBM GB/s fraction at 1/2/4/8 B200 vs CPU ref).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-shapes]

Headline workload (config.workload, `value`): BASELINE.json configs[1] -- a batch of 1,000
synthetic 2 s sweeps at 20 kHz, nperseg=512, hop=128 (Hann), per-sweep spectrograms + the
cross-sweep mean spectrogram.  One step = one pass of the path over that batch.  With N > 1
(launched by torchrun, one rank per GPU) every rank owns 1,000 sweeps (weak scaling); the only
collective is the all-reduce of the [309 x 257] partial sum for the mean.

Every other named shape of BASELINE.json rides along in `config.shapes[]` (outside the timed
region of the headline, each timed on its own with CUDA events, max over ranks): C1 as a single
call and as a batch, the batched 1024-point STFT the north star names (75 % overlap and the
reference's own call form), C3 (one hour @ 48 kHz; frame ranges sharded over the ranks, the
gather to rank 0 timed separately), C4 (16 channels x 60 s; channels sharded, gather timed
separately) and the C5 nperseg x overlap grid.  With N > 1 the line also records, outside
any timed region, that the sharded results equal the unsharded ones bit for bit and that the
peer-memory all-reduce equals a plain recomputation.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "stft_input_samples_per_sec"
UNIT = "samples/s"
WORKLOAD = "configs[1]: 1000 sweeps x 40000 samples @ 20 kHz, nperseg=512, hop=128, hann, detrend=constant, " \
           "per-sweep PSD + cross-sweep mean"
B, NS, FS, NPERSEG, HOP = 1000, 40000, 20000.0, 512, 128


def algorithmic_bytes(batch, n, nperseg, hop):
    F = (n - nperseg) // hop + 1
    K = nperseg // 2 + 1
    return 4 * batch * n + 4 * batch * F * K, F, K


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(fused=True):
    """Per-launch DRAM bytes of the STFT kernel from the committed ncu --set full capture."""
    try:
        key = "stft_psd_sum_c2_bytes_per_launch" if fused else "stft_psd_c2_bytes_per_launch"
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[key]
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max = index, False, [], set(), None
        self.h = None
        try:                                     # NVML set-up happens before the timed region
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:      # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except Exception:
                return index
        return index

    def sample(self):
        if self.h is None:
            return
        try:
            self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception as e:      # pragma: no cover
            self.reasons.add(f"nvml_error:{type(e).__name__}")

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.002)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def make_batch(seed_offset=0):
    from spectrogram_generator_b200 import synth
    x, kw = synth.config2(batch=B, seed=1234 + seed_offset)
    kw.pop("fs")
    return x, kw


def cpu_baseline_leg(x, kw, all_cores):
    """The reference's own path (scipy.signal.spectrogram, the call at PlotEngine.py:113
    with the config's window/overlap) on the host cores, float64, + the cross-sweep mean."""
    from oracle import reference_path
    cores = reference_path.host_cores() if all_cores else 1
    x64 = x.astype(np.float64)
    t0 = time.perf_counter()
    if cores == 1:
        f, t, S = reference_path.reference_call_kw(x64, FS, **kw)
        S.mean(axis=0)
    else:
        reference_path.run_sharded(x64, FS, kw, cores)
    dt = time.perf_counter() - t0
    return x.size / dt, cores, dt


def run_reference(args):
    """The reference arm: scipy.signal.spectrogram (the reference's call, float64) on all
    host cores, rows sharded over forked worker processes that inherit the input (set-up
    outside the timed steps, like the GPU arm's device-resident set-up)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import reference_path
    x, kw = make_batch()
    cores = reference_path.host_cores()
    x64 = x.astype(np.float64)
    steps = max(1, min(args.steps, 5))
    warm = max(0, min(args.warmup, 2))
    with reference_path.ShardedRunner(x64, FS, kw, cores) as runner:
        for _ in range(warm):
            runner.step()
        dts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            parts = runner.step()
            sum(p[1] for p in parts) / B            # the cross-sweep mean
            dts.append(time.perf_counter() - t0)
    dt = float(np.median(dts))
    v = x.size / dt
    sample = f"full workload (1000 sweeps) per step, rows sharded over {cores} forked processes (pool and input " \
             "set up outside the timed steps), each running scipy.signal.spectrogram (SciPy default: 1 FFT thread) " \
             "in float64 and returning its partial cross-sweep sum"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "note": f"CPU arm: a step takes ~0.25 s, so at most 5 timed steps and 2 warm-ups are run whatever "
                           f"--steps/--warmup ask for (asked: {args.steps}/{args.warmup}); the value is the median step"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------
# the other named shapes (config.shapes[])
# --------------------------------------------------------------------------------------------

class ShapeBench:
    """Device-resident timing of one launch of the path per shape: CUDA events on the launching
    stream around every iteration, median over the iterations, max over the ranks."""

    def __init__(self, torch, dist, sg, dev, world, rank):
        self.torch, self.dist, self.sg, self.dev, self.world, self.rank = torch, dist, sg, dev, world, rank
        self.eng = sg.engine()
        self.peak, _ = hbm_peak()
        self.flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)       # > the 126 MB L2
        self.launches = 0

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def time_ms(self, fn, iters=7, warm_ms=25.0, flush=True):
        """Median of `iters` launches, each bracketed by CUDA events, after at least 3 warm-up launches and
        `warm_ms` of back-to-back launches.  Before every timed launch the L2 is flushed (a 192 MB memset in stream
        order); besides defeating the cache this keeps the device busy while the host enqueues event - launch -
        event, so the bracket does not contain the ~15 us the Python / ctypes call needs to reach cudaLaunchKernel
        on an idle stream (measured: every shape reads a constant 15 us longer without it)."""
        torch = self.torch
        t0 = time.perf_counter()
        n_warm = 0
        while n_warm < 3 or (time.perf_counter() - t0) * 1e3 < warm_ms:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            n_warm += 3
            if n_warm >= 300:
                break
        if self.world > 1:
            self.dist.barrier()
        ts = []
        for _ in range(iters):
            if flush:
                self.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        self.launches += iters + n_warm
        return self.max_over_ranks(float(np.median(ts)))

    def plan(self, n, nperseg, hop, window="hann"):
        return self.sg.triage(n, 1.0, window, nperseg, nperseg - hop, None, "constant", True, "density", "psd")

    def stft(self, name, batch, n, nperseg, hop, window="hann", iters=5, flush=True, global_batch=None,
             scaling="weak", note=None):
        """`batch` signals of `n` samples on THIS rank; global_batch: signals over all ranks."""
        torch = self.torch
        from spectrogram_generator_b200 import _lib
        plan = self.plan(n, nperseg, hop, window)
        x = torch.randn((batch, n), device=self.dev, dtype=torch.float32)
        out = torch.empty((batch, plan.nframes, plan.nbins), device=self.dev, dtype=torch.float32)
        ms = self.time_ms(lambda: self.eng.stft_psd(x, plan, out=out), iters=iters, flush=flush)
        gb = batch * self.world if global_batch is None else global_batch
        bytes_alg = 4 * gb * n + 4 * gb * plan.nframes * plan.nbins
        d = {"name": name, "signals": gb, "samples_per_signal": n, "nperseg": nperseg, "hop": hop,
             "frames": plan.nframes, "ms": round(ms, 5), "gsamples_s": round(gb * n / ms / 1e6, 2),
             "gbs_per_gpu": round(bytes_alg / self.world / ms / 1e6, 1),
             "frac": round(bytes_alg / self.world / ms / 1e6 / self.peak, 4), "scaling": scaling,
             "kernel": _lib.last_kernel()}
        if note:
            d["note"] = note
        del x, out
        return d

    def stft_mean(self, name, batch, n, nperseg, hop, window="hann", iters=5):
        """Per-sweep spectrograms AND their cross-sweep sum in one call (Engine.stft_psd_sum), timed in the fused
        form and -- same call, b2s_set_option("no_fused_sum") -- as per-sweep kernel + two-pass sum."""
        torch = self.torch
        from spectrogram_generator_b200 import _lib
        plan = self.plan(n, nperseg, hop, window)
        x = torch.randn((batch, n), device=self.dev, dtype=torch.float32)
        out = torch.empty((batch, plan.nframes, plan.nbins), device=self.dev, dtype=torch.float32)
        tot = torch.empty((plan.nframes, plan.nbins), device=self.dev, dtype=torch.float32)
        call = lambda: self.eng.stft_psd_sum(x, plan, post_scale=1.0 / batch, out=out, sum_out=tot)
        ms = self.time_ms(call, iters=iters)
        kernel = _lib.last_kernel()
        _lib.set_option("no_fused_sum", 1)
        try:
            ms2 = self.time_ms(call, iters=iters)
        finally:
            _lib.set_option("no_fused_sum", 0)
        gb = batch * self.world
        bytes_alg = 4 * gb * n + 4 * gb * plan.nframes * plan.nbins
        d = {"name": name, "signals": gb, "samples_per_signal": n, "nperseg": nperseg, "hop": hop,
             "frames": plan.nframes, "ms": round(ms, 5), "gsamples_s": round(gb * n / ms / 1e6, 2),
             "gbs_per_gpu": round(bytes_alg / self.world / ms / 1e6, 1),
             "frac": round(bytes_alg / self.world / ms / 1e6 / self.peak, 4), "scaling": "weak", "kernel": kernel,
             "ms_two_pass": round(ms2, 5),
             "note": "rows + cross-sweep sum in one call; ms_two_pass: the same call as per-sweep kernel + two-pass sum "
                     "(the rows read back once)"}
        del x, out, tot
        return d

    # C3: one long recording, contiguous frame ranges per rank (distributed.shard_frames / sample_span)
    def c3(self, n_total=172_800_000, nperseg=2048, hop=512):
        torch, sg = self.torch, self.sg
        from spectrogram_generator_b200 import _lib, distributed as D
        plan_g = self.plan(n_total, nperseg, hop)
        f0, cnt = D.shard_frames(plan_g.nframes, self.world, self.rank)
        lo, hi = D.sample_span(f0, cnt, hop, nperseg)
        sub = sg.Plan(**{**plan_g.__dict__, "n": hi - lo, "nframes": cnt})
        x = torch.randn((1, hi - lo), device=self.dev, dtype=torch.float32)
        out = torch.empty((1, cnt, plan_g.nbins), device=self.dev, dtype=torch.float32)
        ms = self.time_ms(lambda: self.eng.stft_psd(x, sub, out=out))
        bytes_alg = 4 * n_total + 4 * plan_g.nframes * plan_g.nbins
        d = {"name": "C3 configs[2]: 1 h @ 48 kHz, frame ranges sharded over the ranks (halo nperseg - hop read-only)",
             "signals": 1, "samples_per_signal": n_total, "nperseg": nperseg, "hop": hop, "frames": plan_g.nframes,
             "frames_per_rank": cnt, "ms": round(ms, 5), "gsamples_s": round(n_total / ms / 1e6, 2),
             "gbs_per_gpu": round(bytes_alg / self.world / ms / 1e6, 1),
             "frac": round(bytes_alg / self.world / ms / 1e6 / self.peak, 4), "scaling": "strong",
             "kernel": _lib.last_kernel()}
        if self.world > 1:
            counts = [c for _, c in sg.split_frames(plan_g.nframes, self.world)]
            gms = self.time_gather(out[0], counts)
            d["gather_to_rank0_ms"] = round(gms, 4)
            d["gather_bytes"] = int(4 * (plan_g.nframes - counts[0]) * plan_g.nbins)
            d["gsamples_s_gather_inclusive"] = round(n_total / (ms + gms) / 1e6, 2)
        del x, out
        return d

    # C4: channels sharded over the ranks (distributed.shard_rows)
    def c4(self, channels=16, n=5_760_000, nperseg=4096, hop=1024):
        torch = self.torch
        from spectrogram_generator_b200 import _lib, distributed as D
        plan = self.plan(n, nperseg, hop)
        lo, hi = D.shard_rows(channels, self.world, self.rank)
        rows = max(hi - lo, 0)
        x = torch.randn((max(rows, 1), n), device=self.dev, dtype=torch.float32)[:rows]
        out = torch.empty((rows, plan.nframes, plan.nbins), device=self.dev, dtype=torch.float32)
        fn = (lambda: self.eng.stft_psd(x, plan, out=out)) if rows else (lambda: None)
        ms = self.time_ms(fn)
        bytes_alg = 4 * channels * n + 4 * channels * plan.nframes * plan.nbins
        d = {"name": "C4 configs[3]: 16 channels x 60 s @ 96 kHz, PSD scaling, channels sharded over the ranks",
             "signals": channels, "samples_per_signal": n, "nperseg": nperseg, "hop": hop, "frames": plan.nframes,
             "channels_per_rank": rows, "ms": round(ms, 5), "gsamples_s": round(channels * n / ms / 1e6, 2),
             "gbs_per_gpu": round(bytes_alg / self.world / ms / 1e6, 1),
             "frac": round(bytes_alg / self.world / ms / 1e6 / self.peak, 4), "scaling": "strong",
             "kernel": _lib.last_kernel()}
        if self.world > 1:
            counts = [D.shard_rows(channels, self.world, r)[1] - D.shard_rows(channels, self.world, r)[0]
                      for r in range(self.world)]
            gms = self.time_gather(out, counts)
            d["gather_to_rank0_ms"] = round(gms, 4)
            d["gather_bytes"] = int(4 * (channels - counts[0]) * plan.nframes * plan.nbins)
            d["gsamples_s_gather_inclusive"] = round(channels * n / (ms + gms) / 1e6, 2)
        del x, out
        return d

    def time_gather(self, local, counts, iters=3):
        """The 'final gather to the exporting rank' (distributed.gather_slabs): NCCL send / recv of every
        rank's slab to rank 0, timed on the device from a barrier to rank 0 holding everything."""
        torch, dist = self.torch, self.dist
        from spectrogram_generator_b200 import distributed as D
        ts = []
        for i in range(iters + 1):
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g = D.gather_slabs(local, counts, dst=0)
            b.record()
            torch.cuda.synchronize()
            if i:
                ts.append(a.elapsed_time(b))
            del g
        return self.max_over_ranks(float(np.median(ts)))

    def c5(self):
        """configs[4]: nperseg x overlap grid on 100 000-sample signals, 1024 signals per GPU so that the
        device is saturated (a single 100 k signal is launch-bound)."""
        from spectrogram_generator_b200 import _lib
        rows = []
        for nperseg in (256, 512, 1024, 2048, 4096, 8192, 16384):
            for ov in (0.5, 0.75, 0.875):
                hop = int(nperseg * (1 - ov))
                d = self.stft("c5", 1024, 100_000, nperseg, hop, iters=3)
                rows.append([nperseg, hop, d["ms"], d["gsamples_s"], d["frac"], d["kernel"].split(" ")[0]])
        return {"name": "C5 configs[4]: nperseg x overlap grid, 1024 signals x 100 000 samples per GPU",
                "columns": ["nperseg", "hop", "ms", "gsamples_s", "frac", "kernel"], "rows": rows,
                "scaling": "weak"}

    def run(self):
        shapes = []
        shapes.append(self.stft("C1 configs[0]: one 10 s chirp-length signal @ 44.1 kHz, single call (launch-bound), "
                                "replicated per rank", 1, 441_000, 1024, 256, iters=10, flush=True))
        shapes.append(self.stft("C1 batched: 256 such signals per GPU in one launch", 256, 441_000, 1024, 256, iters=5))
        shapes.append(self.stft("north-star target: batched 1024-point STFT, 1000 x 40 000 per GPU, 75 % overlap",
                                1000, 40_000, 1024, 256, iters=7))
        shapes.append(self.stft_mean("north-star target shape with its mean: 1000 x 40 000 per GPU @ 1024/256, per-sweep "
                                     "spectrograms + cross-sweep sum", 1000, 40_000, 1024, 256))
        shapes.append(self.stft_mean("the same batch @ 256/64, per-sweep spectrograms + cross-sweep sum", 1000, 40_000, 256, 64))
        shapes.append(self.stft_mean("the same batch @ 2048/512, per-sweep spectrograms + cross-sweep sum", 1000, 40_000, 2048, 512))
        shapes.append(self.stft("north-star target, the reference's own call form (PlotEngine.py:113: Tukey(0.25), "
                                "noverlap = nperseg//8, GUI default nperseg 1024), 1000 x 200 000 per GPU",
                                1000, 200_000, 1024, 896, window=("tukey", .25), iters=5))
        shapes.append(self.stft("batched 1024-point STFT without overlap, 1000 x 200 704 per GPU",
                                1000, 200_704, 1024, 1024, iters=5))
        shapes.append(self.c3())
        shapes.append(self.c4())
        shapes.append(self.c5())
        return shapes


def multi_gpu_checks(torch, dist, sg, dev, world, rank):
    """Outside every timed region: frame-range sharded == unsharded and channel-sharded == unsharded,
    bit for bit, through the host-facing sharding API (distributed.spectrogram_time_sharded, shard_rows,
    gather_slabs).  Every rank builds the same seeded recording; rank 0 holds the verdict."""
    from spectrogram_generator_b200 import distributed as D
    eng = sg.engine()
    out = {}
    rng = np.random.default_rng(2025)
    n = 6_000_000
    x = (0.1 * rng.standard_normal(n, dtype=np.float32)
         + np.sin(2 * np.pi * 1000.0 * np.arange(n, dtype=np.float64) / 48000.0).astype(np.float32))
    kw = dict(window="hann", nperseg=2048, noverlap=1536)
    plan = sg.triage(n, 48000.0, "hann", 2048, 1536, None, "constant", True, "density", "psd")
    f0, cnt = D.shard_frames(plan.nframes, world, rank)
    lo, hi = D.sample_span(f0, cnt, plan.hop, plan.nperseg)
    f, t, S_loc, _ = D.spectrogram_time_sharded(x[lo:hi], n, 48000.0, **kw)
    counts = [c for _, c in sg.split_frames(plan.nframes, world)]
    full = D.gather_slabs(S_loc, counts, dst=0)
    t_ok = np.array_equal(t, sg.windows.time_axis(n, 2048, 1536, 48000.0)[f0:f0 + cnt])
    ok = torch.tensor([1 if t_ok else 0], device=dev)
    if rank == 0:
        ref = eng.stft_psd(torch.from_numpy(x).to(dev).view(1, -1), plan)[0]
        ok[0] = int(bool(torch.equal(full, ref)) and t_ok)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out["time_sharded"] = (f"C3 form, {n} samples @ 2048/512 over {world} ranks through spectrogram_time_sharded + "
                           f"gather_slabs: {'bit-identical to the unsharded launch, time axis identical' if int(ok[0]) else 'MISMATCH'}")
    assert int(ok[0]) == 1, out["time_sharded"]
    # channels
    ch, n4 = 16, 400_000
    x4 = (0.1 * rng.standard_normal((ch, n4), dtype=np.float32))
    plan4 = sg.triage(n4, 96000.0, "hann", 4096, 3072, None, "constant", True, "density", "psd")
    lo, hi = D.shard_rows(ch, world, rank)
    S4 = eng.stft_psd(torch.from_numpy(x4[lo:hi]).to(dev), plan4) if hi > lo else \
        torch.empty((0, plan4.nframes, plan4.nbins), device=dev)
    counts4 = [D.shard_rows(ch, world, r)[1] - D.shard_rows(ch, world, r)[0] for r in range(world)]
    full4 = D.gather_slabs(S4, counts4, dst=0)
    ok4 = torch.tensor([1], device=dev)
    if rank == 0:
        ok4[0] = int(bool(torch.equal(full4, eng.stft_psd(torch.from_numpy(x4).to(dev), plan4))))
    dist.all_reduce(ok4, op=dist.ReduceOp.MIN)
    out["channel_sharded"] = (f"C4 form, {ch} channels x {n4} samples @ 4096/1024 over {world} ranks through shard_rows + "
                              f"gather_slabs: {'bit-identical to the unsharded launch' if int(ok4[0]) else 'MISMATCH'}")
    assert int(ok4[0]) == 1, out["channel_sharded"]
    return out


def pcie_rates(torch, dev, h2d_bytes, d2h_bytes):
    """Plain pinned copies of the e2e leg's sizes: each direction alone and both at once (GB/s)."""
    hin = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
    din = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    hout = torch.empty(d2h_bytes, dtype=torch.uint8, pin_memory=True)
    dout = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(do_in, do_out):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if do_in:
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if do_out:
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    for _ in range(2):
        run(True, True)
    a = min(run(True, False) for _ in range(3))
    b = min(run(False, True) for _ in range(3))
    c = min(run(True, True) for _ in range(3))
    return {"h2d_gbs": round(h2d_bytes / a / 1e9, 1), "d2h_gbs": round(d2h_bytes / b / 1e9, 1),
            "duplex_s": c, "duplex_gbs": round((h2d_bytes + d2h_bytes) / c / 1e9, 1)}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import spectrogram_generator_b200 as sg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # The all-reduce of the partial mean is issued in stream order after the cross-sweep sum.
    # Diagnostics (measured on 2 GPUs, 200 steps: none 0.222 ms/step, sync 0.237, async 1.23 -- async
    # NCCL kernels spin beside the persistent STFT grids of their peers; reserving SMs for them with
    # b2s_set_reserved_sms(8) only brings it back to 0.63): B2S_BENCH_ALLREDUCE = overlap (default) | peer | sync | async | none.
    # "overlap": the peer kernel on a high-priority side stream, so the all-reduce of step i runs beside the
    # STFT of step i + 1, in the CTA slots of 5 SMs the STFT grid leaves free (b2s_set_reserved_sms).
    mode = os.environ.get("B2S_BENCH_ALLREDUCE", "overlap")
    overlap = (mode == "overlap")
    if overlap:
        mode = "peer"
    reserve = max(0, int(os.environ.get("B2S_BENCH_RESERVE_SMS", "5" if (overlap and world > 1) else "0")))
    from spectrogram_generator_b200 import _lib
    _lib.load().b2s_set_reserved_sms(reserve)

    x_host, kw = make_batch(seed_offset=rank)
    plan = sg.triage(NS, FS, kw["window"], kw["nperseg"], kw["noverlap"], None, "constant", True, "density", "psd")
    eng = sg.engine()
    x = torch.from_numpy(x_host).to(dev)
    S = torch.empty((B, plan.nframes, plan.nbins), dtype=torch.float32, device=dev)
    total_sweeps = B * world

    # the mean all-reduce: one kernel over NVLink peer memory (distributed.PeerMeanReducer); NCCL if
    # symmetric memory cannot be set up on this box (B2S_BENCH_ALLREDUCE=sync forces NCCL)
    peer = None
    if world > 1 and mode == "peer":
        try:
            from spectrogram_generator_b200.distributed import PeerMeanReducer
            peer = PeerMeanReducer(plan.nframes * plan.nbins, dev, overlap=overlap,
                                   coresident=os.environ.get("B2S_BENCH_PEER_CORESIDENT", "0") == "1",
                                   nbuf=int(os.environ.get("B2S_BENCH_PEER_NBUF", "3")))
        except Exception as e:          # pragma: no cover
            if rank == 0:
                print(f"peer all-reduce unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
            mode = "sync"
    collective = {"peer": "b2s_peer_allreduce_f32 (one kernel over NVLink peer memory" +
                          ("; on a side stream, overlapping the next step's STFT)" if overlap else ")"), "sync": "NCCL all_reduce",
                  "async": "NCCL all_reduce (async)", "none": "none"}[mode] if world > 1 else "none"

    # per-sweep spectrograms + cross-sweep sum: ONE kernel keeps the running sums on chip while it
    # writes the rows (b2s_stft_psd_sum_f32), a 22-row fold finishes the sum.  B2S_BENCH_FUSED_SUM=0 runs
    # the earlier two-kernel form (b2s_stft_psd_f32, then b2s_batch_sum_f32 reading all rows back).
    fused = os.environ.get("B2S_BENCH_FUSED_SUM", "1") != "0"
    mean_buf = torch.empty((plan.nframes, plan.nbins), dtype=torch.float32, device=dev)

    def spectrograms():
        """The STFT launch of the step; in the fused form it also leaves the (partial) sum."""
        if not fused:
            eng.stft_psd(x, plan, out=S)
        elif peer is not None:
            eng.stft_psd_sum(x, plan, out=S, sum_out=peer.partial())
        else:
            eng.stft_psd_sum(x, plan, post_scale=1.0 / total_sweeps, out=S, sum_out=mean_buf)

    def partial_mean():
        if fused:
            return mean_buf
        return eng.batch_sum(S, 1.0 / total_sweeps, out=mean_buf)

    peer_out = [torch.empty(plan.nframes * plan.nbins, dtype=torch.float32, device=dev) for _ in range(16)]

    def reduce_mean():
        if peer is not None:
            if not fused:
                eng.batch_sum(S, 1.0, out=peer.partial())
            return peer.reduce(1.0 / total_sweeps, out=peer_out[peer.epoch & 15])
        mean = partial_mean()                            # partial mean of this rank's sweeps
        if world > 1 and mode == "sync":
            dist.all_reduce(mean)                        # sum of partial means == global mean
        return mean

    def step():
        spectrograms()
        return reduce_mean()

    # untimed: bring the device to its steady state first (clocks, TLBs, the peers' rhythm) -- a 20-step timed
    # region is only 3.4 ms long, and measured right after 5 warm-up steps it reads 3-4 % slower than the same
    # 20 steps taken after 100 (1 GPU 0.1724 vs 0.1696 ms/step; 2 GPUs 0.1818 vs 0.1696) -- then the W warm-up steps
    pre_warm = max(0, int(os.environ.get("B2S_BENCH_PREWARM_STEPS", "150")))
    for _ in range(pre_warm):
        mean = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for _ in range(max(3, args.warmup)):
        mean = step()
    torch.cuda.synchronize()

    # --- timed region: K steps, CUDA events on the launching stream, max over ranks ---
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    start.record()
    pending = []
    t_host = time.perf_counter()
    for i in range(args.steps):
        k_ev[i][0].record()
        spectrograms()
        k_ev[i][1].record()
        if world > 1 and mode == "async":
            mean = partial_mean().clone()
            pending.append((dist.all_reduce(mean, async_op=True), mean))
        else:
            mean = reduce_mean()
    for w, _ in pending:
        w.wait()
    if peer is not None:
        peer.wait()                  # overlap mode: the side stream's reduces join the timed stream
    end.record()
    host_ms = 1e3 * (time.perf_counter() - t_host) / args.steps      # host time to enqueue one step
    sampler.sample()                 # all K steps are enqueued: this sample is taken under load
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    elapsed_ms = start.elapsed_time(end)
    # outside the timed region: the last step's mean against a plain recomputation (two-pass sum + NCCL)
    check = None
    if peer is not None:
        peer.check()                 # raises if any reduce gave up waiting for a late rank
    if world > 1 and mode in ("peer", "sync"):
        ref = eng.batch_sum(S, 1.0 / total_sweeps)
        dist.all_reduce(ref)
        err = float((mean.view(-1) - ref.view(-1)).abs().max() / ref.abs().max())
        same = [torch.empty_like(mean.view(-1)) for _ in range(world)]
        dist.all_gather(same, mean.view(-1).contiguous())
        identical = all(bool(torch.equal(same[0], s)) for s in same)
        check = (f"last step's mean vs batch_sum + NCCL all_reduce: max abs diff {err:.1e} of max; "
                 f"bit-identical on all {world} ranks: {identical}")
        assert err < 1e-5 and identical, check
    kern_all = [float(a.elapsed_time(b)) for a, b in k_ev]
    kern_ms_local = float(np.mean(kern_all))
    kern_ms, kern_per_rank = kern_ms_local, [round(kern_ms_local, 5)]
    if world > 1:
        tt = torch.tensor([elapsed_ms, kern_ms_local], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms, kern_ms = float(tt[0]), float(tt[1])
        g = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(g, torch.tensor([kern_ms_local], device=dev, dtype=torch.float64))
        kern_per_rank = [round(float(v[0]), 5) for v in g]
    value = world * B * NS * args.steps / (elapsed_ms * 1e-3)

    # --- end to end through the public API: pinned host buffers in, NumPy arrays out; with N > 1 the
    # partial sums are all-reduced inside the call (mean_spectrogram(..., total_sweeps=, group=)) ---
    xp = sg.pinned_empty(x_host.shape, np.float32)
    xp[...] = x_host
    out_buf = sg.pinned_empty((B, plan.nframes, plan.nbins), np.float32)       # caller-owned result buffer, reused
    api_kw = dict(fs=FS, window=kw["window"], nperseg=kw["nperseg"], noverlap=kw["noverlap"], out=out_buf)
    if world > 1:
        api_kw.update(total_sweeps=total_sweeps, group=dist.group.WORLD)
    res = None
    for _ in range(3):
        res = sg.mean_spectrogram(xp, return_per_sweep=True, **api_kw)
    e2e_steps = max(3, min(args.steps, 10))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = sg.mean_spectrogram(xp, return_per_sweep=True, **api_kw)
    torch.cuda.synchronize()
    f, t, m, Sx = res
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt[0])
    e2e_value = world * B * NS * e2e_steps / e2e_s
    d2h = int(Sx.nbytes + m.nbytes)
    pcie = pcie_rates(torch, dev, int(x_host.nbytes), d2h)
    e2e_step_s = e2e_s / e2e_steps
    del res, Sx, out_buf, xp

    # --- the other named shapes, each timed on its own (config.shapes[]) ---
    shapes, mg = None, None
    if not args.no_shapes:
        torch.cuda.empty_cache()
        _lib.load().b2s_set_reserved_sms(0)
        sb = ShapeBench(torch, dist, sg, dev, world, rank)
        shapes = sb.run()
        if world > 1:
            mg = multi_gpu_checks(torch, dist, sg, dev, world, rank)

    if rank == 0:
        bytes_alg, F, K = algorithmic_bytes(B, NS, NPERSEG, HOP)
        peak, peak_src = hbm_peak()
        achieved = bytes_alg / (kern_ms * 1e-3) / 1e9
        cpu_v, cpu_cores, cpu_dt = cpu_baseline_leg(x_host, kw, all_cores=False)
        if fused:
            kernel_name = ("stft_psd_duo_sum_kernel<float,S=4,ACC_TMEM=1> (nperseg 512, hop 128: two frames per lane "
                           "group, packed fp32x2, walks a block of sweeps and keeps their running sums in tensor "
                           "memory); kernel_ms brackets this launch plus the ~5 us fold of its 22 partial sums")
        else:
            kernel_name = ("stft_psd_duo_kernel<float,S=4,EPI_PLAIN> (nperseg 512, hop 128: two frames per lane group, "
                           "packed fp32x2)")
        config = {"workload": WORKLOAD, "sweeps_per_gpu": B, "global_sweeps": total_sweeps,
                  "frames_per_sweep": F, "bins": K,
                  "allreduce_check": check, "host_enqueue_ms_per_step": round(host_ms, 4),
                  "kernel_ms_per_rank": kern_per_rank, "pre_warm_steps": pre_warm,
                  "l2": "inputs+outputs per step (478 MB) exceed the 126 MB L2; no explicit flush",
                  "parallelism": f"sweeps sharded, {world} rank(s); all_reduce of the [F,K] partial sum only "
                                 f"({collective}, in stream order after the cross-sweep sum, inside the timed region)"}
        if shapes is not None:
            config["shapes"] = shapes
            config["shapes_note"] = ("each shape: one launch of the path on device-resident synthetic input, CUDA events, median "
                                     "of 3-10 iterations after >= 25 ms of warm-up launches, max over ranks; frac = algorithmic bytes (4 B/sample in + "
                                     "4 B/bin out) per GPU / ms / measured HBM peak; the L2 is flushed (192 MB memset in stream order) "
                                     "before every timed launch; 'weak' = the stated batch per GPU, 'strong' = one problem "
                                     "cut over the ranks; gather_to_rank0_ms = the final gather to the exporting rank "
                                     "(NCCL send/recv), reported beside, not inside, the kernel time")
        if mg is not None:
            config["multi_gpu_checks"] = mg
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(fused),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_alg,
                         "kernel_ms": kern_ms, "frac_of_nominal_8000": achieved / 8000.0},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x_host.nbytes),
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "spectrogram_generator_b200.mean_spectrogram(x_pinned, return_per_sweep=True, out=pinned_result"
                           + (", total_sweeps=, group=WORLD)  [all-reduce of the partial sums inside the call]" if world > 1 else ")"),
                    "pcie": {"h2d_gbs_alone": pcie["h2d_gbs"], "d2h_gbs_alone": pcie["d2h_gbs"],
                             "both_directions_gbs": pcie["duplex_gbs"],
                             "note": "plain pinned cudaMemcpyAsync of the same byte counts on this rank, outside the timed region"},
                    "roofline": {"bound": "pcie", "achieved": round((int(x_host.nbytes) + d2h) / e2e_step_s / 1e9, 1),
                                 "peak": pcie["duplex_gbs"], "unit": "GB/s",
                                 "frac": round(((int(x_host.nbytes) + d2h) / e2e_step_s) / ((int(x_host.nbytes) + d2h) / pcie["duplex_s"]), 3)}},
            "gpu_launches": args.steps * ((2 if fused else 3) + (1 if peer is not None else 0)),
            "clocks": sampler.summary(),
            "cpu_baseline": {
                "value": cpu_v, "unit": UNIT, "cores": cpu_cores, "kind": "reference",
                "sample": f"full workload once ({cpu_dt:.2f} s): scipy.signal.spectrogram on the [1000,40000] float64 "
                          "batch + mean over sweeps, 1 process, SciPy default 1 FFT thread (the reference's call as-is), "
                          "on rank 0's host cores"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-shapes", action="store_true", help="skip config.shapes[] (the headline only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

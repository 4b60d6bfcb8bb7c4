#!/usr/bin/env python
"""Benchmark of the spectrogram hot path (BASELINE.json metric: STFT input
samples/sec and HBM GB/s fraction at 1/2/4/8 B200 vs CPU ref).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- a batch of 1,000 synthetic
2 s sweeps at 20 kHz, nperseg=512, hop=128 (Hann), per-sweep spectrograms + the
cross-sweep mean spectrogram.  One step = one pass of the path over that batch.
With N > 1 (launched by torchrun, one rank per GPU) every rank owns 1,000 sweeps
(weak scaling); the only collective is the all-reduce of the [309 x 257] partial
sum for the mean.  One JSON line is printed by rank 0.
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
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


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
            peer = PeerMeanReducer(plan.nframes * plan.nbins, dev, overlap=overlap)
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

    peer_out = [torch.empty(plan.nframes * plan.nbins, dtype=torch.float32, device=dev) for _ in range(4)]

    def reduce_mean():
        if peer is not None:
            if not fused:
                eng.batch_sum(S, 1.0, out=peer.partial())
            return peer.reduce(1.0 / total_sweeps, out=peer_out[peer.epoch & 3])
        mean = partial_mean()                            # partial mean of this rank's sweeps
        if world > 1 and mode == "sync":
            dist.all_reduce(mean)                        # sum of partial means == global mean
        return mean

    def step():
        spectrograms()
        return reduce_mean()

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
    if world > 1 and mode in ("peer", "sync"):
        ref = eng.batch_sum(S, 1.0 / total_sweeps)
        dist.all_reduce(ref)
        err = float((mean.view(-1) - ref.view(-1)).abs().max() / ref.abs().max())
        check = f"last step's mean vs batch_sum + NCCL all_reduce: max abs diff {err:.1e} of max"
        assert err < 1e-5, check
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    if world > 1:
        tt = torch.tensor([elapsed_ms, kern_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms, kern_ms = float(tt[0]), float(tt[1])
    value = world * B * NS * args.steps / (elapsed_ms * 1e-3)

    # --- end to end through the public API: pinned host buffers in, NumPy arrays out ---
    xp = sg.pinned_empty(x_host.shape, np.float32)
    xp[...] = x_host
    api_kw = dict(fs=FS, window=kw["window"], nperseg=kw["nperseg"], noverlap=kw["noverlap"])
    res = None
    for _ in range(4):      # warm-up holding the previous result, as the timed loop does
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

    if rank == 0:
        bytes_alg, F, K = algorithmic_bytes(B, NS, NPERSEG, HOP)
        peak, peak_src = hbm_peak()
        achieved = bytes_alg / (kern_ms * 1e-3) / 1e9
        cpu_v, cpu_cores, cpu_dt = cpu_baseline_leg(x_host, kw, all_cores=False) if world == 1 else (None, None, None)
        if fused:
            kernel_name = ("stft_psd_duo_sum_kernel<float,S=4,ACC_TMEM=1> (nperseg 512, hop 128: two frames per lane "
                           "group, packed fp32x2, walks a block of sweeps and keeps their running sums in tensor "
                           "memory); kernel_ms brackets this launch plus the ~5 us fold of its 22 partial sums")
        else:
            kernel_name = ("stft_psd_duo_kernel<float,S=4,EPI_PLAIN> (nperseg 512, hop 128: two frames per lane group, "
                           "packed fp32x2)")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sweeps_per_gpu": B, "global_sweeps": total_sweeps,
                       "frames_per_sweep": F, "bins": K,
                       "allreduce_check": check, "host_enqueue_ms_per_step": round(host_ms, 4),
                       "l2": "inputs+outputs per step (478 MB) exceed the 126 MB L2; no explicit flush",
                       "parallelism": f"sweeps sharded, {world} rank(s); all_reduce of the [F,K] partial sum only "
                                      f"({collective}, in stream order after the cross-sweep sum, inside the timed region)"},
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(fused),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_alg,
                         "kernel_ms": kern_ms, "frac_of_nominal_8000": achieved / 8000.0},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x_host.nbytes),
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "spectrogram_generator_b200.mean_spectrogram(x_pinned, return_per_sweep=True)"},
            "gpu_launches": args.steps * ((2 if fused else 3) + (1 if peer is not None else 0)),
            "clocks": sampler.summary(),
        }
        if cpu_v is not None:
            line["cpu_baseline"] = {
                "value": cpu_v, "unit": UNIT, "cores": cpu_cores, "kind": "reference",
                "sample": f"full workload once ({cpu_dt:.2f} s): scipy.signal.spectrogram on the [1000,40000] float64 "
                          "batch + mean over sweeps, 1 process, SciPy default 1 FFT thread (the reference's call as-is)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

"""Shared test helpers: tolerances, golden loading, the emulator wrapper."""
from __future__ import annotations

import ast
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")

# ---- tolerances (BASELINE.md section 5 / SURVEY.md 8c) -------------------------------
REL_TOL = 1e-4          # linear power, relative, on bins >= FLOOR * max(S)
FLOOR = 1e-6            # peak - 60 dB
ABS_TOL = 1e-6          # |dS| <= ABS_TOL * max(S) everywhere
DB_TOL = 1e-3           # dB above the floor


def parity_report(S, So, floor=FLOOR):
    """S: engine result (fp32 arithmetic), So: float64 oracle, same shape."""
    S = np.asarray(S, dtype=np.float64)
    So = np.asarray(So, dtype=np.float64)
    assert S.shape == So.shape, (S.shape, So.shape)
    peak = So.max() if So.size else 0.0
    big = (So >= floor * peak) & (So > 0)
    d = np.abs(S - So)
    rel = float((d[big] / So[big]).max()) if big.any() else 0.0
    db = float(np.abs(10 * np.log10(np.maximum(S[big], 1e-300)) - 10 * np.log10(So[big])).max()) if big.any() else 0.0
    return dict(rel=rel, abs=float(d.max() / peak) if peak > 0 else float(d.max()), db=db,
                finite=bool(np.isfinite(S).all()))


def assert_parity(S, So, rel=REL_TOL, floor=FLOOR, abs_tol=ABS_TOL, db=DB_TOL, what="", tail=0.0):
    """The stated bar.  ``tail`` (default 0: every bin) is for multi-million-bin cases
    only: the fraction of above-floor bins allowed between ``rel`` and ``2*rel`` -- the
    worst of ~10^7 bins sitting 60 dB under a tone is a 5-sigma event of the fp32
    rounding noise (DESIGN.md "fp32 floor"); nothing may exceed ``2*rel`` / ``2*db``."""
    r = parity_report(S, So, floor)
    assert r["finite"], f"{what}: non-finite values"
    assert r["abs"] <= abs_tol, f"{what}: abs err {r['abs']:.3e} * max > {abs_tol}"
    if tail > 0.0:
        S64, So64 = np.asarray(S, dtype=np.float64), np.asarray(So, dtype=np.float64)
        big = (So64 >= floor * So64.max()) & (So64 > 0)
        relv = np.abs(S64[big] - So64[big]) / So64[big]
        frac = float(np.mean(relv > rel))
        r["tail_frac"] = frac
        assert frac <= tail, f"{what}: {frac:.2e} of above-floor bins exceed rel {rel} (allowed {tail:.0e})"
        assert r["rel"] <= 2 * rel and r["db"] <= 2 * db, f"{what}: worst bin rel {r['rel']:.3e} / {r['db']:.2e} dB"
        return r
    assert r["rel"] <= rel, f"{what}: rel err {r['rel']:.3e} > {rel} above the {floor:g}*max floor"
    assert r["db"] <= db, f"{what}: dB err {r['db']:.3e} > {db}"
    return r


class _Stream:
    """Streaming max / rms / quantile (log-spaced histogram, 0.25 % resolution) of relative errors."""
    EDGES = np.logspace(-10, -1, 3601)

    def __init__(self):
        self.n, self.sumsq, self.max, self.hist = 0, 0.0, 0.0, np.zeros(3602, dtype=np.int64)

    def add(self, r, rel):
        r = np.asarray(r, dtype=np.float64).ravel()
        self.n += r.size
        self.sumsq += float(np.dot(r, r))
        if r.size:
            self.max = max(self.max, float(r.max()))
        self.hist += np.bincount(np.searchsorted(self.EDGES, r), minlength=3602)
        self.over = getattr(self, "over", 0) + int(np.count_nonzero(r > rel))

    def quantile(self, q):
        c = np.cumsum(self.hist)
        i = int(np.searchsorted(c, q * self.n))
        return float(self.EDGES[min(i, 3600)])                   # upper edge of the bin

    def summary(self):
        return dict(max=self.max, rms=float(np.sqrt(self.sumsq / max(self.n, 1))), p9999=self.quantile(0.9999),
                    frac_over=self.over / max(self.n, 1))


class FullSizeStats:
    """Error statistics of the engine and of SciPy's own float32 pipeline against the float64 oracle,
    accumulated over chunks of a full-size configuration (tests/test_gpu_full_size.py)."""

    def __init__(self, what, rel=REL_TOL):
        self.what, self.rel = what, rel
        self.ours, self.theirs = _Stream(), _Stream()
        self.abs_ours = 0.0
        self.n_bins = 0

    def add(self, S, So, S32, floor=FLOOR):
        So = np.asarray(So, dtype=np.float64)
        assert S.shape == So.shape == S32.shape, (S.shape, So.shape, S32.shape)
        if So.ndim > 2:                           # a batch: every sweep / channel has its own floor
            for i in range(So.shape[0]):
                self.add(S[i], So[i], S32[i], floor)
            return
        S, S32 = np.asarray(S, dtype=np.float64), np.asarray(S32, dtype=np.float64)
        assert np.isfinite(S).all()
        peak = So.max()
        big = (So >= floor * peak) & (So > 0)
        d = np.abs(S - So)
        self.ours.add(d[big] / So[big], self.rel)
        self.theirs.add(np.abs(S32 - So)[big] / So[big], self.rel)
        self.abs_ours = max(self.abs_ours, float(d.max() / peak))
        self.n_bins += So.size

    def check(self, abs_tol=ABS_TOL, tail=1e-5):
        rel = self.rel
        rep = dict(what=self.what, bins=self.n_bins, above_floor=self.ours.n, ours=self.ours.summary(),
                   scipy_f32=self.theirs.summary())
        print("\nFULLSIZE", {k: ({kk: float(f"{vv:.4g}") for kk, vv in v.items()} if isinstance(v, dict) else v)
                             for k, v in rep.items()})
        assert self.abs_ours <= abs_tol, f"{self.what}: abs err {self.abs_ours:.3e} * max"
        assert rep["ours"]["frac_over"] <= tail, f"{self.what}: {rep['ours']['frac_over']:.2e} of the bins miss rel {rel}"
        bar = 1.5 * max(rel, rep["scipy_f32"]["max"])
        assert rep["ours"]["max"] <= bar, f"{self.what}: worst bin {rep['ours']['max']:.3e} > {bar:.3e}"
        assert rep["ours"]["rms"] <= 1.10 * rep["scipy_f32"]["rms"], f"{self.what}: rms {rep}"
        assert rep["ours"]["p9999"] <= 1.15 * rep["scipy_f32"]["p9999"], f"{self.what}: p99.99 {rep}"
        return rep


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["kw"] = ast.literal_eval(str(d["kw"]))
    d["fs"] = float(d["fs"])
    return d


GOLDEN_NAMES = ["ref_call_256", "ref_call_1024", "c1_chirp_025s", "c2_sweeps_4x02s", "int16_512", "clamp_64"]


class Emulator:
    """Runs csrc/b2s_kernels.cuh on the CPU (tests/emu) through the same launch
    planning code as the CUDA library.  Test tooling only."""

    def __init__(self):
        import __graft_entry__ as ge
        # B2S_EMU_LIB: an alternative build of the harness (e.g. -fsanitize=address, see tests/emu/README)
        self.lib = ctypes.CDLL(os.environ.get("B2S_EMU_LIB") or ge.build_emulator())
        c = ctypes
        self.lib.emu_stft_psd.restype = c.c_int
        self.lib.emu_stft_psd.argtypes = [c.c_void_p, c.c_int, c.c_longlong, c.c_longlong, c.c_longlong,
                                          c.c_int, c.c_int, c.c_void_p, c.c_int, c.c_double, c.c_int,
                                          c.c_float, c.c_int, c.c_int, c.c_longlong, c.c_longlong,
                                          c.c_void_p, c.c_longlong, c.c_int, c.c_int, c.c_int]
        self.lib.emu_batch_sum.restype = c.c_int
        self.lib.emu_batch_sum.argtypes = [c.c_void_p, c.c_longlong, c.c_int, c.c_int, c.c_longlong,
                                           c.c_void_p, c.c_float]

    @staticmethod
    def taps(plan, prescale):
        """(fp32 window table, scale) as the host layer passes them (raw taps, true scale), or --
        prescale=True -- taps with sqrt(scale/2) folded in float64 and scale = 2, which the kernels
        take as well (their own fold is then exactly 1; measured: no accuracy gain worth the change)."""
        if prescale:
            return (plan.win64 * np.sqrt(0.5 * plan.scale)).astype(np.float32), 2.0
        return plan.win64.astype(np.float32), plan.scale

    def stft_psd(self, x2d, plan, out_mode=0, db_floor=0.0, kmin=0, kmax=None, frame0=0, nframes=None,
                 grid=2, chunk=0, prescale=False):
        x2d = np.ascontiguousarray(x2d)
        assert x2d.dtype in (np.float32, np.float64) and x2d.ndim == 2
        B, n = x2d.shape
        kmax = plan.nbins - 1 if kmax is None else kmax
        nframes = plan.nframes - frame0 if nframes is None else nframes
        kout = kmax - kmin + 1
        w, scale = self.taps(plan, prescale)
        out = np.full((B, nframes, kout), np.nan, np.float32)
        rc = self.lib.emu_stft_psd(x2d.ctypes.data, int(x2d.dtype == np.float64), B, n, n, plan.nperseg,
                                   plan.hop, w.ctypes.data, plan.detrend, scale, out_mode, db_floor,
                                   kmin, kmax, frame0, nframes, out.ctypes.data, nframes * kout, grid, chunk, 0)
        assert rc == 0, rc
        return out

    def band_power(self, x2d, plan, kmin, kmax, frame0=0, nframes=None, grid=2, chunk=0):
        x2d = np.ascontiguousarray(x2d)
        B, n = x2d.shape
        nframes = plan.nframes - frame0 if nframes is None else nframes
        w, scale = self.taps(plan, False)
        out = np.full((B, nframes), np.nan, np.float32)
        rc = self.lib.emu_stft_psd(x2d.ctypes.data, int(x2d.dtype == np.float64), B, n, n, plan.nperseg,
                                   plan.hop, w.ctypes.data, plan.detrend, scale, 0, 0.0,
                                   kmin, kmax, frame0, nframes, out.ctypes.data, nframes, grid, chunk, 1)
        assert rc == 0, rc
        return out

    FAMILIES = {11: "mixed", 10: "pairq", 9: "pair", 1: "duo256", 2: "duo4", 3: "duo_cta", 4: "duo", 5: "warp", 6: "big", 7: "cta", 8: "dft"}

    def last_family(self):
        """Kernel family the last stft_psd / band_power call ran."""
        self.lib.emu_last_family.restype = ctypes.c_int
        return self.FAMILIES.get(self.lib.emu_last_family())

    def stft_psd_sum(self, x2d, plan, post_scale=1.0, grid=2, max_blocks=64):
        """The sum-fused frame-duo kernel + fold: (rows[B, F, K], sum[F, K], blocks used)."""
        x2d = np.ascontiguousarray(x2d)
        B, n = x2d.shape
        F, K = plan.nframes, plan.nbins
        w, scale = self.taps(plan, False)
        out = np.full((B, F, K), np.nan, np.float32)
        tot = np.full((F, K), np.nan, np.float32)
        c = ctypes
        fn = self.lib.emu_stft_psd_sum
        fn.restype = c.c_int
        fn.argtypes = [c.c_void_p, c.c_int, c.c_longlong, c.c_longlong, c.c_longlong, c.c_int, c.c_int, c.c_void_p,
                       c.c_int, c.c_double, c.c_longlong, c.c_longlong, c.c_void_p, c.c_longlong, c.c_void_p,
                       c.c_float, c.c_int, c.c_int]
        rc = fn(x2d.ctypes.data, int(x2d.dtype == np.float64), B, n, n, plan.nperseg, plan.hop, w.ctypes.data,
                plan.detrend, scale, 0, F, out.ctypes.data, F * K, tot.ctypes.data, post_scale, grid, max_blocks)
        assert rc > 0, rc
        return out, tot, rc

    def batch_sum(self, s, post_scale=1.0, rows_per_slab=64):
        s = np.ascontiguousarray(s, dtype=np.float32)
        B = s.shape[0]
        elems = s[0].size
        slabs = (B + rows_per_slab - 1) // rows_per_slab
        if slabs == 1:
            out = np.empty(s.shape[1:], np.float32)
            self.lib.emu_batch_sum(s.ctypes.data, elems, B, B, elems, out.ctypes.data, post_scale)
            return out
        part = np.empty((slabs,) + s.shape[1:], np.float32)
        self.lib.emu_batch_sum(s.ctypes.data, elems, B, rows_per_slab, elems, part.ctypes.data, 1.0)
        out = np.empty(s.shape[1:], np.float32)
        self.lib.emu_batch_sum(part.ctypes.data, elems, slabs, slabs, elems, out.ctypes.data, post_scale)
        return out

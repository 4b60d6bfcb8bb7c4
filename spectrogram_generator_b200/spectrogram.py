"""Host side of the spectrogram path: SciPy's call surface over the CUDA library.

``spectrogram`` mirrors ``scipy.signal.spectrogram`` (the function the
reference imports at PlotEngine.py:8 and calls at :113 / :232): same signature,
defaults, validation messages, warnings and return layout -- ``(f, t, Sxx)``
NumPy arrays with ``Sxx`` a transposed view of a C-contiguous
``[..., frame, bin]`` buffer, exactly what SciPy hands back.  The arithmetic runs
in ``libb200stft.so`` (fp32, sm_100a); PyTorch is used for device memory,
streams and pinned staging only.  Unsupported keyword combinations raise --
there is no CPU fallback.
"""
from __future__ import annotations

import threading
import warnings
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .windows import frame_count, get_window, rfftfreq, time_axis

_MODES = ["psd", "complex", "magnitude", "angle", "phase"]


# --------------------------------------------------------------------------
# argument triage (scipy/signal/_spectral_py.py:1119-1129, 2209-2230, 2400-2461)
# --------------------------------------------------------------------------

@dataclass
class Plan:
    """Everything the kernel launch and the axis arrays need for one call."""
    n: int                 # samples along the time axis
    nperseg: int
    noverlap: int
    hop: int
    fs: float
    win64: np.ndarray      # float64 window (host)
    scale: float           # 1/(fs*sum(w^2)) or 1/sum(w)^2, float64
    detrend: int           # 0 / 1
    nframes: int
    nbins: int
    window_key: tuple


def _window_key(window, nperseg):
    if isinstance(window, (str, tuple)):
        return ("spec", window, int(nperseg))
    a = np.ascontiguousarray(np.asarray(window, dtype=np.float64))
    return ("array", a.tobytes(), int(nperseg))


def triage(n, fs, window, nperseg, noverlap, nfft, detrend, return_onesided, scaling, mode,
           is_complex=False) -> Plan:
    """Validate like SciPy does and derive the launch plan.  Raises the same
    ``ValueError``s (and the ``nperseg > len(x)`` ``UserWarning``); combinations
    the engine does not implement raise ``NotImplementedError``."""
    if mode not in _MODES:
        raise ValueError(f"unknown value for mode {mode}, must be one of {_MODES}")
    if mode != "psd":
        raise NotImplementedError(f"mode={mode!r}: the B200 engine implements mode='psd' "
                                  "(the reference's only mode, PlotEngine.py:113)")
    if is_complex:
        raise NotImplementedError("complex input is not on the reference's path")
    if not return_onesided:
        raise NotImplementedError("return_onesided=False is not implemented (real input only)")
    # spectrogram() runs _triage_segments first (:1124), so a non-positive or
    # non-integral nperseg surfaces as get_window's ValueError, as in SciPy
    if isinstance(window, (str, tuple)):
        if nperseg is None:
            nperseg = 256
        if nperseg > n:
            warnings.warn(f"nperseg = {nperseg:d} is greater than input length "
                          f" = {n:d}, using nperseg = {n:d}", stacklevel=4)
            nperseg = n
        win = get_window(window, nperseg)          # raises ValueError for nperseg == 0 like SciPy
    else:
        win = np.asarray(window)
        if win.ndim != 1:
            raise ValueError("window must be 1-D")
        if n < win.shape[-1]:
            raise ValueError("window is longer than input signal")
        if nperseg is None:
            nperseg = win.shape[0]
        elif nperseg != win.shape[0]:
            raise ValueError("value specified for nperseg is different from length of window")
        win = np.asarray(win, dtype=np.float64)
    nperseg = int(nperseg)
    if nperseg < 1:                                  # _spectral_helper, :2209-2212
        raise ValueError("nperseg must be a positive integer")
    if nfft is None:
        nfft = nperseg
    elif nfft < nperseg:
        raise ValueError("nfft must be greater than or equal to nperseg.")
    elif int(nfft) != nperseg:
        raise NotImplementedError("nfft != nperseg (zero-padded FFT) is not implemented")
    if noverlap is None:
        noverlap = nperseg // 8                      # spectrogram's default, :1128-1129
    else:
        noverlap = int(noverlap)
    if noverlap >= nperseg:
        raise ValueError("noverlap must be less than nperseg.")
    if not detrend:
        det = 0
    elif detrend == "constant":
        det = 1
    else:
        raise NotImplementedError(f"detrend={detrend!r}: only 'constant' (the reference's) and False")
    if scaling == "density":
        scale = 1.0 / (fs * (win * win).sum())
    elif scaling == "spectrum":
        scale = 1.0 / win.sum() ** 2
    else:
        raise ValueError(f"Unknown scaling: {scaling!r}")
    hop = nperseg - noverlap
    return Plan(n=n, nperseg=nperseg, noverlap=noverlap, hop=hop, fs=fs, win64=win, scale=float(scale),
                detrend=det, nframes=frame_count(n, nperseg, hop), nbins=nperseg // 2 + 1,
                window_key=_window_key(window, nperseg))


# --------------------------------------------------------------------------
# device engine
# --------------------------------------------------------------------------

def _dev_index(device) -> int:
    """Explicit CUDA device index (``torch.device('cuda')`` has index None: the current device)."""
    device = torch.device(device)
    return torch.cuda.current_device() if device.index is None else int(device.index)


class Engine:
    """Per-process state: the loaded library and per-device window tables."""

    _MAX_WINDOWS = 64

    def __init__(self):
        self._lock = threading.Lock()
        self._windows = {}            # insertion-ordered: least recently used first

    @staticmethod
    def require_cuda():
        if not torch.cuda.is_available():
            raise _lib.B2SError("no CUDA device: the spectrogram engine has no CPU fallback")

    def window_table(self, plan: Plan, device: torch.device) -> torch.Tensor:
        idx = _dev_index(device)
        key = (plan.window_key, idx)
        with self._lock:
            t = self._windows.pop(key, None)
            if t is None:
                # evict the least recently used table; kernels still reading it keep it alive through
                # the caching allocator's stream ordering (it was created and is used on this device's streams)
                while len(self._windows) >= self._MAX_WINDOWS:
                    self._windows.pop(next(iter(self._windows)))
                t = torch.from_numpy(plan.win64.astype(np.float32)).to(torch.device("cuda", idx))
            self._windows[key] = t                      # (re)insert as most recently used
            return t

    def stft_psd(self, x: torch.Tensor, plan: Plan, *, out=None, out_mode=0, db_floor=0.0,
                 kmin=0, kmax=None, frame0=0, nframes=None) -> torch.Tensor:
        """x: CUDA tensor [B, n] (float32 or float64, last dim contiguous).
        Returns CUDA float32 [B, nframes, kmax-kmin+1].  Enqueued on the current stream."""
        lib = _lib.load()
        if x.dim() != 2 or not x.is_cuda:
            raise ValueError("stft_psd expects a CUDA tensor of shape [batch, n]")
        if x.stride(1) != 1:
            x = x.contiguous()
        if x.dtype not in (torch.float32, torch.float64):
            raise TypeError("stft_psd expects float32 or float64 samples")
        B, n = x.shape
        if n != plan.n:
            raise ValueError("plan was made for a different signal length")
        kmax = plan.nbins - 1 if kmax is None else int(kmax)
        nframes = plan.nframes - frame0 if nframes is None else int(nframes)
        kout = kmax - kmin + 1
        if out is None:
            out = torch.empty((B, nframes, kout), dtype=torch.float32, device=x.device)
        elif out.shape != (B, nframes, kout) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 [B, nframes, nbins] tensor")
        if B == 0 or nframes == 0:
            return out
        support = lib.b2s_nperseg_support(plan.nperseg)
        if support == 0:
            raise NotImplementedError(
                f"nperseg={plan.nperseg} is not supported by the B200 engine (1..16384)")
        win = self.window_table(plan, x.device)
        fn = lib.b2s_stft_psd_f32 if x.dtype == torch.float32 else lib.b2s_stft_psd_f64
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = fn(x.data_ptr(), B, n, x.stride(0) if B > 1 else n, plan.nperseg, plan.hop,
                    win.data_ptr(), plan.detrend, plan.scale, int(out_mode), float(db_floor),
                    int(kmin), int(kmax), int(frame0), int(nframes), out.data_ptr(),
                    nframes * kout, stream)
        _lib.check(rc, "b2s_stft_psd")
        return out

    def stft_psd_sum(self, x: torch.Tensor, plan: Plan, *, post_scale: float = 1.0, out: torch.Tensor = None,
                     sum_out: torch.Tensor = None):
        """Per-sweep spectrograms and their cross-sweep sum in one pass (BASELINE config 2).
        x: CUDA tensor [B, n]; returns ``(S[B, nframes, nbins], post_scale * S.sum(0))`` -- the
        rows are bit-identical to :meth:`stft_psd`'s, the sum is added in a fixed order.  For the
        shapes of the sum-fused kernels (nperseg 512 with hop 64/128/256, nperseg 256 with any even
        hop, nperseg 1024 with any hop that is a multiple of 4, nperseg 2048 with hop 256/512/1024) the rows are written once and never
        read back; other shapes run :meth:`stft_psd` followed by :meth:`batch_sum` inside the
        library.  ``sum_out``: any contiguous CUDA float32 tensor of nframes*nbins elements."""
        lib = _lib.load()
        if x.dim() != 2 or not x.is_cuda or x.dtype not in (torch.float32, torch.float64):
            raise ValueError("stft_psd_sum expects a CUDA float32/float64 tensor of shape [batch, n]")
        if x.stride(1) != 1:
            x = x.contiguous()
        B, n = x.shape
        if n != plan.n:
            raise ValueError("plan was made for a different signal length")
        if B == 0 or plan.nframes == 0:
            raise ValueError("stft_psd_sum needs at least one sweep and one frame")
        F, K = plan.nframes, plan.nbins
        if out is None:
            out = torch.empty((B, F, K), dtype=torch.float32, device=x.device)
        elif out.shape != (B, F, K) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 [B, nframes, nbins] tensor")
        if sum_out is None:
            sum_out = torch.empty((F, K), dtype=torch.float32, device=x.device)
        elif not (sum_out.is_cuda and sum_out.dtype == torch.float32 and sum_out.is_contiguous()
                  and sum_out.numel() == F * K):
            raise ValueError("sum_out must be a contiguous CUDA float32 tensor of nframes*nbins elements")
        if lib.b2s_nperseg_support(plan.nperseg) == 0:
            raise NotImplementedError(f"nperseg={plan.nperseg} is not supported by the B200 engine (1..16384)")
        win = self.window_table(plan, x.device)
        scratch = self._sum_scratch(lib.b2s_stft_psd_sum_scratch_elems(B, F * K), x.device)
        fn = lib.b2s_stft_psd_sum_f32 if x.dtype == torch.float32 else lib.b2s_stft_psd_sum_f64
        with torch.cuda.device(x.device):
            rc = fn(x.data_ptr(), B, n, x.stride(0) if B > 1 else n, plan.nperseg, plan.hop, win.data_ptr(),
                    plan.detrend, plan.scale, 0, F, out.data_ptr(), F * K, sum_out.data_ptr(), float(post_scale),
                    scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "b2s_stft_psd_sum")
        return out, sum_out

    def _sum_scratch(self, elems: int, device) -> torch.Tensor:
        """Partial-sum scratch of the sum-fused path, kept per device (stream-ordered reuse: the
        kernels that touch it are enqueued on the caller's current stream)."""
        key = (_dev_index(device), torch.cuda.current_stream(device).cuda_stream)
        cache = self.__dict__.setdefault("_scratch", {})
        t = cache.pop(key, None)
        if t is None or t.numel() < elems:
            t = torch.empty(max(int(elems), 1), dtype=torch.float32, device=device)
        while len(cache) >= 16:                         # streams come and go: keep the most recent few
            cache.pop(next(iter(cache)))
        cache[key] = t
        return t

    def band_power(self, x: torch.Tensor, plan: Plan, kmin: int, kmax: int, *, frame0=0, nframes=None) -> torch.Tensor:
        """Fused band-power feature: sum of bins kmin..kmax of every frame, CUDA float32
        [B, nframes]; the spectrogram itself is never written (PlotEngine.py:238-239)."""
        lib = _lib.load()
        if x.dim() != 2 or not x.is_cuda or x.dtype not in (torch.float32, torch.float64):
            raise ValueError("band_power expects a CUDA float32/float64 tensor of shape [batch, n]")
        if x.stride(1) != 1:
            x = x.contiguous()
        B, n = x.shape
        if n != plan.n:
            raise ValueError("plan was made for a different signal length")
        nframes = plan.nframes - frame0 if nframes is None else int(nframes)
        out = torch.empty((B, nframes), dtype=torch.float32, device=x.device)
        if B == 0 or nframes == 0:
            return out
        if lib.b2s_nperseg_support(plan.nperseg) == 0:
            raise NotImplementedError(f"nperseg={plan.nperseg} is not supported by the B200 engine (1..16384)")
        win = self.window_table(plan, x.device)
        fn = lib.b2s_stft_band_power_f32 if x.dtype == torch.float32 else lib.b2s_stft_band_power_f64
        with torch.cuda.device(x.device):
            rc = fn(x.data_ptr(), B, n, x.stride(0) if B > 1 else n, plan.nperseg, plan.hop, win.data_ptr(),
                    plan.detrend, plan.scale, int(kmin), int(kmax), int(frame0), int(nframes), out.data_ptr(),
                    nframes, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "b2s_stft_band_power")
        return out

    def display_scale(self, s: torch.Tensor, log_scale: bool, global_max=None) -> torch.Tensor:
        """PlotEngine.py:126-131 on the device; ``s`` is any contiguous CUDA float32 tensor."""
        lib = _lib.load()
        if not s.is_cuda or s.dtype != torch.float32 or not s.is_contiguous() or s.numel() == 0:
            raise ValueError("display_scale expects a non-empty contiguous CUDA float32 tensor")
        out = torch.empty_like(s)
        scratch = torch.empty(2, dtype=torch.int32, device=s.device)
        gm = float(global_max) if (global_max is not None and global_max > 0) else 0.0
        with torch.cuda.device(s.device):
            rc = lib.b2s_display_scale_f32(s.data_ptr(), s.numel(), int(bool(log_scale)), gm, out.data_ptr(),
                                           scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "b2s_display_scale_f32")
        return out

    def band_sums(self, s: torch.Tensor, ranges) -> np.ndarray:
        """Power summaries on the device (PlotEngine.py:686-719): ``s`` is a contiguous CUDA float32
        ``[frames, bins]`` spectrogram, ``ranges`` up to 16 half-open bin ranges ``(k0, k1)``.  Returns a
        float64 array of ``len(ranges) + 1`` numbers: ``sum(max(0, s[:, k0:k1]))`` per range, then the
        total over all bins.  Only these few numbers cross PCIe."""
        import ctypes
        lib = _lib.load()
        if not s.is_cuda or s.dtype != torch.float32 or not s.is_contiguous() or s.dim() != 2 or s.numel() == 0:
            raise ValueError("band_sums expects a non-empty contiguous CUDA float32 [frames, bins] tensor")
        ranges = [(int(a), int(b)) for a, b in ranges]
        nb = len(ranges)
        k0 = (ctypes.c_int * max(nb, 1))(*[a for a, _ in ranges])
        k1 = (ctypes.c_int * max(nb, 1))(*[b for _, b in ranges])
        out = torch.empty(nb + 1, dtype=torch.float64, device=s.device)
        scratch = torch.empty(int(lib.b2s_band_sums_scratch_elems()), dtype=torch.float64, device=s.device)
        with torch.cuda.device(s.device):
            rc = lib.b2s_band_sums_f32(s.data_ptr(), s.shape[0], s.shape[1], k0, k1, nb, out.data_ptr(),
                                       scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "b2s_band_sums_f32")
        return out.cpu().numpy()

    def batch_sum(self, s: torch.Tensor, post_scale: float = 1.0, out: torch.Tensor = None) -> torch.Tensor:
        """Deterministic sum over dim 0 of a contiguous CUDA float32 [B, ...] tensor (optionally
        into ``out``: any contiguous CUDA float32 tensor with as many elements as one row)."""
        lib = _lib.load()
        if not s.is_cuda or s.dtype != torch.float32 or not s.is_contiguous() or s.dim() < 2:
            raise ValueError("batch_sum expects a contiguous CUDA float32 tensor [B, ...]")
        B = s.shape[0]
        elems = s[0].numel()
        if out is None:
            out = torch.empty(s.shape[1:], dtype=torch.float32, device=s.device)
        elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == elems):
            raise ValueError("batch_sum: out must be a contiguous CUDA float32 tensor with one row's elements")
        ns = lib.b2s_batch_sum_scratch_elems(B, elems)
        scratch = torch.empty(ns, dtype=torch.float32, device=s.device) if ns else None
        with torch.cuda.device(s.device):
            rc = lib.b2s_batch_sum_f32(s.data_ptr(), B, elems, elems, out.data_ptr(),
                                       scratch.data_ptr() if scratch is not None else None,
                                       float(post_scale), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "b2s_batch_sum_f32")
        return out


_ENGINE = Engine()


def engine() -> Engine:
    return _ENGINE


# --------------------------------------------------------------------------
# host staging helpers
# --------------------------------------------------------------------------

def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """A NumPy array backed by page-locked memory.  Signals placed in such an
    array are copied to the GPU asynchronously at full PCIe rate."""
    t = torch.empty(tuple(shape), dtype=torch.from_numpy(np.empty(0, dtype=dtype)).dtype, pin_memory=True)
    return t.numpy()


def _as_host_tensor(a: np.ndarray) -> torch.Tensor:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # read-only arrays: we never write through it
        return torch.from_numpy(a)


def _result_dtype(x: np.ndarray):
    """SciPy's rule: Sxx has the real dtype of result_type(x, complex64) (_spectral_py.py:2169)."""
    return np.float64 if np.result_type(x.dtype, np.complex64) == np.complex128 else np.float32


def _prepare_input(x, axis):
    x = np.asarray(x)
    if np.iscomplexobj(x):
        return x, None, True
    out_dtype = _result_dtype(x)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(out_dtype)
    if x.ndim == 0:
        raise ValueError("input must be at least 1-D")
    axis = int(axis)
    if axis != -1 and axis != x.ndim - 1:
        x = np.moveaxis(x, axis, -1)
    return x, out_dtype, False


def _to_device(x2d: np.ndarray, device) -> torch.Tensor:
    if not x2d.flags.c_contiguous:
        x2d = np.ascontiguousarray(x2d)
    h = _as_host_tensor(x2d)
    B, n = h.shape
    if B > 1 and (n & 1):                       # rows of even stride: sample-pair aligned frames for the packed kernels
        d = torch.empty((B, n + 1), dtype=h.dtype, device=device)[:, :n]
        d.copy_(h, non_blocking=h.is_pinned())
        return d
    return h.to(device, non_blocking=h.is_pinned())


def _to_host(d: torch.Tensor, dtype) -> np.ndarray:
    h = torch.empty(d.shape, dtype=d.dtype, pin_memory=True)
    h.copy_(d, non_blocking=True)
    torch.cuda.current_stream(d.device).synchronize()
    a = h.numpy()
    return a if a.dtype == dtype else a.astype(dtype)


_PIPE_CHUNK_BYTES = 32 << 20       # output bytes per pipeline stage
_PIPE_RAMP_DIV = 16                # the first stage is 1/_PIPE_RAMP_DIV of a full one ...
_PIPE_RAMP_GROWTH = 2.0            # ... and every next one this many times the one before, up to a full stage


class _Streams:
    """Copy-in / copy-out side streams per device, created on first use."""
    _by_dev = {}

    @classmethod
    def get(cls, dev: torch.device):
        idx = _dev_index(dev)
        s = cls._by_dev.get(idx)
        if s is None:
            with torch.cuda.device(idx):
                s = (torch.cuda.Stream(), torch.cuda.Stream())
            cls._by_dev[idx] = s
        return s


def _host_pipeline(x2d: np.ndarray, plan: Plan, dev, out_dtype, *, per_sweep=True, want_sum=False,
                   sum_scale=1.0, kmin=0, kmax=None, out_mode=0, db_floor=0.0, host_out=None):
    """Host arrays in, host arrays out: H2D copy, kernels and D2H copy of successive
    chunks overlap on three streams (PCIe is full duplex), so the end-to-end time tends
    to max(H2D, D2H) instead of their sum.  Chunks are runs of sweeps for batches and
    frame ranges (with their nperseg-hop halo of samples) for a single long recording.
    Returns (S_host [B, F, Kout] or None, sum_scale * sum_dev [F, Kout] or None)."""
    eng = engine()
    B, n = x2d.shape
    kmax = plan.nbins - 1 if kmax is None else kmax
    kout = kmax - kmin + 1
    F = plan.nframes
    if not x2d.flags.c_contiguous:
        x2d = np.ascontiguousarray(x2d)
    h_in = _as_host_tensor(x2d)
    pinned = h_in.is_pinned()
    with torch.cuda.device(dev):
        cur = torch.cuda.current_stream()
        s_in, s_out = _Streams.get(dev)
        # rows of even stride: every frame of every row starts on a sample-pair boundary, which the packed
        # kernels need (free here, the rows are being copied anyway; a device-resident batch with an odd row
        # stride runs the scalar-load kernels instead -- re-striding it costs more than they do: 0.25 vs 0.21 ms
        # for 1000 x 40001 samples)
        x_d = torch.empty((B, n + (n & 1)), dtype=h_in.dtype, device=dev)[:, :n]
        S_d = torch.empty((B, F, kout), dtype=torch.float32, device=dev)
        h_out = None
        if per_sweep:
            if host_out is not None:
                # caller-owned result buffer (ideally page-locked, see pinned_empty): reused across calls
                if host_out.shape != (B, F, kout) or host_out.dtype != np.float32 or not host_out.flags.c_contiguous:
                    raise ValueError(f"out must be a C-contiguous float32 array of shape {(B, F, kout)}")
                h_out = torch.from_numpy(host_out)
            else:
                h_out = torch.empty((B, F, kout), dtype=torch.float32, pin_memory=True)
        row_bytes = F * kout * 4
        # work items: (b0, b1, f0, f1)
        items = []
        # with a sum the chunking depends on the sizes only, so that the order of its additions --
        # and with it every bit of the mean -- does not depend on whether the rows are copied back
        if B * row_bytes <= _PIPE_CHUNK_BYTES or not (per_sweep or want_sum):
            items.append((0, B, 0, F))
        elif B >= 4:
            # the copy back is the long pole (PCIe, more bytes out than in): start it early -- the first
            # chunks are small (1/16 of a stage, doubling), so the D2H stream idles for ~0.05 ms instead of a
            # full stage's copy-in + compute
            step = max(1, _PIPE_CHUNK_BYTES // row_bytes)
            b, cur_step = 0, max(1, step // _PIPE_RAMP_DIV)
            while b < B:
                items.append((b, min(B, b + cur_step), 0, F))
                b += cur_step
                cur_step = min(step, max(cur_step + 1, int(cur_step * _PIPE_RAMP_GROWTH)))
        else:
            fstep = max(1, _PIPE_CHUNK_BYTES // (kout * 4))
            for b in range(B):
                f, cur_step = 0, max(1, fstep // _PIPE_RAMP_DIV)
                while f < F:
                    items.append((b, b + 1, f, min(F, f + cur_step)))
                    f += cur_step
                    cur_step = min(fstep, max(cur_step + 1, int(cur_step * _PIPE_RAMP_GROWTH)))
        if len(items) == 1:
            # one stage: nothing to overlap -- copy in, compute, copy out on the caller's stream
            x_d.copy_(h_in, non_blocking=pinned)
            if want_sum and kout == plan.nbins and out_mode == 0:
                # rows and their cross-sweep sum in one pass (sum-fused kernel where there is one)
                _, total = eng.stft_psd_sum(x_d, plan, post_scale=sum_scale, out=S_d)
            else:
                eng.stft_psd(x_d, plan, out=S_d, kmin=kmin, kmax=kmax, out_mode=out_mode, db_floor=db_floor)
                total = eng.batch_sum(S_d, sum_scale) if want_sum else None
            if per_sweep:
                h_out.copy_(S_d, non_blocking=True)
            cur.synchronize()
            S = None
            if per_sweep:
                S = h_out.numpy()
                if S.dtype != out_dtype:
                    S = S.astype(out_dtype)
            return S, total
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        # chunks of whole sweeps: rows and the chunk's sum in one pass, the chunk sums folded at the end
        fused_sum = (want_sum and kout == plan.nbins and out_mode == 0
                     and all(f0 == 0 and f1 == F for (_, _, f0, f1) in items))
        parts = torch.empty((len(items), F, kout), dtype=torch.float32, device=dev) if fused_sum else None
        for i, (b0, b1, f0, f1) in enumerate(items):
            lo, hi = (0, n) if (f0 == 0 and f1 == F) else (f0 * plan.hop, (f1 - 1) * plan.hop + plan.nperseg)
            with torch.cuda.stream(s_in):
                x_d[b0:b1, lo:hi].copy_(h_in[b0:b1, lo:hi], non_blocking=pinned)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            cur.wait_event(ev_in)
            if fused_sum:
                eng.stft_psd_sum(x_d[b0:b1], plan, out=S_d[b0:b1], sum_out=parts[i])
            elif f0 == 0 and f1 == F:
                eng.stft_psd(x_d[b0:b1], plan, out=S_d[b0:b1], kmin=kmin, kmax=kmax, out_mode=out_mode,
                             db_floor=db_floor)
            else:
                eng.stft_psd(x_d[b0:b1], plan, out=S_d[b0:b1, f0:f1], kmin=kmin, kmax=kmax,
                             out_mode=out_mode, db_floor=db_floor, frame0=f0, nframes=f1 - f0)
            if per_sweep:
                ev_k = torch.cuda.Event()
                ev_k.record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_k)
                    h_out[b0:b1, f0:f1].copy_(S_d[b0:b1, f0:f1], non_blocking=True)
        if fused_sum:
            total = eng.batch_sum(parts, sum_scale)
        else:
            total = eng.batch_sum(S_d, sum_scale) if want_sum else None
        if per_sweep:
            s_out.synchronize()
        cur.synchronize()             # x_d / S_d were used on side streams: keep them alive until here
    S = None
    if per_sweep:
        S = h_out.numpy()
        if S.dtype != out_dtype:
            S = S.astype(out_dtype)
    return S, total


# --------------------------------------------------------------------------
# public API
# --------------------------------------------------------------------------

def spectrogram(x, fs=1.0, window=("tukey", .25), nperseg=None, noverlap=None, nfft=None,
                detrend="constant", return_onesided=True, scaling="density", axis=-1, mode="psd",
                *, device=None):
    """Drop-in for ``scipy.signal.spectrogram`` on the path the reference uses.

    Returns ``(f, t, Sxx)``: ``f`` float64 ``(nperseg//2+1,)``, ``t`` float64
    ``(n_frames,)`` (both bit-identical to SciPy's), ``Sxx`` of shape
    ``(..., n_bins, n_frames)`` with SciPy's dtype rule (float32 for
    float32/int16 input, float64 for float64 input -- the arithmetic is fp32 in
    both cases).
    """
    x, out_dtype, is_complex = _prepare_input(x, axis)
    plan = triage(x.shape[-1], fs, window, nperseg, noverlap, nfft, detrend, return_onesided,
                  scaling, mode, is_complex)
    f = rfftfreq(plan.nperseg, fs)
    t = time_axis(plan.n, plan.nperseg, plan.noverlap, fs)
    lead = x.shape[:-1]
    B = int(np.prod(lead)) if lead else 1
    eng = engine()
    eng.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if B == 0 or plan.nframes == 0:
        S = np.empty(lead + (plan.nframes, plan.nbins), dtype=out_dtype)
    else:
        S, _ = _host_pipeline(x.reshape(B, plan.n), plan, dev, out_dtype)
        S = S.reshape(lead + (plan.nframes, plan.nbins))
    # SciPy rolls the frequency axis back to where the data axis was and leaves
    # the segment-time axis last (:2334-2341); for axis=-1 that is the
    # [..., bin, frame] transposed view of the [..., frame, bin] buffer.
    ax = int(axis)
    if ax < 0:
        ax -= 1
    return f, t, np.moveaxis(S, -1, ax)


def spectrogram_batch(x, fs=1.0, **kw):
    """Per-sweep spectrograms of a batch ``x[B, N]`` -> ``(f, t, Sxx[B, K, F])``
    (same as ``spectrogram`` on a 2-D array; named for BASELINE config 2)."""
    x = np.asarray(x)
    if x.ndim != 2:
        raise ValueError("spectrogram_batch expects x of shape [B, N]")
    return spectrogram(x, fs=fs, **kw)


def mean_spectrogram(x, fs=1.0, window=("tukey", .25), nperseg=None, noverlap=None, nfft=None,
                     detrend="constant", return_onesided=True, scaling="density", mode="psd",
                     *, return_per_sweep=False, device=None, out=None, group=None, total_sweeps=None):
    """Cross-sweep mean spectrogram of ``x[B, N]``: ``mean_b Sxx_b`` (BASELINE
    config 2; the reference has no code for it -- SURVEY.md 8 a-15).  The sum
    over sweeps runs on the device in a fixed order.  Returns ``(f, t, Smean[K, F])``
    or, with ``return_per_sweep``, ``(f, t, Smean, Sxx[B, K, F])``.

    ``out``: a caller-owned float32 ``[B, F, K]`` result buffer for the per-sweep spectrograms (page-locked
    memory from :func:`pinned_empty` makes the copy back asynchronous); reuse it across calls instead
    of paying for a fresh 318 MB pinned allocation every time.  ``group`` / ``total_sweeps``: ``x`` holds
    this rank's sweeps of a batch sharded over the ranks of a ``torch.distributed`` group; the partial
    sums are all-reduced before the division, every rank returns the global mean."""
    x, out_dtype, is_complex = _prepare_input(x, -1)
    if x.ndim != 2:
        raise ValueError("mean_spectrogram expects x of shape [B, N]")
    plan = triage(x.shape[-1], fs, window, nperseg, noverlap, nfft, detrend, return_onesided,
                  scaling, mode, is_complex)
    if x.shape[0] == 0 or plan.nframes == 0:
        raise ValueError("mean_spectrogram needs at least one sweep and one frame")
    f = rfftfreq(plan.nperseg, fs)
    t = time_axis(plan.n, plan.nperseg, plan.noverlap, fs)
    eng = engine()
    eng.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    import torch.distributed as dist
    sharded = group is not None or (total_sweeps is not None and dist.is_available() and dist.is_initialized())
    n_total = int(total_sweeps) if total_sweeps is not None else x.shape[0]
    S, total = _host_pipeline(x, plan, dev, out_dtype, per_sweep=return_per_sweep, want_sum=True,
                              sum_scale=1.0 / n_total, host_out=out)
    with torch.cuda.device(dev):
        if sharded and dist.get_world_size(group) > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)      # sum of partial means
        mean = np.moveaxis(_to_host(total, out_dtype), -1, -2)
    if return_per_sweep:
        return f, t, mean, np.moveaxis(S, -1, -2)
    return f, t, mean


def split_frames(nframes: int, parts: int):
    """Contiguous frame ranges ``[(f0, count), ...]`` for ``parts`` chunks / ranks."""
    parts = max(1, int(parts))
    base, rem = divmod(nframes, parts)
    out, f0 = [], 0
    for r in range(parts):
        c = base + (1 if r < rem else 0)
        out.append((f0, c))
        f0 += c
    return out


def spectrogram_chunked(x, fs=1.0, window=("tukey", .25), nperseg=None, noverlap=None, nfft=None,
                        detrend="constant", return_onesided=True, scaling="density", mode="psd",
                        *, n_chunks=8, device=None):
    """Time-chunked spectrogram of one long 1-D recording (BASELINE config 3 on
    one GPU): the frame range is cut into ``n_chunks`` contiguous ranges; chunk r
    is given only its own samples ``[f0*hop, (f0+c-1)*hop + nperseg)`` -- the
    ``nperseg - hop`` halo it shares with its neighbour is read-only input, so no
    exchange is needed -- and the results are concatenated.  Bit-identical to the
    unchunked call (the per-frame arithmetic does not depend on the chunking)."""
    x, out_dtype, is_complex = _prepare_input(x, -1)
    if x.ndim != 1:
        raise ValueError("spectrogram_chunked expects a 1-D recording")
    plan = triage(x.shape[-1], fs, window, nperseg, noverlap, nfft, detrend, return_onesided,
                  scaling, mode, is_complex)
    f = rfftfreq(plan.nperseg, fs)
    t = time_axis(plan.n, plan.nperseg, plan.noverlap, fs)
    eng = engine()
    eng.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    S = np.empty((plan.nframes, plan.nbins), dtype=out_dtype)
    with torch.cuda.device(dev):
        for f0, c in split_frames(plan.nframes, n_chunks):
            if c == 0:
                continue
            lo, hi = f0 * plan.hop, (f0 + c - 1) * plan.hop + plan.nperseg
            sub = Plan(**{**plan.__dict__, "n": hi - lo, "nframes": c})
            xd = _to_device(x[lo:hi].reshape(1, -1), dev)
            S[f0:f0 + c] = _to_host(eng.stft_psd(xd, sub)[0], out_dtype)
    return f, t, S.T

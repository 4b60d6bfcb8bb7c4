"""Host-side mirror of the reference's ``PlotEngine`` spectrogram path.

``SpectrogramPath`` carries the compute half of the ``PlotEngine`` methods that
sit on the hot path (/root/reference/PlotEngine.py) with the same names,
argument meaning and state (``last_f / last_t / last_Sxx``), minus everything
that draws:

  * ``_plot_spectrogram``      PlotEngine.py:110-131 (call, band mask, display scaling)
  * ``_calculate_features``    PlotEngine.py:229-242 (band power, log10, first difference)
  * ``calculate_absolute_power`` / ``calculate_band_powers``   PlotEngine.py:686-719
  * ``combine``                PlotEngine.py:162-200 (time concatenation + segment map)

The spectrogram itself is computed on the GPU; the band mask
``(f >= fmin) & (f <= fmax)`` becomes a bin crop inside the kernel's store, so
only the kept bins are written and copied back.
"""
from __future__ import annotations

import numpy as np
import torch

from .spectrogram import _as_host_tensor, _prepare_input, _to_device, _to_host, engine, triage
from .windows import rfftfreq, time_axis

DEFAULT_BANDS = {
    # PlotEngine.py:698-706
    "Delta (δ)": (0, 4), "Theta (θ)": (4, 8),
    "Alpha (α)": (8, 13), "Beta (β)": (13, 30),
    "Gamma (γ)": (30, 80), "HFO (ripples)": (80, 250),
}


def band_to_bins(f: np.ndarray, fmin: float, fmax: float):
    """Bin range selected by the reference's mask ``(f >= fmin) & (f <= fmax)``
    (PlotEngine.py:114).  ``f`` is increasing, so the mask is a contiguous range;
    returns ``(kmin, kmax)`` inclusive, or ``None`` when the mask is empty."""
    idx = np.nonzero((f >= fmin) & (f <= fmax))[0]
    if idx.size == 0:
        return None
    return int(idx[0]), int(idx[-1])


class SpectrogramPath:
    def __init__(self, device=None):
        self.device = device
        self.spec_data_source = None
        self.last_fs = None
        self.last_settings = None
        self.last_t = np.array([])
        self.last_f = None
        self.last_Sxx = None
        self.segment_map = []
        self._last_Sxx_dev = None
        self._combined = None            # (host array returned by combine(), its copy on the device)

    def _dev(self):
        return torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)

    def _device_row(self, data, x, dev):
        """The 1-D signal as a [1, n] device tensor.  The array combine() returned was assembled on the device
        sweep by sweep while the host concatenated: it is not uploaded a second time."""
        if self._combined is not None and data is self._combined[0] and self._combined[1].device == dev:
            return self._combined[1].reshape(1, -1)
        return _to_device(x.reshape(1, -1), dev)

    # -- the call at PlotEngine.py:113 / :232, with the band mask fused as a crop --
    def _spectrogram_cropped(self, data, fs, nperseg, fmin, fmax):
        self._last_Sxx_dev = None
        x, out_dtype, is_complex = _prepare_input(data, -1)
        if x.ndim != 1:
            raise ValueError("the reference passes one 1-D sweep per call")
        plan = triage(x.shape[-1], fs, ("tukey", .25), nperseg, None, None, "constant", True,
                      "density", "psd", is_complex)
        f = rfftfreq(plan.nperseg, fs)
        t = time_axis(plan.n, plan.nperseg, plan.noverlap, fs)
        bins = band_to_bins(f, fmin, fmax)
        if bins is None or plan.nframes == 0:
            return f[:0], t, np.empty((0, plan.nframes), dtype=out_dtype)
        eng = engine()
        eng.require_cuda()
        dev = self._dev()
        with torch.cuda.device(dev):
            xd = self._device_row(data, x, dev)
            Sd = eng.stft_psd(xd, plan, kmin=bins[0], kmax=bins[1])
            self._last_Sxx_dev = Sd[0]                       # [frame][bin], kept for the display scaling
            S = _to_host(Sd[0], out_dtype)
        return f[bins[0]:bins[1] + 1], t, S.T

    def _plot_spectrogram(self, data, fs, settings, global_max=None):
        """PlotEngine.py:110-131.  Returns the normalised image the reference hands
        to ``pcolormesh`` (``None`` when the band mask is empty)."""
        nperseg, fmin, fmax, log_scale = (settings["nperseg"], settings["fmin"], settings["fmax"],
                                          settings["log_scale"])
        f, t, Sxx = self._spectrogram_cropped(data, fs, nperseg, fmin, fmax)
        self.last_f = f.copy()
        self.last_t = t.copy()
        self.last_Sxx = Sxx.copy()
        if Sxx.size == 0:
            self.last_t = np.array([])
            return None
        # display scaling (PlotEngine.py:126-131) on the device: global max/min reduction,
        # clip, 10*log10(. + 1e-12), min-max -- the image comes back already normalised
        dev_S = self._last_Sxx_dev
        with torch.cuda.device(dev_S.device):
            img = engine().display_scale(dev_S, bool(log_scale), global_max)
            return _to_host(img, Sxx.dtype).T

    def plot_extra(self, signal_raw, signal_proc, fs, settings, global_max=None):
        """Source selection of PlotEngine.plot_extra (PlotEngine.py:95-105)."""
        self.last_fs = fs
        source = None
        if settings["mode_proc"] in ["Spectrogram", "Both"] and signal_proc is not None:
            source = signal_proc
        elif settings["mode_raw"] in ["Spectrogram", "Both"] and signal_raw is not None:
            source = signal_raw
        if source is None:
            return None
        self.spec_data_source = source
        self.last_fs = fs
        self.last_settings = settings
        return self._plot_spectrogram(source, fs, settings, global_max)

    def _calculate_features(self, signal, fs=None, settings=None):
        """PlotEngine.py:229-242."""
        fs = fs or self.last_fs
        settings = settings or self.last_settings
        x, out_dtype, is_complex = _prepare_input(signal, -1)
        if x.ndim != 1:
            raise ValueError("the reference passes one 1-D sweep per call")
        plan = triage(x.shape[-1], fs, ("tukey", .25), settings["nperseg"], None, None, "constant", True,
                      "density", "psd", is_complex)
        f = rfftfreq(plan.nperseg, fs)
        t = time_axis(plan.n, plan.nperseg, plan.noverlap, fs)
        if t.size == 0:
            return None, None
        bins = band_to_bins(f, settings["fmin"], settings["fmax"])
        if bins is None:
            # an empty band mask gives a sum over zero bins == 0 for every frame
            power_feature = np.zeros(t.shape, dtype=out_dtype)
        else:
            eng = engine()
            eng.require_cuda()
            dev = torch.device("cuda", torch.cuda.current_device()) if self.device is None \
                else torch.device(self.device)
            with torch.cuda.device(dev):
                # fused epilogue: only F numbers leave the kernel (no [F][K] spectrogram)
                band = eng.band_power(self._device_row(signal, x, dev), plan, bins[0], bins[1])
                power_feature = _to_host(band[0], out_dtype)
        log_power = np.log10(power_feature + 1e-20)
        delta_log_power = np.diff(log_power, prepend=log_power[0])
        return t, np.column_stack([log_power, delta_log_power])

    def _band_ranges(self, bands):
        """Half-open bin ranges of ``last_f`` selected by the reference's masks ``(f >= low) & (f < high)``
        (PlotEngine.py:714); ``last_f`` is increasing, so every mask is a contiguous range."""
        out = []
        for low, high in bands.values():
            idx = np.nonzero((self.last_f >= low) & (self.last_f < high))[0]
            out.append((int(idx[0]), int(idx[-1]) + 1) if idx.size else (0, 0))
        return out

    def calculate_absolute_power(self):
        """PlotEngine.py:686-690: ``sum(last_Sxx)``, reduced on the device from the spectrogram the last
        ``_plot_spectrogram`` call left there (``b2s_band_sums_f32``; double accumulators)."""
        if self.last_Sxx is None:
            return None
        if self._last_Sxx_dev is None or self.last_Sxx.size == 0:
            return np.sum(self.last_Sxx)
        # the reference sums the signed values; the kernel clamps at 0 like calculate_band_powers -- a PSD has none
        return self.last_Sxx.dtype.type(engine().band_sums(self._last_Sxx_dev, [])[-1])

    def calculate_band_powers(self, bands=None):
        """PlotEngine.py:692-719, the sums taken on the device: only ``len(bands) + 1`` numbers are read back."""
        if self.last_Sxx is None or self.last_f is None:
            return None
        if bands is None:
            bands = DEFAULT_BANDS
        if self._last_Sxx_dev is None or self.last_Sxx.size == 0 or len(bands) > 16:
            Sxx_linear = np.maximum(0, self.last_Sxx)
            total_power = np.sum(Sxx_linear)
            sums = [np.sum(Sxx_linear[(self.last_f >= low) & (self.last_f < high), :]) for low, high in bands.values()]
        else:
            res = engine().band_sums(self._last_Sxx_dev, self._band_ranges(bands))
            sums, total_power = res[:-1], res[-1]
        if total_power < 1e-18:
            return {name: 0.0 for name in bands}
        return {name: np.clip(band_power / total_power, 0.0, None) for name, band_power in zip(bands, sums)}

    def combine(self, sweeps_info, settings):
        """'Combine all sweeps': time concatenation with a segment map
        (PlotEngine.py:162-200).  Returns the concatenated signal (the reference plots it as the trace).
        When a GPU is there the sweeps are also copied, one by one, to their offsets in ONE device buffer
        (asynchronously where they are page-locked, beside the host's ``np.concatenate``); handing the returned
        array to ``_plot_spectrogram`` / ``_calculate_features`` then uses that buffer instead of uploading the
        concatenated signal again.  Same samples, same kernels: bit-identical to the upload."""
        self._combined = None
        self.segment_map = []
        offset, parts = 0.0, []
        use_proc = settings.get("draw_proc", True)
        for info in sweeps_info:
            sig_raw = info["signal_raw"]
            sig_proc = info["signal_proc"] if info["signal_proc"] is not None else info["signal_raw"]
            sig = sig_proc if use_proc else sig_raw
            if sig is None:
                continue
            duration = len(sig) / info["fs"]
            self.segment_map.append({"start_time_combined": offset,
                                     "end_time_combined": offset + duration,
                                     "source_item": info.get("item")})
            parts.append(sig)
            offset += duration
        if not parts:
            return None
        staged = None
        if torch.cuda.is_available():
            arrs = [np.asarray(q) for q in parts]
            dt = np.result_type(*arrs)
            if dt in (np.float32, np.float64) and all(a.ndim == 1 for a in arrs):
                dev = self._dev()
                total = sum(a.shape[0] for a in arrs)
                with torch.cuda.device(dev):
                    staged = torch.empty((total + (total & 1),), dtype=torch.from_numpy(np.empty(0, dt)).dtype, device=dev)[:total]
                    at = 0
                    for a in arrs:
                        h = _as_host_tensor(np.ascontiguousarray(a, dtype=dt))
                        staged[at:at + a.shape[0]].copy_(h, non_blocking=h.is_pinned())
                        at += a.shape[0]
        final = np.concatenate(parts)
        if staged is not None:
            self._combined = (final, staged)
        return final

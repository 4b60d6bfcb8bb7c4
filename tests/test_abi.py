"""The C-ABI library loads without a GPU and exports every symbol include/b2s.h declares."""
import ctypes
import re

import numpy as np
import pytest

from spectrogram_generator_b200 import _lib


def declared_symbols():
    src = open(_lib.HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    _lib.build()
    lib = _lib.load()
    names = declared_symbols()
    assert "b2s_stft_psd_f32" in names and len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b2s.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)


def test_host_only_entries():
    lib = _lib.load()
    assert lib.b2s_version() == 1
    for n, want in [(32, 1), (512, 1), (16384, 1), (16, 2), (1000, 3), (8191, 2), (96, 2), (352, 3), (34, 2), (1, 2), (0, 0), (32768, 0)]:
        assert lib.b2s_nperseg_support(n) == want
    assert lib.b2s_frame_count(40000, 512, 128) == 309
    assert lib.b2s_frame_count(441000, 1024, 256) == 1719
    assert lib.b2s_frame_count(100, 512, 128) == 0
    assert lib.b2s_batch_sum_scratch_elems(1000, 79413) == 8 * 79413      # slabs of 128 sweeps
    assert lib.b2s_batch_sum_scratch_elems(128, 10) == 0
    # the sum-fused call: up to 64 sweep blocks of partial sums, never less than the two-pass sum needs
    assert lib.b2s_stft_psd_sum_scratch_elems(1000, 79413) == 64 * 79413
    assert lib.b2s_stft_psd_sum_scratch_elems(3, 10) == 30
    assert lib.b2s_stft_psd_sum_scratch_elems(20000, 10) == 157 * 10
    assert lib.b2s_stft_psd_sum_scratch_elems(0, 10) == 0


def test_bad_arguments_are_reported_before_any_device_work():
    lib = _lib.load()
    x = np.zeros(1024, np.float32)
    w = np.ones(256, np.float32)
    o = np.zeros(10 * 129, np.float32)
    def call(**kw):
        a = dict(x=x.ctypes.data, batch=1, n=1024, xs=1024, nperseg=256, hop=128, w=w.ctypes.data, det=1,
                 scale=1.0, mode=0, floor=0.0, kmin=0, kmax=128, f0=0, nf=7, out=o.ctypes.data, os=7 * 129)
        a.update(kw)
        return lib.b2s_stft_psd_f32(a["x"], a["batch"], a["n"], a["xs"], a["nperseg"], a["hop"], a["w"],
                                    a["det"], a["scale"], a["mode"], a["floor"], a["kmin"], a["kmax"],
                                    a["f0"], a["nf"], a["out"], a["os"], None)
    assert call(hop=0) == _lib.B2S_ERR_BAD_ARG
    assert call(nf=8) == _lib.B2S_ERR_BAD_ARG and b"frame range" in lib.b2s_last_error()
    assert call(kmax=129) == _lib.B2S_ERR_BAD_ARG
    assert call(nperseg=40000) == _lib.B2S_ERR_UNSUPPORTED
    assert call(x=None) == _lib.B2S_ERR_BAD_ARG
    with pytest.raises(ValueError):
        _lib.check(call(kmin=5, kmax=4), "b2s_stft_psd_f32")
    with pytest.raises(NotImplementedError):
        _lib.check(call(nperseg=40000), "b2s_stft_psd_f32")


def test_sum_call_reports_bad_arguments_before_any_device_work():
    """b2s_stft_psd_sum_f32: same validation as b2s_stft_psd_f32, plus its own buffers."""
    lib = _lib.load()
    x = np.zeros(2 * 1024, np.float32)
    w = np.ones(512, np.float32)
    o = np.zeros(2 * 5 * 257, np.float32)
    s = np.zeros(5 * 257, np.float32)
    scr = np.zeros(2 * 5 * 257, np.float32)

    def call(**kw):
        a = dict(x=x.ctypes.data, batch=2, n=1024, xs=1024, nperseg=512, hop=128, w=w.ctypes.data, det=1,
                 scale=1.0, f0=0, nf=5, out=o.ctypes.data, os=5 * 257, s=s.ctypes.data, ps=0.5, scr=scr.ctypes.data)
        a.update(kw)
        return lib.b2s_stft_psd_sum_f32(a["x"], a["batch"], a["n"], a["xs"], a["nperseg"], a["hop"], a["w"], a["det"],
                                        a["scale"], a["f0"], a["nf"], a["out"], a["os"], a["s"], a["ps"], a["scr"], None)
    assert call(hop=0) == _lib.B2S_ERR_BAD_ARG
    assert call(nf=6) == _lib.B2S_ERR_BAD_ARG and b"frame range" in lib.b2s_last_error()
    assert call(s=None) == _lib.B2S_ERR_BAD_ARG and b"sum_out" in lib.b2s_last_error()
    assert call(scr=None) == _lib.B2S_ERR_BAD_ARG
    assert call(x=None) == _lib.B2S_ERR_BAD_ARG
    assert call(nperseg=40000) == _lib.B2S_ERR_UNSUPPORTED

"""The kernel source itself (csrc/b2s_kernels.cuh), compiled for the CPU SIMT
emulator (tests/emu) and driven through the library's own launch planning,
compared against the oracle.  This exercises the index math, shuffles, barrier
discipline, bin crop, frame ranges and the unaligned load path in the no-GPU
container; the GPU tests repeat the comparison on the real device."""
import numpy as np
import pytest

import spectrogram_generator_b200 as sg
from oracle import stft_oracle
from util import assert_parity, load_golden


def plan_for(n, fs, **kw):
    a = dict(window=("tukey", .25), nperseg=None, noverlap=None, nfft=None, detrend="constant",
             return_onesided=True, scaling="density", mode="psd")
    a.update(kw)
    return sg.triage(n, fs, a["window"], a["nperseg"], a["noverlap"], a["nfft"], a["detrend"],
                     a["return_onesided"], a["scaling"], a["mode"])


def signal(B, n, seed, dc=0.25):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    x = 0.3 * rng.standard_normal((B, n)) + np.sin(2 * np.pi * 0.0371 * t) + 0.5 * np.sin(2 * np.pi * 0.21 * t + 1.0) + dc
    return x.astype(np.float32)


@pytest.mark.parametrize("nperseg", [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
def test_all_sizes_reference_default_overlap(emu, nperseg):
    """SciPy defaults as the reference uses them: Tukey(.25), noverlap = nperseg//8, detrend."""
    n = nperseg * 4 + 13
    x = signal(2, n, nperseg)
    plan = plan_for(n, 1000.0, nperseg=nperseg)
    got = emu.stft_psd(x, plan, chunk=2)
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=1000.0, nperseg=nperseg)
    assert_parity(np.moveaxis(got, -1, -2), So, what=f"nperseg={nperseg}")


@pytest.mark.parametrize("nperseg,hop", [(512, 128), (512, 256), (512, 64), (1024, 256), (256, 37),
                                         (256, 1), (64, 64), (2048, 512), (1024, 333)])
@pytest.mark.parametrize("detrend", ["constant", False])
def test_hops_and_detrend(emu, nperseg, hop, detrend):
    n = nperseg + hop * 9 + 5
    x = signal(3, n, hop, dc=2.0 if detrend else 0.0)
    kw = dict(window="hann", nperseg=nperseg, noverlap=nperseg - hop, detrend=detrend)
    plan = plan_for(n, 20000.0, **kw)
    got = emu.stft_psd(x, plan)
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=20000.0, **kw)
    assert_parity(np.moveaxis(got, -1, -2), So, what=str(kw))


def test_float64_samples_and_spectrum_scaling(emu):
    n = 5000
    x = signal(2, n, 7).astype(np.float64)       # fp32-representable values held as float64
    kw = dict(window="blackman", nperseg=512, noverlap=100, scaling="spectrum")
    plan = plan_for(n, 48000.0, **kw)
    got = emu.stft_psd(x, plan)
    _, _, So = stft_oracle.spectrogram(x, fs=48000.0, **kw)
    assert_parity(np.moveaxis(got, -1, -2), So)


def test_unaligned_base_and_odd_stride_take_the_scalar_path(emu):
    n = 3001
    big = signal(1, 2 * n + 8, 11)[0]
    x = np.stack([big[1:1 + n], big[n + 2:2 * n + 2]])         # odd offsets -> 4-byte aligned only
    xs = np.lib.stride_tricks.as_strided(big[1:], shape=(2, n), strides=((n + 1) * 4, 4))
    assert np.array_equal(x, xs)
    kw = dict(window="hann", nperseg=256, noverlap=192)
    plan = plan_for(n, 1.0, **kw)
    got = emu.stft_psd(np.ascontiguousarray(xs), plan)
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=1.0, **kw)
    assert_parity(np.moveaxis(got, -1, -2), So)


def test_bin_crop_frame_range_and_db(emu):
    n = 9000
    x = signal(2, n, 5)
    kw = dict(window="hann", nperseg=1024, noverlap=768)
    plan = plan_for(n, 8000.0, **kw)
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=8000.0, **kw)
    So = np.moveaxis(So, -1, -2)
    full = emu.stft_psd(x, plan)
    part = emu.stft_psd(x, plan, kmin=17, kmax=300, frame0=5, nframes=11)
    assert np.array_equal(part, full[:, 5:16, 17:301])          # crop/range do not change the arithmetic
    floor = 1e-6 * So.max()
    db = emu.stft_psd(x, plan, out_mode=1, db_floor=floor)
    want = stft_oracle.to_db(So, floor)
    big = So >= floor
    assert np.max(np.abs(db[big] - want[big])) <= 1e-3
    assert np.all(db >= 10 * np.log10(floor) - 1e-3)


def test_chunking_is_bit_identical(emu):
    n = 20000
    x = signal(1, n, 3)
    kw = dict(window="hann", nperseg=512, noverlap=384)
    plan = plan_for(n, 1.0, **kw)
    a = emu.stft_psd(x, plan, chunk=1, grid=1)
    b = emu.stft_psd(x, plan, chunk=7, grid=3)
    assert np.array_equal(a, b)
    # time-chunking with a halo: each chunk sees only its own samples
    pieces = []
    for f0, c in sg.split_frames(plan.nframes, 4):
        lo, hi = f0 * plan.hop, (f0 + c - 1) * plan.hop + plan.nperseg
        sub = plan_for(hi - lo, 1.0, **kw)
        pieces.append(emu.stft_psd(x[:, lo:hi], sub))
    assert np.array_equal(np.concatenate(pieces, axis=1), a)


@pytest.mark.parametrize("name", ["ref_call_256", "ref_call_1024", "c1_chirp_025s", "c2_sweeps_4x02s", "clamp_64"])
def test_golden_fixtures(emu, name):
    import warnings
    g = load_golden(name)
    x = np.atleast_2d(g["x"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        plan = plan_for(x.shape[-1], g["fs"], **g["kw"])
    got = np.moveaxis(emu.stft_psd(x, plan), -1, -2)
    assert_parity(got.reshape(g["Sxx"].shape), g["Sxx"], what=name)
    if "mean" in g:
        m = emu.batch_sum(np.moveaxis(got, -1, -2), 1.0 / x.shape[0]).T
        assert np.max(np.abs(m - g["mean"])) <= 1e-6 * g["mean"].max()


def test_batch_sum_is_deterministic_and_slabbed(emu):
    rng = np.random.default_rng(0)
    s = rng.random((200, 7, 33)).astype(np.float32)
    a = emu.batch_sum(s, 1.0 / 200)
    b = emu.batch_sum(s, 1.0 / 200)
    assert np.array_equal(a, b)
    np.testing.assert_allclose(a, s.astype(np.float64).mean(axis=0), rtol=2e-6)


def test_parseval_and_linearity(emu):
    """Size-independent properties: with detrend=False the one-sided density
    integrates to the windowed frame energy; the power scales with amplitude^2."""
    n = 6000
    x = signal(1, n, 9, dc=0.0)
    kw = dict(window="hann", nperseg=256, noverlap=128, detrend=False)
    fs = 1234.0
    plan = plan_for(n, fs, **kw)
    S = emu.stft_psd(x, plan)[0].astype(np.float64)                     # [F, K]
    w = plan.win64
    frames = np.lib.stride_tricks.sliding_window_view(x[0].astype(np.float64), 256)[::128]
    energy = ((frames * w) ** 2).sum(axis=1)
    lhs = S.sum(axis=1) * fs / 256 * (w * w).sum()
    np.testing.assert_allclose(lhs, energy, rtol=2e-6)
    S4 = emu.stft_psd(2.0 * x, plan)[0]
    np.testing.assert_array_equal(S4, 4.0 * emu.stft_psd(x, plan)[0])   # exact: power-of-two scaling


@pytest.mark.parametrize("nperseg,hop", [(1000, 875), (96, 84), (100, 25), (33, 7), (8, 4), (1, 1), (250, 250)])
def test_direct_dft_kernel_any_nperseg(emu, nperseg, hop):
    """Lengths the radix-16 kernels do not take (GUI spin box: any integer 32..8192;
    SciPy's clamp to len(x); SciPy's own tests use nperseg=8) run on the direct-DFT kernel."""
    n = nperseg + hop * 6 + 3
    x = signal(2, n, nperseg + hop, dc=1.5)
    kw = dict(window=("tukey", .25), nperseg=nperseg, noverlap=nperseg - hop)
    plan = plan_for(n, 777.0, **kw)
    got = emu.stft_psd(x, plan)
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=777.0, **kw)
    assert_parity(np.moveaxis(got, -1, -2), So, what=f"dft nperseg={nperseg}")
    if nperseg >= 8:
        part = emu.stft_psd(x, plan, kmin=1, kmax=nperseg // 4, frame0=2, nframes=3)
        assert np.array_equal(part, got[:, 2:5, 1:nperseg // 4 + 1])


def test_direct_dft_known_answer_welch(emu):
    """scipy/signal/tests/test_spectral.py:246-256 through the kernel (nperseg = 8)."""
    x = np.zeros((1, 16), np.float32)
    x[0, 0] = 1
    x[0, 8] = 1
    plan = plan_for(16, 1.0, window="hann", nperseg=8, noverlap=4)
    S = emu.stft_psd(x, plan)[0]
    q = np.array([0.08333333, 0.15277778, 0.22222222, 0.22222222, 0.11111111])
    np.testing.assert_allclose(S.mean(axis=0), q, atol=1e-7, rtol=1e-6)


@pytest.mark.parametrize("nperseg", [64, 512, 1024, 2048, 1000])
def test_fused_band_power(emu, nperseg):
    """Band-power epilogue == sum over the cropped spectrogram (PlotEngine.py:238-239)."""
    n = nperseg * 5 + 17
    x = signal(2, n, nperseg + 1, dc=0.5)
    plan = plan_for(n, 1000.0, nperseg=nperseg)              # reference call: default overlap
    kmin, kmax = 1, max(2, nperseg // 8)
    full = emu.stft_psd(x, plan).astype(np.float64)
    band = emu.band_power(x, plan, kmin, kmax)
    np.testing.assert_allclose(band, full[:, :, kmin:kmax + 1].sum(axis=-1), rtol=2e-6)
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=1000.0, nperseg=nperseg)
    np.testing.assert_allclose(band, So[:, kmin:kmax + 1, :].sum(axis=1), rtol=1e-5)


@pytest.mark.parametrize("nperseg,hop", [(512, 64), (512, 128), (512, 256), (512, 448), (512, 512), (512, 300), (512, 10),
                                         (256, 32), (256, 64), (256, 128), (256, 224), (256, 256), (256, 100)])
@pytest.mark.parametrize("nframes_extra", [0, 1])
@pytest.mark.parametrize("detrend", ["constant", False])
def test_frame_duo_kernel(emu, nperseg, hop, nframes_extra, detrend):
    """nperseg 512 and 256 with any even hop run on the packed two-frames-per-lane kernels (sliding
    register window for hop = nperseg/8, /4, /2 and 7/8 nperseg, both frames loaded whole otherwise): odd and even frame counts, runs cut at odd lengths, crop / frame range / band power,
    float64 samples, against the oracle; the result of a frame must not depend on the chunking."""
    nfr = 6 + nframes_extra
    n = nperseg + hop * (nfr - 1) + 6          # even rows: the packed kernels need 8-byte aligned frames
    x = signal(3, n, nperseg + hop + nframes_extra, dc=-3.0 if detrend else 0.0)
    kw = dict(window="hann", nperseg=nperseg, noverlap=nperseg - hop, detrend=detrend)
    plan = plan_for(n, 5000.0, **kw)
    assert plan.nframes == nfr
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=5000.0, **kw)
    So = np.moveaxis(So, -1, -2)
    a = emu.stft_psd(x, plan, chunk=3, grid=1)           # odd run length: last duo of a run is half empty
    assert emu.last_family() == ("duo" if nperseg == 512 else "duo256")
    b = emu.stft_psd(x, plan, chunk=2, grid=2)
    assert_parity(a, So, what=f"duo {nperseg}/{hop}")
    assert np.array_equal(a, b)
    assert np.array_equal(emu.stft_psd(x.astype(np.float64), plan, chunk=4), a)
    part = emu.stft_psd(x, plan, kmin=3, kmax=40, frame0=1, nframes=nfr - 2, chunk=4)
    assert np.array_equal(part, a[:, 1:nfr - 1, 3:41])
    band = emu.band_power(x, plan, 3, 40, chunk=3)
    np.testing.assert_allclose(band, a[:, :, 3:41].astype(np.float64).sum(axis=-1), rtol=2e-6)
    edge = emu.band_power(x, plan, 0, nperseg // 2, chunk=2)
    np.testing.assert_allclose(edge, a.astype(np.float64).sum(axis=-1), rtol=2e-6)


@pytest.mark.parametrize("hop", [64, 128, 256])
@pytest.mark.parametrize("nframes,batch,grid,max_blocks", [(7, 5, 1, 64), (6, 9, 2, 4), (1, 3, 1, 2), (5, 2, 3, 1)])
def test_sum_fused_duo_kernel(emu, hop, nframes, batch, grid, max_blocks):
    """The sum-fused frame-duo kernel (per-sweep rows + cross-sweep sum in one pass): the rows
    are bit-identical to the per-sweep kernel's, the sum equals the float64 sum of the rows to
    fp32 rounding, whatever the split into sweep blocks (odd duo counts leave a lane group idle,
    odd frame counts a half-empty duo, the last block is ragged); float64 samples likewise."""
    n = 512 + hop * (nframes - 1) + 4          # even: rows stay 8-byte aligned (else the library takes the two-pass path)
    x = signal(batch, n, hop + nframes + batch, dc=-2.0)
    kw = dict(window="hann", nperseg=512, noverlap=512 - hop)
    plan = plan_for(n, 20000.0, **kw)
    assert plan.nframes == nframes
    rows = emu.stft_psd(x, plan, chunk=2)
    got, tot, blocks = emu.stft_psd_sum(x, plan, post_scale=0.5, grid=grid, max_blocks=max_blocks)
    assert 1 <= blocks <= min(max_blocks, batch)
    assert np.array_equal(got, rows)
    want = 0.5 * rows.astype(np.float64).sum(axis=0)
    np.testing.assert_allclose(tot, want, rtol=1e-6, atol=0)
    got64, tot64, _ = emu.stft_psd_sum(x.astype(np.float64), plan, post_scale=0.5, grid=grid, max_blocks=max_blocks)
    assert np.array_equal(got64, rows) and np.array_equal(tot64, tot)
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=20000.0, **kw)
    assert_parity(np.moveaxis(got, -1, -2), So, what=f"sum-fused duo 512/{hop}")


@pytest.mark.parametrize("hop,detrend", [(256, "constant"), (128, False), (896, "constant"), (1024, False), (36, "constant")])
@pytest.mark.parametrize("nframes,batch,grid,max_blocks", [(7, 5, 1, 64), (6, 9, 2, 4), (1, 3, 1, 2), (5, 2, 3, 1)])
def test_sum_fused_pair_kernel(emu, hop, detrend, nframes, batch, grid, max_blocks):
    """nperseg 1024: the SUM mode of the staged-sample pair kernel (a warp keeps one pair of frames and walks
    over a block of sweeps, one bulk copy per sweep): the rows are bit-identical to the per-sweep kernel's, the
    sum equals the float64 sum of the rows to fp32 rounding whatever the split into sweep blocks (odd frame
    counts leave a half-empty pair, the last block is ragged); float64 samples likewise."""
    n = 1024 + hop * (nframes - 1) + 4          # rows stay 16-byte aligned
    x = signal(batch, n, hop + nframes + batch, dc=-2.0 if detrend else 0.0)
    kw = dict(window=("tukey", 0.25), nperseg=1024, noverlap=1024 - hop, detrend=detrend)
    plan = plan_for(n, 20000.0, **kw)
    assert plan.nframes == nframes
    rows = emu.stft_psd(x, plan, chunk=2)
    assert emu.last_family() == "pair"
    got, tot, blocks = emu.stft_psd_sum(x, plan, post_scale=0.5, grid=grid, max_blocks=max_blocks)
    assert 1 <= blocks <= min(max_blocks, batch)
    assert np.array_equal(got, rows)
    want = 0.5 * rows.astype(np.float64).sum(axis=0)
    np.testing.assert_allclose(tot, want, rtol=1e-6, atol=0)
    if hop not in (128, 256, 512):      # float64 samples at these hops take the four-step frame-duo kernel: no fused form
        got64, tot64, _ = emu.stft_psd_sum(x.astype(np.float64), plan, post_scale=0.5, grid=grid, max_blocks=max_blocks)
        assert np.array_equal(got64, emu.stft_psd(x.astype(np.float64), plan, chunk=2))
        np.testing.assert_allclose(tot64, want, rtol=2e-6, atol=0)
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=20000.0, **kw)
    assert_parity(np.moveaxis(got, -1, -2), So, what=f"sum-fused pair 1024/{hop}")


@pytest.mark.parametrize("hop,detrend", [(64, "constant"), (32, False), (128, "constant"), (224, "constant"), (256, False), (100, "constant")])
@pytest.mark.parametrize("nframes,batch,grid,max_blocks", [(9, 5, 1, 64), (6, 9, 2, 4), (1, 3, 1, 2), (15, 2, 3, 1)])
def test_sum_fused_duo256_kernel(emu, hop, detrend, nframes, batch, grid, max_blocks):
    """nperseg 256: the SUM mode of the 256-point frame-duo kernel (a lane group keeps one duo and walks over a
    block of sweeps; four duos per warp share the block): rows bit-identical to the per-sweep kernel's -- also at
    the hops whose per-sweep kernel is the S = 14 / 16 variant -- and the sum equals the float64 sum of the rows
    to fp32 rounding whatever the split into sweep blocks; float64 samples likewise."""
    n = 256 + hop * (nframes - 1) + 4
    x = signal(batch, n, hop + nframes + batch, dc=-2.0 if detrend else 0.0)
    kw = dict(window=("tukey", 0.25), nperseg=256, noverlap=256 - hop, detrend=detrend)
    plan = plan_for(n, 20000.0, **kw)
    assert plan.nframes == nframes
    rows = emu.stft_psd(x, plan, chunk=2)
    assert emu.last_family() == "duo256"
    got, tot, blocks = emu.stft_psd_sum(x, plan, post_scale=0.5, grid=grid, max_blocks=max_blocks)
    assert 1 <= blocks <= min(max_blocks, batch)
    assert np.array_equal(got, rows)
    want = 0.5 * rows.astype(np.float64).sum(axis=0)
    np.testing.assert_allclose(tot, want, rtol=1e-6, atol=0)
    got64, tot64, _ = emu.stft_psd_sum(x.astype(np.float64), plan, post_scale=0.5, grid=grid, max_blocks=max_blocks)
    assert np.array_equal(got64, rows) and np.array_equal(tot64, tot)
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=20000.0, **kw)
    assert_parity(np.moveaxis(got, -1, -2), So, what=f"sum-fused duo256 256/{hop}")


@pytest.mark.parametrize("nperseg,hop,detrend", [(2048, 512, "constant"), (2048, 256, False), (2048, 1024, "constant"),
                                                 (4096, 1024, "constant"), (4096, 2048, False)])
@pytest.mark.parametrize("nframes,batch,grid,max_blocks", [(5, 5, 1, 64), (4, 7, 2, 3), (1, 3, 1, 2)])
def test_sum_fused_duo4_kernel(emu, nperseg, hop, detrend, nframes, batch, grid, max_blocks):
    """nperseg 2048 / 4096: the SUM mode of the four-step frame-duo kernel (a thread group keeps one duo and walks
    over a block of sweeps; every thread keeps the sums of its own final-stage tasks): rows bit-identical to the
    per-sweep kernel's, the sum equals the float64 sum of the rows to fp32 rounding whatever the split into sweep
    blocks (odd frame counts, ragged last block); float64 samples likewise."""
    n = nperseg + hop * (nframes - 1) + 4
    x = signal(batch, n, hop + nframes + batch, dc=-2.0 if detrend else 0.0)
    kw = dict(window=("tukey", 0.25), nperseg=nperseg, noverlap=nperseg - hop, detrend=detrend)
    plan = plan_for(n, 20000.0, **kw)
    assert plan.nframes == nframes
    rows = emu.stft_psd(x, plan, chunk=2)
    assert emu.last_family() == "duo4"
    got, tot, blocks = emu.stft_psd_sum(x, plan, post_scale=0.5, grid=grid, max_blocks=max_blocks)
    assert 1 <= blocks <= min(max_blocks, batch)
    assert np.array_equal(got, rows)
    want = 0.5 * rows.astype(np.float64).sum(axis=0)
    np.testing.assert_allclose(tot, want, rtol=1e-6, atol=0)
    got64, tot64, _ = emu.stft_psd_sum(x.astype(np.float64), plan, post_scale=0.5, grid=grid, max_blocks=max_blocks)
    assert np.array_equal(got64, rows) and np.array_equal(tot64, tot)
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=20000.0, **kw)
    assert_parity(np.moveaxis(got, -1, -2), So, what=f"sum-fused duo4 {nperseg}/{hop}")


def test_sum_fused_plan_fills_the_grid():
    """plan_stft_sum on BASELINE config 2 with a B200's resident groups: one round, 22 blocks of 46."""
    import ctypes
    import __graft_entry__ as ge
    lib = ctypes.CDLL(ge.build_emulator())
    fn = lib.emu_plan_sum
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    o = (ctypes.c_longlong * 3)()
    assert fn(1000, 309, 148 * 3 * 8, 64, o) == 22 and list(o) == [46, 156, 22 * 156]
    assert fn(3, 309, 148 * 3 * 8, 64, o) == 3 and o[0] == 1
    b = fn(1000, 100001, 148 * 3 * 8, 64, o)                # more duos than lane groups: few, long blocks
    assert 1 <= b <= 16 and (b - 1) * o[0] < 1000 <= b * o[0]
    assert fn(100000, 7, 148 * 3 * 8, 64, o) == 64


@pytest.mark.parametrize("nperseg,hop", [(2048, 512), (2048, 333), (4096, 1024), (4096, 3584)])
@pytest.mark.parametrize("detrend", ["constant", False])
def test_frame_duo_cta_kernel(emu, nperseg, hop, detrend):
    """nperseg 2048 / 4096 run on the packed two-frames-per-group CTA kernel (any hop, the odd
    one takes the scalar-load path): odd frame counts, runs cut at odd lengths, crop / frame
    range / band power, float64 samples; a frame's result must not depend on the chunking."""
    nfr = 5
    n = nperseg + hop * (nfr - 1) + 3
    x = signal(2, n, nperseg + hop + 1, dc=-3.0 if detrend else 0.0)
    kw = dict(window=("tukey", .25), nperseg=nperseg, noverlap=nperseg - hop, detrend=detrend)
    plan = plan_for(n, 48000.0, **kw)
    assert plan.nframes == nfr
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=48000.0, **kw)
    So = np.moveaxis(So, -1, -2)
    a = emu.stft_psd(x, plan, chunk=3, grid=1)
    assert emu.last_family() == "duo_cta"       # odd rows / hops outside the four-step kernel's set
    b = emu.stft_psd(x, plan, chunk=2, grid=2)
    assert_parity(a, So, what=f"duo cta {nperseg}/{hop}")
    assert np.array_equal(a, b)
    assert np.array_equal(emu.stft_psd(x.astype(np.float64), plan, chunk=4), a)
    part = emu.stft_psd(x, plan, kmin=3, kmax=900, frame0=1, nframes=nfr - 2, chunk=4)
    assert np.array_equal(part, a[:, 1:nfr - 1, 3:901])
    band = emu.band_power(x, plan, 0, nperseg // 2, chunk=3)
    np.testing.assert_allclose(band, a.astype(np.float64).sum(axis=-1), rtol=2e-6)


@pytest.mark.parametrize("nperseg,hop", [(1024, 256), (1024, 128), (1024, 512), (1024, 896), (1024, 1024),
                                         (2048, 512), (2048, 1024), (4096, 1024), (4096, 512),
                                         (8192, 2048), (16384, 4096)])
@pytest.mark.parametrize("detrend", ["constant", False])
@pytest.mark.parametrize("pair", [True, False])
def test_four_step_duo_kernel(emu, nperseg, hop, detrend, pair, monkeypatch):
    """nperseg 1024 / 2048 / 4096 with hop = S * nperseg/16 run on the four-step duo kernel
    (256-point sub-transforms per half-warp + fused radix-R final stage): odd frame counts, runs
    cut at odd lengths, crop / frame range / band power, float64 samples; chunking-invariant.
    On 16-byte aligned rows these shapes go to the staged-sample kernels (b2s_pair_kernel.cuh for 1024,
    b2s_pairq_kernel.cuh above, the latter opt-in with B2S_PAIRQ=1) unless B2S_NO_PAIR=1; all are held to the
    same checks."""
    if not pair:
        monkeypatch.setenv("B2S_NO_PAIR", "1")
    else:
        monkeypatch.setenv("B2S_PAIRQ", "1")            # nperseg >= 2048: the staged-sample kernel is opt-in
    nfr = 5
    n = nperseg + hop * (nfr - 1) + 4          # even rows: the packed kernels need 8-byte aligned frames
    x = signal(2, n, nperseg + hop + 1, dc=-3.0 if detrend else 0.0)
    kw = dict(window=("tukey", .25), nperseg=nperseg, noverlap=nperseg - hop, detrend=detrend)
    plan = plan_for(n, 48000.0, **kw)
    assert plan.nframes == nfr
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=48000.0, **kw)
    So = np.moveaxis(So, -1, -2)
    a = emu.stft_psd(x, plan, chunk=3, grid=1)
    assert emu.last_family() == (("pair" if nperseg == 1024 else "pairq") if pair else
                                 ("duo4" if nperseg <= 4096 else "big"))
    b = emu.stft_psd(x, plan, chunk=2, grid=2)
    assert_parity(a, So, what=f"{emu.last_family()} {nperseg}/{hop}")
    assert np.array_equal(a, b)
    if pair and nperseg == 1024 and hop in (128, 256, 512):
        # float64 samples at the overlapping hops take the four-step frame-duo kernel (every frame of the staged
        # kernel would re-convert its doubles): the same bits as that kernel gives float samples
        a64 = emu.stft_psd(x.astype(np.float64), plan, chunk=4)
        assert emu.last_family() == "duo4"
        monkeypatch.setenv("B2S_NO_PAIR", "1")
        assert np.array_equal(a64, emu.stft_psd(x, plan, chunk=3))
        monkeypatch.delenv("B2S_NO_PAIR")
        assert_parity(a64, So, what=f"duo4 float64 {nperseg}/{hop}")
    elif not (pair and nperseg == 16384):    # (a float64 ring of 16384 samples does not fit: float64 takes the round-1 kernel)
        assert np.array_equal(emu.stft_psd(x.astype(np.float64), plan, chunk=4), a)
    part = emu.stft_psd(x, plan, kmin=3, kmax=500, frame0=1, nframes=nfr - 2, chunk=4)
    assert np.array_equal(part, a[:, 1:nfr - 1, 3:501])
    band = emu.band_power(x, plan, 0, nperseg // 2, chunk=3)
    np.testing.assert_allclose(band, a.astype(np.float64).sum(axis=-1), rtol=2e-6)


@pytest.mark.parametrize("nt", [256, 512])
@pytest.mark.parametrize("hop,detrend", [(256, "constant"), (896, False), (36, "constant")])
def test_pair_kernel_cta_shapes(emu, nt, hop, detrend, monkeypatch):
    """The staged-sample 1024 kernel's CTA shape is a launch parameter: the 384-thread shape and the wide one
    (which transposes real and imaginary parts one after the other through a half-size buffer) give the same
    bits as the 128-thread default."""
    n = (1024 + hop * 10 + 11) // 4 * 4
    x = signal(3, n, hop + 3, dc=5.0 if detrend else 0.0)
    kw = dict(window="hann", nperseg=1024, noverlap=1024 - hop, detrend=detrend)
    plan = plan_for(n, 20000.0, **kw)
    a = emu.stft_psd(x, plan, grid=2)
    assert emu.last_family() == "pair"
    monkeypatch.setenv("B2S_PAIR_NT", str(nt))
    b = emu.stft_psd(x, plan, grid=1)
    assert np.array_equal(a, b)
    if hop not in (128, 256, 512):      # (float64 samples at these hops take the four-step frame-duo kernel)
        assert np.array_equal(emu.stft_psd(x.astype(np.float64), plan, grid=1), a)


@pytest.mark.parametrize("nperseg,hop", [(1000, 875), (288, 72), (260, 65), (2000, 500), (8000, 2000), (315, 100), (1001, 300),
                                         (4800, 1200), (1100, 275), (8190, 4000), (16380, 16380)])
@pytest.mark.parametrize("detrend", ["constant", False])
def test_mixed_radix_kernel(emu, nperseg, hop, detrend, monkeypatch):
    """Lengths that are not powers of two but factor into 2, 3, 5, 7, 11, 13 (the GUI accepts any integer
    32..8192, GUI.py:87-89) run the mixed-radix Stockham kernel (b2s_mixed_kernel.cuh): even lengths through
    the real-FFT trick, odd ones as a complex transform; crop / frame range / dB / band power / float64
    samples; the direct-DFT kernel (B2S_NO_MIXED=1) as a second opinion."""
    nfr = 4
    n = nperseg + hop * (nfr - 1) + 3
    x = signal(2, n, nperseg + hop, dc=-3.0 if detrend else 0.0)
    kw = dict(window=("tukey", .25), nperseg=nperseg, noverlap=nperseg - hop, detrend=detrend)
    plan = plan_for(n, 48000.0, **kw)
    assert plan.nframes == nfr
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=48000.0, **kw)
    So = np.moveaxis(So, -1, -2)
    a = emu.stft_psd(x, plan, grid=3)
    assert emu.last_family() == "mixed"
    assert_parity(a, So, what=f"mixed {nperseg}/{hop}")
    assert np.array_equal(emu.stft_psd(x.astype(np.float64), plan), a)
    K = nperseg // 2 + 1
    part = emu.stft_psd(x, plan, kmin=2, kmax=K - 3, frame0=1, nframes=nfr - 2)
    assert np.array_equal(part, a[:, 1:nfr - 1, 2:K - 2])
    band = emu.band_power(x, plan, 2, K - 3)
    np.testing.assert_allclose(band, a[:, :, 2:K - 2].astype(np.float64).sum(axis=-1), rtol=3e-6)
    floor = float(1e-6 * So.max())
    db = emu.stft_psd(x, plan, out_mode=1, db_floor=floor)
    big = So >= floor
    assert np.max(np.abs(db - 10 * np.log10(np.maximum(So, floor)))[big]) <= 1e-3
    monkeypatch.setenv("B2S_NO_MIXED", "1")
    if nperseg <= 2000:                      # (the O(N^2) kernel is slow in the emulator)
        d = emu.stft_psd(x, plan, grid=3)
        assert emu.last_family() == "dft"
        assert_parity(d, So, what=f"dft {nperseg}/{hop}")


def test_mixed_radix_support_table():
    """b2s_nperseg_support: 1 radix-16 kernels, 3 mixed radix, 2 direct DFT, 0 unsupported."""
    from spectrogram_generator_b200 import _lib
    lib = _lib.load()
    for n, want in [(1000, 3), (2000, 3), (8000, 3), (96, 2), (288, 3), (4800, 3), (8190, 3), (1001, 3), (33, 2), (8191, 2),
                    (34, 2), (31, 2), (62, 2), (1024, 1), (17 * 64, 2), (13 * 64, 3), (16383, 2), (16380, 3), (255, 2), (260, 3)]:
        assert lib.b2s_nperseg_support(n) == want, n


def test_sum_fused_plan_properties():
    """plan_stft_sum over random shapes: the blocks cover every sweep exactly once (ragged last
    block only), never exceed the cap, and the unit count matches blocks x even duo slots."""
    import ctypes
    import __graft_entry__ as ge
    lib = ctypes.CDLL(ge.build_emulator())
    fn = lib.emu_plan_sum
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    o = (ctypes.c_longlong * 3)()
    rng = np.random.default_rng(5)
    for _ in range(300):
        batch = int(rng.integers(1, 5000))
        nframes = int(rng.integers(1, 2000))
        resident = int(rng.integers(1, 5000))
        cap = int(rng.integers(1, 65))
        blocks = fn(batch, nframes, resident, cap, o)
        rows, ups, units = o[0], o[1], o[2]
        assert 1 <= blocks <= min(cap, batch)
        assert (blocks - 1) * rows < batch <= blocks * rows
        nduos = (nframes + 1) // 2
        assert ups == (nduos + 1) // 2 * 2 and units == blocks * ups

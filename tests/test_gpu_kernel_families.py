"""Every kernel family through the device API (Engine.stft_psd -> C ABI), seeded random shapes:
parity with the oracle, bit-identity of frame-range sub-calls / strided batches / float64
samples with the full float32 call, bin crop, dB, band power.  The families and the shapes
that select them are listed in csrc/b2s_dispatch.hpp."""
import numpy as np
import pytest
import torch

import spectrogram_generator_b200 as sg
from oracle import stft_oracle
from util import assert_parity

pytestmark = pytest.mark.gpu

# (nperseg, hop) -> family that runs it
CASES = [
    (512, 128), (512, 64), (512, 256), (512, 512),        # frame-duo kernel ((512, 448) below: even rows permitting)
    (256, 64), (256, 32), (256, 128), (256, 224), (256, 256), (256, 100),   # frame-duo kernel, 8 lanes (even rows permitting)
    (512, 300), (512, 10),                                # frame-duo kernel, any even hop
    (1024, 256), (1024, 128), (1024, 512),                # four-step duo, R = 2
    (1024, 1024),                                         # four-step duo without overlap (1024/896 below: odd rows permitting)
    (2048, 512), (2048, 1024), (4096, 1024), (4096, 512), # four-step duo, R = 4 / 8
    (1024, 896), (2048, 333), (4096, 3584),               # duo CTA kernel (reference default overlap, odd hop)
    (512, 448), (256, 37), (128, 32), (64, 16),           # frame-duo (512/448, even rows) / warp kernel
    (8192, 2048), (16384, 4096),                          # three-pass CTA kernel
    (1000, 875), (600, 150),                              # direct DFT
]


def _signal(rng, B, n, dc):
    t = np.arange(n)
    x = 0.3 * rng.standard_normal((B, n)) + np.sin(2 * np.pi * 0.0371 * t) + 0.5 * np.sin(2 * np.pi * 0.21 * t + 1.0) + dc
    return x.astype(np.float32)


@pytest.mark.parametrize("nperseg,hop", CASES)
def test_family(nperseg, hop):
    rng = np.random.default_rng(nperseg * 31 + hop)
    B = int(rng.integers(1, 4))
    nfr = int(rng.integers(5, 12))
    n = nperseg + hop * (nfr - 1) + int(rng.integers(0, hop))
    detrend = "constant" if rng.random() < 0.7 else False
    window = ("tukey", .25) if rng.random() < 0.5 else "hann"
    x = _signal(rng, B, n, dc=float(rng.choice([0.0, -3.0, 40.0])) if detrend else 0.0)
    kw = dict(window=window, nperseg=nperseg, noverlap=nperseg - hop, detrend=detrend)
    plan = sg.triage(n, 5000.0, window, nperseg, nperseg - hop, None, detrend, True, "density", "psd")
    assert plan.nframes == nfr
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=5000.0, **kw)
    So = np.moveaxis(So, -1, -2)
    eng = sg.engine()
    xd = torch.from_numpy(x).cuda()
    full = eng.stft_psd(xd, plan).cpu().numpy()
    assert_parity(full, So, what=f"{nperseg}/{hop} {kw}")
    # float64 samples holding the same values: same arithmetic, same bits -- except nperseg 1024 at the overlapping
    # hops, where float64 rows that could take the staged kernel run the four-step frame-duo kernel instead
    full64 = eng.stft_psd(xd.double(), plan).cpu().numpy()
    if nperseg == 1024 and hop in (128, 256, 512):
        assert_parity(full64, So, what=f"{nperseg}/{hop} float64 {kw}")
    else:
        assert np.array_equal(full64, full)
    # a strided batch (rows of a wider, 16-byte aligned buffer)
    wide = torch.zeros((B, n + 8), dtype=torch.float32, device="cuda")
    wide[:, :n] = xd
    assert np.array_equal(eng.stft_psd(wide[:, :n], plan).cpu().numpy(), full)
    # frame range + bin crop: bit-identical to the slice of the full result
    K = nperseg // 2 + 1
    k0, k1 = int(rng.integers(0, K // 3)), int(rng.integers(2 * K // 3, K))
    part = eng.stft_psd(xd, plan, kmin=k0, kmax=k1, frame0=1, nframes=nfr - 2).cpu().numpy()
    assert np.array_equal(part, full[:, 1:nfr - 1, k0:k1 + 1])
    # dB with a floor
    floor = float(1e-6 * So.max())
    db = eng.stft_psd(xd, plan, out_mode=1, db_floor=floor).cpu().numpy()
    ref_db = 10.0 * np.log10(np.maximum(So, floor))
    big = So >= floor
    assert np.max(np.abs(db[big] - ref_db[big])) <= 1e-3
    # fused band power == sum of the cropped bins
    band = eng.band_power(xd, plan, k0, k1).cpu().numpy()
    np.testing.assert_allclose(band, full[:, :, k0:k1 + 1].astype(np.float64).sum(axis=-1), rtol=3e-6)


@pytest.mark.parametrize("hop", [256, 128, 512, 896, 1024, 36, 300, 768])
@pytest.mark.parametrize("detrend", ["constant", False])
def test_staged_sample_pair_kernel(hop, detrend):
    """nperseg 1024 on 16-byte aligned rows: the staged-sample pair kernel (b2s_pair_kernel.cuh -- TMA
    bulk copies into a per-warp ring, any hop that is a multiple of 4).  Parity with the oracle, and
    bit-identity across run lengths (rings wrap at different places), the static schedule, float64
    samples, frame ranges / crops, with the four-step duo kernel as a second opinion."""
    from spectrogram_generator_b200 import _lib
    rng = np.random.default_rng(hop)
    B, nfr = 3, 23
    n = (1024 + hop * (nfr - 1) + 5 + 3) // 4 * 4
    x = _signal(rng, B, n, dc=-70.0 if detrend else 0.0)
    kw = dict(window=("tukey", .25), nperseg=1024, noverlap=1024 - hop, detrend=detrend)
    plan = sg.triage(n, 20000.0, kw["window"], 1024, 1024 - hop, None, detrend, True, "density", "psd")
    assert plan.nframes == nfr
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=20000.0, **kw)
    So = np.moveaxis(So, -1, -2)
    eng = sg.engine()
    xd = torch.from_numpy(x).cuda()
    full = eng.stft_psd(xd, plan)
    assert_parity(full.cpu().numpy(), So, what=f"pair 1024/{hop}")
    try:
        for units in (1, 3, 50):                   # runs of 23, 8 and 4 frames per warp
            _lib.set_option("pair_units", units)
            assert torch.equal(eng.stft_psd(xd, plan), full)
        _lib.set_option("static_units", 1)
        assert torch.equal(eng.stft_psd(xd, plan), full)
    finally:
        _lib.set_option("pair_units", 0)
        _lib.set_option("static_units", 0)
    full64 = eng.stft_psd(xd.double(), plan)
    k64 = _lib.last_kernel()
    part = eng.stft_psd(xd, plan, kmin=7, kmax=400, frame0=3, nframes=nfr - 5)
    assert torch.equal(part, full[:, 3:nfr - 2, 7:401])
    band = eng.band_power(xd, plan, 7, 400).cpu().numpy()
    np.testing.assert_allclose(band, full[:, :, 7:401].double().sum(dim=-1).cpu().numpy(), rtol=3e-6)
    _lib.set_option("no_pair", 1)
    try:
        other = eng.stft_psd(xd, plan)
    finally:
        _lib.set_option("no_pair", 0)
    assert not torch.equal(other, full)            # a different kernel did run
    assert_parity(other.cpu().numpy(), So, what=f"duo 1024/{hop}")
    if hop in (128, 256, 512):
        # float64 samples at the overlapping hops take the four-step frame-duo kernel (the staged kernel would
        # re-convert every frame's doubles: 1.6 x slower there, profiles/r2_f64_1024.md): that kernel's bits
        assert k64.startswith("stft_psd_duo4_kernel") and torch.equal(full64, other)
    else:
        assert k64.startswith("stft_psd_pair_kernel") and torch.equal(full64, full)


@pytest.mark.parametrize("nperseg,hop", [(2048, 512), (2048, 1792), (2048, 40), (4096, 1024), (4096, 4096),
                                         (8192, 2048), (8192, 1056), (16384, 4096), (16384, 14336)])
@pytest.mark.parametrize("detrend", ["constant", False])
def test_staged_sample_kernel_2048_to_16384(nperseg, hop, detrend):
    """b2s_pairq_kernel.cuh (TMA ring per frame group, Q = nperseg/256 real sub-sequences in packed pairs,
    fused untangle + radix-Q final stage): parity, bit-identity across run lengths / schedules / float64
    samples / frame ranges and crops, with the round-1 kernels as a second opinion."""
    from spectrogram_generator_b200 import _lib
    rng = np.random.default_rng(nperseg + hop)
    B, nfr = 2, 11
    n = (nperseg + hop * (nfr - 1) + 5 + 3) // 4 * 4
    x = _signal(rng, B, n, dc=-70.0 if detrend else 0.0)
    kw = dict(window=("tukey", .25), nperseg=nperseg, noverlap=nperseg - hop, detrend=detrend)
    plan = sg.triage(n, 20000.0, kw["window"], nperseg, nperseg - hop, None, detrend, True, "density", "psd")
    assert plan.nframes == nfr
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=20000.0, **kw)
    So = np.moveaxis(So, -1, -2)
    eng = sg.engine()
    xd = torch.from_numpy(x).cuda()
    other = eng.stft_psd(xd, plan)                  # the default: a round-1 kernel
    assert "pairq" not in _lib.last_kernel()
    # (16384-point frames on a -70 baseline: the worst of 1.8e5 bins may sit just over 1e-4, as for SciPy-float32)
    tail = 2e-5 if nperseg >= 8192 else 0.0
    assert_parity(other.cpu().numpy(), So, what=f"round-1 kernel {nperseg}/{hop}", tail=tail)
    _lib.set_option("no_pairq", 0)                  # opt in (measured slower than the round-1 kernels: off by default)
    try:
        _pairq_checks(eng, xd, plan, So, nperseg, hop, nfr, tail)
    finally:
        _lib.set_option("no_pairq", 1)


def _pairq_checks(eng, xd, plan, So, nperseg, hop, nfr, tail):
    from spectrogram_generator_b200 import _lib
    full = eng.stft_psd(xd, plan)
    assert "pairq" in _lib.last_kernel()
    assert_parity(full.cpu().numpy(), So, what=f"pairq {nperseg}/{hop}", tail=tail)
    try:
        for units in (1, 3, 50):
            _lib.set_option("pair_units", units)
            assert torch.equal(eng.stft_psd(xd, plan), full)
        _lib.set_option("static_units", 1)
        assert torch.equal(eng.stft_psd(xd, plan), full)
    finally:
        _lib.set_option("pair_units", 0)
        _lib.set_option("static_units", 0)
    if nperseg <= 8192:                      # (a float64 ring of 16384 samples does not fit: float64 takes the round-1 kernel there)
        assert torch.equal(eng.stft_psd(xd.double(), plan), full)
    K = nperseg // 2
    part = eng.stft_psd(xd, plan, kmin=7, kmax=K - 9, frame0=3, nframes=nfr - 5)
    assert torch.equal(part, full[:, 3:nfr - 2, 7:K - 8])
    band = eng.band_power(xd, plan, 7, K - 9).cpu().numpy()
    np.testing.assert_allclose(band, full[:, :, 7:K - 8].double().sum(dim=-1).cpu().numpy(), rtol=3e-6)
    db = eng.stft_psd(xd, plan, out_mode=1, db_floor=float(1e-6 * So.max())).cpu().numpy()
    big = So >= 1e-6 * So.max()
    assert np.max(np.abs(db - 10 * np.log10(np.maximum(So, 1e-6 * So.max())))[big]) <= 1e-3


def test_float64_samples_on_a_large_dc_level_keep_their_signal():
    """float64 input (neo / NIX sweeps, SweepManager.py:135-136): the pair kernel subtracts the frame's
    pivot in double before the cast, so a 1e-3 signal on a DC level of 1e4 -- which float32 samples
    cannot even represent -- comes out within the bar."""
    rng = np.random.default_rng(3)
    n = 1024 + 896 * 20
    t = np.arange(n)
    x = 1e4 + 1e-3 * (np.sin(2 * np.pi * 0.05 * t) + 0.1 * rng.standard_normal(n))
    f, tt, So = stft_oracle.spectrogram(x, fs=1000.0, nperseg=1024)
    plan = sg.triage(n, 1000.0, ("tukey", .25), 1024, None, None, "constant", True, "density", "psd")
    S = sg.engine().stft_psd(torch.from_numpy(x.reshape(1, -1)).cuda(), plan)[0].cpu().numpy()
    assert_parity(S.T, So, what="float64 samples, DC 1e4")


@pytest.mark.parametrize("nperseg,hop", [(512, 128), (256, 64), (1024, 256), (2048, 512), (1024, 896)])
def test_large_batches_take_the_dynamic_schedule_and_stay_deterministic(nperseg, hop):
    """Enough work for the atomic work counter (b2s_api.cu: launch_any_impl): two launches give the
    same bits, and every row equals the one-signal call (which takes the static schedule)."""
    rng = np.random.default_rng(nperseg + hop)
    B, n = 2000, nperseg + hop * 40
    x = torch.from_numpy(_signal(rng, B, n, dc=1.0)).cuda()
    plan = sg.triage(n, 1.0, "hann", nperseg, nperseg - hop, None, "constant", True, "density", "psd")
    eng = sg.engine()
    a = eng.stft_psd(x, plan)
    b = eng.stft_psd(x, plan)
    assert torch.equal(a, b)
    for row in (0, 777, B - 1):
        assert torch.equal(eng.stft_psd(x[row:row + 1], plan)[0], a[row])


@pytest.mark.parametrize("nperseg,hop,B,nfr,odd", [
    (512, 128, 1000, 309, 0),       # BASELINE config 2: the sum-fused frame-duo kernel, one round of 22 blocks
    (512, 128, 37, 10, 0), (512, 64, 5, 7, 0), (512, 256, 130, 6, 0), (512, 128, 2, 1, 0),
    (512, 128, 9, 8, 1),            # odd row length: rows only 4-byte aligned -> two-pass path inside the library
    (512, 128, 1, 12, 0),           # one sweep
    (1024, 256, 40, 9, 0), (600, 150, 7, 5, 0), (512, 448, 20, 6, 0),   # no fused kernel for these shapes
    # nperseg 256: the SUM mode of the 256-point frame-duo kernel (S = 2 / 4 / 8, and S = 16 for every other even hop)
    (256, 64, 300, 11, 3), (256, 64, 1000, 622, 3), (256, 32, 77, 40, 3), (256, 128, 1000, 9, 3), (256, 224, 50, 9, 3),
    (256, 100, 20, 7, 3), (256, 256, 3, 1, 3),
    # nperseg 2048 at hop 256 / 512 / 1024: the SUM mode of the four-step frame-duo kernel
    (2048, 512, 500, 75, 4), (2048, 256, 33, 9, 4), (2048, 1024, 64, 6, 4), (2048, 512, 2, 1, 4),
    (4096, 1024, 30, 7, 0), (2048, 300, 9, 5, 0),       # no fused kernel (4096: measured slower; 2048 at another hop)
    # nperseg 1024, rows 16-byte aligned: the SUM mode of the staged-sample pair kernel, any staged hop
    (1024, 256, 1000, 153, 2),      # the north-star target shape (1000 x 40 000 @ 1024/256) with its mean
    (1024, 896, 33, 7, 2), (1024, 1024, 64, 5, 2), (1024, 36, 9, 20, 2), (1024, 256, 2, 1, 2), (1024, 128, 300, 31, 2),
])
def test_rows_and_cross_sweep_sum_in_one_call(nperseg, hop, B, nfr, odd):
    """Engine.stft_psd_sum (b2s_stft_psd_sum_f32/_f64): the rows are bit-identical to stft_psd's,
    the sum equals the float64 sum of the rows to fp32 rounding and is the same on every run;
    float64 samples and a caller-provided flat sum buffer likewise."""
    rng = np.random.default_rng(nperseg + hop + B)
    n = nperseg + hop * (nfr - 1) + {0: 2, 1: 3, 2: 4, 3: 2, 4: 2}[odd]
    x = _signal(rng, B, n, dc=-3.0)
    plan = sg.triage(n, 20000.0, "hann", nperseg, nperseg - hop, None, "constant", True, "density", "psd")
    assert plan.nframes == nfr
    eng = sg.engine()
    xd = torch.from_numpy(x).cuda()
    rows = eng.stft_psd(xd, plan)
    S, tot = eng.stft_psd_sum(xd, plan, post_scale=1.0 / B)
    from spectrogram_generator_b200 import _lib
    if odd == 2:
        assert _lib.last_kernel().startswith("stft_psd_pair_sum_kernel")
    if odd == 3:
        assert _lib.last_kernel().startswith("stft_psd_duo256_sum_kernel")
    if odd == 4:
        assert _lib.last_kernel().startswith("stft_psd_duo4_sum_kernel")
    assert torch.equal(S, rows)
    want = rows.double().sum(dim=0) / B
    torch.testing.assert_close(tot.double(), want, rtol=2e-6, atol=0)
    flat = torch.full((nfr * plan.nbins,), float("nan"), device="cuda")
    S2, tot2 = eng.stft_psd_sum(xd, plan, post_scale=1.0 / B, sum_out=flat)
    assert tot2 is flat and torch.equal(flat.view(nfr, -1), tot) and torch.equal(S2, rows)
    S3, tot3 = eng.stft_psd_sum(xd.double(), plan, post_scale=1.0 / B)
    if nperseg == 1024 and hop in (128, 256, 512) and odd == 2:
        # float64 samples at these hops run the four-step frame-duo kernel + two-pass sum: its rows, its order
        assert torch.equal(S3, eng.stft_psd(xd.double(), plan))
        torch.testing.assert_close(tot3.double(), S3.double().sum(dim=0) / B, rtol=2e-6, atol=0)
    else:
        assert torch.equal(S3, rows) and torch.equal(tot3, tot)


def test_fused_sum_matches_the_two_pass_sum_and_the_public_mean(monkeypatch):
    """The sum-fused kernel against the earlier form (stft_psd, then batch_sum) on a C2-shaped batch,
    and mean_spectrogram (which takes the fused call) against the oracle's float64 mean."""
    from spectrogram_generator_b200 import synth
    x, kw = synth.config2(batch=200)
    fs = kw.pop("fs")
    plan = sg.triage(x.shape[-1], fs, kw["window"], kw["nperseg"], kw["noverlap"], None, "constant", True,
                     "density", "psd")
    eng = sg.engine()
    xd = torch.from_numpy(x).cuda()
    S, tot = eng.stft_psd_sum(xd, plan)
    # running sums in shared memory instead of tensor memory: the same additions in the same order
    from spectrogram_generator_b200 import _lib
    _lib.set_option("sum_acc_smem", 1)
    try:
        S1, tot1 = eng.stft_psd_sum(xd, plan)
    finally:
        _lib.set_option("sum_acc_smem", 0)
    assert torch.equal(S1, S) and torch.equal(tot1, tot)
    _lib.set_option("no_fused_sum", 1)
    try:
        S0, tot0 = eng.stft_psd_sum(xd, plan)
    finally:
        _lib.set_option("no_fused_sum", 0)
    assert torch.equal(S, S0)
    torch.testing.assert_close(tot, tot0, rtol=2e-6, atol=0)
    assert torch.equal(tot0, eng.batch_sum(eng.stft_psd(xd, plan)))
    f, t, m = sg.mean_spectrogram(x, fs=fs, **kw)
    _, _, mo = stft_oracle.mean_spectrogram(x.astype(np.float64), fs=fs, **kw)
    assert_parity(m, mo, what="mean spectrogram through the sum-fused kernel")


def test_odd_row_lengths_through_the_host_api_take_the_packed_kernels():
    """A batch with an odd number of samples per row: the host API stages it into rows of even stride
    (same bits as handing the engine such rows); a device-resident batch with an odd row stride runs the
    scalar-load kernels; both match the oracle."""
    rng = np.random.default_rng(77)
    n = 512 + 128 * 9 + 1
    x = _signal(rng, 3, n, dc=-3.0)
    kw = dict(window="hann", nperseg=512, noverlap=384)
    plan = sg.triage(n, 20000.0, "hann", 512, 384, None, "constant", True, "density", "psd")
    _, _, So = stft_oracle.spectrogram(x.astype(np.float64), fs=20000.0, **kw)
    eng = sg.engine()
    xd = torch.from_numpy(x).cuda()
    assert xd.stride(0) % 2 == 1
    scalar = eng.stft_psd(xd, plan)
    wide = torch.zeros((3, n + 1), dtype=torch.float32, device="cuda")
    wide[:, :n] = xd
    packed = eng.stft_psd(wide[:, :n], plan)
    assert_parity(np.moveaxis(scalar.cpu().numpy(), -1, -2), So, what="odd rows, scalar loads")
    assert_parity(np.moveaxis(packed.cpu().numpy(), -1, -2), So, what="odd rows staged at even stride")
    f, t, S = sg.spectrogram(x, fs=20000.0, **kw)
    assert np.array_equal(np.moveaxis(S, -1, -2), packed.cpu().numpy())
    f, t, m, Sm = sg.mean_spectrogram(x, fs=20000.0, return_per_sweep=True, **kw)
    assert np.array_equal(np.moveaxis(Sm, -1, -2), packed.cpu().numpy())

"""world_size-2 gloo tests of the sharding / collective host logic.  The per-rank
compute is injected (the CPU emulator build of the kernels), so no GPU is needed;
on GPUs the same functions run the CUDA engine and NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _emu_compute():
    import util
    emu = util.Emulator()

    def compute(x2d, plan):
        return torch.from_numpy(emu.stft_psd(np.ascontiguousarray(x2d, dtype=np.float32), plan))
    return compute


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spectrogram_generator_b200 as sg
        from spectrogram_generator_b200 import distributed as D
        from oracle import stft_oracle
        compute = _emu_compute()
        # --- sweeps sharded over ranks + all-reduce for the mean (config 2 shape, small)
        rng = np.random.default_rng(0)
        x = (rng.standard_normal((7, 3000)) + np.sin(0.2 * np.arange(3000))).astype(np.float32)
        kw = dict(fs=20000.0, window="hann", nperseg=512, noverlap=384)
        lo, hi = D.shard_rows(7, world, rank)
        f, t, mean, S_local = D.mean_spectrogram_sharded(x[lo:hi], 7, compute=compute, return_local=True, **kw)
        _, _, mo = stft_oracle.mean_spectrogram(x.astype(np.float64), **kw)
        assert np.max(np.abs(mean.numpy() - mo)) <= 1e-6 * mo.max()
        counts = [D.shard_rows(7, world, r)[1] - D.shard_rows(7, world, r)[0] for r in range(world)]
        allS = D.gather_slabs(S_local, counts, dst=0)
        if rank == 0:
            _, _, So = stft_oracle.spectrogram(x.astype(np.float64), **kw)
            assert allS.shape == (7, So.shape[2], So.shape[1])
            assert np.max(np.abs(allS.numpy() - np.moveaxis(So, -1, -2))) <= 1e-6 * So.max()
        else:
            assert allS is None
        # --- one long recording, frame ranges sharded (config 3 shape, small)
        y = (rng.standard_normal(50000) * 0.1 + np.sin(0.13 * np.arange(50000))).astype(np.float32)
        kw3 = dict(fs=48000.0, window="hann", nperseg=2048, noverlap=1536)
        plan = sg.triage(len(y), 48000.0, "hann", 2048, 1536, None, "constant", True, "density", "psd")
        f0, c = D.shard_frames(plan.nframes, world, rank)
        slo, shi = D.sample_span(f0, c, plan.hop, plan.nperseg)
        f, t_loc, S_loc, (g0, gc) = D.spectrogram_time_sharded(y[slo:shi], len(y), compute=compute, **kw3)
        fo, to, So = stft_oracle.spectrogram(y.astype(np.float64), **kw3)
        assert (g0, gc) == (f0, c) and np.array_equal(t_loc, to[f0:f0 + c]) and np.array_equal(f, fo)
        whole = compute(y.reshape(1, -1), plan)[0]
        assert torch.equal(S_loc, whole[f0:f0 + c]), "sharded frames must equal the unsharded run bit for bit"
        counts = [D.shard_frames(plan.nframes, world, r)[1] for r in range(world)]
        full = D.gather_slabs(S_loc, counts, dst=0)
        if rank == 0:
            assert torch.equal(full, whole)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_single_process_defaults():
    from spectrogram_generator_b200 import distributed as D
    assert D.world() == (1, 0)
    t = torch.arange(6.0).reshape(3, 2)
    assert D.gather_slabs(t, [3]) is t
    with pytest.raises(ValueError):
        D.spectrogram_time_sharded(np.zeros(10, np.float32), 5000, nperseg=256, compute=lambda a, b: None)

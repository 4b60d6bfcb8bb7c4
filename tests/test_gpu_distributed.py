"""NCCL path on >= 2 GPUs of one box (skipped on a single-GPU box): sweeps sharded with
the all-reduce for the mean, a long recording sharded by frame ranges, gather to rank 0."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import spectrogram_generator_b200 as sg
        from spectrogram_generator_b200 import distributed as D, synth
        from oracle import stft_oracle
        x, kw = synth.config2(batch=37)
        fs = kw.pop("fs")
        lo, hi = D.shard_rows(37, world, rank)
        f, t, mean, S_local = D.mean_spectrogram_sharded(x[lo:hi], 37, fs=fs, return_local=True, **kw)
        _, _, mo = stft_oracle.mean_spectrogram(x.astype(np.float64), fs=fs, **kw)
        assert np.max(np.abs(mean.cpu().numpy() - mo)) <= 1e-6 * mo.max()
        counts = [D.shard_rows(37, world, r)[1] - D.shard_rows(37, world, r)[0] for r in range(world)]
        allS = D.gather_slabs(S_local, counts, dst=0)
        if rank == 0:
            _, _, S1 = sg.spectrogram(x, fs=fs, **kw)
            assert np.array_equal(np.moveaxis(allS.cpu().numpy(), -1, -2), S1), "sharded != single-GPU bits"
        y, kw3 = synth.config3(n=48000 * 20)
        fs3 = kw3.pop("fs")
        plan = sg.triage(len(y), fs3, kw3["window"], kw3["nperseg"], kw3["noverlap"], None, "constant", True,
                         "density", "psd")
        f0, c = D.shard_frames(plan.nframes, world, rank)
        slo, shi = D.sample_span(f0, c, plan.hop, plan.nperseg)
        f, t_loc, S_loc, _ = D.spectrogram_time_sharded(y[slo:shi], len(y), fs=fs3, **kw3)
        counts = [D.shard_frames(plan.nframes, world, r)[1] for r in range(world)]
        full = D.gather_slabs(S_loc, counts, dst=0)
        if rank == 0:
            fw, tw, Sw = sg.spectrogram(y, fs=fs3, **kw3)
            assert np.array_equal(full.cpu().numpy().T, Sw), "time-sharded != unsharded bits"
        assert np.array_equal(t_loc, sg.windows.time_axis(len(y), 2048, 1536, fs3)[f0:f0 + c])
        # the mean all-reduce as one kernel over NVLink peer memory == NCCL's, on every rank, repeatedly
        dev = torch.device("cuda", rank)
        elems = 309 * 257
        red = D.PeerMeanReducer(elems, dev)
        g = torch.Generator(device="cpu").manual_seed(100 + rank)
        for it in range(5):
            part = torch.rand(elems, generator=g).to(dev) * (it + 1)
            red.partial().copy_(part)
            got = red.reduce(0.25)
            ref = part.clone()
            dist.all_reduce(ref)
            torch.testing.assert_close(got, ref * 0.25, rtol=1e-6, atol=0)
            both = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(both, got)
            assert all(torch.equal(both[0], b) for b in both), "peer all-reduce differs between ranks"
        # overlap mode: the reduces run on a side stream (three buffers), results as NCCL's
        red3 = D.PeerMeanReducer(elems, dev, overlap=True)
        outs, refs = [], []
        for it in range(7):
            part = torch.rand(elems, generator=g).to(dev) * (it + 1)
            red3.partial().copy_(part)
            outs.append(red3.reduce(0.5))
            ref = part.clone()
            dist.all_reduce(ref)
            refs.append(ref * 0.5)
        red3.wait()
        torch.cuda.synchronize()
        for o, r in zip(outs, refs):
            torch.testing.assert_close(o, r, rtol=1e-6, atol=0)
        # ... and the same reducer behind the sharded mean
        red2 = D.PeerMeanReducer(309 * 257, dev)
        _, _, mean_p = D.mean_spectrogram_sharded(x[lo:hi], 37, fs=fs, reducer=red2, **kw)
        assert np.max(np.abs(mean_p.cpu().numpy() - mo)) <= 1e-6 * mo.max()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_ranks_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on one box (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]

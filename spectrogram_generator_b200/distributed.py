"""Sharding the spectrogram path across GPUs (one process per GPU).

The forward STFT has no data dependency between frames, so the path shards
without any exchange:

  * sweeps / channels  -> contiguous blocks of rows per rank (configs 2, 4);
  * one long recording -> contiguous *frame* ranges per rank; rank r reads the
    samples ``[f0*hop, (f0+c-1)*hop + nperseg)`` -- the ``nperseg-hop`` halo it
    shares with its neighbour is read-only input (config 3).

Two collectives exist, both after the kernels: the all-reduce of the per-rank
partial *sum* spectrogram for the cross-sweep mean, and the gather of per-rank
slabs to the exporting rank.  They go through ``torch.distributed`` (NCCL on
GPUs; gloo in the CPU tests, where the per-rank compute is injected).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

from .spectrogram import Plan, _to_device, engine, split_frames, triage
from .windows import rfftfreq, time_axis


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_rows(n_rows: int, world_size: int, rank: int):
    """Contiguous block ``[lo, hi)`` of rows (sweeps / channels) owned by ``rank``."""
    base, rem = divmod(n_rows, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_frames(nframes: int, world_size: int, rank: int):
    """Contiguous frame range ``(f0, count)`` owned by ``rank``."""
    return split_frames(nframes, world_size)[rank]


def sample_span(f0: int, count: int, hop: int, nperseg: int):
    """Samples ``[lo, hi)`` a rank needs for frames ``[f0, f0+count)`` (halo included)."""
    if count <= 0:
        return f0 * hop, f0 * hop
    return f0 * hop, (f0 + count - 1) * hop + nperseg


Compute = Callable[[np.ndarray, Plan], torch.Tensor]


def _cuda_compute(x2d: np.ndarray, plan: Plan) -> torch.Tensor:
    eng = engine()
    eng.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    return eng.stft_psd(_to_device(x2d, dev), plan)


def mean_spectrogram_sharded(x_local, total_sweeps: int, fs=1.0, window=("tukey", .25), nperseg=None,
                             noverlap=None, detrend="constant", scaling="density", *, group=None,
                             compute: Optional[Compute] = None, return_local=False, reducer=None):
    """Cross-sweep mean with sweeps sharded over ranks.

    ``x_local[B_r, N]`` holds this rank's sweeps (``shard_rows``).  Each rank computes
    its spectrograms and their sum in one pass on the device, one all-reduce (sum) of
    ``[F, K]`` fp32 follows, then the division by ``total_sweeps``.  The all-reduce is NCCL's, or --
    with ``reducer=PeerMeanReducer(F * K, device)``, created once and reused -- one
    kernel over NVLink peer memory.  Returns ``(f, t, Smean[K, F])`` as a torch
    tensor on the compute device (identical on every rank)."""
    x_local = np.asarray(x_local)
    plan = triage(x_local.shape[-1], fs, window, nperseg, noverlap, None, detrend, True, scaling, "psd")
    if x_local.shape[0] == 0:
        raise ValueError("every rank needs at least one sweep")
    ws, _ = world(group)
    if compute is None:
        # rows and this rank's partial sum in one pass (Engine.stft_psd_sum), straight into the
        # reducer's symmetric buffer when there is one
        eng = engine()
        eng.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        use_peer = reducer is not None and ws > 1
        S, part = eng.stft_psd_sum(_to_device(x_local.reshape(-1, plan.n), dev), plan,
                                   sum_out=reducer.partial() if use_peer else None)
        if use_peer:
            mean = reducer.reduce(1.0 / float(total_sweeps)).view(S.shape[1:])
            reducer.check()                      # joins the side stream; raises if a rank missed the time-out
        else:
            if ws > 1:
                dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
            mean = part * (1.0 / float(total_sweeps))
    else:                                        # injected compute (the gloo tests run it on the CPU)
        S = compute(x_local.reshape(-1, plan.n), plan)
        part = engine().batch_sum(S, 1.0) if S.is_cuda else S.sum(dim=0)
        if ws > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
        mean = part * (1.0 / float(total_sweeps))
    f = rfftfreq(plan.nperseg, fs)
    t = time_axis(plan.n, plan.nperseg, plan.noverlap, fs)
    out = (f, t, mean.transpose(-1, -2))
    return out + (S,) if return_local else out


def spectrogram_time_sharded(x_span, n_total: int, fs=1.0, window=("tukey", .25), nperseg=None,
                             noverlap=None, detrend="constant", scaling="density", *, group=None,
                             compute: Optional[Compute] = None):
    """One long recording, frame ranges sharded over ranks.

    ``x_span`` is this rank's sample slice ``x[lo:hi]`` with ``(lo, hi) =
    sample_span(*shard_frames(F, world, rank), hop, nperseg)``; ``n_total`` is the
    length of the whole recording (for the global time axis).  Returns
    ``(f, t_local, S_local[F_r, K], (f0, count))`` -- ``t_local`` is cut from the
    global axis so the sharded run equals the unsharded one bit for bit."""
    plan_g = triage(int(n_total), fs, window, nperseg, noverlap, None, detrend, True, scaling, "psd")
    ws, rk = world(group)
    f0, count = shard_frames(plan_g.nframes, ws, rk)
    lo, hi = sample_span(f0, count, plan_g.hop, plan_g.nperseg)
    x_span = np.asarray(x_span)
    if x_span.shape[-1] != hi - lo:
        raise ValueError(f"rank {rk} expects samples [{lo}, {hi}) of the recording")
    f = rfftfreq(plan_g.nperseg, fs)
    t = time_axis(plan_g.n, plan_g.nperseg, plan_g.noverlap, fs)[f0:f0 + count]
    compute = compute or _cuda_compute
    if count == 0:
        return f, t, None, (f0, count)
    sub = Plan(**{**plan_g.__dict__, "n": hi - lo, "nframes": count})
    S = compute(x_span.reshape(1, -1), sub)[0]
    return f, t, S, (f0, count)


def gather_slabs(local: torch.Tensor, counts, dst: int = 0, *, group=None):
    """Gather per-rank slabs ``[rows_r, ...]`` (rows_r = counts[r]) to rank ``dst``
    and concatenate along dim 0; other ranks return ``None``.  This is the "final
    gather to the exporting rank"."""
    ws, rk = world(group)
    if ws == 1:
        return local
    tail = tuple(local.shape[1:])
    # one batched group of point-to-point operations: NCCL runs the ws - 1 receives of the exporting rank
    # concurrently (issued one by one on the default group they are serialised: measured 389 GB/s into rank 0
    # at 8 GPUs)
    batched = hasattr(dist, "batch_isend_irecv") and dist.get_backend(group) == "nccl"
    if rk == dst:
        # receive straight into the slabs of the result (no concatenation pass afterwards)
        full = torch.empty((int(sum(int(c) for c in counts)),) + tail, dtype=local.dtype, device=local.device)
        offs = np.concatenate([[0], np.cumsum([int(c) for c in counts])]).astype(int)
        bufs = [full[offs[r]:offs[r + 1]] for r in range(ws)]
        bufs[rk].copy_(local)
        src = [r for r in range(ws) if r != dst and counts[r] > 0]
        if batched and src:
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.irecv, bufs[r], r, group) for r in src])
        else:
            reqs = [dist.irecv(bufs[r], src=r, group=group) for r in src]
        for q in reqs:
            q.wait()
        return full
    if counts[rk] > 0:
        if batched:
            for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst, group)]):
                q.wait()
        else:
            dist.send(local.contiguous(), dst=dst, group=group)
    return None


class PeerMeanReducer:
    """The mean-spectrogram all-reduce as ONE kernel over NVLink peer memory instead of an NCCL
    call: each rank's cross-sweep sum lands in a symmetric buffer (torch's symmetric-memory
    rendezvous maps it into every peer), and ``b2s_peer_allreduce_f32`` announces, waits for the
    peers' announcements and adds the partials in rank order straight from peer memory -- the
    result is bit-identical on every rank.  For the [F, K] partial of config 2 (318 KB) this
    replaces ~30 us of collective latency per step by one ~10 us launch.

    With ``overlap=True`` the reduce kernel runs on a high-priority side stream, so the all-reduce of
    one step overlaps the kernels of the next (leave it CTA slots with ``b2s_set_reserved_sms``).

    Usage (every rank, same order):  ``r = PeerMeanReducer(F * K, device)``; per step
    ``engine().stft_psd_sum(x, plan, sum_out=r.partial())`` (or ``batch_sum(S, 1.0, out=r.partial())``)
    then ``mean = r.reduce(1.0 / total_sweeps)``.
    """

    def __init__(self, elems: int, device, group=None, overlap: bool = False, coresident: bool = False,
                 nbuf: Optional[int] = None):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib
        self._ct = ctypes
        self._lib = _lib.load()
        self.elems = int(elems)
        self.device = torch.device(device)
        group = group if group is not None else dist.group.WORLD
        # overlap: the all-reduce of step i runs on a side stream beside the kernels of step i + 1.
        # A partial may then be overwritten only two reduces later, hence three buffers instead of two:
        # slot e % 3 is rewritten for epoch e + 3 after this rank's reduce(e + 1) has completed, which
        # it only does once every peer has announced e + 1, i.e. has finished reading epoch e.
        self.overlap = bool(overlap)
        # coresident: launch the reduce as one-warp CTAs (B2S_PEER_CORESIDENT) instead of a few 256-thread CTAs in
        # reserved slots.  Measured on 2 x B200 beside the 168-register STFT CTAs: registers are handed out to
        # four warps at a time, so even a one-warp CTA does not fit the 1024 registers three STFT CTAs leave and
        # displaces one of them on every SM (0.176 -> 0.194 ms per step).  Off by default.
        self.coresident = bool(coresident)
        # nbuf > 3 (overlap mode): the launching stream may run up to nbuf - 2 steps ahead of the slowest rank before it
        # has to wait for a slot, so the step-to-step jitter of the ranks is not paid at every step (each step would
        # otherwise cost the slowest rank's time of that step); the reduces themselves stay in lockstep on the side stream.
        self.nbuf = (max(3, int(nbuf)) if nbuf else 3) if self.overlap else 2
        self.stride = (self.elems + 3) // 4 * 4           # every slot 16-byte aligned: 128-bit peer loads
        self.buf = symm_mem.empty(self.nbuf * self.stride, dtype=torch.float32, device=self.device)
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.world, self.rank = int(self.handle.world_size), int(self.handle.rank)
        if self.world > 16:
            raise _lib.B2SError("PeerMeanReducer supports up to 16 ranks")
        self._bufs = [int(p) for p in self.handle.buffer_ptrs]
        self._pads = (ctypes.c_ulonglong * self.world)(*[int(p) for p in self.handle.signal_pad_ptrs])
        self.epoch = 0
        self._done = {}                        # epoch -> event recorded after its reduce (overlap mode)
        if self.overlap:
            with torch.cuda.device(self.device):
                self.side = torch.cuda.Stream(priority=-1)       # its few CTAs go first when slots free up
                # events are reused round-robin (a wait captures the record that precedes it): the step
                # loop must stay cheap on the host, one step is ~0.17 ms of GPU time
                self._nev = self.nbuf + 2
                self._ev_ready = [torch.cuda.Event() for _ in range(self._nev)]
                self._ev_done = [torch.cuda.Event() for _ in range(self._nev)]
        self.handle.barrier()                  # pads and buffers exist on every rank before the first kernel

    def _slot(self, epoch: int) -> int:
        return epoch % self.nbuf

    def partial(self) -> torch.Tensor:
        """Where this rank writes its partial of the coming ``reduce`` call.  In overlap mode the
        current stream first waits until the slot's previous content has been read by every peer."""
        nxt = self.epoch + 1
        if self.overlap:
            # slot nxt % nbuf held epoch nxt - nbuf; every peer has read it once this rank's reduce(nxt - nbuf + 1) is
            # done.  Normally that was several steps ago: then nothing is put into the stream at all
            ev = self._done.pop(nxt - (self.nbuf - 1), None)
            if ev is not None and not ev.query():
                torch.cuda.current_stream(self.device).wait_event(ev)
        k = self._slot(nxt)
        return self.buf[k * self.stride:k * self.stride + self.elems]

    def reduce(self, post_scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Sum of all ranks' partials times ``post_scale`` -> ``out`` (allocated if None).  Must be
        called by every rank, in the same order.  Enqueued on the current stream -- in overlap mode on
        the reducer's side stream, after what the current stream has done so far; ``wait()`` (or the
        next-but-one ``partial()``) joins it."""
        self.epoch += 1
        k = self._slot(self.epoch)
        caller_out = out is not None       # the caller keeps it alive until wait() (overlap mode)
        if out is None:
            out = torch.empty(self.elems, dtype=torch.float32, device=self.device)
        bufs = (self._ct.c_ulonglong * self.world)(*[b + 4 * k * self.stride for b in self._bufs])
        from . import _lib
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            if self.overlap:
                ready = self._ev_ready[self.epoch % self._nev]
                ready.record(cur)
                self.side.wait_event(ready)
                stream = self.side
            else:
                stream = cur
            rc = self._lib.b2s_peer_allreduce_ex_f32(bufs, self._pads, self.world, self.rank, self.epoch, self.elems,
                                                     out.data_ptr(), float(post_scale), 1 if self.coresident else 0,
                                                     stream.cuda_stream)
            _lib.check(rc, "b2s_peer_allreduce_f32")
            if self.overlap:
                if not caller_out:
                    out.record_stream(self.side)
                done = self._ev_done[self.epoch % self._nev]
                done.record(self.side)
                self._done[self.epoch] = done
        return out

    def wait(self):
        """Overlap mode: make the current stream wait for every reduce issued so far."""
        if self.overlap:
            torch.cuda.current_stream(self.device).wait_stream(self.side)

    def check(self):
        """Synchronise and raise ``TimeoutError`` if a reduce gave up waiting for a late rank
        (``B2S_PEER_TIMEOUT_MS``, default two minutes of wall-clock time; its output was not
        written).  Every rank must enqueue ``reduce`` within that window of its peers."""
        from . import _lib
        self.wait()
        with torch.cuda.device(self.device):
            rc = self._lib.b2s_peer_allreduce_status(torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "b2s_peer_allreduce_status")

// Auxiliary (non-template) kernels: the deterministic cross-sweep sum and the display scaling.
// Included by b2s_api.cu and by the emulator harness only (one definition per library).
#pragma once

#include "b2s_fft.cuh"

namespace b2s {

// ---- deterministic sum over the batch axis (cross-sweep mean) ----
// in: [batch][elems] (row stride in_stride), out[slab][elems] partial sums over
// a slab of rows; a second launch with batch = n_slabs folds the partials.
B2S_GLOBAL void batch_sum_kernel(const float* __restrict__ in, long long in_stride, int batch,
                                 int rows_per_slab, long long elems, float* __restrict__ out,
                                 float post_scale) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // the last slabs are the rows the STFT kernel wrote last: take them first, while they are
    // still in L2
    const int slab = (int)(gridDim.y - 1 - blockIdx.y);
    if (e >= elems) return;
    const int r0 = slab * rows_per_slab;
    const int r1 = (r0 + rows_per_slab < batch) ? r0 + rows_per_slab : batch;
    const float* q = in + (long long)r0 * in_stride + e;
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = 0.f;
    int r = r0;
    for (; r + 16 <= r1; r += 16) {       // 16 independent loads in flight per thread
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] += q[j * in_stride];
        q += 16 * in_stride;
    }
    for (; r < r1; ++r) { a[0] += q[0]; q += in_stride; }
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
        for (int j = 0; j < w; ++j) a[j] += a[j + w];
    out[(long long)slab * elems + e] = a[0] * post_scale;
}


#ifndef B2S_EMU
// ---- one-shot all-reduce of the partial mean spectrogram over NVLink peer memory ----------
// Every rank has written its partial sum into its own symmetric buffer (mapped into every peer
// by torch's symmetric-memory rendezvous).  One launch per rank: announce "my partial of this
// epoch is complete" in every peer's signal pad, wait until every peer has announced in mine,
// then read all partials straight from peer memory and add them in rank order -- the result is
// bit-identical on every rank and from run to run.  The buffers alternate between two halves
// (epoch parity), so one announcement per epoch is enough: a rank can only overwrite the half of
// epoch e at epoch e + 2, which it reaches after every peer has announced epoch e + 1, i.e. has
// finished reading epoch e.
struct PeerPtrs {
    const float* buf[16];       // this epoch's partial of every rank
    unsigned* pad[16];          // every rank's signal pad, at the slot block of this epoch's parity
};

// Waiting for a peer is bounded by WALL-CLOCK time (%globaltimer against `timeout_ns`, minutes by
// default: a rank that is legitimately late -- still loading sweeps, first-call allocations, a
// debugger pause -- must not take the job down), polls back off with nanosleep so that the few
// waiting CTAs leave the issue slots of their SMs to the STFT CTAs they share them with, and on
// expiry the kernel sets `*err_flag` (checked by b2s_peer_allreduce_status) and returns without
// writing `out` instead of trapping (a trap poisons the CUDA context of every waiting rank).
B2S_DEVICE unsigned long long b2s_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

B2S_GLOBAL void peer_allreduce_kernel(const PeerPtrs pp, int world, int rank, unsigned epoch, long long elems,
                                      float* __restrict__ out, float post_scale, int vec_ok,
                                      unsigned long long timeout_ns, int* err_flag) {
    if (blockIdx.x == 0 && (int)threadIdx.x < world) {
        __threadfence_system();
        unsigned* dst = pp.pad[threadIdx.x] + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
    }
    int late = 0;
    if ((int)threadIdx.x < world) {
        const unsigned* src = pp.pad[rank] + threadIdx.x;
        unsigned v;
        unsigned ns = 32;
        const unsigned long long t0 = b2s_globaltimer_ns();
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
            if ((int)(v - epoch) >= 0) break;
            if (b2s_globaltimer_ns() - t0 > timeout_ns) {      // a peer never arrived within the window
                late = 1;
                atomicExch(err_flag, (int)(threadIdx.x + 1));     // which rank was missing
                break;
            }
            __nanosleep(ns);
            if (ns < 1024) ns *= 2;
        }
    }
    if (__syncthreads_or(late)) return;
    // peer memory is read with ld.cv (never from a stale L1 line), 16 bytes at a time when the
    // buffers allow it, all ranks' loads in flight before the adds; the adds go in rank order
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    const long long n4 = vec_ok ? elems / 4 : 0;
    for (long long q = tid; q < n4; q += nth) {
        float4 v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r)
            if (r < world) v[r] = __ldcv(reinterpret_cast<const float4*>(pp.buf[r]) + q);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 16; ++r)
            if (r < world) {
                acc.x += v[r].x;
                acc.y += v[r].y;
                acc.z += v[r].z;
                acc.w += v[r].w;
            }
        reinterpret_cast<float4*>(out)[q] = make_float4(acc.x * post_scale, acc.y * post_scale, acc.z * post_scale,
                                                        acc.w * post_scale);
    }
    for (long long e = 4 * n4 + tid; e < elems; e += nth) {
        float acc = 0.f;
        for (int r = 0; r < world; ++r) acc += __ldcv(pp.buf[r] + e);
        out[e] = acc * post_scale;
    }
}

// The same all-reduce as one-warp CTAs of at most 32 registers per thread, for the launch that runs
// BESIDE the persistent STFT grid of the next step (PeerMeanReducer(overlap=True)): three 128-thread STFT
// CTAs of 168 registers leave an SM exactly 1024 registers, i.e. room for one such warp, so these CTAs
// never take the slot of an STFT CTA -- a displaced STFT CTA starts late, and with the sum-fused
// kernel's one-round static schedule a late CTA is a late kernel (measured: +6 us on 170).
// 64-bit peer loads, ranks added in rank order: bit-identical to peer_allreduce_kernel.
B2S_GLOBAL void __maxnreg__(32) peer_allreduce_warp_kernel(const PeerPtrs pp, int world, int rank, unsigned epoch,
                                                                  long long elems, float* __restrict__ out,
                                                                  float post_scale, int vec_ok,
                                                                  unsigned long long timeout_ns, int* err_flag) {
    const int lane = (int)threadIdx.x;
    if (blockIdx.x == 0 && lane < world) {
        __threadfence_system();
        unsigned* dst = pp.pad[lane] + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
    }
    int late = 0;
    if (lane < world) {
        const unsigned* src = pp.pad[rank] + lane;
        unsigned v;
        unsigned ns = 64;
        const unsigned long long t0 = b2s_globaltimer_ns();
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
            if ((int)(v - epoch) >= 0) break;
            if (b2s_globaltimer_ns() - t0 > timeout_ns) {
                late = 1;
                atomicExch(err_flag, lane + 1);
                break;
            }
            __nanosleep(ns);
            if (ns < 2048) ns *= 2;
        }
    }
    if (__any_sync(0xffffffffu, late)) return;
    const long long tid = (long long)blockIdx.x * 32 + lane;
    const long long nth = (long long)gridDim.x * 32;
    const long long n2 = vec_ok ? elems / 2 : 0;
    for (long long q = tid; q < n2; q += nth) {
        float2 acc = make_float2(0.f, 0.f);
        for (int r0 = 0; r0 < world; r0 += 4) {          // four ranks' loads in flight, added in rank order
            float2 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (r0 + j < world) v[j] = __ldcv(reinterpret_cast<const float2*>(pp.buf[r0 + j]) + q);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (r0 + j < world) {
                    acc.x += v[j].x;
                    acc.y += v[j].y;
                }
        }
        reinterpret_cast<float2*>(out)[q] = make_float2(acc.x * post_scale, acc.y * post_scale);
    }
    for (long long e = 2 * n2 + tid; e < elems; e += nth) {
        float acc = 0.f;
        for (int r = 0; r < world; ++r) acc += __ldcv(pp.buf[r] + e);
        out[e] = acc * post_scale;
    }
}
#endif

// ---- display scaling (PlotEngine._plot_spectrogram, PlotEngine.py:126-131) -----------------
// Sxx_norm = clip(S / (base + 1e-20), 0, 1), base = max(S) unless a positive global_max is
// given; with log_scale: 10*log10(Sxx_norm + 1e-12), nan_to_num, min-max to [0, 1] (zeros if
// the dB range is <= 1e-6).  min/max of the dB image follow from min/max of S (monotone), so
// one reduction pass + one elementwise pass suffice.  S >= 0, so float bits order like uints.
B2S_GLOBAL void minmax_init_kernel(unsigned* mm) {
    mm[0] = 0u;             // max
    mm[1] = 0x7f800000u;    // min (+inf)
}

B2S_GLOBAL void minmax_kernel(const float* __restrict__ s, long long elems, unsigned* mm) {
    float mx = 0.f, mn = __uint_as_float(0x7f800000u);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems;
         i += (long long)gridDim.x * blockDim.x) {
        const float v = fmaxf(s[i], 0.f);        // NaN / negative -> 0, like the reference's clip
        mx = fmaxf(mx, v);
        mn = fminf(mn, v);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(mm, __float_as_uint(mx));
        atomicMin(mm + 1, __float_as_uint(mn));
    }
}

B2S_DEVICE float display_norm(float v, float inv_base) { return fminf(fmaxf(v * inv_base, 0.f), 1.f); }

B2S_GLOBAL void display_scale_kernel(const float* __restrict__ s, long long elems, const unsigned* mm, int log_scale,
                                     float global_max, float* __restrict__ out) {
    const float base = (global_max > 0.f) ? global_max : __uint_as_float(mm[0]);
    const float inv_base = 1.0f / (base + 1e-20f);
    float lo = 0.f, inv_range = 0.f;
    if (log_scale) {
        const float hi = 10.0f * log10f(display_norm(__uint_as_float(mm[0]), inv_base) + 1e-12f);
        lo = 10.0f * log10f(display_norm(__uint_as_float(mm[1]), inv_base) + 1e-12f);
        inv_range = (hi - lo > 1e-6f) ? 1.0f / (hi - lo) : 0.f;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems;
         i += (long long)gridDim.x * blockDim.x) {
        float v = display_norm(s[i], inv_base);
        if (log_scale) v = (10.0f * log10f(v + 1e-12f) - lo) * inv_range;
        out[i] = v;
    }
}


// ---- power summaries (PlotEngine.calculate_absolute_power / calculate_band_powers, PlotEngine.py:686-719) ----
// s is the [frames][bins] spectrogram the path left on the device (already cropped to the displayed
// band); band b is the bin range [k0[b], k1[b]) (may be empty), band nb the whole row (the total).
// Negative values are clamped to 0 as the reference does (np.maximum(0, Sxx)).  Two launches, fixed
// summation order, double accumulators: block x sums the frames x, x + gridDim.x, ... (one warp per
// band, lanes stride the bins, xor-butterfly), then one block folds the per-block partials.
constexpr int kMaxBands = 16;
struct BandRanges {
    int k0[kMaxBands + 1];
    int k1[kMaxBands + 1];
};

B2S_GLOBAL void band_sums_kernel(const float* __restrict__ s, long long frames, int bins, const BandRanges br,
                                 int nb, double* __restrict__ partial) {
    const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
    const int nwarps = (int)blockDim.x >> 5;
    for (int b = warp; b <= nb; b += nwarps) {
        double acc = 0.0;
        for (long long f = blockIdx.x; f < frames; f += gridDim.x) {
            const float* row = s + f * bins;
            double a = 0.0;
            for (int k = br.k0[b] + lane; k < br.k1[b]; k += 32) a += (double)fmaxf(row[k], 0.f);
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            acc += a;
        }
        if (lane == 0) partial[(long long)blockIdx.x * (kMaxBands + 1) + b] = acc;
    }
}

B2S_GLOBAL void band_sums_fold_kernel(const double* __restrict__ partial, int nblocks, int nb, double* __restrict__ out) {
    const int b = (int)threadIdx.x;
    if (b > nb) return;
    double acc = 0.0;
    for (int i = 0; i < nblocks; ++i) acc += partial[(long long)i * (kMaxBands + 1) + b];
    out[b] = acc;
}

}  // namespace b2s

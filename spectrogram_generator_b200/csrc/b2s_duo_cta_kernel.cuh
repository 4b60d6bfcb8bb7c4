// Frame-duo STFT -> PSD kernel for nperseg 1024, 2048 and 4096 (any hop): the generic radix-16
// Stockham scheme of stft_psd_kernel (b2s_kernels.cuh) with two CONSECUTIVE frames of a run
// packed in fp32x2 registers (frame A in the low, frame B in the high halves), as in
// b2s_duo_kernel.cuh:
//   * a group of G = nperseg/32 threads (1, 2 or 4 warps) owns a duo, 16 complex points per
//     frame per thread; every butterfly, twiddle multiply and split instruction serves both
//     frames, twiddles / window taps are broadcast scalar operands fetched once per duo;
//   * the exchange buffer holds float4 = (reA, reB, imA, imB) per point, padded one slot per
//     16 (conflict-free for the 128-bit row stores, the strided loads and the Stockham scatter);
//   * the window lives in shared memory (one copy per CTA), not in registers: the data
//     registers are 64 per thread and the kernel runs 3 CTAs x 128 threads per SM.
// Frames are gathered straight from global memory (the overlap between A, B and the next duo
// is served by L1/L2); detrend is the two-pass fp32 scheme of the other kernels, with the sums
// taken in a fixed order per frame so that a frame's result does not depend on the chunking.
#pragma once

#include "b2s_duo_kernel.cuh"

namespace b2s {

template <int R> struct SmallFft2;
template <> struct SmallFft2<1> { B2S_HD static void run(cpx2 (&)[1]) {} };
template <> struct SmallFft2<2> {
    B2S_HD static void run(cpx2 (&v)[2]) {
        const cpx2 t = v[0];
        v[0] = c2add(t, v[1]);
        v[1] = c2sub(t, v[1]);
    }
};
template <> struct SmallFft2<4> { B2S_HD static void run(cpx2 (&v)[4]) { c2radix4(v[0], v[1], v[2], v[3]); } };
template <> struct SmallFft2<8> {
    B2S_HD static void run(cpx2 (&v)[8]) {
        c2radix4(v[0], v[2], v[4], v[6]);   // E[q] in v[2q]
        c2radix4(v[1], v[3], v[5], v[7]);   // O[q] in v[2q+1]
        const cpx2 o0 = v[1], o1 = c2mul_w8_1(v[3]), o2 = c2mul_mi(v[5]), o3 = c2mul_w8_3(v[7]);
        const cpx2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
        v[0] = c2add(e0, o0); v[4] = c2sub(e0, o0);
        v[1] = c2add(e1, o1); v[5] = c2sub(e1, o1);
        v[2] = c2add(e2, o2); v[6] = c2sub(e2, o2);
        v[3] = c2add(e3, o3); v[7] = c2sub(e3, o3);
    }
};

template <int LOG2N>
struct DuoCtaPlan {
    using PL = Plan<LOG2N>;
    static constexpr int N = PL::N, M = PL::M, G = PL::G;
    static constexpr int NT = (G > 128) ? G : 128;       // threads per CTA
    static constexpr int MINB = 3;
    static constexpr int FPC = NT / G;                   // duos in flight per CTA
    static constexpr int RED = (G > 32) ? G / 32 : 1;    // warps per group
    static constexpr int BUF = M + (M >> 4);             // padded float4 slots per duo
    // shared memory: [window M float2][FPC * BUF float4][FPC * 3 * RED float2 reduction slots]
    static constexpr size_t OFF_BUF = (size_t)M * sizeof(float2);
    static constexpr size_t OFF_RED = OFF_BUF + (size_t)FPC * BUF * sizeof(float4);
    static constexpr size_t SMEM = OFF_RED + (size_t)FPC * (3 * RED + 1) * sizeof(float2);   // + the unit draw
    static_assert(G >= 32 && G <= 128, "duo CTA kernel: nperseg 1024 .. 4096");
    static_assert(PL::P == 2, "two radix-16 passes");
};

B2S_HD int phys4(int e) { return e + (e >> 4); }

// real-FFT split of the bin pair (k, M - k) of both frames and its power: Z[k] = zk, Z[M-k] = zm, w = W_N^k
B2S_DEVICE void duo_pair_power(cpx2 zk, cpx2 zm, float2 w, float sc, float2& pa, float2& pb) {
    const cpx2 e{pk_add(zk.re, zm.re), pk_sub(zk.im, zm.im)};      // 2E = zk + conj(zm)
    const cpx2 o{pk_add(zk.im, zm.im), pk_sub(zm.re, zk.re)};      // 2O = -i (zk - conj(zm))
    const cpx2 t = c2mul(o, w);
    const cpx2 a = c2add(e, t), bq = c2sub(e, t);
    pa = pk_fma(a.re, a.re, pk_mul(a.im, a.im));
    pb = pk_fma(bq.re, bq.re, pk_mul(bq.im, bq.im));
    if (sc != 1.0f) {
        pa = pk_muls(pa, sc);
        pb = pk_muls(pb, sc);
    }
}

template <int MODE>
struct EpiDuo {
    float* rowA;        // frame A row (already offset by -kmin); frame B row is rowA + kout
    int kout;
    float floor;
    float2 band;        // MODE 2 partial sums (A, B)
    int kmin, kmax, db;
    bool actA, actB;
    B2S_DEVICE void put(int k, float2 p) {
        if constexpr (MODE == EPI_GENERAL) {
            if (db) p = cmk(10.0f * log10f(fmaxf(p.x, floor)), 10.0f * log10f(fmaxf(p.y, floor)));
            if (k >= kmin && k <= kmax) {
                if (actA) rowA[k] = p.x;
                if (actB) rowA[k + kout] = p.y;
            }
        } else if constexpr (MODE == EPI_BAND) {
            if (k >= kmin && k <= kmax) band = pk_add(band, p);
        } else {
            if (actA) rowA[k] = p.x;
            if (actB) rowA[k + kout] = p.y;
        }
    }
    // Z[k] = zk, Z[M-k] = zm, w = W_N^k; the window carries sqrt(scale/2), so |2 X|^2 is the PSD
    B2S_DEVICE void pair(int k, int mk, cpx2 zk, cpx2 zm, float2 w, float sc) {
        float2 pa, pb;
        duo_pair_power(zk, zm, w, sc, pa, pb);
        put(k, pa);
        put(mk, pb);
    }
};

// The plain epilogue that also keeps what it stored (SUM mode of b2s_duo4_kernel.cuh): the NREC packed power
// values of a final-stage task, in the order they were put -- indices are compile-time after unrolling.
template <int NREC>
struct EpiDuoRec {
    float* rowA;
    int kout;
    float floor;
    float2 band;
    int kmin, kmax, db;
    bool actA, actB;
    int nrec;
    float2 rec[NREC];
    B2S_DEVICE void put(int k, float2 p) {
        if (actA) rowA[k] = p.x;
        if (actB) rowA[k + kout] = p.y;
        rec[nrec++] = p;
    }
    B2S_DEVICE void pair(int k, int mk, cpx2 zk, cpx2 zm, float2 w, float sc) {
        float2 pa, pb;
        duo_pair_power(zk, zm, w, sc, pa, pb);
        put(k, pa);
        put(mk, pb);
    }
};

// Sum over the group of a packed (A, B) value: xor-butterfly inside the warp, then a fixed-order
// sum of the per-warp partials through `red` (RED float2 slots).
template <int LOG2N>
B2S_DEVICE float2 duo_group_sum(float2 v, int grp, int j, unsigned lane, float2* red) {
    using DP = DuoCtaPlan<LOG2N>;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        v = pk_add(v, cmk(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o)));
    if constexpr (DP::G > 32) {
        if (lane == 0) red[j >> 5] = v;
        b2s_bar_sync(grp + 1, DP::G);
        v = red[0];
#pragma unroll
        for (int w = 1; w < DP::RED; ++w) v = pk_add(v, red[w]);
    }
    return v;
}

template <int G>
B2S_DEVICE void duo_group_sync(int grp) {
    if constexpr (G <= 32) __syncwarp();
    else b2s_bar_sync(grp + 1, G);
}

template <int LOG2N, typename Tin, int MODE>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(DuoCtaPlan<LOG2N>::NT, DuoCtaPlan<LOG2N>::MINB)
stft_psd_duo_cta_kernel(const StftParams p) {
    using PL = Plan<LOG2N>;
    using DP = DuoCtaPlan<LOG2N>;
    constexpr int M = PL::M, G = PL::G, NS = PL::NS, GF = PL::GF, N = PL::N;

    B2S_DYN_SMEM(smem_raw);
    const int tid = (int)threadIdx.x;
    const int grp = tid / G;
    const int j = tid - grp * G;
    const unsigned lane = (unsigned)tid & 31u;
    float2* const swin = reinterpret_cast<float2*>(smem_raw);
    float4* const buf = reinterpret_cast<float4*>(smem_raw + DP::OFF_BUF) + (size_t)grp * DP::BUF;
    float2* const red = reinterpret_cast<float2*>(smem_raw + DP::OFF_RED) + grp * (3 * DP::RED + 1);

    // window (times sqrt(scale/2): |2 X|^2 is then the PSD of an interior bin), once per CTA
    {
        const float csc = sqrtf(0.5f * p.scale);
        const float2* w2 = reinterpret_cast<const float2*>(p.window);
        for (int i = tid; i < M; i += DP::NT) {
            const float2 w = __ldg(w2 + i);
            swin[i] = cmk(w.x * csc, w.y * csc);
        }
    }
    __syncthreads();

    const int kout = p.kmax - p.kmin + 1;
    EpiDuo<MODE> epi;
    epi.kout = kout;
    epi.floor = p.db_floor;
    epi.kmin = p.kmin;
    epi.kmax = p.kmax;
    epi.db = p.out_mode;
    epi.band = cmk(0.f, 0.f);

    // work units: static round-robin over the grid, or (p.work) an atomic counter every duo draws
    // from; the next draw is issued a whole unit ahead, so its latency is hidden
    const bool dyn = p.work != nullptr;
    auto draw = [&]() -> long long {
        int b0 = 0;
        if constexpr (G <= 32) {
            if (j == 0) b0 = atomicAdd(p.work, 1);
            b0 = __shfl_sync(0xffffffffu, b0, 0);
        } else {
            int* const slot = reinterpret_cast<int*>(red + 3 * DP::RED);
            if (j == 0) *slot = atomicAdd(p.work, 1);
            b2s_bar_sync(grp + 1, G);
            b0 = *slot;
            b2s_bar_sync(grp + 1, G);
        }
        return (long long)b0;
    };
    long long u_next = dyn ? draw() : (long long)blockIdx.x * DP::FPC + grp;
    while (u_next < p.n_units) {
        const long long u = u_next;
        u_next = dyn ? draw() : u + (long long)gridDim.x * DP::FPC;
        const long long b = u / p.units_per_signal;
        const int c = (int)(u - b * p.units_per_signal);
        const int f_begin = c * p.chunk_frames;
        const int f_end = (f_begin + p.chunk_frames < p.nframes) ? f_begin + p.chunk_frames : p.nframes;
        const Tin* const xb = reinterpret_cast<const Tin*>(p.x) + b * p.x_batch_stride + p.frame0 * (long long)p.hop;
        float* const ob = p.out + b * p.out_batch_stride - ((MODE == EPI_BAND) ? 0 : p.kmin);

        for (int f = f_begin; f < f_end; f += 2) {
            const bool actB = f + 1 < f_end;
            epi.actA = true;
            epi.actB = actB;
            epi.rowA = ob + (long long)f * kout;
            const Tin* const xa = xb + (long long)f * p.hop + 2 * j;
            const Tin* const xq = actB ? xa + p.hop : xa;       // no frame B: recompute A, stores off

            // ---- gather both frames: z[n] = x[2n] + i x[2n+1], n = j + G r ----
            cpx2 v[16];
            if (p.vec_ok) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float2 a = Loader<Tin>::ld2(xa + 2 * G * r), bq = Loader<Tin>::ld2(xq + 2 * G * r);
                    v[r].re = cmk(a.x, bq.x);
                    v[r].im = cmk(a.y, bq.y);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    v[r].re = cmk(Loader<Tin>::ld1(xa + 2 * G * r), Loader<Tin>::ld1(xq + 2 * G * r));
                    v[r].im = cmk(Loader<Tin>::ld1(xa + 2 * G * r + 1), Loader<Tin>::ld1(xq + 2 * G * r + 1));
                }
            }

            // ---- detrend (two fp32 passes, see stft_psd_kernel) + window ----
            if (p.detrend) {
                float2 s[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) s[r] = pk_add(v[r].re, v[r].im);
#pragma unroll
                for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                    for (int r = 0; r < w; ++r) s[r] = pk_add(s[r], s[r + w]);
                const float2 m1 = pk_muls(duo_group_sum<LOG2N>(s[0], grp, j, lane, red), 1.0f / (float)N);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    v[r].re = pk_sub(v[r].re, m1);
                    v[r].im = pk_sub(v[r].im, m1);
                }
#pragma unroll
                for (int r = 0; r < 16; ++r) s[r] = pk_add(v[r].re, v[r].im);
#pragma unroll
                for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                    for (int r = 0; r < w; ++r) s[r] = pk_add(s[r], s[r + w]);
                const float2 nr = pk_muls(duo_group_sum<LOG2N>(s[0], grp, j, lane, red + DP::RED), -1.0f / (float)N);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float2 w = swin[j + G * r];
                    v[r].re = pk_fmas(v[r].re, w.x, pk_muls(nr, w.x));
                    v[r].im = pk_fmas(v[r].im, w.y, pk_muls(nr, w.y));
                }
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float2 w = swin[j + G * r];
                    v[r].re = pk_muls(v[r].re, w.x);
                    v[r].im = pk_muls(v[r].im, w.y);
                }
            }

            // ---- pass 0: radix-16 over r (stride G), Ns 1 -> 16 ----
            c2radix16(v);
            duo_group_sync<G>(grp);            // the previous duo's final-stage reads are done
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const cpx2 z = v[perm16(q)];
                buf[17 * j + q] = make_float4(z.re.x, z.re.y, z.im.x, z.im.y);
            }
            duo_group_sync<G>(grp);

            // ---- pass 1: radix-16 Stockham, Ns = 16 -> 256 ----
            {
                const int jm = j & 15;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float4 q4 = buf[phys4(j + r * G)];
                    v[r] = cpx2{cmk(q4.x, q4.y), cmk(q4.z, q4.w)};
                }
                const float2* const twp = p.tw + PL::OFF_P1 + jm;
#pragma unroll
                for (int r = 1; r < 16; ++r) v[r] = c2mul(v[r], __ldg(twp + (r - 1) * 16));
                c2radix16(v);
                duo_group_sync<G>(grp);
                const int base = (j - jm) * 16 + jm;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const cpx2 z = v[perm16(r)];
                    buf[phys4(base + r * 16)] = make_float4(z.re.x, z.re.y, z.im.x, z.im.y);
                }
                duo_group_sync<G>(grp);
            }

            // ---- fused final stage: radix-GF butterflies + real-FFT split + PSD, both frames ----
#pragma unroll
            for (int cc = 0; cc < PL::TPT; ++cc) {
                const int kap = j + G * cc;            // task id == kappa in [0, NS/2)
                cpx2 U[GF], V[GF];
                if (kap != 0) {
                    const int kap2 = NS - kap;
#pragma unroll
                    for (int r = 0; r < GF; ++r) {
                        const float4 qa = buf[phys4(kap + r * NS)], qb = buf[phys4(kap2 + r * NS)];
                        U[r] = cpx2{cmk(qa.x, qa.y), cmk(qa.z, qa.w)};
                        V[r] = cpx2{cmk(qb.x, qb.y), cmk(qb.z, qb.w)};
                    }
#pragma unroll
                    for (int r = 1; r < GF; ++r) {
                        U[r] = c2mul(U[r], __ldg(p.tw + PL::OFF_FIN + (r - 1) * NS + kap));
                        V[r] = c2mul(V[r], __ldg(p.tw + PL::OFF_FIN + (r - 1) * NS + kap2));
                    }
                    SmallFft2<GF>::run(U);
                    SmallFft2<GF>::run(V);
#pragma unroll
                    for (int a = 0; a < GF; ++a) {
                        const int k = kap + a * NS;
                        epi.pair(k, M - k, U[a], V[GF - 1 - a], __ldg(p.tw + PL::OFF_POST + k), 1.0f);
                    }
                } else {
                    // kappa = 0 and kappa = NS/2 are their own mirrors (thread 0 of the group)
#pragma unroll
                    for (int r = 0; r < GF; ++r) {
                        const float4 qa = buf[phys4(r * NS)], qb = buf[phys4(NS / 2 + r * NS)];
                        U[r] = cpx2{cmk(qa.x, qa.y), cmk(qa.z, qa.w)};
                        V[r] = cpx2{cmk(qb.x, qb.y), cmk(qb.z, qb.w)};
                    }
#pragma unroll
                    for (int r = 1; r < GF; ++r) V[r] = c2mul(V[r], __ldg(p.tw + PL::OFF_FIN + (r - 1) * NS + NS / 2));
                    SmallFft2<GF>::run(U);
                    SmallFft2<GF>::run(V);
                    // DC / Nyquist: the pair formula with zm = zk gives 2 (Re +- Im); they carry scale, not 2 scale
                    epi.pair(0, M, U[0], U[0], cmk(1.f, 0.f), 0.5f);
#pragma unroll
                    for (int a = 1; 2 * a < GF; ++a)
                        epi.pair(a * NS, M - a * NS, U[a], U[GF - a], __ldg(p.tw + PL::OFF_POST + a * NS), 1.0f);
                    if constexpr (GF % 2 == 0) {       // k = M/2: X = conj(Z)
                        const cpx2 z = U[GF / 2];
                        epi.put(M / 2, pk_muls(pk_fma(z.re, z.re, pk_mul(z.im, z.im)), 4.0f));
                    }
#pragma unroll
                    for (int a = 0; 2 * a < GF - 1; ++a) {
                        const int k = NS / 2 + a * NS;
                        epi.pair(k, M - k, V[a], V[GF - 1 - a], __ldg(p.tw + PL::OFF_POST + k), 1.0f);
                    }
                    if constexpr (GF % 2 == 1) {
                        const cpx2 z = V[(GF - 1) / 2];
                        epi.put(NS / 2 + ((GF - 1) / 2) * NS, pk_muls(pk_fma(z.re, z.re, pk_mul(z.im, z.im)), 4.0f));
                    }
                }
            }
            if constexpr (MODE == EPI_BAND) {
                const float2 bs = duo_group_sum<LOG2N>(epi.band, grp, j, lane, red + 2 * DP::RED);
                epi.band = cmk(0.f, 0.f);
                if (j == 0) {
                    ob[f] = bs.x;
                    if (actB) ob[f + 1] = bs.y;
                }
            }
        }
    }
    if (dyn) {      // the last CTA to finish re-arms the counters for the next launch that uses them
        __syncthreads();
        if (tid == 0) {
            const int done = atomicAdd(p.work + 1, 1);
            if (done == (int)gridDim.x - 1) {
                p.work[0] = 0;
                p.work[1] = 0;
            }
        }
    }
}

}  // namespace b2s

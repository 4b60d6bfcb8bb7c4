// Frame-pair STFT -> PSD kernel for nperseg 256 and 512 (float samples, hop a multiple of
// nperseg/8): two CONSECUTIVE frames of a signal are transformed together, one in the low and
// one in the high half of packed fp32x2 registers.
//
// Why (ncu, profiles/r1_c2_v3*): the warp kernel is issue-bound (0.75 of the issue slots, 520
// warp-instructions per frame, 360 of them FP32) with the L1/shared data pipe at 73 %.  On
// sm_100a FADD2/FMUL2/FFMA2 do two fp32 operations per lane per issue slot and accept a scalar
// (broadcast) operand, so with frame A in .x and frame B in .y of every value
//   * every arithmetic instruction of the transform serves two frames (twiddles and window
//     taps are the same for both and enter as broadcast operands);
//   * every LDS.128 / STS.128 of the exchange moves one complex point of both frames;
//   * each lane holds 8 complex points per frame (M = N/2 = 8 * 8 * g, g = 2 or 4), so the data
//     registers stay at 32 and the window taps, pass-1 twiddles, final-stage twiddles and
//     real-FFT-split twiddles of the lane fit in registers -- no table reads per frame at all;
//   * consecutive frames overlap, so the raw samples of the pair are 8 + S register slots
//     (S = hop / (nperseg/8)); the next pair needs 2 S new loads per lane, issued one
//     iteration ahead (each sample is loaded from global memory exactly once per run).
//
// Arithmetic per frame is the same Stockham / fused-final scheme as the other kernels (radix-8
// passes instead of radix-16), rounding is IEEE round-to-nearest in both halves.
#pragma once

#include "b2s_kernels.cuh"

namespace b2s {

// ---- packed (two-frame) helpers -----------------------------------------------------------
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
B2S_HD float2 pk_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
B2S_HD float2 pk_sub(float2 a, float2 b) { return __fadd2_rn(a, cmk(-b.x, -b.y)); }
B2S_HD float2 pk_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
B2S_HD float2 pk_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
B2S_HD float2 pk_muls(float2 a, float s) { return __fmul2_rn(a, cmk(s, s)); }
B2S_HD float2 pk_fmas(float2 a, float s, float2 c) { return __ffma2_rn(a, cmk(s, s), c); }
#else
B2S_HD float2 pk_add(float2 a, float2 b) { return cmk(a.x + b.x, a.y + b.y); }
B2S_HD float2 pk_sub(float2 a, float2 b) { return cmk(a.x - b.x, a.y - b.y); }
B2S_HD float2 pk_mul(float2 a, float2 b) { return cmk(a.x * b.x, a.y * b.y); }
B2S_HD float2 pk_fma(float2 a, float2 b, float2 c) { return cmk(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
B2S_HD float2 pk_muls(float2 a, float s) { return cmk(a.x * s, a.y * s); }
B2S_HD float2 pk_fmas(float2 a, float s, float2 c) { return cmk(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y)); }
#endif
B2S_HD float2 pk_neg(float2 a) { return cmk(-a.x, -a.y); }

// one complex value of frame A (.x halves) and frame B (.y halves)
struct cpx2 {
    float2 re, im;
};
B2S_HD cpx2 c2add(cpx2 a, cpx2 b) { return cpx2{pk_add(a.re, b.re), pk_add(a.im, b.im)}; }
B2S_HD cpx2 c2sub(cpx2 a, cpx2 b) { return cpx2{pk_sub(a.re, b.re), pk_sub(a.im, b.im)}; }
// * (w.x + i w.y), the same twiddle for both frames
B2S_HD cpx2 c2mul(cpx2 a, float2 w) {
    return cpx2{pk_fmas(a.re, w.x, pk_neg(pk_muls(a.im, w.y))), pk_fmas(a.re, w.y, pk_muls(a.im, w.x))};
}
B2S_HD cpx2 c2mul_mi(cpx2 a) { return cpx2{a.im, pk_neg(a.re)}; }   // * (-i)
B2S_HD cpx2 c2mul_w8_1(cpx2 a) {
    return cpx2{pk_muls(pk_add(a.re, a.im), B2S_SQRT1_2), pk_muls(pk_sub(a.im, a.re), B2S_SQRT1_2)};
}
B2S_HD cpx2 c2mul_w8_3(cpx2 a) {
    return cpx2{pk_muls(pk_sub(a.im, a.re), B2S_SQRT1_2), pk_muls(pk_add(a.re, a.im), -B2S_SQRT1_2)};
}

B2S_HD void c2radix2(cpx2& a0, cpx2& a1) {
    const cpx2 t = a0;
    a0 = c2add(t, a1);
    a1 = c2sub(t, a1);
}
B2S_HD void c2radix4(cpx2& a0, cpx2& a1, cpx2& a2, cpx2& a3) {
    const cpx2 t0 = c2add(a0, a2), t1 = c2sub(a0, a2);
    const cpx2 t2 = c2add(a1, a3), t3 = c2sub(a1, a3);
    a0 = c2add(t0, t2);
    a2 = c2sub(t0, t2);
    a1 = cpx2{pk_add(t1.re, t3.im), pk_sub(t1.im, t3.re)};
    a3 = cpx2{pk_sub(t1.re, t3.im), pk_add(t1.im, t3.re)};
}
// 8-point DFT in place; X[k] is left in v[perm8(k)]
B2S_HD constexpr int perm8(int k) { return (k < 4) ? 2 * k : 2 * (k - 4) + 1; }
B2S_HD void c2radix8(cpx2 (&v)[8]) {
    c2radix4(v[0], v[2], v[4], v[6]);      // E[q] in v[2q]
    c2radix4(v[1], v[3], v[5], v[7]);      // O[q] in v[2q+1]
    v[3] = c2mul_w8_1(v[3]);
    v[5] = c2mul_mi(v[5]);
    v[7] = c2mul_w8_3(v[7]);
#pragma unroll
    for (int q = 0; q < 4; ++q) c2radix2(v[2 * q], v[2 * q + 1]);   // X[q] -> v[2q], X[q+4] -> v[2q+1]
}
template <int R> struct SmallFft2;
template <> struct SmallFft2<2> { B2S_HD static void run(cpx2 (&v)[2]) { c2radix2(v[0], v[1]); } };
template <> struct SmallFft2<4> { B2S_HD static void run(cpx2 (&v)[4]) { c2radix4(v[0], v[1], v[2], v[3]); } };

template <int LOG2N>
struct PairPlan {
    static constexpr int N = 1 << LOG2N;
    static constexpr int M = N / 2;
    static constexpr int G = M / 8;                 // lanes per frame pair (16 or 32)
    static constexpr int NS = 64;                   // sub-transform length before the final stage
    static constexpr int GF = M / NS;               // final radix (2 or 4)
    static constexpr int TPT = (NS / 2) / G;        // final tasks per lane (2 or 1)
    static constexpr int NT = 256;
    static constexpr int FPC = NT / G;              // pair groups per CTA
    static constexpr int BUF = M + (M >> 3);        // padded 16-byte slots (one complex of both frames)
    static constexpr size_t SMEM = (size_t)FPC * BUF * sizeof(float4);
    static_assert(G == 16 || G == 32, "pair kernel: nperseg 256 or 512");
    static_assert(GF == 2 || GF == 4, "pair kernel plan");
};

B2S_HD int phys8(int e) { return e + (e >> 3); }

template <int MODE>
struct Epi2 {
    float* rowA;        // frame A row (already offset by -kmin); frame B row is rowA + kout
    int kout;
    float s_edge, s_int, floor;
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
    B2S_DEVICE void pair(int k, int mk, cpx2 zk, cpx2 zm, float2 w, float sc) {
        const cpx2 e{pk_add(zk.re, zm.re), pk_sub(zk.im, zm.im)};      // 2E = zk + conj(zm)
        const cpx2 o{pk_add(zk.im, zm.im), pk_sub(zm.re, zk.re)};      // 2O = -i (zk - conj(zm))
        const cpx2 t = c2mul(o, w);
        const cpx2 a = c2add(e, t), bq = c2sub(e, t);
        put(k, pk_muls(pk_fma(a.re, a.re, pk_mul(a.im, a.im)), sc));
        put(mk, pk_muls(pk_fma(bq.re, bq.re, pk_mul(bq.im, bq.im)), sc));
    }
};

// S = hop / (nperseg / 8): slots the frame start advances per frame (1, 2, 4, 7 or 8)
template <int LOG2N, int S, int MODE>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(256, 2) stft_psd_pair_kernel(const StftParams p) {
    using PP = PairPlan<LOG2N>;
    constexpr int N = PP::N, M = PP::M, G = PP::G, NS = PP::NS, GF = PP::GF;
    constexpr int NCUR = 8 + S;                      // raw complex slots of the pair

    B2S_DYN_SMEM_F4(sm4);
    const int tid = (int)threadIdx.x;
    const int grp = tid / G;
    const int t = tid - grp * G;
    float4* const buf = sm4 + grp * PP::BUF;
    const float2* const W = p.tw;                    // W_N^j, j < N (direct table)

    // ---- per-lane constants, loaded once per kernel ----
    float2 win[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) win[r] = __ldg(reinterpret_cast<const float2*>(p.window) + (t + G * r));
    float2 tw1[8];                                   // W_64^(r (t mod 8))
#pragma unroll
    for (int r = 1; r < 8; ++r) tw1[r] = __ldg(W + ((r * (t & 7)) * (N / 64)) % N);
    float2 twf[PP::TPT][GF];                         // W_M^(r kappa)
    float2 twp[PP::TPT][GF];                         // W_N^(kappa + a NS)
#pragma unroll
    for (int c = 0; c < PP::TPT; ++c) {
        const int kap = t + G * c;
#pragma unroll
        for (int r = 0; r < GF; ++r) {
            twf[c][r] = __ldg(W + (2 * r * kap) % N);
            twp[c][r] = __ldg(W + (kap + r * NS));
        }
    }

    const int kout = p.kmax - p.kmin + 1;
    Epi2<MODE> epi;
    epi.kout = kout;
    epi.s_edge = p.scale;
    epi.s_int = 2.0f * p.scale;
    epi.floor = p.db_floor;
    epi.kmin = p.kmin;
    epi.kmax = p.kmax;
    epi.db = p.out_mode;
    epi.band = cmk(0.f, 0.f);
    const float sc_int = 0.25f * epi.s_int;

    constexpr int GW = 32 / G;
    const long long ustride = (long long)gridDim.x * PP::FPC;
    for (long long ub = (long long)blockIdx.x * PP::FPC + (grp - (grp % GW)); ub < p.n_units; ub += ustride) {
        long long u = ub + (grp % GW);
        const bool uvalid = u < p.n_units;
        if (!uvalid) u = p.n_units - 1;
        const long long b = u / p.units_per_signal;
        const int c = (int)(u - b * p.units_per_signal);
        const int f_begin = c * p.chunk_frames;
        const int f_end = (f_begin + p.chunk_frames < p.nframes) ? f_begin + p.chunk_frames : p.nframes;
        const float* const xb = reinterpret_cast<const float*>(p.x) + b * p.x_batch_stride + p.frame0 * (long long)p.hop;
        float* const ob = p.out + b * p.out_batch_stride - p.kmin;

        // raw samples of the pair: slot i <-> complex index t + G i relative to frame f
        float2 cur[NCUR];
        {
            const float* const xf = xb + (long long)f_begin * p.hop;
            const bool hasB = f_begin + 1 < f_end;
#pragma unroll
            for (int i = 0; i < NCUR; ++i)
                cur[i] = (i < 8 || hasB) ? __ldg(reinterpret_cast<const float2*>(xf + 2 * (t + G * i))) : cmk(0.f, 0.f);
        }

        for (int f = f_begin;; f += 2) {
            const bool actA = uvalid && (f < f_end);
            const bool actB = uvalid && (f + 1 < f_end);
            if (!__any_sync(0xffffffffu, actA)) break;
            epi.actA = actA;
            epi.actB = actB;
            epi.rowA = ob + (long long)f * kout;

            // ---- detrend + window, packing frame A (slots 0..7) and B (slots S..S+7) ----
            cpx2 v[8];
            if (p.detrend) {
                float sl[NCUR];
#pragma unroll
                for (int i = 0; i < NCUR; ++i) sl[i] = cur[i].x + cur[i].y;
                float sa = 0.f, sb = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) { sa += sl[i]; sb += sl[i + S]; }
#pragma unroll
                for (int o = G / 2; o >= 1; o >>= 1) {
                    sa += __shfl_xor_sync(0xffffffffu, sa, o);
                    sb += __shfl_xor_sync(0xffffffffu, sb, o);
                }
                const float mA = sa * (1.0f / (float)N), mB = sb * (1.0f / (float)N);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    v[r].re = cmk(cur[r].x - mA, cur[r + S].x - mB);
                    v[r].im = cmk(cur[r].y - mA, cur[r + S].y - mB);
                }
                float2 s2 = pk_add(v[0].re, v[0].im);
#pragma unroll
                for (int r = 1; r < 8; ++r) s2 = pk_add(s2, pk_add(v[r].re, v[r].im));
#pragma unroll
                for (int o = G / 2; o >= 1; o >>= 1)
                    s2 = pk_add(s2, cmk(__shfl_xor_sync(0xffffffffu, s2.x, o), __shfl_xor_sync(0xffffffffu, s2.y, o)));
                const float2 nr = pk_muls(s2, -1.0f / (float)N);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    v[r].re = pk_fmas(v[r].re, win[r].x, pk_muls(nr, win[r].x));
                    v[r].im = pk_fmas(v[r].im, win[r].y, pk_muls(nr, win[r].y));
                }
            } else {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    v[r].re = cmk(cur[r].x * win[r].x, cur[r + S].x * win[r].x);
                    v[r].im = cmk(cur[r].y * win[r].y, cur[r + S].y * win[r].y);
                }
            }

            // ---- next pair (frames f+2, f+3): keep the overlap, prefetch 2 S new slots ----
            if (actA && f + 2 < f_end) {
                const float* const xn = xb + (long long)(f + 2) * p.hop;
                const bool hasB = f + 3 < f_end;
#pragma unroll
                for (int i = 0; i + 2 * S < NCUR; ++i) cur[i] = cur[i + 2 * S];
#pragma unroll
                for (int i = (NCUR > 2 * S ? NCUR - 2 * S : 0); i < NCUR; ++i)
                    cur[i] = (i < 8 || hasB) ? __ldg(reinterpret_cast<const float2*>(xn + 2 * (t + G * i))) : cmk(0.f, 0.f);
            }

            // ---- pass 0: radix-8 over r (stride G), Ns 1 -> 8 ----
            c2radix8(v);
            __syncwarp();                        // previous pair's final-stage reads are done
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const cpx2 z = v[perm8(r)];
                buf[9 * t + r] = make_float4(z.re.x, z.re.y, z.im.x, z.im.y);
            }
            __syncwarp();

            // ---- pass 1: radix-8 Stockham, Ns = 8 -> 64 ----
            {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float4 q = buf[phys8(t + r * G)];
                    v[r] = cpx2{cmk(q.x, q.y), cmk(q.z, q.w)};
                }
#pragma unroll
                for (int r = 1; r < 8; ++r) v[r] = c2mul(v[r], tw1[r]);
                c2radix8(v);
                __syncwarp();
                const int jm = t & 7;
                const int base = (t - jm) * 8 + jm;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const cpx2 z = v[perm8(r)];
                    buf[phys8(base + r * 8)] = make_float4(z.re.x, z.re.y, z.im.x, z.im.y);
                }
                __syncwarp();
            }

            // ---- fused final stage: radix-GF butterflies + real-FFT split + PSD, both frames ----
#pragma unroll
            for (int cc = 0; cc < PP::TPT; ++cc) {
                const int kap = t + G * cc;
                cpx2 U[GF], V[GF];
                if (kap != 0) {
                    const int kap2 = NS - kap;
#pragma unroll
                    for (int r = 0; r < GF; ++r) {
                        const float4 qa = buf[phys8(kap + r * NS)], qb = buf[phys8(kap2 + r * NS)];
                        U[r] = cpx2{cmk(qa.x, qa.y), cmk(qa.z, qa.w)};
                        V[r] = cpx2{cmk(qb.x, qb.y), cmk(qb.z, qb.w)};
                    }
#pragma unroll
                    for (int r = 1; r < GF; ++r) {
                        U[r] = c2mul(U[r], twf[cc][r]);
                        // W_M^(r (NS - kappa)) = W_GF^r * conj(W_M^(r kappa))
                        cpx2 vv = c2mul(V[r], cmk(twf[cc][r].x, -twf[cc][r].y));
                        if constexpr (GF == 2) {
                            V[r] = cpx2{pk_neg(vv.re), pk_neg(vv.im)};                  // * (-1)
                        } else {
                            if (r == 1) V[r] = c2mul_mi(vv);                             // * (-i)
                            else if (r == 2) V[r] = cpx2{pk_neg(vv.re), pk_neg(vv.im)};  // * (-1)
                            else V[r] = cpx2{pk_neg(vv.im), vv.re};                      // * (+i)
                        }
                    }
                    SmallFft2<GF>::run(U);
                    SmallFft2<GF>::run(V);
#pragma unroll
                    for (int a = 0; a < GF; ++a) {
                        const int k = kap + a * NS;
                        epi.pair(k, M - k, U[a], V[GF - 1 - a], twp[cc][a], sc_int);
                    }
                } else {
                    // kappa = 0 and kappa = NS/2 are their own mirrors (lane 0 of the group)
#pragma unroll
                    for (int r = 0; r < GF; ++r) {
                        const float4 qa = buf[phys8(r * NS)], qb = buf[phys8(NS / 2 + r * NS)];
                        U[r] = cpx2{cmk(qa.x, qa.y), cmk(qa.z, qa.w)};
                        V[r] = cpx2{cmk(qb.x, qb.y), cmk(qb.z, qb.w)};
                    }
#pragma unroll
                    for (int r = 1; r < GF; ++r) V[r] = c2mul(V[r], __ldg(W + (2 * r * (NS / 2)) % N));
                    SmallFft2<GF>::run(U);
                    SmallFft2<GF>::run(V);
                    {
                        const float2 a = pk_add(U[0].re, U[0].im), bq = pk_sub(U[0].re, U[0].im);
                        epi.put(0, pk_muls(pk_mul(a, a), epi.s_edge));
                        epi.put(M, pk_muls(pk_mul(bq, bq), epi.s_edge));
                    }
#pragma unroll
                    for (int a = 1; 2 * a < GF; ++a)
                        epi.pair(a * NS, M - a * NS, U[a], U[GF - a], __ldg(W + a * NS), sc_int);
                    {
                        const cpx2 z = U[GF / 2];                                   // k = M/2: X = conj(Z)
                        epi.put(M / 2, pk_muls(pk_fma(z.re, z.re, pk_mul(z.im, z.im)), epi.s_int));
                    }
#pragma unroll
                    for (int a = 0; 2 * a < GF - 1; ++a) {
                        const int k = NS / 2 + a * NS;
                        epi.pair(k, M - k, V[a], V[GF - 1 - a], __ldg(W + k), sc_int);
                    }
                }
            }
            if constexpr (MODE == EPI_BAND) {
                float2 bs = epi.band;
                epi.band = cmk(0.f, 0.f);
#pragma unroll
                for (int o = G / 2; o >= 1; o >>= 1)
                    bs = pk_add(bs, cmk(__shfl_xor_sync(0xffffffffu, bs.x, o), __shfl_xor_sync(0xffffffffu, bs.y, o)));
                if (t == 0) {
                    if (actA) p.out[b * p.out_batch_stride + f] = bs.x;
                    if (actB) p.out[b * p.out_batch_stride + f + 1] = bs.y;
                }
            }
        }
    }
}

}  // namespace b2s

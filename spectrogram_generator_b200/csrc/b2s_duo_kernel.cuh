// Frame-duo STFT -> PSD kernel, nperseg = 512, hop = S * 32 samples (S = 2, 4, 8: 87.5 %, 75 %
// and 50 % overlap): two CONSECUTIVE frames of a signal are transformed together, frame A in
// the low and frame B in the high half of packed fp32x2 registers.
//
// Why (ncu, profiles/r1_c2_v3*): at 75 % overlap the one-frame warp kernel is issue-bound
// (520 warp-instructions per frame, 360 of them FP32) with the L1/shared pipe at 73 %; the
// HBM budget is only ~66 SM-cycles per frame.  On sm_100a FADD2 / FMUL2 / FFMA2 do two fp32
// operations per lane per issue slot, take a scalar register or an immediate as a broadcast
// operand and have negate modifiers, so with (A, B) packed
//   * every arithmetic instruction of the transform serves two frames, and twiddles / window
//     taps enter as broadcast scalars (one LDS.64 serves both frames and both groups of a warp);
//   * every LDS.128 / STS.128 of the exchange moves one complex point of both frames;
//   * 16 lanes own a duo and hold 16 complex points per frame each: M = 256 = 16 x 16, so there
//     is ONE shared-memory exchange (the 16 x 16 transpose between the two radix-16 passes);
//   * after pass 1 lane q holds Z[q + 16 p]; the mirror bin Z[256 - k] the real-FFT split needs
//     sits in lane (16 - q) & 15 at p' = 15 - p: it comes over with 32 shuffles (half of the
//     points, each lane does the 8 pairs whose low bin it owns) instead of a second trip
//     through shared memory;
//   * consecutive frames overlap, so the raw samples of the duo are 16 + S register slots; the
//     next duo needs 2 S new LDG.64 per lane, issued one iteration ahead.
//
// Detrend (scipy _signaltools.py:4288-4290) is two fp32 passes as in the other kernels: a coarse
// per-frame mean is subtracted first (exact or nearly so when the recording sits on a large
// DC level), then the mean of the residual is removed inside the window multiply.  Every sum
// is formed in a fixed, frame-relative order, so the result of a frame depends on that frame's
// samples only (any chunking gives bit-identical output).
// The PSD scale is folded into the window taps (sqrt(scale/2), once per CTA).
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
// * (w.x + i w.y), the same twiddle for both frames: 2 FMUL2 + 2 FFMA2
B2S_HD cpx2 c2mul(cpx2 a, float2 w) {
    return cpx2{pk_fmas(a.re, w.x, pk_muls(a.im, -w.y)), pk_fmas(a.re, w.y, pk_muls(a.im, w.x))};
}
B2S_HD cpx2 c2mul_mi(cpx2 a) { return cpx2{a.im, pk_neg(a.re)}; }   // * (-i)
B2S_HD cpx2 c2mul_w8_1(cpx2 a) {
    return cpx2{pk_muls(pk_add(a.re, a.im), B2S_SQRT1_2), pk_muls(pk_sub(a.im, a.re), B2S_SQRT1_2)};
}
B2S_HD cpx2 c2mul_w8_3(cpx2 a) {
    return cpx2{pk_muls(pk_sub(a.im, a.re), B2S_SQRT1_2), pk_muls(pk_add(a.re, a.im), -B2S_SQRT1_2)};
}
B2S_HD void c2radix4(cpx2& a0, cpx2& a1, cpx2& a2, cpx2& a3) {
    const cpx2 t0 = c2add(a0, a2), t1 = c2sub(a0, a2);
    const cpx2 t2 = c2add(a1, a3), t3 = c2sub(a1, a3);
    a0 = c2add(t0, t2);
    a2 = c2sub(t0, t2);
    a1 = cpx2{pk_add(t1.re, t3.im), pk_sub(t1.im, t3.re)};
    a3 = cpx2{pk_sub(t1.re, t3.im), pk_add(t1.im, t3.re)};
}
// 16-point DFT in place, both frames; X[k] is left in v[perm16(k)] (same scheme as radix16)
B2S_HD void c2radix16(cpx2 (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) c2radix4(v[c], v[c + 4], v[c + 8], v[c + 12]);
    const float2 W1 = cmk(B2S_COS_PI_8, -B2S_SIN_PI_8);
    const float2 W3 = cmk(B2S_SIN_PI_8, -B2S_COS_PI_8);
    const float2 W9 = cmk(-B2S_COS_PI_8, B2S_SIN_PI_8);
    v[5] = c2mul(v[5], W1);   v[9] = c2mul_w8_1(v[9]);   v[13] = c2mul(v[13], W3);
    v[6] = c2mul_w8_1(v[6]);  v[10] = c2mul_mi(v[10]);   v[14] = c2mul_w8_3(v[14]);
    v[7] = c2mul(v[7], W3);   v[11] = c2mul_w8_3(v[11]); v[15] = c2mul(v[15], W9);
#pragma unroll
    for (int q = 0; q < 4; ++q) c2radix4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

struct DuoPlan {
    static constexpr int LOG2N = 9, N = 512, M = 256;
    static constexpr int G = 16;                         // lanes per frame duo
    static constexpr int NT = 128;                       // threads per CTA
    static constexpr int MINB = 3;                       // CTAs per SM (168 registers per thread)
    static constexpr int FPC = NT / G;                   // duos in flight per CTA
    static constexpr int ROW = 17;                       // float4 slots per lane row (16 + 1 pad)
    static constexpr int BUF = 16 * ROW;                 // exchange buffer of one duo, float4 units
    // shared memory: float2 tables, then the float4 exchange buffers
    static constexpr int OFF_WIN = 0;                    // [256]    window taps (w[2n], w[2n+1]) * sqrt(scale/2)
    static constexpr int OFF_TW1 = OFF_WIN + M;          // [15][16] W_256^(t q)
    static constexpr int OFF_TWP = OFF_TW1 + 15 * 16;    // [8][16]  W_512^(q + 16 p)
    static constexpr int TAB = OFF_TWP + 8 * 16;         // float2 units (a multiple of 2)
    static constexpr size_t SMEM = (size_t)TAB * sizeof(float2) + (size_t)FPC * BUF * sizeof(float4);
    static_assert(TAB % 2 == 0, "exchange buffers are 16-byte aligned");
};

// S = hop / 32: raw complex slots the frame start advances per frame
template <typename Tin, int S, int MODE>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(DuoPlan::NT, DuoPlan::MINB) stft_psd_duo_kernel(const StftParams p) {
    using DP = DuoPlan;
    using PL = Plan<9>;
    constexpr int M = DP::M, G = DP::G, ROW = DP::ROW;
    constexpr int NCUR = 16 + S;                         // raw complex slots of the duo
    constexpr int KEEP = (NCUR > 2 * S) ? NCUR - 2 * S : 0;   // slots shared with the next duo

    B2S_DYN_SMEM(smem_raw);
    float2* const smc = reinterpret_cast<float2*>(smem_raw);
    const int tid = (int)threadIdx.x;
    const int grp = tid / G;
    const int t = tid & (G - 1);
    const int lane0 = tid & 16;                          // first lane of this duo inside the warp
    float4* const buf = reinterpret_cast<float4*>(smem_raw + (size_t)DP::TAB * sizeof(float2)) + grp * DP::BUF;

    // ---- stage the constant tables (once per CTA); the PSD scale goes into the window ----
    {
        const float csc = sqrtf(0.5f * p.scale);
        const float2* w2 = reinterpret_cast<const float2*>(p.window);
        for (int i = tid; i < M; i += DP::NT) {
            const float2 w = __ldg(w2 + i);
            smc[DP::OFF_WIN + i] = cmk(w.x * csc, w.y * csc);
        }
        for (int i = tid; i < 15 * 16; i += DP::NT) smc[DP::OFF_TW1 + i] = __ldg(p.tw + PL::OFF_P1 + i);
        for (int i = tid; i < 8 * 16; i += DP::NT) smc[DP::OFF_TWP + i] = __ldg(p.tw + PL::OFF_POST + i);
    }
    __syncthreads();

    const int kout = p.kmax - p.kmin + 1;
    const int partner = lane0 | ((16 - t) & 15);
    const bool is0 = (t == 0);
    const float edge = is0 ? 0.5f : 1.0f;                // DC / Nyquist carry scale, not 2 scale

    const long long ustride = (long long)gridDim.x * DP::FPC;
    for (long long ub = (long long)blockIdx.x * DP::FPC + (grp & ~1); ub < p.n_units; ub += ustride) {
        long long u = ub + (grp & 1);
        const bool uvalid = u < p.n_units;
        if (!uvalid) u = p.n_units - 1;
        const long long b = u / p.units_per_signal;
        const int c = (int)(u - b * p.units_per_signal);
        const int f_begin = c * p.chunk_frames;
        const int f_end = (f_begin + p.chunk_frames < p.nframes) ? f_begin + p.chunk_frames : p.nframes;
        const Tin* const xb = reinterpret_cast<const Tin*>(p.x) + b * p.x_batch_stride + p.frame0 * (long long)p.hop;
        float* const ob = p.out + b * p.out_batch_stride - ((MODE == EPI_BAND) ? 0 : p.kmin);

        // raw samples of the duo: slot i <-> complex index t + 16 i relative to frame f
        float2 cur[NCUR];
        {
            const Tin* const xf = xb + (long long)f_begin * p.hop + 2 * t;
            const bool hasB = f_begin + 1 < f_end;
#pragma unroll
            for (int i = 0; i < NCUR; ++i)
                cur[i] = (i < 16 || hasB) ? Loader<Tin>::ld2(xf + 32 * i) : cmk(0.f, 0.f);
        }

        for (int f = f_begin;; f += 2) {
            const bool actA = uvalid && (f < f_end);
            const bool actB = uvalid && (f + 1 < f_end);
            if (!__any_sync(0xffffffffu, actA)) break;

            // ---- detrend + window, packing frame A (slots 0..15) and B (slots S..S+15) ----
            cpx2 v[16];
            if (p.detrend) {
                // pivots: the coarse mean of each frame, from per-slot sums added in a fixed
                // frame-relative order (blocks of S slots first: A and B share NB - 1 blocks)
                constexpr int NB = 16 / S;
                float blk[NB + 1];
#pragma unroll
                for (int bi = 0; bi <= NB; ++bi) {
                    float ss[S];
#pragma unroll
                    for (int i = 0; i < S; ++i) ss[i] = cur[bi * S + i].x + cur[bi * S + i].y;
#pragma unroll
                    for (int w = S / 2; w >= 1; w >>= 1)
#pragma unroll
                        for (int i = 0; i < w; ++i) ss[i] += ss[i + w];
                    blk[bi] = ss[0];
                }
                float cA = blk[0], cB = blk[1];
                if constexpr (NB == 2) {
                    cA += blk[1];
                    cB += blk[2];
                } else if constexpr (NB == 4) {
                    cA = (blk[0] + blk[1]) + (blk[2] + blk[3]);
                    cB = (blk[1] + blk[2]) + (blk[3] + blk[4]);
                } else {
                    cA = ((blk[0] + blk[1]) + (blk[2] + blk[3])) + ((blk[4] + blk[5]) + (blk[6] + blk[7]));
                    cB = ((blk[1] + blk[2]) + (blk[3] + blk[4])) + ((blk[5] + blk[6]) + (blk[7] + blk[8]));
                }
#pragma unroll
                for (int o = G / 2; o >= 1; o >>= 1) {
                    cA += __shfl_xor_sync(0xffffffffu, cA, o);
                    cB += __shfl_xor_sync(0xffffffffu, cB, o);
                }
                cA *= 1.0f / (float)DP::N;
                cB *= 1.0f / (float)DP::N;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i].re = cmk(cur[i].x - cA, cur[i + S].x - cB);
                    v[i].im = cmk(cur[i].y - cA, cur[i + S].y - cB);
                }
                float2 s[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) s[i] = pk_add(v[i].re, v[i].im);
#pragma unroll
                for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                    for (int i = 0; i < w; ++i) s[i] = pk_add(s[i], s[i + w]);
                float2 tot = s[0];
#pragma unroll
                for (int o = G / 2; o >= 1; o >>= 1)
                    tot = pk_add(tot, cmk(__shfl_xor_sync(0xffffffffu, tot.x, o), __shfl_xor_sync(0xffffffffu, tot.y, o)));
                const float2 nr = pk_muls(tot, -1.0f / (float)DP::N);     // - mean of the residual
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float2 w = smc[DP::OFF_WIN + t + 16 * i];
                    v[i].re = pk_fmas(v[i].re, w.x, pk_muls(nr, w.x));     // (x' - r) w, one rounding
                    v[i].im = pk_fmas(v[i].im, w.y, pk_muls(nr, w.y));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float2 w = smc[DP::OFF_WIN + t + 16 * i];
                    v[i].re = cmk(cur[i].x * w.x, cur[i + S].x * w.x);
                    v[i].im = cmk(cur[i].y * w.y, cur[i + S].y * w.y);
                }
            }

            // ---- next duo (frames f+2, f+3): keep the overlap, prefetch the 2 S new slots ----
            {
#pragma unroll
                for (int i = 0; i < KEEP; ++i) cur[i] = cur[i + 2 * S];
                const bool nextA = actA && (f + 2 < f_end);
                const bool nextB = actA && (f + 3 < f_end);
                const Tin* const xn = xb + (long long)(f + 2) * p.hop + 2 * t;
#pragma unroll
                for (int i = KEEP; i < NCUR; ++i) {
                    const bool need = (i < 16) ? nextA : nextB;
                    cur[i] = need ? Loader<Tin>::ld2(xn + 32 * i) : cmk(0.f, 0.f);
                }
            }

            // ---- pass 0: radix-16 over r (n = t + 16 r), then the 16 x 16 transpose ----
            c2radix16(v);
            __syncwarp();                        // the previous duo's pass-1 reads are done
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const cpx2 z = v[perm16(q)];
                buf[ROW * t + q] = make_float4(z.re.x, z.re.y, z.im.x, z.im.y);
            }
            __syncwarp();

            // ---- pass 1: lane q = t; twiddle W_256^(t' q), radix-16 over t' -> Z[q + 16 p] ----
#pragma unroll
            for (int tt = 0; tt < 16; ++tt) {
                const float4 q4 = buf[ROW * tt + t];
                v[tt] = cpx2{cmk(q4.x, q4.y), cmk(q4.z, q4.w)};
            }
#pragma unroll
            for (int tt = 1; tt < 16; ++tt) v[tt] = c2mul(v[tt], smc[DP::OFF_TW1 + (tt - 1) * 16 + t]);
            c2radix16(v);

            // ---- real-FFT split + PSD: this lane owns bins k = t + 16 pp (pp < 8) and 256 - k ----
            float* const rowA = ob + (long long)f * kout;
            float2 band = cmk(0.f, 0.f);
            auto put = [&](int k, float2 pw) {
                if constexpr (MODE == EPI_GENERAL) {
                    if (p.out_mode) pw = cmk(10.0f * log10f(fmaxf(pw.x, p.db_floor)), 10.0f * log10f(fmaxf(pw.y, p.db_floor)));
                    if (k >= p.kmin && k <= p.kmax) {
                        if (actA) rowA[k] = pw.x;
                        if (actB) rowA[k + kout] = pw.y;
                    }
                } else if constexpr (MODE == EPI_BAND) {
                    if (k >= p.kmin && k <= p.kmax) band = pk_add(band, pw);
                } else {
                    if (actA) rowA[k] = pw.x;
                    if (actB) rowA[k + kout] = pw.y;
                }
            };
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
                const cpx2 zk = v[perm16(pp)];
                const cpx2 snd = v[perm16(15 - pp)];
                cpx2 zm;
                zm.re = cmk(__shfl_sync(0xffffffffu, snd.re.x, partner), __shfl_sync(0xffffffffu, snd.re.y, partner));
                zm.im = cmk(__shfl_sync(0xffffffffu, snd.im.x, partner), __shfl_sync(0xffffffffu, snd.im.y, partner));
                {   // lane 0 pairs k = 16 pp with 16 (16 - pp), which it holds itself
                    const cpx2 own = v[perm16((16 - pp) & 15)];
                    zm.re = is0 ? own.re : zm.re;
                    zm.im = is0 ? own.im : zm.im;
                }
                const float2 w = smc[DP::OFF_TWP + pp * 16 + t];
                const cpx2 e{pk_add(zk.re, zm.re), pk_sub(zk.im, zm.im)};      // 2E = zk + conj(zm)
                const cpx2 o{pk_add(zk.im, zm.im), pk_sub(zm.re, zk.re)};      // 2O = -i (zk - conj(zm))
                const cpx2 tw = c2mul(o, w);
                const cpx2 a = c2add(e, tw), bq = c2sub(e, tw);                // 2 X[k], 2 conj(X[256 - k])
                float2 pk = pk_fma(a.re, a.re, pk_mul(a.im, a.im));
                float2 pm = pk_fma(bq.re, bq.re, pk_mul(bq.im, bq.im));
                if (pp == 0) {
                    pk = pk_muls(pk, edge);
                    pm = pk_muls(pm, edge);
                }
                const int k = t + 16 * pp;
                put(k, pk);
                put(M - k, pm);
            }
            {   // k = 128: X = conj(Z[128]), held by lane 0
                const cpx2 z = v[perm16(8)];
                const float2 pw = pk_muls(pk_fma(z.re, z.re, pk_mul(z.im, z.im)), 4.0f);
                if (is0) put(M / 2, pw);
            }
            if constexpr (MODE == EPI_BAND) {
#pragma unroll
                for (int o = G / 2; o >= 1; o >>= 1)
                    band = pk_add(band, cmk(__shfl_xor_sync(0xffffffffu, band.x, o), __shfl_xor_sync(0xffffffffu, band.y, o)));
                if (is0) {
                    if (actA) ob[f] = band.x;
                    if (actB) ob[f + 1] = band.y;
                }
            }
        }
    }
}

}  // namespace b2s

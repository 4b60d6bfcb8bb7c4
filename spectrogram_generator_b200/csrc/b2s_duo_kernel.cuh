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
    // shared memory, float4 units.  Every constant table is [j][lane] with the entries for the
    // two consecutive indices 2j, 2j+1 in one float4, so one LDS.128 fetches two of them and the
    // 16 lanes of a duo read 256 consecutive bytes (both duos of a warp read the same ones).
    static constexpr int OFF_WIN = 0;                    // [8][16] window taps of slots 2j, 2j+1, * sqrt(scale/2)
    static constexpr int OFF_TW1 = OFF_WIN + 8 * 16;     // [8][16] W_256^(t' q), t' = 2j, 2j+1
    static constexpr int OFF_TWP = OFF_TW1 + 8 * 16;     // [4][16] W_512^(q + 16 p), p = 2j, 2j+1
    static constexpr int TAB = OFF_TWP + 4 * 16;
    static constexpr size_t SMEM = (size_t)(TAB + FPC * BUF) * sizeof(float4);
};

// Coarse per-frame means of the duo (the detrend pivots), from per-slot sums added in a
// fixed frame-relative order: blocks of S slots first (frames A and B share 16/S - 1 blocks).
template <int S>
B2S_DEVICE void duo_coarse_means(const float2 (&cur)[16 + S], float& cA, float& cB) {
    if constexpr (S >= 14) {
        // frames A and B share (almost) nothing: the same fixed tree over each frame's own 16 slots
        float sa[16], sb[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            sa[i] = cur[i].x + cur[i].y;
            sb[i] = cur[i + S].x + cur[i + S].y;
        }
#pragma unroll
        for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
            for (int i = 0; i < w; ++i) {
                sa[i] += sa[i + w];
                sb[i] += sb[i + w];
            }
        cA = sa[0];
        cB = sb[0];
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            cA += __shfl_xor_sync(0xffffffffu, cA, o);
            cB += __shfl_xor_sync(0xffffffffu, cB, o);
        }
        cA *= 1.0f / 512.0f;
        cB *= 1.0f / 512.0f;
        return;
    }
    constexpr int NB = (S >= 14) ? 1 : 16 / S;
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
    if constexpr (NB == 2) {
        cA = blk[0] + blk[1];
        cB = blk[1] + blk[2];
    } else if constexpr (NB == 4) {
        cA = (blk[0] + blk[1]) + (blk[2] + blk[3]);
        cB = (blk[1] + blk[2]) + (blk[3] + blk[4]);
    } else {
        cA = ((blk[0] + blk[1]) + (blk[2] + blk[3])) + ((blk[4] + blk[5]) + (blk[6] + blk[7]));
        cB = ((blk[1] + blk[2]) + (blk[3] + blk[4])) + ((blk[5] + blk[6]) + (blk[7] + blk[8]));
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
        cA += __shfl_xor_sync(0xffffffffu, cA, o);
        cB += __shfl_xor_sync(0xffffffffu, cB, o);
    }
    cA *= 1.0f / 512.0f;
    cB *= 1.0f / 512.0f;
}

// S = hop / 32: raw complex slots the frame start advances per frame
template <typename Tin, int S, int MODE, int MINB = DuoPlan::MINB, int OPT = 0>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(DuoPlan::NT, MINB) stft_psd_duo_kernel(const StftParams p) {
    using DP = DuoPlan;
    // timing ablations (tools/ubench/duo_bench only; results are wrong with any of these set)
    constexpr bool ABL_NOSTG = (OPT & 8) != 0, ABL_NOXCHG = (OPT & 16) != 0, ABL_NOSHFL = (OPT & 32) != 0;
    using PL = Plan<9>;
    constexpr int M = DP::M, G = DP::G, ROW = DP::ROW;
    constexpr int NCUR = 16 + S;                         // raw complex slots of the duo
    constexpr int KEEP = (NCUR > 2 * S) ? NCUR - 2 * S : 0;   // slots shared with the next duo
    // S == 16 keeps both frames of a duo whole, so it serves ANY hop (up to nperseg): frame B's slots
    // then start hop samples after frame A's instead of exactly 16 slots (512 samples) after
    const long long bskew = (S == 16) ? (long long)p.hop - 16 * 32 : 0;

    B2S_DYN_SMEM_F4(sm4);
    const int tid = (int)threadIdx.x;
    const int grp = tid / G;
    const int t = tid & (G - 1);
    float4* const buf = sm4 + DP::TAB + grp * DP::BUF;

    // ---- stage the constant tables (once per CTA); the PSD scale goes into the window ----
    {
        const float csc = sqrtf(0.5f * p.scale);
        const float2* w2 = reinterpret_cast<const float2*>(p.window);
        for (int i = tid; i < 8 * 16; i += DP::NT) {
            const int j = i >> 4, l = i & 15;
            const float2 wa = __ldg(w2 + l + 16 * (2 * j)), wb = __ldg(w2 + l + 16 * (2 * j + 1));
            sm4[DP::OFF_WIN + i] = make_float4(wa.x * csc, wa.y * csc, wb.x * csc, wb.y * csc);
            // W_256^(t' q): t' = 0 is 1
            const float2 ta = (j == 0) ? cmk(1.f, 0.f) : __ldg(p.tw + PL::OFF_P1 + (2 * j - 1) * 16 + l);
            const float2 tb = __ldg(p.tw + PL::OFF_P1 + (2 * j) * 16 + l);
            sm4[DP::OFF_TW1 + i] = make_float4(ta.x, ta.y, tb.x, tb.y);
            if (j < 4) {
                const float2 pa = __ldg(p.tw + PL::OFF_POST + l + 16 * (2 * j));
                const float2 pb = __ldg(p.tw + PL::OFF_POST + l + 16 * (2 * j + 1));
                sm4[DP::OFF_TWP + i] = make_float4(pa.x, pa.y, pb.x, pb.y);
            }
        }
    }
    __syncthreads();

    const int kout = p.kmax - p.kmin + 1;
    const int partner = (tid & 16) | ((16 - t) & 15);
    const bool is0 = (t == 0);
    const float edge = is0 ? 0.5f : 1.0f;                // DC / Nyquist carry scale, not 2 scale

    // work units: static round-robin over the grid, or (p.work) an atomic counter the warps draw
    // 2 units at a time from -- the next draw is issued a whole unit ahead, so its latency is hidden
    const long long ustride = (long long)gridDim.x * DP::FPC;
    const bool dyn = p.work != nullptr;
    auto draw = [&]() -> long long {
        int b0 = 0;
        if ((tid & 31) == 0) b0 = atomicAdd(p.work, 2);
        return (long long)__shfl_sync(0xffffffffu, b0, 0);
    };
    long long ub_next = dyn ? draw() : (long long)blockIdx.x * DP::FPC + (grp & ~1);
    while (ub_next < p.n_units) {
        const long long ub = ub_next;
        ub_next = dyn ? draw() : ub + ustride;
        long long u = ub + (grp & 1);
        const bool uvalid = u < p.n_units;
        if (!uvalid) u = p.n_units - 1;
        const long long b = u / p.units_per_signal;
        const int c = (int)(u - b * p.units_per_signal);
        const int f_begin = c * p.chunk_frames;
        const int f_end = (f_begin + p.chunk_frames < p.nframes) ? f_begin + p.chunk_frames : p.nframes;
        const Tin* const xb = reinterpret_cast<const Tin*>(p.x) + b * p.x_batch_stride + p.frame0 * (long long)p.hop;
        float* const ob = p.out + b * p.out_batch_stride - ((MODE == EPI_BAND) ? 0 : p.kmin);

        // raw samples of the duo: slot i <-> complex index t + 16 i relative to frame f.
        // Slots 16.. belong to frame B only; when the run has no frame B they re-read the
        // previous S slots instead (valid addresses, values unused).
        float2 cur[NCUR];
        {
            const Tin* const xf = xb + (long long)f_begin * p.hop + 2 * t;
            const Tin* const xfB = xf + bskew - ((f_begin + 1 < f_end) ? 0 : p.hop);
#pragma unroll
            for (int i = 0; i < NCUR; ++i) cur[i] = Loader<Tin>::ld2(((i < 16) ? xf : xfB) + 32 * i);
        }
        // the two duos of a warp run the same trip count (the longer run's); a duo past its
        // last frame recomputes with its stores predicated off
        int ntrip = uvalid ? (f_end - f_begin + 1) >> 1 : 0;
        {
            const int other = __shfl_xor_sync(0xffffffffu, ntrip, 16);
            ntrip = ntrip > other ? ntrip : other;
        }
        int f = f_begin;
        for (int it = 0; it < ntrip; ++it, f += 2) {
            const bool actA = uvalid && (f < f_end);
            const bool actB = uvalid && (f + 1 < f_end);

            // ---- detrend + window, packing frame A (slots 0..15) and B (slots S..S+15) ----
            // x' = x - (coarse mean) is exact or nearly so whatever the DC level; the mean r of
            // x' is then removed inside the window multiply with a single rounding.
            // (Measured and lost on B200: removing r after pass 0 through the transformed taps,
            //  computing the next duo's pivots one iteration ahead, and carrying per-hop-block
            //  means / residual sums across duos so that no reduction precedes the transform
            //  (fewer operations, no shuffle chain in front) -- all lengthen live ranges at the
            //  168-register cap; 2-7 % slower.)
            cpx2 v[16];
            if (p.detrend) {
                float cA, cB;
                duo_coarse_means<S>(cur, cA, cB);
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
                const float2 nr = pk_muls(tot, -1.0f / (float)DP::N);      // - mean of x'
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 w = sm4[DP::OFF_WIN + j * 16 + t];
                    v[2 * j].re = pk_fmas(v[2 * j].re, w.x, pk_muls(nr, w.x));         // (x' - r) w, one rounding
                    v[2 * j].im = pk_fmas(v[2 * j].im, w.y, pk_muls(nr, w.y));
                    v[2 * j + 1].re = pk_fmas(v[2 * j + 1].re, w.z, pk_muls(nr, w.z));
                    v[2 * j + 1].im = pk_fmas(v[2 * j + 1].im, w.w, pk_muls(nr, w.w));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 w = sm4[DP::OFF_WIN + j * 16 + t];
                    v[2 * j].re = cmk(cur[2 * j].x * w.x, cur[2 * j + S].x * w.x);
                    v[2 * j].im = cmk(cur[2 * j].y * w.y, cur[2 * j + S].y * w.y);
                    v[2 * j + 1].re = cmk(cur[2 * j + 1].x * w.z, cur[2 * j + 1 + S].x * w.z);
                    v[2 * j + 1].im = cmk(cur[2 * j + 1].y * w.w, cur[2 * j + 1 + S].y * w.w);
                }
            }

            // ---- next duo (frames f+2, f+3): keep the overlap, prefetch the 2 S new slots ----
            // (unpredicated loads: past the end of the run the addresses are clamped to the
            //  run's last frame, the values are never used)
            {
#pragma unroll
                for (int i = 0; i < KEEP; ++i) cur[i] = cur[i + 2 * S];
                const int fa = (f + 2 < f_end) ? f + 2 : f_end - 1;
                const Tin* const xn = xb + (long long)fa * p.hop + 2 * t;
                const Tin* const xnB = xn + bskew - ((fa + 1 < f_end) ? 0 : p.hop);
#pragma unroll
                for (int i = KEEP; i < NCUR; ++i) cur[i] = Loader<Tin>::ld2(((i < 16) ? xn : xnB) + 32 * i);
            }

            // ---- pass 0: radix-16 over r (n = t + 16 r), then the 16 x 16 transpose ----
            c2radix16(v);
            if constexpr (!ABL_NOXCHG) {
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
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 w = sm4[DP::OFF_TW1 + j * 16 + t];
                if (j > 0) v[2 * j] = c2mul(v[2 * j], cmk(w.x, w.y));
                v[2 * j + 1] = c2mul(v[2 * j + 1], cmk(w.z, w.w));
            }
            c2radix16(v);
            __syncwarp();                        // every lane has consumed its exchange reads

            // ---- real-FFT split + PSD: this lane owns bins k = t + 16 pp (pp < 8) and 256 - k ----
            float* const rowA = ob + (long long)f * kout;
            float* const pA = rowA + t;                    // bin k = t + 16 pp of frame A
            float* const pmA = rowA + (M - t);             // bin 256 - k
            const int koutc = (MODE == EPI_PLAIN) ? (M + 1) : kout;
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
                    if (actB) rowA[k + koutc] = pw.y;
                }
            };
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
                const cpx2 zk = v[perm16(pp)];
                // the mirror Z[256 - k] sits in the partner lane at p' = 15 - pp; lane 0 pairs
                // k = 16 pp with 16 (16 - pp), which it holds itself (it is its own partner)
                const cpx2 s15 = v[perm16(15 - pp)], s16 = v[perm16((16 - pp) & 15)];
                const float s0 = is0 ? s16.re.x : s15.re.x, s1 = is0 ? s16.re.y : s15.re.y;
                const float s2 = is0 ? s16.im.x : s15.im.x, s3 = is0 ? s16.im.y : s15.im.y;
                cpx2 zm;
                if constexpr (ABL_NOSHFL) {
                    zm.re = cmk(s0, s1);
                    zm.im = cmk(s2, s3);
                } else {
                    zm.re = cmk(__shfl_sync(0xffffffffu, s0, partner), __shfl_sync(0xffffffffu, s1, partner));
                    zm.im = cmk(__shfl_sync(0xffffffffu, s2, partner), __shfl_sync(0xffffffffu, s3, partner));
                }
                const float4 w4 = sm4[DP::OFF_TWP + (pp >> 1) * 16 + t];
                const float2 w = (pp & 1) ? cmk(w4.z, w4.w) : cmk(w4.x, w4.y);
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
                if constexpr (ABL_NOSTG) {
                    band = pk_add(band, pk_add(pk, pm));
                    if (pp == 7 && band.x == 123.456f) pA[0] = band.y;
                } else if constexpr (MODE == EPI_PLAIN) {
                    if (actA) {
                        pA[16 * pp] = pk.x;
                        pmA[-16 * pp] = pm.x;
                    }
                    if (actB) {
                        pA[koutc + 16 * pp] = pk.y;
                        pmA[koutc - 16 * pp] = pm.y;
                    }
                } else {
                    const int k = t + 16 * pp;
                    put(k, pk);
                    put(M - k, pm);
                }
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

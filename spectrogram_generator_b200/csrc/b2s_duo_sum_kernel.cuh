// Frame-duo STFT -> PSD kernel with the cross-sweep sum fused in (nperseg 512, hop 64 / 128 / 256,
// all bins, linear power): the arithmetic of b2s_duo_kernel.cuh, walked in the other direction.
//
// stft_psd_duo_kernel keeps a lane group on one signal and slides along its frames; the mean
// spectrogram of a batch of sweeps (BASELINE config 2) then needs a second
// kernel that reads every per-sweep spectrogram back (318 MB for C2: 30 % of the step).  Here a
// lane group keeps ONE frame duo (f, f + 1) and walks over a block of consecutive sweeps: each
// per-sweep row is stored exactly as before (bit-identical: the per-frame arithmetic is the
// same), and the 2 x 257 power values are also added, in sweep order, into running sums that
// never leave the SM: 34 values per lane, kept in tensor memory (ACC_TMEM = 1, the product
// path) or in shared memory (the CPU emulator's path and the twin the launch takes its
// residency from).  A unit ends by writing its block's partial sums; a fold over the sweep
// blocks (batch_sum_kernel, a few MB) finishes the sum.  The order of the additions is fixed
// (sweep order inside a block, block order in the fold): deterministic.  The price is the
// sliding register window -- every duo loads its 16 + S slots afresh, one sweep ahead --
// against reading the whole output a second time.
#pragma once

#include "b2s_duo_kernel.cuh"
#include "b2s_tmem.cuh"

namespace b2s {



struct DuoSumPlan {
    static constexpr int ACC = 9 * DuoPlan::G;                          // float4 per duo
    static constexpr size_t SMEM = DuoPlan::SMEM + (size_t)DuoPlan::FPC * ACC * sizeof(float4);
    static constexpr int TMEM_COLS = 64;                                // 34 used: 8 x (k: A, B; 256 - k: A, B) + bin 128
};

// units: (sweep block, duo) with the duo index fastest, so that the groups working at the same
// time are on neighbouring frames of the same sweeps (their overlapping samples meet in L1 / L2).
// p.units_per_signal = duos per block rounded up to even, p.n_units = blocks * that
// (plan_stft_sum); p.acc_rows sweeps per block, p.acc_batch sweeps in all.
template <typename Tin, int S, int ACC_TMEM = 0>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(DuoPlan::NT, DuoPlan::MINB) stft_psd_duo_sum_kernel(const StftParams p) {
    using DP = DuoPlan;
    using PL = Plan<9>;
    constexpr int M = DP::M, G = DP::G, ROW = DP::ROW;
    constexpr int NCUR = 16 + S;
    constexpr int KOUT = M + 1;

    B2S_DYN_SMEM_F4(sm4);
    const int tid = (int)threadIdx.x;
    const int grp = tid / G;
    const int t = tid & (G - 1);
    // exchange buffer of the duo: a plane of real and a plane of imaginary parts, [16][ROW] float2 (A, B) each
    float2* const xre = reinterpret_cast<float2*>(sm4 + DP::TAB + grp * DP::BUF);
    float2* const xim = xre + 16 * ROW;
    // the block sums of the duo: [slot][lane] float4 = (bin k: A, B; bin 256 - k: A, B), slot 8 = bin 128
    float4* const sacc = sm4 + DP::TAB + DP::FPC * DP::BUF + grp * DuoSumPlan::ACC + t;

    // ---- stage the constant tables (once per CTA); the PSD scale goes into the window ----
    {
        const float csc = sqrtf(0.5f * p.scale);
        const float2* w2 = reinterpret_cast<const float2*>(p.window);
        for (int i = tid; i < 8 * 16; i += DP::NT) {
            const int j = i >> 4, l = i & 15;
            const float2 wa = __ldg(w2 + l + 16 * (2 * j)), wb = __ldg(w2 + l + 16 * (2 * j + 1));
            sm4[DP::OFF_WIN + i] = make_float4(wa.x * csc, wa.y * csc, wb.x * csc, wb.y * csc);
            const float2 ta = (j == 0) ? cmk(1.f, 0.f) : __ldg(p.tw + PL::OFF_P1 + (2 * j - 1) * 16 + l);
            const float2 tb = __ldg(p.tw + PL::OFF_P1 + (2 * j) * 16 + l);
            sm4[DP::OFF_TW1 + i] = make_float4(ta.x, ta.y, tb.x, tb.y);
            // split twiddles as [pp][lane] float2: one conflict-free 64-bit read per pair (the
            // [pp / 2][lane] float4 rows of the per-sweep kernel were re-read here half by half,
            // 16 bytes apart: two-way bank conflicts, +5 % shared-memory wavefronts)
            reinterpret_cast<float2*>(sm4 + DP::OFF_TWP)[i] = __ldg(p.tw + PL::OFF_POST + l + 16 * j);
        }
    }
#ifndef B2S_EMU
    unsigned tacc = 0;
    if constexpr (ACC_TMEM) {
        __shared__ unsigned tmem_base_s;
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         : : "r"((unsigned)__cvta_generic_to_shared(&tmem_base_s)), "n"(DuoSumPlan::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" : : : "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" : : : "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" : : : "memory");
        tacc = tmem_base_s + ((unsigned)((tid >> 5) & 3) << 21);        // lane field: 32 (warp % 4) << 16
    } else {
        __syncthreads();
    }
#else
    __syncthreads();
#endif

    const int partner = (tid & 16) | ((16 - t) & 15);
    const bool is0 = (t == 0);
    const float edge = is0 ? 0.5f : 1.0f;
    const int nduos = (p.nframes + 1) >> 1;
    const int batch = p.acc_batch;

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
        // both duos of a warp belong to the same sweep block (units_per_signal is even; the odd
        // slot past the last duo of a block stays idle), so the trip count is warp-uniform
        const long long u = ub + (grp & 1);
        const int blk = (int)(u / p.units_per_signal);
        int c = (int)(u - (long long)blk * p.units_per_signal);
        const bool uvalid = c < nduos;
        if (!uvalid) c = nduos - 1;
        const int f = 2 * c;                                           // frames f, f + 1
        const int b_begin = blk * p.acc_rows;
        const int ntrip = ((b_begin + p.acc_rows < batch) ? p.acc_rows : batch - b_begin);
        const bool hasB = f + 1 < p.nframes;
        const bool actA = uvalid, actB = uvalid && hasB;
        const long long offB = hasB ? 0 : -(long long)p.hop;         // no frame B: re-read valid samples
        const Tin* xn = reinterpret_cast<const Tin*>(p.x) + (p.frame0 + f) * (long long)p.hop + 2 * t +
                        (long long)b_begin * p.x_batch_stride;
        float* rowA = p.out + (long long)b_begin * p.out_batch_stride + (long long)f * KOUT;

        // raw samples of the duo of the first sweep of the block
        float2 cur[NCUR];
#pragma unroll
        for (int i = 0; i < NCUR; ++i) cur[i] = Loader<Tin>::ld2(xn + ((i < 16) ? 0 : offB) + 32 * i);
        // the block sums start at zero; only this lane touches its slots, no barrier needed
#ifndef B2S_EMU
        if constexpr (ACC_TMEM) {
#pragma unroll
            for (int j = 0; j < 8; ++j) tm_st4(tacc + 4 * j, 0.f, 0.f, 0.f, 0.f);
            tm_st2(tacc + 32, 0.f, 0.f);
        } else
#endif
        {
#pragma unroll
            for (int j = 0; j < 9; ++j) sacc[j * G] = make_float4(0.f, 0.f, 0.f, 0.f);
        }

        for (int it = 0; it < ntrip; ++it) {

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


            // ---- the same duo of the next sweep: all 16 + S slots, one sweep ahead ----
            // (Measured and dropped, round 2: loading the sweep when it is needed instead -- no 16 + S slots live
            //  across the transform, 128 registers, FOUR CTAs per SM -- with nothing, an L2 or an L1 prefetch of
            //  the next sweep's lines in their place: 0.171-0.175 ms against 0.169.  The kernel sits at the same
            //  time with three or four warps per scheduler; profiles/r2_c2_duo_sum_full.ncu_summary.txt.)
            if (it + 1 < ntrip) xn += p.x_batch_stride;
#pragma unroll
            for (int i = 0; i < NCUR; ++i) cur[i] = Loader<Tin>::ld2(xn + ((i < 16) ? 0 : offB) + 32 * i);

            // ---- pass 0: radix-16 over r (n = t + 16 r), then the 16 x 16 transpose ----
            c2radix16(v);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const cpx2 z = v[perm16(q)];
                xre[ROW * t + q] = z.re;           // two conflict-free 64-bit stores straight from the register
                xim[ROW * t + q] = z.im;           // pairs: a 128-bit store first gathers them with 4 MOVs.
                                                   // (1-2 % here; the per-sweep kernel loses 4 % with it)
            }
            __syncwarp();

            // ---- pass 1: lane q = t; twiddle W_256^(t' q), radix-16 over t' -> Z[q + 16 p] ----
#pragma unroll
            for (int tt = 0; tt < 16; ++tt) {
                v[tt] = cpx2{xre[ROW * tt + t], xim[ROW * tt + t]};
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 w = sm4[DP::OFF_TW1 + j * 16 + t];
                if (j > 0) v[2 * j] = c2mul(v[2 * j], cmk(w.x, w.y));
                v[2 * j + 1] = c2mul(v[2 * j + 1], cmk(w.z, w.w));
            }
            c2radix16(v);
            __syncwarp();                        // every lane has consumed its exchange reads


            // ---- real-FFT split + PSD: store the sweep's rows, add to the block sums ----
            float* const pA = rowA + t;                    // bin k = t + 16 pp of frame A
            float* const pmA = rowA + (M - t);             // bin 256 - k
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
                const cpx2 zk = v[perm16(pp)];
                float4 a4;
#ifndef B2S_EMU
                if constexpr (ACC_TMEM) {
                    if (pp == 0) tm_st_wait();                         // the previous sweep's updates have landed
                    tm_ld4(tacc + 4 * pp, a4.x, a4.y, a4.z, a4.w);     // in flight during the butterfly below
                }
#endif
                // the mirror Z[256 - k] sits in the partner lane at p' = 15 - pp; lane 0 pairs
                // k = 16 pp with 16 (16 - pp), which it holds itself (it is its own partner)
                const cpx2 s15 = v[perm16(15 - pp)], s16 = v[perm16((16 - pp) & 15)];
                const float s0 = is0 ? s16.re.x : s15.re.x, s1 = is0 ? s16.re.y : s15.re.y;
                const float s2 = is0 ? s16.im.x : s15.im.x, s3 = is0 ? s16.im.y : s15.im.y;
                cpx2 zm;
                zm.re = cmk(__shfl_sync(0xffffffffu, s0, partner), __shfl_sync(0xffffffffu, s1, partner));
                zm.im = cmk(__shfl_sync(0xffffffffu, s2, partner), __shfl_sync(0xffffffffu, s3, partner));
                const float2 w = reinterpret_cast<const float2*>(sm4 + DP::OFF_TWP)[pp * 16 + t];
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
                if (actA) {
                    pA[16 * pp] = pk.x;
                    pmA[-16 * pp] = pm.x;
                }
                if (actB) {
                    pA[KOUT + 16 * pp] = pk.y;
                    pmA[KOUT - 16 * pp] = pm.y;
                }
                // (34 more live registers do not fit beside the transform: measured with spills 2 % slower)
#ifndef B2S_EMU
                if constexpr (ACC_TMEM) {
                    tm_ld_wait4(a4.x, a4.y, a4.z, a4.w);
                    const float2 a0 = pk_add(cmk(a4.x, a4.y), pk), a1 = pk_add(cmk(a4.z, a4.w), pm);
                    tm_st4(tacc + 4 * pp, a0.x, a0.y, a1.x, a1.y);
                } else
#endif
                {
                    a4 = sacc[pp * G];
                    const float2 a0 = pk_add(cmk(a4.x, a4.y), pk), a1 = pk_add(cmk(a4.z, a4.w), pm);
                    sacc[pp * G] = make_float4(a0.x, a0.y, a1.x, a1.y);
                }
            }
            {   // k = 128: X = conj(Z[128]), held by lane 0
                const cpx2 z = v[perm16(8)];
                const float2 pw = pk_muls(pk_fma(z.re, z.re, pk_mul(z.im, z.im)), 4.0f);
                if (is0) {
                    if (actA) rowA[M / 2] = pw.x;
                    if (actB) rowA[M / 2 + KOUT] = pw.y;
                }
#ifndef B2S_EMU
                if constexpr (ACC_TMEM) {
                    float2 am;
                    tm_ld2(tacc + 32, am.x, am.y);
                    tm_ld_wait2(am.x, am.y);
                    am = pk_add(am, pw);
                    tm_st2(tacc + 32, am.x, am.y);
                } else
#endif
                if (is0) {
                    float2* const q = reinterpret_cast<float2*>(sacc + 8 * G);
                    *q = pk_add(*q, pw);
                }
            }
            rowA += p.out_batch_stride;
        }

        // ---- the block's partial sums: p.acc[blk][frame][bin] ----
        float4 a9[9];
#ifndef B2S_EMU
        if constexpr (ACC_TMEM) {        // (straight from tensor memory: this kernel may be launched without the
            tm_st_wait();                //  twin's shared-memory sums)
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
                tm_ld4(tacc + 4 * pp, a9[pp].x, a9[pp].y, a9[pp].z, a9[pp].w);
                tm_ld_wait4(a9[pp].x, a9[pp].y, a9[pp].z, a9[pp].w);
            }
            a9[8] = make_float4(0.f, 0.f, 0.f, 0.f);
            tm_ld2(tacc + 32, a9[8].x, a9[8].y);
            tm_ld_wait2(a9[8].x, a9[8].y);
        } else
#endif
        {
#pragma unroll
            for (int pp = 0; pp < 9; ++pp) a9[pp] = sacc[pp * G];
        }
        if (uvalid) {
            float* const sA = p.acc + ((long long)blk * p.nframes + f) * KOUT;
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
                const float4 a4 = a9[pp];
                sA[t + 16 * pp] = a4.x;
                sA[M - t - 16 * pp] = a4.z;
                if (hasB) {
                    sA[KOUT + t + 16 * pp] = a4.y;
                    sA[KOUT + M - t - 16 * pp] = a4.w;
                }
            }
            if (is0) {
                sA[M / 2] = a9[8].x;
                if (hasB) sA[KOUT + M / 2] = a9[8].y;
            }
        }
    }
#ifndef B2S_EMU
    if constexpr (ACC_TMEM) {
        asm volatile("tcgen05.fence::before_thread_sync;" : : : "memory");
        __syncthreads();
        if (tid < 32)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                         : : "r"(tacc & 0xffffu), "n"(DuoSumPlan::TMEM_COLS) : "memory");
    }
#endif
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

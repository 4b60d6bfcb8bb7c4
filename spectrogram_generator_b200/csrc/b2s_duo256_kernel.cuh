// Frame-duo STFT -> PSD kernel, nperseg = 256, hop = S * 16 samples (S = 2, 4, 8, 14, 16): the scheme of
// b2s_duo_kernel.cuh (two consecutive frames packed in fp32x2 registers, sliding register
// window, one shared-memory transpose, shuffle mirror exchange) for M = 128 = 16 x 8:
//   * 8 lanes own a duo (4 duos per warp), 16 complex points per frame per lane;
//   * pass 0 is the radix-16 over the lane's own points; after the 8 x 16 transpose lane q holds
//     the columns q and q + 8 and does two 8-point DFTs -> Z[q + 16 p] and Z[q + 8 + 16 p];
//   * the mirror of Z[q + 16 p] is Z[(8 - q) + 8 + 16 (7 - p)]: the second column of lane
//     (8 - q) & 7, which comes over with 32 shuffles; each lane does the 8 pairs whose low bin
//     is in its first column.  Lane 4 is its own partner and needs nothing special; lane 0 pairs
//     inside its own columns (16 p with 16 (8 - p), 8 + 16 p with 8 + 16 (7 - p)) and selects
//     its operands, twiddles and bins accordingly.
//
// SUM mode (stft_psd_duo256_sum_kernel: per-sweep spectrograms AND their cross-sweep sum in one pass,
// SURVEY.md 8 a-15): as in b2s_duo_sum_kernel.cuh a lane group keeps ONE frame duo and walks over a
// block of consecutive sweeps, loading all its 16 + S slots afresh one sweep ahead; rows bit-identical
// to the per-sweep kernel's, the 2 x 129 power values added in sweep order into 34 running sums per
// lane -- tensor memory (SUM = 2, the product path) or shared memory (SUM = 1: emulator / residency twin).
#pragma once

#include "b2s_duo_cta_kernel.cuh"
#include "b2s_tmem.cuh"

namespace b2s {

struct Duo256Plan {
    static constexpr int N = 256, M = 128;
    static constexpr int G = 8;                          // lanes per frame duo
    static constexpr int NT = 128;
    static constexpr int MINB = 3;
    static constexpr int FPC = NT / G;                   // duos in flight per CTA
    static constexpr int ROW = 17;
    static constexpr int BUF = 8 * ROW;                  // float4 slots per duo
    // shared memory (float4 units)
    static constexpr int OFF_WIN = 0;                    // [8][8]  window taps of slots 2j, 2j+1, * sqrt(scale/2)
    static constexpr int OFF_TW = OFF_WIN + 8 * 8;       // [8][8]  (W_128^(t q), W_128^(t (q + 8))) for row t, lane q
    static constexpr int OFF_TWP = OFF_TW + 8 * 8;       // [4][8]  split twiddles of pairs 2j, 2j+1 (lane 0: its own bins)
    static constexpr int TAB = OFF_TWP + 4 * 8;
    static constexpr size_t SMEM = (size_t)(TAB + FPC * BUF) * sizeof(float4);
    // SUM mode: the shared-memory twin's running sums, [9][NT] float4 behind the transpose buffers
    static constexpr int ACC_SLOTS = 9;                  // 8 pairs x (k: A, B; 128 - k: A, B) + bin 64
    static constexpr size_t SUM_SMEM = SMEM + (size_t)ACC_SLOTS * NT * sizeof(float4);
    static constexpr int TMEM_COLS = 64;                 // 34 used
};

// low bin of pair pp for lane q: q + 16 pp, except lane 0 (16 pp for pp < 4, then 8 + 16 (pp - 4))
B2S_HD int duo256_low_bin(int q, int pp) { return (q != 0) ? q + 16 * pp : ((pp < 4) ? 16 * pp : 8 + 16 * (pp - 4)); }

template <typename Tin, int S, int MODE, int SUM>
B2S_DEVICE void stft_psd_duo256_body(const StftParams& p);

template <typename Tin, int S, int MODE>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(Duo256Plan::NT, Duo256Plan::MINB) stft_psd_duo256_kernel(const StftParams p) {
    stft_psd_duo256_body<Tin, S, MODE, 0>(p);
}
// per-sweep rows + cross-sweep block sums (SUM = 1: sums in shared memory, 2: in tensor memory); p.units_per_signal =
// duos per block rounded up to a multiple of 4 (the four duos of a warp share a block), plan_stft_sum(..., 4)
template <typename Tin, int S, int SUM>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(Duo256Plan::NT, Duo256Plan::MINB) stft_psd_duo256_sum_kernel(const StftParams p) {
    stft_psd_duo256_body<Tin, S, EPI_PLAIN, SUM>(p);
}

template <typename Tin, int S, int MODE, int SUM>
B2S_DEVICE void stft_psd_duo256_body(const StftParams& p) {
    static_assert(SUM == 0 || MODE == EPI_PLAIN, "SUM mode: plain epilogue");
    using DP = Duo256Plan;
    using PL = Plan<8>;
    constexpr int M = DP::M, G = DP::G, ROW = DP::ROW;
    constexpr int NCUR = 16 + S;
    constexpr int KEEP = (NCUR > 2 * S) ? NCUR - 2 * S : 0;
    // S == 16 keeps both frames of a duo whole, so it serves ANY hop (up to nperseg): frame B's slots
    // then start hop samples after frame A's instead of exactly 16 slots (256 samples) after
    const long long bskew = (S == 16) ? (long long)p.hop - 16 * 16 : 0;
    static_assert(PL::NS == 16 && PL::GF == 8, "plan tables: FIN = W_128^(r kappa)");

    B2S_DYN_SMEM_F4(sm4);
    const int tid = (int)threadIdx.x;
    const int grp = tid / G;
    const int t = tid & (G - 1);
    float4* const buf = sm4 + DP::TAB + grp * DP::BUF;

    // ---- constant tables, once per CTA; the PSD scale goes into the window ----
    {
        const float csc = sqrtf(0.5f * p.scale);
        const float2* w2 = reinterpret_cast<const float2*>(p.window);
        if (tid < 64) {
            const int jj = tid >> 3, l = tid & 7;
            const float2 wa = __ldg(w2 + l + 8 * (2 * jj)), wb = __ldg(w2 + l + 8 * (2 * jj + 1));
            sm4[DP::OFF_WIN + tid] = make_float4(wa.x * csc, wa.y * csc, wb.x * csc, wb.y * csc);
            // row t = jj of the transpose: W_128^(t q) and W_128^(t (q + 8)); row 0 is 1
            const float2 ta = (jj == 0) ? cmk(1.f, 0.f) : __ldg(p.tw + PL::OFF_FIN + (jj - 1) * 16 + l);
            const float2 tb = (jj == 0) ? cmk(1.f, 0.f) : __ldg(p.tw + PL::OFF_FIN + (jj - 1) * 16 + l + 8);
            sm4[DP::OFF_TW + tid] = make_float4(ta.x, ta.y, tb.x, tb.y);
            if (jj < 4) {
                const float2 pa = __ldg(p.tw + PL::OFF_POST + duo256_low_bin(l, 2 * jj));
                const float2 pb = __ldg(p.tw + PL::OFF_POST + duo256_low_bin(l, 2 * jj + 1));
                sm4[DP::OFF_TWP + tid] = make_float4(pa.x, pa.y, pb.x, pb.y);
            }
        }
    }
    [[maybe_unused]] float4* const sacc = sm4 + DP::TAB + DP::FPC * DP::BUF + tid;     // [slot][NT]
    [[maybe_unused]] unsigned tacc = 0;
#ifndef B2S_EMU
    if constexpr (SUM == 2) {
        __shared__ unsigned tmem_base_s;
        tacc = tm_alloc_cta<DP::TMEM_COLS>(&tmem_base_s, tid);
    } else
#endif
    {
        __syncthreads();
    }

    const int kout = p.kmax - p.kmin + 1;
    const int partner = (tid & 24) | ((8 - t) & 7);
    const bool is0 = (t == 0);
    const float edge = is0 ? 0.5f : 1.0f;

    // work units: static round-robin over the grid, or (p.work) an atomic counter the warps draw
    // 4 units at a time from -- the next draw is issued a whole unit ahead, so its latency is hidden
    const long long ustride = (long long)gridDim.x * DP::FPC;
    const bool dyn = p.work != nullptr;
    auto draw = [&]() -> long long {
        int b0 = 0;
        if ((tid & 31) == 0) b0 = atomicAdd(p.work, 4);
        return (long long)__shfl_sync(0xffffffffu, b0, 0);
    };
    long long ub_next = dyn ? draw() : (long long)blockIdx.x * DP::FPC + (grp & ~3);
    while (ub_next < p.n_units) {
        const long long ub = ub_next;
        ub_next = dyn ? draw() : ub + ustride;
        long long u = ub + (grp & 3);
        const bool uvalid = u < p.n_units;
        if (!uvalid) u = p.n_units - 1;
        long long b = u / p.units_per_signal;
        int c = (int)(u - b * p.units_per_signal);
        int f_begin = c * p.chunk_frames;
        int f_end = (f_begin + p.chunk_frames < p.nframes) ? f_begin + p.chunk_frames : p.nframes;
        [[maybe_unused]] int blk = 0, nsweeps = 0;
        [[maybe_unused]] bool svalid = true;
        if constexpr (SUM != 0) {
            // unit = (sweep block, duo c), the duo index fastest; the run is the one duo (frames 2 c, 2 c + 1)
            // of the block's first sweep, and the loop below walks the sweeps instead of the frames
            blk = (int)b;
            const int nduos = (p.nframes + 1) >> 1;
            svalid = c < nduos;
            if (!svalid) c = nduos - 1;
            f_begin = 2 * c;
            f_end = (f_begin + 2 < p.nframes) ? f_begin + 2 : p.nframes;
            b = (long long)blk * p.acc_rows;
            nsweeps = (int)((b + p.acc_rows < p.acc_batch) ? p.acc_rows : p.acc_batch - b);
        }
        const Tin* xb = reinterpret_cast<const Tin*>(p.x) + b * p.x_batch_stride + p.frame0 * (long long)p.hop + 2 * t;
        float* ob = p.out + b * p.out_batch_stride - ((MODE == EPI_BAND) ? 0 : p.kmin);

        // raw samples of the duo: slot i <-> complex index t + 8 i relative to frame f
        float2 cur[NCUR];
        {
            const Tin* const xf = xb + (long long)f_begin * p.hop;
            const Tin* const xfB = xf + bskew - ((f_begin + 1 < f_end) ? 0 : p.hop);
#pragma unroll
            for (int i = 0; i < NCUR; ++i) cur[i] = Loader<Tin>::ld2(((i < 16) ? xf : xfB) + 16 * i);
        }
        // the four duos of a warp run the same trip count (the longest run's)
        int ntrip = uvalid ? (f_end - f_begin + 1) >> 1 : 0;
        if constexpr (SUM != 0) {
            ntrip = nsweeps;                     // (warp-uniform: the four duos of a warp share the block)
#ifndef B2S_EMU
            if constexpr (SUM == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) tm_st4(tacc + 4 * j, 0.f, 0.f, 0.f, 0.f);
                tm_st2(tacc + 32, 0.f, 0.f);
            } else
#endif
            {
#pragma unroll
                for (int j = 0; j < DP::ACC_SLOTS; ++j) sacc[j * DP::NT] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            int o = __shfl_xor_sync(0xffffffffu, ntrip, 8);
            ntrip = ntrip > o ? ntrip : o;
            o = __shfl_xor_sync(0xffffffffu, ntrip, 16);
            ntrip = ntrip > o ? ntrip : o;
        }

        int f = f_begin;
        for (int it = 0; it < ntrip; ++it, f += (SUM != 0 ? 0 : 2)) {
            const bool actA = uvalid && svalid && (f < f_end);
            const bool actB = uvalid && svalid && (f + 1 < f_end);

            // ---- detrend + window, packing frame A (slots 0..15) and B (slots S..S+15) ----
            cpx2 v[16];
            if (p.detrend) {
                float2 cs;
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
                    cs = cmk(sa[0], sb[0]);
                } else {
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
                    if constexpr (NB == 2) cs = cmk(blk[0] + blk[1], blk[1] + blk[2]);
                    else if constexpr (NB == 4) cs = cmk((blk[0] + blk[1]) + (blk[2] + blk[3]), (blk[1] + blk[2]) + (blk[3] + blk[4]));
                    else cs = cmk(((blk[0] + blk[1]) + (blk[2] + blk[3])) + ((blk[4] + blk[5]) + (blk[6] + blk[7])),
                                  ((blk[1] + blk[2]) + (blk[3] + blk[4])) + ((blk[5] + blk[6]) + (blk[7] + blk[8])));
                }
#pragma unroll
                for (int o = G / 2; o >= 1; o >>= 1)
                    cs = pk_add(cs, cmk(__shfl_xor_sync(0xffffffffu, cs.x, o), __shfl_xor_sync(0xffffffffu, cs.y, o)));
                const float2 cm = pk_muls(cs, 1.0f / (float)DP::N);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i].re = cmk(cur[i].x - cm.x, cur[i + S].x - cm.y);
                    v[i].im = cmk(cur[i].y - cm.x, cur[i + S].y - cm.y);
                }
                float2 sr[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) sr[i] = pk_add(v[i].re, v[i].im);
#pragma unroll
                for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                    for (int i = 0; i < w; ++i) sr[i] = pk_add(sr[i], sr[i + w]);
                float2 tot = sr[0];
#pragma unroll
                for (int o = G / 2; o >= 1; o >>= 1)
                    tot = pk_add(tot, cmk(__shfl_xor_sync(0xffffffffu, tot.x, o), __shfl_xor_sync(0xffffffffu, tot.y, o)));
                const float2 nr = pk_muls(tot, -1.0f / (float)DP::N);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float4 w = sm4[DP::OFF_WIN + jj * 8 + t];
                    v[2 * jj].re = pk_fmas(v[2 * jj].re, w.x, pk_muls(nr, w.x));
                    v[2 * jj].im = pk_fmas(v[2 * jj].im, w.y, pk_muls(nr, w.y));
                    v[2 * jj + 1].re = pk_fmas(v[2 * jj + 1].re, w.z, pk_muls(nr, w.z));
                    v[2 * jj + 1].im = pk_fmas(v[2 * jj + 1].im, w.w, pk_muls(nr, w.w));
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float4 w = sm4[DP::OFF_WIN + jj * 8 + t];
                    v[2 * jj].re = cmk(cur[2 * jj].x * w.x, cur[2 * jj + S].x * w.x);
                    v[2 * jj].im = cmk(cur[2 * jj].y * w.y, cur[2 * jj + S].y * w.y);
                    v[2 * jj + 1].re = cmk(cur[2 * jj + 1].x * w.z, cur[2 * jj + 1 + S].x * w.z);
                    v[2 * jj + 1].im = cmk(cur[2 * jj + 1].y * w.w, cur[2 * jj + 1 + S].y * w.w);
                }
            }

            // ---- next duo (frames f+2, f+3): keep the overlap, prefetch the 2 S new slots ----
            if constexpr (SUM != 0) {            // the same duo of the next sweep: all 16 + S slots, one sweep ahead
                if (it + 1 < ntrip) xb += p.x_batch_stride;
                const Tin* const xn = xb + (long long)f * p.hop;
                const Tin* const xnB = xn + bskew - ((f + 1 < f_end) ? 0 : p.hop);
#pragma unroll
                for (int i = 0; i < NCUR; ++i) cur[i] = Loader<Tin>::ld2(((i < 16) ? xn : xnB) + 16 * i);
            } else {
#pragma unroll
                for (int i = 0; i < KEEP; ++i) cur[i] = cur[i + 2 * S];
                const int fa = (f + 2 < f_end) ? f + 2 : f_end - 1;
                const Tin* const xn = xb + (long long)fa * p.hop;
                const Tin* const xnB = xn + bskew - ((fa + 1 < f_end) ? 0 : p.hop);
#pragma unroll
                for (int i = KEEP; i < NCUR; ++i) cur[i] = Loader<Tin>::ld2(((i < 16) ? xn : xnB) + 16 * i);
            }

            // ---- pass 0: radix-16 over the lane's points, then the 8 x 16 transpose ----
            c2radix16(v);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const cpx2 z = v[perm16(q)];
                buf[ROW * t + q] = make_float4(z.re.x, z.re.y, z.im.x, z.im.y);
            }
            __syncwarp();

            // ---- pass 1: lane q = t holds columns q and q + 8; twiddle, two 8-point DFTs ----
            cpx2 c1[8], c2[8];
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) {
                const float4 qa = buf[ROW * tt + t], qb = buf[ROW * tt + t + 8];
                c1[tt] = cpx2{cmk(qa.x, qa.y), cmk(qa.z, qa.w)};
                c2[tt] = cpx2{cmk(qb.x, qb.y), cmk(qb.z, qb.w)};
            }
#pragma unroll
            for (int tt = 1; tt < 8; ++tt) {
                const float4 w = sm4[DP::OFF_TW + tt * 8 + t];
                c1[tt] = c2mul(c1[tt], cmk(w.x, w.y));
                c2[tt] = c2mul(c2[tt], cmk(w.z, w.w));
            }
            SmallFft2<8>::run(c1);               // Z[t + 16 pp]     in c1[pp]
            SmallFft2<8>::run(c2);               // Z[t + 8 + 16 pp] in c2[pp]
            __syncwarp();                        // every lane has consumed its exchange reads

            // ---- real-FFT split + PSD ----
            float* const rowA = ob + (long long)f * kout;
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
                // general lane: Z[t + 16 pp] with the partner's second column at 7 - pp.
                // lane 0 (its own partner): pp < 4: Z[16 pp] with Z[16 ((8 - pp) & 7)] (pp = 0: itself);
                //                           pp >= 4: Z[8 + 16 (pp - 4)] with Z[8 + 16 (11 - pp)]
                [[maybe_unused]] float4 a4;
#ifndef B2S_EMU
                if constexpr (SUM == 2) {
                    if (pp == 0) tm_st_wait();                         // the previous sweep's updates have landed
                    tm_ld4(tacc + 4 * pp, a4.x, a4.y, a4.z, a4.w);     // in flight during the butterfly below
                }
#endif
                cpx2 zk = c1[pp];
                const cpx2 sg = c2[7 - pp];
                const cpx2 s0 = (pp < 4) ? c1[(8 - pp) & 7] : c2[11 - pp];
                if (pp >= 4) {
                    const cpx2 z0 = c2[pp - 4];
                    zk.re = cmk(is0 ? z0.re.x : zk.re.x, is0 ? z0.re.y : zk.re.y);
                    zk.im = cmk(is0 ? z0.im.x : zk.im.x, is0 ? z0.im.y : zk.im.y);
                }
                const float a0 = is0 ? s0.re.x : sg.re.x, a1 = is0 ? s0.re.y : sg.re.y;
                const float a2 = is0 ? s0.im.x : sg.im.x, a3 = is0 ? s0.im.y : sg.im.y;
                cpx2 zm;
                zm.re = cmk(__shfl_sync(0xffffffffu, a0, partner), __shfl_sync(0xffffffffu, a1, partner));
                zm.im = cmk(__shfl_sync(0xffffffffu, a2, partner), __shfl_sync(0xffffffffu, a3, partner));
                const float4 w4 = sm4[DP::OFF_TWP + (pp >> 1) * 8 + t];
                const float2 w = (pp & 1) ? cmk(w4.z, w4.w) : cmk(w4.x, w4.y);
                const cpx2 e{pk_add(zk.re, zm.re), pk_sub(zk.im, zm.im)};      // 2E = zk + conj(zm)
                const cpx2 o{pk_add(zk.im, zm.im), pk_sub(zm.re, zk.re)};      // 2O = -i (zk - conj(zm))
                const cpx2 tw = c2mul(o, w);
                const cpx2 a = c2add(e, tw), bq = c2sub(e, tw);
                float2 pk = pk_fma(a.re, a.re, pk_mul(a.im, a.im));
                float2 pm = pk_fma(bq.re, bq.re, pk_mul(bq.im, bq.im));
                if (pp == 0) {
                    pk = pk_muls(pk, edge);
                    pm = pk_muls(pm, edge);
                }
                const int k = duo256_low_bin(t, pp);
                put(k, pk);
                put(M - k, pm);
                if constexpr (SUM != 0) {
#ifndef B2S_EMU
                    if constexpr (SUM == 2) {
                        tm_ld_wait4(a4.x, a4.y, a4.z, a4.w);
                        const float2 a0 = pk_add(cmk(a4.x, a4.y), pk), a1 = pk_add(cmk(a4.z, a4.w), pm);
                        tm_st4(tacc + 4 * pp, a0.x, a0.y, a1.x, a1.y);
                    } else
#endif
                    {
                        a4 = sacc[pp * DP::NT];
                        const float2 a0 = pk_add(cmk(a4.x, a4.y), pk), a1 = pk_add(cmk(a4.z, a4.w), pm);
                        sacc[pp * DP::NT] = make_float4(a0.x, a0.y, a1.x, a1.y);
                    }
                }
            }
            {   // k = 64: X = conj(Z[64]), held by lane 0 (first column, pp = 4)
                const cpx2 z = c1[4];
                const float2 pw = pk_muls(pk_fma(z.re, z.re, pk_mul(z.im, z.im)), 4.0f);
                if (is0) put(M / 2, pw);
                if constexpr (SUM != 0) {
#ifndef B2S_EMU
                    if constexpr (SUM == 2) {        // (warp-wide accesses: every lane keeps such a sum, lane 0's is used)
                        float2 am;
                        tm_ld2(tacc + 32, am.x, am.y);
                        tm_ld_wait2(am.x, am.y);
                        am = pk_add(am, pw);
                        tm_st2(tacc + 32, am.x, am.y);
                    } else
#endif
                    {
                        float2* const q = reinterpret_cast<float2*>(sacc + 8 * DP::NT);
                        *q = pk_add(*q, pw);
                    }
                }
            }
            if constexpr (SUM != 0) ob += p.out_batch_stride;
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
        if constexpr (SUM != 0) {
            // ---- the block's partial sums: p.acc[blk][frame][bin] ----
            float4 a[DP::ACC_SLOTS];
#ifndef B2S_EMU
            if constexpr (SUM == 2) {
                tm_st_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    tm_ld4(tacc + 4 * j, a[j].x, a[j].y, a[j].z, a[j].w);
                    tm_ld_wait4(a[j].x, a[j].y, a[j].z, a[j].w);
                }
                a[8] = make_float4(0.f, 0.f, 0.f, 0.f);
                tm_ld2(tacc + 32, a[8].x, a[8].y);
                tm_ld_wait2(a[8].x, a[8].y);
            } else
#endif
            {
#pragma unroll
                for (int j = 0; j < DP::ACC_SLOTS; ++j) a[j] = sacc[j * DP::NT];
            }
            if (uvalid && svalid) {
                const bool hasB = f_begin + 1 < p.nframes;
                float* const sA = p.acc + ((long long)blk * p.nframes + f_begin) * (M + 1);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = duo256_low_bin(t, j);
                    sA[k] = a[j].x;
                    sA[M - k] = a[j].z;
                    if (hasB) {
                        sA[M + 1 + k] = a[j].y;
                        sA[M + 1 + M - k] = a[j].w;
                    }
                }
                if (is0) {
                    sA[M / 2] = a[8].x;
                    if (hasB) sA[M + 1 + M / 2] = a[8].y;
                }
            }
        }
    }
#ifndef B2S_EMU
    if constexpr (SUM == 2) tm_free_cta<DP::TMEM_COLS>(tacc, tid);
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

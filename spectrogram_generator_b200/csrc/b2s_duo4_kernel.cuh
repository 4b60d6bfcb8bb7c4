// Four-step frame-duo STFT -> PSD kernel for nperseg 1024, 2048, 4096 with hop = S * nperseg/16
// (S = 2, 4, 8: 87.5 %, 75 %, 50 % overlap).  The frame-duo scheme of b2s_duo_kernel.cuh (two
// consecutive frames packed in fp32x2 registers, sliding register window over the raw
// samples) extended past M = 256 complex points by decimation in time:
//
//   M = 256 R (R = 2, 4, 8).  Sub-transform s (0 <= s < R) takes the points n = s + R m and is
//   owned by 16 lanes, exactly as the nperseg-512 kernel owns its frame: radix-16 over the
//   lane's 16 points, ONE 16 x 16 transpose through a buffer private to the half-warp
//   (__syncwarp only), radix-16 again -> F_s[q + 16 p] in lane q.  The R sub-transforms of a
//   duo (1, 2 or 4 warps) then meet once: F_s goes to shared memory, the group synchronises,
//   and the fused final stage does the radix-R butterflies Z[kap + 256 a] = sum_s W_M^(s kap)
//   W_R^(s a) F_s[kap] for kap and its mirror 256 - kap, the real-FFT split and the PSD, and
//   stores bins that are consecutive across the lanes of a warp (128-byte rows).
//
// Against the generic two-pass scheme (b2s_duo_cta_kernel.cuh) this halves the group-wide
// exchanges and barriers, keeps the raw samples of the overlapping frames in registers (each
// sample is requested once per run) and reads the pass-1 twiddles once per warp.
//
// SUM mode (stft_psd_duo4_sum_kernel, nperseg 2048 / 4096: per-sweep spectrograms AND their cross-sweep sum
// in one pass): as in b2s_duo_sum_kernel.cuh the thread group keeps ONE frame duo and walks
// over a block of consecutive sweeps, loading all its 16 + S slots afresh one sweep ahead; rows bit-identical
// to the per-sweep kernel's.  A thread's final-stage tasks produce 128 / G x 2 R = 16 packed power values
// (+ bin M/2 in thread 0 of the duo); they are added in sweep order into 34 running sums per thread --
// tensor memory (SUM = 2, the product path) or shared memory (SUM = 1: emulator / residency twin).
#pragma once

#include <type_traits>

#include "b2s_duo_cta_kernel.cuh"
#include "b2s_tmem.cuh"

namespace b2s {

template <int LOG2N>
struct Duo4Plan {
    using PL = Plan<LOG2N>;
    static constexpr int N = PL::N, M = PL::M;
    static constexpr int R = M / 256;                    // sub-transforms per frame
    static constexpr int G = 16 * R;                     // threads per duo
    static constexpr int NT = 128;
    static constexpr int MINB = 3;
    static constexpr int FPC = NT / G;                   // duos in flight per CTA
    static constexpr int NSUB = NT / 16;                 // sub-transform buffers per CTA
    static constexpr int RED = (G > 32) ? G / 32 : 1;    // warps per duo
    static constexpr int TPT = 128 / G;                  // final tasks per thread
    static constexpr int ROW = 17;
    // float4 slots per sub-transform buffer: 16 rows of ROW, + 8 / R so that the R buffers of a duo start 8 / R
    // quarter-bank-groups apart -- the pass-0 results are written by threads that hold R CONSECUTIVE sub-transforms
    // (see "thread <-> point" below), and a quarter-warp of their 128-bit stores must not meet in a bank
    static constexpr int BUF = 16 * ROW + 8 / R;
    // shared memory (float4 units): window taps [8][G] (slots 2j, 2j+1 of thread g), pass-1
    // twiddles [8][16], the NSUB exchange buffers, reduction slots
    static constexpr int OFF_WIN = 0;
    static constexpr int OFF_TW1 = OFF_WIN + 8 * G;
    static constexpr int OFF_BUF = OFF_TW1 + 8 * 16;
    static constexpr int OFF_RED = OFF_BUF + NSUB * BUF;
    static constexpr int TOTAL = OFF_RED + FPC * (3 * RED + 1);    // one float4 per reduction slot + the unit draw
    // final-stage twiddles W_M^(r kap), W_N^k (float2 units after the float4 region) for R = 4: they
    // are read once per task and frame, so their L1 latency is exposed (C3: +2 %).  For R = 8 the
    // extra 30 KB would leave too little L1 for the strided sample loads (measured: 4096/1024 0.59 ->
    // 0.72 ms); for R = 2 staging them only lengthens the prologue of small launches.
    static constexpr bool POST_IN_SMEM = (R == 4);
    static constexpr int FIN2 = POST_IN_SMEM ? (R - 1) * 256 : 0;
    static constexpr int POST2 = POST_IN_SMEM ? (M + 2) : 0;
    static constexpr size_t SMEM = (size_t)TOTAL * sizeof(float4) + (size_t)(FIN2 + POST2) * sizeof(float2);
    // SUM mode: packed (A, B) running sums per thread -- TPT tasks x 2 R bins, + bin M/2 (thread 0 of the duo)
    static constexpr int ACC_SLOTS = TPT * 2 * R + 1;    // 17 float2
    static constexpr size_t SUM_SMEM = (SMEM + 15) / 16 * 16 + (size_t)ACC_SLOTS * NT * sizeof(float2);
    static constexpr int TMEM_COLS = 64;                 // 34 used
    static_assert(R == 2 || R == 4 || R == 8, "four-step duo kernel: nperseg 1024, 2048, 4096");
    static_assert(PL::NS == 256 && PL::GF == R, "plan tables");
};

// sum over the G lanes of a duo of a packed value: fixed xor-butterfly inside the warp, then a
// fixed-order sum of the per-warp partials
template <int LOG2N>
B2S_DEVICE float2 duo4_group_sum(float2 v, int grp, int j, float4* red) {
    using DP = Duo4Plan<LOG2N>;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        v = pk_add(v, cmk(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o)));
    if constexpr (DP::G > 32) {
        if ((j & 31) == 0) red[j >> 5] = make_float4(v.x, v.y, 0.f, 0.f);
        b2s_bar_sync(grp + 1, DP::G);
        float4 q = red[0];
        v = cmk(q.x, q.y);
#pragma unroll
        for (int w = 1; w < DP::RED; ++w) {
            q = red[w];
            v = pk_add(v, cmk(q.x, q.y));
        }
    }
    return v;
}

template <int LOG2N, typename Tin, int S, int MODE, int SUM>
B2S_DEVICE void stft_psd_duo4_body(const StftParams& p);

template <int LOG2N, typename Tin, int S, int MODE>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(Duo4Plan<LOG2N>::NT, Duo4Plan<LOG2N>::MINB) stft_psd_duo4_kernel(const StftParams p) {
    stft_psd_duo4_body<LOG2N, Tin, S, MODE, 0>(p);
}
// per-sweep rows + cross-sweep block sums (SUM = 1: sums in shared memory, 2: in tensor memory);
// units: plan_stft_sum(..., duos_per_warp = 1) -- a duo is one or more whole warps
template <int LOG2N, typename Tin, int S, int SUM>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(Duo4Plan<LOG2N>::NT, Duo4Plan<LOG2N>::MINB) stft_psd_duo4_sum_kernel(const StftParams p) {
    stft_psd_duo4_body<LOG2N, Tin, S, EPI_PLAIN, SUM>(p);
}

template <int LOG2N, typename Tin, int S, int MODE, int SUM>
B2S_DEVICE void stft_psd_duo4_body(const StftParams& p) {
    static_assert(SUM == 0 || MODE == EPI_PLAIN, "SUM mode: plain epilogue");
    using PL = Plan<LOG2N>;
    using DP = Duo4Plan<LOG2N>;
    constexpr int M = DP::M, N = DP::N, R = DP::R, G = DP::G, ROW = DP::ROW;
    constexpr int NCUR = 16 + S;
    constexpr int KEEP = (NCUR > 2 * S) ? NCUR - 2 * S : 0;
    constexpr int SLOT = 32 * R;                         // samples between a lane's consecutive slots

    B2S_DYN_SMEM_F4(sm4);
    const int tid = (int)threadIdx.x;
    const int grp = tid / G;                             // duo inside the CTA
    const int j = tid - grp * G;                         // thread inside the duo
    // thread <-> point.  From the transpose on, thread j is lane t = j & 15 of sub-transform s = j >> 4 (16 lanes own a
    // sub-transform, as the 512-point kernel owns its frame).  BEFORE it -- loads, detrend, window, the in-lane
    // radix-16 over the slots -- any assignment works, and thread j takes sub-transform sL = j % R, row tL = j / R:
    // its slot i is complex point sL + R (tL + 16 i) = j + G i, so the G threads of a duo read G CONSECUTIVE points
    // (a warp: 256 contiguous bytes) instead of 16 points R apart per half-warp.  The pass-0 results change hands in
    // the transpose buffer anyway.  (Round 2; before: 8 wavefronts per LDG.64 at nperseg 4096, 4 at 2048.)
    const int s = j >> 4;                                // sub-transform (transpose read side, pass 1, final stage)
    const int t = j & 15;                                // lane inside the sub-transform
    const int sL = j & (R - 1);                          // sub-transform and row of the points this thread LOADS
    const int tL = j / R;
    float4* const bufs = sm4 + DP::OFF_BUF + (grp * R) * DP::BUF;    // the duo's R buffers
    float4* const buf = bufs + s * DP::BUF;
    float4* const bufL = bufs + sL * DP::BUF;
    float4* const red = sm4 + DP::OFF_RED + grp * (3 * DP::RED + 1);
    float2* const stab = reinterpret_cast<float2*>(sm4 + DP::TOTAL);
    auto fin = [&](int i) -> float2 {            // W_M^(r kap) at (r - 1) * 256 + kap
        if constexpr (DP::POST_IN_SMEM) return stab[i];
        else return __ldg(p.tw + PL::OFF_FIN + i);
    };
    auto post = [&](int k) -> float2 {           // W_N^k
        if constexpr (DP::POST_IN_SMEM) return stab[DP::FIN2 + k];
        else return __ldg(p.tw + PL::OFF_POST + k);
    };

    // ---- constant tables, once per CTA; the PSD scale goes into the window ----
    {
        if constexpr (DP::POST_IN_SMEM) {
            for (int i = tid; i < DP::FIN2; i += DP::NT) stab[i] = __ldg(p.tw + PL::OFF_FIN + i);
            for (int i = tid; i <= M; i += DP::NT) stab[DP::FIN2 + i] = __ldg(p.tw + PL::OFF_POST + i);
        }
        const float csc = sqrtf(0.5f * p.scale);
        const float2* w2 = reinterpret_cast<const float2*>(p.window);
        for (int i = tid; i < 8 * G; i += DP::NT) {
            const int jj = i / G, g = i - jj * G;
            const int n0 = g + G * (2 * jj);                                  // slot 2 jj of thread g: point g + G i
            const float2 wa = __ldg(w2 + n0), wb = __ldg(w2 + n0 + 16 * R);
            sm4[DP::OFF_WIN + i] = make_float4(wa.x * csc, wa.y * csc, wb.x * csc, wb.y * csc);
        }
        for (int i = tid; i < 8 * 16; i += DP::NT) {
            const int jj = i >> 4, l = i & 15;
            const float2 ta = (jj == 0) ? cmk(1.f, 0.f) : __ldg(p.tw + PL::OFF_P1 + (2 * jj - 1) * 16 + l);
            const float2 tb = __ldg(p.tw + PL::OFF_P1 + (2 * jj) * 16 + l);
            sm4[DP::OFF_TW1 + i] = make_float4(ta.x, ta.y, tb.x, tb.y);
        }
    }
    [[maybe_unused]] float2* const sacc =
        reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(sm4) + (DP::SMEM + 15) / 16 * 16) + tid;    // [slot][NT]
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
    typename std::conditional<SUM != 0, EpiDuoRec<2 * R + 1>, EpiDuo<MODE>>::type epi;
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
        long long b = u / p.units_per_signal;
        const int c = (int)(u - b * p.units_per_signal);
        int f_begin = c * p.chunk_frames;
        int f_end = (f_begin + p.chunk_frames < p.nframes) ? f_begin + p.chunk_frames : p.nframes;
        [[maybe_unused]] int blk = 0, nsweeps = 1;
        if constexpr (SUM != 0) {
            // unit = (sweep block, duo c), the duo index fastest; the "run" is the one duo (frames 2 c, 2 c + 1) of
            // the block's first sweep, and the loop below walks the sweeps instead of the frames
            blk = (int)b;
            f_begin = 2 * c;
            f_end = (f_begin + 2 < p.nframes) ? f_begin + 2 : p.nframes;
            b = (long long)blk * p.acc_rows;
            nsweeps = (int)((b + p.acc_rows < p.acc_batch) ? p.acc_rows : p.acc_batch - b);
#ifndef B2S_EMU
            if constexpr (SUM == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) tm_st4(tacc + 4 * i, 0.f, 0.f, 0.f, 0.f);
                tm_st2(tacc + 32, 0.f, 0.f);
            } else
#endif
            {
#pragma unroll
                for (int i = 0; i < DP::ACC_SLOTS; ++i) sacc[i * DP::NT] = cmk(0.f, 0.f);
            }
        }
        const Tin* xb = reinterpret_cast<const Tin*>(p.x) + b * p.x_batch_stride + p.frame0 * (long long)p.hop + 2 * j;
        float* ob = p.out + b * p.out_batch_stride - ((MODE == EPI_BAND) ? 0 : p.kmin);

        // raw samples of the duo: slot i <-> complex index s + R (t + 16 i) relative to frame f;
        // slots 16.. belong to frame B only (without a frame B they re-read the previous S slots)
        float2 cur[NCUR];
        {
            const Tin* const xf = xb + (long long)f_begin * p.hop;
            const Tin* const xfB = xf - ((f_begin + 1 < f_end) ? 0 : p.hop);
#pragma unroll
            for (int i = 0; i < NCUR; ++i) cur[i] = Loader<Tin>::ld2(((i < 16) ? xf : xfB) + SLOT * i);
        }

        for (int f = f_begin, it = 0; (SUM != 0) ? (it < nsweeps) : (f < f_end); f += (SUM != 0 ? 0 : 2), ++it) {
            epi.actA = true;
            epi.actB = f + 1 < f_end;
            epi.rowA = ob + (long long)f * kout;

            // ---- detrend + window, packing frame A (slots 0..15) and B (slots S..S+15) ----
            cpx2 v[16];
            if (p.detrend) {
                // coarse per-frame means (pivots): per-slot sums in a fixed frame-relative order
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
                const float2 cm = pk_muls(duo4_group_sum<LOG2N>(cs, grp, j, red), 1.0f / (float)N);
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
                const float2 nr = pk_muls(duo4_group_sum<LOG2N>(sr[0], grp, j, red + DP::RED), -1.0f / (float)N);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float4 w = sm4[DP::OFF_WIN + jj * G + j];
                    v[2 * jj].re = pk_fmas(v[2 * jj].re, w.x, pk_muls(nr, w.x));
                    v[2 * jj].im = pk_fmas(v[2 * jj].im, w.y, pk_muls(nr, w.y));
                    v[2 * jj + 1].re = pk_fmas(v[2 * jj + 1].re, w.z, pk_muls(nr, w.z));
                    v[2 * jj + 1].im = pk_fmas(v[2 * jj + 1].im, w.w, pk_muls(nr, w.w));
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float4 w = sm4[DP::OFF_WIN + jj * G + j];
                    v[2 * jj].re = cmk(cur[2 * jj].x * w.x, cur[2 * jj + S].x * w.x);
                    v[2 * jj].im = cmk(cur[2 * jj].y * w.y, cur[2 * jj + S].y * w.y);
                    v[2 * jj + 1].re = cmk(cur[2 * jj + 1].x * w.z, cur[2 * jj + 1 + S].x * w.z);
                    v[2 * jj + 1].im = cmk(cur[2 * jj + 1].y * w.w, cur[2 * jj + 1 + S].y * w.w);
                }
            }

            // ---- next duo (frames f+2, f+3): keep the overlap, prefetch the 2 S new slots ----
            if constexpr (SUM != 0) {            // the same duo of the next sweep: all 16 + S slots, one sweep ahead
                if (it + 1 < nsweeps) xb += p.x_batch_stride;
                const Tin* const xn = xb + (long long)f * p.hop;
                const Tin* const xnB = xn - ((f + 1 < f_end) ? 0 : p.hop);
#pragma unroll
                for (int i = 0; i < NCUR; ++i) cur[i] = Loader<Tin>::ld2(((i < 16) ? xn : xnB) + SLOT * i);
            } else {
#pragma unroll
                for (int i = 0; i < KEEP; ++i) cur[i] = cur[i + 2 * S];
                const int fa = (f + 2 < f_end) ? f + 2 : f_end - 1;
                const Tin* const xn = xb + (long long)fa * p.hop;
                const Tin* const xnB = xn - ((fa + 1 < f_end) ? 0 : p.hop);
#pragma unroll
                for (int i = KEEP; i < NCUR; ++i) cur[i] = Loader<Tin>::ld2(((i < 16) ? xn : xnB) + SLOT * i);
            }

            // ---- sub-transform: radix-16, 16 x 16 transpose inside the half-warp, radix-16 ----
            c2radix16(v);
            duo_group_sync<G>(grp);              // the previous duo's final-stage reads are done
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const cpx2 z = v[perm16(q)];
                bufL[ROW * tL + q] = make_float4(z.re.x, z.re.y, z.im.x, z.im.y);
            }
            duo_group_sync<G>(grp);              // the rows of a sub-transform come from R different warps
#pragma unroll
            for (int tt = 0; tt < 16; ++tt) {
                const float4 q4 = buf[ROW * tt + t];
                v[tt] = cpx2{cmk(q4.x, q4.y), cmk(q4.z, q4.w)};
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float4 w = sm4[DP::OFF_TW1 + jj * 16 + t];
                if (jj > 0) v[2 * jj] = c2mul(v[2 * jj], cmk(w.x, w.y));
                v[2 * jj + 1] = c2mul(v[2 * jj + 1], cmk(w.z, w.w));
            }
            c2radix16(v);
            __syncwarp();                        // every lane of the half-warp has consumed its reads
            // F_s[kap], kap = t + 16 p, natural order in the sub-transform's own buffer
#pragma unroll
            for (int pp = 0; pp < 16; ++pp) {
                const cpx2 z = v[perm16(pp)];
                buf[t + 16 * pp] = make_float4(z.re.x, z.re.y, z.im.x, z.im.y);
            }
            duo_group_sync<G>(grp);

            // ---- fused final stage: radix-R butterflies + real-FFT split + PSD, both frames ----
#pragma unroll
            for (int cc = 0; cc < DP::TPT; ++cc) {
                const int kap = j + G * cc;            // task id == kappa in [0, 128)
                cpx2 U[R], V[R];
                // SUM: the task's 2 R running sums (packed A, B), fetched while the butterflies run
                [[maybe_unused]] float2 acc[2 * R + 1];
                if constexpr (SUM != 0) {
                    epi.nrec = 0;
#ifndef B2S_EMU
                    if constexpr (SUM == 2) {
                        if (cc == 0) tm_st_wait();                     // the previous sweep's updates have landed
#pragma unroll
                        for (int i = 0; i < R; ++i)
                            tm_ld4(tacc + 4 * (R * cc + i), acc[2 * i].x, acc[2 * i].y, acc[2 * i + 1].x, acc[2 * i + 1].y);
                        if (cc == 0) tm_ld2(tacc + 32, acc[2 * R].x, acc[2 * R].y);
                    }
#endif
                }
                if (kap != 0) {
                    const int kap2 = 256 - kap;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float4 qa = bufs[r * DP::BUF + kap], qb = bufs[r * DP::BUF + kap2];
                        U[r] = cpx2{cmk(qa.x, qa.y), cmk(qa.z, qa.w)};
                        V[r] = cpx2{cmk(qb.x, qb.y), cmk(qb.z, qb.w)};
                    }
#pragma unroll
                    for (int r = 1; r < R; ++r) {
                        U[r] = c2mul(U[r], fin((r - 1) * 256 + kap));
                        V[r] = c2mul(V[r], fin((r - 1) * 256 + kap2));
                    }
                    SmallFft2<R>::run(U);
                    SmallFft2<R>::run(V);
#pragma unroll
                    for (int a = 0; a < R; ++a) {
                        const int k = kap + a * 256;
                        epi.pair(k, M - k, U[a], V[R - 1 - a], post(k), 1.0f);
                    }
                } else {
                    // kappa = 0 and kappa = 128 are their own mirrors (thread 0 of the duo)
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float4 qa = bufs[r * DP::BUF], qb = bufs[r * DP::BUF + 128];
                        U[r] = cpx2{cmk(qa.x, qa.y), cmk(qa.z, qa.w)};
                        V[r] = cpx2{cmk(qb.x, qb.y), cmk(qb.z, qb.w)};
                    }
#pragma unroll
                    for (int r = 1; r < R; ++r) V[r] = c2mul(V[r], fin((r - 1) * 256 + 128));
                    SmallFft2<R>::run(U);
                    SmallFft2<R>::run(V);
                    epi.pair(0, M, U[0], U[0], cmk(1.f, 0.f), 0.5f);           // DC / Nyquist carry scale, not 2 scale
#pragma unroll
                    for (int a = 1; 2 * a < R; ++a)
                        epi.pair(a * 256, M - a * 256, U[a], U[R - a], post(a * 256), 1.0f);
                    {                                                       // k = M/2: X = conj(Z)
                        const cpx2 z = U[R / 2];
                        epi.put(M / 2, pk_muls(pk_fma(z.re, z.re, pk_mul(z.im, z.im)), 4.0f));
                    }
#pragma unroll
                    for (int a = 0; 2 * a < R - 1; ++a) {
                        const int k = 128 + a * 256;
                        epi.pair(k, M - k, V[a], V[R - 1 - a], post(k), 1.0f);
                    }
                }
                if constexpr (SUM != 0) {
                    // records 0 .. 2 R - 1 of this task -> slots 2 R cc + i; thread 0's task kap = 0 has one more
                    // (the bins in the order of duo4_task_bins), kept in the last slot
                    const bool extra = (cc == 0) && (j == 0);
#ifndef B2S_EMU
                    if constexpr (SUM == 2) {
                        if (cc == 0) tm_ld_wait2(acc[2 * R].x, acc[2 * R].y);
#pragma unroll
                        for (int i = 0; i < R; ++i) {
                            tm_ld_wait4(acc[2 * i].x, acc[2 * i].y, acc[2 * i + 1].x, acc[2 * i + 1].y);
                            const float2 a0 = pk_add(acc[2 * i], epi.rec[2 * i]), a1 = pk_add(acc[2 * i + 1], epi.rec[2 * i + 1]);
                            tm_st4(tacc + 4 * (R * cc + i), a0.x, a0.y, a1.x, a1.y);
                        }
                        if (cc == 0) {
                            const float2 ax = extra ? pk_add(acc[2 * R], epi.rec[2 * R]) : acc[2 * R];
                            tm_st2(tacc + 32, ax.x, ax.y);
                        }
                    } else
#endif
                    {
#pragma unroll
                        for (int i = 0; i < 2 * R; ++i)
                            sacc[(2 * R * cc + i) * DP::NT] = pk_add(sacc[(2 * R * cc + i) * DP::NT], epi.rec[i]);
                        if (extra) sacc[(DP::ACC_SLOTS - 1) * DP::NT] = pk_add(sacc[(DP::ACC_SLOTS - 1) * DP::NT], epi.rec[2 * R]);
                    }
                }
            }
            if constexpr (SUM != 0) ob += p.out_batch_stride;
            if constexpr (MODE == EPI_BAND) {
                const float2 bs = duo4_group_sum<LOG2N>(epi.band, grp, j, red + 2 * DP::RED);
                epi.band = cmk(0.f, 0.f);
                if (j == 0) {
                    ob[f] = bs.x;
                    if (epi.actB) ob[f + 1] = bs.y;
                }
            }
        }
        if constexpr (SUM != 0) {
            // ---- the block's partial sums: p.acc[blk][frame][bin], bins in the order the tasks put them ----
            float2 a[DP::ACC_SLOTS];
#ifndef B2S_EMU
            if constexpr (SUM == 2) {
                tm_st_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    tm_ld4(tacc + 4 * i, a[2 * i].x, a[2 * i].y, a[2 * i + 1].x, a[2 * i + 1].y);
                    tm_ld_wait4(a[2 * i].x, a[2 * i].y, a[2 * i + 1].x, a[2 * i + 1].y);
                }
                tm_ld2(tacc + 32, a[16].x, a[16].y);
                tm_ld_wait2(a[16].x, a[16].y);
            } else
#endif
            {
#pragma unroll
                for (int i = 0; i < DP::ACC_SLOTS; ++i) a[i] = sacc[i * DP::NT];
            }
            const bool hasB = f_begin + 1 < p.nframes;
            float* const sA = p.acc + ((long long)blk * p.nframes + f_begin) * (M + 1);
            auto out = [&](int k, float2 v) {
                sA[k] = v.x;
                if (hasB) sA[M + 1 + k] = v.y;
            };
#pragma unroll
            for (int cc = 0; cc < DP::TPT; ++cc) {
                const int kap = j + G * cc;
                if (kap != 0) {
#pragma unroll
                    for (int aa = 0; aa < R; ++aa) {
                        const int k = kap + aa * 256;
                        out(k, a[2 * R * cc + 2 * aa]);
                        out(M - k, a[2 * R * cc + 2 * aa + 1]);
                    }
                } else {
                    // thread 0 of the duo, task kap = 0: (0, M), (256 a, M - 256 a) for 2 a < R, M / 2,
                    // (128 + 256 a, M - 128 - 256 a) for 2 a < R - 1 -- 2 R + 1 values, the last one in the spare slot
                    int i = 0;
                    auto nxt = [&]() -> float2 {
                        const float2 v = (i < 2 * R) ? a[i] : a[DP::ACC_SLOTS - 1];
                        ++i;
                        return v;
                    };
                    out(0, nxt());
                    out(M, nxt());
#pragma unroll
                    for (int aa = 1; 2 * aa < R; ++aa) {
                        out(aa * 256, nxt());
                        out(M - aa * 256, nxt());
                    }
                    out(M / 2, nxt());
#pragma unroll
                    for (int aa = 0; 2 * aa < R - 1; ++aa) {
                        out(128 + aa * 256, nxt());
                        out(M - 128 - aa * 256, nxt());
                    }
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

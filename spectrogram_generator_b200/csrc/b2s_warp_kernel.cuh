// Warp-resident STFT -> PSD kernel for nperseg <= 1024 (G = nperseg/32 <= 32 threads
// per frame).  Same arithmetic as stft_psd_kernel (b2s_kernels.cuh) -- radix-16
// Stockham passes + fused final stage -- restructured around what the ncu
// profiles of that kernel showed (profiles/r1_*):
//
//   * every frame group lives inside one warp, and all groups of a warp run the
//     same trip count (finished groups recompute their last frame with stores
//     predicated off), so every sync is a full-mask __syncwarp and every shuffle
//     has a compile-time mask;
//   * SLIDING: when hop = SHIFT * nperseg/16 (50 %, 75 %, 87.5 % overlap and the
//     reference's default 12.5 %), the thread's 16 complex samples of frame f+1
//     are the samples of frame f shifted by SHIFT registers plus SHIFT new
//     loads.  Each sample is therefore loaded from global memory exactly once
//     per run of frames, and the SHIFT new loads are issued a whole frame ahead
//     (software prefetch into registers) -- the exposed load latency at the top
//     of each frame was the largest single stall of the previous kernel;
//   * the window and all twiddle tables are staged once per CTA in shared memory
//     and read with conflict-free, lane-consecutive LDS.64.
#pragma once

#include "b2s_kernels.cuh"

namespace b2s {

template <int LOG2N, int NT_ = 256>
struct WarpPlan {
    using PL = Plan<LOG2N>;
    static_assert(PL::G <= 32, "warp kernel handles nperseg <= 1024");
    static constexpr int NT = NT_;
    static constexpr int FPC = NT / PL::G;                 // groups per CTA
    static constexpr int GW = 32 / PL::G;                  // groups per warp
    // shared memory, in float2 units: [window N/2][tables TABLE (padded even)][FPC exchange buffers]
    static constexpr int OFF_WIN = 0;
    static constexpr int OFF_TAB = PL::N / 2;
    static constexpr int OFF_BUF = OFF_TAB + ((PL::TABLE + 1) & ~1);
    static constexpr int TOTAL = OFF_BUF + FPC * PL::BUF;
    static constexpr size_t SMEM = (size_t)TOTAL * sizeof(float2);
};

template <int G>
B2S_DEVICE float warp_group_mean(const float2 (&v)[16]) {
    float s[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) s[r] = v[r].x + v[r].y;
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
        for (int r = 0; r < w; ++r) s[r] += s[r + w];
    float tot = s[0];
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    return tot * (1.0f / (float)(32 * G));
}

// Front half of a frame for the warp kernel: detrend + window from the raw samples in
// `cur` into `v`, then slide `cur` to the next frame and issue its new loads.
// ROT rotates the register names instead of moving data: logical sample r of the current
// frame lives in cur[(r + ROT) & 15]; sliding by SHIFT is then just ROT += SHIFT for the
// next frame (the caller switches over the reachable ROT values), and the SHIFT new loads
// land in the slots the oldest samples occupied.
template <int LOG2N, typename Tin, int SHIFT, int ROT>
B2S_DEVICE void warp_front(float2 (&cur)[16], float2 (&v)[16], const float2* wtaps, int detrend, bool do_next,
                           const Tin* xn, int j, int vec_ok) {
    constexpr int G = Plan<LOG2N>::G;
    if (detrend) {
        const float m1 = warp_group_mean<G>(cur);
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = cmk(cur[(r + ROT) & 15].x - m1, cur[(r + ROT) & 15].y - m1);
        const float nr = -warp_group_mean<G>(v);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float2 w = wtaps[G * r];
            v[r].x = fmaf(v[r].x, w.x, nr * w.x);
            v[r].y = fmaf(v[r].y, w.y, nr * w.y);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float2 w = wtaps[G * r];
            v[r] = cmk(cur[(r + ROT) & 15].x * w.x, cur[(r + ROT) & 15].y * w.y);
        }
    }
    if (do_next) {
        if constexpr (SHIFT != 0) {
            // logical r' = 16-SHIFT+i of the next frame -> physical slot (i + ROT) & 15
#pragma unroll
            for (int i = 0; i < SHIFT; ++i)
                cur[(i + ROT) & 15] = Loader<Tin>::ld2(xn + 2 * (j + G * (16 - SHIFT + i)));
        } else if (vec_ok) {
#pragma unroll
            for (int r = 0; r < 16; ++r) cur[r] = Loader<Tin>::ld2(xn + 2 * (j + G * r));
        } else {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const Tin* q = xn + 2 * (j + G * r);
                cur[r] = cmk(Loader<Tin>::ld1(q), Loader<Tin>::ld1(q + 1));
            }
        }
    }
}

// SHIFT: 0 = reload the whole frame every time (any hop; scalar loads when !vec_ok),
//        2/4/8/14 = sliding register window, hop == SHIFT * nperseg / 16.
// (Keeping the window taps / pass-1 twiddles in registers was measured and lost: at 168-225
//  registers per thread the occupancy drop costs more than the 39 LDS per frame it saves.)
template <int LOG2N, typename Tin, int SHIFT, int MODE, int NT = 256, int MINB = 2>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(NT, MINB) stft_psd_warp_kernel(const StftParams p) {
    using PL = Plan<LOG2N>;
    using WP = WarpPlan<LOG2N, NT>;
    constexpr int M = PL::M, G = PL::G, NS = PL::NS, GF = PL::GF;

    B2S_DYN_SMEM_F2(sm);
    const int tid = (int)threadIdx.x;
    const int grp = tid / G;
    const int j = tid - grp * G;
    const int bufo = WP::OFF_BUF + grp * PL::BUF;          // this group's exchange buffer

    // ---- stage window + tables in shared memory (once per CTA) ----
    {
        const float2* w2 = reinterpret_cast<const float2*>(p.window);
        for (int i = tid; i < PL::N / 2; i += WP::NT) sm[WP::OFF_WIN + i] = __ldg(w2 + i);
        for (int i = tid; i < PL::TABLE; i += WP::NT) sm[WP::OFF_TAB + i] = __ldg(p.tw + i);
    }
    __syncthreads();

    const int kout = p.kmax - p.kmin + 1;
    Epi<MODE> epi;
    epi.s_edge = p.scale;
    epi.s_int = 2.0f * p.scale;
    epi.floor = p.db_floor;
    epi.kmin = p.kmin;
    epi.kmax = p.kmax;
    epi.db = p.out_mode;

    const long long ustride = (long long)gridDim.x * WP::FPC;
    // work units: static round-robin over the grid (warp-uniform: the first group of a warp has the
    // smallest unit index of the warp), or (p.work) an atomic counter the warps draw GW units at a
    // time from -- the next draw is issued a whole unit ahead, so its latency is hidden
    const bool dyn = p.work != nullptr;
    auto draw = [&]() -> long long {
        int b0 = 0;
        if ((tid & 31) == 0) b0 = atomicAdd(p.work, WP::GW);
        return (long long)__shfl_sync(0xffffffffu, b0, 0);
    };
    long long ub_next = dyn ? draw() : (long long)blockIdx.x * WP::FPC + (grp - (grp % WP::GW));
    while (ub_next < p.n_units) {
        const long long ub = ub_next;
        ub_next = dyn ? draw() : ub + ustride;
        long long u = ub + (grp % WP::GW);
        const bool uvalid = u < p.n_units;
        if (!uvalid) u = p.n_units - 1;
        const long long b = u / p.units_per_signal;
        const int c = (int)(u - b * p.units_per_signal);
        const int f_begin = c * p.chunk_frames;
        const int f_end = (f_begin + p.chunk_frames < p.nframes) ? f_begin + p.chunk_frames : p.nframes;
        const Tin* const xb = reinterpret_cast<const Tin*>(p.x) + b * p.x_batch_stride + p.frame0 * (long long)p.hop;
        float* const ob = p.out + b * p.out_batch_stride - p.kmin;

        // ---- first frame of the run: load all 16 complex points ----
        float2 cur[16];
        {
            const Tin* const xf = xb + (long long)f_begin * p.hop;
            if (SHIFT != 0 || p.vec_ok) {
#pragma unroll
                for (int r = 0; r < 16; ++r) cur[r] = Loader<Tin>::ld2(xf + 2 * (j + G * r));
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const Tin* q = xf + 2 * (j + G * r);
                    cur[r] = cmk(Loader<Tin>::ld1(q), Loader<Tin>::ld1(q + 1));
                }
            }
        }

        int rot = 0;          // warp-uniform register rotation of `cur` (see warp_front)
        for (int f = f_begin;; ++f) {
            const bool act = uvalid && (f < f_end);
            if (!__any_sync(0xffffffffu, act)) break;
            epi.act = act;
            epi.row = ob + (long long)f * kout;

            // ---- detrend + window -> v; slide / prefetch the next frame's samples ----
            float2 v[16];
            {
                const bool do_next = act && (f + 1 < f_end);
                const Tin* const xn = xb + (long long)(f + 1) * p.hop;
                const float2* const wt = &sm[WP::OFF_WIN + j];
                constexpr int STEP = (SHIFT == 0) ? 16 : ((SHIFT % 4 == 0) ? ((SHIFT % 8 == 0) ? 8 : 4) : 2);
                switch (rot) {
#define B2S_FRONT_CASE(R)                                                                                   \
    case R:                                                                                                 \
        if constexpr ((R) % STEP == 0)                                                                      \
            warp_front<LOG2N, Tin, SHIFT, (R)>(cur, v, wt, p.detrend, do_next, xn, j, p.vec_ok);            \
        break;
                    B2S_FRONT_CASE(0)
                    B2S_FRONT_CASE(2)
                    B2S_FRONT_CASE(4)
                    B2S_FRONT_CASE(6)
                    B2S_FRONT_CASE(8)
                    B2S_FRONT_CASE(10)
                    B2S_FRONT_CASE(12)
                    B2S_FRONT_CASE(14)
#undef B2S_FRONT_CASE
                    default: break;
                }
                rot = (rot + SHIFT) & 15;
            }

            // ---- pass 0: radix-16 over r (stride G), Ns 1 -> 16 ----
            radix16(v);
            __syncwarp();                       // previous frame's final-stage reads are done
            {
                float4* dst = reinterpret_cast<float4*>(&sm[bufo + 18 * j]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 a = v[perm16(2 * i)], bq = v[perm16(2 * i + 1)];
                    dst[i] = make_float4(a.x, a.y, bq.x, bq.y);
                }
            }
            __syncwarp();

            // ---- pass 1 (nperseg >= 512): radix-16 Stockham, Ns = 16 ----
            if constexpr (PL::P == 2) {
                const int jm = j & 15;
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = sm[bufo + phys(j + r * G)];
#pragma unroll
                for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], sm[WP::OFF_TAB + PL::OFF_P1 + (r - 1) * 16 + jm]);
                radix16(v);
                __syncwarp();
                const int base = (j - jm) * 16 + jm;
#pragma unroll
                for (int r = 0; r < 16; ++r) sm[bufo + phys(base + r * 16)] = v[perm16(r)];
                __syncwarp();
            }

            // ---- fused final stage: radix-GF butterflies + real-FFT split + PSD ----
#pragma unroll
            for (int cc = 0; cc < PL::TPT; ++cc) {
                const int kap = j + G * cc;
                float2 U[GF], V[GF];
                if (kap != 0) {
                    const int kap2 = NS - kap;
#pragma unroll
                    for (int r = 0; r < GF; ++r) {
                        U[r] = sm[bufo + phys(kap + r * NS)];
                        V[r] = sm[bufo + phys(kap2 + r * NS)];
                    }
#pragma unroll
                    for (int r = 1; r < GF; ++r) {
                        U[r] = cmul(U[r], sm[WP::OFF_TAB + PL::OFF_FIN + (r - 1) * NS + kap]);
                        V[r] = cmul(V[r], sm[WP::OFF_TAB + PL::OFF_FIN + (r - 1) * NS + kap2]);
                    }
                    SmallFft<GF>::run(U);
                    SmallFft<GF>::run(V);
#pragma unroll
                    for (int a = 0; a < GF; ++a) {
                        const int k = kap + a * NS;
                        epi.pair(k, M - k, U[a], V[GF - 1 - a], sm[WP::OFF_TAB + PL::OFF_POST + k]);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < GF; ++r) {
                        U[r] = sm[bufo + phys(r * NS)];
                        V[r] = sm[bufo + phys(NS / 2 + r * NS)];
                    }
#pragma unroll
                    for (int r = 1; r < GF; ++r)
                        V[r] = cmul(V[r], sm[WP::OFF_TAB + PL::OFF_FIN + (r - 1) * NS + NS / 2]);
                    SmallFft<GF>::run(U);
                    SmallFft<GF>::run(V);
                    epi.dc_nyq(M, U[0]);
#pragma unroll
                    for (int a = 1; 2 * a < GF; ++a)
                        epi.pair(a * NS, M - a * NS, U[a], U[GF - a], sm[WP::OFF_TAB + PL::OFF_POST + a * NS]);
                    if constexpr (GF % 2 == 0) epi.self_mid(M / 2, U[GF / 2]);
#pragma unroll
                    for (int a = 0; 2 * a < GF - 1; ++a) {
                        const int k = NS / 2 + a * NS;
                        epi.pair(k, M - k, V[a], V[GF - 1 - a], sm[WP::OFF_TAB + PL::OFF_POST + k]);
                    }
                    if constexpr (GF % 2 == 1) epi.self_mid(NS / 2 + ((GF - 1) / 2) * NS, V[(GF - 1) / 2]);
                }
            }
            if constexpr (MODE == EPI_BAND) {
                float bs = epi.band;
                epi.band = 0.f;
#pragma unroll
                for (int o = G / 2; o >= 1; o >>= 1) bs += __shfl_xor_sync(0xffffffffu, bs, o);
                if (j == 0 && act) p.out[b * p.out_batch_stride + f] = bs;
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

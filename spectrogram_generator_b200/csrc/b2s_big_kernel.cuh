// Three-pass STFT -> PSD kernel for nperseg 8192 and 16384 with H = 2 columns per thread.
//
// Same arithmetic as stft_psd_kernel (b2s_kernels.cuh): radix-16 Stockham passes over a padded
// shared-memory buffer + fused final stage.  There a frame needs M/16 = 256 / 512 threads, so an
// SM holds 2 / 1 frames whose warps all stop at the same CTA-wide barriers (ncu: 16384 spends
// most of its time there with nothing to overlap).  Here every thread owns two of the M/16
// columns (j and j + GT), i.e. 32 complex points, so a frame takes 128 / 256 threads and an SM
// holds twice as many independent frames at the same number of warps; the window taps are read
// through L1 instead of living in registers, which keeps the kernel at 128 registers.
#pragma once

#include "b2s_kernels.cuh"

namespace b2s {

template <int LOG2N>
struct BigPlan {
    using PL = Plan<LOG2N>;
    static constexpr int H = 2;
    static constexpr int G = PL::G;                      // columns per frame (M / 16)
    static constexpr int GT = G / H;                     // threads per frame
    static constexpr int NT = GT;                        // one frame per CTA
    static constexpr int MINB = 512 / NT;                // 128 registers per thread: 16 warps per SM
    static constexpr int FPC = 1;
    static constexpr int RED = GT / 32;                  // warps per frame
    static constexpr int TPT = (PL::NS / 2) / GT;        // final tasks per thread
    static constexpr size_t SMEM = (size_t)PL::BUF * sizeof(float2) + (size_t)(3 * RED + 1) * sizeof(float);
    static_assert(PL::P == 3 && GT >= 64, "big kernel: nperseg 8192, 16384");
};

template <int LOG2N, typename Tin, int MODE>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(BigPlan<LOG2N>::NT, BigPlan<LOG2N>::MINB) stft_psd_big_kernel(const StftParams p) {
    using PL = Plan<LOG2N>;
    using BP = BigPlan<LOG2N>;
    constexpr int M = PL::M, G = PL::G, NS = PL::NS, GF = PL::GF, H = BP::H, GT = BP::GT;

    B2S_DYN_SMEM(smem_raw);
    const int jt = (int)threadIdx.x;
    float2* const buf = reinterpret_cast<float2*>(smem_raw);
    float* const red = reinterpret_cast<float*>(smem_raw + (size_t)PL::BUF * sizeof(float2));
    const unsigned lane = (unsigned)jt & 31u;
    const float2* const win2 = reinterpret_cast<const float2*>(p.window);

    const int kout = p.kmax - p.kmin + 1;
    Epi<MODE> epi;
    epi.s_edge = p.scale;
    epi.s_int = 2.0f * p.scale;
    epi.floor = p.db_floor;
    epi.kmin = p.kmin;
    epi.kmax = p.kmax;
    epi.db = p.out_mode;

    // sum over the frame of a per-thread partial: xor-butterfly in the warp, fixed order across warps
    auto frame_sum = [&](float tot, float* slot) -> float {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) slot[jt >> 5] = tot;
        __syncthreads();
        tot = 0.f;
#pragma unroll
        for (int w = 0; w < BP::RED; ++w) tot += slot[w];
        return tot;
    };
    auto lane_sum = [&](const float2 (&v)[H][16]) -> float {
        float s[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) s[r] = (v[0][r].x + v[0][r].y) + (v[1][r].x + v[1][r].y);
#pragma unroll
        for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
            for (int r = 0; r < w; ++r) s[r] += s[r + w];
        return s[0];
    };

    // work units: static round-robin over the grid, or (p.work) an atomic counter (see b2s_kernels.cuh)
    const bool dyn = p.work != nullptr;
    auto draw = [&]() -> long long {
        int* const slot = reinterpret_cast<int*>(red + 3 * BP::RED);
        if (jt == 0) *slot = atomicAdd(p.work, 1);
        __syncthreads();
        const int b0 = *slot;
        __syncthreads();
        return (long long)b0;
    };
    long long u_next = dyn ? draw() : (long long)blockIdx.x;
    while (u_next < p.n_units) {
        const long long u = u_next;
        u_next = dyn ? draw() : u + (long long)gridDim.x;
        const long long b = u / p.units_per_signal;
        const int c = (int)(u - b * p.units_per_signal);
        const int f_begin = c * p.chunk_frames;
        const int f_end = (f_begin + p.chunk_frames < p.nframes) ? f_begin + p.chunk_frames : p.nframes;
        const Tin* const xb = reinterpret_cast<const Tin*>(p.x) + b * p.x_batch_stride;
        float* const ob = p.out + b * p.out_batch_stride - ((MODE == EPI_BAND) ? 0 : p.kmin);

        for (int f = f_begin; f < f_end; ++f) {
            const Tin* const xf = xb + (p.frame0 + f) * (long long)p.hop;
            epi.row = ob + (long long)f * kout;

            // ---- gather: z[n] = x[2n] + i x[2n+1], n = j + G r, j = jt + GT h ----
            float2 v[H][16];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const int j = jt + GT * h;
                if (p.vec_ok) {
#pragma unroll
                    for (int r = 0; r < 16; ++r) v[h][r] = Loader<Tin>::ld2(xf + 2 * (j + G * r));
                } else {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const Tin* q = xf + 2 * (j + G * r);
                        v[h][r] = cmk(Loader<Tin>::ld1(q), Loader<Tin>::ld1(q + 1));
                    }
                }
            }

            // ---- detrend (two fp32 passes, see stft_psd_kernel) + window ----
            if (p.detrend) {
                const float m1 = frame_sum(lane_sum(v), red) * (1.0f / (float)PL::N);
#pragma unroll
                for (int h = 0; h < H; ++h)
#pragma unroll
                    for (int r = 0; r < 16; ++r) { v[h][r].x -= m1; v[h][r].y -= m1; }
                const float nr = -frame_sum(lane_sum(v), red + BP::RED) * (1.0f / (float)PL::N);
#pragma unroll
                for (int h = 0; h < H; ++h)
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const float2 w = __ldg(win2 + (jt + GT * h + G * r));
                        v[h][r].x = fmaf(v[h][r].x, w.x, nr * w.x);
                        v[h][r].y = fmaf(v[h][r].y, w.y, nr * w.y);
                    }
            } else {
#pragma unroll
                for (int h = 0; h < H; ++h)
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const float2 w = __ldg(win2 + (jt + GT * h + G * r));
                        v[h][r].x *= w.x;
                        v[h][r].y *= w.y;
                    }
            }

            // ---- pass 0: radix-16 over r (stride G), Ns 1 -> 16 ----
#pragma unroll
            for (int h = 0; h < H; ++h) radix16(v[h]);
            __syncthreads();                    // the previous frame's final-stage reads are done
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float4* dst = reinterpret_cast<float4*>(buf + 18 * (jt + GT * h));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 a = v[h][perm16(2 * i)], bq = v[h][perm16(2 * i + 1)];
                    dst[i] = make_float4(a.x, a.y, bq.x, bq.y);
                }
            }
            __syncthreads();

            // ---- passes 1, 2: radix-16 Stockham, Ns = 16, 256 (all reads before any write) ----
#pragma unroll
            for (int pass = 1; pass < 3; ++pass) {
                const int Ns = (pass == 1) ? 16 : 256;
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const int j = jt + GT * h;
#pragma unroll
                    for (int r = 0; r < 16; ++r) v[h][r] = buf[phys(j + r * G)];
                }
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const int jm = (jt + GT * h) & (Ns - 1);
                    const float2* const twp = p.tw + ((pass == 1) ? PL::OFF_P1 : PL::OFF_P2) + jm;
#pragma unroll
                    for (int r = 1; r < 16; ++r) v[h][r] = cmul(v[h][r], __ldg(twp + (r - 1) * Ns));
                    radix16(v[h]);
                }
                __syncthreads();
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const int j = jt + GT * h;
                    const int jm = j & (Ns - 1);
                    const int base = (j - jm) * 16 + jm;
#pragma unroll
                    for (int r = 0; r < 16; ++r) buf[phys(base + r * Ns)] = v[h][perm16(r)];
                }
                __syncthreads();
            }

            // ---- fused final stage: radix-GF butterflies + real-FFT split + PSD ----
#pragma unroll 1
            for (int cc = 0; cc < BP::TPT; ++cc) {
                const int kap = jt + GT * cc;          // task id == kappa in [0, NS/2)
                float2 U[GF], V[GF];
                if (kap != 0) {
                    const int kap2 = NS - kap;
#pragma unroll
                    for (int r = 0; r < GF; ++r) {
                        U[r] = buf[phys(kap + r * NS)];
                        V[r] = buf[phys(kap2 + r * NS)];
                    }
#pragma unroll
                    for (int r = 1; r < GF; ++r) {
                        U[r] = cmul(U[r], __ldg(p.tw + PL::OFF_FIN + (r - 1) * NS + kap));
                        V[r] = cmul(V[r], __ldg(p.tw + PL::OFF_FIN + (r - 1) * NS + kap2));
                    }
                    SmallFft<GF>::run(U);
                    SmallFft<GF>::run(V);
#pragma unroll
                    for (int a = 0; a < GF; ++a) {
                        const int k = kap + a * NS;
                        epi.pair(k, M - k, U[a], V[GF - 1 - a], __ldg(p.tw + PL::OFF_POST + k));
                    }
                } else {
                    // kappa = 0 and kappa = NS/2 are their own mirrors
#pragma unroll
                    for (int r = 0; r < GF; ++r) {
                        U[r] = buf[phys(r * NS)];
                        V[r] = buf[phys(NS / 2 + r * NS)];
                    }
#pragma unroll
                    for (int r = 1; r < GF; ++r) V[r] = cmul(V[r], __ldg(p.tw + PL::OFF_FIN + (r - 1) * NS + NS / 2));
                    SmallFft<GF>::run(U);
                    SmallFft<GF>::run(V);
                    epi.dc_nyq(M, U[0]);
#pragma unroll
                    for (int a = 1; 2 * a < GF; ++a)
                        epi.pair(a * NS, M - a * NS, U[a], U[GF - a], __ldg(p.tw + PL::OFF_POST + a * NS));
                    if constexpr (GF % 2 == 0) epi.self_mid(M / 2, U[GF / 2]);
#pragma unroll
                    for (int a = 0; 2 * a < GF - 1; ++a) {
                        const int k = NS / 2 + a * NS;
                        epi.pair(k, M - k, V[a], V[GF - 1 - a], __ldg(p.tw + PL::OFF_POST + k));
                    }
                    if constexpr (GF % 2 == 1) epi.self_mid(NS / 2 + ((GF - 1) / 2) * NS, V[(GF - 1) / 2]);
                }
            }
            if constexpr (MODE == EPI_BAND) {
                const float bs = frame_sum(epi.band, red + 2 * BP::RED);
                epi.band = 0.f;
                if (jt == 0) p.out[b * p.out_batch_stride + f] = bs;
            }
        }
    }
    if (dyn) {      // the last CTA to finish re-arms the counters for the next launch that uses them
        __syncthreads();
        if (jt == 0) {
            const int done = atomicAdd(p.work + 1, 1);
            if (done == (int)gridDim.x - 1) {
                p.work[0] = 0;
                p.work[1] = 0;
            }
        }
    }
}

}  // namespace b2s

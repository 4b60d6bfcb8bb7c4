// Fused STFT -> one-sided PSD kernels for sm_100a (power-of-two nperseg).
//
// One launch does everything `_fft_helper` + the PSD epilogue of SciPy's
// `_spectral_helper` do on the CPU (scipy/signal/_spectral_py.py:2374-2395 and
// :2313-2322), i.e. the path the reference takes at PlotEngine.py:113/232:
//
//   frame gather -> per-frame mean removal -> window -> real FFT
//   -> |X|^2 * scale (x2 on interior bins) -> optional 10*log10(max(.,floor))
//   -> [frame][bin] store (optionally cropped to [kmin,kmax]).
//
// Structure (see DESIGN.md "Kernel K-STFT"):
//   * the N real samples of a frame are packed as M = N/2 complex points;
//     a "group" of G = M/16 threads owns a run of consecutive frames of one
//     signal, every thread holding 16 complex points in registers;
//   * M = 16^P * g: P radix-16 Stockham passes (first one straight from the
//     global loads, window/detrend applied in registers) exchanged through a
//     padded shared-memory buffer private to the group, then a fused final
//     stage that does the last radix-g butterflies *and* the real-FFT
//     split (X[k], X[M-k] from Z[k], Z[M-k]) *and* the PSD epilogue, and
//     stores straight to global memory, coalesced over bins;
//   * groups of <= 32 threads synchronise with __syncwarp(group mask) only.
//
// The file compiles for the CPU SIMT emulator (tests/emu) when B2S_EMU is
// defined; that build is test tooling, never a product path.
#pragma once

#include "b2s_fft.cuh"

namespace b2s {

struct StftParams {
    const void* x;               // [batch][n] samples (float or double), device
    long long x_batch_stride;    // elements between signals
    long long n_units;           // work units = batch * units_per_signal
    long long units_per_signal;  // ceil(nframes / chunk_frames)
    long long frame0;            // first (global) frame index of this launch
    long long out_batch_stride;  // elements between signals in `out`
    const float* window;         // [nperseg] fp32 window table, device
    const float2* tw;            // [Plan::TABLE] twiddle tables (see Plan), device
    float* out;                  // [batch][nframes][kmax-kmin+1]
    float* acc;                  // fused cross-sweep sum (stft_psd_duo_sum_kernel): [blocks][nframes][bins] partials
    int acc_rows;                // sweeps per block
    int acc_batch;               // sweeps in the launch
    int* work;                   // dynamic unit scheduling (duo kernels): work[0] = next unit, work[1] = CTAs
                                 // done; both zero at launch, reset by the last CTA.  NULL: static round-robin
    int nframes;                 // frames per signal computed by this launch
    int chunk_frames;            // consecutive frames per work unit
    int hop;
    int detrend;                 // 0 none, 1 constant
    int out_mode;                // 0 linear power, 1 dB
    int kmin, kmax;              // inclusive bin crop
    int vec_ok;                  // 1: frame starts are 2-element aligned (vector loads)
    int ring;                    // pair kernel (b2s_pair_kernel.cuh): samples per warp ring in shared memory
    float scale;                 // 1/(fs*sum(w^2))  or  1/sum(w)^2
    float db_floor;              // linear floor applied before log10 in dB mode
};

template <int LOG2N>
struct Plan {
    static constexpr int N = 1 << LOG2N;
    static constexpr int M = N / 2;
    static constexpr int G = (M / 16 > 0) ? M / 16 : 1;          // threads per frame
    static constexpr int P = (M >= 4096) ? 3 : ((M >= 256) ? 2 : 1);
    static constexpr int NS = (P == 1) ? 16 : ((P == 2) ? 256 : 4096);
    static constexpr int GF = M / NS;                            // final radix 1,2,4,8
    static constexpr int NT = (G > 256) ? G : 256;               // threads per CTA
    static constexpr int FPC = NT / G;                           // groups per CTA
    static constexpr int TPT = (NS / 2) / G;                     // final tasks per thread
    static constexpr int BUF = M + (M >> 4) * 2;                 // padded complex slots
    static constexpr int RED = (G > 32) ? G / 32 : 1;            // mean partials per group
    static constexpr size_t SMEM = (size_t)FPC * BUF * sizeof(float2) + (size_t)FPC * (3 * RED + 1) * sizeof(float);
    // constant tables (complex entries, W = exp(-2 pi i ./.)), laid out so that
    // consecutive lanes read consecutive entries:
    //   P1  [15][16]    W_256^(r jm)           pass 1 (P >= 2)
    //   P2  [15][256]   W_4096^(r jm)          pass 2 (P == 3)
    //   FIN [GF-1][NS]  W_M^(r kappa)          final-stage butterflies
    //   POST[M+1]       W_N^k                  real-FFT split
    static constexpr int OFF_P1 = 0;
    static constexpr int OFF_P2 = OFF_P1 + (P >= 2 ? 15 * 16 : 0);
    static constexpr int OFF_FIN = OFF_P2 + (P == 3 ? 15 * 256 : 0);
    static constexpr int OFF_POST = OFF_FIN + (GF - 1) * NS;
    static constexpr int TABLE = OFF_POST + M + 1;
    static_assert(M >= 16, "nperseg >= 32");
    static_assert(GF == 1 || GF == 2 || GF == 4 || GF == 8, "plan");
};

// padded index inside a group's exchange buffer (16 B pad every 128 B)
B2S_HD int phys(int e) { return e + ((e >> 4) << 1); }

// ---- group synchronisation -------------------------------------------------
template <int G>
B2S_DEVICE void group_sync(unsigned gmask, int grp) {
    if constexpr (G <= 32) {
        __syncwarp(gmask);
    } else {
        b2s_bar_sync(grp + 1, G);      // named barrier private to the group
    }
}

template <typename Tin> struct Loader;
template <> struct Loader<float> {
    B2S_DEVICE static float2 ld2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
    B2S_DEVICE static float ld1(const float* p) { return __ldg(p); }
};
template <> struct Loader<double> {
    B2S_DEVICE static float2 ld2(const double* p) {
        const double2 d = __ldg(reinterpret_cast<const double2*>(p));
        return cmk((float)d.x, (float)d.y);
    }
    B2S_DEVICE static float ld1(const double* p) { return (float)__ldg(p); }
};

// ---- epilogue helpers ---------------------------------------------------------
// MODE 0: linear power, all bins (the reference's call) -- no per-bin branches.
// MODE 1: general -- optional dB, bin crop [kmin, kmax].
// MODE 2: band power -- nothing is stored per bin; the bins in [kmin, kmax] are summed
//         per frame (PlotEngine._calculate_features, PlotEngine.py:238-239).
enum : int { EPI_PLAIN = 0, EPI_GENERAL = 1, EPI_BAND = 2 };
template <int MODE>
struct Epi {
    float* row;        // out + frame row (already offset by -kmin)
    float s_edge;      // scale            (DC / Nyquist)
    float s_int;       // 2*scale          (interior bins)
    float floor;
    float band = 0.f;  // MODE 2: this thread's partial band sum for the current frame
    int kmin, kmax, db;
    bool act = true;   // false: the group is past its last frame (warp kernel), stores are predicated off
    B2S_DEVICE void put(int k, float p) {
        if constexpr (MODE == EPI_GENERAL) {
            if (db) p = 10.0f * log10f(fmaxf(p, floor));
            if (act && k >= kmin && k <= kmax) row[k] = p;
        } else if constexpr (MODE == EPI_BAND) {
            if (k >= kmin && k <= kmax) band += p;
        } else {
            if (act) row[k] = p;
        }
    }
    // Z[k] = zk, Z[M-k] = zm, w = W_N^k : interior bins k and M-k
    B2S_DEVICE void pair(int k, int mk, float2 zk, float2 zm, float2 w) {
        const float2 e = cfma_sign(zm, 1.f, -1.f, zk);                           // 2E = zk + conj(zm)
        const float2 o = cfma_sign(cmk(zk.y, zk.x), 1.f, -1.f, cmk(zm.y, zm.x)); // 2O = -i (zk - conj(zm))
        const float2 t = cmul(o, w);                                             // T = w * 2O
        const float2 a = cadd(e, t);                                             // 2 X[k]
        const float2 bq = csub(e, t);                                            // 2 conj(X[M-k])
        put(k, 0.25f * s_int * fmaf(a.x, a.x, a.y * a.y));
        put(mk, 0.25f * s_int * fmaf(bq.x, bq.x, bq.y * bq.y));
    }
    B2S_DEVICE void self_mid(int k, float2 z) { put(k, s_int * fmaf(z.x, z.x, z.y * z.y)); }
    B2S_DEVICE void dc_nyq(int M, float2 z0) {
        const float a = z0.x + z0.y, b = z0.x - z0.y;
        put(0, s_edge * a * a);
        put(M, s_edge * b * b);
    }
};

// Mean over the N = 32 G samples a group holds (16 complex points per thread).
// Deterministic: fixed pairwise tree per thread, xor-butterfly across lanes, fixed
// order across warps.  `red` (RED floats per group) is only used when G > 32.
template <int LOG2N>
B2S_DEVICE float group_mean(const float2 (&v)[16], unsigned gmask, int grp, int j, unsigned lane, float* red) {
    using PL = Plan<LOG2N>;
    constexpr int G = PL::G;
    float s[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) s[r] = v[r].x + v[r].y;
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
        for (int r = 0; r < w; ++r) s[r] += s[r + w];
    float tot = s[0];
#pragma unroll
    for (int o = (G < 32 ? G : 32) / 2; o >= 1; o >>= 1) tot += __shfl_xor_sync(gmask, tot, o);
    if constexpr (G > 32) {
        if (lane == 0) red[j >> 5] = tot;
        group_sync<G>(gmask, grp);
        tot = 0.f;
#pragma unroll
        for (int w = 0; w < PL::RED; ++w) tot += red[w];
    }
    return tot * (1.0f / (float)PL::N);
}

// ---- the kernel ------------------------------------------------------------------
template <int LOG2N, typename Tin, int MINB, int MODE>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(Plan<LOG2N>::NT, MINB) stft_psd_kernel(const StftParams p) {
    using PL = Plan<LOG2N>;
    constexpr int M = PL::M, G = PL::G, NS = PL::NS, GF = PL::GF;

    B2S_DYN_SMEM(smem_raw);
    const int tid = (int)threadIdx.x;
    const int grp = tid / G;
    const int j = tid - grp * G;
    float2* const buf = reinterpret_cast<float2*>(smem_raw) + (size_t)grp * PL::BUF;
    float* const red = reinterpret_cast<float*>(smem_raw + (size_t)PL::FPC * PL::BUF * sizeof(float2)) + grp * (3 * PL::RED + 1);
    const unsigned lane = (unsigned)tid & 31u;
    const unsigned gmask = (G >= 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(unsigned)(G - 1)));

    // per-thread constants: the window taps this thread always multiplies by
    float2 win[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) win[r] = __ldg(reinterpret_cast<const float2*>(p.window) + (j + G * r));

    const int kout = p.kmax - p.kmin + 1;
    Epi<MODE> epi;
    epi.s_edge = p.scale;
    epi.s_int = 2.0f * p.scale;
    epi.floor = p.db_floor;
    epi.kmin = p.kmin;
    epi.kmax = p.kmax;
    epi.db = p.out_mode;

    // work units: static round-robin over the grid, or (p.work) an atomic counter every group draws
    // from; the next draw is issued a whole unit ahead, so its latency is hidden
    const bool dyn = p.work != nullptr;
    auto draw = [&]() -> long long {
        int b0 = 0;
        if constexpr (G <= 32) {
            if (j == 0) b0 = atomicAdd(p.work, 1);
            b0 = __shfl_sync(gmask, b0, (int)(lane & ~(unsigned)(G - 1)));
        } else {
            int* const slot = reinterpret_cast<int*>(red + 3 * PL::RED);
            if (j == 0) *slot = atomicAdd(p.work, 1);
            group_sync<G>(gmask, grp);
            b0 = *slot;
            group_sync<G>(gmask, grp);
        }
        return (long long)b0;
    };
    long long u_next = dyn ? draw() : (long long)blockIdx.x * PL::FPC + grp;
    while (u_next < p.n_units) {
        const long long u = u_next;
        u_next = dyn ? draw() : u + (long long)gridDim.x * PL::FPC;
        const long long b = u / p.units_per_signal;
        const int c = (int)(u - b * p.units_per_signal);
        const int f_begin = c * p.chunk_frames;
        const int f_end = (f_begin + p.chunk_frames < p.nframes) ? f_begin + p.chunk_frames : p.nframes;
        const Tin* const xb = reinterpret_cast<const Tin*>(p.x) + b * p.x_batch_stride;
        float* const ob = p.out + b * p.out_batch_stride - p.kmin;

        for (int f = f_begin; f < f_end; ++f) {
            const Tin* const xf = xb + (p.frame0 + f) * (long long)p.hop;
            epi.row = ob + (long long)f * kout;

            // ---- gather: z[n] = x[2n] + i x[2n+1], n = j + G r ----
            float2 v[16];
            if (p.vec_ok) {
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = Loader<Tin>::ld2(xf + 2 * (j + G * r));
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const Tin* q = xf + 2 * (j + G * r);
                    v[r] = cmk(Loader<Tin>::ld1(q), Loader<Tin>::ld1(q + 1));
                }
            }

            // ---- detrend='constant': per-frame mean (scipy _signaltools.py:4288-4290) ----
            // Two passes in fp32: a coarse mean m1, then the mean r of the residual
            // x - m1.  The residual is small whatever the DC level of the recording, so
            // r is accurate relative to the signal's own scale; it is removed inside the
            // window multiply with a single rounding: (x' - r) w = fma(x', w, -r w).
            if (p.detrend) {
                const float m1 = group_mean<LOG2N>(v, gmask, grp, j, lane, red);
#pragma unroll
                for (int r = 0; r < 16; ++r) { v[r].x -= m1; v[r].y -= m1; }
                const float nr = -group_mean<LOG2N>(v, gmask, grp, j, lane, red + PL::RED);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    v[r].x = fmaf(v[r].x, win[r].x, nr * win[r].x);
                    v[r].y = fmaf(v[r].y, win[r].y, nr * win[r].y);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) { v[r].x *= win[r].x; v[r].y *= win[r].y; }
            }

            // ---- pass 0: radix-16 over r (stride G), Ns 1 -> 16 ----
            radix16(v);
            group_sync<G>(gmask, grp);          // previous frame's final-stage reads are done
            {
                float4* dst = reinterpret_cast<float4*>(buf + 18 * j);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 a = v[perm16(2 * i)], bq = v[perm16(2 * i + 1)];
                    dst[i] = make_float4(a.x, a.y, bq.x, bq.y);
                }
            }
            group_sync<G>(gmask, grp);

            // ---- passes 1..P-1: radix-16 Stockham, Ns = 16^pass ----
#pragma unroll
            for (int pass = 1; pass < PL::P; ++pass) {
                const int Ns = (pass == 1) ? 16 : 256;
                const int jm = j & (Ns - 1);
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = buf[phys(j + r * G)];
                const float2* const twp = p.tw + ((pass == 1) ? PL::OFF_P1 : PL::OFF_P2) + jm;
#pragma unroll
                for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], __ldg(twp + (r - 1) * Ns));
                radix16(v);
                group_sync<G>(gmask, grp);
                const int base = (j - jm) * 16 + jm;
#pragma unroll
                for (int r = 0; r < 16; ++r) buf[phys(base + r * Ns)] = v[perm16(r)];
                group_sync<G>(gmask, grp);
            }

            // ---- fused final stage: radix-GF butterflies + real-FFT split + PSD ----
#pragma unroll
            for (int cc = 0; cc < PL::TPT; ++cc) {
                const int kap = j + G * cc;            // task id == kappa in [0, NS/2)
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
                // one number per frame: sum the partial band sums of the group (fixed order)
                float bs = epi.band;
                epi.band = 0.f;
#pragma unroll
                for (int o = (G < 32 ? G : 32) / 2; o >= 1; o >>= 1) bs += __shfl_xor_sync(gmask, bs, o);
                if constexpr (G > 32) {
                    if (lane == 0) red[2 * PL::RED + (j >> 5)] = bs;
                    group_sync<G>(gmask, grp);
                    bs = 0.f;
#pragma unroll
                    for (int w = 0; w < PL::RED; ++w) bs += red[2 * PL::RED + w];
                }
                if (j == 0) p.out[b * p.out_batch_stride + f] = bs;
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

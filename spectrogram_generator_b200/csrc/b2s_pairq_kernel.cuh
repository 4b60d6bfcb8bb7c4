// Sub-sequence-pair STFT -> PSD kernel for nperseg 2048 ... 16384: the scheme of b2s_pair_kernel.cuh
// (samples staged in shared memory by the bulk-copy engine, two real sub-sequences per fp32x2
// register, 256-point sub-transforms private to a half-warp) carried past Q = 4 sub-sequences.
//
//   The frame's N real samples are Q = N / 256 real sub-sequences x_r[m] = x[Q m + r].  Half-warp sp
//   (0 <= sp < HW = Q/4) owns r = 4 sp .. 4 sp + 3: z_A = x_{4sp} + i x_{4sp+2} and z_B = x_{4sp+1} +
//   i x_{4sp+3} packed in fp32x2; lane t slot i holds m = t + 16 i, whose four samples are the 16-byte
//   word at sample Q m + 4 sp.  The HW half-warps of a frame read 16 HW contiguous words per slot:
//   conflict-free LDS.128 for every hop that is a multiple of 4.
//   After the two radix-16 passes F^sp[kap] (kap = t + 16 p) goes to shared memory in natural order and
//   the G = 16 HW lanes of the frame meet ONCE.  Task kap (0 <= kap <= 128) then
//     * untangles the Q real spectra from F^sp[kap], F^sp[256 - kap]:
//         (2 X_{4sp}, 2 X_{4sp+1}) = F + conj(F'),  (2 X_{4sp+2}, 2 X_{4sp+3}) = -i (F - conj(F'));
//     * applies Y_r = W_N^(r kap) 2 X_r (packed, twiddle pairs from a table laid out for it);
//     * does the Q-point DFT over r as HW-point DFTs over sp (packed), the twiddles W_Q^(r0 a1) (constant
//       bank) and the 4-point DFT over r0 of b2s_pair_kernel.cuh: outputs a = a1 + HW a0 are the bins
//       kap + 256 a (a < Q/2) and, conjugated, 256 - kap + 256 (Q - 1 - a) (a >= Q/2);
//     * stores |.|^2: for a fixed a, consecutive lanes store consecutive bins.
//   Real-FFT split, last radix and PSD in one step; against the four-step frame-duo kernel one
//   group-wide exchange instead of two, groups of half the threads (no group barrier at all for
//   nperseg 2048: one warp per frame), no register window, no strided global loads.
//
// Staging: every group owns a ring of N samples, stored SUB-SEQUENCE-MAJOR: HW arrays of 256 16-byte words
// (word m of array sp = the four samples x[Q m + 4 sp .. + 3]; float64 samples: two planes of 16-byte halves), each
// padded so that the copies do not collide on banks.  A 1-D bulk copy cannot produce this layout (on the linear
// ring it writes, the 8 words a quarter-warp reads are HW apart: 2- to 4-way bank conflicts on every sample and
// window load -- measured 25 % / 55 % of the shared wavefronts, profiles/r2_pairq_vs_round1.md), so the group's
// threads feed the ring with 16-byte cp.async copies (LDGSTS: coalesced on the global side, permuted on the
// shared side, no registers held) that arrive on an mbarrier: the N samples of a run's first frame, then `hop`
// new samples per frame, issued as soon as every thread of the group has CONSUMED the frame before -- one frame
// ahead of their use.  Lane (sp, t) then reads slot i at word (m0 + t + 16 i) mod 256 of array sp: 16 lanes,
// 256 contiguous bytes.  The window taps sit in shared memory in the same order.  Needs hop % Q == 0.
#pragma once

#include "b2s_pair_kernel.cuh"

namespace b2s {

constexpr int kPairQMaxHW = 16;
// W_Q^(r0 a1) for the twiddle between the HW-point DFTs and the 4-point DFT, laid out for packed use:
// q[2 a1] = (1, Re W^a1, 0, Im W^a1), q[2 a1 + 1] = (Re W^2a1, Re W^3a1, Im W^2a1, Im W^3a1).  Kernel
// parameter: read as constant-bank operands, no loads.
struct PairQConst {
    float4 q[2 * kPairQMaxHW];
};

template <int LOG2N>
struct PairQPlan {
    using PL = Plan<LOG2N>;
    static constexpr int N = PL::N, M = PL::M;
    static constexpr int Q = N / 256;                    // real sub-sequences
    static constexpr int HW = Q / 4;                     // half-warps per frame
    static constexpr int G = 16 * HW;                    // threads per frame (group)
    static constexpr int WG = G / 32;                    // warps per group
    static_assert(HW >= 2 && HW <= kPairQMaxHW, "pairq kernel: nperseg 2048 .. 16384");
    static constexpr int ROW = 17, BUF = 16 * ROW;       // transpose / exchange buffer of one half-warp, float4 units
    static constexpr int NT_MAX = (G <= 128) ? 384 : 256;
    static constexpr int RM = 256;                       // words per sub-sequence array (one frame)
    static constexpr int PADW = (HW == 2) ? 4 : ((HW == 4) ? 2 : 1);     // keeps the 8 words of a quarter-warp's copy on distinct banks
    static constexpr int SUBW = RM + PADW;               // array stride, 16-byte units
    // shared memory, float4 units
    static constexpr int OFF_TW1 = 0;                                    // [8][16] W_256^(t' q)
    static constexpr int OFF_WIN = OFF_TW1 + 8 * 16;                     // [HW][256] window taps, sub-sequence-major
    static constexpr int OFF_BUF = OFF_WIN + N / 4;                      // nt/16 half-warp buffers
    B2S_HD static int off_red(int nt) { return OFF_BUF + (nt / 16) * BUF; }          // per group: 3 x WG partial sums
    B2S_HD static int red_f4() { return (3 * WG + 3) / 4 + 1; }                      // + the unit draw
    B2S_HD static int off_bar(int nt) { return off_red(nt) + (nt / G) * red_f4(); }  // per group: 2 mbarriers
    B2S_HD static int off_ring(int nt) { return off_bar(nt) + (nt / G); }
    B2S_HD static int ring_f4(int esz) { return HW * (esz / 4) * SUBW; }             // 16-byte units per group
    static size_t smem_bytes(int esz, int nt) {
        return (size_t)(off_ring(nt) + (nt / G) * ring_f4(esz)) * sizeof(float4);
    }
    // twiddle table behind Plan::TABLE (float2 units): [2 HW][129] float4
    static constexpr int OFF_PQ = PL::TABLE + (PL::TABLE & 1);           // 16-byte aligned
    static constexpr int TABLE_PQ = OFF_PQ + 129 * 4 * HW;
};

inline bool pairq_kernel_ok(const void* x, long long batch, long long x_batch_stride, int nperseg, int hop) {
    if (nperseg < 2048 || nperseg > 16384) return false;
    if (reinterpret_cast<uintptr_t>(x) % 16) return false;
    if (hop % (nperseg / 256) || hop < 32 || hop > nperseg) return false;      // frames start on a sub-sequence boundary
    if (batch > 1 && (x_batch_stride % 4)) return false;
    return true;
}

template <int HW>
inline void make_pairq_const(PairQConst& c) {
    const int Q = 4 * HW;
    for (int a1 = 0; a1 < HW; ++a1) {
        double wr[4], wi[4];
        for (int r0 = 0; r0 < 4; ++r0) {
            const double ang = -2.0 * 3.14159265358979323846 * (double)((r0 * a1) % Q) / (double)Q;
            wr[r0] = std::cos(ang);
            wi[r0] = std::sin(ang);
        }
        c.q[2 * a1] = float4{(float)wr[0], (float)wr[1], (float)wi[0], (float)wi[1]};
        c.q[2 * a1 + 1] = float4{(float)wr[2], (float)wr[3], (float)wi[2], (float)wi[3]};
    }
}

// natural-order packed DFT of HW points (HW = 2, 4, 8: SmallFft2; 16: the radix-16 butterfly, whose
// output k sits in slot perm16(k))
template <int HW> B2S_HD constexpr int pq_slot(int k) { return (HW == 16) ? perm16(k) : k; }
template <int HW>
B2S_DEVICE void pq_dft(cpx2 (&v)[HW]) {
    if constexpr (HW == 16) c2radix16(v);
    else SmallFft2<HW>::run(v);
}
// 16-byte asynchronous copy global -> shared (LDGSTS) and "my copies so far have landed" on an mbarrier
#ifdef B2S_EMU
B2S_DEVICE void ring_cp16(void* dst, const void* src) { std::memcpy(dst, src, 16); }
B2S_DEVICE void ring_cp_arrive(void*) {}
#else
B2S_DEVICE void ring_cp16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
B2S_DEVICE void ring_cp_arrive(void* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#endif

// (a.re + i a.im) * (w.re + i w.im), element-wise on the two packed halves
B2S_HD cpx2 c2mul_vec(cpx2 a, float2 wr, float2 wi) {
    return cpx2{pk_fma(a.re, wr, pk_neg(pk_mul(a.im, wi))), pk_fma(a.re, wi, pk_mul(a.im, wr))};
}

template <int LOG2N, typename Tin, int MODE>
B2S_DEVICE void stft_psd_pairq_body(const StftParams& p, const PairQConst& qc) {
    using PL = Plan<LOG2N>;
    using PP = PairQPlan<LOG2N>;
    constexpr int N = PP::N, M = PP::M, HW = PP::HW, G = PP::G, WG = PP::WG, ROW = PP::ROW, Q = PP::Q;
    constexpr int ESZ = (int)sizeof(Tin);

    B2S_DYN_SMEM_F4(sm4);
    const int tid = (int)threadIdx.x;
    const int nt = (int)blockDim.x;
    const int ngroups = nt / G;
    const int grp = tid / G;
    const int j = tid - grp * G;                         // thread inside the frame's group
    // Half-warp sp owns sub-sequences 4 sp .. 4 sp + 3.  (A shared-memory LDS.128 is served a quarter-warp at a
    // time and the words of a quarter are HW apart on the linear ring: 2-way bank conflicts on the sample and
    // window loads at HW = 2, 4-way above -- measured 25 % / 55 % of the shared wavefronts,
    // profiles/r2_pairq_*.ncu_summary.txt.  Interleaving the two sub-transforms of a warp removes them for
    // HW = 2 but moves the conflicts to the transposes; a conflict-free layout needs the ring stored
    // sub-sequence-major, i.e. 16-byte cp.async copies instead of one bulk copy.  See DESIGN.md.)
    const int sp = j >> 4;
    const int t = j & 15;
    const int lane = tid & 31;
    float4* const buf = sm4 + PP::OFF_BUF + (grp * HW + sp) * PP::BUF;            // this sub-transform's
    float4* const xb = sm4 + PP::OFF_BUF + (grp * HW) * PP::BUF;                  // the group's HW buffers
    float* const red = reinterpret_cast<float*>(sm4 + PP::off_red(nt) + grp * PP::red_f4());
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(sm4 + PP::off_bar(nt) + grp);
    constexpr int NP = ESZ / 4;                          // 16-byte pieces per word of four samples (1 float, 2 double)
    constexpr int SPP = 16 / ESZ;                        // samples per piece
    float4* const ring = sm4 + PP::off_ring(nt) + grp * PP::ring_f4(ESZ);        // [HW * NP][SUBW] pieces
    auto gsync = [&]() {
        if constexpr (G <= 32) __syncwarp();
        else b2s_bar_sync(grp + 1, G);
    };

    // ---- constant tables, once per CTA ----
    const float csc = sqrtf(0.5f * p.scale);             // the PSD scale goes into the window taps
    {
        for (int i = tid; i < 8 * 16; i += nt) {
            const int jj = i >> 4, l = i & 15;
            const float2 ta = (jj == 0) ? cmk(1.f, 0.f) : __ldg(p.tw + PL::OFF_P1 + (2 * jj - 1) * 16 + l);
            const float2 tb = __ldg(p.tw + PL::OFF_P1 + (2 * jj) * 16 + l);
            sm4[PP::OFF_TW1 + i] = make_float4(ta.x, ta.y, tb.x, tb.y);
        }
        {
            // taps in the order the lanes read them: [sp][m] = w[Q m + 4 sp .. + 3]
            const float4* w4 = reinterpret_cast<const float4*>(p.window);
            for (int i = tid; i < N / 4; i += nt) {
                const float4 w = __ldg(w4 + i);                       // word i: sp = i % HW, m = i / HW
                sm4[PP::OFF_WIN + (i % HW) * 256 + i / HW] = make_float4(w.x * csc, w.y * csc, w.z * csc, w.w * csc);
            }
        }
        if (j == 0) {
            ring_bar_init(bars, G);              // every thread of the group arrives once per chunk
            ring_bar_init(bars + 1, G);
        }
        ring_fence_init();
    }
    __syncthreads();
    const float4* const pq = reinterpret_cast<const float4*>(p.tw + PP::OFF_PQ);

    const int kout = p.kmax - p.kmin + 1;
    EpiOne<MODE> epi;
    epi.floor = p.db_floor;
    epi.kmin = p.kmin;
    epi.kmax = p.kmax;
    epi.db = p.out_mode;
    epi.band = 0.f;
    epi.act = true;

    // sum of one float per thread over the group, fixed order: xor-butterfly inside the warp, then the
    // per-warp partials in warp order (three slots -- coarse mean, residual mean, band power -- so that a
    // group-wide barrier always lies between two uses of the same slot)
    auto group_sum = [&](float v, int slot) -> float {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if constexpr (WG > 1) {
            if (lane == 0) red[slot * WG + (j >> 5)] = v;
            gsync();
            v = red[slot * WG];
#pragma unroll
            for (int w = 1; w < WG; ++w) v += red[slot * WG + w];
        }
        return v;
    };

    // ---- work units: runs of consecutive frames of one signal, one run per group at a time ----
    const bool dyn = p.work != nullptr;
    auto draw = [&]() -> long long {
        int b0 = 0;
        if constexpr (G <= 32) {
            if (lane == 0) b0 = atomicAdd(p.work, 1);
            b0 = __shfl_sync(0xffffffffu, b0, 0);
        } else {
            int* const slot = reinterpret_cast<int*>(red + 4 * ((3 * WG + 3) / 4));
            if (j == 0) *slot = atomicAdd(p.work, 1);
            gsync();
            b0 = *slot;
            gsync();
        }
        return (long long)b0;
    };
    struct Unit {
        const Tin* x;
        float* out;
        int nf;
    };
    auto unit_of = [&](long long u) -> Unit {
        const long long b = u / p.units_per_signal;
        const int c = (int)(u - b * p.units_per_signal);
        const int f_begin = c * p.chunk_frames;
        const int f_end = (f_begin + p.chunk_frames < p.nframes) ? f_begin + p.chunk_frames : p.nframes;
        Unit r;
        r.x = reinterpret_cast<const Tin*>(p.x) + b * p.x_batch_stride + (p.frame0 + f_begin) * (long long)p.hop;
        r.out = p.out + b * p.out_batch_stride +
                ((MODE == EPI_BAND) ? (long long)f_begin : (long long)f_begin * kout - p.kmin);
        r.nf = f_end - f_begin;
        return r;
    };
    const int hop = p.hop;
    // chunk jc of a run: the N samples of its first frame, then `hop` samples per further frame.  Piece u (16
    // bytes) of the run is piece u % NP of word w = u / NP, i.e. of sub-sequence array sp = w % HW at m = w / HW.
    unsigned issued = 0, waited = 0;
    auto issue = [&](const Unit& un, int jc) {
        void* const bar = bars + (issued & 1u);
        const int lo = jc == 0 ? 0 : N + (jc - 1) * hop;                 // samples
        const int len = jc == 0 ? N : hop;
        const int u1 = (lo + len) / SPP;
        const unsigned char* const src = reinterpret_cast<const unsigned char*>(un.x);
        for (int u = lo / SPP + j; u < u1; u += G) {
            const int w = u / NP, pc = u - w * NP;
            const int s_ = w % HW, m = (w / HW) & (PP::RM - 1);
            ring_cp16(ring + (s_ * NP + pc) * PP::SUBW + m, src + (size_t)u * 16);
        }
        ring_cp_arrive(bar);
        ++issued;
    };
    auto wait_chunk = [&]() {
#ifdef B2S_EMU
        gsync();
#endif
        ring_wait(bars + (waited & 1u), (waited >> 1) & 1u);
        ++waited;
    };

    const long long ustride = (long long)gridDim.x * ngroups;
    long long u_cur = dyn ? draw() : (long long)blockIdx.x * ngroups + grp;
    long long u_next = 0;
    Unit un{}, unn{};
    if (u_cur < p.n_units) {
        un = unit_of(u_cur);
        issue(un, 0);
        u_next = dyn ? draw() : u_cur + ustride;
    }
    while (u_cur < p.n_units) {
        const bool have_next = u_next < p.n_units;
        if (have_next) unn = unit_of(u_next);
        int pos = 0;                                     // ring position (words of a sub-sequence array) of the current frame
        for (int it = 0; it < un.nf; ++it) {
            epi.row = un.out + (long long)it * kout;
            wait_chunk();

            // ---- the frame's samples: slot i = x[Q (t + 16 i) + 4 sp .. + 3] ----
            // word m of sub-sequence array sp; the frame starts at m0 = (it hop / Q) mod 256
            const int m0 = pos + t;
            auto slot_word = [&](int i) -> int { return (m0 + 16 * i) & (PP::RM - 1); };
            auto ld4 = [&](int i) -> float4 {
                if constexpr (sizeof(Tin) == 4) {
                    return ring[sp * PP::SUBW + slot_word(i)];
                } else {
                    const double2 a = reinterpret_cast<const double2*>(ring)[(2 * sp) * PP::SUBW + slot_word(i)];
                    const double2 bq = reinterpret_cast<const double2*>(ring)[(2 * sp + 1) * PP::SUBW + slot_word(i)];
                    return make_float4((float)a.x, (float)a.y, (float)bq.x, (float)bq.y);
                }
            };
            cpx2 v[16];
            bool need_sync = true;
            float c = 0.f;
            if constexpr (sizeof(Tin) == 8) {
                double cd = 0.0;
                if (p.detrend) {
                    float s[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 q = ld4(i);
                        s[i] = (q.x + q.y) + (q.z + q.w);
                        if ((i & 3) == 3) B2S_SCHED_FENCE();
                    }
#pragma unroll
                    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                        for (int i = 0; i < w; ++i) s[i] += s[i + w];
                    cd = (double)(group_sum(s[0], 0) * (1.0f / (float)N));
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const double2 a = reinterpret_cast<const double2*>(ring)[(2 * sp) * PP::SUBW + slot_word(i)];
                    const double2 bq = reinterpret_cast<const double2*>(ring)[(2 * sp + 1) * PP::SUBW + slot_word(i)];
                    const float4 q = make_float4((float)(a.x - cd), (float)(a.y - cd), (float)(bq.x - cd), (float)(bq.y - cd));
                    v[i].re = cmk(q.x, q.y);
                    v[i].im = cmk(q.z, q.w);
                    if ((i & 3) == 3) B2S_SCHED_FENCE();
                }
            } else {
                float4 raw[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) raw[i] = ld4(i);
                if (p.detrend) {
                    float s[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) s[i] = (raw[i].x + raw[i].y) + (raw[i].z + raw[i].w);
#pragma unroll
                    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                        for (int i = 0; i < w; ++i) s[i] += s[i + w];
                    c = group_sum(s[0], 0) * (1.0f / (float)N);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        v[i].re = pk_add(cmk(raw[i].x, raw[i].y), cmk(-c, -c));
                        v[i].im = pk_add(cmk(raw[i].z, raw[i].w), cmk(-c, -c));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        v[i].re = cmk(raw[i].x, raw[i].y);
                        v[i].im = cmk(raw[i].z, raw[i].w);
                    }
                }
            }
            auto taps = [&](int i) -> float4 { return sm4[PP::OFF_WIN + sp * 256 + t + 16 * i]; };
            if (p.detrend) {
                float2 sr[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) sr[i] = pk_add(v[i].re, v[i].im);
#pragma unroll
                for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                    for (int i = 0; i < w; ++i) sr[i] = pk_add(sr[i], sr[i + w]);
                const float nr = group_sum(sr[0].x + sr[0].y, 1) * (-1.0f / (float)N);
                if constexpr (WG > 1) need_sync = false;  // that sum's barrier came after every thread had consumed its samples
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 w = taps(i);
                    v[i].re = pk_fma(v[i].re, cmk(w.x, w.y), pk_muls(cmk(w.x, w.y), nr));     // (x' - r) w, one rounding
                    v[i].im = pk_fma(v[i].im, cmk(w.z, w.w), pk_muls(cmk(w.z, w.w), nr));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 w = taps(i);
                    v[i].re = pk_mul(v[i].re, cmk(w.x, w.y));
                    v[i].im = pk_mul(v[i].im, cmk(w.z, w.w));
                }
            }
            // Every thread of the group has CONSUMED its samples (a barrier alone does not wait for shared-memory
            // loads that are still in flight -- measured: frames built from half-overwritten samples when the copy
            // below was issued right after the loads): the first `hop` positions of this frame are free, feed
            // the next frame's samples (or the next run's first frame) into them.
            if (need_sync) gsync();
            if (it + 1 < un.nf) issue(un, it + 1);
            else if (have_next) issue(unn, 0);

            // ---- sub-transforms: radix-16, 16 x 16 transpose inside the half-warp, radix-16 ----
            c2radix16(v);
            // (the group-wide barrier after the sample reads also separates the previous frame's final-stage
            //  reads of these buffers from the stores below)
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const cpx2 z = v[perm16(q)];
                buf[ROW * t + q] = make_float4(z.re.x, z.re.y, z.im.x, z.im.y);
            }
            __syncwarp();
#pragma unroll
            for (int tt = 0; tt < 16; ++tt) {
                const float4 q4 = buf[ROW * tt + t];
                v[tt] = cpx2{cmk(q4.x, q4.y), cmk(q4.z, q4.w)};
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float4 w = sm4[PP::OFF_TW1 + jj * 16 + t];
                if (jj > 0) v[2 * jj] = c2mul(v[2 * jj], cmk(w.x, w.y));
                v[2 * jj + 1] = c2mul(v[2 * jj + 1], cmk(w.z, w.w));
            }
            c2radix16(v);
            __syncwarp();                        // every lane of the half-warp has consumed its transpose reads
            // F^sp[kap], kap = t + 16 pp, natural order in the half-warp's own buffer
#pragma unroll
            for (int pp = 0; pp < 16; ++pp) {
                const cpx2 z = v[perm16(pp)];
                buf[t + 16 * pp] = make_float4(z.re.x, z.re.y, z.im.x, z.im.y);
            }
            gsync();

            // ---- fused final stage: task kap ----
            auto task = [&](int kap, bool dc, bool mid) {
                const int kapm = (256 - kap) & 255;
                cpx2 U[HW], V[HW];
                const float4* const tw = pq + kap;         // [2 HW][129]: consecutive lanes read consecutive entries
#pragma unroll
                for (int s = 0; s < HW; ++s) {
                    const float4 f4 = xb[s * PP::BUF + kap], g4 = xb[s * PP::BUF + kapm];
                    const cpx2 F{cmk(f4.x, f4.y), cmk(f4.z, f4.w)}, Gm{cmk(g4.x, g4.y), cmk(g4.z, g4.w)};
                    const cpx2 u{pk_add(F.re, Gm.re), pk_sub(F.im, Gm.im)};       // (2 X_4s, 2 X_4s+1) = F + conj(F')
                    const cpx2 w{pk_add(F.im, Gm.im), pk_sub(Gm.re, F.re)};       // (2 X_4s+2, 2 X_4s+3) = -i (F - conj(F'))
                    const float4 tu = __ldg(tw + (2 * s) * 129), tv = __ldg(tw + (2 * s + 1) * 129);
                    U[s] = c2mul_vec(u, cmk(tu.x, tu.y), cmk(tu.z, tu.w));        // Y_r = W_N^(r kap) 2 X_r
                    V[s] = c2mul_vec(w, cmk(tv.x, tv.y), cmk(tv.z, tv.w));
                }
                pq_dft<HW>(U);                   // over sp: index a1
                pq_dft<HW>(V);
#pragma unroll
                for (int a1 = 0; a1 < HW; ++a1) {
                    cpx2 u = U[pq_slot<HW>(a1)], w = V[pq_slot<HW>(a1)];
                    if (a1 > 0) {
                        const float4 cu = qc.q[2 * a1], cv = qc.q[2 * a1 + 1];
                        u = c2mul_vec(u, cmk(cu.x, cu.y), cmk(cu.z, cu.w));
                        w = c2mul_vec(w, cmk(cv.x, cv.y), cmk(cv.z, cv.w));
                    }
                    const float2 sr = pk_add(u.re, w.re), si = pk_add(u.im, w.im);
                    const float2 dr = pk_sub(u.re, w.re), di = pk_sub(u.im, w.im);
                    const float2 pr = cmk(sr.x + sr.y, sr.x - sr.y), pi = cmk(si.x + si.y, si.x - si.y);   // a0 = 0, 2
                    const float2 qr = cmk(dr.x + di.y, dr.x - di.y), qi = cmk(di.x - dr.y, di.x + dr.y);   // a0 = 1, 3
                    float2 ps = pk_fma(pr, pr, pk_mul(pi, pi));
                    const float2 pd = pk_fma(qr, qr, pk_mul(qi, qi));
                    if (dc && a1 == 0) ps = pk_muls(ps, 0.5f);              // DC / Nyquist carry scale, not 2 scale
                    epi.put(kap + 256 * a1, ps.x);
                    epi.put(kap + 256 * (a1 + HW), pd.x);
                    if (!mid && !dc) {
                        epi.put((256 - kap) + 256 * (2 * HW - 1 - a1), ps.y);
                        epi.put((256 - kap) + 256 * (HW - 1 - a1), pd.y);
                    } else if (dc && a1 == 0) {
                        epi.put(M, ps.y);        // Nyquist; every other mirrored output of kap = 0 is a duplicate
                    }
                }
            };
            constexpr int TPL = (G <= 128) ? 128 / G : 1;
#pragma unroll
            for (int cc = 0; cc < TPL; ++cc) {
                const int kap = j + G * cc;
                if (G <= 128 || kap < 128) task(kap, kap == 0, false);
            }
            if (j == G - 1) task(128, false, true);       // kap = 128 is its own mirror
            if constexpr (MODE == EPI_BAND) {
                const float bs = group_sum(epi.band, 2);
                epi.band = 0.f;
                if (j == 0) un.out[it] = bs;
            }
            pos = (pos + hop / Q) & (PP::RM - 1);
        }
        u_cur = u_next;
        un = unn;
        if (have_next) u_next = dyn ? draw() : u_next + ustride;
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

// one CTA of up to 384 threads (nperseg <= 8192: 12 warps) or 256 threads (16384: one frame) per SM
template <int LOG2N, typename Tin, int MODE>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(PairQPlan<LOG2N>::NT_MAX, 1) stft_psd_pairq_kernel(const StftParams p, const PairQConst qc) {
    stft_psd_pairq_body<LOG2N, Tin, MODE>(p, qc);
}

}  // namespace b2s

// Sub-sequence-pair STFT -> PSD kernel for nperseg 1024 (any hop that is a multiple of 4
// samples), samples staged through shared memory by the bulk-copy engine (TMA).
//
// The frame-duo kernels pack two FRAMES in fp32x2 registers and keep the raw samples of the
// overlapping frames in a sliding register window; at nperseg >= 1024 that window (40 registers),
// the strided sample loads and the group-wide exchanges make them LSU-bound
// (profiles/r1_c3_duo4_v11.ncu_summary.txt).  This kernel packs two SUB-TRANSFORMS of ONE frame:
//
//   the frame's N = 1024 real samples are four real sub-sequences x_r[m] = x[4 m + r] of 256
//   points.  z_A = x_0 + i x_2 and z_B = x_1 + i x_3 are transformed together, A in the low and B
//   in the high halves of fp32x2 registers: both use the SAME pass-0 / pass-1 twiddles (broadcast
//   scalars, as in the frame-duo kernels), and the four samples x[4 m .. 4 m + 3] of a point are
//   ONE 16-byte word whose register quad is already (re_A, re_B | im_A, im_B) -- no shuffling of
//   registers, window taps in natural order.  16 lanes own a frame, lane t slot i holds
//   m = t + 16 i; the 16 lanes read 256 contiguous bytes: conflict-free LDS.128 straight from the
//   staged samples, with any hop.
//   * one 16 x 16 transpose per frame through a buffer private to the half-warp (__syncwarp only);
//   * the mirror F[256 - kap] comes from lane (16 - t) & 15 with 32 shuffles; with it the fused
//     final stage untangles the four real spectra (2 X_0, 2 X_1 = F + conj(F'), 2 X_2, 2 X_3 =
//     -i (F - conj(F'))), applies W_1024^(r kap) and one radix-4 butterfly gives the bins kap,
//     kap + 256, 256 - kap and 512 - kap -- real-FFT split and last radix in one step;
//   * no group-wide barrier anywhere: the two frames of a warp only share instructions.
//
// Sample staging (north-star kernel 1): every warp owns a ring of N + hop samples in shared
// memory.  Lane 0 feeds it with cp.async.bulk (global -> shared, completion on an mbarrier):
// the first N + hop samples of a run of frames, then 2 hop new samples per pair of frames,
// issued as soon as the pair before has read its samples -- a whole iteration ahead of their
// use, with no registers and no LSU instructions spent on global loads.  Every sample of a run
// crosses L2 -> SM once.  The ring of the next run is primed during the last iteration of the
// current one (units are drawn one ahead).
//
// Detrend: coarse mean, then the mean of the residual, in a fixed frame-relative order (the
// result of a frame depends on its samples only).  float64 samples are staged as float64 and the
// coarse mean is subtracted in double before the cast, so a recording that sits on a large DC
// level keeps its small signal (SciPy detrends float64 input in float64).
//
// SUM mode (stft_psd_pair_sum_kernel; per-sweep spectrograms AND their cross-sweep sum in one pass,
// SURVEY.md 8 a-15, for nperseg 1024 at any staged hop): the same body walked in the other
// direction.  A warp keeps ONE pair of frames (2 c, 2 c + 1) and walks over a block of consecutive
// sweeps; the ring then holds the N + hop samples of that pair of one sweep, refilled by one bulk
// copy per sweep, issued as soon as the sweep before has been consumed.  Every per-sweep row is
// stored exactly as the per-sweep kernel stores it (same arithmetic, bit-identical), and the 33
// power values a lane produces per frame are added, in sweep order, to running sums that never
// leave the SM: tensor memory (SUM = 2, the product path: 34 columns of the lane's own TMEM lane)
// or shared memory (SUM = 1: the CPU emulator's path and the twin the launch takes its residency
// from).  A unit ends by writing its block's partial sums; batch_sum_kernel folds the blocks.
// The order of the additions is fixed (sweep order in a block, block order in the fold).
#pragma once

#include "b2s_duo_cta_kernel.cuh"
#include "b2s_tmem.cuh"

namespace b2s {

template <int LOG2N>
struct PairPlan {
    using PL = Plan<LOG2N>;
    static constexpr int N = PL::N, M = PL::M;
    static constexpr int R = M / 256;
    static_assert(R == 2, "pair kernel: nperseg 1024 (four real sub-sequences of 256 points)");
    // CTA shapes: 128 threads x 3 CTAs per SM (<= 168 registers), or ONE CTA of up to 480 threads per SM
    // (<= 136 registers): nothing in the main loop is CTA-wide, so the CTA is only the unit that shares
    // the constant tables.  `nt` below is the launch's block size.
    static constexpr int NT = 128;
    static constexpr int NT_MID = 384;
    static constexpr int NT_WIDE = 512;
    static constexpr int ROW = 17, BUF = 16 * ROW;       // transpose buffer of one frame, float4 units
    // shared memory, float4 units
    static constexpr int OFF_WIN = 0;                    // [16][16] taps of slot i, lane t: w[4 (t + 16 i) .. + 3] * sqrt(scale/2)
    static constexpr int OFF_TW1 = OFF_WIN + 16 * 16;    // [8][16] W_256^(t' q), t' = 2j, 2j+1
    static constexpr int OFF_W12 = OFF_TW1 + 8 * 16;     // [8][16] (W_1024^kap, W_1024^(2 kap)), kap = t + 16 pp
    static constexpr int OFF_W3 = OFF_W12 + 8 * 16;      // [8][16] float2 W_1024^(3 kap)  (half a float4 each)
    static constexpr int OFF_BUF = OFF_W3 + 4 * 16;      // nt/16 transpose buffers
    // the wide CTA shape (nt > NT_MID) transposes real and imaginary parts one after the other through a
    // half-size buffer: 12 KB of shared memory per warp instead of 16, i.e. 16 warps per SM where 13 fit
    B2S_HD static int buf_f4(int nt) { return nt > NT_MID ? BUF / 2 : BUF; }
    B2S_HD static int off_bar(int nt) { return OFF_BUF + (nt / 16) * buf_f4(nt); }   // nt/32 x 2 mbarriers (one float4 per warp)
    B2S_HD static int off_ring(int nt) { return off_bar(nt) + nt / 32; }     // nt/32 rings of ring_samples elements each
    // ring: the N + hop samples of a pair of frames, rounded up to 128 bytes
    B2S_HD static int ring_samples(int hop) { return (N + hop + 31) / 32 * 32; }
    B2S_HD static size_t smem_bytes(int hop, int esz, int nt = NT) {
        return (size_t)off_ring(nt) * sizeof(float4) + (size_t)(nt / 32) * ring_samples(hop) * esz;
    }
    // SUM mode (128-thread CTAs): the running sums of the shared-memory twin, [9][NT] float4, behind the rings
    static constexpr int ACC_SLOTS = 9;                  // 8 tasks x (kap, kap + 256, 512 - kap, 256 - kap) + (128, 384)
    static constexpr int TMEM_COLS = 64;                 // 34 used
    static size_t sum_smem_bytes(int hop, int esz) {
        return smem_bytes(hop, esz, NT) + (size_t)ACC_SLOTS * NT * sizeof(float4);
    }
};

// hop must keep every frame start on a 16-byte boundary of the row
inline bool pair_kernel_ok(const void* x, int x_is_f64, long long batch, long long x_batch_stride, int nperseg, int hop,
                           long long frame0) {
    (void)frame0;
    if (nperseg != 1024) return false;
    (void)x_is_f64;              // the same rule for float and double rows: both take the same kernel
    if (reinterpret_cast<uintptr_t>(x) % 16) return false;
    if (hop % 4 || hop < 32 || hop > nperseg) return false;
    if (batch > 1 && (x_batch_stride % 4)) return false;
    return true;
}

#ifdef B2S_EMU
#define B2S_MAXNREG(n)
#define B2S_SCHED_FENCE() do {} while (0)
#else
#define B2S_MAXNREG(n) __maxnreg__(n)
// keeps the compiler from hoisting every load of an unrolled loop above their uses (register pressure)
#define B2S_SCHED_FENCE() asm volatile("" ::: "memory")
#endif

// ---- bulk-copy / mbarrier primitives -----------------------------------------------------------
#ifdef B2S_EMU
struct RingBar { int dummy; };
B2S_DEVICE void ring_bar_init(void*, int) {}
B2S_DEVICE void ring_fence_init() {}
B2S_DEVICE void ring_expect(void*, unsigned) {}
B2S_DEVICE void ring_copy(void* dst, const void* src, unsigned bytes, void*) { std::memcpy(dst, src, bytes); }
B2S_DEVICE void ring_wait(void*, unsigned) {}
#else
B2S_DEVICE unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
B2S_DEVICE void ring_bar_init(void* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
B2S_DEVICE void ring_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
B2S_DEVICE void ring_expect(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (TMA, 1-D): 16-byte aligned addresses, size a multiple of 16
B2S_DEVICE void ring_copy(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
B2S_DEVICE void ring_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif

// four consecutive samples of the ring as floats; `idx16` counts 16-byte words of float samples
template <typename Tin> struct RingLoad;
template <> struct RingLoad<float> {
    B2S_DEVICE static float4 ld4(const void* ring, int word) { return reinterpret_cast<const float4*>(ring)[word]; }
};
template <> struct RingLoad<double> {
    B2S_DEVICE static float4 ld4(const void* ring, int word) {
        const double2 a = reinterpret_cast<const double2*>(ring)[2 * word];
        const double2 b = reinterpret_cast<const double2*>(ring)[2 * word + 1];
        return make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
    }
    // (x - pivot) in double, rounded once
    B2S_DEVICE static float4 ld4_minus(const void* ring, int word, double pivot) {
        const double2 a = reinterpret_cast<const double2*>(ring)[2 * word];
        const double2 b = reinterpret_cast<const double2*>(ring)[2 * word + 1];
        return make_float4((float)(a.x - pivot), (float)(a.y - pivot), (float)(b.x - pivot), (float)(b.y - pivot));
    }
};

template <int MODE>
struct EpiOne {
    float* row;         // frame row (already offset by -kmin)
    float floor;
    float band;
    int kmin, kmax, db;
    bool act;
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
};

template <int LOG2N, typename Tin, int MODE, int MINB, int SUM = 0>
B2S_DEVICE void stft_psd_pair_body(const StftParams& p);

// 128 threads x 3 CTAs per SM
template <int LOG2N, typename Tin, int MODE>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(PairPlan<LOG2N>::NT, 3) stft_psd_pair_kernel(const StftParams p) {
    stft_psd_pair_body<LOG2N, Tin, MODE, 3>(p);
}
// one CTA of up to 512 threads per SM: 128 registers (registers are handed out to four warps at a time,
// so 15 warps of 136 do not fit)
template <int LOG2N, typename Tin, int MODE>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(512, 1) stft_psd_pair_wide_kernel(const StftParams p) {
    stft_psd_pair_body<LOG2N, Tin, MODE, 1>(p);
}
// one CTA of up to 384 threads per SM at 136 registers: twelve warps where three 128-thread CTAs no longer
// fit their rings (hop > 768).  Measured against 192 threads x 2 CTAs at 168 registers: 0.75 vs 0.69 of the
// HBM peak at hop 896, 0.81 vs 0.75 at hop 1024.
template <int LOG2N, typename Tin, int MODE>
B2S_GLOBAL void B2S_MAXNREG(136) stft_psd_pair_mid_kernel(const StftParams p) {
    stft_psd_pair_body<LOG2N, Tin, MODE, 2>(p);
}

// per-sweep rows + cross-sweep block sums (SUM = 1: sums in shared memory, 2: in tensor memory)
template <int LOG2N, typename Tin, int SUM>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(PairPlan<LOG2N>::NT, 3) stft_psd_pair_sum_kernel(const StftParams p) {
    stft_psd_pair_body<LOG2N, Tin, EPI_PLAIN, 3, SUM>(p);
}

template <int LOG2N, typename Tin, int MODE, int MINB, int SUM>
B2S_DEVICE void stft_psd_pair_body(const StftParams& p) {
    static_assert(SUM == 0 || (MODE == EPI_PLAIN && MINB == 3), "SUM mode: plain epilogue, 128-thread CTAs");
    using PL = Plan<LOG2N>;
    using PP = PairPlan<LOG2N>;
    constexpr int N = PP::N, M = PP::M, ROW = PP::ROW;
    constexpr int ESZ = (int)sizeof(Tin);

    B2S_DYN_SMEM_F4(sm4);
    const int tid = (int)threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int h = lane >> 4;                             // frame of the pair this half-warp owns
    const int t = lane & 15;
    constexpr bool HALF = (MINB == 1);                   // wide CTA: half-size transpose buffer
    float4* const buf = sm4 + PP::OFF_BUF + (tid >> 4) * (HALF ? PP::BUF / 2 : PP::BUF);
    const int nt = (int)blockDim.x;
    const int nwarps = nt >> 5;
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(sm4 + PP::off_bar(nt) + warp);
    const int RS = p.ring;                               // ring length in samples
    unsigned char* const ring = reinterpret_cast<unsigned char*>(sm4 + PP::off_ring(nt)) + (size_t)warp * RS * ESZ;

    // ---- constant tables, once per CTA; the PSD scale goes into the window ----
    {
        const float csc = sqrtf(0.5f * p.scale);
        const float4* w4 = reinterpret_cast<const float4*>(p.window);
        for (int i = tid; i < 16 * 16; i += nt) {
            const float4 w = __ldg(w4 + i);             // taps of samples 4 (t + 16 i) .. + 3, i * 16 + t == i
            sm4[PP::OFF_WIN + i] = make_float4(w.x * csc, w.y * csc, w.z * csc, w.w * csc);
        }
        float2* const w3tab = reinterpret_cast<float2*>(sm4 + PP::OFF_W3);
        for (int i = tid; i < 8 * 16; i += nt) {
            const int j = i >> 4, l = i & 15;
            const float2 ta = (j == 0) ? cmk(1.f, 0.f) : __ldg(p.tw + PL::OFF_P1 + (2 * j - 1) * 16 + l);
            const float2 tb = __ldg(p.tw + PL::OFF_P1 + (2 * j) * 16 + l);
            sm4[PP::OFF_TW1 + i] = make_float4(ta.x, ta.y, tb.x, tb.y);
            const int kap = l + 16 * j;                 // final stage, task pp = j of lane l
            const float2 w1 = __ldg(p.tw + PL::OFF_POST + kap), w2 = __ldg(p.tw + PL::OFF_POST + 2 * kap);
            sm4[PP::OFF_W12 + i] = make_float4(w1.x, w1.y, w2.x, w2.y);
            w3tab[i] = __ldg(p.tw + PL::OFF_POST + 3 * kap);
        }
        if (lane == 0) {
            ring_bar_init(bars, 1);
            ring_bar_init(bars + 1, 1);
        }
        ring_fence_init();
    }
    // SUM: the running sums of this lane, slot j = task pp (bins kap, kap + 256, 512 - kap, 256 - kap), slot 8 = (128, 384)
    [[maybe_unused]] float4* const sacc = reinterpret_cast<float4*>(
        reinterpret_cast<unsigned char*>(sm4) + PP::smem_bytes(p.hop, ESZ, PP::NT)) + tid;
    [[maybe_unused]] unsigned tacc = 0;
#ifndef B2S_EMU
    if constexpr (SUM == 2) {
        __shared__ unsigned tmem_base_s;
        tacc = tm_alloc_cta<PP::TMEM_COLS>(&tmem_base_s, tid);
    } else
#endif
    {
        __syncthreads();
    }

    const int kout = p.kmax - p.kmin + 1;
    EpiOne<MODE> epi;
    epi.floor = p.db_floor;
    epi.kmin = p.kmin;
    epi.kmax = p.kmax;
    epi.db = p.out_mode;
    epi.band = 0.f;
    const int partner = (lane & 16) | ((16 - t) & 15);
    const bool is0 = (t == 0);

    // ---- work units: runs of consecutive frames of one signal, one run per warp at a time ----
    const bool dyn = p.work != nullptr;
    auto draw = [&]() -> long long {
        int b0 = 0;
        if (lane == 0) b0 = atomicAdd(p.work, 1);
        return (long long)__shfl_sync(0xffffffffu, b0, 0);
    };
    struct Unit {
        const Tin* x;        // first sample of the run
        float* out;          // row of the run's first frame (already offset by -kmin)
        int nf;              // frames in the run
        int nit;             // iterations: pairs of frames of the run (SUM: sweeps of the block)
        int blk, f0;         // SUM: sweep block, first frame of the pair
    };
    auto unit_of = [&](long long u) -> Unit {
        const long long b = u / p.units_per_signal;
        const int c = (int)(u - b * p.units_per_signal);
        if constexpr (SUM != 0) {
            // unit = (sweep block b, frame pair c), the pair index fastest: warps at work at the same time
            // are on neighbouring frames of the same sweeps (their overlapping samples meet in L2)
            const long long b_begin = b * p.acc_rows;
            Unit r;
            r.x = reinterpret_cast<const Tin*>(p.x) + b_begin * p.x_batch_stride + (p.frame0 + 2 * c) * (long long)p.hop;
            r.out = p.out + b_begin * p.out_batch_stride + (long long)(2 * c) * kout;
            r.nf = (2 * c + 1 < p.nframes) ? 2 : 1;
            r.nit = (int)((b_begin + p.acc_rows < p.acc_batch) ? p.acc_rows : p.acc_batch - b_begin);
            r.blk = (int)b;
            r.f0 = 2 * c;
            return r;
        }
        const int f_begin = c * p.chunk_frames;
        const int f_end = (f_begin + p.chunk_frames < p.nframes) ? f_begin + p.chunk_frames : p.nframes;
        Unit r;
        r.x = reinterpret_cast<const Tin*>(p.x) + b * p.x_batch_stride + (p.frame0 + f_begin) * (long long)p.hop;
        r.out = p.out + b * p.out_batch_stride +
                ((MODE == EPI_BAND) ? (long long)f_begin : (long long)f_begin * kout - p.kmin);
        r.nf = f_end - f_begin;
        r.nit = (r.nf + 1) >> 1;
        r.blk = 0;
        r.f0 = f_begin;
        return r;
    };
    // chunk j of a run of nf frames: j == 0 the first N + hop samples, then 2 hop per pair of frames,
    // cut at the run's last sample (span = (nf - 1) hop + N).  Returns the sample count (0: none).
    const int hop = p.hop;
    // (SUM: chunk j is the whole pair of sweep j of the block -- span samples at ring position 0)
    auto chunk_lo = [&](int j) -> int { return j == 0 ? 0 : N + hop + 2 * hop * (j - 1); };
    auto chunk_len = [&](int nf, int j) -> int {
        const int span = (nf - 1) * hop + N;
        if constexpr (SUM != 0) return span;
        const int lo = chunk_lo(j), hi = (j == 0) ? N + hop : lo + 2 * hop;
        const int e = hi < span ? hi : span;
        return e > lo ? e - lo : 0;
    };
    // lane 0: feed chunk j of the run into the ring (sample s of the run lives at s mod RS)
    unsigned issued = 0, waited = 0;             // chunks issued / waited for: barrier = count & 1, parity = (count >> 1) & 1
    auto issue = [&](const Unit& un, int j) {
        const int len = chunk_len(un.nf, j);
        if (len == 0) return;                    // (warp-uniform)
        if constexpr (SUM != 0) {
            if (lane == 0) {
                void* const bar = bars + (issued & 1u);
                ring_expect(bar, (unsigned)(len * ESZ));
                ring_copy(ring, un.x + (long long)j * p.x_batch_stride, (unsigned)(len * ESZ), bar);
            }
            ++issued;
            return;
        }
        if (lane == 0) {
            void* const bar = bars + (issued & 1u);
            const int lo = chunk_lo(j);
            const int pos = lo % RS;
            const int first = (pos + len <= RS) ? len : RS - pos;
            ring_expect(bar, (unsigned)(len * ESZ));
            ring_copy(ring + (size_t)pos * ESZ, un.x + lo, (unsigned)(first * ESZ), bar);
            if (first < len) ring_copy(ring, un.x + lo + first, (unsigned)((len - first) * ESZ), bar);
        }
        ++issued;
    };
    auto wait_chunk = [&](const Unit& un, int j) {
        if (chunk_len(un.nf, j) == 0) return;
#ifdef B2S_EMU
        __syncwarp();
#endif
        ring_wait(bars + (waited & 1u), (waited >> 1) & 1u);
        ++waited;
    };

    const long long ustride = (long long)gridDim.x * nwarps;
    long long u_cur = dyn ? draw() : (long long)blockIdx.x * nwarps + warp;
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
        const int nit = un.nit;
        int pos = 0;                                     // ring position (samples) of frame 2 it
        if constexpr (SUM != 0) {                        // the block sums start at zero; only this lane touches its slots
#ifndef B2S_EMU
            if constexpr (SUM == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) tm_st4(tacc + 4 * j, 0.f, 0.f, 0.f, 0.f);
                tm_st2(tacc + 32, 0.f, 0.f);
            } else
#endif
            {
#pragma unroll
                for (int j = 0; j < PP::ACC_SLOTS; ++j) sacc[j * PP::NT] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        for (int it = 0; it < nit; ++it) {
            const int fr = (SUM != 0) ? h : 2 * it + h;  // this half-warp's frame of the run
            epi.act = fr < un.nf;
            epi.row = un.out + (long long)fr * kout;
            if constexpr (SUM != 0) epi.row += (long long)it * p.out_batch_stride;
            wait_chunk(un, it);

            // ---- the frame's samples: slot i = x[4 (t + 16 i) .. + 3], 16 lanes read 256 contiguous bytes ----
            int w0 = pos + h * hop;                      // frame start in the ring (samples)
            if constexpr (SUM != 0) w0 = epi.act ? h * hop : 0;    // (no frame B: transform frame A twice)
            if (w0 >= RS) w0 -= RS;
            w0 = w0 / 4 + t;                             // in float4s of samples, this lane's first slot
            const int RW4 = RS / 4;
            auto slot_word = [&](int i) -> int {
                int w = w0 + 16 * i;
                return (w >= RW4) ? w - RW4 : w;
            };
            cpx2 v[16];
            if constexpr (sizeof(Tin) == 8) {
                // float64 samples: two trips over the staged samples instead of 128 live registers; the
                // pivot is subtracted in double and the difference rounded once
                double cd = 0.0;
                if (p.detrend) {
                    float s[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 q = RingLoad<double>::ld4(ring, slot_word(i));
                        s[i] = (q.x + q.y) + (q.z + q.w);
                        if ((i & 3) == 3) B2S_SCHED_FENCE();
                    }
#pragma unroll
                    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                        for (int i = 0; i < w; ++i) s[i] += s[i + w];
                    float c = s[0];
#pragma unroll
                    for (int o = 8; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    cd = (double)(c * (1.0f / (float)N));
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 q = RingLoad<double>::ld4_minus(ring, slot_word(i), cd);
                    v[i].re = cmk(q.x, q.y);
                    v[i].im = cmk(q.z, q.w);
                    if ((i & 3) == 3) B2S_SCHED_FENCE();
                }
            } else {
                float4 raw[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) raw[i] = RingLoad<Tin>::ld4(ring, slot_word(i));
                if (p.detrend) {
                    // coarse mean (pivot): fixed tree over the lane's 64 samples, xor-butterfly over the 16 lanes
                    float s[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) s[i] = (raw[i].x + raw[i].y) + (raw[i].z + raw[i].w);
#pragma unroll
                    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                        for (int i = 0; i < w; ++i) s[i] += s[i + w];
                    float c = s[0];
#pragma unroll
                    for (int o = 8; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    c *= 1.0f / (float)N;
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
            if (p.detrend) {
                // mean of the residual, removed inside the window multiply
                float2 sr[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) sr[i] = pk_add(v[i].re, v[i].im);
#pragma unroll
                for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                    for (int i = 0; i < w; ++i) sr[i] = pk_add(sr[i], sr[i + w]);
                float r = sr[0].x + sr[0].y;
#pragma unroll
                for (int o = 8; o >= 1; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
                const float nr = r * (-1.0f / (float)N);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 w = sm4[PP::OFF_WIN + i * 16 + t];
                    // (x' - r) w with ONE rounding: at this error level (rms 6e-6 of a bin 60 dB under the
                    // peak) a second rounding per sample is measurable (7.7e-6, worst bin 1.09e-4 vs 0.82e-4)
                    v[i].re = pk_fma(v[i].re, cmk(w.x, w.y), pk_muls(cmk(w.x, w.y), nr));
                    v[i].im = pk_fma(v[i].im, cmk(w.z, w.w), pk_muls(cmk(w.z, w.w), nr));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 w = sm4[PP::OFF_WIN + i * 16 + t];
                    v[i].re = pk_mul(v[i].re, cmk(w.x, w.y));
                    v[i].im = pk_mul(v[i].im, cmk(w.z, w.w));
                }
            }

            // Every lane has CONSUMED its samples (a barrier alone does not wait for shared-memory loads still in
            // flight): the ring positions before frame 2 (it + 1) are free -- feed the next pair's samples, or the
            // next run's first pair, into them.
            __syncwarp();
            if (it + 1 < nit) issue(un, it + 1);
            else if (have_next) issue(unn, 0);

            // ---- sub-transforms: radix-16, 16 x 16 transpose inside the half-warp, radix-16 ----
            c2radix16(v);
            if constexpr (HALF) {
                float2* const b2 = reinterpret_cast<float2*>(buf);
                float2 rr[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) b2[ROW * t + q] = v[perm16(q)].re;
                __syncwarp();
#pragma unroll
                for (int tt = 0; tt < 16; ++tt) rr[tt] = b2[ROW * tt + t];
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 16; ++q) b2[ROW * t + q] = v[perm16(q)].im;
                __syncwarp();
#pragma unroll
                for (int tt = 0; tt < 16; ++tt) v[tt] = cpx2{rr[tt], b2[ROW * tt + t]};
            } else {
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
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 w = sm4[PP::OFF_TW1 + j * 16 + t];
                if (j > 0) v[2 * j] = c2mul(v[2 * j], cmk(w.x, w.y));
                v[2 * j + 1] = c2mul(v[2 * j + 1], cmk(w.z, w.w));
            }
            c2radix16(v);
            __syncwarp();                        // every lane of the half-warp has consumed its reads

            // ---- fused final stage.  v[perm16(pp)] = F[kap] = (F_A, F_B)[kap], kap = t + 16 pp, with
            // F_A = X_0 + i X_2, F_B = X_1 + i X_3 (X_r the spectrum of x[4 m + r]).  Task pp < 8 takes the
            // mirror F' = F[256 - kap] from lane 16 - t (its slot 15 - pp; lane 0 pairs kap = 16 pp with its own
            // slot 16 - pp, and kap = 0 with itself) and produces bins kap, kap + 256, 256 - kap, 512 - kap.
            auto task = [&](cpx2 F, cpx2 G, float2 w1, float2 w2, float2 w3, int kap, bool dc, bool mid, int slot) {
                [[maybe_unused]] float4 a4;
#ifndef B2S_EMU
                if constexpr (SUM == 2) {
                    if (slot == 0) tm_st_wait();                       // the previous sweep's updates have landed
                    if (slot < 8) tm_ld4(tacc + 4 * slot, a4.x, a4.y, a4.z, a4.w);      // in flight during the butterfly below
                    else tm_ld2(tacc + 32, a4.x, a4.y);
                }
#endif
                const cpx2 U{pk_add(F.re, G.re), pk_sub(F.im, G.im)};       // (2 X_0, 2 X_1) = F + conj(F')
                const cpx2 V{pk_add(F.im, G.im), pk_sub(G.re, F.re)};       // (2 X_2, 2 X_3) = -i (F - conj(F'))
                // Y_r = W_1024^(r kap) 2 X_r
                const float y1r = fmaf(-U.im.y, w1.y, U.re.y * w1.x), y1i = fmaf(U.im.y, w1.x, U.re.y * w1.y);
                const float y2r = fmaf(-V.im.x, w2.y, V.re.x * w2.x), y2i = fmaf(V.im.x, w2.x, V.re.x * w2.y);
                const float y3r = fmaf(-V.im.y, w3.y, V.re.y * w3.x), y3i = fmaf(V.im.y, w3.x, V.re.y * w3.y);
                const float2 ar = cmk(U.re.x, y1r), ai = cmk(U.im.x, y1i);  // (Y_0, Y_1)
                const float2 br = cmk(y2r, y3r), bi = cmk(y2i, y3i);        // (Y_2, Y_3)
                const float2 sr = pk_add(ar, br), si = pk_add(ai, bi);      // (Y_0 + Y_2, Y_1 + Y_3)
                const float2 dr = pk_sub(ar, br), di = pk_sub(ai, bi);      // (Y_0 - Y_2, Y_1 - Y_3)
                // 2 X[kap] = S0 + S1, 2 conj X[512 - kap] = S0 - S1, 2 X[kap + 256] = D0 - i D1, 2 conj X[256 - kap] = D0 + i D1
                const float2 pr = cmk(sr.x + sr.y, sr.x - sr.y), pi = cmk(si.x + si.y, si.x - si.y);
                const float2 qr = cmk(dr.x + di.y, dr.x - di.y), qi = cmk(di.x - dr.y, di.x + dr.y);
                float2 ps = pk_fma(pr, pr, pk_mul(pi, pi));                 // bins kap, 512 - kap
                const float2 pd = pk_fma(qr, qr, pk_mul(qi, qi));           // bins kap + 256, 256 - kap
                if (dc) ps = pk_muls(ps, 0.5f);                             // DC / Nyquist carry scale, not 2 scale
                epi.put(kap, ps.x);
                epi.put(kap + M / 2, pd.x);
                if (!mid) {
                    epi.put(M - kap, ps.y);
                    if (!dc) epi.put(M / 2 - kap, pd.y);
                }
                if constexpr (SUM != 0) {
                    // (kap, kap + 256 | 512 - kap, 256 - kap); slot 8 (lane 0's kap = 128): only the first two
#ifndef B2S_EMU
                    if constexpr (SUM == 2) {
                        if (slot < 8) {
                            tm_ld_wait4(a4.x, a4.y, a4.z, a4.w);
                            const float2 a0 = pk_add(cmk(a4.x, a4.y), cmk(ps.x, pd.x)), a1 = pk_add(cmk(a4.z, a4.w), cmk(ps.y, pd.y));
                            tm_st4(tacc + 4 * slot, a0.x, a0.y, a1.x, a1.y);
                        } else {
                            tm_ld_wait2(a4.x, a4.y);
                            const float2 a0 = pk_add(cmk(a4.x, a4.y), cmk(ps.x, pd.x));
                            tm_st2(tacc + 32, a0.x, a0.y);
                        }
                    } else
#endif
                    {
                        a4 = sacc[slot * PP::NT];
                        const float2 a0 = pk_add(cmk(a4.x, a4.y), cmk(ps.x, pd.x)), a1 = pk_add(cmk(a4.z, a4.w), cmk(ps.y, pd.y));
                        sacc[slot * PP::NT] = make_float4(a0.x, a0.y, a1.x, a1.y);
                    }
                }
            };
            const float2* const w3tab = reinterpret_cast<const float2*>(sm4 + PP::OFF_W3);
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
                const cpx2 F = v[perm16(pp)];
                const cpx2 s15 = v[perm16(15 - pp)], s16 = v[perm16((16 - pp) & 15)];
                const float s0 = is0 ? s16.re.x : s15.re.x, s1 = is0 ? s16.re.y : s15.re.y;
                const float s2 = is0 ? s16.im.x : s15.im.x, s3 = is0 ? s16.im.y : s15.im.y;
                cpx2 G;
                G.re = cmk(__shfl_sync(0xffffffffu, s0, partner), __shfl_sync(0xffffffffu, s1, partner));
                G.im = cmk(__shfl_sync(0xffffffffu, s2, partner), __shfl_sync(0xffffffffu, s3, partner));
                const float4 w12 = sm4[PP::OFF_W12 + pp * 16 + t];
                const float2 w3 = w3tab[pp * 16 + t];
                task(F, G, cmk(w12.x, w12.y), cmk(w12.z, w12.w), w3, t + 16 * pp, pp == 0 && is0, false, pp);
            }
            if (is0 || SUM == 2) {               // kap = 128 is its own mirror: bins 128 and 384 (lane 0's; the
                                                 // tensor-memory accesses are warp-wide, so with SUM = 2 every lane goes
                                                 // through the motions and only lane 0's stores and sums are used)
                const cpx2 F = v[perm16(8)];
                const bool keep = epi.act;
                if constexpr (SUM == 2) epi.act = keep && is0;
                task(F, F, cmk(B2S_SQRT1_2, -B2S_SQRT1_2), cmk(0.f, -1.f), cmk(-B2S_SQRT1_2, -B2S_SQRT1_2), 128, false, true, 8);
                if constexpr (SUM == 2) epi.act = keep;
            }
            if constexpr (MODE == EPI_BAND) {
                float bs = epi.band;
                epi.band = 0.f;
#pragma unroll
                for (int o = 8; o >= 1; o >>= 1) bs += __shfl_xor_sync(0xffffffffu, bs, o);
                if (is0 && epi.act) un.out[fr] = bs;
            }
            if constexpr (SUM == 0) {
                pos += 2 * hop;
                while (pos >= RS) pos -= RS;
            }
        }
        if constexpr (SUM != 0) {
            // ---- the block's partial sums: p.acc[blk][frame][bin] ----
            float4 a[PP::ACC_SLOTS];
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
                for (int j = 0; j < PP::ACC_SLOTS; ++j) a[j] = sacc[j * PP::NT];
            }
            if (h < un.nf) {
                float* const sA = p.acc + ((long long)un.blk * p.nframes + un.f0 + h) * (M + 1);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int kap = t + 16 * j;
                    sA[kap] = a[j].x;
                    sA[kap + M / 2] = a[j].y;
                    sA[M - kap] = a[j].z;
                    if (kap != 0) sA[M / 2 - kap] = a[j].w;
                }
                if (is0) {
                    sA[128] = a[8].x;
                    sA[128 + M / 2] = a[8].y;
                }
            }
        }
        u_cur = u_next;
        un = unn;
        if (have_next) u_next = dyn ? draw() : u_next + ustride;
    }
#ifndef B2S_EMU
    if constexpr (SUM == 2) tm_free_cta<PP::TMEM_COLS>(tacc, tid);
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

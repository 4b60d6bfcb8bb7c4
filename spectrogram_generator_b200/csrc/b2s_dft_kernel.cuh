// Direct-DFT STFT -> PSD kernel for every nperseg the radix-16 kernels do not take
// (non-powers of two, and lengths below 32).  The GUI's nperseg spin box accepts any
// integer in 32..8192 (GUI.py:87-89) and SciPy clamps nperseg to len(x)
// (_spectral_py.py:2443-2447), so a drop-in has to produce these sizes too -- on the
// GPU, not through a CPU fallback.
//
// One CTA per frame (grid-stride over batch x frames).  The frame is detrended (two
// fp32 passes, as in the FFT kernels), windowed and parked in shared memory next to
// the table W_N^j; thread t then evaluates bins k = t, t+NT, ... with
//     X[k] = sum_n y[n] W_N^(n k mod N),   n k mod N kept incrementally,
// accumulating blocks of 8 terms in fp32 and the block sums in fp64, which keeps the
// rounding error at the level of the FFT kernels (the sqrt(N) growth of a plain fp32
// running sum would not meet the 1e-4 bar).  O(N^2/2) per frame: microseconds for the
// sizes the GUI can ask for, and never the fast path.
#pragma once

#include "b2s_kernels.cuh"

namespace b2s {

struct DftParams {
    const void* x;
    long long x_batch_stride;
    long long frame0;
    long long out_batch_stride;
    long long total_frames;      // batch * nframes
    const float* window;         // [nperseg]
    const float2* tw;            // [nperseg] W_N^j
    float* out;
    int nframes;
    int nperseg;
    int hop;
    int detrend;
    int out_mode;
    int kmin, kmax;
    int band;                    // 1: write the per-frame sum of bins kmin..kmax to out[b][f]
    float scale;
    float db_floor;
};

constexpr int kDftThreads = 256;

B2S_DEVICE float dft_block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int tid = (int)threadIdx.x;
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kDftThreads / 32; ++w) t += red[w];
    __syncthreads();
    return t;
}

template <typename Tin>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(kDftThreads, 2) dft_psd_kernel(const DftParams p) {
    B2S_DYN_SMEM_F2(sm);
    const int N = p.nperseg;
    const int K = N / 2 + 1;
    const int nyq = (N % 2 == 0) ? N / 2 : -1;
    const int tid = (int)threadIdx.x;
    float2* const W = sm;                                          // [N]
    float* const y = reinterpret_cast<float*>(sm + N);             // [N] (+ pad to 4)
    float* const red = y + ((N + 3) & ~3);                         // [8]
    for (int i = tid; i < N; i += kDftThreads) W[i] = __ldg(p.tw + i);
    const int kout = p.kmax - p.kmin + 1;
    const float inv_n = 1.0f / (float)N;

    for (long long fi = blockIdx.x; fi < p.total_frames; fi += gridDim.x) {
        const long long b = fi / p.nframes;
        const int f = (int)(fi - b * p.nframes);
        const Tin* const xf = reinterpret_cast<const Tin*>(p.x) + b * p.x_batch_stride + (p.frame0 + f) * (long long)p.hop;
        float* const row = p.out + b * p.out_batch_stride + (long long)f * kout - p.kmin;
        __syncthreads();                                           // previous frame's readers are done with y
        // ---- gather + detrend + window -> y ----
        float part = 0.f;
        for (int i = tid; i < N; i += kDftThreads) {
            const float v = Loader<Tin>::ld1(xf + i);
            y[i] = v;
            part += v;
        }
        if (p.detrend) {
            const float m1 = dft_block_sum(part, red) * inv_n;
            float part2 = 0.f;
            for (int i = tid; i < N; i += kDftThreads) {
                const float v = y[i] - m1;
                y[i] = v;
                part2 += v;
            }
            const float nr = -dft_block_sum(part2, red) * inv_n;
            for (int i = tid; i < N; i += kDftThreads) {
                const float w = __ldg(p.window + i);
                y[i] = fmaf(y[i], w, nr * w);
            }
        } else {
            for (int i = tid; i < N; i += kDftThreads) y[i] *= __ldg(p.window + i);
        }
        __syncthreads();
        // ---- bins ----
        float band = 0.f;
        for (int k = tid; k < K; k += kDftThreads) {
            double re = 0.0, im = 0.0;
            int idx = 0;
            int n = 0;
            for (; n + 8 <= N; n += 8) {
                float br = 0.f, bi = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float2 w = W[idx];
                    const float v = y[n + q];
                    br = fmaf(v, w.x, br);
                    bi = fmaf(v, w.y, bi);
                    idx += k;
                    if (idx >= N) idx -= N;
                }
                re += (double)br;
                im += (double)bi;
            }
            float br = 0.f, bi = 0.f;
            for (; n < N; ++n) {
                const float2 w = W[idx];
                const float v = y[n];
                br = fmaf(v, w.x, br);
                bi = fmaf(v, w.y, bi);
                idx += k;
                if (idx >= N) idx -= N;
            }
            re += (double)br;
            im += (double)bi;
            const float xr = (float)re, xi = (float)im;
            float pw = fmaf(xr, xr, xi * xi) * ((k == 0 || k == nyq) ? p.scale : 2.0f * p.scale);
            if (p.band) {
                if (k >= p.kmin && k <= p.kmax) band += pw;
            } else {
                if (p.out_mode) pw = 10.0f * log10f(fmaxf(pw, p.db_floor));
                if (k >= p.kmin && k <= p.kmax) row[k] = pw;
            }
        }
        if (p.band) {
            const float bs = dft_block_sum(band, red);
            if (tid == 0) p.out[b * p.out_batch_stride + f] = bs;
        }
    }
}

inline size_t dft_smem_bytes(int nperseg) {
    return (size_t)nperseg * sizeof(float2) + (size_t)((nperseg + 3) & ~3) * sizeof(float) + 8 * sizeof(float);
}

}  // namespace b2s

// Mixed-radix STFT -> PSD kernel for the nperseg that are not powers of two but factor into small
// primes (2, 3, 5, 7, 11, 13): the GUI's spin box steps by 32 but accepts any integer 32..8192
// (GUI.py:87-89), and typed values such as 1000, 2000, 4800 or 8000 are 5-smooth.  Round 1 ran all of
// them on the O(N^2) direct-DFT kernel (b2s_dft_kernel.cuh), which stays the fallback for lengths with
// a prime factor above 13 (and for nperseg < 32).
//
// One CTA per frame (grid-stride over batch x frames).  The frame is detrended (two fp32 passes, as in
// every other kernel), windowed and parked in shared memory.  Even N: the real-FFT trick -- the N real
// samples are Mc = N/2 complex points z[n] = x[2n] + i x[2n+1]; odd N: Mc = N points with zero imaginary
// parts.  A Stockham autosort FFT of Mc points follows, one pass per radix (4 and 2 hard-coded, the odd
// primes as table-driven r-point DFTs), ping-ponging between two shared-memory buffers, twiddles
// W_Mc^j from a table built in double; even N then takes the split X[k] = E[k] + W_N^k O[k].  The PSD
// epilogue is the direct-DFT kernel's (scale, x2 on interior bins, optional dB / crop / band sum).
// O(N log N) per frame instead of O(N^2).
#pragma once

#include "b2s_dft_kernel.cuh"

namespace b2s {

constexpr int kMixedMaxPasses = 16;
constexpr int kMixedThreads = kDftThreads;

struct MixedParams {
    DftParams d;                 // d.tw: [Mc] W_Mc^j, then (even N) [Mc + 1] W_N^k
    int mc;                      // complex points of the transform
    int npass;
    int radix[kMixedMaxPasses];
};

// radices of an Mc-point transform, largest first where it matters little: 4s, a 2, then the odd primes.
// Returns the number of passes, 0 if Mc has a prime factor above 13 (or needs too many passes).
inline int mixed_radices(int mc, int* radix) {
    int n = 0;
    int m = mc;
    if (m < 1) return 0;
    while (m % 4 == 0 && n < kMixedMaxPasses) { radix[n++] = 4; m /= 4; }
    while (m % 2 == 0 && n < kMixedMaxPasses) { radix[n++] = 2; m /= 2; }
    const int primes[5] = {3, 5, 7, 11, 13};
    for (int p : primes)
        while (m % p == 0 && n < kMixedMaxPasses) { radix[n++] = p; m /= p; }
    return (m == 1 && n > 0) ? n : 0;
}

// true if nperseg runs on this kernel: not a power of two (those have the radix-16 kernels), smooth, and at
// least 256 -- below that a 256-thread CTA per frame is mostly idle and the direct-DFT kernel is as fast or
// faster (measured on B200, profiles/r2_nonpow2.md: nperseg 96 3.5 vs 5.2 ms, 352 5.7 vs 2.3 ms)
inline bool mixed_supported(int nperseg) {
    if (nperseg < 256 || nperseg > 16384) return false;
    if ((nperseg & (nperseg - 1)) == 0) return false;
    int r[kMixedMaxPasses];
    return mixed_radices((nperseg % 2 == 0) ? nperseg / 2 : nperseg, r) > 0;
}

inline size_t mixed_smem_bytes(int nperseg) {
    const int mc = (nperseg % 2 == 0) ? nperseg / 2 : nperseg;
    return (size_t)2 * (mc + 1) * sizeof(float2) + 8 * sizeof(float);
}

// one Stockham pass of radix R over `mc` points: sub-transforms of length ns -> ns * R
template <int R>
B2S_DEVICE void mixed_pass(const float2* __restrict__ in, float2* __restrict__ out, int mc, int ns, const float2* __restrict__ W) {
    const int T = mc / R;
    const int tstep = mc / (ns * R);                     // W_(ns R)^1 = W_mc^tstep
    const int rstep = mc / R;                            // W_R^1 = W_mc^rstep
    for (int j = (int)threadIdx.x; j < T; j += kMixedThreads) {
        const int k = j % ns;
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            v[q] = in[j + q * T];
            if (q > 0 && k > 0) v[q] = cmul(v[q], __ldg(W + (int)(((long long)q * k * tstep) % mc)));
        }
        float2 o[R];
        if constexpr (R == 2) {
            o[0] = cadd(v[0], v[1]);
            o[1] = csub(v[0], v[1]);
        } else if constexpr (R == 4) {
            radix4(v[0], v[1], v[2], v[3]);
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = v[q];
        } else {
            // table-driven R-point DFT: o[a] = sum_b v[b] W_R^(a b)
#pragma unroll
            for (int a = 0; a < R; ++a) {
                float2 acc = v[0];
#pragma unroll
                for (int b = 1; b < R; ++b) {
                    const float2 w = __ldg(W + ((a * b) % R) * rstep);
                    acc.x = fmaf(v[b].x, w.x, fmaf(-v[b].y, w.y, acc.x));
                    acc.y = fmaf(v[b].x, w.y, fmaf(v[b].y, w.x, acc.y));
                }
                o[a] = acc;
            }
        }
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; ++q) out[j0 + q * ns] = o[q];
    }
}

template <typename Tin>
B2S_GLOBAL void B2S_LAUNCH_BOUNDS(kMixedThreads, 2) mixed_psd_kernel(const MixedParams mp) {
    const DftParams& p = mp.d;
    B2S_DYN_SMEM_F2(sm);
    const int N = p.nperseg;
    const int mc = mp.mc;
    const bool even = (N % 2) == 0;
    const int K = N / 2 + 1;
    const int nyq = even ? N / 2 : -1;
    const int tid = (int)threadIdx.x;
    float2* bufA = sm;                                   // [mc + 1]
    float2* bufB = sm + (mc + 1);                        // [mc + 1]
    float* const red = reinterpret_cast<float*>(sm + 2 * (mc + 1));
    const float2* const W = p.tw;                        // [mc] W_mc^j
    const float2* const WN = p.tw + mc;                  // [mc + 1] W_N^k (even N)
    const int kout = p.kmax - p.kmin + 1;
    const float inv_n = 1.0f / (float)N;

    for (long long fi = blockIdx.x; fi < p.total_frames; fi += gridDim.x) {
        const long long b = fi / p.nframes;
        const int f = (int)(fi - b * p.nframes);
        const Tin* const xf = reinterpret_cast<const Tin*>(p.x) + b * p.x_batch_stride + (p.frame0 + f) * (long long)p.hop;
        float* const row = p.out + b * p.out_batch_stride + (long long)f * kout - p.kmin;
        __syncthreads();                                 // the previous frame's readers are done with the buffers
        // ---- gather + detrend + window: y[i] is sample i; even N packs them as z[n] = (y[2n], y[2n+1]) ----
        float* const y = reinterpret_cast<float*>(bufA);
        auto put_y = [&](int i, float v) {
            if (even) y[i] = v;
            else bufA[i] = cmk(v, 0.f);
        };
        auto get_y = [&](int i) -> float { return even ? y[i] : bufA[i].x; };
        float part = 0.f;
        for (int i = tid; i < N; i += kMixedThreads) {
            const float v = Loader<Tin>::ld1(xf + i);
            put_y(i, v);
            part += v;
        }
        if (p.detrend) {
            const float m1 = dft_block_sum(part, red) * inv_n;
            float part2 = 0.f;
            for (int i = tid; i < N; i += kMixedThreads) {
                const float v = get_y(i) - m1;
                put_y(i, v);
                part2 += v;
            }
            const float nr = -dft_block_sum(part2, red) * inv_n;
            for (int i = tid; i < N; i += kMixedThreads) {
                const float w = __ldg(p.window + i);
                put_y(i, fmaf(get_y(i), w, nr * w));
            }
        } else {
            for (int i = tid; i < N; i += kMixedThreads) put_y(i, get_y(i) * __ldg(p.window + i));
        }
        __syncthreads();
        // ---- Stockham passes ----
        float2* in = bufA;
        float2* out = bufB;
        int ns = 1;
        for (int ps = 0; ps < mp.npass; ++ps) {
            const int r = mp.radix[ps];
            switch (r) {
                case 2: mixed_pass<2>(in, out, mc, ns, W); break;
                case 3: mixed_pass<3>(in, out, mc, ns, W); break;
                case 4: mixed_pass<4>(in, out, mc, ns, W); break;
                case 5: mixed_pass<5>(in, out, mc, ns, W); break;
                case 7: mixed_pass<7>(in, out, mc, ns, W); break;
                case 11: mixed_pass<11>(in, out, mc, ns, W); break;
                default: mixed_pass<13>(in, out, mc, ns, W); break;
            }
            ns *= r;
            __syncthreads();
            float2* const tmp = in;
            in = out;
            out = tmp;
        }
        const float2* const Z = in;                      // natural order
        // ---- bins ----
        float band = 0.f;
        for (int k = tid; k < K; k += kMixedThreads) {
            float xr, xi;
            if (even) {
                const float2 zk = Z[(k == mc) ? 0 : k], zm = Z[(k == 0 || k == mc) ? 0 : mc - k];
                const float2 w = __ldg(WN + k);
                const float er = zk.x + zm.x, ei = zk.y - zm.y;         // 2E = zk + conj(zm)
                const float orr = zk.y + zm.y, oi = zm.x - zk.x;        // 2O = -i (zk - conj(zm))
                const float tr = fmaf(-oi, w.y, orr * w.x), ti = fmaf(oi, w.x, orr * w.y);
                xr = 0.5f * (er + tr);
                xi = 0.5f * (ei + ti);
            } else {
                xr = Z[k].x;
                xi = Z[k].y;
            }
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

}  // namespace b2s

// Register-level FFT butterflies and complex helpers (fp32).
//
// Everything here is pure register arithmetic, callable from device code and
// (under B2S_EMU) from the CPU SIMT emulator used by tests/emu.  Forward
// transform convention: W_n = exp(-2*pi*i/n), as in the r2c transform SciPy
// reaches from _spectral_py.py:2395 (sp_fft.rfft).
#pragma once

#ifdef B2S_EMU
#include "emu_cuda.hpp"
#else
#include <cuda_runtime.h>
#endif

#ifdef B2S_EMU
#define B2S_HD inline
#define B2S_DEVICE inline
#define B2S_GLOBAL inline
#define B2S_LAUNCH_BOUNDS(a, b)
#define B2S_DYN_SMEM(name) unsigned char* const name = emu::dyn_smem()
#define B2S_DYN_SMEM_F2(name) float2* const name = reinterpret_cast<float2*>(emu::dyn_smem())
#define B2S_DYN_SMEM_F4(name) float4* const name = reinterpret_cast<float4*>(emu::dyn_smem())
#else
#define B2S_HD __host__ __device__ __forceinline__
#define B2S_DEVICE __device__ __forceinline__
#define B2S_GLOBAL __global__
#define B2S_LAUNCH_BOUNDS(a, b) __launch_bounds__(a, b)
#define B2S_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define B2S_DYN_SMEM_F2(name) extern __shared__ __align__(16) float2 name[]
#define B2S_DYN_SMEM_F4(name) extern __shared__ __align__(16) float4 name[]
// named barrier over `n` threads (n a multiple of 32), id 1..15
__device__ __forceinline__ void b2s_bar_sync(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
#endif

namespace b2s {

B2S_HD float2 cmk(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
// Complex add / subtract.  On sm_100a these are single packed FADD2 instructions (two fp32
// adds per issue slot, IEEE round-to-nearest like the scalar form -- results are identical).
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
B2S_HD float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
B2S_HD float2 csub(float2 a, float2 b) { return __fadd2_rn(a, cmk(-b.x, -b.y)); }
// a + s * b with s = (sx, sy) a compile-time sign pattern
B2S_HD float2 cfma_sign(float2 b, float sx, float sy, float2 a) { return __ffma2_rn(b, cmk(sx, sy), a); }
#else
B2S_HD float2 cadd(float2 a, float2 b) { return cmk(a.x + b.x, a.y + b.y); }
B2S_HD float2 csub(float2 a, float2 b) { return cmk(a.x - b.x, a.y - b.y); }
B2S_HD float2 cfma_sign(float2 b, float sx, float sy, float2 a) { return cmk(fmaf(b.x, sx, a.x), fmaf(b.y, sy, a.y)); }
#endif
// (a.x + i a.y) * (w.x + i w.y): 2 mul + 2 fma
B2S_HD float2 cmul(float2 a, float2 w) {
    return cmk(fmaf(-a.y, w.y, a.x * w.x), fmaf(a.y, w.x, a.x * w.y));
}
B2S_HD float2 cmul_mi(float2 a) { return cmk(a.y, -a.x); }   // * (-i)

// 4-point DFT in place, natural order out.
B2S_HD void radix4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
    const float2 t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cmk(t1.x + t3.y, t1.y - t3.x);
    a3 = cmk(t1.x - t3.y, t1.y + t3.x);
}

B2S_HD void radix2(float2& a0, float2& a1) {
    const float2 t = a0;
    a0 = cadd(t, a1);
    a1 = csub(t, a1);
}

#define B2S_SQRT1_2 0.70710678118654752440f
#define B2S_COS_PI_8 0.92387953251128675613f
#define B2S_SIN_PI_8 0.38268343236508977173f

// * W8^1 = sqrt(1/2) (1 - i)
B2S_HD float2 cmul_w8_1(float2 a) { return cmk(B2S_SQRT1_2 * (a.x + a.y), B2S_SQRT1_2 * (a.y - a.x)); }
// * W8^3 = sqrt(1/2) (-1 - i)
B2S_HD float2 cmul_w8_3(float2 a) { return cmk(B2S_SQRT1_2 * (a.y - a.x), -B2S_SQRT1_2 * (a.x + a.y)); }

// 16-point DFT in place.  Output X[k] is left in v[perm16(k)].
B2S_HD constexpr int perm16(int k) { return 4 * (k & 3) + (k >> 2); }

B2S_HD void radix16(float2 (&v)[16]) {
    // columns: n = c + 4m  ->  u_c[q] left in v[c + 4q]
#pragma unroll
    for (int c = 0; c < 4; ++c) radix4(v[c], v[c + 4], v[c + 8], v[c + 12]);
    // twiddle u_c[q] *= W16^(c q)
    const float2 W1 = cmk(B2S_COS_PI_8, -B2S_SIN_PI_8);
    const float2 W3 = cmk(B2S_SIN_PI_8, -B2S_COS_PI_8);
    const float2 W9 = cmk(-B2S_COS_PI_8, B2S_SIN_PI_8);
    v[5] = cmul(v[5], W1);   v[9] = cmul_w8_1(v[9]);   v[13] = cmul(v[13], W3);
    v[6] = cmul_w8_1(v[6]);  v[10] = cmul_mi(v[10]);   v[14] = cmul_w8_3(v[14]);
    v[7] = cmul(v[7], W3);   v[11] = cmul_w8_3(v[11]); v[15] = cmul(v[15], W9);
    // rows: over c for each q; X[q + 4p] left in v[4q + p]
#pragma unroll
    for (int q = 0; q < 4; ++q) radix4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// Small natural-order DFTs used by the fused final stage (g = 1, 2, 4, 8).
template <int R> struct SmallFft;
template <> struct SmallFft<1> { B2S_HD static void run(float2 (&)[1]) {} };
template <> struct SmallFft<2> { B2S_HD static void run(float2 (&v)[2]) { radix2(v[0], v[1]); } };
template <> struct SmallFft<4> {
    B2S_HD static void run(float2 (&v)[4]) { radix4(v[0], v[1], v[2], v[3]); }
};
template <> struct SmallFft<8> {
    B2S_HD static void run(float2 (&v)[8]) {
        // even / odd 4-point DFTs, then a radix-2 combine with W8^q
        radix4(v[0], v[2], v[4], v[6]);   // E[q] in v[2q]
        radix4(v[1], v[3], v[5], v[7]);   // O[q] in v[2q+1]
        const float2 o0 = v[1], o1 = cmul_w8_1(v[3]), o2 = cmul_mi(v[5]), o3 = cmul_w8_3(v[7]);
        const float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
        v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
        v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
        v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
        v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
    }
};

}  // namespace b2s

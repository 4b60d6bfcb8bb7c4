// Minimal CPU SIMT emulator -- TEST TOOLING ONLY.
//
// Lets tests compile the *same* kernel source (csrc/b2s_kernels.cuh) with g++
// and run one CTA at a time with one OS thread per CUDA thread, so the index
// math, shuffles and barrier discipline can be checked against the oracle
// in the no-GPU build container.  It is never linked into the product library.
#pragma once

#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct uint3_ { unsigned x = 0, y = 0, z = 0; };
struct dim3_ { unsigned x = 1, y = 1, z = 1; };

inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }

extern thread_local uint3_ threadIdx;
extern thread_local uint3_ blockIdx;
extern dim3_ blockDim;
extern dim3_ gridDim;

namespace emu {

struct Cta {
    std::vector<unsigned char> smem;
    std::mutex mu;
    std::map<uint64_t, std::unique_ptr<std::barrier<>>> bars;   // key -> barrier
    std::vector<std::array<unsigned char, 16>> slots;          // shuffle slots per thread
    std::barrier<>& get(uint64_t key, int count) {
        std::lock_guard<std::mutex> g(mu);
        auto it = bars.find(key);
        if (it == bars.end()) it = bars.emplace(key, std::make_unique<std::barrier<>>(count)).first;
        return *it->second;
    }
};
extern Cta* g_cta;

inline unsigned char* dyn_smem() { return g_cta->smem.data(); }

inline void warp_sync(unsigned mask) {
    const unsigned warp = threadIdx.x >> 5;
    g_cta->get((uint64_t(1) << 62) | (uint64_t(warp) << 32) | mask, __builtin_popcount(mask)).arrive_and_wait();
}

template <typename T>
inline T shfl_idx(unsigned mask, T v, int src_lane) {
    static_assert(sizeof(T) <= 16, "shuffle payload");
    const unsigned base = threadIdx.x & ~31u;
    std::memcpy(g_cta->slots[threadIdx.x].data(), &v, sizeof(T));
    warp_sync(mask);
    T r;
    std::memcpy(&r, g_cta->slots[base + (unsigned)(src_lane & 31)].data(), sizeof(T));
    warp_sync(mask);
    return r;
}

// launch `kernel(args...)` over grid x block, one CTA at a time
template <typename F>
void launch(unsigned grid, unsigned block, size_t smem_bytes, F&& body) {
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b) {
        Cta cta;
        cta.smem.assign(smem_bytes + 16, 0);
        cta.slots.resize(block);
        g_cta = &cta;
        std::vector<std::thread> th;
        th.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([&, t, b] {
                threadIdx.x = t;
                blockIdx.x = b;
                body();
            });
        for (auto& x : th) x.join();
        g_cta = nullptr;
    }
}

}  // namespace emu

inline void __syncthreads() { emu::g_cta->get(uint64_t(1) << 61, (int)blockDim.x).arrive_and_wait(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_sync(mask); }
inline void b2s_bar_sync(int id, int n) { emu::g_cta->get((uint64_t(1) << 60) | (uint64_t)id, n).arrive_and_wait(); }

template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
    return emu::shfl_idx(mask, v, (int)((threadIdx.x & 31u) ^ (unsigned)lane_mask));
}
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src_lane) { return emu::shfl_idx(mask, v, src_lane); }

inline int __any_sync(unsigned mask, int pred) {
    const unsigned base = threadIdx.x & ~31u;
    const int mine = pred ? 1 : 0;
    std::memcpy(emu::g_cta->slots[threadIdx.x].data(), &mine, sizeof(int));
    emu::warp_sync(mask);
    int any = 0;
    for (unsigned l = 0; l < 32; ++l) {
        if (!((mask >> l) & 1u)) continue;
        int v;
        std::memcpy(&v, emu::g_cta->slots[base + l].data(), sizeof(int));
        any |= v;
    }
    emu::warp_sync(mask);
    return any;
}

template <typename T>
inline T __ldg(const T* p) { return *p; }

inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline int atomicAdd(int* a, int v) {
    std::lock_guard<std::mutex> g(emu::g_cta->mu);
    const int old = *a;
    *a = old + v;
    return old;
}
inline unsigned atomicMax(unsigned* a, unsigned v) {
    std::lock_guard<std::mutex> g(emu::g_cta->mu);
    const unsigned old = *a;
    if (v > old) *a = v;
    return old;
}
inline unsigned atomicMin(unsigned* a, unsigned v) {
    std::lock_guard<std::mutex> g(emu::g_cta->mu);
    const unsigned old = *a;
    if (v < old) *a = v;
    return old;
}

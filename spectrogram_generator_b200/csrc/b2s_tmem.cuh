// Tensor memory (tcgen05.alloc / ld / st) used as per-thread scratch for running sums: shared by the
// sum-fused kernels (b2s_duo_sum_kernel.cuh, the SUM mode of b2s_pair_kernel.cuh).
#pragma once

namespace b2s {

#ifndef B2S_EMU
// Tensor memory as per-thread scratch (no tensor-core math involved): with the 32x32b shape a
// thread reads / writes N consecutive 32-bit columns of its own TMEM lane (warp w of the CTA owns
// lanes 32 (w % 4) ... + 31).  The accesses go over the tensor-memory datapath, not the L1 /
// shared-memory data pipe the rest of the kernel keeps busy.
__device__ __forceinline__ void tm_ld4(unsigned addr, float& a, float& b, float& c, float& d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr));
}
__device__ __forceinline__ void tm_ld2(unsigned addr, float& a, float& b) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(addr));
}
__device__ __forceinline__ void tm_ld_wait4(float& a, float& b, float& c, float& d) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(a), "+f"(b), "+f"(c), "+f"(d) : : "memory");
}
__device__ __forceinline__ void tm_ld_wait2(float& a, float& b) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(a), "+f"(b) : : "memory");
}
__device__ __forceinline__ void tm_st4(unsigned addr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 : : "r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tm_st2(unsigned addr, float a, float b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" : : "r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void tm_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" : : : "memory"); }
#endif

#ifndef B2S_EMU
// one warp of the CTA allocates NCOLS columns (a power of two >= 32); every thread then gets the address of
// its own lane (warp w of the CTA owns lanes 32 (w % 4) ... + 31).  Ends with a CTA-wide barrier.
template <int NCOLS>
__device__ __forceinline__ unsigned tm_alloc_cta(unsigned* base_slot_smem, int tid) {
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     : : "r"((unsigned)__cvta_generic_to_shared(base_slot_smem)), "n"(NCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" : : : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" : : : "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" : : : "memory");
    return *base_slot_smem + ((unsigned)((tid >> 5) & 3) << 21);        // lane field: 32 (warp % 4) << 16
}
// CTA-wide: every thread has finished with its columns; one warp frees them
template <int NCOLS>
__device__ __forceinline__ void tm_free_cta(unsigned taddr, int tid) {
    asm volatile("tcgen05.fence::before_thread_sync;" : : : "memory");
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" : : "r"(taddr & 0xffffu), "n"(NCOLS) : "memory");
}
#endif

}  // namespace b2s

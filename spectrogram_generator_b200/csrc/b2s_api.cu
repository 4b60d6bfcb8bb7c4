// C ABI of libb200stft.so (include/b2s.h): argument checking, twiddle cache,
// launch configuration.  No torch types, no allocations other than the
// per-(device, nperseg) twiddle tables.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b2s.h"
#include "b2s_aux_kernels.cuh"
#include "b2s_duo_sum_kernel.cuh"
#include "b2s_launcher.hpp"

namespace {

thread_local std::string g_err;
thread_local std::string g_last_kernel = "none";

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail(B2S_ERR_CUDA, std::string("b2s: CUDA error in ") + what + ": " + cudaGetErrorString(e));
}

struct DeviceInfo {
    int sm_count = 0;
    int smem_optin = 0;
};

std::mutex g_mu;
std::atomic<int> g_reserved_sms{0};
std::map<int, DeviceInfo> g_dev;
std::map<std::pair<int, int>, float2*> g_tw;     // (device, +-nperseg) -> tables (negative: direct-DFT table; + 2^20: mixed-radix table)

int device_info(DeviceInfo& out, int& dev) {
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_dev.find(dev);
    if (it == g_dev.end()) {
        DeviceInfo di;
        e = cudaDeviceGetAttribute(&di.sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
        e = cudaDeviceGetAttribute(&di.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
        it = g_dev.emplace(dev, di).first;
    }
    out = it->second;
    return B2S_OK;
}

// Twiddle table for nperseg on the current device; built on the host in double,
// uploaded once (synchronously, first use only) and cached for the process.
int twiddles(int dev, int nperseg, bool dft, const float2** out, bool mixed = false) {
    std::lock_guard<std::mutex> g(g_mu);
    auto key = std::make_pair(dev, mixed ? nperseg + (1 << 20) : (dft ? -nperseg : nperseg));
    auto it = g_tw.find(key);
    if (it == g_tw.end()) {
        std::vector<float> host;
        if (mixed) b2s::make_mixed_table(nperseg, host);
        else if (dft) b2s::make_dft_table(nperseg, host);
        else b2s::make_tables(nperseg, host);
        float2* d = nullptr;
        cudaError_t e = cudaMalloc(&d, host.size() * sizeof(float));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(twiddles)");
        e = cudaMemcpy(d, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(d);
            return cuda_fail(e, "cudaMemcpy(twiddles)");
        }
        it = g_tw.emplace(key, d).first;
    }
    *out = it->second;
    return B2S_OK;
}

// error flag of the peer all-reduce (set by the kernel when a peer misses the time-out), one per device
std::map<int, int*> g_peer_err;     // guarded by g_mu
int peer_err_flag(int dev, int** out) {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_peer_err.find(dev);
    if (it == g_peer_err.end()) {
        int* d = nullptr;
        cudaError_t e = cudaMalloc(&d, 32 * sizeof(int));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(peer error flag)");
        e = cudaMemset(d, 0, 32 * sizeof(int));
        if (e != cudaSuccess) {
            cudaFree(d);
            return cuda_fail(e, "cudaMemset(peer error flag)");
        }
        it = g_peer_err.emplace(dev, d).first;
    }
    *out = it->second;
    return B2S_OK;
}

struct KernelState {
    int occ = 0;
    int dev = -1;
};
std::map<const void*, KernelState> g_kern;     // guarded by g_mu

// Work counters of the dynamically scheduled kernels: one (next unit, CTAs done) pair per
// (device, stream), zero when idle -- the last CTA of a launch re-arms its pair, and launches of one
// stream do not overlap, so a pair is never shared by two launches in flight.  A launch that is
// being captured into a CUDA graph takes the static schedule instead (a replayed graph may run
// beside eager launches of the stream it was captured on).
std::map<std::pair<int, cudaStream_t>, int*> g_work;     // guarded by g_mu

int work_counters(int dev, cudaStream_t stream, int** out) {
    std::lock_guard<std::mutex> g(g_mu);
    auto key = std::make_pair(dev, stream);
    auto it = g_work.find(key);
    if (it == g_work.end()) {
        int* d = nullptr;
        cudaError_t e = cudaMalloc(&d, 32 * sizeof(int));       // one 128-byte line per stream
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(work counters)");
        e = cudaMemsetAsync(d, 0, 32 * sizeof(int), stream);    // ordered before the first launch that uses it
        if (e != cudaSuccess) {
            cudaFree(d);
            return cuda_fail(e, "cudaMemsetAsync(work counters)");
        }
        it = g_work.emplace(key, d).first;
    }
    *out = it->second;
    return B2S_OK;
}

bool stream_capturing(cudaStream_t stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st != cudaStreamCaptureStatusNone;
}

// diagnostic switches, read once per process
struct EnvSwitches {
    bool allow_duo = true, duo1024 = true, allow_duo4 = true, allow_big = true, dynamic_units = true;
    bool allow_pair = true, allow_pairq = false, fused_sum = true, sum_acc_smem = false, allow_mixed = true;
    int pair_units = 0;          // B2S_PAIR_UNITS: work units per resident warp of the pair kernel (0: default)
    int pair_nt = 0;             // B2S_PAIR_NT: threads per CTA of the pair kernel (0: default)
    int sum_blocks = 0;          // B2S_SUM_BLOCKS: sweep blocks of the sum-fused kernels (0: plan_stft_sum's choice)
    bool sum_dynamic = false;    // B2S_SUM_DYNAMIC: the sum-fused kernels draw their units from an atomic counter
    int peer_timeout_ms = 0;     // B2S_PEER_TIMEOUT_MS: how long the peer all-reduce waits for a late rank (0: 120 s)
    EnvSwitches() {
        auto on = [](const char* name) { const char* v = getenv(name); return v && atoi(v) != 0; };
        auto off0 = [](const char* name, bool dflt) { const char* v = getenv(name); return v ? atoi(v) != 0 : dflt; };
        allow_duo = !on("B2S_NO_DUO");
        duo1024 = off0("B2S_DUO1024", true);
        allow_duo4 = !on("B2S_NO_DUO4");
        allow_big = !on("B2S_NO_BIG");
        dynamic_units = !on("B2S_STATIC_UNITS");
        allow_pair = !on("B2S_NO_PAIR");
        // the staged-sample kernel for nperseg >= 2048 is correct but measured slower than the round-1 kernels
        // (profiles/r2_pairq_vs_round1.md): opt-in
        allow_pairq = on("B2S_PAIRQ");
        fused_sum = !on("B2S_NO_FUSED_SUM");
        allow_mixed = !on("B2S_NO_MIXED");
        sum_acc_smem = on("B2S_SUM_ACC_SMEM");
        if (const char* v = getenv("B2S_PAIR_UNITS")) pair_units = atoi(v);
        if (const char* v = getenv("B2S_PAIR_NT")) pair_nt = atoi(v);
        if (const char* v = getenv("B2S_SUM_BLOCKS")) sum_blocks = atoi(v);
        sum_dynamic = on("B2S_SUM_DYNAMIC");
        if (const char* v = getenv("B2S_PEER_TIMEOUT_MS")) peer_timeout_ms = atoi(v);
    }
};
EnvSwitches& env_mut() {
    static EnvSwitches e;        // the B2S_* environment is read once, at the first call
    return e;
}
const EnvSwitches& env() { return env_mut(); }

// One launch of an STFT kernel (either family): persistent grid sized from the
// occupancy, work units sized from the grid (b2s::plan_stft).
int launch_any_impl(const void* kern, int nt, size_t smem, int fpc, const b2s::StftArgs& a, cudaStream_t stream,
                    bool direct_table = false, bool dynamic = false) {
    DeviceInfo di;
    int dev = 0;
    int rc = device_info(di, dev);
    if (rc != B2S_OK) return rc;
    int occ = 0;
    {
        std::lock_guard<std::mutex> g(g_mu);
        KernelState& ks = g_kern[kern];
        if (ks.dev != dev) {
            if ((int)smem > di.smem_optin)
                return fail(B2S_ERR_UNSUPPORTED, "b2s: shared memory per block too small for this nperseg");
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ks.occ, kern, nt, smem);
            if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
            if (ks.occ < 1) ks.occ = 1;
            ks.dev = dev;
        }
        occ = ks.occ;
    }
    // b2s_set_reserved_sms(): leave a few SMs' worth of CTA slots to kernels of other streams (a
    // collective that should run beside the persistent grid); 0 by default
    const int reserve = (g_reserved_sms.load() < di.sm_count) ? g_reserved_sms.load() : di.sm_count - 1;
    const long long resident_ctas = (long long)(di.sm_count - reserve) * occ;
    b2s::StftParams p{};
    std::string err;
    // small launches (less than ~4 duos per resident group) keep the static schedule: the two
    // counter round trips cost more than any imbalance they could remove
    dynamic = dynamic && (a.batch * a.nframes >= 8 * resident_ctas * fpc) && !stream_capturing(stream);
    rc = b2s::plan_stft(a, fpc, resident_ctas * fpc, p, err, dynamic);
    if (rc < 0) return fail(rc, err);
    if (p.n_units == 0) return B2S_OK;
    if (dynamic) {
        rc = work_counters(dev, stream, &p.work);
        if (rc != B2S_OK) return rc;
    }
    rc = twiddles(dev, a.nperseg, direct_table, &p.tw);
    if (rc != B2S_OK) return rc;
    // persistent grid: never more CTAs than work
    const long long need = (p.n_units + fpc - 1) / fpc;
    const long long grid = (need < resident_ctas) ? need : resident_ctas;
    void* args[] = {&p};
    cudaError_t e = cudaLaunchKernel(kern, dim3((unsigned)grid), dim3((unsigned)nt), args, smem, stream);
    if (e != cudaSuccess) return cuda_fail(e, "stft kernel launch");
    return B2S_OK;
}

// the staged-sample pair kernel (b2s_pair_kernel.cuh): shared memory -- hence residency -- depends on the hop
struct PairState {
    int occ = 0;
    int max_smem = 0;
};
std::map<std::pair<const void*, long long>, PairState> g_pair;     // (kernel, device << 32 | smem) -> residency

int launch_pair_impl(const void* kern, const void* kern_mid, const void* kern_wide, int esz, const b2s::StftArgs& a,
                     cudaStream_t stream, bool dynamic) {
    using PP = b2s::PairPlan<10>;
    DeviceInfo di;
    int dev = 0;
    int rc = device_info(di, dev);
    if (rc != B2S_OK) return rc;
    // CTA shape (B2S_PAIR_NT / b2s_set_option("pair_nt") overrides).  Nothing in the main loop is CTA-wide, so
    // the shape only decides how many warps share an SM and how many registers each gets.  Measured on B200
    // (tools/microbench.py --set n1024x, profiles/r2_pair_cta_width.md):
    //   128 x 3 CTAs (164 registers)  the default;
    //   one CTA of 384 (136 registers) hop > 768: three 128-thread CTAs no longer fit their rings there;
    //   one CTA of 480 (128 registers) hop <= 128 and 257..768: 15 warps per SM buy up to 10 %.
    int nt = PP::NT;
    const bool big_launch = a.batch * a.nframes >= 64LL * di.sm_count;
    if (big_launch) {
        if (a.hop > 768) nt = PP::NT_MID;
        else if (a.hop <= 128 || a.hop > 256) nt = 480;
    }
    if (env().pair_nt > 0) nt = env().pair_nt / 32 * 32;
    if (nt < 32) nt = 32;
    if (nt > PP::NT_WIDE) nt = PP::NT_WIDE;
    while (nt > PP::NT && (int)PP::smem_bytes(a.hop, esz, nt) > di.smem_optin) nt -= 32;
    if (nt > PP::NT_MID) kern = kern_wide;
    else if (nt > PP::NT) kern = kern_mid;
    const size_t smem = PP::smem_bytes(a.hop, esz, nt);
    if ((int)smem > di.smem_optin) return fail(B2S_ERR_UNSUPPORTED, "b2s: shared memory per block too small for this hop");
    int occ = 0;
    {
        std::lock_guard<std::mutex> g(g_mu);
        PairState& ks = g_pair[std::make_pair(kern, ((long long)dev << 40) | ((long long)nt << 24) | (long long)smem)];
        if (ks.occ == 0) {
            // the attribute is the maximum any launch of this kernel may ask for
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ks.occ, kern, nt, smem);
            if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
            if (ks.occ < 1) ks.occ = 1;
        }
        occ = ks.occ;
    }
    const int fpc = nt / 32;                 // runs in flight per CTA: one per warp
    const int reserve = (g_reserved_sms.load() < di.sm_count) ? g_reserved_sms.load() : di.sm_count - 1;
    const long long resident_ctas = (long long)(di.sm_count - reserve) * occ;
    const long long resident_warps = resident_ctas * fpc;
    b2s::StftParams p{};
    std::string err;
    dynamic = dynamic && (a.batch * a.nframes >= 16 * resident_warps) && !stream_capturing(stream);
    rc = b2s::plan_stft(a, fpc, resident_warps, p, err, dynamic);
    if (rc < 0) return fail(rc, err);
    if (p.n_units == 0) return B2S_OK;
    b2s::plan_pair_units(a, resident_warps, env().pair_units, dynamic, p);
    p.ring = PP::ring_samples(a.hop);
    if (dynamic) {
        rc = work_counters(dev, stream, &p.work);
        if (rc != B2S_OK) return rc;
    }
    rc = twiddles(dev, a.nperseg, false, &p.tw);
    if (rc != B2S_OK) return rc;
    const long long need = (p.n_units + fpc - 1) / fpc;
    const long long grid = (need < resident_ctas) ? need : resident_ctas;
    void* args[] = {&p};
    cudaError_t e = cudaLaunchKernel(kern, dim3((unsigned)grid), dim3((unsigned)nt), args, smem, stream);
    if (e != cudaSuccess) return cuda_fail(e, "pair stft kernel launch");
    return B2S_OK;
}

// the staged-sample kernel for nperseg >= 2048: one CTA per SM, as many frame groups as its shared memory holds
int launch_pairq_impl(const void* kern, int group_threads, int nt_max, size_t (*smem_fn)(int, int), int esz,
                      const b2s::PairQConst& qc, const b2s::StftArgs& a, cudaStream_t stream, bool dynamic) {
    DeviceInfo di;
    int dev = 0;
    int rc = device_info(di, dev);
    if (rc != B2S_OK) return rc;
    int nt = nt_max / group_threads * group_threads;
    if (env().pair_nt > 0) nt = env().pair_nt / group_threads * group_threads;
    if (nt < group_threads) nt = group_threads;
    if (nt > nt_max) nt = nt_max / group_threads * group_threads;
    while (nt > group_threads && (long long)smem_fn(esz, nt) > di.smem_optin) nt -= group_threads;
    const size_t smem = smem_fn(esz, nt);
    if ((int)smem > di.smem_optin) return fail(B2S_ERR_UNSUPPORTED, "b2s: shared memory per block too small for this nperseg");
    {
        std::lock_guard<std::mutex> g(g_mu);
        PairState& ks = g_pair[std::make_pair(kern, ((long long)dev << 40) | ((long long)nt << 24) | (long long)smem)];
        if (ks.occ == 0) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
            ks.occ = 1;
        }
    }
    const int fpc = nt / group_threads;
    const int reserve = (g_reserved_sms.load() < di.sm_count) ? g_reserved_sms.load() : di.sm_count - 1;
    const long long resident_ctas = (long long)(di.sm_count - reserve);
    const long long resident_groups = resident_ctas * fpc;
    b2s::StftParams p{};
    std::string err;
    dynamic = dynamic && (a.batch * a.nframes >= 8 * resident_groups) && !stream_capturing(stream);
    rc = b2s::plan_stft(a, fpc, resident_groups, p, err, dynamic);
    if (rc < 0) return fail(rc, err);
    if (p.n_units == 0) return B2S_OK;
    b2s::plan_pair_units(a, resident_groups, env().pair_units, dynamic, p);
    p.ring = a.nperseg;
    if (dynamic) {
        rc = work_counters(dev, stream, &p.work);
        if (rc != B2S_OK) return rc;
    }
    rc = twiddles(dev, a.nperseg, false, &p.tw);
    if (rc != B2S_OK) return rc;
    const long long need = (p.n_units + fpc - 1) / fpc;
    const long long grid = (need < resident_ctas) ? need : resident_ctas;
    b2s::PairQConst qcc = qc;
    void* args[] = {&p, &qcc};
    cudaError_t e = cudaLaunchKernel(kern, dim3((unsigned)grid), dim3((unsigned)nt), args, smem, stream);
    if (e != cudaSuccess) return cuda_fail(e, "pairq stft kernel launch");
    return B2S_OK;
}

// direct-DFT family: one CTA per frame, grid-stride
template <typename Tin>
int launch_dft(const b2s::StftArgs& a, cudaStream_t stream) {
    DeviceInfo di;
    int dev = 0;
    int rc = device_info(di, dev);
    if (rc != B2S_OK) return rc;
    const void* kern = (const void*)b2s::dft_psd_kernel<Tin>;
    const size_t smem = b2s::dft_smem_bytes(a.nperseg);
    {
        std::lock_guard<std::mutex> g(g_mu);
        KernelState& ks = g_kern[kern];
        if (ks.dev != dev) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
            ks.dev = dev;
            ks.occ = 1;
        }
    }
    if ((int)smem > di.smem_optin) return fail(B2S_ERR_UNSUPPORTED, "b2s: nperseg too large for the direct-DFT kernel");
    b2s::DftParams p{};
    b2s::fill_dft_params(a, p);
    if (p.total_frames == 0) return B2S_OK;
    rc = twiddles(dev, a.nperseg, true, &p.tw);
    if (rc != B2S_OK) return rc;
    const long long cap = (long long)di.sm_count * 8;
    const long long grid = p.total_frames < cap ? p.total_frames : cap;
    void* args[] = {&p};
    cudaError_t e = cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(b2s::kDftThreads), args, smem, stream);
    if (e != cudaSuccess) return cuda_fail(e, "dft kernel launch");
    return B2S_OK;
}

// mixed-radix family: one CTA per frame, grid-stride
template <typename Tin>
int launch_mixed(const b2s::StftArgs& a, cudaStream_t stream) {
    DeviceInfo di;
    int dev = 0;
    int rc = device_info(di, dev);
    if (rc != B2S_OK) return rc;
    const void* kern = (const void*)b2s::mixed_psd_kernel<Tin>;
    const size_t smem = b2s::mixed_smem_bytes(a.nperseg);
    if ((int)smem > di.smem_optin) return fail(B2S_ERR_UNSUPPORTED, "b2s: nperseg too large for the mixed-radix kernel");
    {
        std::lock_guard<std::mutex> g(g_mu);
        KernelState& ks = g_kern[kern];
        if (ks.dev != dev) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
            ks.dev = dev;
            ks.occ = 1;
        }
    }
    b2s::MixedParams mp{};
    if (!b2s::fill_mixed_params(a, mp)) return fail(B2S_ERR_UNSUPPORTED, "b2s: nperseg has a prime factor above 13");
    if (mp.d.total_frames == 0) return B2S_OK;
    rc = twiddles(dev, a.nperseg, false, &mp.d.tw, true);
    if (rc != B2S_OK) return rc;
    const long long cap = (long long)di.sm_count * 8;
    const long long grid = mp.d.total_frames < cap ? mp.d.total_frames : cap;
    void* args[] = {&mp};
    cudaError_t e = cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(b2s::kMixedThreads), args, smem, stream);
    if (e != cudaSuccess) return cuda_fail(e, "mixed-radix kernel launch");
    return B2S_OK;
}

template <typename Tin>
int stft_entry(const Tin* x, long long batch, long long n, long long x_batch_stride, int nperseg, int hop,
               const float* window, int detrend, double scale, int out_mode, float db_floor, int kmin,
               int kmax, long long frame0, long long nframes, float* out, long long out_batch_stride,
               void* stream, int band_mode = 0) {
    b2s::StftArgs a{x, (int)(sizeof(Tin) == 8), batch, n, x_batch_stride, nperseg, hop, window, detrend,
                    scale, out_mode, db_floor, kmin, kmax, frame0, nframes, out, out_batch_stride, band_mode};
    // validate before touching the device so bad arguments are reported as such
    {
        std::string err;
        int rc = b2s::validate_args(a, err);
        if (rc < 0) return fail(rc, err);
    }
    const int support = b2s::nperseg_support(nperseg);
    if (support == 3 && env().allow_mixed) {
        b2s_note_kernel("mixed_psd_kernel (mixed-radix Stockham, one frame per CTA)", a);
        return launch_mixed<Tin>(a, (cudaStream_t)stream);
    }
    if (support == 2 || support == 3) {
        b2s_note_kernel("dft_psd_kernel (direct DFT)", a);
        return launch_dft<Tin>(a, (cudaStream_t)stream);
    }
    b2s::CudaLauncher L{(cudaStream_t)stream};
    const EnvSwitches& sw = env();
    L.allow_duo = sw.allow_duo;
    L.duo1024 = sw.duo1024;
    L.allow_duo4 = sw.allow_duo4;
    L.allow_big = sw.allow_big;
    L.dynamic_units = sw.dynamic_units;
    L.allow_pair = sw.allow_pair && sw.allow_duo;
    L.allow_pairq = sw.allow_pairq && sw.allow_duo;
    // the reference's call (linear power, every bin) takes the branch-free epilogue
    const bool general = (a.out_mode != B2S_OUT_LINEAR) || a.kmin != 0 || a.kmax != a.nperseg / 2;
    const int mode = a.band_mode ? b2s::EPI_BAND : (general ? b2s::EPI_GENERAL : b2s::EPI_PLAIN);
    if (sizeof(Tin) == 8) {
        if (mode == b2s::EPI_BAND) return b2s::dispatch_f64_band(a, L);
        return mode == b2s::EPI_GENERAL ? b2s::dispatch_f64_general(a, L) : b2s::dispatch_f64_plain(a, L);
    }
    if (mode == b2s::EPI_BAND) return b2s::dispatch_f32_band(a, L);
    return mode == b2s::EPI_GENERAL ? b2s::dispatch_f32_general(a, L) : b2s::dispatch_f32_plain(a, L);
}

constexpr int kMaxSumBlocks = 64;       // sweep blocks of the sum-fused kernel: bounds its scratch

// the sum-fused frame-duo kernel (per-sweep rows and their cross-sweep sum in one pass), then the
// fold over its sweep blocks.  Static schedule: plan_stft_sum makes the units fill the grid evenly.
// diagnostic overrides of plan_stft_sum's choice (b2s_set_option "sum_blocks" / "sum_dynamic")
int apply_sum_overrides(const b2s::StftArgs& a, int blocks, int dev, cudaStream_t stream, b2s::StftParams& p) {
    if (env().sum_blocks > 0) {
        long long nb = env().sum_blocks < kMaxSumBlocks ? env().sum_blocks : kMaxSumBlocks;
        if (nb > a.batch) nb = a.batch;
        const long long rows = (a.batch + nb - 1) / nb;
        blocks = (int)((a.batch + rows - 1) / rows);
        p.acc_rows = (int)rows;
        p.n_units = (long long)blocks * p.units_per_signal;
    }
    if (env().sum_dynamic && !stream_capturing(stream)) {
        const int rc = work_counters(dev, stream, &p.work);
        if (rc != B2S_OK) return rc;
    }
    return blocks;
}

struct SumKernelShape {
    const void* kern;        // the launch's kernel (running sums in tensor memory, or the twin)
    const void* twin;        // the shared-memory twin: residency is taken from it
    int nt, fpc;             // CTA threads, lane groups (duos in flight) per CTA
    size_t smem;             // dynamic shared memory (both)
    int tmem_cols;           // tensor-memory columns a CTA allocates
    int duos_per_warp;       // plan_stft_sum's unit multiple
};

int launch_sum_kernel(const b2s::StftArgs& a, const SumKernelShape& ks_in, float* sum_out, float post_scale, float* scratch,
                      cudaStream_t stream) {
    DeviceInfo di;
    int dev = 0;
    int rc = device_info(di, dev);
    if (rc != B2S_OK) return rc;
    const void* const kern = ks_in.kern;
    const void* const twin = ks_in.twin;
    const size_t smem = ks_in.smem;
    int occ = 0;
    {
        std::lock_guard<std::mutex> g(g_mu);
        KernelState& ks = g_kern[kern];
        if (ks.dev != dev) {
            // residency from the shared-memory twin (same registers bound, same shared memory): the
            // query answers 1 for a kernel that allocates tensor memory
            cudaError_t e = cudaFuncSetAttribute(twin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ks.occ, twin, ks_in.nt, smem);
            if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
            if (ks.occ < 1) ks.occ = 1;
            if (ks.occ * ks_in.tmem_cols > 512) ks.occ = 512 / ks_in.tmem_cols;
            ks.dev = dev;
        }
        occ = ks.occ;
    }
    const int reserve = (g_reserved_sms.load() < di.sm_count) ? g_reserved_sms.load() : di.sm_count - 1;
    const long long resident_ctas = (long long)(di.sm_count - reserve) * occ;
    b2s::StftParams p{};
    std::string err;
    int blocks = b2s::plan_stft_sum(a, resident_ctas * ks_in.fpc, kMaxSumBlocks, p, err, ks_in.duos_per_warp);
    if (blocks < 0) return fail(blocks, err);
    if (p.n_units == 0) return B2S_OK;
    blocks = apply_sum_overrides(a, blocks, dev, stream, p);
    if (blocks < 0) return blocks;
    p.acc = scratch;
    rc = twiddles(dev, a.nperseg, false, &p.tw);
    if (rc != B2S_OK) return rc;
    const long long need = (p.n_units + ks_in.fpc - 1) / ks_in.fpc;
    const long long grid = (need < resident_ctas) ? need : resident_ctas;
    void* args[] = {&p};
    cudaError_t e = cudaLaunchKernel(kern, dim3((unsigned)grid), dim3((unsigned)ks_in.nt), args, smem, stream);
    if (e != cudaSuccess) return cuda_fail(e, "sum-fused stft kernel launch");
    const long long elems = a.nframes * (a.nperseg / 2 + 1);
    const int block = 256;
    b2s::batch_sum_kernel<<<dim3((unsigned)((elems + block - 1) / block), 1), block, 0, stream>>>(
        scratch, elems, blocks, blocks, elems, sum_out, post_scale);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "batch_sum_kernel launch");
    return B2S_OK;
}

// the sum-fused frame-duo kernels (per-sweep rows and their cross-sweep sum in one pass), then the
// fold over the sweep blocks.  Static schedule: plan_stft_sum makes the units fill the grid evenly.
int launch_duo_sum(const b2s::StftArgs& a, int slots, float* sum_out, float post_scale, float* scratch,
                   cudaStream_t stream) {
    const int tmem = env().sum_acc_smem ? 0 : 1;
    SumKernelShape ks{};
    if (a.nperseg == 2048) {
        // (the tensor-memory kernel is launched without the twin's shared-memory sums)
        using DP = b2s::Duo4Plan<11>;
        ks.twin = b2s::duo4_sum_kernel_for(11, a.x_is_f64, slots, 0);
        ks.kern = tmem ? b2s::duo4_sum_kernel_for(11, a.x_is_f64, slots, 1) : ks.twin;
        ks.nt = DP::NT, ks.fpc = DP::FPC, ks.smem = tmem ? DP::SMEM : DP::SUM_SMEM, ks.tmem_cols = DP::TMEM_COLS;
        ks.duos_per_warp = 1;
    } else if (a.nperseg == 256) {
        using DP = b2s::Duo256Plan;
        ks.twin = b2s::duo256_sum_kernel_for(a.x_is_f64, slots, 0);
        ks.kern = tmem ? b2s::duo256_sum_kernel_for(a.x_is_f64, slots, 1) : ks.twin;
        ks.nt = DP::NT, ks.fpc = DP::FPC, ks.smem = DP::SUM_SMEM, ks.tmem_cols = DP::TMEM_COLS, ks.duos_per_warp = 4;
    } else {
        using DP = b2s::DuoPlan;
        ks.twin = b2s::duo_sum_kernel_for(a.x_is_f64, slots, 0);
        ks.kern = tmem ? b2s::duo_sum_kernel_for(a.x_is_f64, slots, 1) : ks.twin;
        ks.nt = DP::NT, ks.fpc = DP::FPC, ks.smem = b2s::DuoSumPlan::SMEM, ks.tmem_cols = b2s::DuoSumPlan::TMEM_COLS;
        ks.duos_per_warp = 2;
    }
    if (!ks.kern || !ks.twin) return fail(B2S_ERR_UNSUPPORTED, "b2s: no sum-fused kernel for this hop");
    return launch_sum_kernel(a, ks, sum_out, post_scale, scratch, stream);
}

// nperseg 1024: the SUM mode of the staged-sample pair kernel (one pair of frames per warp over a block of
// sweeps), then the fold over the sweep blocks.  Static schedule, planned like launch_duo_sum's.
int launch_pair_sum(const b2s::StftArgs& a, float* sum_out, float post_scale, float* scratch, cudaStream_t stream) {
    using PP = b2s::PairPlan<10>;
    DeviceInfo di;
    int dev = 0;
    int rc = device_info(di, dev);
    if (rc != B2S_OK) return rc;
    const int tmem = env().sum_acc_smem ? 0 : 1;
    const void* twin = b2s::pair_sum_kernel_for(a.x_is_f64, 0);
    const void* kern = tmem ? b2s::pair_sum_kernel_for(a.x_is_f64, 1) : twin;
    const int esz = a.x_is_f64 ? 8 : 4;
    // (the tensor-memory kernel does not touch the twin's shared-memory sums: without them three CTAs share an
    //  SM up to hop 256 where the twin's 18 KB more leave room for two)
    const size_t smem = tmem ? PP::smem_bytes(a.hop, esz, PP::NT) : PP::sum_smem_bytes(a.hop, esz);
    const int smem_cap = di.smem_optin - 64;     // the tensor-memory kernel keeps its base-address slot in static shared memory
    if ((int)smem > smem_cap) return fail(B2S_ERR_UNSUPPORTED, "b2s: shared memory per block too small for this hop");
    int occ = 0;
    {
        std::lock_guard<std::mutex> g(g_mu);
        PairState& ks = g_pair[std::make_pair(kern, ((long long)dev << 40) | ((long long)PP::NT << 24) | (long long)smem)];
        if (ks.occ == 0) {
            // residency from the shared-memory twin (same register bound, same shared memory): the query
            // answers 1 for a kernel that allocates tensor memory
            cudaError_t e = cudaFuncSetAttribute(twin, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ks.occ, twin, PP::NT, smem);
            if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
            if (ks.occ < 1) ks.occ = 1;
            if (ks.occ * PP::TMEM_COLS > 512) ks.occ = 512 / PP::TMEM_COLS;
        }
        occ = ks.occ;
    }
    const int fpc = PP::NT / 32;
    const int reserve = (g_reserved_sms.load() < di.sm_count) ? g_reserved_sms.load() : di.sm_count - 1;
    const long long resident_ctas = (long long)(di.sm_count - reserve) * occ;
    b2s::StftParams p{};
    std::string err;
    // The split into sweep blocks fixes the order of the additions.  It is planned from the residency a launch
    // on float samples would have (three CTAs per SM where their rings fit), so that float64 samples -- bigger
    // rings, sometimes one CTA fewer per SM -- give the same sums bit for bit.
    int occ_plan = 3;
    while (occ_plan > 1 && (long long)occ_plan * (PP::smem_bytes(a.hop, 4, PP::NT) + 1024) > (long long)di.smem_optin + 1024) --occ_plan;
    const long long plan_groups = (long long)(di.sm_count - reserve) * occ_plan * fpc;
    int blocks = b2s::plan_stft_sum(a, plan_groups, kMaxSumBlocks, p, err, 1);
    if (blocks < 0) return fail(blocks, err);
    if (p.n_units == 0) return B2S_OK;
    blocks = apply_sum_overrides(a, blocks, dev, stream, p);
    if (blocks < 0) return blocks;
    p.acc = scratch;
    p.ring = PP::ring_samples(a.hop);
    rc = twiddles(dev, a.nperseg, false, &p.tw);
    if (rc != B2S_OK) return rc;
    const long long need = (p.n_units + fpc - 1) / fpc;
    const long long grid = (need < resident_ctas) ? need : resident_ctas;
    void* args[] = {&p};
    cudaError_t e = cudaLaunchKernel(kern, dim3((unsigned)grid), dim3((unsigned)PP::NT), args, smem, stream);
    if (e != cudaSuccess) return cuda_fail(e, "sum-fused pair kernel launch");
    const long long elems = a.nframes * (a.nperseg / 2 + 1);
    const int block = 256;
    b2s::batch_sum_kernel<<<dim3((unsigned)((elems + block - 1) / block), 1), block, 0, stream>>>(
        scratch, elems, blocks, blocks, elems, sum_out, post_scale);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "batch_sum_kernel launch");
    return B2S_OK;
}

template <typename Tin>
int stft_sum_entry(const Tin* x, long long batch, long long n, long long x_batch_stride, int nperseg, int hop,
                          const float* window, int detrend, double scale, long long frame0, long long nframes,
                          float* out, long long out_batch_stride, float* sum_out, float post_scale, float* scratch,
                          void* stream) {
    b2s::StftArgs a{x, (int)(sizeof(Tin) == 8), batch, n, x_batch_stride, nperseg, hop, window, detrend,
                    scale, B2S_OUT_LINEAR, 0.f, 0, nperseg / 2, frame0, nframes, out, out_batch_stride, 0};
    {
        std::string err;
        int rc = b2s::validate_args(a, err);
        if (rc < 0) return fail(rc, err);
    }
    if (!sum_out || !scratch) return fail(B2S_ERR_BAD_ARG, "b2s_stft_psd_sum: sum_out and scratch are required");
    if (nframes < 1 || batch < 1) return fail(B2S_ERR_BAD_ARG, "b2s_stft_psd_sum: nothing to sum");
    const long long elems = nframes * (nperseg / 2 + 1);
    int slots = (!env().fused_sum || !env().allow_duo || batch < 2) ? 0 : b2s::duo_slots(a, b2s::ilog2_exact(nperseg));
    if (slots > 8) slots = 0;           // hop 7/8 nperseg and nperseg: per-sweep frame-duo kernel, no sum-fused variant
    if (nperseg == 256 && env().fused_sum && env().allow_duo && batch >= 2) slots = b2s::duo256_sum_slots(a, 8);
    // (nperseg 4096: the fused form measured slower than kernel + two-pass sum, b2s_inst_sum4.cu)
    if (nperseg == 2048 && env().fused_sum && env().allow_duo && env().allow_duo4 && !env().allow_pairq &&
        batch >= 2) {
        slots = b2s::duo4_slots(a);         // the per-sweep call takes the four-step frame-duo kernel for these hops
        if (slots != 2 && slots != 4 && slots != 8) slots = 0;
    }
    if (slots) {
        b2s_note_kernel(nperseg == 256 ? "stft_psd_duo256_sum_kernel (frame duo, 8 lanes, running cross-sweep sums in tensor memory) + fold"
                        : nperseg >= 2048 ? "stft_psd_duo4_sum_kernel (four-step frame duo, running cross-sweep sums in tensor memory) + fold"
                                          : "stft_psd_duo_sum_kernel (frame duo, running cross-sweep sums in tensor memory) + fold", a);
        return launch_duo_sum(a, slots, sum_out, post_scale, scratch, (cudaStream_t)stream);
    }
    b2s::CudaLauncher Lsel{(cudaStream_t)stream};
    Lsel.allow_duo = env().allow_duo;
    Lsel.duo1024 = env().duo1024;
    Lsel.allow_duo4 = env().allow_duo4;
    Lsel.allow_pair = env().allow_pair && env().allow_duo;
    if (env().fused_sum && batch >= 2 && b2s::pair_preferred(a, Lsel)) {
        b2s_note_kernel("stft_psd_pair_sum_kernel (staged samples, running cross-sweep sums in tensor memory) + fold", a);
        return launch_pair_sum(a, sum_out, post_scale, scratch, (cudaStream_t)stream);
    }
    // every other shape: the per-sweep kernel of its family, then the two-pass sum
    int rc = stft_entry<Tin>(x, batch, n, x_batch_stride, nperseg, hop, window, detrend, scale, B2S_OUT_LINEAR, 0.f, 0,
                             nperseg / 2, frame0, nframes, out, out_batch_stride, stream);
    if (rc != B2S_OK) return rc;
    return b2s_batch_sum_f32(out, batch, elems, out_batch_stride, sum_out, scratch, post_scale, stream);
}

}  // namespace

void b2s_note_kernel(const char* family, const b2s::StftArgs& a) {
    g_last_kernel = std::string(family) + " nperseg " + std::to_string(a.nperseg) + " hop " + std::to_string(a.hop) +
                    (a.x_is_f64 ? " f64" : " f32");
}

int b2s_launch_any(const void* kern, int nt, size_t smem, int fpc, const b2s::StftArgs& a, cudaStream_t stream,
                   bool dynamic) {
    return launch_any_impl(kern, nt, smem, fpc, a, stream, false, dynamic);
}

int b2s_launch_pairq(const void* kern, int group_threads, int nt_max, size_t (*smem_fn)(int, int), int esz,
                     const b2s::PairQConst& qc, const b2s::StftArgs& a, cudaStream_t stream, bool dynamic) {
    return launch_pairq_impl(kern, group_threads, nt_max, smem_fn, esz, qc, a, stream, dynamic);
}

int b2s_launch_pair(const void* kern, const void* kern_mid, const void* kern_wide, int esz, const b2s::StftArgs& a,
                    cudaStream_t stream, bool dynamic) {
    return launch_pair_impl(kern, kern_mid, kern_wide, esz, a, stream, dynamic);
}

extern "C" {

int b2s_version(void) { return B2S_ABI_VERSION; }

int b2s_set_reserved_sms(int n) {
    const int old = g_reserved_sms.load();
    g_reserved_sms.store(n < 0 ? 0 : n);
    return old;
}

const char* b2s_last_error(void) { return g_err.c_str(); }

const char* b2s_last_kernel(void) { return g_last_kernel.c_str(); }

int b2s_set_option(const char* name, int value) {
    if (!name) return fail(B2S_ERR_BAD_ARG, "b2s_set_option: null name");
    EnvSwitches& e = env_mut();
    const std::string n(name);
    const bool on = value != 0;
    if (n == "no_duo") e.allow_duo = !on;
    else if (n == "duo1024") e.duo1024 = on;
    else if (n == "no_duo4") e.allow_duo4 = !on;
    else if (n == "no_big") e.allow_big = !on;
    else if (n == "static_units") e.dynamic_units = !on;
    else if (n == "no_pair") e.allow_pair = !on;
    else if (n == "no_pairq") e.allow_pairq = !on;
    else if (n == "no_fused_sum") e.fused_sum = !on;
    else if (n == "no_mixed") e.allow_mixed = !on;
    else if (n == "sum_acc_smem") e.sum_acc_smem = on;
    else if (n == "pair_units") e.pair_units = value;
    else if (n == "pair_nt") e.pair_nt = value;
    else if (n == "sum_blocks") e.sum_blocks = value;
    else if (n == "sum_dynamic") e.sum_dynamic = on;
    else if (n == "peer_timeout_ms") e.peer_timeout_ms = value;
    else return fail(B2S_ERR_BAD_ARG, "b2s_set_option: unknown option " + n);
    return B2S_OK;
}

int b2s_nperseg_support(int nperseg) { return b2s::nperseg_support(nperseg); }

long long b2s_frame_count(long long n, int nperseg, int hop) {
    if (nperseg < 1 || hop < 1) return 0;
    return b2s::frames_available(n, nperseg, hop);
}

int b2s_stft_psd_f32(const float* x, long long batch, long long n, long long x_batch_stride, int nperseg,
                     int hop, const float* window, int detrend, double scale, int out_mode, float db_floor,
                     int kmin, int kmax, long long frame0, long long nframes, float* out,
                     long long out_batch_stride, void* stream) {
    return stft_entry<float>(x, batch, n, x_batch_stride, nperseg, hop, window, detrend, scale, out_mode,
                             db_floor, kmin, kmax, frame0, nframes, out, out_batch_stride, stream);
}

int b2s_stft_psd_f64(const double* x, long long batch, long long n, long long x_batch_stride, int nperseg,
                     int hop, const float* window, int detrend, double scale, int out_mode, float db_floor,
                     int kmin, int kmax, long long frame0, long long nframes, float* out,
                     long long out_batch_stride, void* stream) {
    return stft_entry<double>(x, batch, n, x_batch_stride, nperseg, hop, window, detrend, scale, out_mode,
                              db_floor, kmin, kmax, frame0, nframes, out, out_batch_stride, stream);
}

int b2s_stft_band_power_f32(const float* x, long long batch, long long n, long long x_batch_stride, int nperseg,
                            int hop, const float* window, int detrend, double scale, int kmin, int kmax,
                            long long frame0, long long nframes, float* out, long long out_batch_stride,
                            void* stream) {
    return stft_entry<float>(x, batch, n, x_batch_stride, nperseg, hop, window, detrend, scale, B2S_OUT_LINEAR,
                             0.f, kmin, kmax, frame0, nframes, out, out_batch_stride, stream, 1);
}

int b2s_stft_band_power_f64(const double* x, long long batch, long long n, long long x_batch_stride, int nperseg,
                            int hop, const float* window, int detrend, double scale, int kmin, int kmax,
                            long long frame0, long long nframes, float* out, long long out_batch_stride,
                            void* stream) {
    return stft_entry<double>(x, batch, n, x_batch_stride, nperseg, hop, window, detrend, scale, B2S_OUT_LINEAR,
                              0.f, kmin, kmax, frame0, nframes, out, out_batch_stride, stream, 1);
}

static const int kSumSlabRows = 128;    // measured best on B200 with 16 loads in flight (tools/ubench/bsum)

long long b2s_batch_sum_scratch_elems(long long batch, long long elems) {
    if (batch <= kSumSlabRows) return 0;
    const long long slabs = (batch + kSumSlabRows - 1) / kSumSlabRows;
    return slabs * elems;
}

int b2s_batch_sum_f32(const float* in, long long batch, long long elems, long long in_batch_stride,
                      float* out, float* scratch, float post_scale, void* stream) {
    if (!in || !out || batch < 1 || elems < 1 || in_batch_stride < elems || batch > 0x7fffffffLL)
        return fail(B2S_ERR_BAD_ARG, "b2s_batch_sum_f32: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int block = 256;
    const unsigned gx = (unsigned)((elems + block - 1) / block);
    if (batch <= kSumSlabRows) {
        b2s::batch_sum_kernel<<<dim3(gx, 1), block, 0, st>>>(in, in_batch_stride, (int)batch, (int)batch, elems,
                                                              out, post_scale);
    } else {
        if (!scratch) return fail(B2S_ERR_BAD_ARG, "b2s_batch_sum_f32: scratch required for batch > 64");
        const long long slabs = (batch + kSumSlabRows - 1) / kSumSlabRows;
        b2s::batch_sum_kernel<<<dim3(gx, (unsigned)slabs), block, 0, st>>>(in, in_batch_stride, (int)batch,
                                                                          kSumSlabRows, elems, scratch, 1.0f);
        b2s::batch_sum_kernel<<<dim3(gx, 1), block, 0, st>>>(scratch, elems, (int)slabs, (int)slabs, elems, out,
                                                              post_scale);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "batch_sum_kernel launch");
    return B2S_OK;
}


long long b2s_stft_psd_sum_scratch_elems(long long batch, long long elems) {
    if (batch < 1 || elems < 1) return 0;
    const long long fused = (batch < kMaxSumBlocks ? batch : kMaxSumBlocks) * elems;
    const long long two_pass = b2s_batch_sum_scratch_elems(batch, elems);
    return fused > two_pass ? fused : two_pass;
}

int b2s_stft_psd_sum_f32(const float* x, long long batch, long long n, long long x_batch_stride, int nperseg, int hop,
                         const float* window, int detrend, double scale, long long frame0, long long nframes,
                         float* out, long long out_batch_stride, float* sum_out, float post_scale, float* scratch,
                         void* stream) {
    return stft_sum_entry<float>(x, batch, n, x_batch_stride, nperseg, hop, window, detrend, scale, frame0, nframes,
                                 out, out_batch_stride, sum_out, post_scale, scratch, stream);
}

int b2s_stft_psd_sum_f64(const double* x, long long batch, long long n, long long x_batch_stride, int nperseg, int hop,
                         const float* window, int detrend, double scale, long long frame0, long long nframes,
                         float* out, long long out_batch_stride, float* sum_out, float post_scale, float* scratch,
                         void* stream) {
    return stft_sum_entry<double>(x, batch, n, x_batch_stride, nperseg, hop, window, detrend, scale, frame0, nframes,
                                  out, out_batch_stride, sum_out, post_scale, scratch, stream);
}


int b2s_peer_allreduce_f32(const unsigned long long* peer_bufs, const unsigned long long* peer_pads, int world,
                           int rank, unsigned int epoch, long long elems, float* out, float post_scale, void* stream) {
    return b2s_peer_allreduce_ex_f32(peer_bufs, peer_pads, world, rank, epoch, elems, out, post_scale, 0, stream);
}

int b2s_peer_allreduce_ex_f32(const unsigned long long* peer_bufs, const unsigned long long* peer_pads, int world,
                              int rank, unsigned int epoch, long long elems, float* out, float post_scale, int flags,
                              void* stream) {
    if (!peer_bufs || !peer_pads || !out || world < 1 || world > 16 || rank < 0 || rank >= world || elems < 1)
        return fail(B2S_ERR_BAD_ARG, "b2s_peer_allreduce_f32: bad argument");
    b2s::PeerPtrs pp{};
    int vec_ok = (reinterpret_cast<uintptr_t>(out) % 16 == 0) ? 1 : 0;
    for (int r = 0; r < world; ++r) {
        pp.buf[r] = reinterpret_cast<const float*>(peer_bufs[r]);
        pp.pad[r] = reinterpret_cast<unsigned*>(peer_pads[r]) + (epoch & 1u) * 16u;
        if (peer_bufs[r] % 16) vec_ok = 0;
    }
    DeviceInfo di;
    int dev = 0;
    int rc = device_info(di, dev);
    if (rc != B2S_OK) return rc;
    const int block = 256;
    long long want = (elems + block - 1) / block;
    // every CTA waits for the peers, so never more CTAs than can be resident; with SMs reserved
    // (b2s_set_reserved_sms: the all-reduce runs beside a persistent STFT grid of the next step) only
    // as many as those SMs' CTA slots of the STFT kernels (3 per SM)
    const int reserve = g_reserved_sms.load();
    const long long cap = (reserve > 0 && reserve < di.sm_count) ? 3LL * reserve : di.sm_count;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    int* err_flag = nullptr;
    rc = peer_err_flag(dev, &err_flag);
    if (rc != B2S_OK) return rc;
    const unsigned long long timeout_ns = (unsigned long long)(env().peer_timeout_ms > 0 ? env().peer_timeout_ms : 120000) * 1000000ULL;
    if (flags & B2S_PEER_CORESIDENT) {
        // one warp per SM: fits the registers three STFT CTAs leave, never displaces one of them
        long long want32 = (elems / 2 + 31) / 32;
        const unsigned g32 = (unsigned)(want32 < di.sm_count ? (want32 < 1 ? 1 : want32) : di.sm_count);
        b2s::peer_allreduce_warp_kernel<<<g32, 32, 0, (cudaStream_t)stream>>>(pp, world, rank, epoch, elems, out,
                                                                              post_scale, vec_ok, timeout_ns, err_flag);
    } else {
        b2s::peer_allreduce_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(pp, world, rank, epoch, elems, out,
                                                                             post_scale, vec_ok, timeout_ns, err_flag);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "peer_allreduce_kernel launch");
    return B2S_OK;
}

long long b2s_band_sums_scratch_elems(void) { return 256LL * (b2s::kMaxBands + 1); }

int b2s_band_sums_f32(const float* s, long long frames, int bins, const int* k0, const int* k1, int nb, double* out,
                      double* scratch, void* stream) {
    if (!s || !out || !scratch || frames < 1 || bins < 1 || nb < 0 || nb > b2s::kMaxBands || (nb > 0 && (!k0 || !k1)))
        return fail(B2S_ERR_BAD_ARG, "b2s_band_sums_f32: bad argument");
    b2s::BandRanges br{};
    for (int b = 0; b < nb; ++b) {
        if (k0[b] < 0 || k1[b] > bins || k0[b] > k1[b])
            return fail(B2S_ERR_BAD_ARG, "b2s_band_sums_f32: band outside [0, bins]");
        br.k0[b] = k0[b];
        br.k1[b] = k1[b];
    }
    br.k0[nb] = 0;
    br.k1[nb] = bins;
    const int blocks = (int)(frames < 256 ? frames : 256);
    cudaStream_t st = (cudaStream_t)stream;
    b2s::band_sums_kernel<<<blocks, 256, 0, st>>>(s, frames, bins, br, nb, scratch);
    b2s::band_sums_fold_kernel<<<1, 32, 0, st>>>(scratch, blocks, nb, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "band_sums kernels");
    return B2S_OK;
}

int b2s_peer_allreduce_status(void* stream) {
    DeviceInfo di;
    int dev = 0;
    int rc = device_info(di, dev);
    if (rc != B2S_OK) return rc;
    int* err_flag = nullptr;
    rc = peer_err_flag(dev, &err_flag);
    if (rc != B2S_OK) return rc;
    int host = 0;
    cudaError_t e = cudaMemcpyAsync(&host, err_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "b2s_peer_allreduce_status");
    if (host == 0) return B2S_OK;
    cudaMemsetAsync(err_flag, 0, sizeof(int), (cudaStream_t)stream);
    return fail(B2S_ERR_TIMEOUT, "b2s_peer_allreduce_f32: rank " + std::to_string(host - 1) +
                                     " did not announce its partial within the time-out (B2S_PEER_TIMEOUT_MS); the result "
                                     "of that reduce was not written");
}

int b2s_display_scale_f32(const float* s, long long elems, int log_scale, float global_max, float* out,
                          unsigned int* scratch, void* stream) {
    if (!s || !out || !scratch || elems < 1) return fail(B2S_ERR_BAD_ARG, "b2s_display_scale_f32: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceInfo di;
    int dev = 0;
    int rc = device_info(di, dev);
    if (rc != B2S_OK) return rc;
    const int block = 256;
    long long want = (elems + block - 1) / block;
    const unsigned grid = (unsigned)(want < (long long)di.sm_count * 8 ? want : (long long)di.sm_count * 8);
    b2s::minmax_init_kernel<<<1, 1, 0, st>>>(scratch);
    b2s::minmax_kernel<<<grid, block, 0, st>>>(s, elems, scratch);
    b2s::display_scale_kernel<<<grid, block, 0, st>>>(s, elems, scratch, log_scale ? 1 : 0, global_max, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "display_scale kernels");
    return B2S_OK;
}

}  // extern "C"

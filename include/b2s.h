/*
 * b2s.h -- C ABI of libb200stft.so, the B200 (sm_100a) spectrogram engine.
 *
 * This is the drop-in boundary for ONE path of Karmotr1ne/Spectrogram-Generator:
 * the per-sweep spectrogram
 *
 *     f, t, Sxx = spectrogram(data, fs=fs, nperseg=nperseg,
 *                             scaling="density", mode="psd")
 *
 * made at /root/reference/PlotEngine.py:113 (plot) and :232 (HMM features),
 * imported at PlotEngine.py:8.  The reference's boundary is that Python call
 * (there is no FFI in the reference); the entry points below are what a
 * Python/ctypes shim behind that call binds, and each one
 * names the reference / SciPy code whose work it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; every *device* pointer is owned by the
 *    caller (PyTorch tensors in the shipped host code);
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    all work is enqueued on it and nothing synchronises unless stated;
 *  - return value: 0 = success, <0 = error (B2S_ERR_*), text via
 *    b2s_last_error() (thread-local);
 *  - the library keeps one small device cache: twiddle tables keyed by
 *    (device, nperseg).  It allocates nothing else persistent.
 *  - there is no CPU fallback: without a CUDA device every compute entry
 *    returns B2S_ERR_CUDA.
 */
#ifndef B2S_H_
#define B2S_H_

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_ABI_VERSION 1

#define B2S_OK 0
#define B2S_ERR_BAD_ARG (-1)     /* invalid argument (mirrors SciPy's ValueError cases) */
#define B2S_ERR_UNSUPPORTED (-2) /* nperseg not handled by this entry */
#define B2S_ERR_CUDA (-3)        /* CUDA runtime error; see b2s_last_error() */
#define B2S_ERR_TIMEOUT (-4)     /* b2s_peer_allreduce_status: a peer missed the all-reduce's time-out */

#define B2S_OUT_LINEAR 0 /* PSD / power spectrum, linear */
#define B2S_OUT_DB 1     /* 10*log10(max(S, db_floor)) */

int b2s_version(void);

/* Size the persistent grids of the STFT kernels for (SM count - n) SMs, so that kernels of other
 * streams -- e.g. the NCCL all-reduce of the partial mean spectrogram -- find free CTA slots while
 * a launch is running.  Process-wide; returns the previous value.  Default 0.  (No counterpart in
 * the reference: it has no device or collective layer) */
int b2s_set_reserved_sms(int n);

/* Diagnostic switches of the kernel selection.  Their defaults come from the
 * B2S_* environment variables, read ONCE at the first call into the library (never on the per-call
 * path); this entry changes one at run time: "no_duo", "duo1024", "no_duo4", "no_big", "no_pair",
 * "no_pairq", "no_mixed", "static_units", "no_fused_sum", "sum_acc_smem", "sum_dynamic" (value 0 / 1),
 * "pair_units" (work units per resident warp of the staged-sample kernel), "pair_nt" (its CTA width),
 * "sum_blocks" (sweep blocks of the sum-fused kernels) and "peer_timeout_ms" -- 0 = default each.
 * Not needed in normal use; the GPU tests
 * use it to compare kernel variants bit for bit.  Returns B2S_OK or B2S_ERR_BAD_ARG.  (No
 * counterpart in the reference.) */
int b2s_set_option(const char* name, int value);

/* Name of the STFT kernel family the calling thread launched last ("none" before the first launch),
 * with nperseg / hop / sample type -- what bench.py prints next to each measured shape. */
const char* b2s_last_kernel(void);

/* One-shot all-reduce (sum) of `elems` floats over NVLink peer memory: every rank has written its
 * partial into a buffer that is mapped into all peers (peer_bufs[r] = the address of rank r's
 * partial in THIS process, e.g. from torch's symmetric-memory rendezvous), and owns a zero-
 * initialised signal pad of at least 128 bytes mapped likewise (peer_pads[r]).  One launch per rank
 * on `stream`: announce, wait for every peer's announcement, add the partials in rank order
 * (bit-identical on all ranks), write post_scale * sum to `out`.  `epoch` starts at 1 and grows by
 * one per call on every rank; the partials of odd and even epochs must live in different buffers.
 * Used for the mean-spectrogram all-reduce; the reference has no counterpart. */
int b2s_peer_allreduce_f32(const unsigned long long* peer_bufs, const unsigned long long* peer_pads, int world,
                           int rank, unsigned int epoch, long long elems, float* out, float post_scale,
                           void* stream);

/* The same, with flags: B2S_PEER_CORESIDENT launches the reduce as one-warp CTAs of <= 32 registers, one per
 * SM, meant to run on a side stream BESIDE the persistent STFT grid of the next step: such a CTA fits the
 * registers three STFT CTAs leave on an SM, so it never displaces one of them (no b2s_set_reserved_sms
 * needed).  Same additions in the same order: bit-identical to the plain entry. */
#define B2S_PEER_CORESIDENT 1
int b2s_peer_allreduce_ex_f32(const unsigned long long* peer_bufs, const unsigned long long* peer_pads, int world,
                              int rank, unsigned int epoch, long long elems, float* out, float post_scale, int flags,
                              void* stream);

/* The all-reduce waits for a late peer for B2S_PEER_TIMEOUT_MS (default 120 000) of wall-clock time
 * (%globaltimer), polling with a nanosleep back-off; every rank must therefore enqueue its call
 * within that window.  A kernel whose wait expires does NOT trap: it leaves `out` unwritten and
 * raises a per-device flag.  This entry synchronises `stream`, reads the flag and clears it:
 * B2S_OK, or B2S_ERR_TIMEOUT with the missing rank in b2s_last_error(). */
int b2s_peer_allreduce_status(void* stream);
const char* b2s_last_error(void);

/* 1 if `nperseg` runs on the fused radix-16 Stockham kernels (powers of two in
 * [32, 16384]), 3 if it runs on the mixed-radix kernel (other lengths >= 256 whose prime
 * factors are all <= 13, e.g. 1000, 2000, 4800, 8000), 2 if it runs on the direct-DFT kernel
 * (every other length in 1..16384: the GUI spin box allows any integer 32..8192,
 * GUI.py:87-89, and SciPy clamps nperseg to len(x), _spectral_py.py:2443-2447), 0 if
 * unsupported. */
int b2s_nperseg_support(int nperseg);

/* Frames SciPy produces for a signal of n samples: (n - nperseg)//hop + 1
 * (sliding_window_view(...)[..., ::step, :], _spectral_py.py:2377-2381). */
long long b2s_frame_count(long long n, int nperseg, int hop);

/*
 * The hot path.  Replaces, for real input, `_fft_helper` (_spectral_py.py:
 * 2346-2397: framing, detrend, window, rfft) and the PSD epilogue of
 * `_spectral_helper` (_spectral_py.py:2313-2322: conj(X)*X, *= scale, one-sided
 * doubling) -- i.e. everything the call at PlotEngine.py:113/232 computes
 * except the two axis arrays (host side, bit-exact recipes).
 *
 *   x                 device, [batch][n] samples; signal b starts at x + b*x_batch_stride
 *   nperseg, hop      frame length and step (hop = nperseg - noverlap)
 *   window            device, nperseg fp32 taps (built on the host in float64
 *                     with SciPy's recipe, rounded once)
 *   detrend           0 = False, 1 = 'constant' (per-frame mean removal,
 *                     _signaltools.py:4288-4290)
 *   scale             1/(fs*sum(w^2)) for scaling='density', 1/sum(w)^2 for
 *                     'spectrum' (_spectral_py.py:2274-2277), computed in double
 *   out_mode/db_floor B2S_OUT_LINEAR, or B2S_OUT_DB with a linear floor
 *   kmin,kmax         inclusive bin crop (0, nperseg/2 for all bins); the
 *                     reference masks bins to [fmin,fmax] right after the call
 *                     (PlotEngine.py:114-115)
 *   frame0,nframes    frame range [frame0, frame0+nframes) of every signal
 *                     (time-chunking: a rank computes its own range; frame j
 *                     covers samples [j*hop, j*hop+nperseg))
 *   out               device, [batch][nframes][kmax-kmin+1] fp32; signal b at
 *                     out + b*out_batch_stride.  [frame][bin] is the layout of
 *                     SciPy's own result buffer (Sxx is its transposed view).
 */
int b2s_stft_psd_f32(const float* x, long long batch, long long n, long long x_batch_stride,
                     int nperseg, int hop, const float* window, int detrend, double scale,
                     int out_mode, float db_floor, int kmin, int kmax,
                     long long frame0, long long nframes,
                     float* out, long long out_batch_stride, void* stream);

/* Same, float64 samples in device memory (converted to fp32 on load; SciPy's
 * dtype rule `result_type(x, complex64)`, _spectral_py.py:2169, is honoured by
 * the host shim, which widens the fp32 result when x is float64). */
int b2s_stft_psd_f64(const double* x, long long batch, long long n, long long x_batch_stride,
                     int nperseg, int hop, const float* window, int detrend, double scale,
                     int out_mode, float db_floor, int kmin, int kmax,
                     long long frame0, long long nframes,
                     float* out, long long out_batch_stride, void* stream);

/*
 * Fused band-power feature (PlotEngine._calculate_features, PlotEngine.py:229-239):
 * the same spectrogram, but instead of storing [frame][bin] the bins kmin..kmax of
 * every frame are summed in the epilogue -- out is [batch][nframes] fp32 (signal b at
 * out + b*out_batch_stride).  The HMM path needs only these F numbers, so the
 * 4*F*K bytes of spectrogram output are never written.  log10 and the first
 * difference (PlotEngine.py:240-242) stay on the host (F values).
 */
int b2s_stft_band_power_f32(const float* x, long long batch, long long n, long long x_batch_stride,
                            int nperseg, int hop, const float* window, int detrend, double scale,
                            int kmin, int kmax, long long frame0, long long nframes,
                            float* out, long long out_batch_stride, void* stream);
int b2s_stft_band_power_f64(const double* x, long long batch, long long n, long long x_batch_stride,
                            int nperseg, int hop, const float* window, int detrend, double scale,
                            int kmin, int kmax, long long frame0, long long nframes,
                            float* out, long long out_batch_stride, void* stream);

/*
 * Cross-sweep sum / mean (BASELINE config 2; the reference has no code for it,
 * SURVEY.md 8 a-15): out[e] = post_scale * sum_b in[b*in_batch_stride + e],
 * summed in a fixed order (deterministic).  `scratch` must hold
 * b2s_batch_sum_scratch_elems(batch, elems) floats (may be NULL when that is 0).
 */
long long b2s_batch_sum_scratch_elems(long long batch, long long elems);
int b2s_batch_sum_f32(const float* in, long long batch, long long elems, long long in_batch_stride,
                      float* out, float* scratch, float post_scale, void* stream);

/*
 * Per-sweep spectrograms AND their cross-sweep sum in one pass (BASELINE config 2:
 * "per-sweep spectrograms + mean spectrogram").  Arguments as
 * b2s_stft_psd_f32 with linear power and every bin; in addition
 *
 *   sum_out     device, [nframes][nperseg/2+1] fp32:
 *               post_scale * sum_b out[b][frame][bin], added in a fixed order
 *               (sweep order inside a block of sweeps, then block order)
 *   scratch     device, b2s_stft_psd_sum_scratch_elems(batch, nframes*(nperseg/2+1)) floats
 *
 * For nperseg 512 with hop 64 / 128 / 256, nperseg 256 with any even hop, nperseg 1024 with
 * any hop that is a multiple of 4 (rows and window 16-byte aligned) and nperseg 2048 with hop
 * 256 / 512 / 1024 one kernel walks a frame
 * pair over a block of sweeps and keeps the running sums on chip (tensor memory), so the
 * [batch][nframes][bins] result is written once and never read back; every other shape runs
 * b2s_stft_psd_* followed by b2s_batch_sum_f32.  The per-sweep rows are bit-identical to
 * b2s_stft_psd_*'s either way.
 */
long long b2s_stft_psd_sum_scratch_elems(long long batch, long long elems);
int b2s_stft_psd_sum_f32(const float* x, long long batch, long long n, long long x_batch_stride,
                         int nperseg, int hop, const float* window, int detrend, double scale,
                         long long frame0, long long nframes, float* out, long long out_batch_stride,
                         float* sum_out, float post_scale, float* scratch, void* stream);
int b2s_stft_psd_sum_f64(const double* x, long long batch, long long n, long long x_batch_stride,
                         int nperseg, int hop, const float* window, int detrend, double scale,
                         long long frame0, long long nframes, float* out, long long out_batch_stride,
                         float* sum_out, float post_scale, float* scratch, void* stream);

/*
 * Display scaling of a (cropped) spectrogram on the device -- PlotEngine._plot_spectrogram,
 * PlotEngine.py:126-131: out = clip(S/(base+1e-20), 0, 1) with base = max(S) (or global_max
 * when > 0); with log_scale, 10*log10(out+1e-12) min-max normalised to [0,1] (zeros when the
 * dB range is <= 1e-6).  `scratch` holds 2 unsigned ints (device).  S and out: `elems` floats.
 */
int b2s_display_scale_f32(const float* s, long long elems, int log_scale, float global_max, float* out,
                          unsigned int* scratch, void* stream);

/* Power summaries of the spectrogram the path left on the device -- the reductions of
 * PlotEngine.calculate_absolute_power (PlotEngine.py:686-690: sum(last_Sxx)) and
 * PlotEngine.calculate_band_powers (PlotEngine.py:692-719: sum over the rows of last_f in
 * [low, high) of max(0, last_Sxx), divided by the total on the host).  s: device [frames][bins]
 * (the cropped result of b2s_stft_psd_*); k0/k1: HOST arrays of nb <= 16 half-open bin ranges
 * [k0, k1) (empty ranges allowed); out: device, nb + 1 doubles -- the band sums, then the total
 * over all bins; scratch: device, b2s_band_sums_scratch_elems() doubles.  Fixed summation order,
 * double accumulators; negative values are clamped to 0 like the reference's np.maximum. */
long long b2s_band_sums_scratch_elems(void);
int b2s_band_sums_f32(const float* s, long long frames, int bins, const int* k0, const int* k1, int nb, double* out,
                      double* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2S_H_ */

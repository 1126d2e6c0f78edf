/* dspb200.h -- C ABI of libdspb200.so: the B200-native (sm_100a) implementation of DSP.jl's
 * data-parallel hot path.  Plain pointers and sizes only; callable from Julia `ccall`, Python `ctypes`, C.
 *
 * Each entry point names the reference interface it replaces (paths relative to the DSP.jl tree,
 * v0.8.5 @ 2d57c27).  The reference has no FFI of its own: its only native boundary is FFTW.jl's plan
 * API called from src/dspbase.jl, src/Filters/filt.jl, src/Filters/stream_filt.jl, src/periodograms.jl.
 * This library replaces those FFTW calls *and* the Julia inner loops around them.
 *
 * Conventions
 *  - every function returns a status (DSPB200_OK == 0, negative on error) and never throws/aborts;
 *    dspb200_last_error() returns a thread-local message for the last failure.
 *  - dtype: element type of the signal (DSPB200_F32/F64/C32/C64); complex = interleaved (re, im),
 *    the memory layout of Julia's Complex{T}.  Arrays are column-major: dim 1 is time, every column
 *    is an independent channel (src/dspbase.jl:55, src/Filters/filt.jl:504).
 *  - `*_exec`     : HOST pointers (caller-owned, never retained); copies in, computes on the GPU, copies
 *                   out; synchronous on return.  Pinned host memory (dspb200_host_alloc) is streamed in
 *                   chunks so copies overlap compute.
 *  - `*_exec_dev` : DEVICE pointers; enqueued on `stream` (a cudaStream_t, NULL = default stream);
 *                   asynchronous.
 *  - plans own device scratch, twiddle tables and cuFFT plans; one caller at a time per plan (the
 *    reference's WelchConfig / FIRFilter / ArraySplit scratch is equally non-reentrant:
 *    src/periodograms.jl:88-90,525-526; src/Filters/stream_filt.jl:137-142).
 *  - there is no CPU fallback: without a CUDA device every exec call fails with DSPB200_ECUDA.
 */
#ifndef DSPB200_H
#define DSPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSPB200_VERSION 100

#if defined(__GNUC__)
#define DSPB200_API __attribute__((visibility("default")))
#else
#define DSPB200_API
#endif

enum { DSPB200_F32 = 0, DSPB200_F64 = 1, DSPB200_C32 = 2, DSPB200_C64 = 3 };

enum {
    DSPB200_OK = 0,
    DSPB200_EINVALID = -1,      /* argument check failed (the Julia glue raises the reference's exception types first) */
    DSPB200_ECUDA = -2,         /* CUDA runtime error / no device */
    DSPB200_ECUFFT = -3,        /* cuFFT error (generic-size path) */
    DSPB200_ENOMEM = -4,        /* device or pinned-host allocation failed */
    DSPB200_EUNSUPPORTED = -5   /* combination outside the hot-path scope */
};

/* ------------------------------------------------------------------------------------------ runtime */
DSPB200_API int dspb200_version(void);
DSPB200_API const char* dspb200_last_error(void);
DSPB200_API int dspb200_device_count(int* count);
DSPB200_API int dspb200_set_device(int device);                 /* device used by plans created afterwards on this thread */
DSPB200_API int dspb200_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem, size_t* l2_bytes);
DSPB200_API int dspb200_malloc(void** dptr, size_t bytes);      /* device memory, for hosts without a CUDA binding */
DSPB200_API int dspb200_free(void* dptr);
DSPB200_API int dspb200_host_alloc(void** hptr, size_t bytes);  /* pinned host memory */
DSPB200_API int dspb200_host_free(void* hptr);
DSPB200_API int dspb200_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
DSPB200_API int dspb200_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
DSPB200_API int dspb200_stream_sync(void* stream);

/* Number of kernels this library has launched in this process (bench.py's `gpu_launches`). */
DSPB200_API int64_t dspb200_launch_count(void);

/* ------------------------------------------------------------------------------------------ FIR, time domain
 * filt(b, 1, x) / filt!(out, b, 1, x) / tdfilt(h, x): src/dspbase.jl:14-15, 26-66, 95-154;
 * src/Filters/filt.jl:431-443.  y[i] = sum_k b[k] x[i-k+1] per column, evaluated with the reference's
 * accumulation order (oldest tap first, one fused multiply-add per tap).  b, x and y share `dtype`
 * (the host promotes, src/dspbase.jl:15); b is normalised by a[1] on the host (src/dspbase.jl:43-47). */
typedef struct dspb200_fir_plan dspb200_fir_plan;
DSPB200_API int dspb200_fir_plan_create(dspb200_fir_plan** plan, int dtype, const void* b_host, int64_t nb);
DSPB200_API int dspb200_fir_exec(dspb200_fir_plan* plan, const void* x, int64_t nx, int64_t ncols, void* out);
DSPB200_API int dspb200_fir_exec_dev(dspb200_fir_plan* plan, const void* x, int64_t nx, int64_t ncols, void* out, void* stream);
DSPB200_API int dspb200_fir_plan_destroy(dspb200_fir_plan* plan);

/* ------------------------------------------------------------------------------------------ overlap-save
 * conv(u, v; algorithm=:fft_overlapsave) / unsafe_conv_kern_os! (src/dspbase.jl:490-609, 299-356) and
 * fftfilt / _fftfilt! / filt(b, x) (src/Filters/filt.jl:458-555).
 * out[m] = sum_j u[j] v[m-j], m = 0 .. nout-1, per column; nout = nu for fftfilt/filt (src/Filters/filt.jl:517),
 * nout = nu+nv-1 for conv; samples of `out` beyond nu+nv-1 are zero-filled (src/dspbase.jl:733-735).
 * u, v and out share `dtype` (real: two blocks ride one complex FFT; complex: one block per FFT).
 * nfft: 0 = library choice (largest shared-memory transform that amortises the nv-1 halo; the reference's
 *       optimalfftfiltlength, src/dspbase.jl:268-291, is a CPU cost model -- any nfft >= nv gives the same
 *       convolution); otherwise a power of two in [32, 16384] (F64/C64: 8192) runs the fused kernel and
 *       any other value >= nv runs gather -> cuFFT -> multiply -> cuFFT -> scatter. */
typedef struct dspb200_os_plan dspb200_os_plan;
DSPB200_API int dspb200_os_plan_create(dspb200_os_plan** plan, int dtype, const void* v_host, int64_t nv, int64_t nfft);
DSPB200_API int dspb200_os_plan_nfft(const dspb200_os_plan* plan, int64_t* nfft, int* fused);
DSPB200_API int dspb200_os_plan_geometry(const dspb200_os_plan* plan, int* dtype, int64_t* nv, int64_t* nfft);
DSPB200_API int dspb200_os_exec(dspb200_os_plan* plan, const void* u, int64_t nu, int64_t ncols, void* out, int64_t nout);
DSPB200_API int dspb200_os_exec_dev(dspb200_os_plan* plan, const void* u, int64_t nu, int64_t ncols, void* out, int64_t nout,
                        void* stream);
/* Range form for sharding one long column across GPUs: compute outputs
 * [out_begin, out_begin+out_count) of the convolution of the virtual signal whose samples
 * [u_begin, u_begin+nu_local) are stored at `u_local` (everything outside is zero).  No collective. */
DSPB200_API int dspb200_os_exec_range_dev(dspb200_os_plan* plan, const void* u_local, int64_t u_begin, int64_t nu_local,
                              void* out_local, int64_t out_begin, int64_t out_count, void* stream);
DSPB200_API int dspb200_os_plan_destroy(dspb200_os_plan* plan);

/* conv(u, v; algorithm=:fft_simple) / _conv_kern_fft!: src/dspbase.jl:611-644 -- one FFT pair of size
 * nfft >= nu+nv-1 (the host passes nextfastfft(nu+nv-1), src/util.jl:134); out has nu+nv-1 samples. */
DSPB200_API int dspb200_conv_fft_exec(int dtype, const void* u, int64_t nu, const void* v, int64_t nv, int64_t nfft, void* out);
/* conv(u, v; algorithm=:direct) / _conv_td!: src/dspbase.jl:646-660 -- direct muladd convolution. */
DSPB200_API int dspb200_conv_direct_exec(int dtype, const void* u, int64_t nu, const void* v, int64_t nv, void* out);

/* conv(u, v; algorithm) for matrices and rank-3 arrays: src/dspbase.jl:611-660 (_conv_kern_fft!, _conv_td!), 709-757.
 * Column-major arrays of `rank` <= 3 dimensions, sizes usize / vsize; out has usize + vsize - 1 per dimension.
 * nffts != NULL: one N-D FFT pair of size nffts (the host passes nextfastfft.(usize .+ vsize .- 1), :618, :632);
 * nffts == NULL: direct muladd convolution (:646-660).  The _dev forms take device pointers and a cudaStream_t and return
 * after the work on that stream has completed (plans and scratch come from the library's cache). */
DSPB200_API int dspb200_conv_nd_exec(int dtype, int rank, const int64_t* usize, const void* u, const int64_t* vsize, const void* v,
                                     const int64_t* nffts, void* out);
DSPB200_API int dspb200_conv_nd_exec_dev(int dtype, int rank, const int64_t* usize, const void* d_u, const int64_t* vsize,
                                         const void* d_v, const int64_t* nffts, void* d_out, void* stream);
/* conv(u, v; algorithm=:fft_overlapsave) for arrays of rank <= 3: unsafe_conv_kern_os! with its perimeter blocks
 * unsafe_conv_kern_os_edge!, src/dspbase.jl:371-609.  u is the array with more elements (:746-751); nffts[d] >= vsize[d] is
 * the block transform per dimension (the host passes optimalfftfiltlength.(vsize, usize), :736); every block contributes
 * save_blocksize = nffts - vsize + 1 outputs per dimension (:500-506).  Blocks are gathered (zero outside u), transformed
 * by ONE batched N-D cuFFT plan, multiplied by the filter spectrum and scattered, as many per batch as fit the block-buffer
 * budget (default 1 GiB; dspb200_conv_nd_os_set_budget) -- so arrays whose single transform would not fit are convolved
 * in bounded memory, which is what the reference's blocking is for. */
DSPB200_API int dspb200_conv_nd_os_exec(int dtype, int rank, const int64_t* usize, const void* u, const int64_t* vsize, const void* v,
                                        const int64_t* nffts, void* out);
DSPB200_API int dspb200_conv_nd_os_exec_dev(int dtype, int rank, const int64_t* usize, const void* d_u, const int64_t* vsize,
                                            const void* d_v, const int64_t* nffts, void* d_out, void* stream);
DSPB200_API int dspb200_conv_nd_os_set_budget(size_t bytes);
/* hilbert(x): src/util.jl:31-75 -- analytic signal of a real [n x ncols] column-major array along dim 1 (rfft, bins
 * 2 .. n/2+isodd(n) doubled, the rest of the negative half zero, normalised inverse FFT).  dtype F32 -> ComplexF32 out,
 * F64 -> ComplexF64 (integers are converted by the host, src/util.jl:43).  Any n (cuFFT).  The _dev form takes device
 * pointers and a cudaStream_t and returns after the work on that stream has completed. */
DSPB200_API int dspb200_hilbert_exec(int dtype, const void* x, int64_t n, int64_t ncols, void* out);
DSPB200_API int dspb200_hilbert_exec_dev(int dtype, const void* d_x, int64_t n, int64_t ncols, void* d_out, void* stream);

/* ------------------------------------------------------------------------------------------ Welch / STFT
 * One plan per (dtype, n, noverlap, nfft, onesided, window): the analogue of WelchConfig
 * (src/periodograms.jl:516-576) = ArraySplit segmenter (:32-73) + forward_plan (:511-514) + fft2pow! (:142-172)
 * + fft2oneortwosided! (:234-244).  `window`: n Float64 values or NULL for `nothing` (:248-257; the sample *
 * window product is formed in Float64 and rounded to the buffer eltype, :66).  Power-of-two nfft in
 * [32, 16384] (F64/C64: 8192) runs fused shared-memory kernels; other sizes use cuFFT. */
typedef struct dspb200_spec_plan dspb200_spec_plan;
DSPB200_API int dspb200_spec_plan_create(dspb200_spec_plan** plan, int dtype, int64_t n, int64_t noverlap, int64_t nfft,
                             int onesided, const double* window_host);
DSPB200_API int dspb200_spec_plan_info(const dspb200_spec_plan* plan, int64_t* nout, int* fused);
DSPB200_API int dspb200_spec_plan_geometry(const dspb200_spec_plan* plan, int* dtype, int64_t* n, int64_t* hop, int64_t* nout);
DSPB200_API int64_t dspb200_spec_nsegments(const dspb200_spec_plan* plan, int64_t len);   /* k, src/periodograms.jl:49-50 */

/* welch_pgram / welch_pgram! / welch_pgram_helper!: src/periodograms.jl:647-759.
 * out[nout] (real eltype of dtype) = sum over segments of fft2pow!(.., r, onesided) with r = k*fs*norm2 (:751).
 * periodogram (:393-417) is the k = 1 case (n = length(s)). */
DSPB200_API int dspb200_welch_exec(dspb200_spec_plan* plan, const void* s, int64_t len, double r, void* out);
DSPB200_API int dspb200_welch_exec_dev(dspb200_spec_plan* plan, const void* s, int64_t len, double r, void* out, void* stream);
/* Segment-range form for multi-GPU sharding: accumulates only segments [seg_begin, seg_end) of the signal
 * whose sample `sample_offset` is s[0]; the caller sums the partial spectra (NCCL all-reduce). */
DSPB200_API int dspb200_welch_exec_range_dev(dspb200_spec_plan* plan, const void* s, int64_t len, int64_t sample_offset,
                                 int64_t seg_begin, int64_t seg_end, double r, void* out, void* stream);

/* Streaming form of welch_pgram_helper! (src/periodograms.jl:746-759): begin zeroes the accumulator, accumulate adds the
 * segments [seg_begin, seg_end) found in a buffer whose first sample is `sample_offset` (any number of calls, any chunking),
 * finalize applies the fft2pow! scaling with r = k*fs*norm2 and writes out[nout]. */
DSPB200_API int dspb200_welch_begin_dev(dspb200_spec_plan* plan, void* stream);
DSPB200_API int dspb200_welch_accumulate_dev(dspb200_spec_plan* plan, const void* s, int64_t len, int64_t sample_offset,
                                 int64_t seg_begin, int64_t seg_end, void* stream);
DSPB200_API int dspb200_welch_finalize_dev(dspb200_spec_plan* plan, double r, void* out, void* stream);
/* welch_pgram(filt(b, x), config) as ONE host-pointer call (src/dspbase.jl:14-15 + src/periodograms.jl:702-759): x[n] on the
 * host (pinned memory lets the copies overlap), out[nout] on the host.  The stream goes through the GPU in chunks: the H2D copy
 * of chunk c+1 overlaps the overlap-save convolution of chunk c (os plan: taps b, dtype of x) and the Welch accumulation of
 * the segments chunk c completes; the filter output y = (b * x)[0:n] stays in HBM.  r = k*fs*norm2 with k = segments of n. */
DSPB200_API int dspb200_filt_welch_exec(dspb200_os_plan* os, dspb200_spec_plan* spec, const void* x_host, int64_t n, double r,
                            void* out_host);

/* stft / spectrogram: src/periodograms.jl:828-897.  `s` holds nchan columns of `len` samples; out holds
 * nchan matrices of nout x k (column-major, column = segment).  psd_only != 0: PSD columns (real eltype)
 * scaled with r = fs*norm2 (:883,890); psd_only == 0: raw spectra (complex eltype), two-sided real input
 * completed by conjugate symmetry (:234-244).  nchan > 1 is the batched form of the per-vector reference
 * signature (SURVEY.md hard part 4).  psd_only == 3 (device form, fused sizes): the PSD columns are ADDED to the contents
 * of `out` -- how mt_spectrogram sums its tapers (src/multitaper.jl:362-377) without a separate pass. */
DSPB200_API int dspb200_stft_exec(dspb200_spec_plan* plan, const void* s, int64_t len, int64_t nchan, double r, int psd_only,
                      void* out);
DSPB200_API int dspb200_stft_exec_dev(dspb200_spec_plan* plan, const void* s, int64_t len, int64_t nchan, double r, int psd_only,
                          void* out, void* stream);
/* arraysplit(s, n, noverlap, nfft, window) / ArraySplit: src/periodograms.jl:32-73, 134-137.  out = k x nfft matrix, row i =
 * [window .* s[i*hop .. i*hop+n) ; zeros(nfft-n)] (the reference yields the rows one at a time into one reused buffer). */
DSPB200_API int dspb200_arraysplit_exec(dspb200_spec_plan* plan, const void* s, int64_t len, void* out);
/* periodogram(s::AbstractMatrix; nfft, fs, radialsum, radialavg): src/periodograms.jl:473-509 with fft2pow2! (:175-181) and
 * fft2pow2radial! (:183-232).  s is an n1 x n2 real matrix (column-major), zero-padded to nfft1 x nfft2; r = fs * length(s).
 * ptype 0: out = real[nfft1 x nfft2] two-dimensional PSD; 1 (radialsum) / 2 (radialavg): out = real[min(nfft)>>1 + 1]. */
DSPB200_API int dspb200_periodogram2_exec(int dtype, const void* s, int64_t n1, int64_t n2, int64_t nfft1, int64_t nfft2, double r,
                                          int ptype, void* out);
/* device pointers + cudaStream_t; returns after the work on that stream has completed */
DSPB200_API int dspb200_periodogram2_exec_dev(int dtype, const void* d_s, int64_t n1, int64_t n2, int64_t nfft1, int64_t nfft2,
                                              double r, int ptype, void* d_out, void* stream);

/* Multitaper ("next" rank 1): mt_pgram / mt_spectrogram, src/multitaper.jl:117-242, 262-404.
 * `tapers` = ntapers rows of n Float64 samples, each pre-scaled by the host with 1/sqrt(r_t),
 * r_t = fs * sum|w_t|^2 / weight_t (:135-139); the library then sums fft2pow!(FFT(w_t .* segment), 1) over tapers.
 * mt_pgram: len must equal n (DimensionMismatch :226); out = nout values.  mt_spectrogram: out = nout x k. */
DSPB200_API int dspb200_mt_plan_create(dspb200_spec_plan** plan, int dtype, int64_t n, int64_t noverlap, int64_t nfft,
                           int onesided, const double* tapers_host, int64_t ntapers);
DSPB200_API int dspb200_mt_pgram_exec(dspb200_spec_plan* plan, const void* s, int64_t len, void* out);
DSPB200_API int dspb200_mt_spectrogram_exec(dspb200_spec_plan* plan, const void* s, int64_t len, void* out);
/* device pointers + cudaStream_t; return after the work on that stream has completed */
DSPB200_API int dspb200_mt_pgram_exec_dev(dspb200_spec_plan* plan, const void* d_s, int64_t len, void* d_out, void* stream);
DSPB200_API int dspb200_mt_spectrogram_exec_dev(dspb200_spec_plan* plan, const void* d_s, int64_t len, void* d_out, void* stream);
/* mt_cross_power_spectra! / mt_coherence!: src/multitaper.jl:553-616, 672-693, 722-790.  `signal` is the reference's
 * n_channels x n_samples matrix (column-major: channel index fastest), n_samples = the plan's n; the plan must be real and
 * one-sided (:411-416) with noverlap = 0.  demean != 0 subtracts the channel means (:566-570).  [f_lo, f_lo+nf) is the
 * 0-based range of retained frequency bins (freq_range, :497-503).  coherence == 0: out = Complex[n_channels x n_channels
 * x nf] cross power spectra; != 0: out = real[n_channels x n_channels x nf] pairwise coherences. */
DSPB200_API int dspb200_mt_cross_spectra_exec(dspb200_spec_plan* plan, const void* signal, int64_t nchan, int demean,
                                              int64_t f_lo, int64_t nf, int coherence, void* out);
DSPB200_API int dspb200_mt_cross_spectra_exec_dev(dspb200_spec_plan* plan, const void* d_signal, int64_t nchan, int demean,
                                                  int64_t f_lo, int64_t nf, int coherence, void* d_out, void* stream);
DSPB200_API int dspb200_spec_plan_destroy(dspb200_spec_plan* plan);

/* ------------------------------------------------------------------------------------------ polyphase resample
 * resample(x, rate, h) with Integer / Rational rate = interp // decim: FIRFilter{FIRRational | FIRInterpolator |
 * FIRDecimator} + filt! loops (src/Filters/stream_filt.jl:8-78, 137-178, 294-307, 431-560) and _resample!
 * (:696-725).  y[j] = sum_t hp[phi + t*interp] x[n - t], p = phi0 + j*decim, n = n0 + p / interp,
 * phi = p % interp (closed form of the (inputIdx, phiIdx) recurrence); x is zero outside [0, nx).
 * (n0, phi0) come from setphase!(timedelay) on the host (:223-229, 400-403, 706-714).
 * dtype_x in {F32,F64,C32,C64}, dtype_h in {F32,F64}; output eltype = promote_type(dtype_h, dtype_x) (:654). */
typedef struct dspb200_resample_plan dspb200_resample_plan;
DSPB200_API int dspb200_resample_plan_create(dspb200_resample_plan** plan, int dtype_x, int dtype_h, const void* h_host,
                                 int64_t hlen, int64_t interp, int64_t decim);
DSPB200_API int dspb200_resample_out_dtype(const dspb200_resample_plan* plan, int* dtype_out);
DSPB200_API int dspb200_resample_exec(dspb200_resample_plan* plan, const void* x, int64_t nx, int64_t ncols, int64_t n0,
                          int64_t phi0, void* out, int64_t nout);
DSPB200_API int dspb200_resample_exec_dev(dspb200_resample_plan* plan, const void* x, int64_t nx, int64_t ncols, int64_t n0,
                              int64_t phi0, void* out, int64_t nout, void* stream);
/* Range form: outputs [j_begin, j_begin+nout_local) of the virtual input whose samples
 * [x_begin, x_begin+nx_local) are stored at x_local (zero elsewhere). */
DSPB200_API int dspb200_resample_exec_range_dev(dspb200_resample_plan* plan, const void* x_local, int64_t x_begin,
                                    int64_t nx_local, int64_t n0, int64_t phi0, void* out_local, int64_t j_begin,
                                    int64_t nout_local, void* stream);
/* Arbitrary (floating-point) rate: FIRArbitrary / filt!(buffer, ::FIRFilter{FIRArbitrary}, x),
 * src/Filters/stream_filt.jl:92-134, 567-625.  The plan holds pfb = taps2pfb(h, nphases) and dpfb = taps2pfb([diff(h); 0],
 * nphases).  Output j (0-based) of a call sits at total phase P_j = acc0 + j*delta (delta = nphases / rate, acc0 = the
 * kernel's phiAccumulator in [0, nphases)): newest input sample n0 + floor(P_j / nphases) (0-based index into x, which the
 * host passes as [history; x]), phase floor(P_j mod nphases), alpha = its fraction; y_j = muladd(yUpper, alpha, yLower)
 * (:606-616).  Samples outside [0, nx) are zero.  Out eltype = promote(eltype(h), eltype(x)) as for the rational plan. */
DSPB200_API int dspb200_resample_arb_plan_create(dspb200_resample_plan** plan, int dtype_x, int dtype_h, const void* h_host,
                                                 int64_t hlen, int64_t nphases);
DSPB200_API int dspb200_resample_arb_exec(dspb200_resample_plan* plan, const void* x, int64_t nx, int64_t n0, double acc0,
                                          double delta, void* out, int64_t nout);
DSPB200_API int dspb200_resample_arb_exec_dev(dspb200_resample_plan* plan, const void* x, int64_t nx, int64_t n0, double acc0,
                                              double delta, void* out, int64_t nout, void* stream);
DSPB200_API int dspb200_resample_plan_destroy(dspb200_resample_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* DSPB200_H */

// dspb200 -- host-pointer pipelines over the device-pointer entry points of the other modules.
//
// dspb200_filt_welch_exec: welch_pgram(filt(b, x), config) (src/dspbase.jl:14-15, src/periodograms.jl:702-759) for a stream
// that lives in host memory.  The filter output never crosses PCIe: per chunk c the copy stream uploads the chunk's samples
// (plus the nv-1 sample halo) while the execute stream runs, for chunk c-1, the overlap-save convolution of its output range
// (dspb200_os_exec_range_dev) and the Welch accumulation of the segments that range completes
// (dspb200_welch_accumulate_dev).  End to end the step costs the H2D copy plus one chunk's kernels.
#include "common.cuh"
#include <mutex>

namespace dspb200 {

struct FiltWelchState {
    int device = -1;
    cudaStream_t s_copy = nullptr, s_exec = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    DevBuf in[2], y, out;
};
// one scratch set per device (the C ABI is used from one host thread per device; guarded for safety)
static FiltWelchState g_fw[16];
static std::mutex g_fw_mutex;

static int fw_state(FiltWelchState** st) {
    int dev = 0;
    DSP_CUDA(cudaGetDevice(&dev));
    DSP_REQUIRE(dev >= 0 && dev < 16, "device index %d out of range", dev);
    FiltWelchState* s = &g_fw[dev];
    if (s->device < 0) {
        DSP_CUDA(cudaStreamCreateWithFlags(&s->s_copy, cudaStreamNonBlocking));
        DSP_CUDA(cudaStreamCreateWithFlags(&s->s_exec, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            DSP_CUDA(cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming));
            DSP_CUDA(cudaEventCreateWithFlags(&s->ev_free[i], cudaEventDisableTiming));
        }
        s->device = dev;
    }
    *st = s;
    return DSPB200_OK;
}

}  // namespace dspb200

using namespace dspb200;

extern "C" {

int dspb200_filt_welch_exec(dspb200_os_plan* os, dspb200_spec_plan* spec, const void* x_host, int64_t n, double r,
                            void* out_host) {
    DSP_RANGE("dspb200_filt_welch_exec");
    DSP_REQUIRE(os && spec && out_host, "NULL argument");
    DSP_REQUIRE(n >= 0, "n must be >= 0");
    DSP_REQUIRE(r != 0.0, "r must be nonzero");
    int64_t nv = 0, nfft = 0, nseg_len = 0, hop = 0, nout = 0;
    int dt_os = 0, dt_spec = 0;
    DSP_TRY(dspb200_os_plan_geometry(os, &dt_os, &nv, &nfft));
    DSP_TRY(dspb200_spec_plan_geometry(spec, &dt_spec, &nseg_len, &hop, &nout));
    DSP_REQUIRE(dt_os == dt_spec, "filter plan dtype (%d) and Welch plan dtype (%d) differ", dt_os, dt_spec);
    std::lock_guard<std::mutex> lock(g_fw_mutex);
    FiltWelchState* st = nullptr;
    DSP_TRY(fw_state(&st));
    const size_t esz = dtype_size(dt_os);
    const size_t out_bytes = (size_t)nout * (dtype_is_f64(dt_os) ? 8 : 4);
    const int64_t halo = nv - 1;
    const int64_t k = n >= nseg_len ? (n - nseg_len) / hop + 1 : 0;
    DSP_TRY(st->out.reserve(out_bytes));
    DSP_TRY(dspb200_welch_begin_dev(spec, st->s_exec));
    if (k > 0) {
        DSP_REQUIRE(x_host != nullptr, "x is NULL");
        // chunk = a whole number of overlap-save blocks (L outputs each), ~32 MiB of new samples
        const int64_t L = nfft - nv + 1;
        int64_t chunk = ((int64_t(32) << 20) / (int64_t)esz) / L * L;
        if (chunk < L) chunk = L;
        if (chunk > n) chunk = n;
        DSP_TRY(st->y.reserve((size_t)n * esz));
        DSP_TRY(st->in[0].reserve((size_t)(chunk + halo) * esz));
        if (chunk < n) DSP_TRY(st->in[1].reserve((size_t)(chunk + halo) * esz));
        int64_t seg_done = 0;
        bool used[2] = {false, false};
        int slot = 0;
        for (int64_t c0 = 0; c0 < n; c0 += chunk, slot ^= 1) {
            const int64_t c1 = c0 + chunk < n ? c0 + chunk : n;
            const int64_t in0 = c0 - halo > 0 ? c0 - halo : 0;
            if (used[slot]) DSP_CUDA(cudaStreamWaitEvent(st->s_copy, st->ev_free[slot], 0));
            DSP_CUDA(cudaMemcpyAsync(st->in[slot].p, (const char*)x_host + (size_t)in0 * esz, (size_t)(c1 - in0) * esz,
                                     cudaMemcpyHostToDevice, st->s_copy));
            DSP_CUDA(cudaEventRecord(st->ev_in[slot], st->s_copy));
            DSP_CUDA(cudaStreamWaitEvent(st->s_exec, st->ev_in[slot], 0));
            // y[c0, c1) = (b * x)[c0, c1): same-length filter output, the rank-local buffer holds x[in0, c1)
            DSP_TRY(dspb200_os_exec_range_dev(os, st->in[slot].p, in0, c1 - in0, (char*)st->y.p + (size_t)c0 * esz, c0, c1 - c0, st->s_exec));
            DSP_CUDA(cudaEventRecord(st->ev_free[slot], st->s_exec));
            used[slot] = true;
            // segments that end inside [0, c1) and have not been transformed yet
            int64_t seg_hi = c1 >= nseg_len ? (c1 - nseg_len) / hop + 1 : 0;
            if (seg_hi > k) seg_hi = k;
            if (seg_hi > seg_done) {
                DSP_TRY(dspb200_welch_accumulate_dev(spec, st->y.p, c1, 0, seg_done, seg_hi, st->s_exec));
                seg_done = seg_hi;
            }
        }
    }
    DSP_TRY(dspb200_welch_finalize_dev(spec, r, st->out.p, st->s_exec));
    DSP_CUDA(cudaMemcpyAsync(out_host, st->out.p, out_bytes, cudaMemcpyDeviceToHost, st->s_exec));
    DSP_CUDA(cudaStreamSynchronize(st->s_exec));
    return DSPB200_OK;
}

}  // extern "C"

// dspb200 -- runtime entry points: errors, device selection, memory helpers.
#include "common.cuh"
#include <atomic>

namespace dspb200 {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? DSPB200_ENOMEM : DSPB200_ECUDA;
}

int device_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace dspb200

using namespace dspb200;

extern "C" {

int dspb200_version(void) { return DSPB200_VERSION; }
const char* dspb200_last_error(void) { return g_err; }
int64_t dspb200_launch_count(void) { return g_launches.load(); }

int dspb200_device_count(int* count) {
    DSP_REQUIRE(count != nullptr, "count is NULL");
    *count = 0;
    DSP_CUDA(cudaGetDeviceCount(count));
    return DSPB200_OK;
}

int dspb200_set_device(int device) {
    DSP_CUDA(cudaSetDevice(device));
    return DSPB200_OK;
}

int dspb200_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem, size_t* l2_bytes) {
    int dev = 0;
    DSP_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    DSP_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    if (l2_bytes) *l2_bytes = (size_t)p.l2CacheSize;
    return DSPB200_OK;
}

int dspb200_malloc(void** dptr, size_t bytes) {
    DSP_REQUIRE(dptr != nullptr, "dptr is NULL");
    *dptr = nullptr;
    if (bytes == 0) return DSPB200_OK;
    DSP_CUDA(cudaMalloc(dptr, bytes));
    return DSPB200_OK;
}
int dspb200_free(void* dptr) {
    if (dptr) DSP_CUDA(cudaFree(dptr));
    return DSPB200_OK;
}
int dspb200_host_alloc(void** hptr, size_t bytes) {
    DSP_REQUIRE(hptr != nullptr, "hptr is NULL");
    *hptr = nullptr;
    if (bytes == 0) return DSPB200_OK;
    DSP_CUDA(cudaHostAlloc(hptr, bytes, cudaHostAllocDefault));
    return DSPB200_OK;
}
int dspb200_host_free(void* hptr) {
    if (hptr) DSP_CUDA(cudaFreeHost(hptr));
    return DSPB200_OK;
}
int dspb200_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes == 0) return DSPB200_OK;
    DSP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return DSPB200_OK;
}
int dspb200_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes == 0) return DSPB200_OK;
    DSP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return DSPB200_OK;
}
int dspb200_stream_sync(void* stream) {
    DSP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return DSPB200_OK;
}

}  // extern "C"

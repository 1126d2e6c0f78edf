// dspb200 -- common device/host helpers (complex type, dtype traits, error plumbing).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <type_traits>

#include "../../include/dspb200.h"

namespace dspb200 {

// ----------------------------------------------------------------------------------------------
// Interleaved complex (layout-identical to Julia's Complex{T}, float2 / double2).
template <typename T>
struct alignas(2 * sizeof(T)) cx {
    T x, y;
};

template <typename T> __host__ __device__ __forceinline__ cx<T> mkc(T a, T b) { cx<T> r; r.x = a; r.y = b; return r; }
template <typename T> __host__ __device__ __forceinline__ cx<T> operator+(cx<T> a, cx<T> b) { return mkc<T>(a.x + b.x, a.y + b.y); }
template <typename T> __host__ __device__ __forceinline__ cx<T> operator-(cx<T> a, cx<T> b) { return mkc<T>(a.x - b.x, a.y - b.y); }
template <typename T> __host__ __device__ __forceinline__ cx<T> cmul(cx<T> a, cx<T> b) {
    return mkc<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T> __host__ __device__ __forceinline__ cx<T> cscale(cx<T> a, T s) { return mkc<T>(a.x * s, a.y * s); }
template <typename T> __host__ __device__ __forceinline__ cx<T> cconj(cx<T> a) { return mkc<T>(a.x, -a.y); }
template <typename T> __host__ __device__ __forceinline__ cx<T> cswap(cx<T> a) { return mkc<T>(a.y, a.x); }
// multiply by -i
template <typename T> __host__ __device__ __forceinline__ cx<T> mul_mi(cx<T> a) { return mkc<T>(a.y, -a.x); }
template <typename T> __host__ __device__ __forceinline__ T cabs2(cx<T> a) { return a.x * a.x + a.y * a.y; }

// ----------------------------------------------------------------------------------------------
// dtype traits
template <typename E> struct elt_traits;
template <> struct elt_traits<float>       { using real = float;  static constexpr bool is_cplx = false; };
template <> struct elt_traits<double>      { using real = double; static constexpr bool is_cplx = false; };
template <> struct elt_traits<cx<float>>   { using real = float;  static constexpr bool is_cplx = true; };
template <> struct elt_traits<cx<double>>  { using real = double; static constexpr bool is_cplx = true; };

inline size_t dtype_size(int dt) {
    switch (dt) {
        case DSPB200_F32: return 4;
        case DSPB200_F64: return 8;
        case DSPB200_C32: return 8;
        case DSPB200_C64: return 16;
    }
    return 0;
}
inline bool dtype_is_cplx(int dt) { return dt == DSPB200_C32 || dt == DSPB200_C64; }
inline bool dtype_is_f64(int dt) { return dt == DSPB200_F64 || dt == DSPB200_C64; }
inline bool dtype_valid(int dt) { return dt >= 0 && dt <= 3; }

// ----------------------------------------------------------------------------------------------
// error plumbing (thread-local message; the C ABI never throws)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define DSP_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) return ::dspb200::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define DSP_REQUIRE(cond, ...)                    \
    do {                                          \
        if (!(cond)) {                            \
            ::dspb200::set_error(__VA_ARGS__);    \
            return DSPB200_EINVALID;              \
        }                                         \
    } while (0)

#define DSP_TRY(expr)               \
    do {                            \
        int rc__ = (expr);          \
        if (rc__ != DSPB200_OK) return rc__; \
    } while (0)

// Device scratch buffer that only grows (owned by plans).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return DSPB200_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return DSPB200_ENOMEM; }
        cap = bytes;
        return DSPB200_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

int device_sm_count();
void count_launch(int n = 1);

// after every kernel launch
#define DSP_LAUNCH_OK()                                                              \
    do {                                                                            \
        ::dspb200::count_launch(1);                                                 \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) return ::dspb200::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace dspb200

// dspb200 -- time-domain FIR: filt(b, 1, x) / tdfilt (src/dspbase.jl:26-66, 95-154; src/Filters/filt.jl:431-443).
//
// The reference runs a transposed direct-form loop whose state update is si[j] = muladd(x_i, b[j+1], si[j+1])
// (:95-105, and the unrolled NTuple form :118-141), which unrolls to
//     y[i] = fma(x[i], b[1], fma(x[i-1], b[2], ... fma(x[i-nb+2], b[nb-1], x[i-nb+1]*b[nb])))
// i.e. one fused multiply-add per tap, oldest tap first.  This kernel evaluates exactly that chain per
// output, so Float32/Float64 results match the reference bit for bit on FMA hardware.
// Layout: CTA = 256 threads x 4 consecutive outputs; the x tile (+ tap-chunk halo) and the tap chunk are
// staged in shared memory (padded so the stride-4 sliding-window reads are bank-conflict free).
#include "common.cuh"
#include <new>

namespace dspb200 {

template <typename T, bool CPLX> struct fir_elt { using type = T; };
template <typename T> struct fir_elt<T, true> { using type = cx<T>; };

constexpr int FIR_NT = 256;
constexpr int FIR_OPT = 4;                      // outputs per thread
constexpr int FIR_TILE = FIR_NT * FIR_OPT;      // outputs per CTA
constexpr int FIR_KC = 512;                     // taps per chunk

__host__ __device__ __forceinline__ int fir_pad(int j) { return j + (j >> 5); }

template <typename T> __device__ __forceinline__ T fir_fma(T x, T b, T acc) { return fma(x, b, acc); }
// Base.muladd(z::Complex, w::Complex, x::Complex) (base/complex.jl)
template <typename T> __device__ __forceinline__ cx<T> fir_fma(cx<T> z, cx<T> w, cx<T> x) {
    return mkc<T>(fma(z.x, w.x, -fma(z.y, w.y, -x.x)), fma(z.x, w.y, fma(z.y, w.x, x.y)));
}
template <typename T> __device__ __forceinline__ T fir_zero(T*) { return T(0); }
template <typename T> __device__ __forceinline__ cx<T> fir_zero(cx<T>*) { return mkc<T>(T(0), T(0)); }

template <typename E>
__global__ void __launch_bounds__(FIR_NT)
fir_td_kernel(const E* __restrict__ x, int64_t nx, int64_t tiles_per_col, const E* __restrict__ b, int nb,
              E* __restrict__ out) {
    __shared__ E xs[FIR_TILE + FIR_KC + (FIR_TILE + FIR_KC) / 32 + 2];
    __shared__ E bs[FIR_KC];
    const int tid = threadIdx.x;
    const int64_t col = blockIdx.x / tiles_per_col;
    const int64_t tile = blockIdx.x % tiles_per_col;
    const int64_t i0 = tile * FIR_TILE;
    const E* xc = x + col * nx;
    E* oc = out + col * nx;
    E acc[FIR_OPT];
#pragma unroll
    for (int o = 0; o < FIR_OPT; ++o) acc[o] = fir_zero((E*)nullptr);

    // chunks from the oldest taps (largest k) to the newest
    for (int k_hi = nb - 1; k_hi >= 0; k_hi -= FIR_KC) {
        const int kc = k_hi + 1 < FIR_KC ? k_hi + 1 : FIR_KC;   // taps k_hi, k_hi-1, .., k_hi-kc+1
        const int64_t base = i0 - k_hi;                         // global index of xs[0]
        const int cnt = FIR_TILE + kc - 1;
        __syncthreads();
        for (int j = tid; j < cnt; j += FIR_NT) {
            const int64_t g = base + j;
            xs[fir_pad(j)] = (g >= 0 && g < nx) ? xc[g] : fir_zero((E*)nullptr);
        }
        for (int j = tid; j < kc; j += FIR_NT) bs[j] = b[k_hi - j];
        __syncthreads();
        E w[FIR_OPT];
#pragma unroll
        for (int o = 0; o < FIR_OPT; ++o) w[o] = xs[fir_pad(FIR_OPT * tid + o)];
        for (int kk = 0; kk < kc; ++kk) {
            const E bk = bs[kk];
#pragma unroll
            for (int o = 0; o < FIR_OPT; ++o) acc[o] = fir_fma(w[o], bk, acc[o]);
#pragma unroll
            for (int o = 0; o < FIR_OPT - 1; ++o) w[o] = w[o + 1];
            w[FIR_OPT - 1] = xs[fir_pad(FIR_OPT * tid + FIR_OPT + kk)];
        }
    }
#pragma unroll
    for (int o = 0; o < FIR_OPT; ++o) {
        const int64_t i = i0 + FIR_OPT * tid + o;
        if (i < nx) oc[i] = acc[o];
    }
}

// ---------------------------------------------------------------------------------------------- register-tiled kernel
// fir_td_kernel above spends two shared-memory loads (tap + sample) and a register shift per tap for FIR_OPT = 4
// multiply-adds: ncu (profiles/r2i_fir.txt) shows the FMA pipe at 39 % with issue slots at 82 % -- only a third of the
// instructions are FMAs.  Here a thread owns G consecutive outputs and walks the taps eight at a time, oldest first: the
// 8 taps come from two 128-bit broadcast loads, the G + 7 samples those G x 8 products touch are a 16-register sliding
// window that advances by eight samples (two 128-bit loads) per chunk, so a chunk is 4 + a few instructions for 8 G FMAs.
// The FMA chain of every output is the same as above (one fused multiply-add per tap, oldest tap first): bit-identical.
// Taps are zero-padded at the OLD end to a multiple of eight; the padding taps are skipped, never multiplied (0 * Inf).
// Shared-memory layout: 16 bytes of padding after every 128 bytes, which puts the eight lanes of a 128-bit load phase
// (thread stride G elements = 32 / 64 / 128 bytes) on eight different 16-byte bank groups.
template <typename E> struct fir_geom {
    static constexpr int VEC = 16 / (int)sizeof(E);                    // elements per 128-bit load
    static constexpr int G = sizeof(E) == 4 ? 8 : 4;                   // outputs per thread
    static constexpr int TILE = FIR_NT * G;
    static constexpr int KC = 512;                                     // taps per staging round
    __host__ __device__ static constexpr int pad(int j) { return j + VEC * (j / (8 * VEC)); }
    static constexpr int XS = pad(TILE + KC + 16) + VEC;
};

template <typename E> __device__ __forceinline__ void fir_ld8(E (&dst)[16], int at, const E* __restrict__ xs, int j) {
    using Gm = fir_geom<E>;
#pragma unroll
    for (int v = 0; v < 8; v += Gm::VEC)
        *reinterpret_cast<uint4*>(&dst[at + v]) = *reinterpret_cast<const uint4*>(&xs[Gm::pad(j + v)]);
}

template <typename E>
__global__ void __launch_bounds__(FIR_NT)
fir_tile_kernel(const E* __restrict__ x, int64_t nx, int64_t tiles_per_col, const E* __restrict__ b, int nb,
                E* __restrict__ out) {
    using Gm = fir_geom<E>;
    constexpr int G = Gm::G;
    __shared__ __align__(16) E xs[Gm::XS];
    __shared__ __align__(16) E bs[Gm::KC];
    const int tid = threadIdx.x;
    const int64_t col = blockIdx.x / tiles_per_col;
    const int64_t tile = blockIdx.x % tiles_per_col;
    const int64_t i0 = tile * Gm::TILE;
    const E* xc = x + col * nx;
    E* oc = out + col * nx;
    E acc[G];
#pragma unroll
    for (int o = 0; o < G; ++o) acc[o] = fir_zero((E*)nullptr);
    const int nb8 = (nb + 7) & ~7;                                      // taps nb .. nb8-1 are padding (skipped)
    for (int k_hi = nb8 - 1; k_hi >= 0; k_hi -= Gm::KC) {
        const int kc = k_hi + 1 < Gm::KC ? k_hi + 1 : Gm::KC;           // padded taps k_hi, k_hi-1, .., k_hi-kc+1 (a multiple of 8)
        const int64_t base = i0 - k_hi;                                 // global index of xs[0]
        const int cnt = Gm::TILE + kc + 8;
        __syncthreads();
        for (int j = tid; j < cnt; j += FIR_NT) {
            const int64_t g = base + j;
            xs[Gm::pad(j)] = (g >= 0 && g < nx) ? xc[g] : fir_zero((E*)nullptr);
        }
        for (int j = tid; j < kc; j += FIR_NT) bs[j] = (k_hi - j < nb) ? b[k_hi - j] : fir_zero((E*)nullptr);
        __syncthreads();
        // output g, padded tap k_hi - c - q  <->  sample xs[G*tid + c + g + q]
        E w[16];
        fir_ld8<E>(w, 0, xs, G * tid);
        for (int c = 0; c < kc; c += 8) {
            fir_ld8<E>(w, 8, xs, G * tid + c + 8);
            E t[8];
#pragma unroll
            for (int v = 0; v < 8; v += Gm::VEC) *reinterpret_cast<uint4*>(&t[v]) = *reinterpret_cast<const uint4*>(&bs[c + v]);
            if (k_hi - c < nb) {                                        // no padding tap in this chunk (all but the very first)
#pragma unroll
                for (int q = 0; q < 8; ++q)
#pragma unroll
                    for (int o = 0; o < G; ++o) acc[o] = fir_fma(w[o + q], t[q], acc[o]);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (k_hi - c - q < nb) {
#pragma unroll
                        for (int o = 0; o < G; ++o) acc[o] = fir_fma(w[o + q], t[q], acc[o]);
                    }
            }
#pragma unroll
            for (int v = 0; v < 8; ++v) w[v] = w[v + 8];
        }
    }
    const int64_t i = i0 + (int64_t)G * tid;
    if (i + G <= nx && (reinterpret_cast<uintptr_t>(oc + i) & 15) == 0) {
#pragma unroll
        for (int v = 0; v < G; v += Gm::VEC) *reinterpret_cast<uint4*>(oc + i + v) = *reinterpret_cast<const uint4*>(&acc[v]);
    } else {
#pragma unroll
        for (int o = 0; o < G; ++o)
            if (i + o < nx) oc[i + o] = acc[o];
    }
}

struct FirPlanImpl {
    int dtype = 0;
    int64_t nb = 0;
    int device = 0;
    void* d_b = nullptr;
    DevBuf in, out;
    cudaStream_t stream = nullptr;
};

}  // namespace dspb200

using namespace dspb200;

struct dspb200_fir_plan {
    FirPlanImpl impl;
};

extern "C" {

int dspb200_fir_plan_create(dspb200_fir_plan** plan, int dtype, const void* b_host, int64_t nb) {
    DSP_RANGE("dspb200_fir_plan_create");
    DSP_REQUIRE(plan != nullptr, "plan is NULL");
    *plan = nullptr;
    DSP_REQUIRE(dtype_valid(dtype), "invalid dtype %d", dtype);
    DSP_REQUIRE(b_host != nullptr && nb >= 1, "filter vector b must be non-empty");   // ArgumentError src/dspbase.jl:28
    DSP_REQUIRE(nb < (int64_t(1) << 30), "filter too long");
    dspb200_fir_plan* h = new (std::nothrow) dspb200_fir_plan();
    DSP_REQUIRE(h != nullptr, "out of host memory");
    FirPlanImpl* p = &h->impl;
    p->dtype = dtype; p->nb = nb;
    cudaError_t e = cudaGetDevice(&p->device);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_b, (size_t)nb * dtype_size(dtype));
    if (e == cudaSuccess) e = cudaMemcpy(p->d_b, b_host, (size_t)nb * dtype_size(dtype), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { const int rc = cuda_fail(e, "tap upload", __FILE__, __LINE__); dspb200_fir_plan_destroy(h); return rc; }
    *plan = h;
    return DSPB200_OK;
}

int dspb200_fir_exec_dev(dspb200_fir_plan* plan, const void* x, int64_t nx, int64_t ncols, void* out, void* stream) {
    DSP_RANGE("dspb200_fir_exec_dev");
    DSP_REQUIRE(plan != nullptr, "plan is NULL");
    DSP_REQUIRE(nx >= 0 && ncols >= 0, "negative size");
    if (nx == 0 || ncols == 0) return DSPB200_OK;
    DSP_REQUIRE(x && out, "NULL argument");
    FirPlanImpl* p = &plan->impl;
    static const bool tiled = [] { const char* e = getenv("DSPB200_FIR_TILE"); return !(e && e[0] == '0'); }();
    cudaStream_t st = (cudaStream_t)stream;
    if (tiled) {
#define FIR_TILED(E_) do {                                                                                   \
            const int64_t tiles = cdiv(nx, fir_geom<E_>::TILE), blocks = tiles * ncols;                          \
            DSP_REQUIRE(blocks < (int64_t)0x7fffffff, "too many tiles for one launch");                          \
            fir_tile_kernel<E_><<<(unsigned)blocks, FIR_NT, 0, st>>>((const E_*)x, nx, tiles, (const E_*)p->d_b, (int)p->nb, (E_*)out); \
        } while (0)
        switch (p->dtype) {
            case DSPB200_F32: FIR_TILED(float); break;
            case DSPB200_F64: FIR_TILED(double); break;
            case DSPB200_C32: FIR_TILED(cx<float>); break;
            default: FIR_TILED(cx<double>); break;
        }
#undef FIR_TILED
        DSP_LAUNCH_OK();
        return DSPB200_OK;
    }
    const int64_t tiles = cdiv(nx, FIR_TILE);
    const int64_t blocks = tiles * ncols;
    DSP_REQUIRE(blocks < (int64_t)0x7fffffff, "too many tiles for one launch");
    switch (p->dtype) {
        case DSPB200_F32: fir_td_kernel<float><<<(unsigned)blocks, FIR_NT, 0, st>>>((const float*)x, nx, tiles, (const float*)p->d_b, (int)p->nb, (float*)out); break;
        case DSPB200_F64: fir_td_kernel<double><<<(unsigned)blocks, FIR_NT, 0, st>>>((const double*)x, nx, tiles, (const double*)p->d_b, (int)p->nb, (double*)out); break;
        case DSPB200_C32: fir_td_kernel<cx<float>><<<(unsigned)blocks, FIR_NT, 0, st>>>((const cx<float>*)x, nx, tiles, (const cx<float>*)p->d_b, (int)p->nb, (cx<float>*)out); break;
        default: fir_td_kernel<cx<double>><<<(unsigned)blocks, FIR_NT, 0, st>>>((const cx<double>*)x, nx, tiles, (const cx<double>*)p->d_b, (int)p->nb, (cx<double>*)out); break;
    }
    DSP_LAUNCH_OK();
    return DSPB200_OK;
}

int dspb200_fir_exec(dspb200_fir_plan* plan, const void* x, int64_t nx, int64_t ncols, void* out) {
    DSP_RANGE("dspb200_fir_exec");
    DSP_REQUIRE(plan != nullptr, "plan is NULL");
    DSP_REQUIRE(nx >= 0 && ncols >= 0, "negative size");
    if (nx == 0 || ncols == 0) return DSPB200_OK;
    DSP_REQUIRE(x && out, "NULL argument");
    FirPlanImpl* p = &plan->impl;
    DSP_CUDA(cudaSetDevice(p->device));
    if (!p->stream) DSP_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    const size_t bytes = (size_t)(nx * ncols) * dtype_size(p->dtype);
    DSP_TRY(p->in.reserve(bytes));
    DSP_TRY(p->out.reserve(bytes));
    DSP_CUDA(cudaMemcpyAsync(p->in.p, x, bytes, cudaMemcpyHostToDevice, p->stream));
    DSP_TRY(dspb200_fir_exec_dev(plan, p->in.p, nx, ncols, p->out.p, p->stream));
    DSP_CUDA(cudaMemcpyAsync(out, p->out.p, bytes, cudaMemcpyDeviceToHost, p->stream));
    DSP_CUDA(cudaStreamSynchronize(p->stream));
    return DSPB200_OK;
}

int dspb200_fir_plan_destroy(dspb200_fir_plan* plan) {
    if (!plan) return DSPB200_OK;
    FirPlanImpl* p = &plan->impl;
    if (p->d_b) cudaFree(p->d_b);
    p->in.release(); p->out.release();
    if (p->stream) cudaStreamDestroy(p->stream);
    delete plan;
    return DSPB200_OK;
}

}  // extern "C"

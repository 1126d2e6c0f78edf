// Microbenchmark: scalar FFMA/FADD vs packed FFMA2/FADD2 (sm_100 f32x2) issue throughput.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE, int WARPS_PER_BLOCK>
__global__ void kern(float* out, int iters, float a, float b) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { acc[i].x = fmaf(acc[i].x, aa.x, bb.x); acc[i].y = fmaf(acc[i].y, aa.y, bb.y); }     // 2 FFMA
            if (MODE == 1) { acc[i] = __ffma2_rn(acc[i], aa, bb); }                                                // 1 FFMA2
            if (MODE == 2) { acc[i].x = acc[i].x + bb.x; acc[i].y = acc[i].y + bb.y; }                             // 2 FADD
            if (MODE == 3) { acc[i] = __fadd2_rn(acc[i], bb); }                                                    // 1 FADD2
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, int warps_per_sm) {
    const int threads = 256, blocks = 148 * warps_per_sm * 32 / threads, iters = 20000;
    float* out;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    kern<MODE, 8><<<blocks, threads>>>(out, 100, 1.0001f, 0.5f);
    cudaEventRecord(a);
    kern<MODE, 8><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double lane_ops = (double)blocks * threads * iters * 16;   // 16 scalar results per iteration per thread
    printf("%-8s warps/SM %2d  %.3f ms  %.1f G scalar-ops/s  (%.2f per clk per SM @1.965GHz)\n", name, warps_per_sm, ms,
           lane_ops / ms / 1e6, lane_ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}

int main() {
    for (int w : {8, 16, 32, 64}) {
        run<0>("FFMA", w); run<1>("FFMA2", w); run<2>("FADD", w); run<3>("FADD2", w);
    }
    return 0;
}

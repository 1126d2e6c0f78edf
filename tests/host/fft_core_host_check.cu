// Host-side emulation of the shared-memory FFT core (no GPU needed): every pass of fft_forward /
// fft_adjoint is run for all "threads" sequentially, exactly as the kernels sequence them between
// __syncthreads(), and compared with a direct O(N^2)/recursive double-precision DFT.
// Build (host compiler only): g++ -std=c++17 -O2 -x c++ -I/usr/local/cuda/include fft_core_host_check.cu
// (run by tests/test_host_fft.py)
#ifndef __CUDACC__
static inline void __syncthreads() {}
#endif
#include "../../dsp.jl_b200/csrc/fft_core.cuh"
#include <complex>
#include <vector>
#include <cstdio>
#include <cmath>
#include <cstdlib>

using namespace dspb200;
namespace dspb200 { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*, const char*, int) { return -2; } void count_launch(int) {} int device_sm_count() { return 148; } }

typedef std::complex<double> cd;

static void ref_fft(std::vector<cd>& a, bool inv) {   // iterative radix-2, double
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = 2 * M_PI / (double)len * (inv ? 1 : -1);
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                cd w = std::polar(1.0, ang * (double)k);
                cd u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v; a[i + k + len / 2] = u - v;
            }
    }
}

template <typename T, int N> struct Emu {
    using P = fft_plan_traits<N>;
    static constexpr int NT = fft_threads<N>::value;
    static constexpr int R0 = P::R0;
    static constexpr int NP = P::NPASS16;
    static constexpr int M1 = N / R0;
    std::vector<cx<T>> sm, tw, t16, t256;
    FftCtx<T> ctx;
    Emu() : sm(padded_len<T>(N)), tw(N), t16(TW16_LEN), t256(TW256_LEN) {
        fft_fill_wn<T>(tw.data(), N);
        fft_fill_tables<T>(t16.data(), t256.data());
        ctx.sm = sm.data(); ctx.tw = tw.data(); ctx.t16 = t16.data(); ctx.t256 = t256.data();
    }
    // radix-16 passes use the grouped thread -> butterfly map exactly as the kernels do
    template <int M, int R, bool DIT, class Ld, class St> void pass(Ld ld, St st) {
        for (int tid = 0; tid < NT; ++tid) fft_pass<T, N, NT, M, R, DIT, ((R == 16 && (N / 16) % (2 * NT) == 0) ? 2 : 1), (R == 16)>(ctx, tid, ld, st);
    }
    // forward: x natural -> regs[slot] (digit-reversed slots)
    void forward(const std::vector<cx<T>>& x, std::vector<cx<T>>& last) {
        SmemLd<T> sld{sm.data()};
        SmemSt<T> sst{sm.data()};
        auto ld0 = [&](int j, int, int, int) { return x[j]; };
        auto stl = [&](int slot, int, int, int, cx<T> v) { last[slot] = v; };
        pass<N, R0, false>(ld0, sst);
        if constexpr (NP == 1) pass<M1, 16, false>(sld, stl);
        else if constexpr (NP == 2) { pass<M1, 16, false>(sld, sst); pass<M1 / 16, 16, false>(sld, stl); }
        else { pass<M1, 16, false>(sld, sst); pass<M1 / 16, 16, false>(sld, sst); pass<M1 / 256, 16, false>(sld, stl); }
    }
    // adjoint: regs (digit-reversed slots) -> y natural (swapped domain handled by caller)
    void adjoint(const std::vector<cx<T>>& first, std::vector<cx<T>>& y) {
        SmemLd<T> sld{sm.data()};
        SmemSt<T> sst{sm.data()};
        auto ldf = [&](int slot, int, int, int) { return first[slot]; };
        auto st0 = [&](int j, int, int, int, cx<T> v) { y[j] = v; };
        if constexpr (NP == 1) pass<M1, 16, true>(ldf, sst);
        else if constexpr (NP == 2) { pass<M1 / 16, 16, true>(ldf, sst); pass<M1, 16, true>(sld, sst); }
        else { pass<M1 / 256, 16, true>(ldf, sst); pass<M1 / 16, 16, true>(sld, sst); pass<M1, 16, true>(sld, sst); }
        pass<N, R0, true>(sld, st0);
    }
};

template <typename T, int N> static int check(double tol) {
    Emu<T, N>* e = new Emu<T, N>();
    std::vector<cx<T>> x(N), last(N), y(N);
    std::vector<cd> xr(N);
    srand(1234 + N);
    for (int j = 0; j < N; ++j) {
        double a = rand() / (double)RAND_MAX - 0.5, b = rand() / (double)RAND_MAX - 0.5;
        x[j] = mkc<T>((T)a, (T)b);
        xr[j] = cd((double)x[j].x, (double)x[j].y);
    }
    e->forward(x, last);
    std::vector<cd> X = xr;
    ref_fft(X, false);
    double num = 0, den = 0;
    for (int k = 0; k < N; ++k) {
        cx<T> v = last[digit_reverse<N>(k)];
        num += std::norm(cd((double)v.x, (double)v.y) - X[k]);
        den += std::norm(X[k]);
    }
    const double ef = std::sqrt(num / den);
    // inverse through the swap trick: y = swap(adjoint(swap(last)))  == N * x
    std::vector<cx<T>> sw(N);
    for (int i = 0; i < N; ++i) sw[i] = cswap(last[i]);
    e->adjoint(sw, y);
    num = den = 0;
    for (int j = 0; j < N; ++j) {
        cx<T> v = cswap(y[j]);
        num += std::norm(cd((double)v.x, (double)v.y) / (double)N - xr[j]);
        den += std::norm(xr[j]);
    }
    const double ei = std::sqrt(num / den);
    // bank-conflict audit of the padded layout for the stride-1 and stride-16 radix-16 passes (8-byte words)
    printf("N=%5d %s  forward relerr %.3e  roundtrip relerr %.3e  %s\n", N, sizeof(T) == 4 ? "f32" : "f64", ef, ei,
           (ef < tol && ei < tol) ? "ok" : "FAIL");
    delete e;
    return (ef < tol && ei < tol) ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += check<float, 32>(5e-7);
    bad += check<float, 64>(5e-7);
    bad += check<float, 128>(5e-7);
    bad += check<float, 256>(5e-7);
    bad += check<float, 512>(5e-7);
    bad += check<float, 1024>(5e-7);
    bad += check<float, 2048>(5e-7);
    bad += check<float, 4096>(5e-7);
    bad += check<float, 8192>(5e-7);
    bad += check<float, 16384>(5e-7);
    bad += check<double, 256>(1e-15);
    bad += check<double, 1024>(1e-15);
    bad += check<double, 2048>(1e-15);
    bad += check<double, 4096>(1e-15);
    bad += check<double, 8192>(1e-15);
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad;
}

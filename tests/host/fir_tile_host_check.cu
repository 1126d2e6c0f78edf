// Host-side emulation of the register-tiled FIR kernel (dsp.jl_b200/csrc/fir_tile.cuh; no GPU needed): the staging and
// the multiply-add rounds of fir_tile_kernel are run for every "thread" of a CTA in turn, exactly as the kernel sequences
// them between its barriers, and every output is compared BIT FOR BIT with the literal chain of the reference
// (src/dspbase.jl:95-105: one fused multiply-add per tap, oldest tap first).  Shared memory is poisoned with NaNs before
// every round, so a read of a slot the round did not stage shows up as a mismatch; the 128-bit load phases of the layout
// are audited for bank conflicts.
// Build (host compiler only): g++ -std=c++17 -O2 -march=native -x c++ -I/usr/local/cuda/include fir_tile_host_check.cu
// (run by tests/test_host_logic.py)
#include "../../dsp.jl_b200/csrc/fir_tile.cuh"
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <limits>

using namespace dspb200;
namespace dspb200 { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*, const char*, int) { return -2; } void count_launch(int) {} int device_sm_count() { return 148; } }

static unsigned long long rng_state = 88172645463325252ULL;
static double rnd() {                                     // xorshift, uniform in (-1, 1)
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / (double)(1ULL << 53) * 2.0 - 1.0;
}
template <typename T> static void fill(T& v) { v = (T)rnd(); }
template <typename T> static void fill(cx<T>& v) { v.x = (T)rnd(); v.y = (T)rnd(); }
template <typename T> static void poison(T& v) { v = std::numeric_limits<T>::quiet_NaN(); }
template <typename T> static void poison(cx<T>& v) { v.x = v.y = std::numeric_limits<T>::quiet_NaN(); }

template <typename E, int NT> static int run_case(int nb, long long nx) {
    using Gm = fir_geom<E, NT>;
    constexpr int G = Gm::G;
    std::vector<E> x(nx), b(nb), y(nx), ref(nx);
    for (auto& v : x) fill(v);
    for (auto& v : b) fill(v);
    // the literal chain
    for (long long i = 0; i < nx; ++i) {
        E acc = fir_zero((E*)nullptr);
        for (int k = nb - 1; k >= 0; --k) acc = fir_fma(i - k >= 0 ? x[i - k] : fir_zero((E*)nullptr), b[k], acc);
        ref[i] = acc;
    }
    // the kernel, one CTA after the other
    E* xs = (E*)aligned_alloc(16, ((sizeof(E) * Gm::XS + 15) / 16) * 16);
    E* bs = (E*)aligned_alloc(16, sizeof(E) * Gm::KC);
    std::vector<E> acc((size_t)NT * G);
    const long long tiles = (nx + Gm::TILE - 1) / Gm::TILE;
    const int nb8 = (nb + 7) & ~7;
    for (long long tile = 0; tile < tiles; ++tile) {
        const long long i0 = tile * Gm::TILE;
        for (auto& v : acc) v = fir_zero((E*)nullptr);
        for (int k_hi = nb8 - 1; k_hi >= 0; k_hi -= Gm::KC) {
            const int kc = k_hi + 1 < Gm::KC ? k_hi + 1 : Gm::KC;
            for (int j = 0; j < Gm::XS; ++j) poison(xs[j]);
            for (int j = 0; j < Gm::KC; ++j) poison(bs[j]);
            for (int tid = 0; tid < NT; ++tid)
                fir_stage<E, NT>(tid, xs, bs, x.data(), nx, i0 - k_hi, Gm::TILE + kc + 8, b.data(), nb, k_hi, kc);
            for (int tid = 0; tid < NT; ++tid) {
                E (&a)[G] = *reinterpret_cast<E (*)[G]>(&acc[(size_t)tid * G]);
                fir_round<E, NT>(tid, a, xs, bs, nb, k_hi, kc);
            }
        }
        for (int tid = 0; tid < NT; ++tid)
            for (int o = 0; o < G; ++o) {
                const long long i = i0 + (long long)G * tid + o;
                if (i < nx) y[i] = acc[(size_t)tid * G + o];
            }
    }
    free(xs); free(bs);
    int bad = 0;
    for (long long i = 0; i < nx; ++i)
        if (memcmp(&y[i], &ref[i], sizeof(E)) != 0) {
            if (++bad <= 3) printf("  mismatch: sizeof(E)=%d NT=%d nb=%d nx=%lld at i=%lld\n", (int)sizeof(E), NT, nb, nx, i);
        }
    return bad;
}

// eight lanes of a 128-bit load phase (consecutive threads) must hit eight different 16-byte bank groups (mod 128 bytes)
template <typename E, int NT> static int audit_banks() {
    using Gm = fir_geom<E, NT>;
    int bad = 0;
    for (int half = 0; half < 2; ++half)                   // elements 0..3 and 4..7 of a run
        for (int t0 = 0; t0 < NT; t0 += 8) {
            unsigned seen = 0;
            int distinct = 0;
            for (int l = 0; l < 8; ++l) {
                const int j = Gm::G * (t0 + l) + 4 * half;
                const long long byte = (long long)Gm::pos(j) * (long long)sizeof(E);
                if (byte % 16) { ++bad; continue; }
                const unsigned bit = 1u << ((byte % 128) / 16);
                if (!(seen & bit)) ++distinct;
                seen |= bit;
            }
            // 16-byte elements (G = 4): the second half of a run has one two-way conflict per phase (fir_tile.cuh)
            if (distinct < (sizeof(E) == 16 && half == 1 ? 7 : 8)) ++bad;
        }
    if (bad) printf("  bank audit failed: sizeof(E)=%d NT=%d (%d phases)\n", (int)sizeof(E), NT, bad);
    return bad;
}

template <typename E, int NT> static int run_all(const char* name) {
    using Gm = fir_geom<E, NT>;
    static const int nbs[] = {1, 2, 7, 8, 9, 19, 66, 67, 257, 511, 512, 513, 520, 1030};
    const long long nxs[] = {1, 5, Gm::TILE - 1, Gm::TILE, Gm::TILE + 3, 2 * Gm::TILE + 17};
    int bad = audit_banks<E, NT>();
    int cases = 0;
    for (int nb : nbs)
        for (long long nx : nxs) {
            if ((long long)nb * nx > 3000000) continue;    // keeps the whole check to a few seconds
            bad += run_case<E, NT>(nb, nx);
            ++cases;
        }
    bad += run_case<E, NT>(1500, Gm::TILE + 40);           // three staging rounds
    printf("%s NT=%d: %d cases, %d mismatches\n", name, NT, cases + 1, bad);
    return bad;
}

int main() {
    int bad = 0;
    bad += run_all<float, 256>("Float32");
    bad += run_all<float, 128>("Float32");
    bad += run_all<double, 256>("Float64");
    bad += run_all<double, 128>("Float64");
    bad += run_all<cx<float>, 256>("ComplexF32");
    bad += run_all<cx<float>, 128>("ComplexF32");
    bad += run_all<cx<double>, 256>("ComplexF64");
    bad += run_all<cx<double>, 128>("ComplexF64");
    printf(bad ? "FAIL\n" : "OK\n");
    return bad ? 1 : 0;
}

// dspb200 -- block-cooperative power-of-two FFT out of shared memory (sm_100a).
//
// Decimation-in-frequency forward passes / decimation-in-time inverse passes, each pass truly in place
// (a butterfly reads and writes the same R shared-memory slots), so one N-element buffer suffices and a
// thread may own several butterflies per pass.  Radix plan: N = R0 * 16^k with R0 in {2,4,8,16}; every
// pass except possibly the first is radix-16, so the pass strides are >= 16 or exactly 1 and the padded
// layout below is bank-conflict free for the 8-byte (strided passes) and 16-byte (stride-1 runs) accesses used.
//
// After the forward transform X[k] sits at the digit-reversed slot pos(k) (see digit_reverse()).
// The inverse consumes exactly that order and returns natural order, so forward -> pointwise multiply ->
// inverse (overlap-save) needs no reordering pass; consumers that need natural order read smem[pos(k)].
//
// The inverse is computed with the swap trick: IDFT(x) = swap(DFT-style adjoint passes(swap(x))), so only
// forward twiddles W = exp(-2 pi i j / N) and forward butterflies exist.
#pragma once
#include "common.cuh"
#include <math.h>

// butterflies of the shared-memory radix-16 passes of the conv pipeline are unrolled by two where a thread owns
// two of them (LDS of the second overlaps the math of the first): 764 -> 751 us on the 2^26 conv
#ifndef DSP_FFT_UNROLL16
#define DSP_FFT_UNROLL16 2
#endif

namespace dspb200 {

// ---------------------------------------------------------------------------------------------- layout
// Padded slot address: PADK pad elements per 16 and per 256 slots.  Float32 uses PADK = 2 so that every run of
// slots a thread touches together (the 16 contiguous slots of a stride-1 butterfly) starts 16-byte aligned and moves
// as 128-bit pairs; with 128-bit accesses (a quarter warp per wavefront) the lane stride of 18 slots is bank-conflict
// free.  Float64 elements are 16 bytes already and keep PADK = 1 (lane stride 17).
// (Measured: spectrogram -4 %, middle pass of the overlap-save kernel -2 %, profiles/README.md.)
#ifndef DSP_PADK_F32
#define DSP_PADK_F32 2
#endif
template <typename T> struct fft_pad { static constexpr int K = sizeof(T) == 4 ? DSP_PADK_F32 : 1; };
// Pad per 256 slots: one unit (K slots), except for N = 1024 (first radix 4, first-digit stride 256 slots) where it is
// two units.  It matters only where the lanes of one wavefront sit in different 256-slot blocks -- the STFT emit step,
// whose lanes (consecutive bins k) read row (k mod 4) * 16 + ((k / 4) mod 16) of the slot-ordered spectrum: a quarter
// warp covers 4 blocks x 2 rows, and with one unit per block its eight 16-byte reads fall on only 4 distinct bank
// groups (ncu on the 1024-point spectrogram kernel: 20 % of all shared-memory wavefronts were conflict replays).
template <typename T> __host__ __device__ constexpr int fft_pad256(int n) { return fft_pad<T>::K * (n == 1024 ? 2 : 1); }
template <typename T, int N> __host__ __device__ __forceinline__ constexpr int padaddr(int p) {
    return p + fft_pad<T>::K * (p >> 4) + fft_pad256<T>(N) * (p >> 8);
}
template <typename T> __host__ __device__ constexpr int padded_len(int n) {
    // last slot + 1, rounded up to a multiple of 4 so that what follows stays 16-byte aligned (exact: the complex
    // 4096-point Welch kernel fits its window table next to two resident CTAs by 8 bytes)
    return ((n - 1) + fft_pad<T>::K * ((n - 1) >> 4) + fft_pad256<T>(n) * ((n - 1) >> 8) + 1 + 3) & ~3;
}
template <typename T, int N> __host__ __device__ constexpr int padded_stride(int S) {
    return S + fft_pad<T>::K * (S >> 4) + fft_pad256<T>(N) * (S >> 8);
}

// two adjacent complex values (16-byte aligned for Float32) in one shared-memory access
template <typename T> __host__ __device__ __forceinline__ void lds2(const cx<T>* p, cx<T>& a, cx<T>& b) {
#ifdef __CUDA_ARCH__
    if constexpr (sizeof(T) == 4 && fft_pad<T>::K == 2) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        a = mkc<T>(v.x, v.y); b = mkc<T>(v.z, v.w);
        return;
    }
#endif
    a = p[0]; b = p[1];
}
template <typename T> __host__ __device__ __forceinline__ void sts2(cx<T>* p, cx<T> a, cx<T> b) {
#ifdef __CUDA_ARCH__
    if constexpr (sizeof(T) == 4 && fft_pad<T>::K == 2) {
        *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y);
        return;
    }
#endif
    p[0] = a; p[1] = b;
}

template <int N> struct fft_plan_traits {
    static_assert((N & (N - 1)) == 0 && N >= 16, "N must be a power of two >= 16");
    static constexpr int log2n() { int l = 0; for (int n = N; n > 1; n >>= 1) ++l; return l; }
    static constexpr int LOGN = log2n();
    static constexpr int R0 = 1 << (LOGN % 4 == 0 ? 4 : LOGN % 4);   // first radix: 2, 4, 8 or 16
    static constexpr int NPASS16 = (LOGN - (LOGN % 4 == 0 ? 4 : LOGN % 4)) / 4;  // radix-16 passes after it
};

// slot of natural index k after the forward transform: digits of k (least significant first, bases
// R0,16,16,..) become most significant first.
template <int N> __host__ __device__ __forceinline__ int digit_reverse(int k) {
    constexpr int R0 = fft_plan_traits<N>::R0;
    constexpr int NP = fft_plan_traits<N>::NPASS16;
    int pos = (k & (R0 - 1)) * (N / R0);
    int rest = k / R0;
    int sub = N / R0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        sub >>= 4;
        pos += (rest & 15) * sub;
        rest >>= 4;
    }
    return pos;
}

// ---------------------------------------------------------------------------------------------- butterflies
template <typename T> struct fft_const;
template <> struct fft_const<float> {
    static constexpr float SQH = 0.70710678118654752440f;  // sqrt(1/2)
    static constexpr float C8 = 0.92387953251128675613f;   // cos(pi/8)
    static constexpr float S8 = 0.38268343236508977173f;   // sin(pi/8)
};
template <> struct fft_const<double> {
    static constexpr double SQH = 0.70710678118654752440;
    static constexpr double C8 = 0.92387953251128675613;
    static constexpr double S8 = 0.38268343236508977173;
};

template <typename T> __host__ __device__ __forceinline__ void dft2(cx<T>& a, cx<T>& b) {
    cx<T> t = a; a = t + b; b = t - b;
}
template <typename T> __host__ __device__ __forceinline__ void dft4(cx<T>& a0, cx<T>& a1, cx<T>& a2, cx<T>& a3) {
    cx<T> t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = mul_mi(a1 - a3);
    a0 = t0 + t2; a2 = t0 - t2; a1 = t1 + t3; a3 = t1 - t3;
}
// x * W8^1 = x * (1 - i)/sqrt2 ; x * W8^3 = x * (-1 - i)/sqrt2
// written as h * (a + (-i a)) and h * ((-i a) - a) so that the Float32 device build is one FADD2 (with the swap/negate
// operand modifier) plus one FMUL2
template <typename T> __host__ __device__ __forceinline__ cx<T> mul_w8_1(cx<T> a) {
    const T h = fft_const<T>::SQH; return cscale(a + mul_mi(a), h);
}
template <typename T> __host__ __device__ __forceinline__ cx<T> mul_w8_3(cx<T> a) {
    const T h = fft_const<T>::SQH; return cscale(mul_mi(a) - a, h);
}
template <typename T> __host__ __device__ __forceinline__ void dft8(cx<T>& a0, cx<T>& a1, cx<T>& a2, cx<T>& a3,
                                                            cx<T>& a4, cx<T>& a5, cx<T>& a6, cx<T>& a7) {
    dft4(a0, a2, a4, a6);   // E0..E3 in a0,a2,a4,a6
    dft4(a1, a3, a5, a7);   // O0..O3 in a1,a3,a5,a7
    cx<T> o1 = mul_w8_1(a3), o2 = mul_mi(a5), o3 = mul_w8_3(a7);
    cx<T> x0 = a0 + a1, x4 = a0 - a1;
    cx<T> x1 = a2 + o1, x5 = a2 - o1;
    cx<T> x2 = a4 + o2, x6 = a4 - o2;
    cx<T> x3 = a6 + o3, x7 = a6 - o3;
    a0 = x0; a1 = x1; a2 = x2; a3 = x3; a4 = x4; a5 = x5; a6 = x6; a7 = x7;
}
template <typename T> __host__ __device__ __forceinline__ void dft16(cx<T> (&v)[16]) {
    dft8(v[0], v[2], v[4], v[6], v[8], v[10], v[12], v[14]);   // E0..E7 in v[0],v[2],..,v[14]
    dft8(v[1], v[3], v[5], v[7], v[9], v[11], v[13], v[15]);   // O0..O7 in v[1],v[3],..,v[15]
    const T c = fft_const<T>::C8, s = fft_const<T>::S8;
    // O_k *= W16^k, W16 = exp(-i pi/8)
    cx<T> o0 = v[1];
    cx<T> o1 = cmul(v[3], mkc<T>(c, -s));
    cx<T> o2 = mul_w8_1(v[5]);
    cx<T> o3 = cmul(v[7], mkc<T>(s, -c));
    cx<T> o4 = mul_mi(v[9]);
    cx<T> o5 = cmul(v[11], mkc<T>(-s, -c));
    cx<T> o6 = mul_w8_3(v[13]);
    cx<T> o7 = cmul(v[15], mkc<T>(-c, -s));
    cx<T> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], e4 = v[8], e5 = v[10], e6 = v[12], e7 = v[14];
    v[0] = e0 + o0; v[8] = e0 - o0;
    v[1] = e1 + o1; v[9] = e1 - o1;
    v[2] = e2 + o2; v[10] = e2 - o2;
    v[3] = e3 + o3; v[11] = e3 - o3;
    v[4] = e4 + o4; v[12] = e4 - o4;
    v[5] = e5 + o5; v[13] = e5 - o5;
    v[6] = e6 + o6; v[14] = e6 - o6;
    v[7] = e7 + o7; v[15] = e7 - o7;
}
template <typename T, int R> __host__ __device__ __forceinline__ void dftR(cx<T> (&v)[R]) {
    if constexpr (R == 2) dft2(v[0], v[1]);
    else if constexpr (R == 4) dft4(v[0], v[1], v[2], v[3]);
    else if constexpr (R == 8) dft8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    else dft16(v);
}

// ---------------------------------------------------------------------------------------------- twiddles
// Radix-16 passes need W_M^(t*s), s = 1..15, M = 16*S.  Six values per t are tabulated -- exponents
// t*{1,2,3,4,8,12} -- and the other nine are one complex multiply of two table entries, so every factor is at
// most one rounding away from a correctly rounded table value.  Because S is 16 or 256 for every radix-16
// pass of every supported N, two small tables serve all sizes: T16[t][6] (M = 256) and T256[t][6] (M = 4096).
// They are staged in shared memory (0.75 KB + 12 KB for Float32), so twiddles cost LDS.128s, not L1/L2 loads.
// The first pass, when its radix R0 is 2, 4 or 8, reads W_N^(t*s) from the plain W_N table in global memory
// (consecutive threads -> consecutive t: coalesced), issued ahead of the butterfly.
constexpr int TW16_LEN = 16 * 6;
constexpr int TW256_LEN = 256 * 6;

template <typename T> struct FftCtx {
    cx<T>* sm;                      // padded data buffer, padded_len(N) elements
    const cx<T>* tw;                // W_N^j: global (j < N), or -- when tw_smem -- shared, j < N/R0 (first-pass radix 2 / 4)
    const cx<T>* t16;               // shared (or global): T16
    const cx<T>* t256;              // shared (or global): T256
};

// The first pass (radix R0 over the whole transform) needs W_N^(t*s), t < N/R0.  For R0 = 2 or 4 and large N the W_N^t
// column (N/R0 values, 32 KB for N = 16384) is staged in shared memory and the s = 2, 3 powers are formed by complex
// multiplication: the timing probes showed the three gathered LDGs per butterfly of the global W_N table (which no longer
// fits L1 next to the data buffer) to be the most expensive part of the first and last pass.
// Enabled where the CTA is alone on its SM anyway and the table fits: single precision, N = 16384.
template <typename T, int N> __host__ __device__ constexpr bool fft_tw0_in_smem() {
    return sizeof(T) == 4 && fft_plan_traits<N>::R0 == 4 && N == 16384;
}
template <typename T, int N> __host__ __device__ constexpr int fft_tw0_len() { return fft_tw0_in_smem<T, N>() ? N / fft_plan_traits<N>::R0 : 0; }

// shared-memory footprint of a fused transform of size N (data + twiddle tables), in elements of cx<T>
template <typename T, int N> __host__ __device__ constexpr int fft_smem_elems() {
    return padded_len<T>(N) + ((N >= 256) ? TW16_LEN : 0) + ((N >= 4096) ? TW256_LEN : 0) + fft_tw0_len<T, N>();
}
template <int N> __host__ __device__ constexpr bool fft_uses_t16() { return N >= 256; }
template <int N> __host__ __device__ constexpr bool fft_uses_t256() { return N >= 4096; }

template <typename T> __host__ __device__ __forceinline__ cx<T> ldtw(const cx<T>* __restrict__ tw, int j) {
#ifndef __CUDA_ARCH__
    return tw[j];
#else
    if constexpr (sizeof(T) == 4) {
        float2 v = __ldg(reinterpret_cast<const float2*>(tw) + j);
        return mkc<T>(v.x, v.y);
    } else {
        double2 v = __ldg(reinterpret_cast<const double2*>(tw) + j);
        return mkc<T>(v.x, v.y);
    }
#endif
}

// six tabulated twiddles of one butterfly: w[0..5] = W^(t*{1,2,3,4,8,12}); `row` is 16-byte aligned
template <typename T> __host__ __device__ __forceinline__ void load_tw6(const cx<T>* __restrict__ row, cx<T> (&w)[6]) {
#ifdef __CUDA_ARCH__
    if constexpr (sizeof(T) == 4) {
        const float4* q = reinterpret_cast<const float4*>(row);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float4 v = q[i];
            w[2 * i] = mkc<T>(v.x, v.y);
            w[2 * i + 1] = mkc<T>(v.z, v.w);
        }
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < 6; ++i) w[i] = row[i];
}

template <typename T> __host__ __device__ __forceinline__ void apply_tw6(cx<T> (&v)[16], const cx<T> (&w)[6]) {
    const cx<T> w1 = w[0], w2 = w[1], w3 = w[2], w4 = w[3], w8 = w[4], w12 = w[5];
    v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3);
    v[4] = cmul(v[4], w4);
    v[5] = cmul(v[5], cmul(w4, w1)); v[6] = cmul(v[6], cmul(w4, w2)); v[7] = cmul(v[7], cmul(w4, w3));
    v[8] = cmul(v[8], w8);
    v[9] = cmul(v[9], cmul(w8, w1)); v[10] = cmul(v[10], cmul(w8, w2)); v[11] = cmul(v[11], cmul(w8, w3));
    v[12] = cmul(v[12], w12);
    v[13] = cmul(v[13], cmul(w12, w1)); v[14] = cmul(v[14], cmul(w12, w2)); v[15] = cmul(v[15], cmul(w12, w3));
}

// first-pass twiddles for radix 2 / 4 / 8 from the W_N table: w[s-1] = W_N^(t*s)
template <typename T, int R, bool SMEM = false> __host__ __device__ __forceinline__ void load_tw_first(const cx<T>* __restrict__ tw, int t, cx<T> (&w)[R - 1]) {
    if constexpr (SMEM && R == 2) {
        w[0] = tw[t];
    } else if constexpr (SMEM && R == 4) {
        w[0] = tw[t];
        w[1] = cmul(w[0], w[0]);
        w[2] = cmul(w[1], w[0]);
    } else if constexpr (R == 8) {
        w[0] = ldtw(tw, t); w[1] = ldtw(tw, 2 * t); w[2] = ldtw(tw, 3 * t); w[3] = ldtw(tw, 4 * t);
        w[4] = w[5] = w[6] = w[0];     // filled in by apply (products)
    } else {
#pragma unroll
        for (int s = 1; s < R; ++s) w[s - 1] = ldtw(tw, s * t);
    }
}
template <typename T, int R> __host__ __device__ __forceinline__ void apply_tw_first(cx<T> (&v)[R], const cx<T> (&w)[R - 1]) {
    if constexpr (R == 8) {
        v[1] = cmul(v[1], w[0]); v[2] = cmul(v[2], w[1]); v[3] = cmul(v[3], w[2]); v[4] = cmul(v[4], w[3]);
        v[5] = cmul(v[5], cmul(w[3], w[0])); v[6] = cmul(v[6], cmul(w[3], w[1])); v[7] = cmul(v[7], cmul(w[3], w[2]));
    } else {
#pragma unroll
        for (int s = 1; s < R; ++s) v[s] = cmul(v[s], w[s - 1]);
    }
}

// ---------------------------------------------------------------------------------------------- passes
// One pass over sub-transforms of size M with radix R (stride S = M / R).  Butterfly b = (blk, t) owns slots
// blk*M + t + r*S.  `ld(slot, paddr, it, r)` supplies the inputs and `st(slot, paddr, it, r, value)` takes the
// outputs: `slot` is the logical index, `paddr` the padded shared-memory address (padaddr(base) + r * PS with a
// compile-time PS, because S is a power of 16 -- no per-element address arithmetic), `it` the per-thread
// butterfly counter (compile-time when UNROLL == 0: register accumulators).
//   DIF (forward):   out[s] = (sum_r in[r] W_R^(rs)) * W_M^(ts)
//   DIT (adjoint):   out[r] =  sum_s (in[s] W_M^(ts)) W_R^(rs)
// Thread -> butterfly map.  After the first pass the R0 sub-transforms of size N/R0 are independent until the
// adjoint's last pass, so in the radix-16 passes thread group g = tid / G (G = NT / R0 threads) owns exactly the
// butterflies of sub-transform g: groups then synchronise among themselves only (fft_group_sync) and drift apart,
// which de-phases their load / math / store bursts inside one CTA.
template <int N, int NT> struct fft_groups {
    static constexpr int R0 = fft_plan_traits<N>::R0;
    static constexpr int NB16 = N / 16;
    static constexpr bool enabled = (NT % R0 == 0) && (NB16 % NT == 0) && (NB16 / R0 >= 1) && ((NB16 / R0) % (NT / R0) == 0);
    static constexpr int G = enabled ? NT / R0 : NT;          // threads per group
};
template <int N, int NT, bool GROUPED> __host__ __device__ __forceinline__ int fft_bfly16_index(int tid, int it) {
    if constexpr (GROUPED && fft_groups<N, NT>::enabled) {
        constexpr int G = fft_groups<N, NT>::G;
        constexpr int PER = (N / 16) / fft_groups<N, NT>::R0;   // butterflies per sub-transform
        return (tid / G) * PER + (tid % G) + it * G;
    } else {
        return tid + it * NT;
    }
}
// barrier among the threads of one group (full-CTA barrier when grouping is off)
template <int N, int NT> __device__ __forceinline__ void fft_group_sync(int tid) {
#ifdef __CUDA_ARCH__
    if constexpr (!fft_groups<N, NT>::enabled) {
        __syncthreads();
    } else if constexpr (fft_groups<N, NT>::G <= 32) {
        __syncwarp();
    } else {
        constexpr int G = fft_groups<N, NT>::G;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + tid / G), "r"(G) : "memory");
    }
#endif
}

template <typename T> struct SmemLd {
    static constexpr bool is_smem = true;
    const cx<T>* sm;
    __host__ __device__ __forceinline__ cx<T> operator()(int, int paddr, int, int) const { return sm[paddr]; }
};
template <typename T> struct SmemSt {
    static constexpr bool is_smem = true;
    cx<T>* sm;
    __host__ __device__ __forceinline__ void operator()(int, int paddr, int, int, cx<T> v) const { sm[paddr] = v; }
};
template <class F, class = void> struct fft_is_smem : std::false_type {};
template <class F> struct fft_is_smem<F, std::void_t<decltype(F::is_smem)>> : std::true_type {};

// loads / stores of one butterfly's 16 inputs: vectorised when the source is the padded shared-memory buffer and the
// 16 slots are contiguous (S == 1)
template <typename T, int N, int S, class Ld>
__host__ __device__ __forceinline__ void bfly_load(cx<T> (&v)[16], Ld ld, int base, int pbase, int it) {
    constexpr int PS = padded_stride<T, N>(S);
    if constexpr (S == 1 && fft_is_smem<Ld>::value) {
#pragma unroll
        for (int r = 0; r < 16; r += 2) lds2<T>(ld.sm + pbase + r, v[r], v[r + 1]);
    } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = ld(base + r * S, pbase + r * PS, it, r);
    }
}
template <typename T, int N, int S, class St>
__host__ __device__ __forceinline__ void bfly_store(const cx<T> (&v)[16], St st, int base, int pbase, int it) {
    constexpr int PS = padded_stride<T, N>(S);
    if constexpr (S == 1 && fft_is_smem<St>::value) {
#pragma unroll
        for (int r = 0; r < 16; r += 2) sts2<T>(st.sm + pbase + r, v[r], v[r + 1]);
    } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) st(base + r * S, pbase + r * PS, it, r, v[r]);
    }
}

template <typename T, int N, int NT, int M, int R, bool DIT, int UNROLL = 1, bool GROUPED = false, class Ld, class St>
__host__ __device__ __forceinline__ void fft_pass(const FftCtx<T>& c, int tid, Ld ld, St st) {
    constexpr int S = M / R;
    constexpr int PS = padded_stride<T, N>(S);
    constexpr int NB = N / R;
    constexpr int ITERS = (NB + NT - 1) / NT;
    static_assert(!GROUPED || R == 16, "grouped mapping is for the radix-16 passes");
    static_assert(R != 16 || S == 1 || S == 16 || S == 256, "radix-16 pass with an unsupported stride");
    static_assert(R == 16 || M == N, "only the first pass may have a radix below 16");
    static_assert(S == 1 || (S & 15) == 0, "stride must be 1 or a multiple of 16");
    // UNROLL = 0: fully unrolled; otherwise the butterfly loop is unrolled UNROLL times (1 = rolled).
    constexpr int U = UNROLL == 0 ? ITERS : UNROLL;
    if constexpr (R == 16 && UNROLL == 2 && ITERS % 2 == 0 && NB % NT == 0) {
        // software pipelining by hand: the inputs of TWO butterflies are loaded before either is transformed (the
        // compiler cannot move the second butterfly's shared-memory loads above the first one's stores on its own --
        // it cannot prove the slots are distinct), so the second load burst overlaps the first butterfly's math
#pragma unroll 1
        for (int it = 0; it < ITERS; it += 2) {
            const int b0 = fft_bfly16_index<N, NT, GROUPED>(tid, it);
            const int b1 = fft_bfly16_index<N, NT, GROUPED>(tid, it + 1);
            const int t0 = b0 & (S - 1), t1 = b1 & (S - 1);
            const int base0 = (b0 / S) * M + t0, base1 = (b1 / S) * M + t1;
            const int p0 = padaddr<T, N>(base0), p1 = padaddr<T, N>(base1);
            cx<T> v0[16], v1[16], w0[6], w1[6];
            if constexpr (S == 16) { load_tw6<T>(c.t16 + t0 * 6, w0); load_tw6<T>(c.t16 + t1 * 6, w1); }
            if constexpr (S == 256) { load_tw6<T>(c.t256 + t0 * 6, w0); load_tw6<T>(c.t256 + t1 * 6, w1); }
            bfly_load<T, N, S>(v0, ld, base0, p0, it);
            bfly_load<T, N, S>(v1, ld, base1, p1, it + 1);
            if constexpr (DIT && S > 1) apply_tw6<T>(v0, w0);
            dft16<T>(v0);
            if constexpr (!DIT && S > 1) apply_tw6<T>(v0, w0);
            bfly_store<T, N, S>(v0, st, base0, p0, it);
            if constexpr (DIT && S > 1) apply_tw6<T>(v1, w1);
            dft16<T>(v1);
            if constexpr (!DIT && S > 1) apply_tw6<T>(v1, w1);
            bfly_store<T, N, S>(v1, st, base1, p1, it + 1);
        }
        return;
    }
#pragma unroll(U)
    for (int it = 0; it < ITERS; ++it) {
        const int b = (R == 16) ? fft_bfly16_index<N, NT, GROUPED>(tid, it) : tid + it * NT;
        if (NB % NT != 0 && b >= NB) break;
        const int t = b & (S - 1);
        const int base = (b / S) * M + t;
        const int pbase = padaddr<T, N>(base);
        if constexpr (R == 16) {
            cx<T> v[16];
            cx<T> w[6];
            if constexpr (S == 16) load_tw6<T>(c.t16 + t * 6, w);
            if constexpr (S == 256) load_tw6<T>(c.t256 + t * 6, w);
            bfly_load<T, N, S>(v, ld, base, pbase, it);
            if constexpr (DIT && S > 1) apply_tw6<T>(v, w);
            dft16<T>(v);
            if constexpr (!DIT && S > 1) apply_tw6<T>(v, w);
            bfly_store<T, N, S>(v, st, base, pbase, it);
        } else {
            cx<T> v[R];
            cx<T> w[R - 1];
            load_tw_first<T, R, fft_tw0_in_smem<T, N>()>(c.tw, t, w);
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = ld(base + r * S, pbase + r * PS, it, r);
            if constexpr (DIT) apply_tw_first<T, R>(v, w);
            dftR<T, R>(v);
            if constexpr (!DIT) apply_tw_first<T, R>(v, w);
#pragma unroll
            for (int r = 0; r < R; ++r) st(base + r * S, pbase + r * PS, it, r, v[r]);
        }
    }
}

// Copy the twiddle tables a transform of size N needs from global memory into the shared-memory area behind
// the data buffer and return the context.  Must be followed by __syncthreads() before the first pass that
// uses them (every user has one: the first pass is followed by a barrier, and T16/T256 are first read after it
// unless R0 == 16, in which case call sites sync explicitly).
template <typename T, int N, int NT>
__device__ __forceinline__ FftCtx<T> fft_make_ctx(cx<T>* smem, const cx<T>* __restrict__ tw, const cx<T>* __restrict__ g16,
                                                    const cx<T>* __restrict__ g256, int tid) {
    FftCtx<T> c;
    c.sm = smem;
    c.tw = tw;
    cx<T>* s16 = smem + padded_len<T>(N);
    cx<T>* s256 = s16 + (fft_uses_t16<N>() ? TW16_LEN : 0);
    c.t16 = s16;
    c.t256 = s256;
    if constexpr (fft_uses_t16<N>()) {
        for (int i = tid; i < TW16_LEN; i += NT) s16[i] = g16[i];
    }
    if constexpr (fft_uses_t256<N>()) {
        for (int i = tid; i < TW256_LEN; i += NT) s256[i] = g256[i];
    }
    if constexpr (fft_tw0_in_smem<T, N>()) {
        cx<T>* s0 = s256 + (fft_uses_t256<N>() ? TW256_LEN : 0);
        for (int i = tid; i < fft_tw0_len<T, N>(); i += NT) s0[i] = tw[i];
        c.tw = s0;
        __syncthreads();            // the first pass reads this table
    }
    return c;
}

// Forward passes after the first one (which must be followed by __syncthreads()): radix-16 passes inside the
// independent sub-transforms, group-synchronised; the last one hands its results to `stlast` in registers.
template <typename T, int N, int NT, class StLast>
__device__ __forceinline__ void fft_forward_rest(const FftCtx<T>& c, int tid, StLast stlast) {
    using P = fft_plan_traits<N>;
    constexpr int NP = P::NPASS16;
    constexpr int M1 = N / P::R0;
    SmemLd<T> sld{c.sm};
    SmemSt<T> sst{c.sm};
    if constexpr (NP == 1) {
        fft_pass<T, N, NT, M1, 16, false, 0, true>(c, tid, sld, stlast);
    } else if constexpr (NP == 2) {
        fft_pass<T, N, NT, M1, 16, false, 1, true>(c, tid, sld, sst);
        fft_group_sync<N, NT>(tid);
        fft_pass<T, N, NT, M1 / 16, 16, false, 0, true>(c, tid, sld, stlast);
    } else {
        static_assert(NP == 3, "unsupported N");
        fft_pass<T, N, NT, M1, 16, false, 1, true>(c, tid, sld, sst);
        fft_group_sync<N, NT>(tid);
        fft_pass<T, N, NT, M1 / 16, 16, false, 1, true>(c, tid, sld, sst);
        fft_group_sync<N, NT>(tid);
        fft_pass<T, N, NT, M1 / 256, 16, false, 0, true>(c, tid, sld, stlast);
    }
}

// Forward DIF: first pass from `ld0` (natural order input, slot j = sample j) into smem, middle passes in
// smem, last pass out through `stlast` (slot = digit-reversed position; values stay in registers).
// Needs NPASS16 >= 1 (N >= 32).  All threads of the block must call it; contains __syncthreads().
template <typename T, int N, int NT, class Ld0, class StLast>
__device__ __forceinline__ void fft_forward(const FftCtx<T>& c, int tid, Ld0 ld0, StLast stlast) {
    using P = fft_plan_traits<N>;
    constexpr int R0 = P::R0;
    constexpr int NP = P::NPASS16;
    static_assert(NP >= 1, "N too small for the fused FFT");
    SmemLd<T> sld{c.sm};
    SmemSt<T> sst{c.sm};
    fft_pass<T, N, NT, N, R0, false, 2>(c, tid, ld0, sst);
    __syncthreads();
    fft_forward_rest<T, N, NT>(c, tid, stlast);
}

// Adjoint (inverse, swapped-domain) DIT: first pass from `ldfirst` (digit-reversed slots, typically the
// registers left by fft_forward's last pass), last pass out through `st0` (natural order, slot j).
template <typename T, int N, int NT, class LdFirst, class St0>
__device__ __forceinline__ void fft_adjoint(const FftCtx<T>& c, int tid, LdFirst ldfirst, St0 st0) {
    using P = fft_plan_traits<N>;
    constexpr int R0 = P::R0;
    constexpr int NP = P::NPASS16;
    static_assert(NP >= 1, "N too small for the fused FFT");
    SmemLd<T> sld{c.sm};
    SmemSt<T> sst{c.sm};
    constexpr int M1 = N / R0;
    if constexpr (NP == 1) {
        fft_pass<T, N, NT, M1, 16, true, 0, true>(c, tid, ldfirst, sst);
    } else if constexpr (NP == 2) {
        fft_pass<T, N, NT, M1 / 16, 16, true, 0, true>(c, tid, ldfirst, sst);
        fft_group_sync<N, NT>(tid);
        fft_pass<T, N, NT, M1, 16, true, 1, true>(c, tid, sld, sst);
    } else {
        fft_pass<T, N, NT, M1 / 256, 16, true, 0, true>(c, tid, ldfirst, sst);
        fft_group_sync<N, NT>(tid);
        fft_pass<T, N, NT, M1 / 16, 16, true, 1, true>(c, tid, sld, sst);
        fft_group_sync<N, NT>(tid);
        fft_pass<T, N, NT, M1, 16, true, 1, true>(c, tid, sld, sst);
    }
    __syncthreads();
    fft_pass<T, N, NT, N, R0, true>(c, tid, sld, st0);
}

// ---------------------------------------------------------------------------------------------- conv pipeline
// Forward "head": every DIF pass except the last (stride-1) radix-16 pass; ends with a group barrier (the
// middle pass and the adjoint tail use the same thread -> butterfly map).
template <typename T, int N, int NT, int U16 = DSP_FFT_UNROLL16, class Ld0>
__device__ __forceinline__ void fft_forward_head(const FftCtx<T>& c, int tid, Ld0 ld0) {
    using P = fft_plan_traits<N>;
    constexpr int R0 = P::R0;
    constexpr int NP = P::NPASS16;
    static_assert(NP >= 1, "N too small for the fused FFT");
    SmemLd<T> sld{c.sm};
    SmemSt<T> sst{c.sm};
    constexpr int M1 = N / R0;
    // first pass straight from global memory: enough butterflies in flight to cover the (L2) latency
    fft_pass<T, N, NT, N, R0, false, (R0 == 2 ? 8 : (R0 <= 4 ? 4 : 2))>(c, tid, ld0, sst);
    __syncthreads();
    if constexpr (NP >= 2) {
        fft_pass<T, N, NT, M1, 16, false, U16, true>(c, tid, sld, sst);
        fft_group_sync<N, NT>(tid);
    }
    if constexpr (NP >= 3) {
        fft_pass<T, N, NT, M1 / 16, 16, false, U16, true>(c, tid, sld, sst);
        fft_group_sync<N, NT>(tid);
    }
}

// Middle pass of a frequency-domain product: last forward pass (stride 1, no twiddles), `mul(base, X)`,
// swap into the adjoint domain, first adjoint pass (stride 1, no twiddles) -- all in registers, one
// shared-memory round trip instead of three.  `pre(base)` runs before the shared-memory loads (prefetch hook).
template <typename T, int N, int NT, class Mul>
__host__ __device__ __forceinline__ void fft_mid_pass_nosync(cx<T>* sm, int tid, Mul mul) {
    constexpr int NB = N / 16;
    constexpr int ITERS = (NB + NT - 1) / NT;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const int b = fft_bfly16_index<N, NT, true>(tid, it);
        if (NB % NT != 0 && b >= NB) break;
        const int base = b * 16;
        const int pbase = padaddr<T, N>(base);
        cx<T> v[16];
#pragma unroll
        for (int r = 0; r < 16; r += 2) lds2<T>(sm + pbase + r, v[r], v[r + 1]);
        dft16(v);
        mul(base, v);
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = cswap(v[r]);
        dft16(v);
#pragma unroll
        for (int r = 0; r < 16; r += 2) sts2<T>(sm + pbase + r, v[r], v[r + 1]);
    }
}

// Adjoint "tail": every DIT pass except the first (stride-1) one (entered after a group barrier); the last pass,
// which joins the sub-transforms again, runs after a full barrier and goes out through st0 (natural order).
template <typename T, int N, int NT, int U16 = DSP_FFT_UNROLL16, class St0>
__device__ __forceinline__ void fft_adjoint_tail(const FftCtx<T>& c, int tid, St0 st0) {
    using P = fft_plan_traits<N>;
    constexpr int R0 = P::R0;
    constexpr int NP = P::NPASS16;
    SmemLd<T> sld{c.sm};
    SmemSt<T> sst{c.sm};
    constexpr int M1 = N / R0;
    if constexpr (NP >= 3) {
        fft_pass<T, N, NT, M1 / 16, 16, true, U16, true>(c, tid, sld, sst);
        fft_group_sync<N, NT>(tid);
    }
    if constexpr (NP >= 2) {
        fft_pass<T, N, NT, M1, 16, true, U16, true>(c, tid, sld, sst);
    }
    __syncthreads();
    fft_pass<T, N, NT, N, R0, true>(c, tid, sld, st0);
}

// Host side: fill the two universal twiddle tables (long-double trig, rounded once).
template <typename T> inline void fft_fill_tables(cx<T>* t16, cx<T>* t256) {
    const int mult[6] = {1, 2, 3, 4, 8, 12};
    const long double PI2 = 6.283185307179586476925286766559005768L;
    for (int t = 0; t < 16; ++t)
        for (int m = 0; m < 6; ++m) {
            const long double a = -PI2 * (long double)(t * mult[m]) / 256.0L;
            t16[t * 6 + m] = mkc<T>((T)cosl(a), (T)sinl(a));
        }
    for (int t = 0; t < 256; ++t)
        for (int m = 0; m < 6; ++m) {
            const long double a = -PI2 * (long double)(t * mult[m]) / 4096.0L;
            t256[t * 6 + m] = mkc<T>((T)cosl(a), (T)sinl(a));
        }
}
template <typename T> inline void fft_fill_wn(cx<T>* tw, long long n) {
    const long double PI2 = 6.283185307179586476925286766559005768L;
    for (long long j = 0; j < n; ++j) {
        const long double a = -PI2 * (long double)j / (long double)n;
        tw[j] = mkc<T>((T)cosl(a), (T)sinl(a));
    }
}

// Threads per block for a fused transform of size N: one radix-16 butterfly per thread up to 256 threads.
template <int N> struct fft_threads {
    static constexpr int NB16 = N / 16;
    static constexpr int value = NB16 < 64 ? 64 : (NB16 > 512 ? 512 : (NB16 > 256 ? 256 : NB16));
};
// __launch_bounds__ min-blocks: cap Float32 kernels at 128 registers (512 resident threads per SM at least);
// Float64 butterflies need the full register file.
template <typename T, int N> struct fft_minblocks {
    // (an 80-register cap -> 3 Welch CTAs/SM was measured SLOWER: 251 vs 219 us at N = 4096 -- the third CTA's
    //  shared memory leaves no L1 for the window table and the tighter cap adds instructions)
    static constexpr int value = sizeof(T) == 8 ? 1 : (fft_threads<N>::value >= 512 ? 1 : 512 / fft_threads<N>::value);
};

}  // namespace dspb200

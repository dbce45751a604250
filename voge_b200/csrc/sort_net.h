// sort_net.h -- register sorting networks of select_topk (select.cu), plain C++17 so that the CPU test-suite
// can compile and check them (tests/test_sort_networks_cpu.py): Batcher odd-even merge sort on N registers,
// the bitonic merge of S registers, and the index folding of the gather-blend kernels (blend.cu).
#pragma once
#include <cstddef>
#include <utility>

#if defined(__CUDACC__)
#define VOGE_HD __host__ __device__ __forceinline__
#else
#define VOGE_HD inline
#endif

namespace voge {

VOGE_HD unsigned net_min(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return min(a, b);
#else
    return a < b ? a : b;
#endif
}
VOGE_HD unsigned net_max(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return max(a, b);
#else
    return a > b ? a : b;
#endif
}

// Batcher's odd-even merge sort on N registers: the network of the next power of two without the comparators
// that touch the (implicit, +inf) inputs above N -- 63 / 191 / 384 / 543 compare-exchanges for N = 16 / 32 / 48 /
// 64, each a VIMNMX pair on 32-bit keys.  The comparator list is built at compile time and applied through a
// fold expression, so every register index is a literal (no local-memory array).
constexpr int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

template <typename F>
constexpr void odd_even_comparators(int N, F&& f) {
    const int P = next_pow2(N);
    for (int p = 1; p < P; p <<= 1)
        for (int k = p; k >= 1; k >>= 1)
            for (int j = k % p; j + k < P; j += 2 * k)
                for (int i = 0; i < k; ++i)
                    if ((i + j) / (2 * p) == (i + j + k) / (2 * p) && i + j + k < N) f(i + j, i + j + k);
}

constexpr int odd_even_count(int N) {
    int n = 0;
    odd_even_comparators(N, [&](int, int) { ++n; });
    return n;
}

template <int N>
struct OddEvenNet {
    static constexpr int kMax = odd_even_count(N);
    short lo[kMax], hi[kMax];
    int n;
    constexpr OddEvenNet() : lo{}, hi{}, n(0) {
        odd_even_comparators(N, [&](int a, int b) { lo[n] = (short)a; hi[n] = (short)b; ++n; });
    }
};

template <int N, size_t... I>
VOGE_HD void sort_network_impl(unsigned (&r)[N], std::index_sequence<I...>) {
    constexpr OddEvenNet<N> net{};
    static_assert(net.n == OddEvenNet<N>::kMax, "comparator count");
    ((void)([&] {
         const unsigned x = r[net.lo[I]], y = r[net.hi[I]];
         r[net.lo[I]] = net_min(x, y);
         r[net.hi[I]] = net_max(x, y);
     }()),
     ...);
}

template <int N>
VOGE_HD void sort_network(unsigned (&r)[N]) {
    sort_network_impl<N>(r, std::make_index_sequence<OddEvenNet<N>::kMax>{});
}

// bitonic merge network on S registers (log2 S stages of S/2 compare-exchanges): sorts a bitonic sequence
// ascending.  Comparator I of stage I / (S/2) (stride (S/2) >> stage) is the I % (S/2)-th index with the stride
// bit clear.
constexpr int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

template <int S, size_t... I>
VOGE_HD void bitonic_merge_impl(unsigned (&r)[S], std::index_sequence<I...>) {
    ((void)([&] {
         constexpr int stride = (S / 2) >> (I / (S / 2));
         constexpr int q = I % (S / 2);
         constexpr int lo = ((q / stride) * 2 * stride) + (q % stride);
         const unsigned x = r[lo], y = r[lo + stride];
         r[lo] = net_min(x, y);
         r[lo + stride] = net_max(x, y);
     }()),
     ...);
}

template <int S>
VOGE_HD void bitonic_merge(unsigned (&r)[S]) {
    bitonic_merge_impl<S>(r, std::make_index_sequence<(S / 2) * ilog2(S)>{});
}

// g mod d for 0 <= g < 2^31 without the integer-division sequence: with m = floor((2^32 - 1) / d) (computed on the
// host) q = mulhi(g, m) is floor(g / d) or one less, so one conditional subtraction finishes the remainder.
struct FastMod {
    unsigned d, m;     // d == 0: identity
};
inline FastMod make_fastmod(int d) {
    FastMod f;
    f.d = d > 0 ? (unsigned)d : 0u;
    f.m = d > 0 ? 0xffffffffu / (unsigned)d : 0u;
    return f;
}
VOGE_HD int fold_index(int g, const FastMod f) {
    g = g > 0 ? g : 0;                          // Aggregation.py:131  vert_assign += (vert_assign < 0)
    if (f.d == 0u) return g;
#if defined(__CUDA_ARCH__)
    const unsigned q = __umulhi((unsigned)g, f.m);
#else
    const unsigned q = (unsigned)(((unsigned long long)(unsigned)g * f.m) >> 32);
#endif
    unsigned r = (unsigned)g - q * f.d;
    if (r >= f.d) r -= f.d;
    return (int)r;
}

}  // namespace voge

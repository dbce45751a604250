// Host check of voge_b200/csrc/sort_net.h (compiled by tests/test_sort_networks_cpu.py with g++):
//   * the odd-even networks sort (exhaustive 0/1 inputs for N = 16, random keys + 0/1 samples for N = 32, 48, 64),
//   * the bitonic merge sorts every bitonic 0/1 sequence of 8 / 16 / 32 elements,
//   * the lane-pair scheme of select_topk (two lanes sort S = 8 / 16 / 32 interleaved slots each, min / max against
//     the partner's reversed half, bitonic merge per lane) yields the sorted 2 S -- lanes simulated in lockstep,
//   * fold_index == max(g, 0) % d.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

#include "sort_net.h"

using namespace voge;

template <int N>
static bool is_sorted_arr(const unsigned (&r)[N]) {
    for (int i = 1; i < N; ++i)
        if (r[i - 1] > r[i]) return false;
    return true;
}

template <int N>
static int check_sort(std::mt19937& rng, int samples) {
    int bad = 0;
    unsigned r[N], ref[N];
    for (int s = 0; s < samples; ++s) {
        const int mode = s % 3;
        for (int i = 0; i < N; ++i) r[i] = mode == 0 ? rng() : (mode == 1 ? (rng() & 1u) : (rng() % 7u));
        for (int i = 0; i < N; ++i) ref[i] = r[i];
        std::sort(ref, ref + N);
        sort_network<N>(r);
        for (int i = 0; i < N; ++i) bad += r[i] != ref[i];
    }
    return bad;
}

template <int S>
static int check_bitonic() {
    // every 0/1 bitonic sequence: 0^a 1^b 0^c and 1^a 0^b 1^c
    int bad = 0;
    unsigned r[S];
    for (int inv = 0; inv < 2; ++inv)
        for (int a = 0; a <= S; ++a)
            for (int b = 0; a + b <= S; ++b) {
                for (int i = 0; i < S; ++i) r[i] = ((i >= a && i < a + b) ? 1u : 0u) ^ (unsigned)inv;
                bitonic_merge<S>(r);
                bad += !is_sorted_arr<S>(r);
            }
    return bad;
}

// select_pair<S> (select.cu) with the two lanes of a pixel simulated in lockstep
template <int S>
static int check_pair(std::mt19937& rng, int samples) {
    int bad = 0;
    for (int s = 0; s < samples; ++s) {
        const int c = (int)(rng() % (unsigned)(2 * S + 1));    // 0 .. 2 S hits
        unsigned keys[2 * S], ref[2 * S];
        for (int j = 0; j < 2 * S; ++j) keys[j] = j < c ? (((s & 1) ? (rng() % 50u) : (rng() >> 7)) << 6 | (unsigned)j) : 0xffffffffu;
        for (int j = 0; j < 2 * S; ++j) ref[j] = keys[j];
        std::sort(ref, ref + 2 * S);
        unsigned L[2][S];
        for (int sub = 0; sub < 2; ++sub) {
            for (int i = 0; i < S; ++i) L[sub][i] = keys[2 * i + sub];       // interleaved slots
            sort_network<S>(L[sub]);
        }
        for (int x = 0; x < S / 2; ++x) {
            // both lanes shuffle before either writes (SIMT lockstep)
            const unsigned a1 = L[1][S - 1 - x], a2 = L[1][x], b1 = L[0][S - 1 - x], b2 = L[0][x];
            L[0][x] = net_min(L[0][x], a1); L[0][S - 1 - x] = net_min(L[0][S - 1 - x], a2);
            L[1][x] = net_max(L[1][x], b1); L[1][S - 1 - x] = net_max(L[1][S - 1 - x], b2);
        }
        bitonic_merge<S>(L[0]);
        bitonic_merge<S>(L[1]);
        for (int i = 0; i < S; ++i) bad += (L[0][i] != ref[i]) + (L[1][i] != ref[S + i]);
    }
    return bad;
}

// Four lanes per pixel (select_quad of select.cu, S = 32: segments of 65 .. 128 hits; the kernel keeps only the
// lower half, lanes 0 and 1): S slots per lane, slot 4 i + q in lane q.  Round 1 merges lanes (0,1) and (2,3) as in select_pair; round 2 takes min / max against the reversed
// sequence of the other pair (partner lane q ^ 3), then one cross-lane compare-exchange stage inside each pair
// (partner q ^ 1, same register) and a bitonic merge per lane.  Lane q ends with ranks q S .. q S + S - 1.
template <int S>
static int check_quad(std::mt19937& rng, int samples) {
    int bad = 0;
    for (int s = 0; s < samples; ++s) {
        const int c = (int)(rng() % (unsigned)(4 * S + 1));
        unsigned keys[4 * S], ref[4 * S];
        for (int j = 0; j < 4 * S; ++j) keys[j] = j < c ? (((s & 1) ? (rng() % 50u) : (rng() >> 8)) << 7 | (unsigned)j) : 0xffffffffu;
        for (int j = 0; j < 4 * S; ++j) ref[j] = keys[j];
        std::sort(ref, ref + 4 * S);
        unsigned L[4][S], T[4][S];
        for (int q = 0; q < 4; ++q) {
            for (int i = 0; i < S; ++i) L[q][i] = keys[4 * i + q];
            sort_network<S>(L[q]);
        }
        auto reversed_exchange = [&](int xor_mask, int max_bit) {          // lockstep: read all, then write all
            for (int q = 0; q < 4; ++q)
                for (int x = 0; x < S; ++x) {
                    const unsigned other = L[q ^ xor_mask][S - 1 - x];
                    T[q][x] = (q & max_bit) ? net_max(L[q][x], other) : net_min(L[q][x], other);
                }
            for (int q = 0; q < 4; ++q)
                for (int x = 0; x < S; ++x) L[q][x] = T[q][x];
        };
        reversed_exchange(1, 1);
        for (int q = 0; q < 4; ++q) bitonic_merge<S>(L[q]);
        reversed_exchange(3, 2);
        for (int q = 0; q < 4; ++q)
            for (int x = 0; x < S; ++x) {
                const unsigned other = L[q ^ 1][x];
                T[q][x] = (q & 1) ? net_max(L[q][x], other) : net_min(L[q][x], other);
            }
        for (int q = 0; q < 4; ++q) {
            for (int x = 0; x < S; ++x) L[q][x] = T[q][x];
            bitonic_merge<S>(L[q]);
        }
        for (int q = 0; q < 4; ++q)
            for (int i = 0; i < S; ++i) bad += L[q][i] != ref[q * S + i];
    }
    return bad;
}

int main() {
    std::mt19937 rng(1234);
    int bad = 0;
    {   // exhaustive zero-one principle for N = 16
        unsigned r[16];
        for (unsigned m = 0; m < 65536u; ++m) {
            for (int i = 0; i < 16; ++i) r[i] = (m >> i) & 1u;
            sort_network<16>(r);
            bad += !is_sorted_arr<16>(r);
        }
    }
    bad += check_sort<16>(rng, 3000) + check_sort<32>(rng, 30000) + check_sort<48>(rng, 10000) + check_sort<64>(rng, 10000);
    bad += check_sort<20>(rng, 3000) + check_sort<7>(rng, 3000);
    bad += check_bitonic<8>() + check_bitonic<16>() + check_bitonic<32>();
    bad += check_pair<8>(rng, 20000) + check_pair<16>(rng, 20000) + check_pair<32>(rng, 20000);
    bad += check_quad<8>(rng, 10000) + check_quad<16>(rng, 10000) + check_quad<32>(rng, 10000);
    static_assert(odd_even_count(16) == 63 && odd_even_count(32) == 191 && odd_even_count(64) == 543, "comparator counts");
    {   // fold_index
        const int ds[] = {1, 2, 3, 7, 1000, 35947, 1000000, 999983, 1 << 20, (1 << 30) + 7, 2147483647};
        for (int d : ds) {
            const FastMod f = make_fastmod(d);
            for (int t = 0; t < 200000; ++t) {
                const int g = t < 64 ? (t < 32 ? t : 2147483647 - (t - 32)) : (int)(rng() >> 1);
                bad += fold_index(g, f) != g % d;
            }
            bad += fold_index(-1, f) != 0 || fold_index(-2147483647, f) != 0;
        }
        const FastMod id = make_fastmod(0);
        bad += fold_index(12345, id) != 12345 || fold_index(-5, id) != 0;
    }
    std::printf("bad=%d\n", bad);
    return bad != 0;
}

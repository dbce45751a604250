// Micro-benchmarks behind DESIGN.md's "what a hit costs" table: per-SM cost, in cycles per WARP instruction, of the
// memory operations the fused kernels are built from, measured on the box the bench runs on.  Every kernel runs
// 148 x 8 CTAs of 256 threads; each thread repeats the operation ITERS times on addresses from a cheap hash.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_costs l1_costs.cu && ./l1_costs
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

constexpr int ITERS = 256;
constexpr int NT = 256;

__device__ __forceinline__ unsigned mix(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

enum Mode { LDS32_LINEAR, LDS32_RANDOM, LDS128_LINEAR, LDS128_RANDOM, ATOMS_ADD_LINEAR, ATOMS_ADD_RANDOM, ATOMS_ADD_SAME4,
            ATOMS_CAS_RANDOM, LDG128_SCATTER, LDG128_SCATTER3, LDG128_PAIRS, RED128_SCATTER, RED32_SCATTER, MATCH_ANY, SHFL, STS64_RANDOM,
            LDG32_SCATTER, RED128_SCATTER3, LDG128_SCATTER4_64B, ATOMS_CAS_INSERT, NMODES };
const char* kNames[NMODES] = {"LDS.32 conflict-free", "LDS.32 random bank", "LDS.128 linear", "LDS.128 random",
                              "ATOMS.ADD.S32 distinct banks", "ATOMS.ADD.S32 random addr (4K table)", "ATOMS.ADD.S32 4 lanes/addr",
                              "ATOMS.CAS random addr (4K table)", "LDG.128 one random line per lane (48 MB table)",
                              "3 x LDG.128 consecutive 48 B record per lane", "LDG.128 lane pairs share a line",
                              "RED.ADD.F32x4 one random line per lane", "RED.ADD.F32 one random line per lane",
                              "MATCH.ANY", "SHFL.BFLY", "STS.64 random", "LDG.32 one random line per lane",
                              "3 x RED.ADD.F32x4 consecutive 48 B record per lane", "4 x LDG.128 one 64 B record per lane",
                              "hash insert: ATOMS.CAS probe + ATOMS.ADD (512 slots, ~100 keys)"};

template <int MODE>
__global__ void __launch_bounds__(NT) bench_kernel(float4* table, unsigned table_mask, float* sink) {
    __shared__ __align__(16) unsigned s[4096 + 8];
    const int tid = threadIdx.x;
    for (int i = tid; i < 4096; i += NT) s[i] = i;
    __syncthreads();
    unsigned h = mix(blockIdx.x * NT + tid + 1);
    float acc = 0.f;
    unsigned uacc = 0;
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        h = h * 1664525u + 1013904223u;
        const unsigned r = h >> 8;
        if (MODE == LDS32_LINEAR) uacc += s[(tid + it * 32) & 4095];
        if (MODE == LDS32_RANDOM) uacc += s[r & 4095];
        if (MODE == LDS128_LINEAR) { const uint4 v = *reinterpret_cast<const uint4*>(&s[((tid + it * 32) & 1023) * 4]); uacc += v.x + v.w; }
        if (MODE == LDS128_RANDOM) { const uint4 v = *reinterpret_cast<const uint4*>(&s[(r & 1023) * 4]); uacc += v.x + v.w; }
        if (MODE == ATOMS_ADD_LINEAR) uacc += atomicAdd(&s[(tid + it * 32) & 4095], 1u);
        if (MODE == ATOMS_ADD_RANDOM) uacc += atomicAdd(&s[r & 4095], 1u);
        if (MODE == ATOMS_ADD_SAME4) uacc += atomicAdd(&s[((tid >> 2) + it * 8) & 4095], 1u);
        if (MODE == ATOMS_CAS_RANDOM) uacc += atomicCAS(&s[r & 4095], r & 4095, r & 4095);
        if (MODE == STS64_RANDOM) *reinterpret_cast<uint2*>(&s[(r & 2047) * 2]) = make_uint2(r, uacc);
        if (MODE == LDG128_SCATTER) { const float4 v = __ldg(table + (size_t)(r & table_mask) * 8); acc += v.x + v.w; }
        if (MODE == LDG32_SCATTER) { acc += __ldg(reinterpret_cast<const float*>(table + (size_t)(r & table_mask) * 8)); }
        if (MODE == LDG128_SCATTER3) {
            const float4* p = table + (size_t)(r & table_mask) * 3;
            const float4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
            acc += v0.x + v1.y + v2.z;
        }
        if (MODE == LDG128_PAIRS) {
            const unsigned rr = __shfl_sync(0xffffffffu, r, tid & 30);
            const float4 v = __ldg(table + (size_t)(rr & table_mask) * 8 + (tid & 1));
            acc += v.x + v.w;
        }
        if (MODE == RED128_SCATTER) atomicAdd(table + (size_t)(r & table_mask) * 8, make_float4(1.f, 2.f, 3.f, 4.f));
        if (MODE == RED32_SCATTER) atomicAdd(reinterpret_cast<float*>(table + (size_t)(r & table_mask) * 8), 1.f);
        if (MODE == RED128_SCATTER3) {
            float4* p = table + (size_t)(r & table_mask) * 3;
            atomicAdd(p, make_float4(1.f, 2.f, 3.f, 4.f)); atomicAdd(p + 1, make_float4(1.f, 2.f, 3.f, 4.f));
            atomicAdd(p + 2, make_float4(1.f, 2.f, 3.f, 4.f));
        }
        if (MODE == LDG128_SCATTER4_64B) {
            const float4* p = table + (size_t)(r & table_mask) * 4;
            const float4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
            acc += v0.x + v1.y + v2.z + v3.w;
        }
        if (MODE == ATOMS_CAS_INSERT) {
            const unsigned key = 4096u + (r % 97u);              // ~100 distinct keys per CTA, as a tile sees Gaussians
            unsigned hsl = (key * 2654435761u) >> 23;            // 512 slots
            for (int probe = 0; probe < 16; ++probe) {
                const unsigned old = atomicCAS(&s[hsl], hsl, key);   // slot initially holds its own index = "empty"
                if (old == hsl || old == key) break;
                hsl = (hsl + 1) & 511;
            }
            uacc += atomicAdd(&s[512 + hsl], 1u);
        }
        if (MODE == MATCH_ANY) uacc += __match_any_sync(0xffffffffu, r & 15);
        if (MODE == SHFL) { acc += __shfl_xor_sync(0xffffffffu, acc + (float)it, 5); }
    }
    if (acc + (float)uacc == 123.456f) sink[0] = acc;
    __syncthreads();
    if (s[tid] == 0xdeadbeefu) sink[1] = 1.f;
}

template <int MODE>
double run(float4* table, unsigned mask, float* sink, int clock_mhz) {
    const int blocks = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench_kernel<MODE><<<blocks, NT>>>(table, mask, sink);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        bench_kernel<MODE><<<blocks, NT>>>(table, mask, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double warp_insts_per_sm = (double)blocks * (NT / 32) * ITERS / 148.0;
    const double cycles = best * 1e-3 * clock_mhz * 1e6;
    return cycles / warp_insts_per_sm;
}

int main() {
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    const int mhz = clock_khz / 1000;
    const size_t lines = 1u << 19;                 // 512 Ki entries x 128 B = 64 MB (L2-resident like the Gaussian tables)
    float4* table;
    float* sink;
    cudaMalloc(&table, lines * 128);
    cudaMemset(table, 0, lines * 128);
    cudaMalloc(&sink, 16);
    const unsigned mask = (unsigned)lines - 1;
    printf("SM clock %d MHz (max); cycles per warp instruction per SM (all 32 lanes active, 64 warps/SM resident)\n", mhz);
    double v[NMODES];
    v[LDS32_LINEAR] = run<LDS32_LINEAR>(table, mask, sink, mhz);
    v[LDS32_RANDOM] = run<LDS32_RANDOM>(table, mask, sink, mhz);
    v[LDS128_LINEAR] = run<LDS128_LINEAR>(table, mask, sink, mhz);
    v[LDS128_RANDOM] = run<LDS128_RANDOM>(table, mask, sink, mhz);
    v[ATOMS_ADD_LINEAR] = run<ATOMS_ADD_LINEAR>(table, mask, sink, mhz);
    v[ATOMS_ADD_RANDOM] = run<ATOMS_ADD_RANDOM>(table, mask, sink, mhz);
    v[ATOMS_ADD_SAME4] = run<ATOMS_ADD_SAME4>(table, mask, sink, mhz);
    v[ATOMS_CAS_RANDOM] = run<ATOMS_CAS_RANDOM>(table, mask, sink, mhz);
    v[STS64_RANDOM] = run<STS64_RANDOM>(table, mask, sink, mhz);
    v[LDG128_SCATTER] = run<LDG128_SCATTER>(table, mask, sink, mhz);
    v[LDG32_SCATTER] = run<LDG32_SCATTER>(table, mask, sink, mhz);
    v[LDG128_SCATTER3] = run<LDG128_SCATTER3>(table, mask, sink, mhz);
    v[LDG128_PAIRS] = run<LDG128_PAIRS>(table, mask, sink, mhz);
    v[RED128_SCATTER] = run<RED128_SCATTER>(table, mask, sink, mhz);
    v[RED32_SCATTER] = run<RED32_SCATTER>(table, mask, sink, mhz);
    v[MATCH_ANY] = run<MATCH_ANY>(table, mask, sink, mhz);
    v[SHFL] = run<SHFL>(table, mask, sink, mhz);
    v[RED128_SCATTER3] = run<RED128_SCATTER3>(table, mask, sink, mhz);
    v[LDG128_SCATTER4_64B] = run<LDG128_SCATTER4_64B>(table, mask, sink, mhz);
    v[ATOMS_CAS_INSERT] = run<ATOMS_CAS_INSERT>(table, mask, sink, mhz);
    for (int m = 0; m < NMODES; ++m) printf("%-52s %8.2f cycles / warp inst  (%.3f per lane)\n", kNames[m], v[m], v[m] / 32.0);
    return 0;
}

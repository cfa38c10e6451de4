// Evidence for DESIGN.md 3: what it costs to deliver ONE 16-byte node piece to all 32 lanes of a warp.
//   A  LDG.128, all lanes the same address (the hot kernel's gather: a warp sits in one grid cell), L1 resident
//   B  LDS.128, all lanes the same address (what a TMA / cp.async staged tile would be read with)
//   C  LDG.128, every lane its own 16 bytes (coalesced 512 B) - the cost if the data were NOT shared
// 16 warps per SM like the hot kernel, 24 loads per round (= one right-hand side), results folded so that nothing
// is optimised away.  Prints cycles per warp-load per SM.  Build + run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/microbench_broadcast profiles/microbench_broadcast.cu
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int kLoads = 24, kRounds = 4096, kThreads = 128, kBlocksPerSM = 4;

template <int MODE>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM) bench(const double2* __restrict__ g, double* out,
                                                                long long* cycles) {
    __shared__ double2 tile[kLoads * 64];
    for (int i = threadIdx.x; i < kLoads * 64; i += kThreads) tile[i] = g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc0 = 0., acc1 = 0.;
    const long long t0 = clock64();
    for (int r = 0; r < kRounds; ++r) {
        const int base = ((r * 7 + warp * 13 + blockIdx.x) & 31) * kLoads;  // a different "cell" every round
#pragma unroll
        for (int k = 0; k < kLoads; ++k) {
            double2 v;
            if (MODE == 0) {
                const double2* p = g + base + k;
                asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
            } else if (MODE == 1) {
                const unsigned a = (unsigned) __cvta_generic_to_shared(tile + base + k);
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
            } else {
                const double2* p = g + (base + k) * 32 + lane;
                asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
            }
            acc0 += v.x;
            acc1 += v.y;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * kThreads + threadIdx.x] = acc0 + acc1;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
double run(const double2* g, double* out, long long* cycles, int blocks, int sms) {
    bench<MODE><<<blocks, kThreads>>>(g, out, cycles);  // warm-up
    bench<MODE><<<blocks, kThreads>>>(g, out, cycles);
    cudaDeviceSynchronize();
    long long* h = new long long[blocks];
    cudaMemcpy(h, cycles, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double mean = 0.;
    for (int i = 0; i < blocks; ++i) mean += (double) h[i] / blocks;
    delete[] h;
    // one SM runs kBlocksPerSM blocks = 16 warps concurrently; each issues kRounds * kLoads warp-loads in `mean` cycles
    const double warp_loads_per_sm = (double) kBlocksPerSM * (kThreads / 32) * kRounds * kLoads;
    return mean / warp_loads_per_sm;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, blocks = sms * kBlocksPerSM;
    const size_t n = (size_t) 32 * kLoads * 32 + kLoads * 64;
    double2* g;
    double* out;
    long long* cycles;
    cudaMalloc(&g, sizeof(double2) * n);
    cudaMemset(g, 0, sizeof(double2) * n);
    cudaMalloc(&out, sizeof(double) * blocks * kThreads);
    cudaMalloc(&cycles, sizeof(long long) * blocks);
    const double a = run<0>(g, out, cycles, blocks, sms);
    const double b = run<1>(g, out, cycles, blocks, sms);
    const double c = run<2>(g, out, cycles, blocks, sms);
    printf("%s, %d SMs, 16 warps/SM, cycles per warp-level 16-byte load per SM:\n", prop.name, sms);
    printf("  A  LDG.128 broadcast (same address in all lanes, L1 hit): %.2f\n", a);
    printf("  B  LDS.128 broadcast (same address in all lanes)        : %.2f\n", b);
    printf("  C  LDG.128 coalesced (32 different addresses, L1 hit)    : %.2f\n", c);
    printf("  one right-hand side = 24 such loads per warp: %.0f / %.0f / %.0f cycles of the SM's L1 data stage\n",
           24 * a, 24 * b, 24 * c);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// Round 2 companion of microbench_broadcast.cu: what ELSE could deliver a grid cell (24 sixteen-byte pieces = 384 B) to
// the 32 lanes of a warp, and what lanes in different cells cost.  Same frame: 16 warps per SM, cycles of the SM per
// 16 bytes delivered to every lane of a warp.
//   A   LDG.128, same address in all lanes (the hot kernel today)
//   D   LDG.256 (ld.global.nc.v4.f64), same address: 12 loads per cell instead of 24
//   E   LDG.64, same address: 48 loads per cell
//   F2  LDG.128, lanes 0-15 in one cell and 16-31 in another          (two cells, split by half warp)
//   F4  LDG.128, one cell per quarter warp                             (four cells)
//   I2  LDG.128, even lanes in one cell and odd lanes in another       (two cells, interleaved)
//   R   LDG.128, 32 lanes in 32 different cells                        (fully divergent)
//   S   SHFL.IDX of one 32-bit register (what redistributing a cooperatively loaded cell would cost, per 4 bytes)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/microbench_broadcast2 profiles/microbench_broadcast2.cu
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int kPieces = 24, kRounds = 2048, kThreads = 128, kBlocksPerSM = 4, kCells = 64;

template <int MODE>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM) bench(const double2* __restrict__ g, double* out,
                                                                long long* cycles) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc0 = 0., acc1 = 0.;
    unsigned sh = lane;
    const long long t0 = clock64();
    for (int r = 0; r < kRounds; ++r) {
        int cell = (r * 7 + warp * 13 + blockIdx.x) & (kCells - 1);
        if (MODE == 3) cell = (cell + (lane >> 4)) & (kCells - 1);
        if (MODE == 4) cell = (cell + (lane >> 3)) & (kCells - 1);
        if (MODE == 5) cell = (cell + (lane & 1)) & (kCells - 1);
        if (MODE == 6) cell = (cell + lane) & (kCells - 1);
        const double2* base = g + cell * kPieces;
        if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < kPieces / 2; ++k) {
                double a, b, c, d;
                asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(base + 2 * k));
                acc0 += a + c;
                acc1 += b + d;
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < kPieces * 2; ++k) {
                double a;
                asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(a) : "l"(reinterpret_cast<const double*>(base) + k));
                acc0 += a;
            }
        } else if (MODE == 7) {
#pragma unroll
            for (int k = 0; k < kPieces * 4; ++k) {
                sh = __shfl_sync(0xffffffffu, sh + k, (lane + k) & 31);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kPieces; ++k) {
                double2 v;
                asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(base + k));
                acc0 += v.x;
                acc1 += v.y;
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * kThreads + threadIdx.x] = acc0 + acc1 + (double) sh;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
double run(const double2* g, double* out, long long* cycles, int blocks) {
    bench<MODE><<<blocks, kThreads>>>(g, out, cycles);  // warm-up
    bench<MODE><<<blocks, kThreads>>>(g, out, cycles);
    cudaDeviceSynchronize();
    long long* h = new long long[blocks];
    cudaMemcpy(h, cycles, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double mean = 0.;
    for (int i = 0; i < blocks; ++i) mean += (double) h[i] / blocks;
    delete[] h;
    // per SM: 16 warps, each delivering kRounds * kPieces sixteen-byte pieces to its lanes in `mean` cycles
    return mean / ((double) kBlocksPerSM * (kThreads / 32) * kRounds * kPieces);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int blocks = prop.multiProcessorCount * kBlocksPerSM;
    const size_t n = (size_t) (kCells + 32) * kPieces;
    double2* g;
    double* out;
    long long* cycles;
    cudaMalloc(&g, sizeof(double2) * n);
    cudaMemset(g, 0, sizeof(double2) * n);
    cudaMalloc(&out, sizeof(double) * blocks * kThreads);
    cudaMalloc(&cycles, sizeof(long long) * blocks);
    printf("%s, 16 warps/SM, SM cycles per 16 bytes delivered to every lane of a warp (one cell = 24 of them):\n", prop.name);
    printf("  A   LDG.128 same address          : %.2f\n", run<0>(g, out, cycles, blocks));
    printf("  D   LDG.256 same address          : %.2f\n", run<1>(g, out, cycles, blocks));
    printf("  E   LDG.64  same address          : %.2f\n", run<2>(g, out, cycles, blocks));
    printf("  F2  LDG.128 two cells by half warp : %.2f\n", run<3>(g, out, cycles, blocks));
    printf("  F4  LDG.128 four cells by quarter  : %.2f\n", run<4>(g, out, cycles, blocks));
    printf("  I2  LDG.128 two cells interleaved  : %.2f\n", run<5>(g, out, cycles, blocks));
    printf("  R   LDG.128 32 cells               : %.2f\n", run<6>(g, out, cycles, blocks));
    printf("  S   SHFL.IDX, 4 per 16 bytes       : %.2f (per 16 bytes; %.2f per shuffle)\n", run<7>(g, out, cycles, blocks),
           run<7>(g, out, cycles, blocks) / 4);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// Evidence for the roofline of the hot kernel (bench.py `roofline.bound = "fp64"`): what the fp64 pipe
// of a B200 SM sustains, measured - MEASURED_PEAKS.json only holds the HBM copy rate and the bf16 tensor rate.
//   A  peak: every thread runs 8 independent DFMA chains (no memory traffic), 4 / 8 / 16 warps per scheduler;
//   B  the hot kernel's occupancy: 4 warps per scheduler (16 per SM), 1 .. 8 independent chains per thread - how much
//      instruction-level parallelism a thread needs before the pipe saturates at that occupancy;
//   C  latency of one dependent DFMA (1 warp per scheduler, 1 chain).
// Prints fp64 warp-instructions per cycle per SM and the equivalent TFLOP/s (one DFMA = 2 flop) at the SM clock the
// run saw, and writes the same as one JSON line (stdout, last line).  Build + run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench_fp64 profiles/microbench_fp64.cu && /tmp/microbench_fp64
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int kIters = 1 << 14;

template <int CHAINS>
__global__ void dfma_kernel(double* out, long long* cycles, double a, double b) {
    double v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = (double) (threadIdx.x + c);
    const long long t0 = clock64();
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) v[c] = fma(v[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0.;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += v[c];
    out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Result {
    double inst_per_clk_sm, ms;
};

template <int CHAINS>
Result run(int blocks_per_sm, int threads, int sms, double* out, long long* cycles) {
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma_kernel<CHAINS><<<blocks, threads>>>(out, cycles, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    dfma_kernel<CHAINS><<<blocks, threads>>>(out, cycles, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[blocks];
    cudaMemcpy(h, cycles, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double mean = 0.;
    for (int i = 0; i < blocks; ++i) mean += (double) h[i] / blocks;
    delete[] h;
    const double warp_inst_per_sm = (double) blocks_per_sm * (threads / 32) * kIters * CHAINS;
    return Result{warp_inst_per_sm / mean, ms};
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double* out;
    long long* cycles;
    cudaMalloc(&out, sizeof(double) * sms * 16 * 1024);
    cudaMalloc(&cycles, sizeof(long long) * sms * 16);
    printf("%s, %d SMs\n", prop.name, sms);
    printf("A  peak, 8 chains per thread:\n");
    double peak = 0., peak_ms = 0.;
    int peak_warps = 0;
    for (int bps : {4, 8, 16}) {  // blocks of 128 threads per SM -> 16 / 32 / 64 warps per SM
        const Result r = run<8>(bps, 128, sms, out, cycles);
        printf("   %2d warps/SM: %.3f fp64 warp-inst/clk/SM (%.1f lanes/clk/SM), %.3f ms\n", bps * 4, r.inst_per_clk_sm,
               32 * r.inst_per_clk_sm, r.ms);
        if (r.inst_per_clk_sm > peak) {
            peak = r.inst_per_clk_sm;
            peak_ms = r.ms;
            peak_warps = bps * 4;
        }
    }
    // wall-clock rate of the best configuration -> TFLOP/s at the clock the run actually had
    const double total_inst = (double) sms * peak_warps * kIters * 8;  // warp instructions
    const double tflops = total_inst * 32 * 2 / (peak_ms * 1e-3) / 1e12;
    const double clk_ghz = total_inst / peak / sms / (peak_ms * 1e-3) / 1e9;
    printf("   -> %.2f TFLOP/s fp64 (DFMA = 2 flop) at %.3f GHz effective SM clock\n", tflops, clk_ghz);
    printf("B  16 warps/SM (the hot kernel's occupancy), chains per thread:\n");
    const Result b1 = run<1>(4, 128, sms, out, cycles), b2 = run<2>(4, 128, sms, out, cycles);
    const Result b4 = run<4>(4, 128, sms, out, cycles), b8 = run<8>(4, 128, sms, out, cycles);
    printf("   1: %.3f   2: %.3f   4: %.3f   8: %.3f warp-inst/clk/SM\n", b1.inst_per_clk_sm, b2.inst_per_clk_sm,
           b4.inst_per_clk_sm, b8.inst_per_clk_sm);
    // D  sustained: the peak configuration back to back for 3 s (clocks ramp up from idle, then settle under the power
    //    limit); the rate of the last second is what a long fp64-bound kernel can count on
    double sustained = 0., sustained_clk = 0., burst_best = tflops;
    {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        const int blocks = sms * (peak_warps / 4);
        const double flop_per_launch = (double) sms * peak_warps * kIters * 8 * 32 * 2;
        double elapsed = 0.;
        while (elapsed < 3000.) {
            cudaEventRecord(e0);
            for (int i = 0; i < 20; ++i) dfma_kernel<8><<<blocks, 128>>>(out, cycles, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            elapsed += ms;
            const double rate = 20 * flop_per_launch / (ms * 1e-3) / 1e12;
            if (rate > burst_best) burst_best = rate;
            if (elapsed >= 2000.) {
                sustained = rate;  // last batches
                long long* h = new long long[blocks];
                cudaMemcpy(h, cycles, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
                double mean = 0.;
                for (int i = 0; i < blocks; ++i) mean += (double) h[i] / blocks;
                delete[] h;
                sustained_clk = mean / (ms / 20 * 1e-3) / 1e9;
            }
        }
        printf("D  sustained (after 2 s of back-to-back launches): %.2f TFLOP/s at %.3f GHz; best batch %.2f TFLOP/s\n", sustained,
               sustained_clk, burst_best);
    }
    const Result c = run<1>(1, 128, sms, out, cycles);
    const double latency = 4. / c.inst_per_clk_sm;  // 4 warps on 4 schedulers, each one instruction per `latency` cycles
    printf("C  dependent DFMA latency: %.1f cycles\n", latency);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_warp_inst_per_clk_sm\": %.4f, \"fp64_tflops_cold\": %.3f, "
           "\"sm_clock_ghz_cold\": %.4f, \"fp64_tflops_sustained\": %.3f, \"sm_clock_ghz_sustained\": %.4f, "
           "\"fp64_tflops_best_batch\": %.3f, \"at_16_warps_chains_1_2_4_8\": [%.4f, %.4f, %.4f, %.4f], "
           "\"dfma_latency_cycles\": %.2f}\n",
           prop.name, sms, peak, tflops, clk_ghz, sustained, sustained_clk, burst_best, b1.inst_per_clk_sm,
           b2.inst_per_clk_sm, b4.inst_per_clk_sm, b8.inst_per_clk_sm, latency);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

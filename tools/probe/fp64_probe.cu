// fp64_probe.cu -- latency (one dependent chain, one warp) and throughput (many warps, independent chains) of the FP64
// instructions the reduced solve leans on, on sm_100a: DFMA, DMUL, rsqrt(double), 64-bit SHFL, DMMA m8n8k4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// OP: 0 DFMA chain, 1 DMUL chain, 2 rsqrt chain, 3 shfl64 chain, 4 DMMA chain (accumulator dependent), 5 shfl + dmul + dfma (trsm step)
template <int OP, int CHAINS>
__global__ void k(double* out, double seed, int iters, long long* cycles) {
    double x[CHAINS], y[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { x[i] = seed + threadIdx.x * 1e-3 + i; y[i] = 0.5; }
    const double b = seed * 0.999, c = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                if (OP == 0) x[i] = fma(x[i], b, c);
                if (OP == 1) x[i] = x[i] * b;
                if (OP == 2) x[i] = rsqrt(x[i]) + 1.5;
                if (OP == 3) x[i] = __shfl_sync(0xFFFFFFFFu, x[i], (r + i) & 31);
                if (OP == 4) dmma884(x[i], y[i], b, c);
                if (OP == 5) { const double a = __shfl_sync(0xFFFFFFFFu, x[i], r) * b; x[i] = fma(-a, c, x[i]); }
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP, int CHAINS>
void run(const char* name, int threads, int blocks) {
    double* out; long long* cyc; long long h = 0;
    cudaMalloc(&out, sizeof(double) * threads * blocks); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<OP, CHAINS><<<blocks, threads>>>(out, 1.37, 10, cyc);
    k<OP, CHAINS><<<blocks, threads>>>(out, 1.37, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * 8.0);
    printf("%-34s threads %4d chains %d: %7.2f cycles per round (%.2f per op per warp-chain; %.3f warp-ops/clk/SM)\n", name, threads, CHAINS, per, per / CHAINS,
           (threads / 32.0) * CHAINS / per);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    // latency: one warp, one chain
    run<0, 1>("DFMA latency", 32, 1);
    run<1, 1>("DMUL latency", 32, 1);
    run<2, 1>("rsqrt(double)+add latency", 32, 1);
    run<3, 1>("SHFL 64-bit latency", 32, 1);
    run<4, 1>("DMMA m8n8k4 latency (acc chain)", 32, 1);
    run<5, 1>("shfl+dmul+dfma (trsm step)", 32, 1);
    run<5, 4>("trsm step, 4 chains / warp", 32, 1);
    run<5, 4>("trsm step, 4 chains, 8 warps", 256, 1);
    // throughput: one SM, many warps
    run<0, 8>("DFMA throughput", 256, 1);
    run<0, 8>("DFMA throughput", 1024, 1);
    run<4, 4>("DMMA throughput", 128, 1);
    run<4, 4>("DMMA throughput", 256, 1);
    run<4, 8>("DMMA throughput", 256, 1);
    run<4, 4>("DMMA throughput", 512, 1);
    run<4, 4>("DMMA throughput", 1024, 1);
    run<3, 8>("SHFL64 throughput", 256, 1);
    run<4, 4>("DMMA throughput, 148 CTAs", 256, 148);
    return 0;
}

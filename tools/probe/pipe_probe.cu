// pipe_probe.cu -- issue rate of the integer ops the FAST / descriptor kernels lean on (sm_100a).
// Each kernel runs a long dependent-free stream of one op on 8 independent chains per thread;
// reports warp-instructions per clock per SM.   nvcc -arch=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i + 1);
    uint32_t b = seed ^ 0x1234567u, c = seed * 77u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) a[i] = __vimin3_u16x2(a[i], b, c);
                if (OP == 1) a[i] = __vminu2(a[i], b);
                if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
                if (OP == 3) a[i] = a[i] * b + c;                                  // IMAD
                if (OP == 4) a[i] = __byte_perm(a[i], b, c);                        // PRMT
                if (OP == 5) a[i] = __funnelshift_r(a[i], b, c);                    // SHF
                if (OP == 6) a[i] = (uint32_t)min((int)a[i], (int)b);               // VIMNMX scalar
                if (OP == 7) a[i] = __dp2a_lo(a[i], b, c);
                if (OP == 8) a[i] = __dp4a(a[i], b, c);
                if (OP == 9) a[i] = (uint32_t)max(min((int)a[i], (int)b), (int)c);  // VIMNMX3 scalar
                if (OP == 11) { __half2 h = __hmin2(__hmin2(*(__half2*)&a[i], *(__half2*)&b), *(__half2*)&c); a[i] = *(uint32_t*)&h; }   // VHMNMX
                if (OP == 12) { __half2 h = __hmin2(*(__half2*)&a[i], *(__half2*)&b); a[i] = *(uint32_t*)&h; }                             // HMNMX2
                if (OP == 13) {   // VIMNMX3.U16x2 + VHMNMX alternating on independent chains
                    if (i & 1) a[i] = __vimin3_u16x2(a[i], b, c);
                    else { __half2 h = __hmin2(__hmin2(*(__half2*)&a[i], *(__half2*)&b), *(__half2*)&c); a[i] = *(uint32_t*)&h; }
                }
                if (OP == 14) {   // VIMNMX3.U16x2 + HMNMX2
                    if (i & 1) a[i] = __vimin3_u16x2(a[i], b, c);
                    else { __half2 h = __hmin2(*(__half2*)&a[i], *(__half2*)&b); a[i] = *(uint32_t*)&h; }
                }
                if (OP == 15) { if (i & 1) a[i] = __vimin3_u16x2(a[i], b, c); else a[i] = a[i] * b + c; }   // VIMNMX3 + IMAD
                if (OP == 16) { float f = fminf(fminf(__uint_as_float(a[i]), __uint_as_float(b)), __uint_as_float(c)); a[i] = __float_as_uint(f); }   // FMNMX3
                if (OP == 17) { if (i & 1) a[i] = __vimin3_u16x2(a[i], b, c); else a[i] = __vminu2(a[i], b); }   // VIMNMX3 + VIMNMX
                if (OP == 10) { a[i] = a[i] * b + c; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[(i + 4) & 7]) : "r"(b), "r"(c)); }
            }
            b += c;
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, double ops_per_inner) {
    uint32_t* d; cudaMalloc(&d, 148 * 2 * 1024 * 4);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * 2, 1024>>>(d, 3u, 16);
    cudaEventRecord(e0);
    k<OP><<<148 * 2, 1024>>>(d, 3u, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double winstr = 148.0 * 2 * 32 * iters * 4 * 8 * ops_per_inner;
    printf("%-22s %8.3f ms  %.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, winstr / (ms * 1e-3) / (clk * 1e3) / 148.0, clk / 1000);
    cudaFree(d);
}

int main() {
    run<0>("VIMNMX3.U16x2", 1); run<1>("VIMNMX.U16x2", 1); run<2>("LOP3", 1); run<3>("IMAD", 1); run<4>("PRMT", 1);
    run<5>("SHF", 1); run<6>("VIMNMX.S32", 1); run<7>("IDP2A", 1); run<8>("IDP4A", 1); run<9>("VIMNMX3.S32(min,max)", 1);
    run<10>("IMAD+LOP3 mix", 2);
    run<11>("VHMNMX (half2, 3 in)", 1); run<12>("HMNMX2", 1); run<13>("VIMNMX3+VHMNMX mix", 1); run<14>("VIMNMX3+HMNMX2 mix", 1);
    run<15>("VIMNMX3+IMAD mix", 1); run<16>("FMNMX3", 1); run<17>("VIMNMX3+VIMNMX mix", 1);
    return 0;
}

#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k2d(const __grid_constant__ CUtensorMap map, uint8_t* out, int x, int y, int bytes) {
    __shared__ __align__(128) uint8_t tile[8192];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(tile)), "l"(&map), "r"(x), "r"(y), "r"(s32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
__global__ void k3d(const __grid_constant__ CUtensorMap map, uint8_t* out, int x, int y, int z, int bytes, int fence) {
    __shared__ __align__(128) uint8_t tile[8192];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        if (fence) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(tile)), "l"(&map), "r"(x), "r"(y), "r"(z), "r"(s32(&bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
    int dtype = argc > 1 ? atoi(argv[1]) : 0;   // 0 u8, 1 f32
    const int w = 640, h = 480;
    uint8_t *d, *out; cudaMalloc(&d, (size_t)w * h * 4); cudaMalloc(&out, 8192);
    cudaMemset(d, 7, (size_t)w * h * 4);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ee = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry %p err %d q %d\n", p, (int)ee, (int)q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)p;
    CUtensorMap map; memset(&map, 0, sizeof(map));
    cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
    int es = dtype ? 4 : 1;
    cuuint64_t strides[1] = {(cuuint64_t)w * es};
    cuuint32_t box[2] = {(cuuint32_t)(dtype ? 16 : 64), 16};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map, dtype ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d; desc:", (int)r);
    for (int i = 0; i < 16; ++i) printf(" %016llx", (unsigned long long)map.opaque[i]);
    printf("\n");
    int rank = argc > 2 ? atoi(argv[2]) : 2, promo = argc > 3 ? atoi(argv[3]) : 0, bw = argc > 4 ? atoi(argv[4]) : 64, fence = argc > 5 ? atoi(argv[5]) : 0;
    if (rank == 3) {
        cuuint64_t d3[3] = {(cuuint64_t)w, (cuuint64_t)h / 2, 2};
        cuuint64_t s3[2] = {(cuuint64_t)w, (cuuint64_t)w * h / 2};
        cuuint32_t b3[3] = {(cuuint32_t)bw, 16, 1};
        cuuint32_t e3[3] = {1, 1, 1};
        r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("rank3 promo %d bw %d fence %d encode %d\n", promo, bw, fence, (int)r);
        { int cx = argc > 6 ? atoi(argv[6]) : 3, cy = argc > 7 ? atoi(argv[7]) : 5, cz = argc > 8 ? atoi(argv[8]) : 1; k3d<<<1, 128>>>(map, out, cx, cy, cz, bw * 16, fence); }
    } else
    { int cx = argc > 6 ? atoi(argv[6]) : 0, cy = argc > 7 ? atoi(argv[7]) : 0; k2d<<<1, 128>>>(map, out, cx, cy, 64 * 16); }
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    return 0;
}

// Probe which TMA descriptor / addressing variants work on this box.  usage: tma_probe <variant>
#include "../../airdos_b200/csrc/common.cuh"
#include <cstdlib>
#include <vector>
using namespace adb;
struct Maps { CUtensorMap m[16]; };
__global__ void k_single(const __grid_constant__ CUtensorMap map, uint8_t* out, int x, int y, int z, int bytes) {
    __shared__ __align__(128) uint8_t tile[8192];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); fence_proxy_async(); mbar_expect_tx(&bar, bytes); tma_load_3d(tile, &map, &bar, x, y, z); }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
__global__ void k_array(const __grid_constant__ Maps maps, int idx, uint8_t* out, int x, int y, int z, int bytes) {
    __shared__ __align__(128) uint8_t tile[8192];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); fence_proxy_async(); mbar_expect_tx(&bar, bytes); tma_load_3d(tile, &maps.m[idx], &bar, x, y, z); }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
__global__ void k_global(const CUtensorMap* map, uint8_t* out, int x, int y, int z, int bytes) {
    __shared__ __align__(128) uint8_t tile[8192];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); fence_proxy_async(); mbar_expect_tx(&bar, bytes); tma_load_3d(tile, map, &bar, x, y, z); }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
    int v = argc > 1 ? atoi(argv[1]) : 0;
    int bw = argc > 2 ? atoi(argv[2]) : 48, bh = argc > 3 ? atoi(argv[3]) : 38;
    const int w = 640, h = 480, n = 2, pitch = 640;
    std::vector<uint8_t> img((size_t)pitch * h * n);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 7 + i / 640);
    uint8_t *d, *out; cudaMalloc(&d, img.size()); cudaMalloc(&out, 8192);
    cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    Maps maps; memset(&maps, 0, sizeof(maps));
    adb_status s = encode_tma_u8_3d(&maps.m[3], d, w, h, n, pitch, (size_t)pitch * h, bw, bh);
    printf("variant %d box %dx%d encode status %d %s\n", v, bw, bh, s, adb_last_error());
    const int x = argc > 4 ? atoi(argv[4]) : 16, y = 17, z = 1, bytes = bw * bh;
    if (v == 0) k_single<<<1, 128>>>(maps.m[3], out, x, y, z, bytes);
    if (v == 1) k_array<<<1, 128>>>(maps, 3, out, x, y, z, bytes);
    if (v == 2) { CUtensorMap* dm; cudaMalloc(&dm, sizeof(CUtensorMap)); cudaMemcpy(dm, &maps.m[3], sizeof(CUtensorMap), cudaMemcpyHostToDevice); k_global<<<1, 128>>>(dm, out, x, y, z, bytes); }
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint8_t> o(bytes); cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < bh; ++r) for (int c = 0; c < bw; ++c) { uint8_t ref = img[(size_t)z * pitch * h + (size_t)(y + r) * pitch + x + c]; bad += o[r * bw + c] != ref; }
        printf("  mismatches %d\n", bad);
    }
    return 0;
}

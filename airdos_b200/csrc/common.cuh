// common.cuh -- shared helpers for libairdos_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/airdos_b200.h"

namespace adb {

// ---- error plumbing: statuses never abort the host (failure detection) ----
void set_error(const char* fmt, ...);
adb_status cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define ADB_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return adb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define ADB_CHECK(cond, status, ...)  \
    do {                              \
        if (!(cond)) {                \
            adb::set_error(__VA_ARGS__); \
            return (status);          \
        }                             \
    } while (0)

adb_status select_device(int device);  // checks for sm_100, sets current device

// ---- TMA tensor-map encoding through the runtime's driver entry point (no -lcuda needed) ----
// 3-D u8 tensor {w, h, frames} with byte strides {pitch, frame_stride}; box {bw, bh, 1}.
adb_status encode_tma_u8_3d(CUtensorMap* map, const void* base, int w, int h, int frames, size_t pitch,
                            size_t frame_stride, int bw, int bh);

// ---- device-side mbarrier / TMA wrappers (PTX; SASS: SYNCS.*, UTMALDG) ----
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// 3-D tiled TMA load global -> shared::cta, completion on an mbarrier.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"((uint64_t)map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
#endif

struct TmaMaps16 {
    CUtensorMap m[16];
};

}  // namespace adb

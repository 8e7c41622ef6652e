// common.cu -- error plumbing, device selection, TMA descriptor encoding.
#include "common.cuh"

#include <cudaTypedefs.h>

namespace adb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

adb_status cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    cudaGetLastError();  // clear the sticky-less error state
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? ADB_ERR_NO_DEVICE : ADB_ERR_CUDA;
}

adb_status select_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no CUDA device available (%s); libairdos_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count 0");
        cudaGetLastError();
        return ADB_ERR_NO_DEVICE;
    }
    ADB_CHECK(device >= 0 && device < n, ADB_ERR_INVALID, "device %d out of range (have %d)", device, n);
    cudaDeviceProp p;
    ADB_CUDA(cudaGetDeviceProperties(&p, device));
    ADB_CHECK(p.major == 10, ADB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
    ADB_CUDA(cudaSetDevice(device));
    return ADB_OK;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}

adb_status encode_tma_u8_3d(CUtensorMap* map, const void* base, int w, int h, int frames, size_t pitch,
                            size_t frame_stride, int bw, int bh) {
    PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
    ADB_CHECK(enc != nullptr, ADB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    ADB_CHECK(((uintptr_t)base & 15) == 0 && (pitch & 15) == 0 && (frame_stride & 15) == 0, ADB_ERR_INVALID,
              "TMA needs 16-byte aligned base/pitch/frame stride (base %p pitch %zu stride %zu)", base, pitch, frame_stride);
    ADB_CHECK(bw % 16 == 0 && bw <= 256 && bh <= 256, ADB_ERR_INVALID, "bad TMA box %dx%d", bw, bh);
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)(frames > 0 ? frames : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ADB_CHECK(r == CUDA_SUCCESS, ADB_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (w %d h %d pitch %zu box %dx%d)", (int)r, w, h, pitch, bw, bh);
    return ADB_OK;
}

}  // namespace adb

extern "C" {
const char* adb_last_error(void) { return adb::g_err; }
int adb_version(void) { return ADB_VERSION; }
int adb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
}

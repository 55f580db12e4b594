// Status plumbing, device check, TMA descriptor encoding. No compute here.
#include "a2f_common.cuh"
#include <atomic>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace a2f {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
    return A2F_ECUDA;
}
int set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int pdl_enabled() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("A2F_PDL");
        cached = (e && e[0] == '0') ? 0 : 1;
    }
    return cached;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int require_sm100() {
    static int cached[64] = {0};   // 0 unknown, 1 ok, 2 bad
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_error(A2F_ECUDA, "cudaGetDevice failed (no CUDA device?)");
    if (dev < 0 || dev >= 64) return set_error(A2F_EINVAL, "device index out of range");
    if (cached[dev] == 0) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
            return set_error(A2F_ECUDA, "cudaDeviceGetAttribute failed");
        cached[dev] = (major == 10) ? 1 : 2;
    }
    if (cached[dev] != 1) return set_error(A2F_EARCH, "liba2f_sm100 needs a compute capability 10.x device");
    return A2F_OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return set_error(A2F_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(A2F_EINVAL, "TMA base must be 16-byte aligned");
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bx[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) {
        if (strides_bytes[i] % 16 != 0) return set_error(A2F_EINVAL, "TMA global strides must be multiples of 16 bytes");
        gstr[i] = strides_bytes[i];
    }
    CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[160];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
        return set_error(A2F_EINVAL, msg);
    }
    return A2F_OK;
}

}  // namespace a2f

extern "C" {

int a2f_version(void) { return A2F_VERSION; }

const char* a2f_status_string(int status) {
    switch (status) {
        case A2F_OK: return "A2F_OK";
        case A2F_EINVAL: return "A2F_EINVAL: invalid shape, alignment or argument";
        case A2F_EARCH: return "A2F_EARCH: device is not sm_100 (no fallback path exists)";
        case A2F_ECUDA: return "A2F_ECUDA: CUDA runtime error";
        default: return "unknown a2f status";
    }
}

const char* a2f_last_error(void) { return a2f::g_err; }

int a2f_device_check(void) {
    int dev = 0, major = 0;
    A2F_CHECK_CUDA(cudaGetDevice(&dev));
    A2F_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return a2f::set_error(A2F_EARCH, "liba2f_sm100 needs a compute capability 10.x device");
    return A2F_OK;
}

long long a2f_launch_count(void) { return a2f::g_launches.load(); }

}  // extern "C"

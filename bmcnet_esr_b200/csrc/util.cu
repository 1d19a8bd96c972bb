// Error plumbing, device queries and TMA descriptor encoding for libbmc_b200.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace bmc {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return BMC_ERR_CUDA;
}

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) on = measure_env("BMC_PDL", 0);
    return on != 0;
}

bool pdl_small_grid(unsigned ctas) {
    static int on = -1;
    if (on < 0) on = measure_env("BMC_PDL_SMALL", 1);        // measurement builds can switch the rule off for A/B runs
    return on && 2 * (int)ctas <= sm_count();
}

int& PerDevice::cur() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    return v[dev];
}

int sm_count() {
    static PerDevice cache;
    int& n = cache.cur();
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
        n = v;
    }
    return n;
}

#ifdef BMC_MEASURE
int measure_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
#endif

// cuTensorMapEncodeTiled is a driver entry point; resolve it through the runtime so the
// library has no link-time dependency on libcuda (it must load on a GPU-less build host).
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(p);
    }
    return fn;
}

#ifdef BMC_ACT_BF16
static const CUtensorMapDataType kTmapType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
static const CUtensorMapDataType kTmapType = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif

int make_tmap_2d_act(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint32_t box_rows, uint32_t box_cols) {
    encode_tiled_fn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return BMC_ERR_CUDA;
    }
    if (((uintptr_t)base & 15) || (cols * 2) % 16 || box_cols * 2 != 128 || box_rows > 256) {
        set_error("make_tmap_2d_act: bad alignment/box (base %p cols %llu box %ux%u)", base,
                  (unsigned long long)cols, box_rows, box_cols);
        return BMC_ERR_ARG;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, kTmapType, 2, const_cast<void*>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu box %ux%u)",
                  (int)r, (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols);
        return BMC_ERR_CUDA;
    }
    return BMC_OK;
}

// Same tensor behind 32-channel boxes (64-byte rows in shared memory, 64-byte swizzle): the operand
// granule of conv_slab2_tc (gemm_slab2.cu).
int make_tmap_2d_act_sw64(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    encode_tiled_fn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return BMC_ERR_CUDA;
    }
    if (((uintptr_t)base & 15) || (cols * 2) % 16 || cols % 32 || box_rows > 256 || box_rows % 8) {
        set_error("make_tmap_2d_act_sw64: bad alignment/box (base %p cols %llu box rows %u)", base,
                  (unsigned long long)cols, box_rows);
        return BMC_ERR_ARG;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, kTmapType, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (sw64) failed with CUresult %d (rows %llu cols %llu box rows %u)",
                  (int)r, (unsigned long long)rows, (unsigned long long)cols, box_rows);
        return BMC_ERR_CUDA;
    }
    return BMC_OK;
}

}  // namespace bmc

extern "C" BMC_EXPORT int bmc_abi_version(void) { return BMC_ABI_VERSION; }
extern "C" BMC_EXPORT const char* bmc_last_error(void) { return bmc::g_err; }
extern "C" BMC_EXPORT const char* bmc_act_dtype(void) { return BMC_ACT_NAME; }

// TMA load throughput per SM on sm_100a (bring-up tool, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_bench tools/tma_bench.cu
// One thread per CTA streams 2-D boxes of a [rows][cols] fp16 tensor into a ring of shared-memory
// buffers (depth 4); reports bytes per clock per SM for several box shapes / swizzles / grid sizes.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    return ok;
}
constexpr int kDepth = 4;
__global__ void __launch_bounds__(128, 1) bench(const __grid_constant__ CUtensorMap map, int box_bytes, int box_rows, int n_loads,
                                                int rows_total, int lanes, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[kDepth];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kDepth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const long long t0 = clock64();
        // every "load" = `lanes` boxes issued by `lanes` lanes at once into one buffer
        for (int i = 0; i < n_loads; ++i) {
            const int st = i % kDepth;
            if (i >= kDepth) { uint32_t n = 0; while (!mbar_try(&bar[st], ((i / kDepth) - 1) & 1)) if (++n > (1u << 26)) __trap(); }
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[st])), "r"(box_bytes * lanes) : "memory");
            __syncwarp();
            if (lane < lanes) {
                const int row = (int)(((long)(blockIdx.x * 977 + i * lanes + lane) * box_rows) % (rows_total - box_rows));
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(smem + (st * lanes + lane) * box_bytes)), "l"(&map), "r"(smem_u32(&bar[st])), "r"(0), "r"(row) : "memory");
            }
        }
        for (int i = n_loads; i < n_loads + kDepth; ++i) {
            const int st = i % kDepth;
            uint32_t n = 0; while (!mbar_try(&bar[st], ((i / kDepth) - 1) & 1)) if (++n > (1u << 26)) __trap();
        }
        if (lane == 0) cyc[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    encode_fn enc = (encode_fn)fp;
    const uint64_t rows = 1 << 20;
    __half* d; cudaMalloc(&d, rows * 128 * 2); cudaMemset(d, 0, rows * 128 * 2);
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    struct Cfg { const char* name; int cols; int box_cols; int box_rows; CUtensorMapSwizzle sw; int lanes; };
    Cfg cfgs[] = {
        {"act  [.,128] box 64ch x 224 rows SW128, 1 lane ", 128, 64, 224, CU_TENSOR_MAP_SWIZZLE_128B, 1},
        {"act  [.,128] box 64ch x 224 rows SW128, 2 lanes", 128, 64, 224, CU_TENSOR_MAP_SWIZZLE_128B, 2},
        {"act  [.,128] box 32ch x 216 rows SW64,  1 lane ", 128, 32, 216, CU_TENSOR_MAP_SWIZZLE_64B, 1},
        {"act  [.,128] box 32ch x 216 rows SW64,  2 lanes", 128, 32, 216, CU_TENSOR_MAP_SWIZZLE_64B, 2},
        {"wgt  [.,64]  box 64ch x 128 rows SW128, 1 lane ", 64, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 1},
        {"wgt  [.,64]  box 64ch x 128 rows SW128, 2 lanes", 64, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 2},
        {"wgt  [.,64]  box 32ch x 128 rows SW64,  1 lane ", 64, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 1},
        {"wgt  [.,64]  box 32ch x 128 rows SW64,  4 lanes", 64, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 4},
        {"wgt  [.,32]  box 32ch x 128 rows SW64,  1 lane ", 32, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 1},
        {"wgt  [.,32]  box 32ch x 256 rows SW64,  2 lanes", 32, 32, 256, CU_TENSOR_MAP_SWIZZLE_64B, 2},
    };
    for (const Cfg& c : cfgs) {
        CUtensorMap map;
        const uint64_t r = rows * 128 / c.cols;
        cuuint64_t dims[2] = {(cuuint64_t)c.cols, r};
        cuuint64_t strides[1] = {(cuuint64_t)c.cols * 2};
        cuuint32_t box[2] = {(cuuint32_t)c.box_cols, (cuuint32_t)c.box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); continue; }
        const int box_bytes = c.box_cols * 2 * c.box_rows;
        const int smem = kDepth * c.lanes * box_bytes + 1024;
        cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int grid : {1, 148}) {
            const int n_loads = 400;
            // small footprint so that everything is an L2 hit after the first pass
            bench<<<grid, 128, smem>>>(map, box_bytes, c.box_rows, n_loads, 1 << 14, c.lanes, cyc);
            bench<<<grid, 128, smem>>>(map, box_bytes, c.box_rows, n_loads, 1 << 14, c.lanes, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)n_loads * c.lanes * box_bytes;
            printf("tma %s grid=%3d: %6.1f B/clk/SM, %5.2f cycles per row, %6.0f cycles per box\n", c.name, grid, bytes / h[0],
                   (double)h[0] / (n_loads * c.lanes * c.box_rows), (double)h[0] / (n_loads * c.lanes));
        }
    }
    return 0;
}

// SWIZZLE_64B operand check for tcgen05.mma (bring-up tool, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_sw64 tools/mma_sw64.cu
// D[128 x 256] = A[128 x 32] . B[256 x 32]^T with both operands K-major in 64-byte rows
// (what a TMA box of 32 halves with CU_TENSOR_MAP_SWIZZLE_64B writes); B is read through a
// descriptor whose start address is shifted by `shift` rows, which is how the slab convolution
// forms its 3x3 tap views.  Prints max|err| against the host for a list of shifts.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) { uint32_t n = 0; while (!mbar_try(b, par)) if (++n > (1u << 26)) { printf("timeout\n"); __trap(); } }
// descriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | layout <<61 (4 = SWIZZLE_64B)
__device__ __forceinline__ uint64_t desc64(uint32_t addr, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
// element (r, k) of a [rows][32] half tile under SWIZZLE_64B
__device__ __forceinline__ void put(uint8_t* tile, int r, int k, float v) {
    const int off = r * 64 + (((k >> 3) ^ ((r >> 1) & 3)) << 4) + ((k & 7) << 1);
    *reinterpret_cast<__half*>(tile + off) = __float2half(v);
}
__host__ __device__ inline float aval(int m, int k) { return (float)((m * 5 + k * 3) % 7 - 3); }
__host__ __device__ inline float bval(int n, int k) { return (float)((n * 3 + k) % 5 - 2); }

__global__ void __launch_bounds__(128, 1) check(float* out, int shift, int a_shift, int iters, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                         // [128 + 32 rows][32]
    uint8_t* sb = smem + (128 + 32) * 64;       // [256 + 32 rows][32]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < (128 + 32) * 32; i += 128) put(sa, i / 32, i % 32, aval(i / 32 - a_shift, i % 32));
    for (int i = threadIdx.x; i < (256 + 32) * 32; i += 128) put(sb, i / 32, i % 32, bval(i / 32 - shift, i % 32));
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = idesc_f16(128, 256);
        const uint32_t a0 = smem_u32(sa) + a_shift * 64, b0 = smem_u32(sb) + shift * 64;
        for (int k = 0; k < 2; ++k) {
            const uint64_t da = desc64(a0 + k * 32, 512), db = desc64(b0 + k * 32, 512);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        if (iters > 0) {      // timing: back-to-back N=256 K=16 MMAs on SWIZZLE_64B operands, tap view rotating
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                const uint32_t bb = b0 + (it % 9) * 64;
                for (int k = 0; k < 2; ++k) {
                    const uint64_t da = desc64(a0 + k * 32, 512), db = desc64(bb + k * 32, 512);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            mbar_wait(&bar, 1);
            cyc[blockIdx.x] = clock64() - t0;
        }
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* o = out + (size_t)(warp * 32 + lane) * 256;
    for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c * 32) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) o[c * 32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

int main() {
    const int smem = (128 + 32) * 64 + (256 + 32) * 64 + 1024;
    cudaFuncSetAttribute(check, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float* out;
    cudaMalloc(&out, 128 * 256 * 4);
    for (int a_shift : {0, 3})
        for (int shift : {0, 1, 2, 3, 4, 5, 7, 8, 13, 16, 27}) {
            check<<<1, 128, smem>>>(out, shift, a_shift, 0, nullptr);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<float> h(128 * 256);
            cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 256; ++n) {
                    float ref = 0;
                    for (int k = 0; k < 32; ++k) ref += aval(m, k) * bval(n, k);
                    maxerr = fmax(maxerr, fabs(ref - h[(size_t)m * 256 + n]));
                }
            printf("sw64 a_shift=%d b_shift=%2d: max|err| = %g\n", a_shift, shift, maxerr);
        }
    long long* cyc;
    cudaMalloc(&cyc, 148 * 8);
    for (int grid : {1, 148}) {
        check<<<grid, 128, smem>>>(out, 3, 0, 2048, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        printf("sw64 timing grid=%d: %.1f cycles per 128x256x16 MMA\n", grid, (double)h[0] / (2048 * 2));
    }
    return 0;
}

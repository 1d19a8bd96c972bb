// tcgen05.mma micro-benchmark + layout check for sm_100a (bring-up tool, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_bench tools/mma_bench.cu
// Measures cycles per MMA (issue thread blocked on the tensor pipe) for cta_group::1 (M=128) and
// cta_group::2 (M=256 across a CTA pair) with operands in shared memory (K-major SWIZZLE_128B),
// and verifies D = A.B^T against the host for both, i.e. the operand / accumulator split of the
// 2-CTA form: CTA r holds A rows [128r,128r+128), B rows [N/2*r, N/2*(r+1)), D rows [128r,..) x N.
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t n = 0;
    while (!mbar_try(b, par)) if (++n > (1u << 26)) { printf("timeout\n"); __trap(); }
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFF) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16); }
__device__ __forceinline__ uint64_t desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// K-major SWIZZLE_128B tile of [rows][64] halves: element (r, k) lives at
//   r*128 + (((k/8) ^ (r%8)) * 16) + (k%8)*2       (what TMA SWIZZLE_128B produces)
__device__ __forceinline__ void put(__half* tile, int r, int k, float v) {
    const int off = r * 128 + (((k >> 3) ^ (r & 7)) << 4) + ((k & 7) << 1);
    *reinterpret_cast<__half*>(reinterpret_cast<char*>(tile) + off) = __float2half(v);
}
__host__ __device__ inline float aval(int m, int k) { return (float)((m * 5 + k * 3) % 7 - 3); }
__host__ __device__ inline float bval(int n, int k) { return (float)((n * 3 + k) % 5 - 2); }

// grid = clusters of CG CTAs; every cluster runs the same test.  out: [clusters][CG*128][N] fp32, cyc: [clusters]
template <int CG, int N>
__global__ void __launch_bounds__(128, 1) bench(float* out, long long* cyc, int iters, int row_shift, int mode) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __half* sa = (__half*)smem;                          // [128 + 16 rows][64]  (extra rows for the shifted view)
    __half* sb = (__half*)(smem + (128 + 16) * 128);     // [N/CG (+16) rows][64]
    const int b_shift = (mode & 512) ? row_shift : 0;   // mode 512: the row shift applies to B instead of A
    const int a_shift = (mode & 512) ? 0 : row_shift;
    __shared__ uint64_t bar, ring[8], done_bar, done2;
    __shared__ uint32_t tmem_s;
    const int rank = CG == 2 ? (int)cg::this_cluster().block_rank() : 0;
    const int cluster = blockIdx.x / CG;
    constexpr int NB = N / CG;
    for (int i = threadIdx.x; i < (128 + 16) * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        put(sa, r, k, aval(rank * 128 + r - a_shift, k));           // row r of the slab = logical row r - shift
    }
    for (int i = threadIdx.x; i < (NB + 16) * 64; i += 128) put(sb, i / 64, i % 64, bval(rank * NB + i / 64 - b_shift, i % 64));
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1); mbar_init(&done_bar, 1); mbar_init(&done2, 1);
        for (int i = 0; i < 8; ++i) mbar_init(&ring[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar)) : "memory");   // phase 0 complete
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(256) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(256) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cg::this_cluster().sync(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    constexpr uint32_t idesc = idesc_f16(128 * CG, N);
    constexpr uint32_t hi = desc_hi(1024);
    long long t = 0;
    const bool dual = (mode & 64) != 0;                    // two issuing warps, one accumulator each
    const bool split = (mode & 128) != 0;                  // dual issuers, the second one in the peer CTA (CG == 2)
    const bool issuer = split ? (threadIdx.x == 0) : ((threadIdx.x == 0 || (dual && threadIdx.x == 32)) && rank == 0);
    if (issuer) {
        const int who = split ? rank : (threadIdx.x >> 5);
        const uint32_t a_lo = desc_lo(smem_u32(sa) + a_shift * 128, 16), b_lo = desc_lo(smem_u32(sb) + b_shift * 128, 16);
        const int period = (mode & 32) ? 4 : 2;            // wait / commit every 16 or 8 MMAs
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            // mode bits: 1 commit every `period` iterations, 2 fence::after_thread_sync, 4 alternate accumulators,
            //            8 rotate the A start row, 16 try_wait on a completed barrier, 32 period 16 MMAs, 64 dual issuer
            if ((it % period) == 0) {
                if (mode & 16) mbar_wait(&done_bar, 0);
                if (mode & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t acc = tmem + (dual ? who * N : ((mode & 4) ? (it & 1) * N : 0));
            const uint32_t a_it = a_lo + ((mode & 8) ? (it % 9) * 8 : 0);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                mma<CG>(acc, desc(a_it + k * 2, hi), desc(b_lo + k * 2, hi), idesc, (it > 1 || k) ? 1u : 0u);
            if ((mode & 1) && (it % period) == period - 1) { if (mode & 256) commit<CG>(&ring[(it >> 1) & 7]); else commit<1>(&ring[(it >> 1) & 7]); }
        }
        commit<CG>(who ? &done2 : &bar);
        mbar_wait(who ? &done2 : &bar, 0);
        t = clock64() - t0;
        if (who == 0) cyc[cluster] = t;
    } else if (threadIdx.x == 0 && !split) {
        mbar_wait(&bar, 0);                               // peer CTA: multicast commit arrives here too
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // read back: warp w -> lanes 32w..32w+31, all N columns in chunks of 32
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* o = out + ((size_t)cluster * CG * 128 + rank * 128 + warp * 32 + lane) * N;
    for (int c = 0; c < N / 32; ++c) {
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
    if (CG == 2) cg::this_cluster().sync(); else __syncthreads();
    if (threadIdx.x < 32) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
}

template <int CG, int N>
void run(int clusters, int row_shift, int mode = 0) {
    const int smem = (128 + 16) * 128 + (N / CG + 16) * 128 + 1024;
    auto kern = bench<CG, N>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float* out; long long* cyc;
    cudaMalloc(&out, (size_t)clusters * CG * 128 * N * 4);
    cudaMalloc(&cyc, clusters * 8);
    for (int iters : {1, 512}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(clusters * CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {(unsigned)CG, 1, 1};
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, out, cyc, iters, row_shift, mode);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CG=%d N=%d: CUDA error %s\n", CG, N, cudaGetErrorString(e)); exit(1); }
        if (iters == 1 && (mode & ~512) == 0) {
            std::vector<float> h((size_t)CG * 128 * N);
            cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < CG * 128; ++m)
                for (int n = 0; n < N; ++n) {
                    float ref = 0;
                    for (int k = 0; k < 64; ++k) ref += aval(m, k) * bval(n, k);
                    maxerr = fmax(maxerr, fabs(ref - h[(size_t)m * N + n]));
                }
            printf("CG=%d M=%d N=%3d shift=%d: layout check max|err| = %g\n", CG, 128 * CG, N, row_shift, maxerr);
        } else if (iters > 1) {
            std::vector<long long> c(clusters);
            cudaMemcpy(c.data(), cyc, clusters * 8, cudaMemcpyDeviceToHost);
            const double per = (double)c[0] / (iters * 4) / ((mode & (64 | 128)) ? 2 : 1);
            printf("mode=%2d ", mode);
            printf("CG=%d M=%d N=%3d clusters=%3d: %.1f cycles per MMA (K=16) -> %.0f MAC/clk/SM\n", CG, 128 * CG, N, clusters, per,
                   128.0 * N * 16 / per);
        }
    }
    cudaFree(out); cudaFree(cyc);
}

int main(int argc, char** argv) {
    if (argc > 1 && argv[1][0] == 'b') {
        for (int sh : {0, 1, 3, 5, 8, 13}) run<1, 256>(1, sh, 512);
        run<1, 256>(148, 3, 512);
        return 0;
    }
    if (argc > 1) {
        for (int mode : {0, 17, 17 + 32, 64, 64 + 17, 64 + 17 + 32, 64 + 31}) run<1, 128>(148, 0, mode);
        for (int mode : {64 + 17, 256 + 64 + 17, 256 + 64 + 17 + 32, 256 + 17}) run<2, 128>(74, 0, mode);
        return 0;
    }
    run<1, 64>(1, 0);
    run<1, 128>(1, 0);
    run<1, 128>(1, 3);
    run<1, 256>(1, 0);
    run<1, 128>(148, 0);
    run<1, 256>(148, 0);
    run<2, 128>(1, 0);
    run<2, 128>(1, 5);
    run<2, 256>(1, 0);
    run<2, 128>(74, 0);
    run<2, 256>(74, 0);
    return 0;
}

// Shared helpers for libbmc_b200: error plumbing and the sm_100a PTX wrappers
// (mbarrier, TMA, tcgen05/TMEM) used by the kernels.  No torch types anywhere.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <utility>

#include "../../include/bmc_b200.h"

#define BMC_EXPORT __attribute__((visibility("default")))

// 16-bit activation / weight type of the tensor-core path.  Default fp16 (11-bit significand):
// measured against the reference fp32 forward, bf16 operands miss the max-abs 1e-2 bar on the
// full BMCNet (1.2e-2) while fp16 holds it with margin (DESIGN.md "Precision").  Both feed the
// same tcgen05 kind::f16 instruction at the same rate; -DBMC_ACT_BF16 selects bf16.
#include <cuda_fp16.h>
#ifdef BMC_ACT_BF16
typedef __nv_bfloat16 act_t;
#define BMC_ACT_NAME "bf16"
#else
typedef __half act_t;
#define BMC_ACT_NAME "f16"
#endif

namespace bmc {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define BMC_CUDA(expr)                                                           \
    do {                                                                         \
        cudaError_t _e = (expr);                                                 \
        if (_e != cudaSuccess) return ::bmc::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define BMC_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            ::bmc::set_error(__VA_ARGS__); \
            return BMC_ERR_ARG;           \
        }                                 \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (148 on B200); cached per device.
int sm_count();

// Measurement switches exist only in a -DBMC_MEASURE build (`build.py --measure`):
// there `measure_env` reads an integer from the environment; in the product library it is the constant default,
// so no environment variable can change what a kernel computes.
#ifdef BMC_MEASURE
int measure_env(const char* name, int dflt);
#else
inline int measure_env(const char*, int dflt) { return dflt; }
#endif

// Per-device "already done" state for one call site (shared-memory attributes are per device and function; two
// devices driven from one process each need their own cudaFuncSetAttribute).
struct PerDevice {
    int v[64] = {};
    int& cur();                     // slot of the calling thread's current device
};

// Launch with programmatic dependent launch -- OFF by default, BMC_PDL=1 enables it (measured on B200: the
// persistent kernels fill shared memory, a successor cannot become resident before a predecessor CTA exits, and
// the step got 0.6-1 % slower: 3917 vs 3895 us plain, 12621 vs 12502 us BMCNet).  With it the kernel may become resident while its
// predecessor on the stream is still draining; it must execute pdl_wait() before touching global memory the
// predecessor reads or writes.  Inside a captured CUDA graph this becomes a programmatic dependency edge, so the
// prologue (barrier init, TMEM allocation, descriptor prefetch) and the launch latency overlap the previous
// kernel's tail instead of sitting between the two.
bool pdl_enabled();
// Small launches (at most half the SMs get a CTA: batch-1 inference, infer_BMCNet.py:46-68) ARE launched with the
// attribute: there the SMs are mostly idle, the successor's CTAs become resident at once and their prologues
// (barrier init, TMEM allocation, descriptor prefetch, the weight-only part of the pipeline fill) run under the
// predecessor's tail -- the step is a chain of ~23 / 54 dependent kernels of 15-30 CTAs each, and the gaps between
// them are a large part of its latency.
bool pdl_small_grid(unsigned ctas);
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = (pdl_enabled() || pdl_small_grid(grid.x * grid.y * grid.z)) ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---------------------------------------------------------------- geometry
// Padded NHWC activation layout: image b occupies rows [b*R, (b+1)*R); row r <-> padded
// pixel (r / Wp, r % Wp) with Wp = W + 2; rows >= (H+2)*Wp are tail padding.  Halo and tail
// rows always hold zeros, so a 3x3 tap is a constant row shift dy*Wp + dx.
struct Geom {
    int B, H, W, Wp, R;
    __host__ __device__ static inline Geom make(int B, int H, int W) {
        Geom g;
        g.B = B; g.H = H; g.W = W; g.Wp = W + 2;
        g.R = (((H + 2) * (W + 2)) + 127) / 128 * 128;
        return g;
    }
    __host__ __device__ inline long rows() const { return (long)B * R; }
    // row within image -> is it a real pixel?  (y, x) in unpadded coordinates on success
    __device__ inline bool interior(int r, int& y, int& x) const {
        int py = r / Wp, px = r - py * Wp;
        y = py - 1; x = px - 1;
        return (py >= 1) & (py <= H) & (px >= 1) & (px <= W);
    }
};

#ifdef __CUDACC__
// ---------------------------------------------------------------- small device helpers
// Programmatic dependent launch: wait until the prerequisite grids have completed and their writes are visible
// (a no-op when the kernel was not launched with the attribute).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t.reg .b32 r;\n\t"
        "elect.sync r|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (-> CUDA error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            printf("bmc: mbarrier timeout block (%d,%d) thread %d\n", blockIdx.x, blockIdx.y,
                   threadIdx.x);
            __trap();
        }
    }
}

// ---------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// 2-D tiled load: c0 = innermost (channel) coordinate, c1 = row coordinate (may be < 0 or
// past the end: out-of-bounds elements are written as zeros, which is what conv padding needs).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* map, uint64_t* bar, int c0,
                                            int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 2-D tiled store shared -> global (bulk async group); rows / columns outside the tensor are clipped.
__device__ __forceinline__ void tma_store_2d(const void* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
// 2-D tiled REDUCE shared -> global: global[box] += smem[box], element type from the tensor map (act16 add in L2)
__device__ __forceinline__ void tma_reduce_add_2d(const void* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Pull a 2-D box into L2 only (no shared-memory destination, no barrier): used to start the HBM read
// of a tile whose shared-memory buffer is still busy.
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1)
                 : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread for the whole CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                     "r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// 16 lanes x (8 columns x 4 repetitions): the mma-accumulator fragment layout -- per repetition k thread t holds
// (lane t/4, columns 8k + 2(t%4), +1) in v[4k], v[4k+1] and (lane t/4 + 8, same columns) in v[4k+2], v[4k+3].
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// Four 8x8 b16 matrices, transposed on the way: lane l supplies the shared-memory address of row (l % 8) of matrix
// (l / 8); the thread's fragment of matrix m is register m (row t/4, columns 2(t%4), +1 of the UNtransposed matrix).
__device__ __forceinline__ void stmatrix_x4_trans(void* smem_row, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};"
                 ::"r"(smem_u32(smem_row)), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, sm_100):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [49,52) base offset
//   | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// The MMA-issuing thread is alone in its warp, so every scalar instruction between two
// tcgen05.mma costs its full latency.  Descriptors are therefore split: the high word is a
// compile-time constant per layout, the low word (start address >> 4 | LBO field) advances by a
// plain 32-bit add.
__host__ __device__ constexpr uint32_t umma_desc_hi_sw128(uint32_t sbo_bytes) {
    return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
}
// K-major rows of 64 bytes under SWIZZLE_64B (layout code 4): 8-row groups are `sbo_bytes` = 512 apart
__host__ __device__ constexpr uint32_t umma_desc_hi_sw64(uint32_t sbo_bytes) {
    return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (4u << 29);
}
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t lo, uint32_t hi) {
    return ((uint64_t)hi << 32) | lo;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): act_t x act_t -> fp32, M x N tile.
#ifdef BMC_ACT_BF16
#define BMC_UMMA_FMT 1u                    // F16F32Format::BF16
#else
#define BMC_UMMA_FMT 0u                    // F16F32Format::F16
#endif
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, bool a_mn_major,
                                                       bool b_mn_major) {
    return (1u << 4)                       // c_format  = F32
           | (BMC_UMMA_FMT << 7)           // a_format
           | (BMC_UMMA_FMT << 10)          // b_format
           | ((a_mn_major ? 1u : 0u) << 15)
           | ((b_mn_major ? 1u : 0u) << 16)
           | ((uint32_t)(N >> 3) << 17)
           | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- CTA pairs (cluster of 2, tcgen05 cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes)
                 : "memory");
}
// TMA load whose completion is signalled on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 split over the pair; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once this thread's MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                     "r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}

// ---------------------------------------------------------------- 16-bit packing
#ifdef BMC_ACT_BF16
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_act2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}
__device__ __forceinline__ act_t to_act(float f) { return __float2bfloat16(f); }
__device__ __forceinline__ float from_act(act_t a) { return __bfloat162float(a); }
#else
// fp16 saturates at +-65504 instead of overflowing to inf
__device__ __forceinline__ float sat_h(float f) { return fminf(fmaxf(f, -65504.f), 65504.f); }
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi) {
    uint32_t r;       // one F2FP.SATFINITE: round to nearest, clamp to +-65504 (first operand -> upper half)
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 unpack_act2(uint32_t u) {
    __half2 v = *reinterpret_cast<__half2*>(&u);
    return __half22float2(v);
}
__device__ __forceinline__ act_t to_act(float f) { return __float2half_rn(sat_h(f)); }
__device__ __forceinline__ float from_act(act_t a) { return __half2float(a); }
#endif
#endif  // __CUDACC__

// ---------------------------------------------------------------- host: TMA descriptors
// 2-D act_t row-major [rows][cols] tensor, box = [box_rows][box_cols], 128-byte swizzle.
int make_tmap_2d_act(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint32_t box_rows, uint32_t box_cols);
// Same tensor, box = [box_rows][32], 64-byte swizzle (box_rows a multiple of 8).
int make_tmap_2d_act_sw64(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

}  // namespace bmc

// tcgen05 / TMEM / TMA kernels for sm_100a: the implicit-GEMM convolution and the BIE
// attention-logit GEMM.
//
// conv_gemm_tc:  one CTA per 128-row tile of padded pixels, N (=128 or 32) output channels.
//   warp 0      TMA producer: per K step one [128 x 64] activation box (row coordinate shifted
//               by the tap offset -- out-of-range rows arrive as zeros) and one [N x 64] weight box
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue (M=128, N, K=16, bf16 -> fp32)
//   warps 2-5   epilogue: tcgen05.ld (one accumulator row per thread), bias / ReLU / residual /
//               halo masking, bf16 (and optional fp32) stores
//   smem ring of kStages {A 16 KB, B N*128 B} tiles in the 128-byte-swizzled K-major layout, full /
//   empty mbarriers; tcgen05.commit releases a stage when the MMAs that read it are done.
#include "gemm_epi.cuh"

namespace bmc {
namespace {

constexpr int kThreadsTc = 192;

template <int N>
struct TcCfg {
    static constexpr int kABytes = kTileM * kChunkK * 2;   // 16384
    static constexpr int kBBytes = N * kChunkK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = N < 32 ? 32 : N;      // power of two >= 32
};

template <int N, int STAGES>
__global__ void __launch_bounds__(kThreadsTc, 2) conv_gemm_tc(const __grid_constant__ GemmParams p) {
    using Cfg = TcCfg<N>;
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    // dynamic smem base is only guaranteed 16-byte aligned: realign to 1024 for SWIZZLE_128B
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[N], gamma_s[N], beta_s[N];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const GemmJobDev& job = p.jobs[blockIdx.y];
    const int tile = blockIdx.x;
    const long m0 = (long)tile * kTileM;
    const int img = tile / p.tiles_per_img;

    int iters = 0;
    for (int s = 0; s < p.n_seg; ++s) iters += p.n_taps * p.chunks[s];

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        mbar_fence_init();
        for (int s = 0; s < p.n_seg; ++s) tma_prefetch_desc(&p.maps[job.a_map[s]]);
        tma_prefetch_desc(&p.maps[job.w_map]);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, Cfg::kTmemCols);
    if (threadIdx.x >= 64 && threadIdx.x - 64 < N) {
        const int n = threadIdx.x - 64;
        bias_s[n] = job.bias ? job.bias[n] : 0.f;
        gamma_s[n] = job.ln_gamma ? job.ln_gamma[n] : 1.f;
        beta_s[n] = job.ln_gamma ? job.ln_beta[n] : 0.f;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_acc = tmem_base_s;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const int w_row0 = job.w_row_base + img * job.w_img_stride;
            int it = 0, kchunk = 0;
            for (int s = 0; s < p.n_seg; ++s) {
                const CUtensorMap* amap = &p.maps[job.a_map[s]];
                for (int t = 0; t < p.n_taps; ++t) {
                    const int arow = job.a_row_base[s] + (int)m0 + p.tap_off[t];
                    for (int c = 0; c < p.chunks[s]; ++c, ++it, ++kchunk) {
                        const int st = it % STAGES;
                        if (it >= STAGES) mbar_wait(&empty_bar[st], ((it / STAGES) - 1) & 1);
                        uint8_t* sa = smem + st * Cfg::kStageBytes;
                        mbar_expect_tx(&full_bar[st], Cfg::kStageBytes);
                        tma_load_2d(sa, amap, &full_bar[st], job.a_col_base[s] + c * kChunkK, arow);
                        tma_load_2d(sa + Cfg::kABytes, &p.maps[job.w_map], &full_bar[st], 0,
                                    kchunk * job.w_rows + w_row0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(kTileM, N, false, false);
            for (int it = 0; it < iters; ++it) {
                const int st = it % STAGES;
                mbar_wait(&full_bar[st], (it / STAGES) & 1);
                tc_fence_after_sync();
                const uint32_t sa = smem_u32(smem + st * Cfg::kStageBytes);
                // K-major SWIZZLE_128B: 8-row groups 1024 B apart; +32 B (2 x 16 B) per K=16 slice
                const uint32_t a_lo = umma_desc_lo(sa, 16), b_lo = umma_desc_lo(sa + Cfg::kABytes, 16);
                constexpr uint32_t hi = umma_desc_hi_sw128(1024);
                const uint32_t acc_first = it > 0 ? 1u : 0u;
#pragma unroll
                for (int k = 0; k < kChunkK / 16; ++k)
                    umma_f16(tmem_acc, umma_desc(a_lo + k * 2, hi), umma_desc(b_lo + k * 2, hi), idesc,
                             k == 0 ? acc_first : 1u);
                umma_commit(&empty_bar[st]);       // stage reusable once these MMAs retire
            }
            umma_commit(&acc_bar);                 // accumulator complete
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int r_tile = q * 32 + lane;
        const long m = m0 + r_tile;
        const int r_img = (int)(m - (long)img * p.g.R);
        int y, x;
        const bool valid = p.g.interior(r_img, y, x);
        mbar_wait(&acc_bar, 0);
        tc_fence_after_sync();
        EpiRow r;
        r.res = job.residual ? job.residual + (job.res_row_base + m) * N : nullptr;
        r.out = job.out ? job.out + (job.out_row_base + m) * N : nullptr;
        r.outf = job.out_f32 ? job.out_f32 + (job.out_row_base + m) * N : nullptr;
        r.valid = valid; r.store = true; r.relu = job.relu != 0; r.ln_eps = job.ln_eps;
        const uint32_t trow = tmem_acc + ((uint32_t)(q * 32) << 16);
        const bool ln = job.ln_gamma != nullptr;
        float mu = 0.f, rstd = 1.f;
        if (ln) epi_ln_stats<N>(trow, bias_s, job.ln_eps, mu, rstd);
#pragma unroll 1
        for (int c = 0; c < N / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(trow + c * 32, v);
            tmem_ld_wait();
            epi_chunk(v, c, r, bias_s, ln, mu, rstd, gamma_s, beta_s);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_acc, Cfg::kTmemCols);
    }
}

// ---------------------------------------------------------------- attention logits
// D[c, c'] = sum_p centres[p, c] * v[p, c'] over a range of pixels of one image: both operands
// are "MN-major" (the contraction index p is the slow, row index of the [rows][128] tensors).
// smem tile per operand and stage: two [64 pixels x 64 channels] SWIZZLE_128B boxes (8 KB each).
constexpr int kAttStages = 4;
constexpr int kAttBox = 64 * 64 * 2;             // 8192
constexpr int kAttStageBytes = 4 * kAttBox;      // c lo/hi, v lo/hi

__global__ void __launch_bounds__(kThreadsTc) att_tc(const __grid_constant__ AttParams p) {
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[kAttStages], empty_bar[kAttStages], acc_bar;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x, b = blockIdx.y, pair = blockIdx.z;
    const int pix0 = split * p.pix_per_split;
    const int pix1 = min(pix0 + p.pix_per_split, p.g.R);
    const int iters = (pix1 - pix0) / 64;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kAttStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        mbar_fence_init();
        tma_prefetch_desc(&p.map_c);
        tma_prefetch_desc(&p.map_v);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_acc = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            const int crow = (int)(p.c_row_base[pair] + (long)b * p.g.R) + pix0;
            const int vrow = (int)(p.v_row_base[pair] + (long)b * p.g.R) + pix0;
            for (int it = 0; it < iters; ++it) {
                const int st = it % kAttStages;
                if (it >= kAttStages) mbar_wait(&empty_bar[st], ((it / kAttStages) - 1) & 1);
                uint8_t* s0 = smem + st * kAttStageBytes;
                mbar_expect_tx(&full_bar[st], kAttStageBytes);
                tma_load_2d(s0, &p.map_c, &full_bar[st], 0, crow + it * 64);
                tma_load_2d(s0 + kAttBox, &p.map_c, &full_bar[st], 64, crow + it * 64);
                tma_load_2d(s0 + 2 * kAttBox, &p.map_v, &full_bar[st], 0, vrow + it * 64);
                tma_load_2d(s0 + 3 * kAttBox, &p.map_v, &full_bar[st], 64, vrow + it * 64);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, 128, true, true);
            for (int it = 0; it < iters; ++it) {
                const int st = it % kAttStages;
                mbar_wait(&full_bar[st], (it / kAttStages) & 1);
                tc_fence_after_sync();
                const uint32_t sa = smem_u32(smem + st * kAttStageBytes);
                // MN-major SWIZZLE_128B: 64-channel blocks LBO = 8192 B apart, 8-pixel groups
                // SBO = 1024 B apart; one K=16 slice = 16 pixel rows = 2048 B (128 x 16 B).
                const uint32_t a_lo = umma_desc_lo(sa, kAttBox), b_lo = umma_desc_lo(sa + 2 * kAttBox, kAttBox);
                constexpr uint32_t hi = umma_desc_hi_sw128(1024);
                const uint32_t acc_first = it > 0 ? 1u : 0u;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_acc, umma_desc(a_lo + k * 128, hi), umma_desc(b_lo + k * 128, hi), idesc,
                             k == 0 ? acc_first : 1u);
                umma_commit(&empty_bar[st]);
            }
            umma_commit(&acc_bar);
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;             // centre channel c
        float* dst = p.partial + ((((long)pair * p.g.B + b) * p.n_split + split) * 128 + row) * 128;
        if (iters > 0) {
            mbar_wait(&acc_bar, 0);
            tc_fence_after_sync();
        }
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            if (iters > 0) {
                tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + c * 32, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0;
            }
            float4* op = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                op[u] = make_float4(__uint_as_float(v[u * 4]) * p.scale, __uint_as_float(v[u * 4 + 1]) * p.scale,
                                    __uint_as_float(v[u * 4 + 2]) * p.scale, __uint_as_float(v[u * 4 + 3]) * p.scale);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_acc, 128);
    }
}

template <int N, int STAGES>
int launch_tc(const GemmParams& p, cudaStream_t st) {
    auto kern = conv_gemm_tc<N, STAGES>;
    const int smem = STAGES * TcCfg<N>::kStageBytes + 1024;
    static PerDevice configured_dev;
    int& configured = configured_dev.cur();
    if (!configured) {
        BMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = 1;
    }
    dim3 grid((unsigned)(p.g.B * p.tiles_per_img), (unsigned)p.n_jobs);
    kern<<<grid, kThreadsTc, smem, st>>>(p);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

}  // namespace

int launch_conv_gemm_simt(const GemmParams& p, cudaStream_t st);
int launch_att_simt(const AttParams& p, cudaStream_t st);

#ifdef BMC_MEASURE
// conv_pair_tc (gemm_pair.cu, cta_group::2): measured slower than the single-CTA slab kernels;
// compiled into `build.py --measure` libraries only
static bool use_pair_ok(const GemmParams& p) { return p.jobs[0].a_map64[0] >= 0 && slab_supported(p) && pair_supported(p); }
#else
static bool use_pair_ok(const GemmParams&) { return false; }
#endif

int launch_conv_gemm(const GemmParams& p, int impl, cudaStream_t st) {
    if (p.tap1_mask && (impl != 0 || p.jobs[0].a_map64[0] < 0 || !slab_supported(p))) {
        set_error("conv_gemm: centre-tap segments exist only in the slab kernel");
        return BMC_ERR_UNSUPPORTED;
    }
#ifdef BMC_MEASURE
    if (impl == 0 && use_pair_ok(p)) return launch_conv_pair(p, st);
#endif
    // conv_slab2_tc wins on launches made of 3x3 segments only; launches with centre-tap segments (one short
    // phase per 32 channels, each with its own slab) are faster on conv_slabt_tc (measured)
    if (impl == 0 && !p.tap1_mask && slab2_supported(p)) return launch_conv_slab2(p, st);
    for (int j = 0; j < p.n_jobs; ++j)
        if (p.jobs[j].out_accumulate) {
            set_error("conv_gemm: an accumulating output (in-place residual add) exists only in conv_slab2_tc");
            return BMC_ERR_UNSUPPORTED;
        }
    if (impl == 0 && p.jobs[0].a_map64[0] >= 0 && slab_supported(p) && slabt_supported(p)) return launch_conv_slabt(p, st);
    if (p.tap1_mask) return launch_conv_slab(p, st);
    if (impl == 1) return launch_conv_gemm_simt(p, st);
    static int use_slab = -1;
    if (use_slab < 0) use_slab = measure_env("BMC_CONV_SLAB", 1);
    if (impl == 0 && use_slab && p.jobs[0].a_map64[0] >= 0 && slab_supported(p)) return launch_conv_slab(p, st);
    static int stages = 0;
    if (stages == 0) {
        stages = measure_env("BMC_TC_STAGES", 3);
        if (stages != 3 && stages != 4 && stages != 6) stages = 3;
    }
    if (p.n == 128) {
        if (stages == 6) return launch_tc<128, 6>(p, st);
        if (stages == 4) return launch_tc<128, 4>(p, st);
        return launch_tc<128, 3>(p, st);
    }
    if (p.n == 32) return launch_tc<32, 4>(p, st);
    set_error("conv_gemm: unsupported N=%d (128 or 32)", p.n);
    return BMC_ERR_UNSUPPORTED;
}

int launch_att(const AttParams& p, int impl, cudaStream_t st) {
    if (impl == 1) return launch_att_simt(p, st);
    const int smem = kAttStages * kAttStageBytes + 1024;
    static PerDevice configured_dev;
    int& configured = configured_dev.cur();
    if (!configured) {
        BMC_CUDA(cudaFuncSetAttribute(att_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = 1;
    }
    dim3 grid((unsigned)p.n_split, (unsigned)p.g.B, (unsigned)p.n_pairs);
    att_tc<<<grid, kThreadsTc, smem, st>>>(p);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

}  // namespace bmc

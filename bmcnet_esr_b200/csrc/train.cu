// Training-step kernels (reference train.py:202-237, config/train_nfs.yml:28-34).
//
//   wgrad_tc        weight gradient of a 3x3 / 1x1 convolution on the tensor cores:
//                       dW[tap][co][ci] = sum over all padded-pixel rows r of dY[r][co] * X[r + off_tap][ci]
//                   -- per tap a [128 x P] . [P x ci] GEMM whose contraction index is the PIXEL, i.e. both operands
//                   are "MN-major" views of the row-major [rows][channels] activation tensors (the same operand
//                   form as the attention-logit GEMM att_tc, gemm_tc.cu).  The tap is a row shift of the X
//                   operand, applied as a TMA row coordinate (rows outside the tensor arrive as zeros; dY has zero
//                   halo rows, so whatever real row a shift reaches across an image border is multiplied by zero).
//                   Split-K over CTAs: grid = (pixel splits, taps); fp32 partials [split][tap][128][ci].
//   wgrad_reduce    fixed-order sum of the partials (deterministic), un-scaling (loss scale) and accumulation into
//                   the fp32 gradient of the PyTorch-layout weight [co][cin][kh][kw] through a channel map (the K
//                   segments of a `torch.cat` input; padded channels map to -1).
//   dgrad           is NOT a new kernel: it is the forward slab convolution run on dY with the taps mirrored and
//                   the weight matrix transposed (a second weight pack; models/_train.py).
//   adam_amsgrad    torch.optim.Adam(amsgrad=True, weight_decay) as one fused elementwise kernel over the flat
//                   fp32 parameter / gradient / moment buffers.
#include <cmath>
#include <cstring>

#include "gemm.cuh"

namespace bmc {
namespace {

constexpr int kWgThreads = 192;
constexpr int kWgStages = 4;
constexpr int kWgBox = 64 * 64 * 2;              // [64 pixel rows][64 channels] act16, SWIZZLE_128B

struct alignas(64) WgradParams {
    CUtensorMap map_dy;                          // [rows][128], box [64][64]
    CUtensorMap map_x;                           // [rows][x_ch], box [64][64]
    int x_chunks;                                // x_ch / 64 (1 or 2)
    int n_taps;
    int tap_off[9];
    int rows;                                    // B * R (multiple of 128)
    int n_split, rows_per_split;                 // rows_per_split: multiple of 64
    float* partial;                              // [split][tap][128][x_ch]
    const act_t* dy_ptr;                         // bias gradient (optional): column sums of dY, one partial row per split,
    float* bias_partial;                         // [split][128] -- summed by the idle epilogue warps of the tap-0 CTAs
};

template <int XCH>
__global__ void __launch_bounds__(kWgThreads) wgrad_tc(const __grid_constant__ WgradParams p) {
    constexpr int kXChunks = XCH / 64;
    constexpr int kStageBytes = (2 + kXChunks) * kWgBox;
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[kWgStages], empty_bar[kWgStages], acc_bar;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x, tap = blockIdx.y;
    const int r0 = split * p.rows_per_split;
    const int r1 = min(r0 + p.rows_per_split, p.rows);
    const int iters = r1 > r0 ? (r1 - r0) / 64 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kWgStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        mbar_fence_init();
        tma_prefetch_desc(&p.map_dy);
        tma_prefetch_desc(&p.map_x);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_acc = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            const int xoff = p.tap_off[tap];
            for (int it = 0; it < iters; ++it) {
                const int st = it % kWgStages;
                if (it >= kWgStages) mbar_wait(&empty_bar[st], ((it / kWgStages) - 1) & 1);
                uint8_t* s0 = smem + st * kStageBytes;
                mbar_expect_tx(&full_bar[st], kStageBytes);
                const int row = r0 + it * 64;
                tma_load_2d(s0, &p.map_dy, &full_bar[st], 0, row);
                tma_load_2d(s0 + kWgBox, &p.map_dy, &full_bar[st], 64, row);
#pragma unroll
                for (int c = 0; c < kXChunks; ++c)
                    tma_load_2d(s0 + (2 + c) * kWgBox, &p.map_x, &full_bar[st], c * 64, row + xoff);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, XCH, true, true);
            for (int it = 0; it < iters; ++it) {
                const int st = it % kWgStages;
                mbar_wait(&full_bar[st], (it / kWgStages) & 1);
                tc_fence_after_sync();
                const uint32_t sa = smem_u32(smem + st * kStageBytes);
                // MN-major SWIZZLE_128B: 64-channel blocks LBO = 8192 B apart, 8-pixel groups SBO = 1024 B apart;
                // one K=16 slice = 16 pixel rows = 2048 B
                const uint32_t a_lo = umma_desc_lo(sa, kWgBox), b_lo = umma_desc_lo(sa + 2 * kWgBox, kWgBox);
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
        const int row = q * 32 + lane;             // output channel co
        float* dst = p.partial + (((long)split * p.n_taps + tap) * 128 + row) * XCH;
        if (p.bias_partial && tap == 0) {          // dbias partial of this pixel split, while the MMAs run
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            const act_t* d = p.dy_ptr + (long)r0 * 128 + row;
            int r = r0;
            for (; r + 4 <= r1; r += 4, d += 4 * 128) {
                s0 += from_act(d[0]); s1 += from_act(d[128]); s2 += from_act(d[256]); s3 += from_act(d[384]);
            }
            for (; r < r1; ++r, d += 128) s0 += from_act(d[0]);
            p.bias_partial[split * 128 + row] = (s0 + s1) + (s2 + s3);
        }
        if (iters > 0) {
            mbar_wait(&acc_bar, 0);
            tc_fence_after_sync();
        }
#pragma unroll 1
        for (int c = 0; c < XCH / 32; ++c) {
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
                op[u] = make_float4(__uint_as_float(v[u * 4]), __uint_as_float(v[u * 4 + 1]),
                                    __uint_as_float(v[u * 4 + 2]), __uint_as_float(v[u * 4 + 3]));
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_acc, 128);
    }
}

// grad[(co * cin_total + cmap[ci]) * taps + tap] += scale * sum_split partial[split][tap][co][ci]
// (+ the last block, when `bias_partial` is given: grad_b[co] += scale * sum_split bias_partial[split][co])
__global__ void wgrad_reduce(const float* __restrict__ partial, int n_split, int taps, int x_ch,
                             const int* __restrict__ cmap, int cin_total, int n_out, float scale,
                             float* __restrict__ grad, const float* __restrict__ bias_partial, float* __restrict__ grad_b) {
    const int per_split = taps * 128 * x_ch;
    if (bias_partial && blockIdx.x == gridDim.x - 1) {
        const int c = threadIdx.x;
        if (c < n_out) {
            float s = 0.f;
            for (int k = 0; k < n_split; ++k) s += bias_partial[k * 128 + c];
            grad_b[c] += scale * s;
        }
        return;
    }
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_split) return;
    const int ci = idx % x_ch;
    const int co = (idx / x_ch) % 128;
    const int tap = idx / (x_ch * 128);
    const int dst_c = cmap[ci];
    if (dst_c < 0 || co >= n_out) return;
    float s = 0.f;
    for (int k = 0; k < n_split; ++k) s += partial[(long)k * per_split + idx];      // fixed order: deterministic
    grad[((long)co * cin_total + dst_c) * taps + tap] += scale * s;
}

// dx = dy * (y > 0), elementwise on act16 pairs (ReLU backward; F.relu at BMCNet.py:64-73, submodules.py:33)
__global__ void relu_backward(const uint32_t* __restrict__ dy, const uint32_t* __restrict__ y, long n2,
                              uint32_t* __restrict__ dx) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const float2 g = unpack_act2(dy[i]), v = unpack_act2(y[i]);
    dx[i] = pack_act2(v.x > 0.f ? g.x : 0.f, v.y > 0.f ? g.y : 0.f);
}

// LayerNormFunction.backward (submodules.py:142-154) on packed rows: one warp per row, 4 channels per lane.
//   y_hat = (x - mu) * rstd ; g = dy * gamma ; dx = rstd * (g - y_hat * mean(g * y_hat) - mean(g))
//   dgamma = sum_rows dy * y_hat ; dbeta = sum_rows dy          (per-CTA partials, reduced in a fixed order)
// mu / rstd are recomputed from x (fp32 math on the 16-bit input, like the forward kernel in pointwise.cu).
constexpr int kLnbThreads = 256;                 // 8 rows per pass
__global__ void __launch_bounds__(kLnbThreads) layernorm_backward(const act_t* __restrict__ x, const act_t* __restrict__ dy,
                                                                  const float* __restrict__ gamma, float eps, long rows,
                                                                  act_t* __restrict__ dx, float* __restrict__ partial) {
    __shared__ float s_g[8][128], s_b[8][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 gm = *reinterpret_cast<const float4*>(gamma + lane * 4);
    float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
    for (long row = (long)blockIdx.x * 8 + warp; row < rows; row += (long)gridDim.x * 8) {
        const uint2 xr = *reinterpret_cast<const uint2*>(x + row * 128 + lane * 4);
        const uint2 dr = *reinterpret_cast<const uint2*>(dy + row * 128 + lane * 4);
        const float2 x0 = unpack_act2(xr.x), x1 = unpack_act2(xr.y), d0 = unpack_act2(dr.x), d1 = unpack_act2(dr.y);
        const float xv[4] = {x0.x, x0.y, x1.x, x1.y}, dv[4] = {d0.x, d0.y, d1.x, d1.y};
        float s = xv[0] + xv[1] + xv[2] + xv[3];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mu = s * (1.f / 128.f);
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) v += (xv[j] - mu) * (xv[j] - mu);
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const float rstd = 1.f / sqrtf(v * (1.f / 128.f) + eps);
        const float gv[4] = {gm.x, gm.y, gm.z, gm.w};
        float yh[4], g[4], sg = 0.f, sgy = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            yh[j] = (xv[j] - mu) * rstd;
            g[j] = dv[j] * gv[j];
            sg += g[j];
            sgy += g[j] * yh[j];
            ag[j] += dv[j] * yh[j];
            ab[j] += dv[j];
        }
        for (int o = 16; o; o >>= 1) { sg += __shfl_xor_sync(0xffffffffu, sg, o); sgy += __shfl_xor_sync(0xffffffffu, sgy, o); }
        const float mg = sg * (1.f / 128.f), mgy = sgy * (1.f / 128.f);
        uint2 o2;
        o2.x = pack_act2(rstd * (g[0] - yh[0] * mgy - mg), rstd * (g[1] - yh[1] * mgy - mg));
        o2.y = pack_act2(rstd * (g[2] - yh[2] * mgy - mg), rstd * (g[3] - yh[3] * mgy - mg));
        *reinterpret_cast<uint2*>(dx + row * 128 + lane * 4) = o2;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_g[warp][lane * 4 + j] = ag[j]; s_b[warp][lane * 4 + j] = ab[j]; }
    __syncthreads();
    if (threadIdx.x < 128) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 8; ++w) { a += s_g[w][threadIdx.x]; b += s_b[w][threadIdx.x]; }
        partial[(long)blockIdx.x * 256 + threadIdx.x] = a;
        partial[(long)blockIdx.x * 256 + 128 + threadIdx.x] = b;
    }
}
__global__ void layernorm_backward_reduce(const float* __restrict__ partial, int n, float scale, float* __restrict__ dgamma,
                                          float* __restrict__ dbeta) {
    __shared__ float s_part[4][256];             // 1024 threads: 4 interleaved partial sums per column, fixed order
    const int c = threadIdx.x & 255, part = threadIdx.x >> 8;      // columns 0..127 gamma, 128..255 beta
    float s = 0.f;
    for (int k = part; k < n; k += 4) s += partial[(long)k * 256 + c];
    s_part[part][c] = s;
    __syncthreads();
    if (part) return;
    s = (s_part[0][c] + s_part[1][c]) + (s_part[2][c] + s_part[3][c]);
    if (c < 128) dgamma[c] += scale * s; else dbeta[c - 128] += scale * s;
}

// torch.optim.Adam(amsgrad=True) with L2 weight decay (config/train_nfs.yml:28-34), one step over flat buffers:
//   g = grad + wd * p ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; vmax = max(vmax, v)
//   p -= lr / (1 - b1^t) * m / (sqrt(vmax) / sqrt(1 - b2^t) + eps)
__global__ void adam_amsgrad(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ m,
                             float* __restrict__ v, float* __restrict__ vmax, long n, float lr, float b1, float b2,
                             float eps, float wd, float bc1, float bc2_sqrt) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float pi = p[i];
    const float g = grad[i] + wd * pi;
    const float mi = b1 * m[i] + (1.f - b1) * g;
    const float vi = b2 * v[i] + (1.f - b2) * g * g;
    const float vm = fmaxf(vmax[i], vi);
    m[i] = mi; v[i] = vi; vmax[i] = vm;
    const float denom = sqrtf(vm) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
}

}  // namespace
}  // namespace bmc

using namespace bmc;

extern "C" BMC_EXPORT size_t bmc_conv_wgrad_workspace_bytes(int n_split, int taps, int x_ch) {
    return (size_t)n_split * taps * 128 * x_ch * sizeof(float) + (size_t)256 * 128 * sizeof(float);
}

extern "C" BMC_EXPORT int bmc_conv_wgrad(const void* dy_act16, const void* x_act16, int x_ch, int taps, int B, int H, int W,
                                         const int* cmap, int cin_total, int n_out, float scale, float* grad_w, float* grad_b,
                                         void* workspace, size_t workspace_bytes, int n_split, void* stream) {
    BMC_REQUIRE(dy_act16 && x_act16 && cmap && grad_w && workspace, "conv_wgrad: NULL argument");
    BMC_REQUIRE(x_ch == 64 || x_ch == 128, "conv_wgrad: x_ch must be 64 or 128");
    BMC_REQUIRE(taps == 1 || taps == 9, "conv_wgrad: taps must be 1 or 9");
    BMC_REQUIRE(n_split >= 1 && n_split <= 256, "conv_wgrad: 1..256 splits");
    BMC_REQUIRE(n_out >= 1 && n_out <= 128, "conv_wgrad: 1..128 output channels");
    BMC_REQUIRE(workspace_bytes >= bmc_conv_wgrad_workspace_bytes(n_split, taps, x_ch), "conv_wgrad: workspace too small");
    const Geom g = Geom::make(B, H, W);
    BMC_REQUIRE(g.rows() < (1L << 31), "conv_wgrad: too many rows");
    cudaStream_t st = as_stream(stream);
    WgradParams p;
    memset(&p, 0, sizeof(p));
    int rc = make_tmap_2d_act(&p.map_dy, dy_act16, (uint64_t)g.rows(), 128, 64, 64);
    if (!rc) rc = make_tmap_2d_act(&p.map_x, x_act16, (uint64_t)g.rows(), (uint64_t)x_ch, 64, 64);
    if (rc) return rc;
    p.x_chunks = x_ch / 64; p.n_taps = taps;
    for (int t = 0; t < taps; ++t) p.tap_off[t] = taps == 9 ? (t / 3 - 1) * g.Wp + (t % 3 - 1) : 0;
    p.rows = (int)g.rows();
    const int chunks = p.rows / 64;
    p.rows_per_split = (chunks + n_split - 1) / n_split * 64;
    p.n_split = n_split;
    p.partial = static_cast<float*>(workspace);
    const int per_split = taps * 128 * x_ch;
    p.dy_ptr = static_cast<const act_t*>(dy_act16);
    p.bias_partial = grad_b ? p.partial + (size_t)n_split * per_split : nullptr;
    const int smem = kWgStages * (2 + p.x_chunks) * kWgBox + 1024;
    static PerDevice configured_dev;
    int& configured = configured_dev.cur();
    if (!configured) {
        BMC_CUDA(cudaFuncSetAttribute(wgrad_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgStages * 4 * kWgBox + 1024));
        BMC_CUDA(cudaFuncSetAttribute(wgrad_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgStages * 3 * kWgBox + 1024));
        configured = 1;
    }
    dim3 grid((unsigned)n_split, (unsigned)taps);
    if (x_ch == 128) wgrad_tc<128><<<grid, kWgThreads, smem, st>>>(p);
    else wgrad_tc<64><<<grid, kWgThreads, smem, st>>>(p);
    BMC_CUDA(cudaGetLastError());
    wgrad_reduce<<<(per_split + 255) / 256 + (grad_b ? 1 : 0), 256, 0, st>>>(p.partial, n_split, taps, x_ch, cmap, cin_total, n_out,
                                                                             scale, grad_w, p.bias_partial, grad_b);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

extern "C" BMC_EXPORT int bmc_relu_backward(const void* dy_act16, const void* y_act16, int64_t n_elems, void* dx_act16,
                                            void* stream) {
    BMC_REQUIRE(dy_act16 && y_act16 && dx_act16 && n_elems >= 0 && n_elems % 2 == 0, "relu_backward: bad argument");
    const long n2 = n_elems / 2;
    if (n2 == 0) return BMC_OK;
    relu_backward<<<(unsigned)((n2 + 255) / 256), 256, 0, as_stream(stream)>>>(
        static_cast<const uint32_t*>(dy_act16), static_cast<const uint32_t*>(y_act16), n2, static_cast<uint32_t*>(dx_act16));
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

extern "C" BMC_EXPORT int bmc_adam_amsgrad_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                                float* max_exp_avg_sq, int64_t n, int step, float lr, float beta1,
                                                float beta2, float eps, float weight_decay, void* stream) {
    BMC_REQUIRE(params && grads && exp_avg && exp_avg_sq && max_exp_avg_sq && n >= 0 && step >= 1, "adam_amsgrad_step: bad argument");
    if (n == 0) return BMC_OK;
    // bias corrections in double, like torch.optim.Adam's Python floats
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    adam_amsgrad<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, max_exp_avg_sq,
                                                                              (long)n, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

extern "C" BMC_EXPORT size_t bmc_layernorm_rows_backward_workspace_bytes(void) { return (size_t)296 * 256 * sizeof(float); }

extern "C" BMC_EXPORT int bmc_layernorm_rows_backward(const void* x_act16, const void* dy_act16, const float* gamma, float eps,
                                                      int64_t rows, void* dx_act16, float scale, float* grad_gamma,
                                                      float* grad_beta, void* workspace, size_t workspace_bytes, void* stream) {
    BMC_REQUIRE(x_act16 && dy_act16 && gamma && dx_act16 && grad_gamma && grad_beta && workspace && rows >= 0,
                "layernorm_rows_backward: bad argument");
    BMC_REQUIRE(workspace_bytes >= bmc_layernorm_rows_backward_workspace_bytes(), "layernorm_rows_backward: workspace too small");
    if (rows == 0) return BMC_OK;
    cudaStream_t st = as_stream(stream);
    long want = (rows + 7) / 8;
    const int grid = (int)(want < 296 ? want : 296);
    float* partial = static_cast<float*>(workspace);
    layernorm_backward<<<grid, kLnbThreads, 0, st>>>(static_cast<const act_t*>(x_act16), static_cast<const act_t*>(dy_act16), gamma,
                                                     eps, (long)rows, static_cast<act_t*>(dx_act16), partial);
    layernorm_backward_reduce<<<1, 1024, 0, st>>>(partial, grid, scale, grad_gamma, grad_beta);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

// Evaluation tail of the inference loop on the device (reference infer_BMCNet.py:77-87):
//   esr_cnt     = prediction, bicubic-resized to the ground truth's size when the two differ
//   bicubic_cnt = bicubic x-upsampling of the LR count frame to the ground-truth resolution
//   esr_mse, bicubic_mse = nn.MSELoss (mean squared error) of each against the ground truth
// The reference moves the prediction to the host every frame for this; here one kernel reads the
// three tensors once and leaves two sums of squares in device memory, so the loop needs no
// per-frame synchronisation.  HBM-bound: 4 B per ground-truth element + the (smaller) inputs.
//
// Bicubic: F.interpolate(mode='bicubic', align_corners=False) without antialiasing -- ATen's
// upsample_bicubic2d: src = scale*(dst+0.5)-0.5 with scale = in/out, index = floor(src)
// (guarded to in-1), t = clamp(src-index, 0, 1), Keys cubic convolution with A = -0.75 on the
// four neighbours index-1..index+2, each clamped to [0, in-1]; rows first, then columns.
#include "common.cuh"

namespace bmc {
namespace {

struct Cubic {
    int idx[4];
    float w[4];
};

__device__ __forceinline__ float cc1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cc2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__device__ __forceinline__ Cubic cubic_taps(int dst, int in_size, float scale) {
    Cubic c;
    const float src = __fsub_rn(__fmul_rn(scale, (float)dst + 0.5f), 0.5f);
    int i0 = (int)floorf(src);
    if (i0 > in_size - 1) i0 = in_size - 1;
    const float t = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
    c.w[0] = cc2(t + 1.f); c.w[1] = cc1(t);
    const float u = 1.f - t;
    c.w[2] = cc1(u); c.w[3] = cc2(u + 1.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) c.idx[j] = max(min(i0 + j - 1, in_size - 1), 0);
    return c;
}

__device__ __forceinline__ float bicubic_at(const float* __restrict__ plane, int in_w, const Cubic& cy, const Cubic& cx) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float* row = plane + (long)cy.idx[i] * in_w;
        float r = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) r += cx.w[j] * row[cx.idx[j]];
        acc += cy.w[i] * r;
    }
    return acc;
}

// planes = B*C; one thread per ground-truth element (grid-stride), block sums -> two double atomics
__global__ void __launch_bounds__(256) sr_metrics_kernel(const float* __restrict__ pred, int Hp, int Wp,
                                                         const float* __restrict__ inp, int H, int W,
                                                         const float* __restrict__ gt, int Hg, int Wg, long planes,
                                                         double* __restrict__ sums) {
    const float sy_p = (float)Hp / (float)Hg, sx_p = (float)Wp / (float)Wg;
    const float sy_i = (float)H / (float)Hg, sx_i = (float)W / (float)Wg;
    const bool same = Hp == Hg && Wp == Wg;            // the reference resizes the prediction only when the sizes differ
    const long total = planes * Hg * Wg;
    float e_esr = 0.f, e_bic = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wg);
        const long r = i / Wg;
        const int y = (int)(r % Hg);
        const long pl = r / Hg;
        const float g = gt[i];
        float esr;
        if (same) {
            esr = pred[i];
        } else {
            const Cubic cy = cubic_taps(y, Hp, sy_p), cx = cubic_taps(x, Wp, sx_p);
            esr = bicubic_at(pred + pl * Hp * Wp, Wp, cy, cx);
        }
        const Cubic cy = cubic_taps(y, H, sy_i), cx = cubic_taps(x, W, sx_i);
        const float bic = bicubic_at(inp + pl * H * W, W, cy, cx);
        e_esr += (esr - g) * (esr - g);
        e_bic += (bic - g) * (bic - g);
    }
    // warp then block reduction, one double atomic per block and metric
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        e_esr += __shfl_xor_sync(0xffffffffu, e_esr, o);
        e_bic += __shfl_xor_sync(0xffffffffu, e_bic, o);
    }
    __shared__ float s_e[8], s_b[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_e[warp] = e_esr; s_b[warp] = e_bic; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += (double)s_e[w]; b += (double)s_b[w]; }
        atomicAdd(&sums[0], a);
        atomicAdd(&sums[1], b);
    }
}

}  // namespace
}  // namespace bmc

using namespace bmc;

extern "C" BMC_EXPORT int bmc_sr_metrics(const float* pred, int B, int C, int Hp, int Wp, const float* inp, int H, int W,
                                         const float* gt, int Hg, int Wg, double* sums, void* stream) {
    BMC_REQUIRE(pred && inp && gt && sums, "sr_metrics: NULL argument");
    BMC_REQUIRE(B > 0 && C > 0 && Hp > 0 && Wp > 0 && H > 0 && W > 0 && Hg > 0 && Wg > 0, "sr_metrics: bad sizes");
    cudaStream_t st = as_stream(stream);
    BMC_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
    const long total = (long)B * C * Hg * Wg;
    long blocks = (total + 255) / 256;
    const long cap = (long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    sr_metrics_kernel<<<(unsigned)blocks, 256, 0, st>>>(pred, Hp, Wp, inp, H, W, gt, Hg, Wg, (long)B * C, sums);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

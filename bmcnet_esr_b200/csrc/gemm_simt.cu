// Plain SIMT kernels with the arithmetic of gemm_tc.cu, kept ON THE DEVICE as a cross-check for
// the tcgen05 path (bmc_model_set_debug_simt / impl=1 of the per-kernel entry points).  One
// thread per output element, fp32 accumulation in K order; never used by the product path.
#include "gemm.cuh"

namespace bmc {
namespace {

// one CTA per output row, one thread per output channel
__global__ void conv_gemm_simt(const __grid_constant__ GemmParams p) {
    __shared__ float red[128];
    const GemmJobDev& job = p.jobs[blockIdx.y];
    const int N = p.n;
    const long m = blockIdx.x;
    const int n = threadIdx.x;
    const int img = (int)(m / p.g.R);
    const int r_img = (int)(m - (long)img * p.g.R);
    int y, x;
    const bool valid = p.g.interior(r_img, y, x);
    float acc = 0.f;
    const int w_row = job.w_row_base + img * job.w_img_stride + n;
    int kchunk = 0;
    for (int s = 0; s < p.n_seg; ++s) {
        for (int t = 0; t < p.n_taps; ++t) {
            const long arow = job.a_row_base[s] + m + p.tap_off[t];
            const bool in = arow >= 0 && arow < job.a_rows[s];
            for (int c = 0; c < p.chunks[s]; ++c, ++kchunk) {
                if (!in || !valid) continue;
                const act_t* a = job.a_ptr[s] + arow * job.a_ld[s] + job.a_col_base[s] + c * kChunkK;
                const act_t* w = job.w_ptr + ((long)kchunk * job.w_rows + w_row) * kChunkK;
                for (int k = 0; k < kChunkK; ++k) acc += from_act(a[k]) * from_act(w[k]);
            }
        }
    }
    acc += job.bias ? job.bias[n] : 0.f;
    if (job.ln_gamma) {                      // uniform per launch: every thread takes this branch
        red[n] = acc;
        __syncthreads();
        float mu = 0.f;
        for (int k = 0; k < N; ++k) mu += red[k];
        mu /= (float)N;
        float var = 0.f;
        for (int k = 0; k < N; ++k) var += (red[k] - mu) * (red[k] - mu);
        var /= (float)N;
        acc = job.ln_gamma[n] * ((acc - mu) / sqrtf(var + job.ln_eps)) + job.ln_beta[n];
    }
    if (job.relu) acc = fmaxf(acc, 0.f);
    if (job.residual && valid) acc += from_act(job.residual[(job.res_row_base + m) * N + n]);
    if (!valid) acc = 0.f;
    if (job.out) job.out[(job.out_row_base + m) * N + n] = to_act(acc);
    if (job.out_f32) job.out_f32[(job.out_row_base + m) * N + n] = acc;
}

__global__ void att_simt(const __grid_constant__ AttParams p) {
    // one thread per (pair, b, split, c, c')
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)p.n_pairs * p.g.B * p.n_split * 128 * 128;
    if (idx >= total) return;
    const int c2 = (int)(idx & 127);
    const int c1 = (int)((idx >> 7) & 127);
    long rest = idx >> 14;
    const int split = (int)(rest % p.n_split); rest /= p.n_split;
    const int b = (int)(rest % p.g.B);
    const int pair = (int)(rest / p.g.B);
    const int pix0 = split * p.pix_per_split;
    const int pix1 = min(pix0 + p.pix_per_split, p.g.R);
    const act_t* c = p.c_ptr + (p.c_row_base[pair] + (long)b * p.g.R) * 128;
    const act_t* v = p.v_ptr + (p.v_row_base[pair] + (long)b * p.g.R) * 128;
    float acc = 0.f;
    for (int px = pix0; px < pix1; ++px)
        acc += from_act(c[(long)px * 128 + c1]) * from_act(v[(long)px * 128 + c2]);
    p.partial[idx] = acc * p.scale;
}

// Sum the split partials in a fixed order, row softmax (submodules.py:72-73), write bf16
// "dynamic weights" P[c][c'] in the chunk-major weight layout.  One warp per row.
__global__ void att_softmax(const SoftmaxParams p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int total_rows = p.n_pairs * p.B * 128;
    if (warp >= total_rows) return;
    const int c = warp & 127;
    const int b = (warp >> 7) % p.B;
    const int pair = (warp >> 7) / p.B;
    const float* src = p.partial + (((long)pair * p.B + b) * p.n_split * 128 + c) * 128 + lane * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < p.n_split; ++s) {
        const float4 t = *reinterpret_cast<const float4*>(src + (long)s * 128 * 128);
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    float mx = fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w));
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    a.x = expf(a.x - mx); a.y = expf(a.y - mx); a.z = expf(a.z - mx); a.w = expf(a.w - mx);
    float sum = a.x + a.y + a.z + a.w;
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    const int col = lane * 4;                      // c' ; K chunk = col / 64
    const long row = p.w_row_base[pair] + (long)b * p.w_img_stride + (col >> 6) * 128 + c;
    act_t* dst = p.w_base + row * kChunkK + (col & 63);
    uint2 o2;
    o2.x = pack_act2(a.x * inv, a.y * inv);
    o2.y = pack_act2(a.z * inv, a.w * inv);
    *reinterpret_cast<uint2*>(dst) = o2;
}

}  // namespace

int launch_conv_gemm_simt(const GemmParams& p, cudaStream_t st) {
    dim3 grid((unsigned)p.g.rows(), (unsigned)p.n_jobs);
    conv_gemm_simt<<<grid, p.n, 0, st>>>(p);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

int launch_att_simt(const AttParams& p, cudaStream_t st) {
    const long total = (long)p.n_pairs * p.g.B * p.n_split * 128 * 128;
    att_simt<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

int launch_att_softmax(const SoftmaxParams& p, cudaStream_t st) {
    const int rows = p.n_pairs * p.B * 128;
    att_softmax<<<(rows * 32 + 255) / 256, 256, 0, st>>>(p);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

}  // namespace bmc

// Layout / pointwise kernels around the GEMMs: fp32 NCHW <-> padded NHWC bf16, the per-step input
// tensor, channel LayerNorm, x4 PixelShuffle + bilinear reconstruction, weight repacking.
#include "gemm.cuh"

namespace bmc {
namespace {

// ---------------------------------------------------------------- LayerNorm over 128 channels
// submodules.py:127-139: mu, biased var, (x-mu)/sqrt(var+eps)*w+b.  One warp per row, 4 ch / lane.
__global__ void layernorm_rows(const act_t* __restrict__ in, const float* __restrict__ gamma,
                               const float* __restrict__ beta, float eps, long rows,
                               act_t* __restrict__ out) {
    const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const uint2 raw = *reinterpret_cast<const uint2*>(in + row * 128 + lane * 4);
    const float2 a = unpack_act2(raw.x), b = unpack_act2(raw.y);
    float s = a.x + a.y + b.x + b.y;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s * (1.f / 128.f);
    const float d0 = a.x - mu, d1 = a.y - mu, d2 = b.x - mu, d3 = b.y - mu;
    float v = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = 1.f / sqrtf(v * (1.f / 128.f) + eps);
    const float4 g = *reinterpret_cast<const float4*>(gamma + lane * 4);
    const float4 bt = *reinterpret_cast<const float4*>(beta + lane * 4);
    uint2 o2;
    o2.x = pack_act2(g.x * (d0 * rstd) + bt.x, g.y * (d1 * rstd) + bt.y);
    o2.y = pack_act2(g.z * (d2 * rstd) + bt.z, g.w * (d3 * rstd) + bt.w);
    *reinterpret_cast<uint2*>(out + row * 128 + lane * 4) = o2;
}

// ---------------------------------------------------------------- NCHW fp32 -> padded NHWC bf16
// thread = (pixel, 8-channel group); lanes run over pixels so the NCHW reads coalesce.
__global__ void pack_nchw(const float* __restrict__ src, Geom g, int C, act_t* __restrict__ dst,
                          int c_pad, int c_off) {
    const int HW = g.H * g.W;
    const int groups = (C + 7) / 8;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per_img = (long)HW * groups;
    if (idx >= per_img * g.B) return;
    const int b = (int)(idx / per_img);
    const long r = idx - (long)b * per_img;
    const int grp = (int)(r / HW);
    const int pix = (int)(r - (long)grp * HW);
    const int y = pix / g.W, x = pix - y * g.W;
    const long row = (long)b * g.R + (long)(y + 1) * g.Wp + (x + 1);
    const float* s = src + ((long)b * C + grp * 8) * HW + pix;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (grp * 8 + j < C) ? s[(long)j * HW] : 0.f;
    act_t* d = dst + row * c_pad + c_off + grp * 8;
    if (grp * 8 + 8 <= C && ((c_off & 7) == 0)) {
        *reinterpret_cast<uint4*>(d) = make_uint4(pack_act2(f[0], f[1]), pack_act2(f[2], f[3]),
                                                  pack_act2(f[4], f[5]), pack_act2(f[6], f[7]));
    } else {
        for (int j = 0; j < 8 && grp * 8 + j < C; ++j) d[j] = to_act(f[j]);
    }
}

__global__ void unpack_nchw(const act_t* __restrict__ src, Geom g, int C, int c_pad, int c_off,
                            float* __restrict__ dst) {
    const int HW = g.H * g.W;
    const int groups = (C + 7) / 8;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per_img = (long)HW * groups;
    if (idx >= per_img * g.B) return;
    const int b = (int)(idx / per_img);
    const long r = idx - (long)b * per_img;
    const int grp = (int)(r / HW);
    const int pix = (int)(r - (long)grp * HW);
    const int y = pix / g.W, x = pix - y * g.W;
    const long row = (long)b * g.R + (long)(y + 1) * g.Wp + (x + 1);
    const act_t* s = src + row * c_pad + c_off + grp * 8;
    float* d = dst + ((long)b * C + grp * 8) * HW + pix;
    for (int j = 0; j < 8 && grp * 8 + j < C; ++j) d[(long)j * HW] = from_act(s[j]);
}

__global__ void unpack_nchw_f32(const float* __restrict__ src, Geom g, int C, int c_pad, float* __restrict__ dst) {
    const int HW = g.H * g.W;
    const int groups = (C + 7) / 8;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per_img = (long)HW * groups;
    if (idx >= per_img * g.B) return;
    const int b = (int)(idx / per_img);
    const long r = idx - (long)b * per_img;
    const int grp = (int)(r / HW);
    const int pix = (int)(r - (long)grp * HW);
    const int y = pix / g.W, x = pix - y * g.W;
    const long row = (long)b * g.R + (long)(y + 1) * g.Wp + (x + 1);
    const float* s = src + row * c_pad + grp * 8;
    float* d = dst + ((long)b * C + grp * 8) * HW + pix;
    for (int j = 0; j < 8 && grp * 8 + j < C; ++j) d[(long)j * HW] = s[j];
}

// ---------------------------------------------------------------- per-step input tensor "MI"
// 64 bf16 channels per padded pixel:
//   0-2  f1 positive x3 | 3-5  f2 positive x3 | 6-8  f1 negative x3 | 9-11 f2 negative x3
//   (the `.repeat(1, 3, 1, 1)` planes of BMCNet.py:109-112 / BMCNet_plain.py:57-58)
//   12-43 the 32 feedback channels o: x_o itself when init, else pixel_unshuffle(x_o, 4)
//         (BMCNet.py:117, submodules.py:80-92: channel = c*16 + ry*4 + rx)
//   44-63 zero
__global__ void pack_inputs(PackInputsParams p) {
    const Geom g = p.g;
    const int HW = g.H * g.W;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)g.B * HW) return;
    const int b = (int)(idx / HW);
    const int pix = (int)(idx - (long)b * HW);
    const int y = pix / g.W, x = pix - y * g.W;
    act_t* d = p.mi + ((long)b * g.R + (long)(y + 1) * g.Wp + (x + 1)) * 64;
    const float* xb = p.x + b * p.xs[0] + y * p.xs[3] + x * p.xs[4];
    const float f1p = xb[0], f2p = xb[p.xs[2]];
    const float f1n = xb[p.xs[1]], f2n = xb[p.xs[1] + p.xs[2]];
    const uint32_t w1p = pack_act2(f1p, f1p), w2p = pack_act2(f2p, f2p);
    const uint32_t w1n = pack_act2(f1n, f1n), w2n = pack_act2(f2n, f2n);
    uint32_t* dw = reinterpret_cast<uint32_t*>(d);
    // (rows are 128-byte aligned: the six words leave as one 16-byte and one 8-byte store)
    *reinterpret_cast<uint4*>(dw) = make_uint4(w1p, pack_act2(f1p, f2p), w2p, w1n);
    *reinterpret_cast<uint2*>(dw + 4) = make_uint2(pack_act2(f1n, f2n), w2n);
    if (!p.x_o && p.init) {                      // device-resident recurrence, reset: o = 0
#pragma unroll
        for (int c = 6; c < 32; ++c) dw[c] = 0u;
    }
    if (p.x_o) {
        float o[32];
        if (p.init) {
            const float* s = p.x_o + (long)b * 32 * HW + pix;
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = s[(long)c * HW];
        } else {
            const int HW4 = 16 * HW, W4 = 4 * g.W;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int ry = 0; ry < 4; ++ry) {
                    const float4 v = *reinterpret_cast<const float4*>(
                        p.x_o + ((long)b * 2 + c) * HW4 + (long)(4 * y + ry) * W4 + 4 * x);
                    o[c * 16 + ry * 4 + 0] = v.x; o[c * 16 + ry * 4 + 1] = v.y;
                    o[c * 16 + ry * 4 + 2] = v.z; o[c * 16 + ry * 4 + 3] = v.w;
                }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) dw[6 + c] = pack_act2(o[2 * c], o[2 * c + 1]);
#pragma unroll
        for (int c = 22; c < 32; ++c) dw[c] = 0u;
    }
}

// ---------------------------------------------------------------- reconstruction
// x_o = pixel_shuffle(a, 4) + bilinear_x4(f2)  (BMCNet.py:119, BMCNet_plain.py:66), fp32.
// F.interpolate(mode='bilinear', align_corners=False, scale_factor=4): src = (dst+0.5)/4-0.5
// clamped at 0, neighbours clamped at the border.  One thread per LR pixel: 2x4x4 outputs.
// Also writes the next step's feedback channels: pixel_unshuffle of what was just produced.
__device__ __forceinline__ void lerp_taps(int d, int n, int& i0, int& i1, float& l1) {
    float s = (d + 0.5f) * 0.25f - 0.5f;
    if (s < 0.f) s = 0.f;
    i0 = (int)s;
    i1 = i0 + (i0 < n - 1 ? 1 : 0);
    l1 = s - (float)i0;
}
__device__ __forceinline__ float sel3(int k, float a, float b, float c) { return k == 0 ? a : (k == 1 ? b : c); }
__global__ void __launch_bounds__(128) emit_output(EmitParams p) {
    const Geom g = p.g;
    const int HW = g.H * g.W;
    // one thread per (LR pixel, output plane c): 16 of the 32 values -- half the registers and twice the threads of a
    // thread per pixel (128 -> 96 -> ~64 registers; the kernel is latency-bound on its loads)
    const long tidx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long idx = tidx >> 1;
    const int c = (int)(tidx & 1);
    if (idx >= (long)g.B * HW) return;
    const int b = (int)(idx / HW);
    const int pix = (int)(idx - (long)b * HW);
    const int y = pix / g.W, x = pix - y * g.W;
    const long row = (long)b * g.R + (long)(y + 1) * g.Wp + (x + 1);
    // the 4x4 outputs of an LR pixel interpolate inside its 3x3 neighbourhood of f2: 9 loads per plane;
    // tap indices / weights exactly as lerp_taps gives them (same arithmetic as one load per tap)
    int ky0[4], ky1[4], kx0[4], kx1[4];
    float ly[4], lx[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int i0, i1;
        lerp_taps(4 * y + r, g.H, i0, i1, ly[r]); ky0[r] = i0 - (y - 1); ky1[r] = i1 - (y - 1);
        lerp_taps(4 * x + r, g.W, i0, i1, lx[r]); kx0[r] = i0 - (x - 1); kx1[r] = i1 - (x - 1);
    }
    const int yc[3] = {max(y - 1, 0), y, min(y + 1, g.H - 1)}, xc[3] = {max(x - 1, 0), x, min(x + 1, g.W - 1)};
    const int W4 = 4 * g.W;
    {
        float o[16];
        {
            const float4* a4 = reinterpret_cast<const float4*>(p.a + row * 32 + c * 16);     // 64-byte aligned half row
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = a4[i];
                o[4 * i] = v.x; o[4 * i + 1] = v.y; o[4 * i + 2] = v.z; o[4 * i + 3] = v.w;
            }
        }
        const float* f2 = p.x + b * p.xs[0] + c * p.xs[1] + p.xs[2];
        float P[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int l = 0; l < 3; ++l) P[k][l] = f2[yc[k] * p.xs[3] + xc[l] * p.xs[4]];
        float Hx[3][4];                                    // x-interpolated rows
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int rx = 0; rx < 4; ++rx)
                Hx[k][rx] = (1.f - lx[rx]) * sel3(kx0[rx], P[k][0], P[k][1], P[k][2]) + lx[rx] * sel3(kx1[rx], P[k][0], P[k][1], P[k][2]);
#pragma unroll
        for (int ry = 0; ry < 4; ++ry)
#pragma unroll
            for (int rx = 0; rx < 4; ++rx) {
                const float up = (1.f - ly[ry]) * sel3(ky0[ry], Hx[0][rx], Hx[1][rx], Hx[2][rx]) +
                                 ly[ry] * sel3(ky1[ry], Hx[0][rx], Hx[1][rx], Hx[2][rx]);
                o[ry * 4 + rx] += up;
            }
        if (p.out_o) {
#pragma unroll
            for (int ry = 0; ry < 4; ++ry)
                *reinterpret_cast<float4*>(p.out_o + ((long)b * 2 + c) * 16 * HW + (long)(4 * y + ry) * W4 + 4 * x) =
                    make_float4(o[ry * 4], o[ry * 4 + 1], o[ry * 4 + 2], o[ry * 4 + 3]);
        }
        if (p.mi_next) {
            uint2* d2 = reinterpret_cast<uint2*>(p.mi_next + row * 64 + 12 + c * 16);   // channels 12 + 16 c ..: byte offset 24 + 32 c
#pragma unroll
            for (int q = 0; q < 4; ++q) d2[q] = make_uint2(pack_act2(o[4 * q], o[4 * q + 1]), pack_act2(o[4 * q + 2], o[4 * q + 3]));
        }
    }
}

// ---------------------------------------------------------------- weight repack
// fp32 conv weight [n_out][*] -> bf16 chunk-major [K/64][w_rows][64] at rows
// [w_row_base, w_row_base + n_out_pad); kmap[k] = flat offset inside one output-channel row of
// the source, or -1 for zero padding.
// `in_scale` (optional, 1x1 weights only: source offset == input channel): the weight of input channel i is
// multiplied by in_scale[i] -- a channel LayerNorm's gamma folded into the convolution that follows it.
__global__ void repack_weight(const float* __restrict__ src, const int* __restrict__ kmap, int src_row_len,
                              int n_out, int n_out_pad, int K, act_t* __restrict__ dst, int w_rows,
                              int w_row_base, const float* __restrict__ in_scale) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n_out_pad * K) return;
    const int n = (int)(idx / K), k = (int)(idx - (long)n * K);
    const int off = kmap[k];
    float v = (n < n_out && off >= 0) ? src[(long)n * src_row_len + off] : 0.f;
    if (in_scale && off >= 0) v *= in_scale[off];
    dst[((long)(k >> 6) * w_rows + w_row_base + n) * 64 + (k & 63)] = to_act(v);
}

// bias_out[o] = bias[o] + sum_i w[o][i] * beta[i]: the LayerNorm's beta pushed through a 1x1 convolution (fp32)
__global__ void fold_beta_bias(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ beta,
                               int n_out, int cin, float* __restrict__ bias_out) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    float s = bias[o];
    for (int i = 0; i < cin; ++i) s += w[(long)o * cin + i] * beta[i];
    bias_out[o] = s;
}

__global__ void fill_identity(act_t* __restrict__ dst) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // [2][128][64]
    if (idx >= 2 * 128 * 64) return;
    const int kc = idx >> 13, n = (idx >> 6) & 127, kk = idx & 63;
    dst[idx] = to_act(kc * 64 + kk == n ? 1.f : 0.f);
}

}  // namespace

int launch_fill_identity(act_t* dst, cudaStream_t st) {
    fill_identity<<<64, 256, 0, st>>>(dst);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

int launch_layernorm(const act_t* in, const float* gamma, const float* beta, float eps, long rows,
                     act_t* out, cudaStream_t st) {
    if (rows <= 0) return BMC_OK;
    layernorm_rows<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(in, gamma, beta, eps, rows, out);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

int launch_pack_nchw(const float* src, Geom g, int C, act_t* dst, int c_pad, int c_off, cudaStream_t st) {
    const long total = (long)g.B * g.H * g.W * ((C + 7) / 8);
    pack_nchw<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, g, C, dst, c_pad, c_off);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

int launch_unpack_nchw(const act_t* src, Geom g, int C, int c_pad, int c_off, float* dst, cudaStream_t st) {
    const long total = (long)g.B * g.H * g.W * ((C + 7) / 8);
    unpack_nchw<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, g, C, c_pad, c_off, dst);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

int launch_unpack_nchw_f32(const float* src, Geom g, int C, int c_pad, float* dst, cudaStream_t st) {
    const long total = (long)g.B * g.H * g.W * ((C + 7) / 8);
    unpack_nchw_f32<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, g, C, c_pad, dst);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

int launch_pack_inputs(const PackInputsParams& p, cudaStream_t st) {
    const long total = (long)p.g.B * p.g.H * p.g.W;
    pack_inputs<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(p);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

int launch_emit(const EmitParams& p, cudaStream_t st) {
    const long total = 2L * p.g.B * p.g.H * p.g.W;          // (pixel, plane) pairs
    emit_output<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(p);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

int launch_repack_weight(const float* src, const int* kmap, int src_row_len, int n_out, int n_out_pad, int K,
                            act_t* dst, int w_rows, int w_row_base, cudaStream_t st, const float* in_scale) {
    const long total = (long)n_out_pad * K;
    repack_weight<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, kmap, src_row_len, n_out, n_out_pad, K, dst,
                                                                  w_rows, w_row_base, in_scale);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

int launch_fold_beta_bias(const float* w, const float* bias, const float* beta, int n_out, int cin, float* bias_out,
                          cudaStream_t st) {
    fold_beta_bias<<<(n_out + 127) / 128, 128, 0, st>>>(w, bias, beta, n_out, cin, bias_out);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

}  // namespace bmc

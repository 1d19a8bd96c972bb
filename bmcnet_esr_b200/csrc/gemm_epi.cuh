// Epilogue shared by the tcgen05 conv kernels: one thread owns one accumulator row (TMEM lane) and
// walks it in 32-column chunks: bias (+ fused channel LayerNorm) (+ ReLU) (+ residual), halo rows
// forced to zero, 16-bit and optional fp32 stores.
#pragma once
#include "gemm.cuh"

namespace bmc {

struct EpiRow {
    const act_t* res;      // residual row or NULL
    act_t* out;            // 16-bit output row or NULL
    float* outf;           // fp32 output row or NULL
    bool valid;            // interior pixel (else zeros are written)
    bool store;            // row exists (tail tiles of the slab kernel run past the tensor)
    bool relu;
    float ln_eps;
};

// Row statistics for the fused LayerNorm (submodules.py:127-139): two passes over TMEM.
template <int N>
__device__ __forceinline__ void epi_ln_stats(uint32_t trow, const float* bias_s, float eps, float& mu, float& rstd) {
    float s1 = 0.f;
#pragma unroll 1
    for (int c = 0; c < N / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(trow + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) s1 += __uint_as_float(v[j]) + bias_s[c * 32 + j];
    }
    mu = s1 * (1.f / N);
    float s2 = 0.f;
#pragma unroll 1
    for (int c = 0; c < N / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(trow + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float d = __uint_as_float(v[j]) + bias_s[c * 32 + j] - mu;
            s2 += d * d;
        }
    }
    rstd = 1.f / sqrtf(s2 * (1.f / N) + eps);
}

// One 32-column chunk whose accumulator values are already in v[].
__device__ __forceinline__ void epi_chunk(const uint32_t (&v)[32], int c, const EpiRow& r, const float* bias_s,
                                          bool ln, float mu, float rstd, const float* gamma_s, const float* beta_s) {
    float f[32];
    const float4* b4 = reinterpret_cast<const float4*>(bias_s + c * 32);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const float4 bb = b4[u];
        f[u * 4 + 0] = __uint_as_float(v[u * 4 + 0]) + bb.x;
        f[u * 4 + 1] = __uint_as_float(v[u * 4 + 1]) + bb.y;
        f[u * 4 + 2] = __uint_as_float(v[u * 4 + 2]) + bb.z;
        f[u * 4 + 3] = __uint_as_float(v[u * 4 + 3]) + bb.w;
    }
    if (ln) {
        const float4* g4 = reinterpret_cast<const float4*>(gamma_s + c * 32);
        const float4* t4 = reinterpret_cast<const float4*>(beta_s + c * 32);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 g = g4[u], t = t4[u];
            f[u * 4 + 0] = g.x * ((f[u * 4 + 0] - mu) * rstd) + t.x;
            f[u * 4 + 1] = g.y * ((f[u * 4 + 1] - mu) * rstd) + t.y;
            f[u * 4 + 2] = g.z * ((f[u * 4 + 2] - mu) * rstd) + t.z;
            f[u * 4 + 3] = g.w * ((f[u * 4 + 3] - mu) * rstd) + t.w;
        }
    }
    if (r.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (r.res && r.valid) {
        const uint4* rp = reinterpret_cast<const uint4*>(r.res + c * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint4 rv = rp[u];
            const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
                const float2 t2 = unpack_act2(w[k2]);
                f[u * 8 + k2 * 2] += t2.x;
                f[u * 8 + k2 * 2 + 1] += t2.y;
            }
        }
    }
    if (!r.valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = 0.f;
    }
    if (r.out && r.store) {
        uint4* op = reinterpret_cast<uint4*>(r.out + c * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            op[u] = make_uint4(pack_act2(f[u * 8], f[u * 8 + 1]), pack_act2(f[u * 8 + 2], f[u * 8 + 3]),
                               pack_act2(f[u * 8 + 4], f[u * 8 + 5]), pack_act2(f[u * 8 + 6], f[u * 8 + 7]));
    }
    if (r.outf && r.store) {
        float4* op = reinterpret_cast<float4*>(r.outf + c * 32);
#pragma unroll
        for (int u = 0; u < 8; ++u) op[u] = make_float4(f[u * 4], f[u * 4 + 1], f[u * 4 + 2], f[u * 4 + 3]);
    }
}

// ---------------------------------------------------------------------------------------------
// Coalesced variant (slab kernel): a thread owns one ROW, so storing its 64 bytes directly makes
// every warp-wide 16-byte store touch 32 different cache lines (measured: stores were 70 % of the
// epilogue).  Instead the warp's [32 rows x 32 cols] 16-bit block goes through a 2 KB shared-memory
// staging tile (16-byte chunks XOR-swizzled by (row>>1)&3: conflict-free both ways) and is written
// out with lane l covering chunk l%4 of row l/4 + 8i: every store instruction writes 8 rows x 64
// contiguous bytes.  The residual is read through the same tile with the same mapping.
__device__ __forceinline__ uint32_t stage_off(int row, int chunk) {
    return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

struct EpiTile {
    const act_t* res;      // residual tile base (row 0 of this warp's 32 rows, column 0) or NULL
    act_t* out;            // 16-bit output tile base or NULL
    float* outf;           // fp32 output row of THIS thread or NULL
    long rows_left;        // rows of this warp's block that exist (store predicate), may be <= 0 or >= 32
    bool valid;            // this thread's row is an interior pixel
    bool relu;
    int n;                 // row pitch in elements (128 or 32)
};

__device__ __forceinline__ void epi_chunk_staged(const uint32_t (&v)[32], int c, const EpiTile& r, const float* bias_s,
                                                 bool ln, float mu, float rstd, const float* gamma_s,
                                                 const float* beta_s, uint8_t* stage, int lane) {
    float f[32];
    const float4* b4 = reinterpret_cast<const float4*>(bias_s + c * 32);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const float4 bb = b4[u];
        f[u * 4 + 0] = __uint_as_float(v[u * 4 + 0]) + bb.x;
        f[u * 4 + 1] = __uint_as_float(v[u * 4 + 1]) + bb.y;
        f[u * 4 + 2] = __uint_as_float(v[u * 4 + 2]) + bb.z;
        f[u * 4 + 3] = __uint_as_float(v[u * 4 + 3]) + bb.w;
    }
    if (ln) {
        const float4* g4 = reinterpret_cast<const float4*>(gamma_s + c * 32);
        const float4* t4 = reinterpret_cast<const float4*>(beta_s + c * 32);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 g = g4[u], t = t4[u];
            f[u * 4 + 0] = g.x * ((f[u * 4 + 0] - mu) * rstd) + t.x;
            f[u * 4 + 1] = g.y * ((f[u * 4 + 1] - mu) * rstd) + t.y;
            f[u * 4 + 2] = g.z * ((f[u * 4 + 2] - mu) * rstd) + t.z;
            f[u * 4 + 3] = g.w * ((f[u * 4 + 3] - mu) * rstd) + t.w;
        }
    }
    if (r.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    const int crow = lane >> 2, cchunk = lane & 3;          // coalesced mapping: row crow + 8i, chunk cchunk
    if (r.res) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = crow + 8 * i;
            uint4 rv = make_uint4(0, 0, 0, 0);
            if (row < r.rows_left) rv = *reinterpret_cast<const uint4*>(r.res + (long)row * r.n + c * 32 + cchunk * 8);
            *reinterpret_cast<uint4*>(stage + stage_off(row, cchunk)) = rv;
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint4 rv = *reinterpret_cast<const uint4*>(stage + stage_off(lane, u));
            const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
                const float2 t2 = unpack_act2(w[k2]);
                f[u * 8 + k2 * 2] += t2.x;
                f[u * 8 + k2 * 2 + 1] += t2.y;
            }
        }
        __syncwarp();
    }
    if (!r.valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = 0.f;
    }
    if (r.out) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            *reinterpret_cast<uint4*>(stage + stage_off(lane, u)) =
                make_uint4(pack_act2(f[u * 8], f[u * 8 + 1]), pack_act2(f[u * 8 + 2], f[u * 8 + 3]),
                           pack_act2(f[u * 8 + 4], f[u * 8 + 5]), pack_act2(f[u * 8 + 6], f[u * 8 + 7]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = crow + 8 * i;
            const uint4 ov = *reinterpret_cast<const uint4*>(stage + stage_off(row, cchunk));
            if (row < r.rows_left) *reinterpret_cast<uint4*>(r.out + (long)row * r.n + c * 32 + cchunk * 8) = ov;
        }
        __syncwarp();
    }
    if (r.outf && lane < r.rows_left) {
        float4* op = reinterpret_cast<float4*>(r.outf + c * 32);
#pragma unroll
        for (int u = 0; u < 8; ++u) op[u] = make_float4(f[u * 4], f[u * 4 + 1], f[u * 4 + 2], f[u * 4 + 3]);
    }
}

}  // namespace bmc

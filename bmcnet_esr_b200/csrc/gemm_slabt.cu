// conv_slabt_tc: the slab convolution with the operands SWAPPED -- weights are the A operand
// (M = 128 output channels), the 256 pixels of the tile are the N dimension.
//
// Why: an SS-mode tcgen05.mma reads both operands from shared memory.  With M = N = 128 every K=16
// slice reads 8 KB in its 64 cycles, i.e. the full 128 B/clk of an SM's shared memory; the TMA traffic
// that refills the slab / weight rings (~40 B/clk at full rate) then has to steal cycles from the tensor
// core, which caps conv_slab_tc at ~70 % of the tensor peak whatever the pipeline depth (gemm_slab.cu,
// DESIGN.md).  One N = 256 instruction per slice reads 4 KB (weights) + 8 KB (pixels) in 128 cycles:
// 96 B/clk, leaving room for the refills, and a single issuing thread keeps the pipe full
// (tools/mma_bench.cu: 128.3 cycles per 128x256x16 MMA = 4087 MAC/clk/SM).
//
// Accumulator: TMEM lane = output channel, column = pixel of the tile (2 x 256 columns, double
// buffered).  The epilogue therefore owns a CHANNEL per thread and 32 consecutive pixels per
// tcgen05.ld.  Outputs behind a tensor map (the model's arena) leave through the TMA-store epilogue of
// conv_slab2_tc (fragment-layout reads, stmatrix.trans, one bulk store per [32 px][128 ch] chunk; mix launch at
// B=95: 217 -> 209 us); otherwise lane pairs swap halves so that every lane stores two adjacent channels of one
// pixel (a warp store covers two 64-byte runs).
// Supported: N = 128 outputs, 16-bit output only, no epilogue residual / LayerNorm / fp32 side output
// (the fused plan needs none of them: residuals are identity K segments); everything else falls back
// to conv_slab_tc.  K steps, slab views, weight stages and the tile schedule are those of gemm_slab.cu.
#include "gemm_epi.cuh"

namespace bmc {
namespace {

constexpr int kThreadsT = 352;      // warps 0-7 epilogue, 8 slab TMA, 9 weight TMA, 10 MMA
constexpr int kBM = 256;
constexpr int kBoxRows = 64;
constexpr int kBoxBytes = kBoxRows * kChunkK * 2;     // 8192
constexpr int kSlabStages = 2;
constexpr int kWGroup = 2;
constexpr int kWStages = 3;
constexpr int kMaxASteps = 8;
constexpr int kN = 128;                               // output channels
constexpr int kWBytes = kN * kChunkK * 2;             // one tap of one K chunk: 16 KB
constexpr int kWStageBytes = kWGroup * kWBytes;

template <bool TMA_OUT>     // outputs behind a tensor map: TMA-store epilogue (as conv_slab2_tc), else 4-byte stores from registers
__global__ void __launch_bounds__(kThreadsT, 1) conv_slabt_tc(const __grid_constant__ GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int slab_bytes = p.slab_boxes * kBoxBytes;
    uint8_t* smem_w = smem + kSlabStages * slab_bytes;
    uint8_t* smem_stage = smem_w + kWStages * kWStageBytes;      // 2 halves x 8 KB: [32 px][128 ch] chunks for the TMA store

    __shared__ uint64_t a_full[kSlabStages], a_empty[kSlabStages], w_full[kWStages], w_empty[kWStages];
    __shared__ uint64_t acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[2][kN];          // [128-pixel half][channel], rewritten per tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long rows_total = p.g.rows();
    const int n_taps = p.n_taps;
    const int tap1_mask = p.tap1_mask, pimg_mask = p.pimg_mask;
    int a_steps = 0;
    int as_seg[kMaxASteps], as_chunk[kMaxASteps], as_k0[kMaxASteps], as_cs[kMaxASteps];
    uint32_t t1_as = 0, pm_as = 0;                      // per (segment, chunk) step: centre-tap only / per-image weights
    {
        int seg_chunk0 = 0;
        for (int s = 0; s < p.n_seg; ++s) {
            const bool t1 = (tap1_mask >> s) & 1;
            for (int c = 0; c < p.chunks[s]; ++c, ++a_steps) {
                as_seg[a_steps] = s; as_chunk[a_steps] = c; as_k0[a_steps] = seg_chunk0 + c; as_cs[a_steps] = p.chunks[s];
                if (t1 && n_taps != 1) t1_as |= 1u << a_steps;
                if ((pimg_mask >> s) & 1) pm_as |= 1u << a_steps;
            }
            if (!t1) seg_chunk0 += n_taps * p.chunks[s];
        }
    }
    // tile schedule: full 256-row tiles round-robin (no per-image mode here)
    const int G = gridDim.x, cta = blockIdx.x;
    const int n_mine = p.n_full > cta ? (p.n_full - cta + G - 1) / G : 0;
    const int per_job = p.n_full / p.n_jobs;
    auto tile_at = [&](int li, int& job, long& m0) {
        const int t = cta + li * G;
        job = t / per_job;
        m0 = (long)(t - job * per_job) * kBM;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlabStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < kWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
        mbar_fence_init();
    }
    if (warp == 10) tmem_alloc(&tmem_base_s, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    pdl_wait();                  // the prologue above overlaps the previous kernel's tail; no global access before this point
    pdl_launch_dependents();     // the next kernel may take over SMs as CTAs of this one exit (it waits for this grid itself)

    if (warp == 8) {
        // ------------------------------------------------------------ activation slabs (lane b issues box b)
        int it = 0;
        for (int li = 0; li < n_mine; ++li) {
            long m0l; int ji;
            tile_at(li, ji, m0l);
            const GemmJobDev& job = p.jobs[ji];
            const int m0 = (int)m0l;
            for (int as = 0; as < a_steps; ++as, ++it) {
                const int s = as_seg[as];
                const int st = it % kSlabStages;
                if (it >= kSlabStages) mbar_wait(&a_empty[st], ((it / kSlabStages) - 1) & 1);
                uint8_t* dst = smem + st * slab_bytes;
                const CUtensorMap* map = &p.maps[job.a_map64[s]];
                const bool t1 = (tap1_mask >> s) & 1;    // centre tap only: just the tile's own 256 rows
                const int row0 = job.a_row_base[s] + m0 - (t1 ? 0 : p.slab_lead);
                const int abox = p.abox_rows;
                const int boxes = ((t1 ? kBM : p.slab_boxes * kBoxRows) + abox - 1) / abox;
                if (lane == 0) mbar_expect_tx(&a_full[st], boxes * abox * 128);
                __syncwarp();
                if (lane < boxes)
                    tma_load_2d(dst + lane * abox * 128, map, &a_full[st], job.a_col_base[s] + as_chunk[as] * kChunkK, row0 + lane * abox);
            }
        }
    } else if (warp == 9) {
        // ------------------------------------------------------------ weight tiles (as in gemm_slab.cu: two K steps per
        // stage; a per-image step alone in its stage with one copy per 128-pixel half)
        int gi = 0;
        for (int li = 0; li < n_mine; ++li) {
            long m0l; int ji;
            tile_at(li, ji, m0l);
            const GemmJobDev& job = p.jobs[ji];
            const CUtensorMap* map = &p.maps[job.w_map];
            const int img_h0 = min((int)(m0l / p.g.R), p.g.B - 1), img_h1 = min((int)((m0l + 128) / p.g.R), p.g.B - 1);
            int as = 0, tap = 0;
            while (as < a_steps) {
                const int st = gi % kWStages;
                if (gi >= kWStages) mbar_wait(&w_empty[st], ((gi / kWStages) - 1) & 1);
                uint8_t* dst = smem_w + st * kWStageBytes + (lane & 1) * kWBytes;
                if ((pm_as >> as) & 1u) {
                    const int sg = as_seg[as];
                    if (lane == 0) mbar_expect_tx(&w_full[st], 2 * kWBytes);
                    __syncwarp();
                    if (lane < 2)
                        tma_load_2d(dst, &p.maps[job.t1_map[sg]], &w_full[st], 0,
                                    job.t1_row[sg] + as_chunk[as] * 128 + (lane ? img_h1 : img_h0) * job.t1_img_stride[sg]);
                    ++as;
                } else {
                    int as2 = as, tap2 = tap + 1;
                    if (tap2 == (((t1_as >> as) & 1u) ? 1 : n_taps)) { tap2 = 0; ++as2; }
                    const bool two = as2 < a_steps && !((pm_as >> as2) & 1u);
                    if (lane == 0) mbar_expect_tx(&w_full[st], (two ? 2 : 1) * kWBytes);
                    __syncwarp();
                    if (lane < (two ? 2 : 1)) {
                        const int asl = lane ? as2 : as, tapl = lane ? tap2 : tap;
                        if ((t1_as >> asl) & 1u) {
                            const int sj = as_seg[asl];
                            tma_load_2d(dst, &p.maps[job.t1_map[sj]], &w_full[st], 0, job.t1_row[sj] + as_chunk[asl] * 128);
                        } else {
                            const int kchunk = as_k0[asl] + tapl * as_cs[asl];
                            tma_load_2d(dst, map, &w_full[st], 0, kchunk * job.w_rows + job.w_row_base);
                        }
                    }
                    as = as2; tap = tap2;
                    if (two) { if (++tap == (((t1_as >> as) & 1u) ? 1 : n_taps)) { tap = 0; ++as; } }
                }
                ++gi;
            }
        }
    } else if (warp == 10) {
        // ------------------------------------------------------------ MMA issuer: A = weight tile, B = tap view of the slab
        if (lane == 0) {
            constexpr uint32_t idesc256 = umma_idesc_f16(128, 256, false, false);
            constexpr uint32_t idesc128 = umma_idesc_f16(128, 128, false, false);
            constexpr uint32_t hi = umma_desc_hi_sw128(1024);
            const uint32_t tap0_lo = (uint32_t)(p.slab_lead + p.tap_off[0]) * 8u;     // = 0: tap (dy,dx)=(-1,-1) starts at slab row 0
            const uint32_t row_wrap = (uint32_t)(p.g.Wp - 2) * 8u;
            const uint32_t slab_lo0 = umma_desc_lo(smem_u32(smem), 16);
            const uint32_t w_lo0 = umma_desc_lo(smem_u32(smem_w), 16);
            const uint32_t slab_step = (uint32_t)slab_bytes >> 4;
            int ia = 0, gi = 0, lt = 0;
            for (int li = 0; li < n_mine; ++li, ++lt) {
                const int buf = lt & 1;
                if (lt >= 2) mbar_wait(&acc_empty[buf], ((lt >> 1) - 1) & 1);
                tc_fence_after_sync();
                const uint32_t acc = tmem_base + buf * 256;
                uint32_t accumulate = 0;
                int tap = 0, dx = 0, sa = ia % kSlabStages, as = 0, cur_taps = (t1_as & 1u) ? 1 : n_taps;
                uint32_t slab_lo = slab_lo0 + sa * slab_step, tap_lo = tap0_lo;
                while (as < a_steps) {
                    const int sw = gi % kWStages;
                    mbar_wait(&w_full[sw], (gi / kWStages) & 1);
                    const uint32_t w_base = w_lo0 + sw * (kWStageBytes >> 4);
                    const bool pm = (pm_as >> as) & 1u;
                    for (int e = 0; e < 2; ++e) {
                        if (tap == 0) mbar_wait(&a_full[sa], (ia / kSlabStages) & 1);
                        tc_fence_after_sync();
                        const uint32_t b_lo = slab_lo + tap_lo;                    // pixels
                        if (pm) {
                            // per-image weights: the two 128-pixel halves may be different images
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint32_t a_lo = w_base + h * (kWBytes >> 4);
                                const uint32_t bh = b_lo + h * 128 * 8;
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    umma_f16(acc + h * 128, umma_desc(a_lo + ks * 2, hi), umma_desc(bh + ks * 2, hi), idesc128, ks ? 1u : accumulate);
                            }
                        } else {
                            const uint32_t a_lo = w_base + e * (kWBytes >> 4);
                            umma_f16(acc, umma_desc(a_lo, hi), umma_desc(b_lo, hi), idesc256, accumulate);
                            umma_f16(acc, umma_desc(a_lo + 2, hi), umma_desc(b_lo + 2, hi), idesc256, 1u);
                            umma_f16(acc, umma_desc(a_lo + 4, hi), umma_desc(b_lo + 4, hi), idesc256, 1u);
                            umma_f16(acc, umma_desc(a_lo + 6, hi), umma_desc(b_lo + 6, hi), idesc256, 1u);
                        }
                        accumulate = 1u;
                        if (++tap == cur_taps) {           // slab fully consumed
                            umma_commit(&a_empty[sa]);
                            tap = 0; dx = 0; tap_lo = tap0_lo; ++ia; ++as;
                            cur_taps = ((t1_as >> as) & 1u) ? 1 : n_taps;
                            sa = ia % kSlabStages;
                            slab_lo = slab_lo0 + sa * slab_step;
                        } else if (++dx == 3) {
                            dx = 0; tap_lo += row_wrap;
                        } else {
                            tap_lo += 8u;
                        }
                        if (pm || as >= a_steps || ((pm_as >> as) & 1u)) break;
                    }
                    umma_commit(&w_empty[sw]);
                    ++gi;
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 0..7)
        // warp: TMEM lane quarter q = channels [32 q, 32 q + 32), pixel half ph = columns [128 ph, 128 ph + 128)
        const int q = warp & 3, ph = warp >> 2;
        const int ch = q * 32 + lane;
        uint8_t* stage = smem_stage + warp * 2048;         // [32 pixels][32 channels] fp16
        int lt = 0;
        for (int li = 0; li < n_mine; ++li, ++lt) {
            const int buf = lt & 1;
            long m0l; int ji;
            tile_at(li, ji, m0l);
            const GemmJobDev& job = p.jobs[ji];
            {
                const int t = threadIdx.x;                // 0..255
                asm volatile("bar.sync 1, 256;" ::: "memory");      // every warp is done with the previous tile's vectors
                const int hb = t >> 7, n = t & 127;
                const int img_h = min((int)((m0l + hb * 128) / p.g.R), p.g.B - 1);
                bias_s[hb][n] = (job.bias ? job.bias[n] : 0.f) + (job.bias_img ? job.bias_img[img_h * kN + n] : 0.f);
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const float bias_c = bias_s[ph][ch];
            const bool relu = job.relu != 0;
            mbar_wait(&acc_full[buf], (lt >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t trow = tmem_base + buf * 256 + ph * 128 + ((uint32_t)(q * 32) << 16);
            if (TMA_OUT) {
                // The four warps of a 128-pixel half assemble each [32 px][128 ch] chunk in shared memory (fragment-layout
                // TMEM reads, stmatrix.trans, SWIZZLE_128B) and one thread stores it as two [32 x 64] boxes.  One 8 KB
                // buffer per half: the next chunk waits until the store has read it (the accumulators are double
                // buffered, so this epilogue only has to keep up with the MMAs of the next tile).
                const long px_base = m0l + ph * 128;
                const int tr = lane >> 2, tc2 = (lane & 3) * 2;
                float bia[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) bia[g] = bias_s[ph][q * 32 + g * 8 + tr];
                const bool store_half = px_base < rows_total;               // rows_total is a multiple of 128
                const bool leader = q == 0 && lane == 0;
                const CUtensorMap* omap = &p.maps32[job.out_map32];
                const int orow = job.out_map_row + (int)(job.out_row_base + px_base);
                uint8_t* sbuf = smem_stage + ph * 8192;
                uint8_t* box = sbuf + (q >> 1) * 4096;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t f[2][16];
                    tmem_ld_16x256b_x4(trow + c * 32, f[0]);
                    tmem_ld_16x256b_x4(trow + c * 32 + (16u << 16), f[1]);
                    tmem_ld_wait();
                    if (c == 3) {
                        tc_fence_before_sync();
                        if (lane == 0) mbar_arrive(&acc_empty[buf]);
                    }
                    unsigned valid_mask;
                    {
                        const long m = px_base + c * 32 + lane;
                        const int img = (int)(m / p.g.R);
                        const int r_img = (int)(m - (long)img * p.g.R);
                        int y, x;
                        valid_mask = __ballot_sync(0xffffffffu, m < rows_total && p.g.interior(r_img, y, x));
                    }
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t plo[4], phi[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int col = 8 * k + tc2;
                            const bool v0 = (valid_mask >> col) & 1u, v1 = (valid_mask >> (col + 1)) & 1u;
                            float a0 = __uint_as_float(f[hf][4 * k]) + bia[2 * hf], a1 = __uint_as_float(f[hf][4 * k + 1]) + bia[2 * hf];
                            float b0 = __uint_as_float(f[hf][4 * k + 2]) + bia[2 * hf + 1], b1 = __uint_as_float(f[hf][4 * k + 3]) + bia[2 * hf + 1];
                            if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); b0 = fmaxf(b0, 0.f); b1 = fmaxf(b1, 0.f); }
                            plo[k] = pack_act2(v0 ? a0 : 0.f, v1 ? a1 : 0.f);
                            phi[k] = pack_act2(v0 ? b0 : 0.f, v1 ? b1 : 0.f);
                        }
                        const int c16 = (q & 1) * 4 + 2 * hf;
                        stmatrix_x4_trans(box + lane * 128 + (((c16) ^ (lane & 7)) << 4), plo[0], plo[1], plo[2], plo[3]);
                        stmatrix_x4_trans(box + lane * 128 + (((c16 + 1) ^ (lane & 7)) << 4), phi[0], phi[1], phi[2], phi[3]);
                    }
                    fence_proxy_async_smem();
                    asm volatile("bar.sync %0, 128;" ::"r"(3 + ph) : "memory");
                    if (leader) {
                        if (store_half) {
                            tma_store_2d(omap, sbuf, 0, orow + c * 32);
                            tma_store_2d(omap, sbuf + 4096, 64, orow + c * 32);
                        }
                        tma_store_commit();
                        tma_store_wait_read<0>();
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(3 + ph) : "memory");
                }
                if (li + 1 == n_mine && leader) tma_store_wait_all();
                continue;
            }
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const long px0 = m0l + ph * 128 + c * 32;              // first pixel row of this chunk
                uint32_t v[32];
                tmem_ld_32x32(trow + c * 32, v);
                tmem_ld_wait();
                if (c == 3) {
                    // all TMEM reads of this accumulator are done: hand it back to the MMA warp
                    tc_fence_before_sync();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                // halo / tail rows are written as zeros: lane j evaluates pixel px0 + j
                unsigned valid_mask;
                {
                    const long m = px0 + lane;
                    const int img = (int)(m / p.g.R);
                    const int r_img = (int)(m - (long)img * p.g.R);
                    int y, x;
                    valid_mask = __ballot_sync(0xffffffffu, m < rows_total && p.g.interior(r_img, y, x));
                }
                // lane pairs swap halves: even lanes end up with channels (ch, ch+1) of pixel j, odd lanes with
                // (ch-1, ch) of pixel j+1; a warp store covers two 64-byte runs (no shared-memory staging)
                {
                    act_t* out = job.out + (job.out_row_base + px0 + (lane & 1)) * kN + (ch & ~1);
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float f0 = __uint_as_float(v[j]) + bias_c, f1 = __uint_as_float(v[j + 1]) + bias_c;
                        if (relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
                        if (!((valid_mask >> j) & 1u)) f0 = 0.f;
                        if (!((valid_mask >> (j + 1)) & 1u)) f1 = 0.f;
                        const uint32_t mine = pack_act2(f0, f1);
                        const uint32_t theirs = __shfl_xor_sync(0xffffffffu, mine, 1);
                        const uint32_t o = (lane & 1) ? __byte_perm(theirs, mine, 0x7632) : __byte_perm(mine, theirs, 0x5410);
                        if (px0 + j + (lane & 1) < rows_total) *reinterpret_cast<uint32_t*>(out + (long)j * kN) = o;
                    }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 10) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

int slabt_smem(int slab_boxes) { return kSlabStages * slab_boxes * kBoxBytes + kWStages * kWStageBytes + 8 * 2048 + 1024; }

}  // namespace

bool slabt_supported(const GemmParams& p) {
    static int enabled = -1;
    if (enabled < 0) enabled = measure_env("BMC_CONV_SLABT", 1);
    if (!enabled || p.n != kN || (p.n_taps != 9 && p.n_taps != 1) || (p.tap1_mask & 1)) return false;
    for (int j = 0; j < p.n_jobs; ++j) {
        const GemmJobDev& d = p.jobs[j];
        if (d.w_img_stride != 0 || d.residual || d.out_f32 || d.ln_gamma || !d.out) return false;
        for (int s = 0; s < p.n_seg; ++s)
            if (d.a_map64[s] < 0) return false;
    }
    const int lead = p.n_taps == 9 ? p.g.Wp + 1 : 0;
    const int boxes = (kBM + 2 * lead + kBoxRows - 1) / kBoxRows;
    return slabt_smem(boxes) + 2048 <= 227 * 1024;
}

int launch_conv_slabt(GemmParams p, cudaStream_t st) {
    p.slab_lead = p.n_taps == 9 ? p.g.Wp + 1 : 0;
    p.slab_boxes = (kBM + 2 * p.slab_lead + kBoxRows - 1) / kBoxRows;
    if (p.abox_rows <= 0) p.abox_rows = kBoxRows;
    p.per_image = 0;
    p.pimg_mask = 0;
    for (int sg = 0; sg < p.n_seg; ++sg)
        if (((p.tap1_mask >> sg) & 1) && p.jobs[0].t1_img_stride[sg] != 0) p.pimg_mask |= 1 << sg;
    p.n_full = p.n_jobs * (int)((p.g.rows() + kBM - 1) / kBM);
    p.n_half = 0;
    const int smem = slabt_smem(p.slab_boxes);
    static PerDevice configured_dev;
    int& configured = configured_dev.cur();
    if (configured < smem) {
        BMC_CUDA(cudaFuncSetAttribute(conv_slabt_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        BMC_CUDA(cudaFuncSetAttribute(conv_slabt_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    bool tma_out = true;
    for (int j = 0; j < p.n_jobs; ++j) tma_out = tma_out && p.jobs[j].out_map32 >= 0;
    {
        static int direct = -1;        // BMC_SLABT_DIRECT=1: force the register epilogue (measurement)
        if (direct < 0) direct = measure_env("BMC_SLABT_DIRECT", 0);
        if (direct) tma_out = false;
    }
    int grid = p.n_full < sm_count() ? p.n_full : sm_count();
    {   // measurement switch: fewer CTAs -> is a tile's time set by the SM or by the shared L2 fabric?
        static int cap = -1;
        if (cap < 0) cap = measure_env("BMC_SLABT_GRID", 0);
        if (cap > 0 && grid > cap) grid = cap;
    }
    if (tma_out) BMC_CUDA(launch_pdl(conv_slabt_tc<true>, dim3(grid), dim3(kThreadsT), (size_t)smem, st, p));
    else BMC_CUDA(launch_pdl(conv_slabt_tc<false>, dim3(grid), dim3(kThreadsT), (size_t)smem, st, p));
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

}  // namespace bmc

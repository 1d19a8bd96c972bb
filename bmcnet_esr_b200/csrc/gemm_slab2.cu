// conv_slab2_tc: the slab convolution with every weight tile shared by TWO pixel tiles.
//
// conv_slabt_tc (gemm_slabt.cu) streams the whole weight set once per 256-pixel tile: 288 KB of
// weights + 112 KB of activation slab per 37.7 M MAC.  Measured on B200, that kernel's tile time
// falls from 8.9 us with 148 CTAs to 6.9 us with 16 (tools/time_conv.py, BMC_SLABT_GRID): it is bound
// first by the chip-wide L2 -> SM fabric (the ncu capture shows 5.3 KB/clk of xbar traffic, the
// ceiling is ~6.3) and then by shared-memory bandwidth inside the SM (MMA operand reads 96 B/clk +
// TMA fills 43 B/clk + epilogue staging 14 B/clk > 128 B/clk).  Both call for fewer bytes per MAC:
//
//  * K steps are 32 channels wide (64-byte rows, SWIZZLE_64B) instead of 64.  A slab is then 27 KB
//    and a weight tap 8 KB, small enough to keep TWO tiles' slabs resident and prefetched.
//  * Two tiles are in flight, one per 256-column TMEM accumulator, half a tile apart: in every
//    phase (one K step = segment x 32 channels) each weight tap is loaded once and multiplied into
//    BOTH accumulators.  Tile j runs phases [j P/2, j P/2 + P) of an endless cycle over the P steps
//    of the launch (K order is irrelevant to the sum), so one tile completes every P/2 phases and
//    its epilogue overlaps the MMAs of the next phases, as before.
//    L2 -> SM bytes per tile: 144 KB weights + 111 KB slabs = 255 KB (was 400 KB).
//  * The epilogue drains its accumulator into registers first (the accumulator is needed again at
//    once, by the tile after next), in the mma-fragment layout (tcgen05.ld.16x256b); the four warps of a
//    128-pixel half assemble [32 px][128 ch] chunks in shared memory with stmatrix.trans (SWIZZLE_128B,
//    conflict-free) and one thread TMA-stores each chunk: 256 contiguous bytes per pixel instead of four
//    64-byte runs from four warps, no per-thread global stores (3x3 launch at B=95: 162 -> 156.5 us).
//
// Operand roles as in conv_slabt_tc: A = weight tile (M = 128 output channels), B = 256-pixel tap
// view of the slab (N = 256), TMEM lane = output channel, column = pixel.  Same feature set: N = 128
// outputs, 16-bit output, bias (+ per-image bias), ReLU, identity / per-image-mix centre-tap segments.
// Tap views are row-shifted descriptor start addresses; tools/mma_sw64.cu verifies on B200 that this
// holds for SWIZZLE_64B operands as it does for SWIZZLE_128B.
#include "gemm.cuh"

namespace bmc {
namespace {

constexpr int kThreads2 = 384;      // warps 0-7 epilogue, 8 slab TMA, 9 weight TMA, 10 / 11 MMA issuers (one per accumulator)
constexpr int kBM = 256;
constexpr int kSubK = 32;                             // channels per K step
constexpr int kRowBytes = kSubK * 2;                  // 64
constexpr int kN = 128;                               // output channels
constexpr int kWBox = kN * kRowBytes;                 // one tap of one K step: 8 KB
constexpr int kWGroup = 3;                            // taps per weight slot
constexpr int kWSlot = kWGroup * kWBox;               // 24 KB
constexpr int kWSlots = 3;
constexpr int kSlabStages = 4;                        // two per accumulator / MMA issuer
constexpr int kMaxSteps = 16;

// two jobs read the same static weights (then one weight slot serves both tiles in flight)
__device__ __forceinline__ bool same_weights(const GemmJobDev& a, const GemmJobDev& b, int n_seg) {
    bool same = a.w_map32 == b.w_map32 && a.w_row_base == b.w_row_base && a.w_rows == b.w_rows;
    for (int s = 0; s < n_seg; ++s) same = same && a.t1_map32[s] == b.t1_map32[s] && a.t1_row[s] == b.t1_row[s];
    return same;
}

// The walk over the phases of a CTA, shared by the two TMA producers and the two MMA issuers so that all
// four derive the same ring positions.  Phase g multiplies K step (g mod P) into the older tile (li = g/PH - 1)
// and the newer tile (li = g/PH).  Weight slots of a phase, in ring order: for every group of <= 3 taps,
// one slot for both tiles when they share static weights, else the older tile's slot then the newer's;
// a per-image step has one slot per tile (its two 128-pixel halves).
// Barrier discipline: a parity wait is only safe for a thread that has observed every earlier phase of that
// barrier (a thread two fills behind or ahead sees the same parity).  So BOTH MMA threads wait on every
// weight fill and arrive on its release, whether or not they read it, and the slab ring is split in two
// private rings of two stages, one per accumulator: slab k of the tiles with parity X is phase X PH + k.
struct Phase {
    int g, step, jn, r;             // r = g mod PH
    bool has_o, has_n, shared, t1, pm;
    bool same_w;                    // both tiles in flight read the same static weights (changes with jn only)
    int job_o, job_n;
    long m_o, m_n;
    int groups, slots_per_group, n_act;
    int iw_slot, iw_par;            // weight ring position of the phase's first slot
};

// DBG selects the epilogue / measurement build (BMC_SLAB2_DBG): 32 = TMA-store epilogue (the product path whenever the
// outputs sit behind a tensor map), 0 = register epilogue with 4-byte stores (fallback; forced by BMC_SLAB2_DBG=64),
// 8 = stmatrix + 16-byte stores, 1 = no operand movement (MMAs on stale shared memory), 2 = no epilogue stores, 4 = store probe
template <int DBG>
__global__ void __launch_bounds__(kThreads2, 1) conv_slab2_tc(const __grid_constant__ GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    constexpr int S = kSlabStages;
    const int n_boxes = p.slab_boxes;
    const int slab_bytes = n_boxes * p.abox32_rows * kRowBytes;
    uint8_t* smem_w = smem + S * slab_bytes;
    uint8_t* smem_stage = smem_w + kWSlots * kWSlot;               // 32 KB: stmatrix epilogues (8 warps x 2 KB, or [half][buffer][8 KB])

    __shared__ uint64_t a_full[S], a_empty[S], w_full[kWSlots], w_empty[kWSlots];
    __shared__ uint64_t acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long rows_total = p.g.rows();
    const int n_taps = p.n_taps;
    const auto& sp = p.s2;
    const int P = sp.n, PH = P >> 1;

    const int G = gridDim.x, cta = blockIdx.x;
    const int n_mine = p.n_full > cta ? (p.n_full - cta + G - 1) / G : 0;
    const int per_job = p.n_full / p.n_jobs;
    auto tile_at = [&](int li, int& job, long& m0) {
        const int t = cta + li * G;
        job = t / per_job;
        m0 = (long)(t - job * per_job) * kBM;
    };
    const int n_phases = (n_mine + 1) * PH;

    auto phase_first = [&](Phase& f) {
        f.g = 0; f.step = 0; f.jn = 0; f.r = 0; f.iw_slot = 0; f.iw_par = 0;
    };
    // tiles / jobs change only when a tile starts (r == 0): the MMA threads run this between the last MMA of
    // one phase and the first of the next, so per phase it is only a few bit tests
    auto phase_fill = [&](Phase& f) {
        if (f.r == 0) {
            f.has_o = f.jn >= 1 && f.jn - 1 < n_mine;
            f.has_n = f.jn < n_mine;
            f.job_o = f.job_n = -1; f.m_o = f.m_n = 0;
            if (f.has_o) tile_at(f.jn - 1, f.job_o, f.m_o);
            if (f.has_n) tile_at(f.jn, f.job_n, f.m_n);
            f.n_act = (int)f.has_o + (int)f.has_n;
            f.same_w = f.has_o && f.has_n && (f.job_o == f.job_n || same_weights(p.jobs[f.job_o], p.jobs[f.job_n], p.n_seg));
        }
        f.t1 = (sp.t1 >> f.step) & 1u; f.pm = (sp.pm >> f.step) & 1u;
        f.shared = f.same_w && !f.pm;
        f.groups = (f.t1 || n_taps == 1) ? 1 : 3;
        f.slots_per_group = f.shared ? 1 : f.n_act;
    };
    auto phase_next = [&](Phase& f) {
        f.iw_slot += f.groups * f.slots_per_group;
        while (f.iw_slot >= kWSlots) { f.iw_slot -= kWSlots; f.iw_par ^= 1; }
        ++f.g;
        if (++f.step == P) f.step = 0;
        if (++f.r == PH) { f.r = 0; ++f.jn; }
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < kWSlots; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 2); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
        mbar_fence_init();
    }
    if (warp == 10) tmem_alloc(&tmem_base_s, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    pdl_wait();                  // everything above overlaps the previous kernel's tail; no global access before this point
    pdl_launch_dependents();     // the next kernel may take over SMs as CTAs of this one exit (it waits for this grid itself)

    if (warp == 8) {
        // ------------------------------------------------------------ activation slabs: per phase the older tile's, then the newer's
        Phase f;
        phase_first(f);
        for (; f.g < ((DBG & 1) ? 0 : n_phases); phase_next(f)) {
            phase_fill(f);
            const int sg = sp.seg[f.step];
            for (int e = 0; e < 2; ++e) {
                if (!(e ? f.has_n : f.has_o)) continue;
                const GemmJobDev& job = p.jobs[e ? f.job_n : f.job_o];
                const int m0 = (int)(e ? f.m_n : f.m_o);
                const int X = (e ? f.jn : f.jn - 1) & 1, k = f.g - X * PH;      // k-th slab of accumulator X
                const int st = 2 * X + (k & 1);
                if (k >= 2) mbar_wait(&a_empty[st], ((k >> 1) - 1) & 1);
                uint8_t* dst = smem + st * slab_bytes;
                const CUtensorMap* map = &p.maps32[job.a_map32[sg]];
                const int row0 = job.a_row_base[sg] + m0 - (f.t1 ? 0 : p.slab_lead);
                const int abox = p.abox32_rows;
                const int boxes = f.t1 ? (kBM + abox - 1) / abox : n_boxes;
                if (lane == 0) mbar_expect_tx(&a_full[st], boxes * abox * kRowBytes);
                __syncwarp();
                if (lane < boxes)
                    tma_load_2d(dst + lane * abox * kRowBytes, map, &a_full[st], job.a_col_base[sg] + sp.col[f.step], row0 + lane * abox);
            }
        }
    } else if (warp == 9) {
        // ------------------------------------------------------------ weight slots (<= 3 boxes of 8 KB each, one lane per box)
        Phase f;
        phase_first(f);
        int iw = 0;
        for (; f.g < ((DBG & 1) ? 0 : n_phases); phase_next(f)) {
            phase_fill(f);
            const int sg = sp.seg[f.step];
            const int c0 = sp.col[f.step] & 63;               // column inside the 64-wide weight chunk
            int slot = f.iw_slot, par = f.iw_par;
            for (int grp = 0; grp < f.groups; ++grp) {
                for (int e = 0; e < 2; ++e) {
                    if (!(e ? f.has_n : f.has_o)) continue;
                    if (e == 1 && f.shared) continue;
                    const GemmJobDev& job = p.jobs[e ? f.job_n : f.job_o];
                    const long m0 = e ? f.m_n : f.m_o;
                    const int boxes = f.pm ? 2 : ((f.t1 || n_taps == 1) ? 1 : kWGroup);
                    if (iw >= kWSlots) mbar_wait(&w_empty[slot], par ^ 1);
                    if (lane == 0) mbar_expect_tx(&w_full[slot], boxes * kWBox);
                    __syncwarp();
                    if (lane < boxes) {
                        uint8_t* dst = smem_w + slot * kWSlot + lane * kWBox;
                        if (f.pm) {
                            const int img = min((int)((m0 + lane * 128) / p.g.R), p.g.B - 1);
                            tma_load_2d(dst, &p.maps32[job.t1_map32[sg]], &w_full[slot], c0,
                                        job.t1_row[sg] + sp.c64[f.step] * 128 + img * job.t1_img_stride[sg]);
                        } else if (f.t1) {
                            tma_load_2d(dst, &p.maps32[job.t1_map32[sg]], &w_full[slot], c0, job.t1_row[sg] + sp.c64[f.step] * 128);
                        } else {
                            const int tap = grp * kWGroup + lane;
                            tma_load_2d(dst, &p.maps32[job.w_map32], &w_full[slot], c0,
                                        (sp.kchunk0[f.step] + tap * sp.cs[f.step]) * job.w_rows + job.w_row_base);
                        }
                    }
                    ++iw;
                    if (++slot == kWSlots) { slot = 0; par ^= 1; }
                }
            }
        }
    } else if (warp >= 10) {
        // ------------------------------------------------------------ MMA issuers: thread X owns accumulator X, i.e. the tiles
        // with li & 1 == X.  Issuing a tcgen05 instruction costs its thread some 70-100 cycles, so one thread cannot
        // feed the pipe with the N = 256 MMAs of two tiles plus their barrier traffic; two can.
        if (lane == 0) {
            const int X = warp - 10;
            constexpr uint32_t idesc256 = umma_idesc_f16(128, 256, false, false);
            constexpr uint32_t idesc128 = umma_idesc_f16(128, 128, false, false);
            constexpr uint32_t hi = umma_desc_hi_sw64(512);
            const uint32_t row_wrap = (uint32_t)(p.g.Wp - 2) * 4u;          // 64-byte rows: 4 descriptor units per row
            const uint32_t slab_lo0 = umma_desc_lo(smem_u32(smem), 16);
            const uint32_t w_lo0 = umma_desc_lo(smem_u32(smem_w), 16);
            const uint32_t slab_step = (uint32_t)slab_bytes >> 4;
            const uint32_t acc = tmem_base + X * 256;
            Phase f;
            phase_first(f);
            for (; f.g < n_phases; phase_next(f)) {
                phase_fill(f);
                // my tile in this phase: the newer one if its index has my parity, else the older one
                const bool newer = ((f.jn & 1) == X);
                const bool active = newer ? f.has_n : f.has_o;
                const int e = newer ? 1 : 0;
                const int li = newer ? f.jn : f.jn - 1;
                const bool first = newer && f.r == 0;                          // first phase of my tile
                const int k = f.g - X * PH, st = 2 * X + (k & 1);              // my k-th slab
                if (active) {
                    if (first && li >= 2) mbar_wait(&acc_empty[X], ((li >> 1) - 1) & 1);      // drained by the epilogue
                    if (!(DBG & 1)) mbar_wait(&a_full[st], (k >> 1) & 1);
                }
                const uint32_t slab_lo = slab_lo0 + st * slab_step;
                int slot = f.iw_slot, wpar = f.iw_par;
                uint32_t tap_lo = 0, fresh = first ? 0u : 1u;                  // 0: the first MMAs of a tile overwrite
                int dx = 0;
                for (int grp = 0; grp < f.groups; ++grp) {
                    for (int ee = 0; ee < f.slots_per_group; ++ee) {
                        // a slot is read by both tiles (shared) or by the ee-th active tile in (older, newer) order
                        const bool mine = active && (f.shared || (f.has_o ? ee : 1) == e);
                        if (!(DBG & 1)) mbar_wait(&w_full[slot], wpar);
                        if (mine) {
                            tc_fence_after_sync();
                            const uint32_t w_lo = w_lo0 + slot * (kWSlot >> 4);
                            if (f.pm) {
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const uint32_t a_lo = w_lo + h * (kWBox >> 4), bh = slab_lo + h * 128 * 4;
                                    umma_f16(acc + h * 128, umma_desc(a_lo, hi), umma_desc(bh, hi), idesc128, fresh);
                                    umma_f16(acc + h * 128, umma_desc(a_lo + 2, hi), umma_desc(bh + 2, hi), idesc128, 1u);
                                }
                                fresh = 1u;
                            } else {
                                const int gt = (f.t1 || n_taps == 1) ? 1 : kWGroup;
                                for (int t = 0; t < gt; ++t) {
                                    const uint32_t a_lo = w_lo + t * (kWBox >> 4), b_lo = slab_lo + tap_lo;
                                    umma_f16(acc, umma_desc(a_lo, hi), umma_desc(b_lo, hi), idesc256, fresh);
                                    umma_f16(acc, umma_desc(a_lo + 2, hi), umma_desc(b_lo + 2, hi), idesc256, 1u);
                                    fresh = 1u;
                                    if (++dx == 3) { dx = 0; tap_lo += row_wrap; } else tap_lo += 4u;
                                }
                            }
                            umma_commit(&w_empty[slot]);                       // released once my MMAs have read it
                        } else {
                            mbar_arrive(&w_empty[slot]);                       // not mine: seen, release at once
                        }
                        if (++slot == kWSlots) { slot = 0; wpar ^= 1; }
                    }
                }
                if (active) {
                    umma_commit(&a_empty[st]);
                    if (!newer && f.r == PH - 1) umma_commit(&acc_full[X]);   // my tile is complete
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 0..7)
        // warp: TMEM lane quarter q = channels [32 q, 32 q + 32), pixel half ph = columns [128 ph, 128 ph + 128)
        const int q = warp & 3, ph = warp >> 2;
        const int ch = q * 32 + lane;
        for (int li = 0; li < n_mine; ++li) {
            const int buf = li & 1;
            long m0l; int ji;
            tile_at(li, ji, m0l);
            const GemmJobDev& job = p.jobs[ji];
            const int img_h = min((int)((m0l + ph * 128) / p.g.R), p.g.B - 1);
            const float bias_c = (job.bias ? job.bias[ch] : 0.f) + (job.bias_img ? job.bias_img[img_h * kN + ch] : 0.f);
            const bool relu = job.relu != 0;
            const long px_base = m0l + ph * 128;
            // halo / tail rows are written as zeros: lane j evaluates pixel px_base + 32 c + j
            unsigned valid[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const long m = px_base + c * 32 + lane;
                const int img = (int)(m / p.g.R);
                const int r_img = (int)(m - (long)img * p.g.R);
                int y, x;
                valid[c] = __ballot_sync(0xffffffffu, m < rows_total && p.g.interior(r_img, y, x));
            }
            mbar_wait(&acc_full[buf], (li >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t trow = tmem_base + buf * 256 + ph * 128 + ((uint32_t)(q * 32) << 16);
            if (DBG & 32) {
                // TMA-store epilogue: the four warps of a 128-pixel half assemble [32 px][128 ch] chunks in shared memory
                // (fragment-layout TMEM reads, stmatrix.trans, SWIZZLE_128B) and one thread stores each chunk as two
                // [32 x 64] boxes: 256 contiguous bytes per pixel, no per-thread global stores.  Two 8 KB buffers per half.
                uint32_t f[4][2][16];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) tmem_ld_16x256b_x4(trow + c * 32 + ((uint32_t)(hf * 16) << 16), f[c][hf]);
                tmem_ld_wait();
                tc_fence_before_sync();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
                const int tr = lane >> 2, tc2 = (lane & 3) * 2;
                float bia[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int chn = q * 32 + g * 8 + tr;
                    bia[g] = (job.bias ? job.bias[chn] : 0.f) + (job.bias_img ? job.bias_img[img_h * kN + chn] : 0.f);
                }
                const bool store_half = px_base < rows_total;               // rows_total is a multiple of 128: a half is in or out
                const bool leader = q == 0 && lane == 0;
                const CUtensorMap* omap = &p.maps32[job.out_map32];
                const int orow = job.out_map_row + (int)(job.out_row_base + px_base);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint8_t* sbuf = smem_stage + (ph * 2 + (c & 1)) * 8192;  // two boxes: channels 0-63, 64-127
                    uint8_t* box = sbuf + (q >> 1) * 4096;
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t plo[4], phi[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int col = 8 * k + tc2;
                            const bool v0 = (valid[c] >> col) & 1u, v1 = (valid[c] >> (col + 1)) & 1u;
                            float a0 = __uint_as_float(f[c][hf][4 * k]) + bia[2 * hf], a1 = __uint_as_float(f[c][hf][4 * k + 1]) + bia[2 * hf];
                            float b0 = __uint_as_float(f[c][hf][4 * k + 2]) + bia[2 * hf + 1], b1 = __uint_as_float(f[c][hf][4 * k + 3]) + bia[2 * hf + 1];
                            if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); b0 = fmaxf(b0, 0.f); b1 = fmaxf(b1, 0.f); }
                            plo[k] = pack_act2(v0 ? a0 : 0.f, v1 ? a1 : 0.f);
                            phi[k] = pack_act2(v0 ? b0 : 0.f, v1 ? b1 : 0.f);
                        }
                        // lane l addresses pixel l (a 128-byte row of the box); 16-byte chunk = (q & 1) * 4 + channel octet, XOR row & 7
                        const int c16 = (q & 1) * 4 + 2 * hf;
                        stmatrix_x4_trans(box + lane * 128 + (((c16) ^ (lane & 7)) << 4), plo[0], plo[1], plo[2], plo[3]);
                        stmatrix_x4_trans(box + lane * 128 + (((c16 + 1) ^ (lane & 7)) << 4), phi[0], phi[1], phi[2], phi[3]);
                    }
                    fence_proxy_async_smem();
                    asm volatile("bar.sync %0, 128;" ::"r"(3 + ph) : "memory");
                    if (leader) {
                        if (store_half) {
                            if (job.out_accumulate) {        // out += tile: the ResidualBlock identity, added in L2 (gemm.cuh)
                                tma_reduce_add_2d(omap, sbuf, 0, orow + c * 32);
                                tma_reduce_add_2d(omap, sbuf + 4096, 64, orow + c * 32);
                            } else {
                                tma_store_2d(omap, sbuf, 0, orow + c * 32);
                                tma_store_2d(omap, sbuf + 4096, 64, orow + c * 32);
                            }
                        }
                        tma_store_commit();
                        tma_store_wait_read<1>();                             // the other buffer (chunk c - 1) has been read
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(3 + ph) : "memory");
                }
                if (li + 1 == n_mine && leader) tma_store_wait_all();
                continue;
            }
            if (DBG & 8) {
                // stmatrix epilogue: the accumulator is read in the mma-fragment layout (16 lanes x 8 columns per
                // repetition), packed to 16 bits and written TRANSPOSED into a [32 px][32 ch] staging tile with four
                // stmatrix per 32-pixel chunk, then leaves as 16-byte stores (64 contiguous bytes per pixel): 14
                // memory-queue instructions per thread and chunk instead of 16 shuffles + 16 4-byte stores.
                uint32_t f[4][2][16];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) tmem_ld_16x256b_x4(trow + c * 32 + ((uint32_t)(hf * 16) << 16), f[c][hf]);
                tmem_ld_wait();
                tc_fence_before_sync();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
                uint8_t* stage = smem_stage + warp * 2048;
                const int tr = lane >> 2, tc2 = (lane & 3) * 2;            // fragment row (channel) and first column (pixel)
                float bia[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int chn = q * 32 + g * 8 + tr;
                    bia[g] = (job.bias ? job.bias[chn] : 0.f) + (job.bias_img ? job.bias_img[img_h * kN + chn] : 0.f);
                }
                const int crow = lane >> 2, cchunk = lane & 3;              // store mapping: pixel crow + 8 i, 16-byte piece cchunk
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t plo[4], phi[4];                             // channel groups g = 2 hf (lanes 0-7 of the half) and 2 hf + 1
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int col = 8 * k + tc2;
                            const bool v0 = (valid[c] >> col) & 1u, v1 = (valid[c] >> (col + 1)) & 1u;
                            float a0 = __uint_as_float(f[c][hf][4 * k]) + bia[2 * hf], a1 = __uint_as_float(f[c][hf][4 * k + 1]) + bia[2 * hf];
                            float b0 = __uint_as_float(f[c][hf][4 * k + 2]) + bia[2 * hf + 1], b1 = __uint_as_float(f[c][hf][4 * k + 3]) + bia[2 * hf + 1];
                            if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); b0 = fmaxf(b0, 0.f); b1 = fmaxf(b1, 0.f); }
                            plo[k] = pack_act2(v0 ? a0 : 0.f, v1 ? a1 : 0.f);
                            phi[k] = pack_act2(v0 ? b0 : 0.f, v1 ? b1 : 0.f);
                        }
                        // lane l addresses pixel l of the chunk (matrix l/8 = pixel group, row l%8); chunk index swizzled
                        // with the pixel so that the 8 rows of a matrix and the 16-byte reads below are conflict-free
                        const int sw = (lane >> 1) & 3;
                        stmatrix_x4_trans(stage + lane * 64 + (((2 * hf) ^ sw) << 4), plo[0], plo[1], plo[2], plo[3]);
                        stmatrix_x4_trans(stage + lane * 64 + (((2 * hf + 1) ^ sw) << 4), phi[0], phi[1], phi[2], phi[3]);
                    }
                    __syncwarp();
                    act_t* o16 = job.out + (job.out_row_base + px_base + c * 32) * kN + q * 32 + cchunk * 8;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = crow + 8 * i;
                        const uint4 ov = *reinterpret_cast<const uint4*>(stage + row * 64 + ((cchunk ^ ((row >> 1) & 3)) << 4));
                        if (px_base + c * 32 + row < rows_total) *reinterpret_cast<uint4*>(o16 + (long)row * kN) = ov;
                    }
                    __syncwarp();
                }
                continue;
            }
            uint32_t v[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_32x32(trow + c * 32, v[c]);
            tmem_ld_wait();
            // all TMEM reads of this accumulator are done: hand it back to its MMA thread at once
            tc_fence_before_sync();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            // pack pixel pairs (2 i, 2 i + 1) of this lane's channel, swap halves with the neighbour lane:
            // even lanes end up with channels (ch, ch+1) of pixel 2 i, odd lanes with (ch-1, ch) of pixel 2 i + 1
            act_t* out = job.out + (job.out_row_base + px_base + (lane & 1)) * kN + (ch & ~1);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float f0 = __uint_as_float(v[c][j]) + bias_c, f1 = __uint_as_float(v[c][j + 1]) + bias_c;
                    if (relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
                    if (!((valid[c] >> j) & 1u)) f0 = 0.f;
                    if (!((valid[c] >> (j + 1)) & 1u)) f1 = 0.f;
                    const uint32_t mine = pack_act2(f0, f1);                       // lo = pixel j, hi = pixel j+1
                    const uint32_t theirs = __shfl_xor_sync(0xffffffffu, mine, 1);
                    // even lane: (mine.lo, theirs.lo)   odd lane: (theirs.hi, mine.hi)
                    const uint32_t o = (lane & 1) ? __byte_perm(theirs, mine, 0x7632) : __byte_perm(mine, theirs, 0x5410);
                    const long px = px_base + c * 32 + j + (lane & 1);
                    if (DBG & 4) {          // measurement: same bytes in a quarter of the store instructions (16-byte stores of garbage)
                        if ((j & 6) == 0 && px + 6 < rows_total) {
                            uint4* a16 = reinterpret_cast<uint4*>(reinterpret_cast<uintptr_t>(out + (long)(c * 32 + j) * kN) & ~(uintptr_t)15);
                            *a16 = make_uint4(o, o, o, o);
                        }
                    } else
                    if (px < rows_total && !(DBG & 2)) *reinterpret_cast<uint32_t*>(out + (long)(c * 32 + j) * kN) = o;
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

int slab2_rows(const Geom& g, int n_taps) {
    const int lead = n_taps == 9 ? g.Wp + 1 : 0;
    return kBM + 2 * lead;
}

}  // namespace

// Activation slab of a tile: 256 rows + both leads, in as few equal TMA boxes (<= 256 rows, a multiple
// of 8 rows so that every box starts on a swizzle period) as possible.
int slab2_box_rows(const Geom& g, int n_taps) {
    const int rows = slab2_rows(g, n_taps);
    const int n_ops = (rows + 255) / 256;
    return ((rows + n_ops - 1) / n_ops + 7) / 8 * 8;
}

static int slab2_boxes(const Geom& g, int n_taps) {
    const int rows = slab2_rows(g, n_taps);
    return (rows + slab2_box_rows(g, n_taps) - 1) / slab2_box_rows(g, n_taps);
}

static int slab2_smem(const Geom& g, int n_taps) {
    return kSlabStages * slab2_boxes(g, n_taps) * slab2_box_rows(g, n_taps) * kRowBytes + kWSlots * kWSlot + 32 * 1024 + 1024;
}

// geometry-only part of the test below: can a 3x3 128->128 launch at this image size run on this kernel at all?
bool slab2_geom_supported(const Geom& g) {
    if (!measure_env("BMC_CONV_SLAB2", 1)) return false;
    return slab2_boxes(g, 9) <= 32 && slab2_smem(g, 9) + 2048 <= 227 * 1024;
}

bool slab2_supported(const GemmParams& p) {
    static int enabled = -1;
    if (enabled < 0) enabled = measure_env("BMC_CONV_SLAB2", 1);
    if (!enabled || !p.has32 || p.n != kN || (p.n_taps != 9 && p.n_taps != 1) || (p.tap1_mask & 1)) return false;
    int steps = 0;
    for (int s = 0; s < p.n_seg; ++s) steps += 2 * p.chunks[s];
    if (steps > kMaxSteps) return false;
    for (int j = 0; j < p.n_jobs; ++j) {
        const GemmJobDev& d = p.jobs[j];
        if (d.w_img_stride != 0 || d.residual || d.out_f32 || d.ln_gamma || !d.out || d.w_map32 < 0) return false;
        if (d.out_accumulate && d.out_map32 < 0) return false;      // the in-place add exists only in the TMA-store epilogue
        for (int s = 0; s < p.n_seg; ++s) {
            if (d.a_map32[s] < 0) return false;
            if (((p.tap1_mask >> s) & 1) && d.t1_map32[s] < 0) return false;
        }
    }
    if (p.abox32_rows != slab2_box_rows(p.g, p.n_taps)) return false;
    return slab2_boxes(p.g, p.n_taps) <= 32 && slab2_smem(p.g, p.n_taps) + 2048 <= 227 * 1024;
}

int launch_conv_slab2(GemmParams p, cudaStream_t st) {
    p.slab_lead = p.n_taps == 9 ? p.g.Wp + 1 : 0;
    p.slab_boxes = slab2_boxes(p.g, p.n_taps);
    p.slab2_stages = kSlabStages;
    p.per_image = 0;
    p.pimg_mask = 0;
    for (int sg = 0; sg < p.n_seg; ++sg)
        if (((p.tap1_mask >> sg) & 1) && p.jobs[0].t1_img_stride[sg] != 0) p.pimg_mask |= 1 << sg;
    p.n_full = p.n_jobs * (int)((p.g.rows() + kBM - 1) / kBM);
    p.n_half = 0;
    {   // K steps: every 64-channel chunk of every segment as two 32-channel steps
        auto& sp = p.s2;
        sp.n = 0; sp.t1 = 0; sp.pm = 0;
        int seg_chunk0 = 0;
        for (int sg = 0; sg < p.n_seg; ++sg) {
            const bool t1 = (p.tap1_mask >> sg) & 1;
            for (int c = 0; c < p.chunks[sg]; ++c)
                for (int h = 0; h < 2; ++h, ++sp.n) {
                    sp.seg[sp.n] = (unsigned char)sg; sp.c64[sp.n] = (unsigned char)c; sp.cs[sp.n] = (unsigned char)p.chunks[sg];
                    sp.col[sp.n] = (short)(c * 64 + h * kSubK); sp.kchunk0[sp.n] = (short)(seg_chunk0 + c);
                    if (t1 && p.n_taps != 1) sp.t1 |= 1u << sp.n;
                    if ((p.pimg_mask >> sg) & 1) sp.pm |= 1u << sp.n;
                }
            if (!t1) seg_chunk0 += p.n_taps * p.chunks[sg];
        }
    }
    const int smem = slab2_smem(p.g, p.n_taps);
    static PerDevice configured_dev;
    int& configured = configured_dev.cur();
    if (configured < smem) {
        BMC_CUDA(cudaFuncSetAttribute(conv_slab2_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        BMC_CUDA(cudaFuncSetAttribute(conv_slab2_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
#ifdef BMC_MEASURE
        BMC_CUDA(cudaFuncSetAttribute(conv_slab2_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        BMC_CUDA(cudaFuncSetAttribute(conv_slab2_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        BMC_CUDA(cudaFuncSetAttribute(conv_slab2_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        BMC_CUDA(cudaFuncSetAttribute(conv_slab2_tc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        BMC_CUDA(cudaFuncSetAttribute(conv_slab2_tc<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
#endif
        configured = smem;
    }
    int grid = p.n_full < sm_count() ? p.n_full : sm_count();
    {   // measurement switch: fewer CTAs -> is a tile's time set by the SM or by the shared L2 fabric?
        static int cap = -1;
        if (cap < 0) cap = measure_env("BMC_SLABT_GRID", 0);
        if (cap > 0 && grid > cap) grid = cap;
    }
    bool tma_out = true;
    for (int j = 0; j < p.n_jobs; ++j) tma_out = tma_out && p.jobs[j].out_map32 >= 0;
#ifdef BMC_MEASURE
    // measurement builds of the kernel (stale-operand MMAs, dropped stores, ...): they do NOT compute the
    // convolution and exist only in `build.py --measure` libraries
    static int dbg = -1;
    if (dbg < 0) dbg = measure_env("BMC_SLAB2_DBG", 0);
    if (dbg == 1) { BMC_CUDA(launch_pdl(conv_slab2_tc<1>, dim3(grid), dim3(kThreads2), (size_t)smem, st, p)); return BMC_OK; }
    if (dbg == 2) { BMC_CUDA(launch_pdl(conv_slab2_tc<2>, dim3(grid), dim3(kThreads2), (size_t)smem, st, p)); return BMC_OK; }
    if (dbg == 3) { BMC_CUDA(launch_pdl(conv_slab2_tc<3>, dim3(grid), dim3(kThreads2), (size_t)smem, st, p)); return BMC_OK; }
    if (dbg == 4) { BMC_CUDA(launch_pdl(conv_slab2_tc<4>, dim3(grid), dim3(kThreads2), (size_t)smem, st, p)); return BMC_OK; }
    if (dbg == 8) { BMC_CUDA(launch_pdl(conv_slab2_tc<8>, dim3(grid), dim3(kThreads2), (size_t)smem, st, p)); return BMC_OK; }
    if (dbg == 64) tma_out = false;
#endif
    if (tma_out) BMC_CUDA(launch_pdl(conv_slab2_tc<32>, dim3(grid), dim3(kThreads2), (size_t)smem, st, p));     // product path: TMA-store epilogue
    else BMC_CUDA(launch_pdl(conv_slab2_tc<0>, dim3(grid), dim3(kThreads2), (size_t)smem, st, p));              // outputs without a tensor map
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

}  // namespace bmc

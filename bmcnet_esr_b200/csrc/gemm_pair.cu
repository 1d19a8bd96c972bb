// conv_pair_tc: the slab convolution (gemm_slab.cu) on CTA PAIRS with tcgen05 cta_group::2.
//
// The single-CTA slab kernel is bound by what each SM can pull in from L2: per 64-channel K chunk a
// 256-row tile needs a 56 KB activation slab plus 144 KB of weights (9 taps x 16 KB), ~25 B/clk per SM,
// and the MMA issuers spend 30 % of their time waiting for weight stages.  A CTA pair (two SMs of one
// TPC, cluster of 2) runs every MMA as ONE M=256 instruction over both CTAs' rows, each CTA supplying
// its own 128 rows of A and only HALF of the B tile (64 of the 128 output channels): the weight bytes
// per SM halve, and so does the shared memory the weight ring needs.
//
//   pair tile = 512 rows: CTA r of the pair owns rows [M0 + 256 r, M0 + 256 r + 256) -- its slab, its
//   two accumulator halves (TMEM of each CTA holds its own rows) and its epilogue are exactly those of
//   the slab kernel.
//   leader (rank 0): the two MMA-issuing threads (one per 128-row half of both CTAs' tiles).  Operand
//   barriers (a_full, w_full) live in the leader and collect arrive.expect_tx + TMA complete_tx from
//   BOTH CTAs' producers; tcgen05.commit multicasts to the barriers of both CTAs (a_empty, w_empty for
//   the producers, acc_full for the epilogues); acc_empty lives in the leader and collects the 16
//   epilogue warps of both CTAs.
// Restrictions (the launcher falls back to conv_slab_tc otherwise): N = 128, no per-image weights.
#include "gemm_epi.cuh"

namespace bmc {
namespace {

constexpr int kThreadsPair = 384;
constexpr int kBM = 256;
constexpr int kBoxRows = 64;
constexpr int kBoxBytes = kBoxRows * kChunkK * 2;     // 8192
constexpr int kSlabStages = 2;
constexpr int kWGroup = 2;          // K steps per weight stage
constexpr int kWStages = 6;
constexpr int kMaxASteps = 8;
constexpr int kN = 128;
constexpr int kWBytes = (kN / 2) * kChunkK * 2;       // this CTA's half of one [128 x 64] weight tile: 8 KB
constexpr int kWStageBytes = kWGroup * kWBytes;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsPair, 1)
conv_pair_tc(const __grid_constant__ GemmParams p) {
    constexpr int N = kN;
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int slab_bytes = p.slab_boxes * kBoxBytes;
    uint8_t* smem_w = smem + kSlabStages * slab_bytes;
    uint8_t* smem_stage = smem_w + kWStages * kWStageBytes;      // 8 epilogue warps x 2 KB

    __shared__ uint64_t a_full[kSlabStages], a_empty[kSlabStages], w_full[kWStages], w_empty[kWStages];
    __shared__ uint64_t acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[N];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const long rows_total = p.g.rows();
    const int n_taps = p.n_taps;
    const int tap1_mask = p.tap1_mask;
    // (segment, chunk) steps of one tile and where their weights start
    int a_steps = 0;
    int as_seg[kMaxASteps], as_chunk[kMaxASteps], as_k0[kMaxASteps], as_cs[kMaxASteps];
    uint32_t t1_as = 0;
    {
        int seg_chunk0 = 0;
        for (int s = 0; s < p.n_seg; ++s) {
            const bool t1 = (tap1_mask >> s) & 1;
            for (int c = 0; c < p.chunks[s]; ++c, ++a_steps) {
                as_seg[a_steps] = s; as_chunk[a_steps] = c; as_k0[a_steps] = seg_chunk0 + c; as_cs[a_steps] = p.chunks[s];
                if (t1 && n_taps != 1) t1_as |= 1u << a_steps;
            }
            if (!t1) seg_chunk0 += n_taps * p.chunks[s];
        }
    }
    // pair tiles: `pairs_per_job` per job, dealt round-robin to the clusters; both CTAs of a pair walk the same list
    const int n_clusters = gridDim.x >> 1, cid = blockIdx.x >> 1;
    const int pairs_total = p.pairs_per_job * p.n_jobs;
    const int n_mine = pairs_total > cid ? (pairs_total - cid + n_clusters - 1) / n_clusters : 0;
    auto tile_at = [&](int li, int& job, long& m0) {
        const int pt = cid + li * n_clusters;
        job = pt / p.pairs_per_job;
        m0 = (long)(pt - job * p.pairs_per_job) * (2 * kBM) + (long)rank * kBM;
    };

    if (threadIdx.x == 0) {
        // leader: full barriers collect one arrive.expect_tx from each CTA's producer; acc_empty 8 warps of each CTA.
        // both: empty / acc_full barriers collect one (multicast) tcgen05.commit from each of the two MMA issuers.
        for (int s = 0; s < kSlabStages; ++s) { mbar_init(&a_full[s], 2); mbar_init(&a_empty[s], 2); }
        for (int s = 0; s < kWStages; ++s) { mbar_init(&w_full[s], 2); mbar_init(&w_empty[s], 2); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 2); mbar_init(&acc_empty[s], 16); }
        mbar_fence_init();
    }
    if (warp == 10) tmem_alloc_pair(&tmem_base_s, 4 * N);
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                                // both CTAs' barriers are initialised before anyone signals them
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

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
                const int abox = p.abox_rows;                                       // rows per TMA op
                const int boxes = ((t1 ? kBM : p.slab_boxes * kBoxRows) + abox - 1) / abox;
                const uint32_t bar = mapa_rank(smem_u32(&a_full[st]), 0);
                if (lane == 0) mbar_expect_tx_cluster(bar, boxes * abox * 128);
                __syncwarp();
                if (lane < boxes)
                    tma_load_2d_pair(dst + lane * abox * 128, map, bar, job.a_col_base[s] + as_chunk[as] * kChunkK, row0 + lane * abox);
            }
        }
    } else if (warp == 9) {
        // ------------------------------------------------------------ weight tiles: this CTA's 64 output channels of each
        int gi = 0;
        for (int li = 0; li < n_mine; ++li) {
            long m0l; int ji;
            tile_at(li, ji, m0l);
            const GemmJobDev& job = p.jobs[ji];
            const CUtensorMap* map = &p.maps[job.w_map64];
            int as = 0, tap = 0;
            while (as < a_steps) {
                const int st = gi % kWStages;
                if (gi >= kWStages) mbar_wait(&w_empty[st], ((gi / kWStages) - 1) & 1);
                uint8_t* dst = smem_w + st * kWStageBytes + (lane & 1) * kWBytes;
                int as2 = as, tap2 = tap + 1;
                if (tap2 == (((t1_as >> as) & 1u) ? 1 : n_taps)) { tap2 = 0; ++as2; }
                const bool two = as2 < a_steps;
                const uint32_t bar = mapa_rank(smem_u32(&w_full[st]), 0);
                if (lane == 0) mbar_expect_tx_cluster(bar, (two ? 2 : 1) * kWBytes);
                __syncwarp();
                if (lane < (two ? 2 : 1)) {
                    const int asl = lane ? as2 : as, tapl = lane ? tap2 : tap;     // this lane's step
                    int row;
                    if ((t1_as >> asl) & 1u) row = job.t1_row[as_seg[asl]] + as_chunk[asl] * 128;      // static 1-tap weights (identity)
                    else row = (as_k0[asl] + tapl * as_cs[asl]) * job.w_rows + job.w_row_base;        // K order (seg, tap, chunk)
                    tma_load_2d_pair(dst, map, bar, 0, row + 64 * (int)rank);
                }
                as = as2; tap = tap2;
                if (two) { if (++tap == (((t1_as >> as) & 1u) ? 1 : n_taps)) { tap = 0; ++as; } }
                ++gi;
            }
        }
    } else if (warp >= 10) {
        // ------------------------------------------------------------ MMA issuers (leader CTA only)
        if (lane == 0 && leader) {
            const int hf = warp - 10;
            constexpr uint32_t idesc = umma_idesc_f16(256, N, false, false);
            constexpr uint32_t hi = umma_desc_hi_sw128(1024);
            const uint32_t tap0_lo = (uint32_t)(p.slab_lead + p.tap_off[0] + hf * 128) * 8u;
            const uint32_t row_wrap = (uint32_t)(p.g.Wp - 2) * 8u;
            const uint32_t slab_lo0 = umma_desc_lo(smem_u32(smem), 16);
            const uint32_t w_lo0 = umma_desc_lo(smem_u32(smem_w), 16);
            const uint32_t slab_step = (uint32_t)slab_bytes >> 4;
            int ia = 0, gi = 0, lt = 0;
            const bool prof_on = p.prof != nullptr;
            const long long t_begin = prof_on ? clock64() : 0;
            long long pw_a = 0, pw_w = 0, pw_acc = 0;
#define PPROF(var, stmt) do { const long long _t = prof_on ? clock64() : 0; stmt; if (prof_on) var += clock64() - _t; } while (0)
            for (int li = 0; li < n_mine; ++li, ++lt) {
                const int buf = lt & 1;
                if (lt >= 2) PPROF(pw_acc, mbar_wait(&acc_empty[buf], ((lt >> 1) - 1) & 1));
                tc_fence_after_sync();
                const uint32_t acc = tmem_base + buf * 2 * N + hf * N;
                uint32_t accumulate = 0;
                int tap = 0, dx = 0, sa = ia % kSlabStages, as = 0, cur_taps = (t1_as & 1u) ? 1 : n_taps;
                uint32_t slab_lo = slab_lo0 + sa * slab_step, tap_lo = tap0_lo;
                while (as < a_steps) {
                    const int sw = gi % kWStages;
                    PPROF(pw_w, mbar_wait(&w_full[sw], (gi / kWStages) & 1));
                    const uint32_t b_base = w_lo0 + sw * (kWStageBytes >> 4);
                    for (int e = 0; e < 2; ++e) {
                        const uint32_t b_lo = b_base + e * (kWBytes >> 4);
                        if (tap == 0) PPROF(pw_a, mbar_wait(&a_full[sa], (ia / kSlabStages) & 1));
                        tc_fence_after_sync();
                        const uint32_t a_lo = slab_lo + tap_lo;
                        umma_f16_pair(acc, umma_desc(a_lo, hi), umma_desc(b_lo, hi), idesc, accumulate);
                        umma_f16_pair(acc, umma_desc(a_lo + 2, hi), umma_desc(b_lo + 2, hi), idesc, 1u);
                        umma_f16_pair(acc, umma_desc(a_lo + 4, hi), umma_desc(b_lo + 4, hi), idesc, 1u);
                        umma_f16_pair(acc, umma_desc(a_lo + 6, hi), umma_desc(b_lo + 6, hi), idesc, 1u);
                        accumulate = 1u;
                        if (++tap == cur_taps) {           // slab fully consumed (in both CTAs)
                            umma_commit_pair(&a_empty[sa]);
                            tap = 0; dx = 0; tap_lo = tap0_lo; ++ia; ++as;
                            cur_taps = ((t1_as >> as) & 1u) ? 1 : n_taps;
                            sa = ia % kSlabStages;
                            slab_lo = slab_lo0 + sa * slab_step;
                        } else if (++dx == 3) {
                            dx = 0; tap_lo += row_wrap;
                        } else {
                            tap_lo += 8u;
                        }
                        if (as >= a_steps) break;
                    }
                    umma_commit_pair(&w_empty[sw]);
                    ++gi;
                }
                umma_commit_pair(&acc_full[buf]);
            }
            if (prof_on && hf == 0) {
                long long* o = p.prof + (blockIdx.x >> 1) * 16;
                o[0] = pw_a; o[1] = pw_w; o[2] = pw_acc; o[3] = clock64() - t_begin; o[4] = lt;
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 0..7), own rows of the pair tile
        const int hf = warp >> 2, q = warp & 3;
        int lt = 0;
        for (int li = 0; li < n_mine; ++li, ++lt) {
            const int buf = lt & 1;
            long m0l; int ji;
            tile_at(li, ji, m0l);
            const GemmJobDev& job = p.jobs[ji];
            const long m_end = rows_total;
            const long m = m0l + hf * 128 + q * 32 + lane;
            {
                const int t = threadIdx.x;                // 0..255
                asm volatile("bar.sync 1, 256;" ::: "memory");      // every warp is done with the previous tile's vector
                if (t < N) bias_s[t] = job.bias ? job.bias[t] : 0.f;
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const int img = (int)(m / p.g.R);
            const int r_img = (int)(m - (long)img * p.g.R);
            int y, x;
            const bool valid = m < m_end && p.g.interior(r_img, y, x);
            mbar_wait(&acc_full[buf], (lt >> 1) & 1);
            tc_fence_after_sync();
            const long m_warp = m - lane;                  // first row of this warp's 32-row block
            EpiTile r;
            r.res = job.residual ? job.residual + (job.res_row_base + m_warp) * N : nullptr;
            r.out = job.out ? job.out + (job.out_row_base + m_warp) * N : nullptr;
            r.outf = job.out_f32 ? job.out_f32 + (job.out_row_base + m) * N : nullptr;
            r.rows_left = m_end - m_warp;
            r.valid = valid; r.relu = job.relu != 0; r.n = N;
            uint8_t* stage = smem_stage + warp * 2048;
            const uint32_t trow = tmem_base + buf * 2 * N + hf * N + ((uint32_t)(q * 32) << 16);
            const bool ln = job.ln_gamma != nullptr;
            float mu = 0.f, rstd = 1.f;
            if (ln) epi_ln_stats<N>(trow, bias_s, job.ln_eps, mu, rstd);
            const uint32_t acc_empty_leader = mapa_rank(smem_u32(&acc_empty[buf]), 0);
#pragma unroll 1
            for (int c = 0; c < N / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(trow + c * 32, v);
                tmem_ld_wait();
                if (c == N / 32 - 1) {
                    // all TMEM reads of this accumulator are done: hand it back to the leader's MMA threads
                    tc_fence_before_sync();
                    if (lane == 0) mbar_arrive_cluster(acc_empty_leader);
                }
                epi_chunk_staged(v, c, r, bias_s, ln, mu, rstd, job.ln_gamma, job.ln_beta, stage, lane);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                                // the peer may still signal our barriers / read our operands
    if (warp == 10) {
        tc_fence_after_sync();
        tmem_dealloc_pair(tmem_base, 4 * N);
    }
}

int pair_smem(int slab_boxes) { return kSlabStages * slab_boxes * kBoxBytes + kWStages * kWStageBytes + 8 * 2048 + 1024; }

}  // namespace

// The pair kernel applies to N = 128 launches without per-image weights whose operands have the 64-row
// TMA maps (a_map64 / w_map64), on an even grid.
bool pair_supported(const GemmParams& p) {
    static int enabled = -1;
    if (enabled < 0) enabled = measure_env("BMC_CONV_PAIR", 0);   // off by default: measured slower than the single-CTA slab kernel
    if (!enabled || p.n != kN) return false;
    for (int j = 0; j < p.n_jobs; ++j) {
        if (p.jobs[j].w_img_stride != 0 || p.jobs[j].w_map64 <= 0) return false;
        for (int s = 0; s < p.n_seg; ++s) {
            if (p.jobs[j].a_map64[s] < 0) return false;
            if (((p.tap1_mask >> s) & 1) && p.jobs[j].t1_img_stride[s] != 0) return false;
        }
    }
    if (p.tap1_mask & 1) return false;
    const int lead = p.n_taps == 9 ? p.g.Wp + 1 : 0;
    const int boxes = (kBM + 2 * lead + kBoxRows - 1) / kBoxRows;
    return pair_smem(boxes) + 2048 <= 227 * 1024 && sm_count() >= 2;
}

int launch_conv_pair(GemmParams p, cudaStream_t st) {
    p.slab_lead = p.n_taps == 9 ? p.g.Wp + 1 : 0;
    p.slab_boxes = (kBM + 2 * p.slab_lead + kBoxRows - 1) / kBoxRows;
    p.pairs_per_job = (int)((p.g.rows() + 2 * kBM - 1) / (2 * kBM));
    if (p.abox_rows <= 0) p.abox_rows = kBoxRows;
    const int smem = pair_smem(p.slab_boxes);
    static PerDevice configured_dev;
    int& configured = configured_dev.cur();
    if (configured < smem) {
        BMC_CUDA(cudaFuncSetAttribute(conv_pair_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    const int pairs_total = p.pairs_per_job * p.n_jobs;
    const int max_clusters = sm_count() / 2;
    const int clusters = pairs_total < max_clusters ? pairs_total : max_clusters;
    static long long* prof = nullptr;
    static int prof_init = 0, dumped = 0;
    if (!prof_init) {
        prof_init = 1;
        if (measure_env("BMC_PAIR_PROF", 0)) { cudaMalloc(&prof, 74 * 16 * sizeof(long long)); cudaMemset(prof, 0, 74 * 16 * sizeof(long long)); }
    }
    p.prof = prof;
    conv_pair_tc<<<2 * clusters, kThreadsPair, smem, st>>>(p);
    BMC_CUDA(cudaGetLastError());
    if (prof && dumped++ == 3) {
        cudaStreamSynchronize(st);
        long long h[74 * 16];
        cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
        for (int c : {0, 1, 36, 73}) {
            const long long* o = h + c * 16;
            printf("pairprof cluster %2d: tiles %lld | MMA wait_a %lld wait_w %lld wait_acc %lld of %lld\n", c, o[4], o[0], o[1], o[2], o[3]);
        }
    }
    return BMC_OK;
}

}  // namespace bmc

// Launch descriptors shared by the tensor-core kernels, the SIMT cross-check kernels and the
// model plan (model.cu).
//
// One "conv-gemm" launch computes, for every job j and every row m of B padded images,
//     out_j[m, :] = act( sum_seg sum_tap A_{j,seg}[m + off_tap, :] . W_j[:, (seg,tap,:)]^T + bias_j )
//                   (+ residual_j[m, :]),   halo / tail rows forced to zero,
// which covers the 3x3 and 1x1 convolutions of the reference (models/BMCNet.py:40-53,
// models/submodules.py:25-26,44-53), their channel concatenations (torch.cat -> K segments),
// and the `softmax(att) @ v` product of BIE (submodules.py:72-73; per-image dynamic weights).
//
// Weights are bf16 in "chunk-major" form: [K/64][w_rows][64], so the B tile of a K chunk is one
// contiguous [N x 64] box and every weight of the model lives in one tall [rows][64] matrix
// behind a single TMA descriptor.
#pragma once
#include "common.cuh"

namespace bmc {

constexpr int kMaxSeg = 4;
constexpr int kMaxJobs = 8;
constexpr int kMaxMaps = 8;
constexpr int kMaxMaps32 = 6;
constexpr int kTileM = 128;
constexpr int kChunkK = 64;

struct GemmJobDev {
    int a_map[kMaxSeg];        // index into GemmParams::maps: box [128 rows x 64 ch]
    int a_map64[kMaxSeg];      // same tensor, box [64 rows x 64 ch] (slab kernel), or -1
    int a_row_base[kMaxSeg];   // row of the source paired with output row 0
    int a_col_base[kMaxSeg];   // first channel used
    const act_t* a_ptr[kMaxSeg];   // same sources as raw pointers (SIMT kernel)
    int a_ld[kMaxSeg];
    long a_rows[kMaxSeg];
    int w_map;
    int w_map64;               // same weights behind a [64 rows x 64] box (pair kernel: each CTA loads half a tile), or 0
    const act_t* w_ptr;
    int w_rows;                // rows per K chunk of the weight matrix behind w_map / w_ptr
    int w_row_base;
    int w_img_stride;
    const float* bias;
    const act_t* residual;
    long res_row_base;
    act_t* out;
    long out_row_base;
    float* out_f32;
    int relu;
    // fused channel LayerNorm on (acc + bias) before the store (submodules.py:127-139,
    // `norm_s(convf(...))` at :63-64); N = 128 only
    const float* ln_gamma;
    const float* ln_beta;
    float ln_eps;
    // centre-tap-only ("1x1") segments of a 3x3 launch (GemmParams::tap1_mask), each with its own weights:
    //   * the per-image mix matrix M_k[b] of the fused BIE (out_k = softmax(att_k) @ v_k = M_k[b] . x_k,
    //     bie_fused.cu, submodules.py:72-77): t1_img_stride != 0;
    //   * the identity matrix: the residual add `identity + out` (submodules.py:35) done by the tensor core
    //     (1.0 * x accumulates exactly in fp32), which keeps the epilogue free of a second global read.
    int t1_map[kMaxSeg];       // TMA map of the segment's weights [.][64], chunk-major [chunks][128][64]
    int t1_row[kMaxSeg], t1_img_stride[kMaxSeg];
    const float* bias_img;     // per-image extra bias [B][128] added to `bias`, or NULL
    // conv_slab2_tc (gemm_slab2.cu): the same operands behind 32-channel / 64-byte-swizzle boxes,
    // indices into GemmParams::maps32 (valid when GemmParams::has32)
    signed char a_map32[kMaxSeg], t1_map32[kMaxSeg], w_map32;
    signed char out_map32;     // maps32 index of the OUTPUT tensor behind [32 rows x 64 ch] SWIZZLE_128B boxes (TMA-store epilogue), or -1
    int out_map_row;           // row of `out` inside that tensor
    // conv_slab2_tc only: ADD the result onto the 16-bit tensor already in `out` (TMA reduce-add stores) instead of
    // overwriting it: `identity + out` of ResidualBlock_noBN (submodules.py:35) when the block's output takes the slot of
    // its input -- out = act16(x + act16(conv + bias)), one more 16-bit rounding than a single fp32 sum, no identity
    // K segment, no residual read.  Only for launches whose OTHER operands do not include `out` with a spatial halo.
    int out_accumulate;
};

struct alignas(64) GemmParams {
    CUtensorMap maps[kMaxMaps];
    GemmJobDev jobs[kMaxJobs];
    int n_jobs;
    int n_seg;
    int chunks[kMaxSeg];       // 64-channel chunks per segment
    int n_taps;
    int tap_off[9];            // row shift of each tap: dy*Wp + dx
    int n;                     // output channels: 128 or 32
    int tap1_mask;             // bit s: segment s contributes its centre tap only, with weights from GemmJobDev::t1_*
    int pimg_mask;             // subset of tap1_mask whose weights are per image (filled in by the slab launcher)
    Geom g;
    int tiles_per_img;         // R / 128
    // slab kernel (gemm_slab.cu); filled in by its launcher
    int slab_lead, slab_boxes, bo_mode;
    int abox_rows;             // rows per TMA box of the a_map64 maps (slab_box_rows(): few large boxes, TMA cost is per op)
    int per_image;             // tiles never straddle images (per-image weights)
    // tile schedule: `n_full` 256-row tiles, then `n_half` 128-row tiles (the odd half at the end of every
    // image in per-image mode), the halves going to the CTAs that got one full tile less
    int n_full, n_half, full_per_img;
    int pairs_per_job;         // pair kernel (gemm_pair.cu): 512-row pair tiles per job
    long long* prof;           // optional per-CTA cycle counters (tools/gpu_diag.py slabprof), else NULL
    // conv_slab2_tc: 32-channel boxes (activations: abox32_rows x 32, weights: 128 x 32), 64-byte swizzle
    CUtensorMap maps32[kMaxMaps32];
    int has32, abox32_rows, slab2_stages;
    // its K steps (segment x 32 channels), built by the launcher
    struct {
        int n;                             // steps per tile (even)
        unsigned t1, pm;                   // per step: centre tap only / per-image weights
        unsigned char seg[16], c64[16], cs[16];     // segment, 64-channel chunk inside it, chunks of the segment
        short col[16], kchunk0[16];        // first channel inside the segment; first K chunk of the weight matrix
    } s2;
};

// > 4 KB of kernel parameters needs CUDA >= 12.1 on sm_70+ (limit 32764 bytes); sm_100a only here
static_assert(sizeof(GemmParams) <= 8192, "GemmParams grew unexpectedly");

// att[b] = centres[b]^T . v[b] (split over pixel ranges) -- submodules.py:69-70
constexpr int kMaxPairs = 4;
struct alignas(64) AttParams {
    CUtensorMap map_c, map_v;              // box [64 pixel rows][64 channels] over [rows][128]
    long c_row_base[kMaxPairs], v_row_base[kMaxPairs];
    const act_t* c_ptr; const act_t* v_ptr;   // raw bases (SIMT kernel)
    int n_pairs, n_split;
    int pix_per_split;                     // multiple of 64
    float scale;
    float* partial;                        // [pair][B][split][128][128]
    Geom g;
};

struct SoftmaxParams {
    const float* partial;                  // as above
    int n_pairs, n_split, B;
    act_t* w_base;                 // [rows][64] dynamic-weight arena; P of (pair, b) is the
    int w_row_base[kMaxPairs];             // chunk-major block [2][128][64] starting at row
    int w_img_stride;                      // w_row_base[pair] + b * w_img_stride
};

// Fused 1x1 / attention section of BIE.forward (bie_fused.cu; submodules.py:63-75).
constexpr int kMaxInst = 2;
struct BieInst {
    int x1_row, x2_row, xs_row, out_row;   // first arena row of x_1, x_2, x_s and of the x_s_ output (== xs_row: in place)
};
struct alignas(64) BieFrontParams {
    CUtensorMap map_act;                   // activation arena, box [128 rows][64 ch]
    CUtensorMap map_w;                     // static weights, box [128 rows][64]
    CUtensorMap map_out;                   // activation arena, box [32 rows][64 ch]: the in-place x_s' reduce-add stores
    BieInst inst[kMaxInst];
    int n_inst;
    int wf_row, wc_row, wu_row;            // first rows of convf (4 chunks), clustering WITH norm_s's gamma folded in (2), unclustering (4)
    const float* bf; const float* bc; const float* bu;     // bc: clustering bias + clustering weight . norm_s beta
    float ln_eps;
    const act_t* act_base;                 // arena base (residual x_s rows / output rows)
    act_t* out_base;
    float* g_partial;                      // [slot][2][128][128]: sum_px c_k[px,c] * x_k[px,i]
    float* s_partial;                      // [slot][2][128]:      sum_px c_k[px,c]
    Geom g;
    int tiles_per_img, total_tiles, tiles_per_cta, lcm;      // lcm(tiles_per_cta, tiles_per_img): slot numbering
    long long* prof;                       // optional per-CTA cycle counters (BMC_FRONT_PROF), else NULL
};
struct FoldParams {
    const float* g_partial; const float* s_partial;
    int n_inst, B, tiles_per_img, tiles_per_cta, total_tiles, lcm;
    const act_t* w_base;                   // static weight matrix [rows][64]
    int wv_row[2];                         // first row of v1 / v2 (2 chunks of 128 rows each)
    const float* bv[2];
    float scale;
    act_t* m_base;                         // per-image matrices: rows pair*B*256 + b*256 + chunk*128 + c, pair = inst*2 + k
    float* bias_img;                       // [pair][B][128]
    int dbg;                               // BMC_FOLD_PROF: print phase cycle counts
    int pre_reduced;                       // fold_prereduce has summed each image's partial slots into its first one
};
int launch_bie_front(const BieFrontParams& p, cudaStream_t st);
int launch_att_fold(const FoldParams& p, cudaStream_t st);                                   // SIMT version (BMC_FOLD_SIMT=1)
int launch_att_fold_tc(const FoldParams& p, const CUtensorMap& map_w, cudaStream_t st);     // tensor-core version
int bie_front_grid(int total_tiles);
int bie_front_slots(int total_tiles, int tiles_per_img);
int bie_front_lcm(int tiles_per_cta, int tiles_per_img);

// Launchers (host).  `impl`: 0 = tcgen05/TMA (slab kernel when it applies, else the per-tap
// kernel), 1 = SIMT cross-check, 2 = force the per-tap tcgen05 kernel.
bool slab_supported(const GemmParams& p);
int slab_box_rows(const Geom& g, int n_taps);   // TMA box height the slab kernels expect behind a_map64
bool pair_supported(const GemmParams& p);
bool slabt_supported(const GemmParams& p);
int launch_conv_slabt(GemmParams p, cudaStream_t st);
bool slab2_supported(const GemmParams& p);
bool slab2_geom_supported(const Geom& g);
int slab2_box_rows(const Geom& g, int n_taps);   // TMA box height conv_slab2_tc expects behind the activation maps32
int launch_conv_slab2(GemmParams p, cudaStream_t st);
int launch_conv_pair(GemmParams p, cudaStream_t st);
int launch_conv_slab(GemmParams p, cudaStream_t st);
int launch_conv_gemm(const GemmParams& p, int impl, cudaStream_t st);
int launch_att(const AttParams& p, int impl, cudaStream_t st);
int launch_att_softmax(const SoftmaxParams& p, cudaStream_t st);

// Pointwise kernels (pointwise.cu)
int launch_layernorm(const act_t* in, const float* gamma, const float* beta, float eps,
                     long rows, act_t* out, cudaStream_t st);
int launch_pack_nchw(const float* src, Geom g, int C, act_t* dst, int c_pad, int c_off,
                     cudaStream_t st);
int launch_unpack_nchw(const act_t* src, Geom g, int C, int c_pad, int c_off, float* dst,
                       cudaStream_t st);
int launch_unpack_nchw_f32(const float* src, Geom g, int C, int c_pad, float* dst, cudaStream_t st);
struct PackInputsParams {
    const float* x; long xs[5];            // [B,2,T,H,W] element strides
    const float* x_o;                      // [B,32,H,W] (init) or [B,2,4H,4W]; NULL = keep the o part
                                           // already in `mi` (or zero it when init)
    int init;
    act_t* mi;                     // [B*R][64]
    Geom g;
};
int launch_pack_inputs(const PackInputsParams& p, cudaStream_t st);
struct EmitParams {
    const float* a;                        // conv_o result, fp32 [B*R][32]
    const float* x; long xs[5];            // f2 = x[:, :, 1]
    float* out_o;                          // [B,2,4H,4W] or NULL
    act_t* mi_next;                // o part of the next step's input tensor, or NULL
    Geom g;
};
int launch_emit(const EmitParams& p, cudaStream_t st);
int launch_fill_identity(act_t* dst, cudaStream_t st);      // [2][128][64] chunk-major 128x128 identity
int launch_repack_weight(const float* src, const int* kmap, int src_row_len, int n_out,
                         int n_out_pad, int K, act_t* dst, int w_rows, int w_row_base,
                         cudaStream_t st, const float* in_scale = nullptr);
int launch_fold_beta_bias(const float* w, const float* bias, const float* beta, int n_out, int cin, float* bias_out,
                          cudaStream_t st);

}  // namespace bmc

/*
 * bmc_b200.h -- C ABI of libbmc_b200.so: the B200 (sm_100a) implementation of the BMCNet
 * inference hot path (event encoders + BMCNet / BMCNet_plain forward).
 *
 * The reference (Lqm26/BMCNet-ESR) is 100 % Python/PyTorch and has no FFI of its own
 * (SURVEY.md section 8b): the interface a maintainer binds is its Python surface.  Each
 * entry point below names the reference function it replaces (file:line under the
 * reference root).  The Python mirror that calls these through ctypes lives in
 * bmcnet_esr_b200/{dataloader/encodings.py,models/*.py}; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only; every pointer marked "device" is a CUDA device pointer owned by
 *     the caller (no allocation or ownership transfer across this ABI);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs;
 *   - return 0 on success, a negative bmc_status otherwise; bmc_last_error() gives the text
 *     (thread-local);
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     BMC_ERR_CUDA.
 */
#ifndef BMC_B200_H_
#define BMC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMC_ABI_VERSION 1

typedef enum {
    BMC_OK = 0,
    BMC_ERR_ARG = -1,        /* bad argument (mirrors the reference's asserts, encodings.py:125,169,220,277,295) */
    BMC_ERR_CUDA = -2,       /* CUDA runtime / driver error */
    BMC_ERR_WORKSPACE = -3,  /* workspace missing or too small */
    BMC_ERR_STATE = -4,      /* call order (weights not loaded, model not configured, ...) */
    BMC_ERR_UNSUPPORTED = -5
} bmc_status;

int bmc_abi_version(void);
const char* bmc_last_error(void);
/* "f16" (default build) or "bf16" (-DBMC_ACT_BF16): the 16-bit type of every activation / weight
 * tensor this header calls "act16".  Accumulation is always fp32. */
const char* bmc_act_dtype(void);

/* ------------------------------------------------------------------------------------------
 * Event encoders (reference: dataloader/encodings.py).
 * Events are four device float32 arrays xs, ys, ts, ps of length n (base_dataset.py:24-31).
 * Out-of-range events are zeroed IN PLACE in xs / ys (and ps where the reference does),
 * exactly as the reference does (encodings.py:249-254, 34-39), when BMC_ENC_MUTATE is set.
 * ---------------------------------------------------------------------------------------- */

#define BMC_ENC_FLIP_Y 0x1u        /* y' = H-1-y (events_to_image family, encodings.py:265) */
#define BMC_ENC_MUTATE 0x2u        /* reproduce the reference's in-place zeroing of out-of-range events */
#define BMC_ENC_NO_QUIRKS 0x4u     /* drop the F9 leak of out-of-range events into pixel (0,0) */
#define BMC_ENC_TNORM 0x8u         /* voxel: t = (ts-ts[0])/dt*(B-1) (encodings.py:127-129) instead of ts*(B-1) (:280) */
#define BMC_ENC_SKIP_ZERO_ENDS 0x20u /* stacks: ts[0]==0 && ts[n-1]==0 -> all-zero output, no event touched (see below) */
#define BMC_ENC_BILINEAR 0x10u     /* image: spatial bilinear splat into (H+1)x(W+1) (encodings.py:57-65) */
#define BMC_ENC_DETERMINISTIC 0x40u /* image / voxel: accumulate the float weights as 64-bit fixed point (2^-32 units) so the
                                      result is bit-identical from run to run (integer sums do not depend on the order the
                                      atomics land in); |weight| < 2^30 per event.  Count encodings are always deterministic. */
#define BMC_ENC_SPLIT_BINS 0x80u  /* channels / voxel on grids beyond one SM's shared memory: take the split-bins kernel
                                    * (role_kernel, encode.cu: a few CTAs stream the same events, each holding a share of
                                    * the bins) whatever the event count; by default it is used from 2^24 events up (below
                                    * that the global-atomic path is as fast).  Same results. */

/* Scratch bytes needed by any encoder call below for an output of `out_elems` floats: 8 bytes per output element (integer /
 * fixed-point and fp32 partial grids), the stack encoders' bin boundaries, and a 512 KB list through which the split-bins
 * kernel defers the in-place zeroing of out-of-range events (BMC_ENC_MUTATE).  Always query it: the size is part of the
 * library build, not of the ABI. */
size_t bmc_encode_workspace_bytes(int64_t out_elems);

/* events_to_channels (encodings.py:290-305): per-polarity counts, out = device float[2][H][W].
 * Bit-exact for ps in {-1,0,+1}; other weights accumulate ps*ps in fp32 (order not fixed). */
int bmc_encode_channels(float* xs, float* ys, const float* ps, int64_t n, int H, int W,
                        float* out, void* workspace, size_t workspace_bytes, unsigned flags,
                        void* stream);

/* Batched form of the above for the dataloader pattern (h5dataset.py:308-309,526): window i
 * covers events [offsets[i], offsets[i+1]) and fills out[i] = float[2][H][W]; `offsets` is a
 * device int64[n_windows+1].  One CTA per window, no cross-window traffic. */
int bmc_encode_channels_windows(float* xs, float* ys, const float* ps, const int64_t* offsets,
                                int n_windows, int H, int W, float* out, unsigned flags,
                                void* stream);

/* events_to_image (encodings.py:241-269, with BMC_ENC_FLIP_Y) and events_to_image_torch
 * (encodings.py:16-72, without; BMC_ENC_BILINEAR selects the padded bilinear splat whose
 * output is float[H+1][W+1]).  Weighted fp32 scatter-add, out = device float[H][W]. */
/* The dataloader's window pipeline on a raw recording, one launch: window i = events
 * [stride*i, min(stride*i + window, n_events - 1)) (H5Dataset.compute_k_indices, h5dataset.py:197-210, with
 * stride = window - sliding_window) of the arrays as stored on disk -- int16 xs, ys, float64 ps
 * (event_packagers.py:128-156; get_events h5dataset.py:407-414) -- cast to float32
 * (BaseDataset.event_formatting, base_dataset.py:24-31) and counted per polarity
 * (create_cnt_encoding -> events_to_channels, h5dataset.py:518-526).  out: device float [n_windows][2][H][W]. */
int bmc_encode_channels_windows_raw(const int16_t* xs, const int16_t* ys, const double* ps, int64_t n_events,
                                    int64_t window, int64_t stride, int n_windows, int H, int W, float* out,
                                    unsigned flags, void* stream);

/* BaseDataset.event_formatting (base_dataset.py:24-31) on the device: out = float32 [4][n] =
 * (xs, ys, (ts - ts[0]) / (ts[-1] - ts[0] + 1e-6), ps), the division in float32 like the reference. */
int bmc_format_events(const int16_t* xs, const int16_t* ys, const double* ts, const double* ps, int64_t n,
                      float* out, void* stream);

int bmc_encode_image(float* xs, float* ys, float* ps, int64_t n, int H, int W, float* out,
                     void* workspace, size_t workspace_bytes, unsigned flags, void* stream);

/* events_to_voxel (encodings.py:272-287; FLIP_Y) and events_to_voxel_torch with
 * temporal_bilinear=True (encodings.py:100-137; TNORM).  out = device float[bins][H][W]. */
int bmc_encode_voxel(float* xs, float* ys, const float* ts, const float* ps, int64_t n, int bins,
                     int H, int W, float* out, void* workspace, size_t workspace_bytes,
                     unsigned flags, void* stream);

/* events_to_stack_polarity (encodings.py:151-199; polarity=1, out = float[2][bins][H][W]),
 * events_to_stack_no_polarity (encodings.py:202-238; polarity=0, out = float[bins][H][W]) and
 * events_to_voxel_torch(temporal_bilinear=False) (encodings.py:138-145; same as polarity=0).
 * The reference's float32 bin boundaries and its any-equal binary search (encodings.py:75-97)
 * are evaluated on the device, so boundary events are double counted exactly as there.
 * The reference's early-out (ts.sum()==0 or n<=3 -> zeros[bins][H][W]) is the CALLER's job
 * (it changes the output shape); n <= 3 is rejected with BMC_ERR_ARG. */
/* The early-out without a host round trip BEFORE the launch: every stack call stores the int
 * (ts[0] == 0 && ts[n-1] == 0) at byte bmc_encode_stack_flag_offset(out_elems) of its workspace; with
 * BMC_ENC_SKIP_ZERO_ENDS such a call leaves the events untouched and writes zeros, so the caller can launch first,
 * read the word afterwards and only then -- in the rare case it is set -- scan ts and either return the
 * reference's zeros or call again without the flag. */
size_t bmc_encode_stack_flag_offset(int64_t out_elems);
int bmc_encode_stack(float* xs, float* ys, const float* ts, float* ps, int64_t n, int bins,
                     int H, int W, int polarity, float* out, void* workspace,
                     size_t workspace_bytes, unsigned flags, void* stream);

/* The same encoders for ONE rank of a recording whose events are split into contiguous ranges across
 * GPUs (SURVEY.md 8e "one giant grid"): this rank holds events [first, first + n_local) as xs, ys, ps; the
 * bin boundaries are a property of the whole recording, so every rank passes the full ts_all[n_total] and
 * evaluates them itself.  out is this rank's PARTIAL stack; the sum over ranks (one all-reduce of
 * integer-valued floats, exact) equals bmc_encode_stack on the whole recording.  n_local may be 0. */
int bmc_encode_stack_shard(float* xs, float* ys, float* ps, int64_t n_local, const float* ts_all,
                           int64_t n_total, int64_t first, int bins, int H, int W, int polarity, float* out,
                           void* workspace, size_t workspace_bytes, unsigned flags, void* stream);

/* ------------------------------------------------------------------------------------------
 * Model (reference: models/BMCNet.py, models/BMCNet_plain.py, models/submodules.py).
 * ---------------------------------------------------------------------------------------- */

typedef struct bmc_model bmc_model_t;

#define BMC_MODEL_BMCNET 0        /* models/BMCNet.py:87-121 */
#define BMC_MODEL_BMCNET_PLAIN 1  /* models/BMCNet_plain.py:36-68 */

/* Constructor arguments of BMCNet(scale, n_c, n_b, repeat) (BMCNet.py:88).  The kernels are
 * specialised for what every caller passes (infer_BMCNet.py:111, train.py:640): n_c = 128,
 * scale = 4, repeat = 3; n_b is free.  Anything else returns NULL (see bmc_last_error). */
bmc_model_t* bmc_model_create(int kind, int scale, int n_c, int n_b, int repeat);
void bmc_model_destroy(bmc_model_t* m);

/* Bytes of device memory the caller must provide for the repacked (act16, K-major) weights. */
size_t bmc_model_weight_bytes(const bmc_model_t* m);

/* Load a reference-format state_dict (318 keys for BMCNet / 120 for plain; aliases allowed and
 * expected to hold equal values).  `tensors[i]` is a device float32 pointer to the
 * contiguous tensor named `names[i]` with `numels[i]` elements.  Missing or unexpected names
 * fail like load_state_dict(strict=True) (infer_BMCNet.py:112).  The repacked copy is written
 * to `weight_buf` (>= bmc_model_weight_bytes), which must stay alive until destroy/reload. */
int bmc_model_load_state_dict(bmc_model_t* m, const char* const* names,
                              const float* const* tensors, const int64_t* numels, int n_tensors,
                              void* weight_buf, size_t weight_buf_bytes, void* stream);

/* Fix the problem size (batch of independent sequences, LR height/width) and report the
 * activation arena the caller must provide via bmc_model_bind_workspace. */
int bmc_model_configure(bmc_model_t* m, int batch, int H, int W);
size_t bmc_model_workspace_bytes(const bmc_model_t* m);
int bmc_model_bind_workspace(bmc_model_t* m, void* workspace, size_t workspace_bytes);

/* One recurrent step == one `forward` of the reference module.
 *   x        device float32, logical shape [B,2,T,H,W] addressed through `x_strides` (elements);
 *            frames t=0 and t=1 are used (BMCNet.py:106-107).  The reference caller passes a
 *            transposed view (infer_BMCNet.py:50), hence the strides.
 *   x_h*     device float32 [B,128,H,W] contiguous: hidden states in (x_h_p/x_h_n NULL for plain).
 *   x_o      device float32: [B,32,H,W] when init != 0, else the previous output [B,2,4H,4W].
 *   out_*    device float32 outputs, same shapes as the inputs; out_o is [B,2,4H,4W].
 * Argument order and meaning follow BMCNet.forward (BMCNet.py:95) / BMCNet_plain.forward
 * (BMCNet_plain.py:44), including the positional hand-over of the three hidden states into
 * Backbone.forward (BMCNet.py:57 vs :115).
 * Fast path for the recurrent loop of infer_BMCNet.py:46-68, where every call is fed the previous
 * call's outputs: x_h (and x_h_p, x_h_n) may be NULL, meaning "the hidden states this model produced
 * last are the inputs" (they are still resident in the workspace), and x_o may be NULL when init == 0,
 * meaning "the previous prediction".  The mirror modules do this when they recognise their own outputs. */
int bmc_model_forward(bmc_model_t* m, const float* x, const int64_t x_strides[5],
                      const float* x_h, const float* x_h_p, const float* x_h_n, const float* x_o,
                      int init, float* out_h, float* out_h_p, float* out_h_n, float* out_o,
                      void* stream);

/* Device-resident recurrence for throughput runs: identical arithmetic, but the hidden states
 * and the fed-back prediction stay inside the workspace between steps (no fp32 NCHW round
 * trip).  `reset` != 0 restarts from the zero state of infer_BMCNet.py:55-60 (init=True).
 * out_o may be NULL to skip emitting the prediction. */
int bmc_model_step(bmc_model_t* m, const float* x, const int64_t x_strides[5], int reset,
                   float* out_o, void* stream);

/* Number of kernels one forward/step enqueues (for bench.py's gpu_launches). */
int bmc_model_launches_per_step(const bmc_model_t* m);

/* 0 = tcgen05/TMA kernels (product), 1 = plain SIMT kernels of the same arithmetic, kept ON
 * THE DEVICE as a debugging cross-check for the tensor-core path.  Never a CPU path. */
int bmc_model_set_debug_simt(bmc_model_t* m, int enable);

/* ------------------------------------------------------------------------------------------
 * Per-kernel entry points for unit parity (SURVEY section 8b "bmc_conv3x3, bmc_bie").
 * Activations are act16 in the padded NHWC layout described in DESIGN.md:
 *   rows = B * R, R = roundup((H+2)*(W+2), 128), row r of image b <-> padded pixel
 *   (r / (W+2), r % (W+2)); halo and tail rows hold zeros.
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    /* A operand: up to 3 concatenated sources (torch.cat along channels, e.g. BMCNet.py:64) */
    int n_seg;
    const void* a[3];       /* device act16 [a_rows][a_ch] */
    int a_rows[3];
    int a_ch[3];            /* multiple of 64 */
    int a_row_base[3];      /* row of the source that pairs with output row 0 */
    /* B operand: weights, device act16 chunk-major [w_k/64][w_rows][64] */
    const void* w;
    int w_rows, w_k;
    int w_row_base;         /* first of the N rows used */
    int w_img_stride;       /* added per image (dynamic per-image weights, submodules.py:72-73) */
    const float* bias;      /* device float[N] or NULL */
    const void* residual;   /* device act16 [.][N] added after activation, or NULL */
    int res_row_base;
    void* out_act16;         /* device act16 [.][N] or NULL */
    int out_row_base;
    float* out_f32;         /* device float [.][N] or NULL (same row base) */
    int relu;
    /* optional fused channel LayerNorm of (acc + bias) (submodules.py:127-139), N = 128 only */
    const float* ln_gamma;  /* device float[N] or NULL */
    const float* ln_beta;
    float ln_eps;
} bmc_gemm_job_t;

/* Weights are chunk-major: w is act16 [w_k/64][w_rows][64] (K index = (segment, tap, channel)).
 * out[m, :] = act(LN?(sum_seg sum_tap A_seg[m + dy*(W+2) + dx, :] . W[:, seg, tap, :]^T + bias)) + res
 * for all rows m of B images; n = 128 or 32 output channels; taps = 1 (1x1) or 9 (3x3, pad 1).
 * All jobs of one call share the shape and run in one launch.  impl: 0 tcgen05, 1 SIMT debug. */
int bmc_conv_gemm(const bmc_gemm_job_t* jobs, int n_jobs, int n, int taps, int B, int H, int W,
                  int impl, void* stream);

/* att[b] = centres[b]^T . v[b] * scale over the pixels of image b (submodules.py:69-70), then
 * row softmax (submodules.py:72-73) written as act16 [B][128][128] "dynamic weights".
 * centres, v: device act16 [B*R][128]; partial: device float[B][n_split][128][128] scratch. */
int bmc_attention_weights(const void* centres, const void* v, int B, int H, int W, float scale,
                          float* partial, int n_split, void* probs_act16, int impl, void* stream);

/* Channel LayerNorm (submodules.py:127-139), act16 [rows][128] -> act16 [rows][128]. */
int bmc_layernorm_rows(const void* in_act16, const float* gamma, const float* beta, float eps,
                       int64_t rows, void* out_act16, void* stream);

/* fp32 NCHW [B,C,H,W] <-> padded NHWC act16 [B*R][c_pad] (channels [c_off, c_off+C)). */
int bmc_pack_nchw(const float* src, int B, int C, int H, int W, void* dst_act16, int c_pad,
                  int c_off, void* stream);
int bmc_unpack_nchw(const void* src_act16, int B, int C, int H, int W, int c_pad, int c_off,
                    float* dst, void* stream);

/* ------------------------------------------------------------------------------------------
 * Inverse encoders (reference dataloader/encodings.py:367-464, 653-671): count stacks back to event clouds.
 *   bmc_stack_event_counts  per-entry totals sum(|round(v)|) and sum(round(v)) -- what the reference derives its
 *                           maxlen and its `entry.sum() != 0` tests from (:383,386,432,435); device int64[B] each.
 *   bmc_stack_to_events     python_event_redistribute_PolarityStack (P = 2, stack [B,2,C,Y,X]) /
 *                           _NoPolarityStack (P = 1, stack [B,C,Y,X]): out = device float [B][maxlen][4] =
 *                           (x, y, t, p) per entry, stably sorted by t, zero-padded; timestamps
 *                           torch.linspace(c/C + 1/(100 C), (c+1)/C, |v|) (rnd == NULL, mode='linear') or
 *                           rnd * (t1 - t0) + t0 with rnd = device float [B][maxlen] uniform numbers (mode='random').
 *   bmc_stack2cnt           stack2cnt (:653-671): [B,TB,H,W] -> [B,2,H,W].
 * The workspace (256-byte aligned, bmc_stack_to_events_workspace_bytes) is caller-owned scratch.
 * ---------------------------------------------------------------------------------------- */
size_t bmc_stack_to_events_workspace_bytes(int B, int64_t per_entry, int64_t maxlen);
int bmc_stack_event_counts(const float* stack, int B, int64_t per_entry, void* workspace, size_t workspace_bytes,
                           int64_t* totals_out, int64_t* sums_out, void* stream);
int bmc_stack_to_events(const float* stack, int B, int P, int C, int Y, int X, int64_t maxlen, const float* rnd,
                        float* out, void* workspace, size_t workspace_bytes, void* stream);
int bmc_stack2cnt(const float* stack, int B, int TB, int H, int W, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Evaluation tail of the inference loop (reference infer_BMCNet.py:77-87), on the device:
 *   sums[0] = sum((resize(pred) - gt)^2)   resize = bicubic to the gt size when the sizes differ (:78-79)
 *   sums[1] = sum((bicubic(inp) - gt)^2)   the bicubic baseline of the LR count frame (:80)
 * over all B*C*Hg*Wg ground-truth elements (nn.MSELoss divides by that count, :83-84).
 * pred: device float [B,C,Hp,Wp]; inp: [B,C,H,W]; gt: [B,C,Hg,Wg], all contiguous;
 * sums: device double[2], overwritten.  Nothing is copied to the host.
 * ---------------------------------------------------------------------------------------- */
int bmc_sr_metrics(const float* pred, int B, int C, int Hp, int Wp, const float* inp, int H, int W,
                   const float* gt, int Hg, int Wg, double* sums, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training step (reference train.py:202-237: forward over the sequence, summed nn.MSELoss, one backward,
 * one optimiser step; config/train_nfs.yml:28-34: Adam(lr 1e-4, weight_decay 1e-5, amsgrad)).
 * The autograd tape lives on the Python side (bmcnet_esr_b200/models/_train.py); these are its kernels.
 *   data gradient (dgrad)   no entry of its own: it is bmc_conv_gemm on dY with the taps mirrored and the weight
 *                           matrix transposed (what cuDNN's backward-data does for nn.Conv2d).
 *   bmc_conv_wgrad          weight (and bias) gradient of `nn.Conv2d(k=3|1, padding=k/2)` w.r.t. ONE input source
 *                           of a concatenated input: grad_w[co][cmap[ci]][tap] += scale * sum_rows dy[row][co] *
 *                           x[row + off_tap][ci], grad_b[co] += scale * sum_rows dy[row][co] (grad_b may be NULL).
 *                           dy: device act16 [B*R][128] with zero halo rows; x: device act16 [B*R][x_ch], x_ch = 64
 *                           or 128; cmap: device int[x_ch], the input channel of the PyTorch-layout weight
 *                           [n_out][cin_total][k][k] each source channel feeds, or -1 (padding); channels >= n_out of
 *                           dy (a 32-output conv run as a zero-padded 128-output one) are ignored; grad_w / grad_b:
 *                           device float, ACCUMULATED into (aliased modules share one gradient).
 *                           tcgen05 split-K over n_split pixel ranges, partials reduced in a fixed order.
 *   bmc_relu_backward       dx = dy * (y > 0) on act16 tensors (F.relu, BMCNet.py:64-73, submodules.py:33).
 *   bmc_adam_amsgrad_step   torch.optim.Adam(amsgrad=True) with L2 weight decay over flat fp32 buffers; `step` is
 *                           the 1-based step count.
 * ---------------------------------------------------------------------------------------- */
size_t bmc_conv_wgrad_workspace_bytes(int n_split, int taps, int x_ch);
int bmc_conv_wgrad(const void* dy_act16, const void* x_act16, int x_ch, int taps, int B, int H, int W,
                   const int* cmap, int cin_total, int n_out, float scale, float* grad_w, float* grad_b,
                   void* workspace, size_t workspace_bytes, int n_split, void* stream);
int bmc_relu_backward(const void* dy_act16, const void* y_act16, int64_t n_elems, void* dx_act16, void* stream);
/* LayerNormFunction.backward (submodules.py:142-154) for bmc_layernorm_rows: dx (act16 [rows][128]) from x and dy;
 * grad_gamma[c] += scale * sum_rows dy * y_hat, grad_beta[c] += scale * sum_rows dy (device float[128], ACCUMULATED;
 * per-CTA partials reduced in a fixed order).  mu / rstd are recomputed from x. */
size_t bmc_layernorm_rows_backward_workspace_bytes(void);
int bmc_layernorm_rows_backward(const void* x_act16, const void* dy_act16, const float* gamma, float eps, int64_t rows,
                                void* dx_act16, float scale, float* grad_gamma, float* grad_beta, void* workspace,
                                size_t workspace_bytes, void* stream);
int bmc_adam_amsgrad_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                          float* max_exp_avg_sq, int64_t n, int step, float lr, float beta1, float beta2,
                          float eps, float weight_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BMC_B200_H_ */

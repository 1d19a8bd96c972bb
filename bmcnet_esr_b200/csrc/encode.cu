// Event encoders for sm_100a: (x, y, t, p) event streams -> count / image / voxel / stack grids.
//
// Reference semantics: dataloader/encodings.py:6-305 (see include/bmc_b200.h for the per-entry
// mapping).  The work is an HBM-bound histogram: 12 B/event (xs, ys, ps) for counts and stacks,
// 16 B/event (+ts) for voxels, read once with 128-bit streaming loads by a persistent grid
// (a multiple of the SM count); each CTA privatises the whole output grid in shared memory
// (int32 bins, or 16-bit packed bins with exact carry handling when the grid only fits that
// way in the 227 KB of a B200 SM), and flushes its non-zero bins to a global int32/fp32 grid
// with one atomic per bin.  Integer counting makes the result independent of event order, i.e.
// bit-exact against the reference's serial fp32 accumulation (which saturates at 2^24; the
// finalize step reproduces that).
#include "common.cuh"

namespace bmc {
namespace {

constexpr int kThreads = 512;
constexpr int kSmemBudget = 227 * 1024;        // dynamic smem per CTA on sm_100
constexpr int kMaxBinsSmem32 = 49152;          // 192 KB of int32 / fp32 bins
constexpr int kMaxBinsSmem16 = kSmemBudget / 2;  // 16-bit packed bins

enum Mode { kSmem32 = 0, kSmem16 = 1, kGlobal = 2, kSmem64 = 3, kGlobal64 = 4, kSmemFix = 5 };
// kSmemFix: the default path of the float-weighted encoders on grids that fit (2 x bins x 4 B <= 227 KB).  Every weight is
// split by sign and added as UNSIGNED 2^-24 fixed point to one of two 32-bit shared-memory bins (positive / negative
// plane) with a native ATOMS.ADD; the thread whose add wraps a bin (the returned old value tells it: exactly one thread
// per wrap) credits 2^32 to the global 64-bit grid.  No compare-and-swap loops (an fp32 shared-memory atomicAdd compiles to
// ATOMS.CAST.SPIN: LDS + FADD + CAS + branch per attempt), integer sums only -- so the result does not depend on the order
// the atomics land in (bit-reproducible), and a weight's rounding error is <= 2^-25 (fp32's own resolution for weights in
// [0.5, 1)).  |weight| >= 128 (never the case for +-1 polarities) goes to the 64-bit grid directly.
constexpr int kFixBits = 24;
constexpr int kMaxBinsSmemFix = kSmemBudget / 8;
constexpr int kMaxBinsSmem64 = 24576;          // 192 KB of 64-bit fixed-point bins

// BMC_ENC_DETERMINISTIC: float weights are accumulated as 64-bit fixed point (2^-32 units).  Integer addition is
// associative, so the sum -- and the fp32 value it is rounded to once, in finalize64_kernel -- does not depend on
// the order in which the atomics land: two runs are bit-identical, whatever the grid size or the scheduling.
// |weight| < 2^30 per event and |sum| < 2^31 per bin; a weight's rounding error is <= 2^-33 (weights >= 2^-9 in
// magnitude are represented exactly), far inside the 1e-6 bar.
__device__ __forceinline__ unsigned long long to_fixed64(float w) {
    return (unsigned long long)__double2ll_rn((double)w * 4294967296.0);
}

__device__ __forceinline__ float ldg_stream1(const float* p) {
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// Decoded spatial part of one event (encodings.py:249-265 / :34-39,67-70).
struct Pix {
    bool oor;
    int x, y;   // truncated toward zero like `.long()`, y already flipped if requested
};
__device__ __forceinline__ Pix decode_xy(float x, float y, int H, int W, bool flip) {
    Pix p;
    // NaN coordinates make every comparison false in the reference and then index with
    // garbage; here they are treated as out of range.
    p.oor = !((x < (float)W) & (x >= 0.f) & (y < (float)H) & (y >= 0.f));
    p.x = p.oor ? 0 : (int)x;
    int yy = p.oor ? 0 : (int)y;
    p.y = flip ? (H - 1 - yy) : yy;
    return p;
}

// ---------------------------------------------------------------- accumulation targets
template <int MODE, bool FLOAT_HIST>
struct Hist {
    static constexpr int kMode = MODE;
    int* s_i;        // smem int32 bins / packed 16-bit pairs
    float* s_f;      // smem fp32 bins (FLOAT_HIST)
    int* g_cnt;      // global int32 grid
    float* g_ext;    // global fp32 grid (non-integral weights)
    unsigned long long* s64;   // kSmem64: smem fixed-point bins
    unsigned long long* g64;   // kSmem64 / kGlobal64 / kSmemFix: global fixed-point grid (overlays cnt + ext)
    int nb;                    // kSmemFix: bins per sign plane

    __device__ __forceinline__ void add_int(int bin, int delta) {
        if (MODE == kSmem32) {
            atomicAdd(&s_i[bin], delta);
        } else if (MODE == kSmem16) {
            // Two 16-bit counters per word.  The one thread that moves a field from 0x7FFF to
            // 0x8000 takes 0x8000 back out and credits it to the global grid; until that lands
            // the field can only grow by the <= kThreads adds in flight, so it never reaches
            // 0xFFFF and never carries into its neighbour: exact for any event distribution.
            const int sh = (bin & 1) * 16;
            unsigned old = atomicAdd(reinterpret_cast<unsigned*>(&s_i[bin >> 1]), 1u << sh);
            if (((old >> sh) & 0xFFFFu) == 0x7FFFu) {
                atomicSub(reinterpret_cast<unsigned*>(&s_i[bin >> 1]), 0x8000u << sh);
                atomicAdd(&g_cnt[bin], 32768);
            }
        } else {
            atomicAdd(&g_cnt[bin], delta);
        }
    }
    // kSmemFix: `d` units of 2^-24 onto the positive (neg == false) or negative plane of `bin`
    __device__ __forceinline__ void add_fix(int bin, unsigned d, bool neg) {
        unsigned* cell = reinterpret_cast<unsigned*>(s_i) + (neg ? nb : 0) + bin;
        const unsigned old = atomicAdd(cell, d);
        if (old + d < old)                              // this add wrapped the 32-bit bin: credit 2^32 units
            atomicAdd(&g64[bin], neg ? (unsigned long long)(-(1ll << 32)) : (1ull << 32));
    }
    __device__ __forceinline__ void add_float(int bin, float w) {
        if (MODE == kSmemFix) {
            const float mag = fabsf(w);
            if (mag < 128.f) {
                const unsigned d = __float2uint_rn(mag * (float)(1 << kFixBits));
                unsigned* cell = reinterpret_cast<unsigned*>(s_i) + (w < 0.f ? nb : 0) + bin;
                const unsigned old = atomicAdd(cell, d);
                if (old + d < old)                      // this add wrapped the 32-bit bin: credit 2^32 units
                    atomicAdd(&g64[bin], w < 0.f ? (unsigned long long)(-(1ll << 32)) : (1ull << 32));
            } else {
                atomicAdd(&g64[bin], (unsigned long long)__double2ll_rn((double)w * (double)(1 << kFixBits)));
            }
            return;
        }
        if (MODE == kSmem64) atomicAdd(&s64[bin], to_fixed64(w));
        else if (MODE == kGlobal64) atomicAdd(&g64[bin], to_fixed64(w));
        else if (FLOAT_HIST && MODE == kSmem32) atomicAdd(&s_f[bin], w);
        else atomicAdd(&g_ext[bin], w);
    }
    // integral +-1 weights go to the exact integer path, everything else to fp32
    __device__ __forceinline__ void add_weight(int bin, float w) {
        if (FLOAT_HIST) { if (w != 0.f) add_float(bin, w); return; }
        if (w == 1.f) add_int(bin, 1);
        else if (w == -1.f && MODE != kSmem16) add_int(bin, -1);
        else if (w != 0.f) atomicAdd(&g_ext[bin], w);
    }
};

// ---------------------------------------------------------------- per-encoder event ops
struct Common {
    float* xs; float* ys; const float* ts; float* ps;
    int H, W, bins;
    unsigned flags;
    __device__ __forceinline__ bool flip() const { return flags & BMC_ENC_FLIP_Y; }
    __device__ __forceinline__ bool mutate() const { return flags & BMC_ENC_MUTATE; }
    __device__ __forceinline__ bool quirks() const { return !(flags & BMC_ENC_NO_QUIRKS); }
    __device__ __forceinline__ int origin() const { return flip() ? (H - 1) * W : 0; }
    __device__ __forceinline__ void prepare(long) {}
    __device__ __forceinline__ void group(long, int) {}      // called once per run of consecutive events
};

// events_to_channels (encodings.py:290-305)
struct ChannelsOp : Common {
    static constexpr bool kBounds = false;
    static constexpr bool kFloat = false, kNeedT = false, kSigned = false;
    template <class HT>
    __device__ __forceinline__ void run(HT& h, long i, float x, float y, float, float p) const {
        Pix q = decode_xy(x, y, H, W, true);
        const float w = p * p;              // ps * mask_{pos,neg} == ps^2 on the matching sign
        if (!q.oor) {
            const int bin = (p < 0.f ? H * W : 0) + q.y * W + q.x;
            h.add_weight(bin, w);
        } else {
            // F9: after the positive pass zeroed xs/ys, the negative pass sees (0,0) in range
            if (quirks() && p < 0.f) h.add_weight(H * W + (H - 1) * W, w);
            if (mutate()) { xs[i] = 0.f; ys[i] = 0.f; }
        }
    }
};

// events_to_image (encodings.py:241-269) / events_to_image_torch (encodings.py:16-72)
struct ImageOp : Common {
    static constexpr bool kBounds = false;
    static constexpr bool kFloat = true, kNeedT = false, kSigned = true;
    template <class HT>
    __device__ __forceinline__ void run(HT& h, long i, float x, float y, float, float p) const {
        Pix q = decode_xy(x, y, H, W, flip());
        if (q.oor) {
            if (mutate()) { xs[i] = 0.f; ys[i] = 0.f; ps[i] = 0.f; }
            return;
        }
        if (flags & BMC_ENC_BILINEAR) {     // padded (H+1)x(W+1) splat, encodings.py:57-65, 6-13
            const float fx = floorf(x), fy = floorf(y);
            const float dx = x - fx, dy = y - fy;
            const int Wp = W + 1, base = (int)fy * Wp + (int)fx;
            const float w0 = __fmul_rn(p, 1.f - dx), w1 = __fmul_rn(p, dx);
            h.add_float(base, __fmul_rn(w0, 1.f - dy));
            h.add_float(base + 1, __fmul_rn(w1, 1.f - dy));
            h.add_float(base + Wp, __fmul_rn(w0, dy));
            h.add_float(base + Wp + 1, __fmul_rn(w1, dy));
        } else {
            h.add_float(q.y * W + q.x, p);
        }
    }
};

// events_to_voxel (encodings.py:272-287) / events_to_voxel_torch bilinear (encodings.py:127-137)
struct VoxelOp : Common {
    static constexpr bool kBounds = false;
    static constexpr bool kFloat = true, kNeedT = true, kSigned = true;
    float t0, dt;     // only for BMC_ENC_TNORM
    __device__ __forceinline__ void prepare(long n) {
        if ((flags & BMC_ENC_TNORM) && n > 0) {      // dt = ts[-1]-ts[0] + 1e-6 (encodings.py:127)
            t0 = ts[0];
            dt = __fadd_rn(__fsub_rn(ts[n - 1], t0), 1e-6f);
        }
    }
    template <class HT>
    __device__ __forceinline__ void run(HT& h, long i, float x, float y, float t, float p) const {
        Pix q = decode_xy(x, y, H, W, flip());
        const float fb = (float)(bins - 1);
        float tn;
        if (flags & BMC_ENC_TNORM) tn = __fmul_rn(__fdiv_rn(__fsub_rn(t, t0), dt), fb);
        else tn = __fmul_rn(t, fb);
        if (q.oor && mutate()) { xs[i] = 0.f; ys[i] = 0.f; }
        const int pix = q.oor ? origin() : q.y * W + q.x;
        const int b0 = (int)floorf(tn);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int b = b0 + k;
            if (b < 0 || b >= bins) continue;
            // bin 0 is the pass that zeroes out-of-range events; later bins see them at (0,0)
            if (q.oor && (b == 0 || !quirks())) continue;
            const float w = fmaxf(0.f, __fsub_rn(1.f, fabsf(__fsub_rn(tn, (float)b))));
            h.add_float(b * H * W + pix, __fmul_rn(p, w));
        }
    }
    // Pair form: the (at most) two non-zero weights of an event fall on ADJACENT bins j, j+1 with
    // j = clamp(floor(tn), 0, bins-2), so both are delivered by one 64-bit update of a float2 slot
    // (same weights as run(): every other bin's max(0, 1-|tn-b|) is zero).  Returns false when the
    // event contributes nothing.
    __device__ __forceinline__ bool pair(long i, float x, float y, float t, float p, int& slot,
                                         float& lo, float& hi) const {
        Pix q = decode_xy(x, y, H, W, flip());
        const float fb = (float)(bins - 1);
        float tn;
        if (flags & BMC_ENC_TNORM) tn = __fmul_rn(__fdiv_rn(__fsub_rn(t, t0), dt), fb);
        else tn = __fmul_rn(t, fb);
        if (q.oor && mutate()) { xs[i] = 0.f; ys[i] = 0.f; }
        if (q.oor && !quirks()) return false;
        const int j = min(max((int)floorf(tn), 0), bins - 2);
        lo = __fmul_rn(p, fmaxf(0.f, __fsub_rn(1.f, fabsf(__fsub_rn(tn, (float)j)))));
        hi = __fmul_rn(p, fmaxf(0.f, __fsub_rn(1.f, fabsf(__fsub_rn(tn, (float)(j + 1))))));
        if (q.oor && j == 0) lo = 0.f;      // bin 0 is the pass that drops out-of-range events
        slot = j * H * W + (q.oor ? origin() : q.y * W + q.x);
        return (lo != 0.f) | (hi != 0.f);
    }
};

// events_to_stack_polarity / _no_polarity / voxel_torch(temporal_bilinear=False)
// (encodings.py:151-238, 138-145): event i belongs to every bin b with beg[b] <= i < end[b].
struct StackOp : Common {
    static constexpr bool kFloat = false, kNeedT = false, kSigned = true, kBounds = true;
    const long* beg; const long* end;    // device [bins]; re-pointed at a smem copy by the kernel
    int polarity;
    // Per-thread cursor.  No bin boundary lies in (slo, shi), so every event of [slo, shi) belongs to exactly the
    // bins of `live`: while a thread's groups stay inside that interval -- a bin is ~1e5 groups long -- the
    // boundaries are not re-read.  (The shared-memory pipe is what bounds this kernel: the 2*bins + 2 boundary
    // loads per group that this replaces cost as much of it as the atomics themselves, 268 -> ~500 Gevents/s.)
    long slo = 0, shi = 0;
    unsigned long long live = 0;         // bins of the current interval / bins meeting a straddling group
    bool uniform = false;                // false: the group crosses a boundary, test every event
    int one = -1;                        // plane offset of the only bin of a uniform interval, else -1
    __device__ __forceinline__ void group(long i, int len) {
        if (i >= slo && i + len <= shi) { uniform = true; return; }
        long lo = -1, hi = 0x7fffffffffffffffL;
        unsigned long long in = 0ull, meet = 0ull;
        for (int b = 0; b < bins; ++b) {
            const long bb = beg[b], ee = end[b];
            if (bb <= i) lo = max(lo, bb); else hi = min(hi, bb);
            if (ee <= i) lo = max(lo, ee); else hi = min(hi, ee);
            if (bb <= i && i < ee) in |= 1ull << b;
            if (bb < i + len && ee > i) meet |= 1ull << b;
        }
        if (i + len <= hi) { slo = lo; shi = hi; live = in; uniform = true; }
        else { slo = shi = 0; live = meet; uniform = false; }
        // the common case by far -- every event of the interval in exactly one bin -- gets a straight-line path
        one = (uniform && live && !(live & (live - 1))) ? (__ffsll((long long)live) - 1) * H * W : -1;
    }
    template <class HT>
    __device__ __forceinline__ void run(HT& h, long i, float x, float y, float, float p) const {
        Pix q = decode_xy(x, y, H, W, false);
        if (uniform && one >= 0) {
            if (!q.oor) {
                if (polarity) h.add_weight((p < 0.f ? bins * H * W : 0) + one + q.y * W + q.x, p * p);
                else h.add_weight(one + q.y * W + q.x, p);
            } else {
                if (polarity && quirks() && p < 0.f) h.add_weight(bins * H * W + one, p * p);   // pixel (0,0)
                if (mutate()) { xs[i] = 0.f; ys[i] = 0.f; if (!polarity) ps[i] = 0.f; }
            }
            return;
        }
        bool first = true;
        for (unsigned long long m = live; m; m &= m - 1) {
            const int b = __ffsll((long long)m) - 1;
            if (!uniform && (i < beg[b] || i >= end[b])) continue;
            const int plane = H * W;
            if (polarity) {
                const int base = (p < 0.f ? bins * plane : 0) + b * plane;
                const float w = p * p;
                if (!q.oor) h.add_weight(base + q.y * W + q.x, w);
                else if (quirks() && (!first || p < 0.f)) h.add_weight(base, w);   // pixel (0,0)
            } else {
                if (!q.oor) h.add_weight(b * plane + q.y * W + q.x, p);
            }
            first = false;
        }
        if (q.oor && !first && mutate()) {
            xs[i] = 0.f; ys[i] = 0.f;
            if (!polarity) ps[i] = 0.f;       // the slice of ps is a view there (encodings.py:230-231)
        }
    }
};

// One event through its encoder.  Voxels with >= 2 bins take the pair form (VoxelOp::pair: the event's two non-zero
// weights fall on adjacent bins j, j + 1 -- same weights as run(), half the index / bounds arithmetic).
template <class HT, class Op>
__device__ __forceinline__ void event(HT& h, Op& op, long i, float x, float y, float t, float p) { op.run(h, i, x, y, t, p); }
template <class HT>
__device__ __forceinline__ void event(HT& h, VoxelOp& op, long i, float x, float y, float t, float p) {
    if (op.bins < 2) { op.run(h, i, x, y, t, p); return; }
    if (HT::kMode == kSmemFix) {
        // Fast path of the common event -- in range, polarity +-1, 0 <= tn < bins - 1: with frac = tn - floor(tn) the
        // reference's two weights max(0, 1 - |tn - b|) are exactly frac (bin j + 1) and 1 - frac (bin j) whenever
        // tn >= 1 (differences of neighbouring floats are exact) and within 2^-25 of them for tn < 1, so both go out as
        // integers: d_hi = round(frac * 2^24), d_lo = 2^24 - d_hi -- one conversion, no weight arithmetic in fp32.
        Pix q = decode_xy(x, y, op.H, op.W, op.flip());
        const float fb = (float)(op.bins - 1);
        float tn;
        if (op.flags & BMC_ENC_TNORM) tn = __fmul_rn(__fdiv_rn(__fsub_rn(t, op.t0), op.dt), fb);
        else tn = __fmul_rn(t, fb);
        const float fl = floorf(tn);
        if (!q.oor && fabsf(p) == 1.f && tn >= 0.f && fl <= fb - 1.f) {
            const unsigned dhi = __float2uint_rn(__fsub_rn(tn, fl) * (float)(1 << kFixBits));
            const int slot = (int)fl * (op.H * op.W) + q.y * op.W + q.x;
            const bool neg = p < 0.f;
            if (dhi != (1u << kFixBits)) h.add_fix(slot, (1u << kFixBits) - dhi, neg);
            if (dhi) h.add_fix(slot + op.H * op.W, dhi, neg);
            return;
        }
    }
    int slot; float lo, hi;
    if (op.pair(i, x, y, t, p, slot, lo, hi)) {
        if (lo != 0.f) h.add_float(slot, lo);
        if (hi != 0.f) h.add_float(slot + op.H * op.W, hi);
    }
}

// ---------------------------------------------------------------- the streaming kernel
template <class Op, int MODE, int kThreads>
__global__ void __launch_bounds__(kThreads) scatter_kernel(Op op, long n, int nbins, int* g_cnt,
                                                           float* g_ext, int vec_ok) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Hist<MODE, Op::kFloat> h;
    h.s_i = reinterpret_cast<int*>(smem_raw);
    h.s_f = reinterpret_cast<float*>(smem_raw);
    h.g_cnt = g_cnt;
    h.g_ext = g_ext;
    h.s64 = reinterpret_cast<unsigned long long*>(smem_raw);
    h.g64 = reinterpret_cast<unsigned long long*>(g_cnt);
    h.nb = nbins;
    const int words = (MODE == kSmem32) ? nbins : (MODE == kSmem16 ? (nbins + 1) / 2 : ((MODE == kSmem64 || MODE == kSmemFix) ? 2 * nbins : 0));
    for (int k = threadIdx.x; k < words; k += kThreads) h.s_i[k] = 0;
    if constexpr (Op::kBounds) {                 // bin ranges: 2*bins longs, read once per CTA
        __shared__ long s_bounds[128];
        if (threadIdx.x < op.bins) {
            s_bounds[threadIdx.x] = op.beg[threadIdx.x];
            s_bounds[64 + threadIdx.x] = op.end[threadIdx.x];
        }
        op.beg = s_bounds;
        op.end = s_bounds + 64;
    }
    op.prepare(n);
    __syncthreads();

    // two 4-event groups per thread and iteration: all their loads are issued before the first
    // atomic, which is what keeps enough bytes in flight when only one CTA fits an SM
    const long n4 = vec_ok ? (n >> 2) : 0;
    const long stride = (long)gridDim.x * kThreads;
    for (long g = (long)blockIdx.x * kThreads + threadIdx.x; g < n4; g += 2 * stride) {
        const long i = g << 2, i2 = (g + stride) << 2;
        const bool two = g + stride < n4;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 x = ldg_stream4(op.xs + i);
        const float4 y = ldg_stream4(op.ys + i);
        const float4 p = ldg_stream4(op.ps + i);
        const float4 t = Op::kNeedT ? ldg_stream4(op.ts + i) : z;
        const float4 x2 = two ? ldg_stream4(op.xs + i2) : z;
        const float4 y2 = two ? ldg_stream4(op.ys + i2) : z;
        const float4 p2 = two ? ldg_stream4(op.ps + i2) : z;
        const float4 t2 = (Op::kNeedT && two) ? ldg_stream4(op.ts + i2) : z;
        op.group(i, 4);
        event(h, op, i + 0, x.x, y.x, t.x, p.x);
        event(h, op, i + 1, x.y, y.y, t.y, p.y);
        event(h, op, i + 2, x.z, y.z, t.z, p.z);
        event(h, op, i + 3, x.w, y.w, t.w, p.w);
        if (two) {
            op.group(i2, 4);
            event(h, op, i2 + 0, x2.x, y2.x, t2.x, p2.x);
            event(h, op, i2 + 1, x2.y, y2.y, t2.y, p2.y);
            event(h, op, i2 + 2, x2.z, y2.z, t2.z, p2.z);
            event(h, op, i2 + 3, x2.w, y2.w, t2.w, p2.w);
        }
    }
    for (long i = (n4 << 2) + (long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        op.group(i, 1);
        event(h, op, i, op.xs[i], op.ys[i], Op::kNeedT ? op.ts[i] : 0.f, op.ps[i]);
    }

    if (MODE == kGlobal || MODE == kGlobal64) return;
    __syncthreads();
    if (MODE == kSmemFix) {
        const unsigned* su = reinterpret_cast<const unsigned*>(h.s_i);
        for (int k = threadIdx.x; k < nbins; k += kThreads) {
            const long long v = (long long)su[k] - (long long)su[nbins + k];
            if (v) atomicAdd(&h.g64[k], (unsigned long long)v);
        }
    } else if (MODE == kSmem64) {
        for (int k = threadIdx.x; k < nbins; k += kThreads) {
            const unsigned long long v = h.s64[k];
            if (v) atomicAdd(&h.g64[k], v);
        }
    } else if (MODE == kSmem32) {
        for (int k = threadIdx.x; k < nbins; k += kThreads) {
            if (Op::kFloat) { const float v = h.s_f[k]; if (v != 0.f) atomicAdd(&g_ext[k], v); }
            else { const int v = h.s_i[k]; if (v != 0) atomicAdd(&g_cnt[k], v); }
        }
    } else {
        for (int k = threadIdx.x; k < words; k += kThreads) {
            const unsigned v = (unsigned)h.s_i[k];
            if (v & 0xFFFFu) atomicAdd(&g_cnt[2 * k], (int)(v & 0xFFFFu));
            if (v >> 16) atomicAdd(&g_cnt[2 * k + 1], (int)(v >> 16));
        }
    }
}

// Time-interpolated voxels on grids too large for shared memory, pair form (events_to_voxel /
// events_to_voxel_torch bilinear).  Global fp32 reductions run in the L2 atomic units (~140 G
// reductions/s on spread addresses, measured); `red.global.add.v2.f32` updates BOTH adjacent
// bins of an event with one 8-byte reduction when the grid is laid out as float2 slots
// [bins-1][H*W] (.x -> bin j, .y -> bin j+1): one reduction per event instead of two
// (180x320x5: 74 -> 140 Gevents/s).  Each CTA streams a CONTIGUOUS range of events: with
// time-sorted input the CTAs then work on different bins at any moment, which spreads the
// reductions over the whole grid.  finalize_pairs_kernel folds the slots into [bins][H][W].
// (Grids that fit shared memory stay on scatter_kernel: the shared and the L2 atomic paths share
// the SM's load/store issue, so splitting events between them only lowered the rate -- DESIGN.md.)
constexpr int kPairThreads = 512;

__device__ __forceinline__ void red_global_f32x2(float2* addr, float lo, float hi) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(lo), "f"(hi) : "memory");
}

__global__ void __launch_bounds__(kPairThreads) voxel_pair_kernel(VoxelOp op, long n, float2* g_pairs,
                                                                   int vec_ok) {
    op.prepare(n);
    auto add = [&](long i, float x, float y, float t, float p) {
        int slot; float lo, hi;
        if (op.pair(i, x, y, t, p, slot, lo, hi)) red_global_f32x2(&g_pairs[slot], lo, hi);
    };
    const long n4 = vec_ok ? (n >> 2) : 0;
    const long per_cta = (n4 + gridDim.x - 1) / gridDim.x;
    const long g_end = min(n4, (long)(blockIdx.x + 1) * per_cta);
    for (long g = (long)blockIdx.x * per_cta + threadIdx.x; g < g_end; g += kPairThreads) {
        const long i = g << 2;
        const float4 x = ldg_stream4(op.xs + i);
        const float4 y = ldg_stream4(op.ys + i);
        const float4 p = ldg_stream4(op.ps + i);
        const float4 t = ldg_stream4(op.ts + i);
        add(i + 0, x.x, y.x, t.x, p.x);
        add(i + 1, x.y, y.y, t.y, p.y);
        add(i + 2, x.z, y.z, t.z, p.z);
        add(i + 3, x.w, y.w, t.w, p.w);
    }
    const long stride = (long)gridDim.x * kPairThreads;
    for (long i = (n4 << 2) + (long)blockIdx.x * kPairThreads + threadIdx.x; i < n; i += stride)
        add(i, op.xs[i], op.ys[i], op.ts[i], op.ps[i]);
}

// out[b][pix] = pairs[b][pix].x + pairs[b-1][pix].y
__global__ void finalize_pairs_kernel(const float2* __restrict__ pairs, float* __restrict__ out,
                                      int bins, int plane) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)bins * plane) return;
    const int b = (int)(i / plane);
    float v = 0.f;
    if (b < bins - 1) v = pairs[i].x;
    if (b > 0) v = __fadd_rn(v, pairs[i - plane].y);
    out[i] = v;
}

// Stack encoders on grids whose [2][bins][H][W] planes do not fit shared memory but ONE time bin's do (180x320:
// 2 x 57,600 16-bit counters, or 57,600 signed int32, = 230 KB).  A time bin is a contiguous event range
// [beg[b], end[b]), so every CTA is given one bin and a share of that range: its shared memory holds just that
// bin's planes and the per-event path is the count kernel's (one shared atomic), instead of one L2 atomic per event
// (145 -> ~450 Gevents/s).  Events on a boundary belong to two bins and are visited by CTAs of both, which is the
// reference's double count (F10).
constexpr int kBinThreads = 1024;
template <bool POL>
__global__ void __launch_bounds__(kBinThreads) stack_bins_kernel(StackOp op, long n, int* __restrict__ g_cnt,
                                                                 float* __restrict__ g_ext, int vec_ok) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* s = reinterpret_cast<unsigned*>(smem_raw);      // POL: [2*plane] 16-bit counters; else [plane] int32
    __shared__ long s_beg[64], s_end[64];
    const int plane = op.H * op.W;
    const int G = gridDim.x, c = blockIdx.x;
    const int b = (int)((long)c * op.bins / G);
    const int c0 = (int)(((long)b * G + op.bins - 1) / op.bins);            // first CTA of bin b
    const int c1 = (int)(((long)(b + 1) * G + op.bins - 1) / op.bins);      // first CTA of bin b + 1
    for (int k = threadIdx.x; k < plane; k += kBinThreads) s[k] = 0u;
    if (threadIdx.x < op.bins) { s_beg[threadIdx.x] = op.beg[threadIdx.x]; s_end[threadIdx.x] = op.end[threadIdx.x]; }
    __syncthreads();
    const long lo = max(0L, s_beg[b]), hi = min(n, s_end[b]);
    const long gbase = (long)b * plane;                        // bin b of polarity plane 0 in the output
    const long gneg = (long)op.bins * plane;                   // offset of the negative planes (POL)

    auto event = [&](long i, float x, float y, float p) {
        if (i < lo || i >= hi) return;
        Pix q = decode_xy(x, y, op.H, op.W, false);
        if (!q.oor) {
            const int pix = q.y * op.W + q.x;
            if (POL) {
                const float w = p * p;
                const int neg = p < 0.f;
                if (w == 1.f) {
                    // two 16-bit counters per word; the thread that moves a field from 0x7FFF to 0x8000 takes
                    // 0x8000 back out and credits the global grid (see Hist::add_int)
                    const int L = neg * plane + pix, sh = (L & 1) * 16;
                    const unsigned old = atomicAdd(&s[L >> 1], 1u << sh);
                    if (((old >> sh) & 0xFFFFu) == 0x7FFFu) {
                        atomicSub(&s[L >> 1], 0x8000u << sh);
                        atomicAdd(&g_cnt[gbase + (neg ? gneg : 0) + pix], 32768);
                    }
                } else if (w != 0.f) atomicAdd(&g_ext[gbase + (neg ? gneg : 0) + pix], w);
            } else {
                if (p == 1.f) atomicAdd(reinterpret_cast<int*>(&s[pix]), 1);
                else if (p == -1.f) atomicAdd(reinterpret_cast<int*>(&s[pix]), -1);
                else if (p != 0.f) atomicAdd(&g_ext[gbase + pix], p);
            }
            return;
        }
        // out of range (rare): the first bin that holds the event zeroes it in place, later bins see (0,0)
        bool first = true, multi = false;
        for (int e = 0; e < op.bins; ++e)
            if (e != b && s_beg[e] <= i && i < s_end[e]) { multi = true; first = first && e > b; }
        if (POL && op.quirks() && (!first || p < 0.f)) {
            const float w = p * p;
            const long gi = gbase + (p < 0.f ? gneg : 0);
            if (w == 1.f) atomicAdd(&g_cnt[gi], 1); else if (w != 0.f) atomicAdd(&g_ext[gi], w);
        }
        // An event of ONE bin is used by this thread alone and is zeroed here.  An event of several bins is also
        // read by other CTAs: nobody touches it while this kernel runs (stack_bins_mutate_kernel does afterwards),
        // so every CTA sees the original coordinates whatever the order they run in.
        if (op.mutate() && !multi) { op.xs[i] = 0.f; op.ys[i] = 0.f; if (!POL) op.ps[i] = 0.f; }
    };

    if (hi > lo) {
        const long n4 = vec_ok ? (n >> 2) : 0;
        const long g_lo = lo >> 2, g_hi = min(n4, (hi + 3) >> 2);       // 4-event groups that meet [lo, hi)
        const long stride = (long)(c1 - c0) * kBinThreads;
        for (long g = g_lo + (long)(c - c0) * kBinThreads + threadIdx.x; g < g_hi; g += 2 * stride) {
            const long i = g << 2, i2 = (g + stride) << 2;
            const bool two = g + stride < g_hi;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 x = ldg_stream4(op.xs + i), y = ldg_stream4(op.ys + i), p = ldg_stream4(op.ps + i);
            const float4 x2 = two ? ldg_stream4(op.xs + i2) : z, y2 = two ? ldg_stream4(op.ys + i2) : z;
            const float4 p2 = two ? ldg_stream4(op.ps + i2) : z;
            event(i + 0, x.x, y.x, p.x); event(i + 1, x.y, y.y, p.y);
            event(i + 2, x.z, y.z, p.z); event(i + 3, x.w, y.w, p.w);
            if (two) {
                event(i2 + 0, x2.x, y2.x, p2.x); event(i2 + 1, x2.y, y2.y, p2.y);
                event(i2 + 2, x2.z, y2.z, p2.z); event(i2 + 3, x2.w, y2.w, p2.w);
            }
        }
        // events past the last whole group (or all of them when the arrays are not 16-byte aligned)
        for (long i = max(lo, n4 << 2) + (long)(c - c0) * kBinThreads + threadIdx.x; i < hi; i += stride)
            event(i, op.xs[i], op.ys[i], op.ps[i]);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < plane; k += kBinThreads) {
        const unsigned v = s[k];
        if (POL) {
            const int L0 = 2 * k, L1 = 2 * k + 1;
            if (v & 0xFFFFu) atomicAdd(&g_cnt[gbase + (L0 >= plane ? gneg + L0 - plane : L0)], (int)(v & 0xFFFFu));
            if (v >> 16) atomicAdd(&g_cnt[gbase + (L1 >= plane ? gneg + L1 - plane : L1)], (int)(v >> 16));
        } else if (v) {
            atomicAdd(&g_cnt[gbase + k], (int)v);
        }
    }
}

// The in-place zeroing (encodings.py:37-39) of out-of-range events that lie in MORE than one time bin, left alone by
// stack_bins_kernel: CTA b walks the intersections of bin b with every later bin.
template <bool POL>
__global__ void stack_bins_mutate_kernel(StackOp op, long n) {
    const int b = blockIdx.x;
    for (int e = b + 1; e < op.bins; ++e) {
        const long lo = max(0L, max(op.beg[b], op.beg[e])), hi = min(n, min(op.end[b], op.end[e]));
        for (long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            if (decode_xy(op.xs[i], op.ys[i], op.H, op.W, false).oor) {
                op.xs[i] = 0.f; op.ys[i] = 0.f;
                if (!POL) op.ps[i] = 0.f;
            }
        }
    }
}

// ---------------------------------------------------------------- grids beyond one SM: bins split over CTAs ("roles")
// A [2,360,640] count grid (921 KB of 16-bit counters) or the per-interval state of a 180x320 voxel grid (460 KB) does
// not fit the 227 KB of one SM, and one L2 reduction per event caps those encoders at ~150 Gevents/s.  Here `roles`
// CTAs stream the SAME contiguous event range, each holding a different 230 KB share of the bins in shared memory:
//   voxels (2 roles)    per pixel and time slot j the state is U = sum of 2^-24 fixed-point weights (frac for p = +1,
//                       1 - frac for p = -1) and the two 16-bit event counts; bin j = Npos * 2^24 - U, bin j+1 = U -
//                       Nneg * 2^24 (the same integers as the kSmemFix path, order-independent).  Role 0 keeps the U
//                       plane, role 1 the count plane: ONE native shared-memory atomic per event and CTA, no divergence.
//                       Time-sorted input keeps a CTA inside one slot for a long stretch; the plane is flushed to the
//                       global 64-bit grid when the slot of an iteration's first event changes, and events of another
//                       slot, out-of-range events and non-unit polarities take the global-atomic path in role 0.
// Every event is read once from HBM and again out of L2 by the other role, which runs in step because it does the
// same work on the same range.
// (Counts on 360x640 -- 4 roles owning a quarter of the pixels each -- were built the same way and measured at 81-108
// Gevents/s against the 157 of one L2 reduction per event: with spatially random events a quarter of a warp's lanes own
// an event, so every CTA pays the whole path for every event group.  That grid stays on the global-atomic path.)  The reference's in-place zeroing of out-of-range events cannot happen while other CTAs
// still read those events: role 0 appends their indices to a list and oor_zero_kernel applies it afterwards.
// (An exchange of per-event records between the CTAs of a cluster through distributed shared memory, so that every
// event is decoded once, was built first and measured at 55 Gevents/s: 10.9 warp instructions per event against
// 2.3 here -- slot allocation, per-lane remote stores and the per-source drain cost more than the second decode.)
constexpr int kRoleThreads = 1024;
constexpr int kRoleWords = 57600;                 // 230,400 B of bins per CTA
constexpr long kOorCap = 65536;                   // out-of-range event indices kept for the deferred zeroing

struct OorList {
    unsigned long long* count;                    // appended so far (may exceed kOorCap: then the fix-up scans everything)
    long* idx;
    __device__ __forceinline__ void add(long i) const {
        const unsigned long long k = atomicAdd(count, 1ull);
        if (k < (unsigned long long)kOorCap) idx[k] = i;
    }
};

struct VoxelRole {
    static constexpr bool kNeedT = true;
    VoxelOp op;                                   // flags WITHOUT BMC_ENC_MUTATE (zeroing is deferred)
    unsigned long long* g64;                      // [bins][plane], 2^-24 units
    OorList oor;
    int mutate;
    __device__ __forceinline__ int words() const { return op.H * op.W; }
    __device__ __forceinline__ float tnorm(float t) const {
        const float fb = (float)(op.bins - 1);
        if (op.flags & BMC_ENC_TNORM) return __fmul_rn(__fdiv_rn(__fsub_rn(t, op.t0), op.dt), fb);
        return __fmul_rn(t, fb);
    }
    __device__ __forceinline__ int slot_from(float t) const {
        const float tn = tnorm(t), fl = floorf(tn);
        return (tn >= 0.f && fl <= (float)(op.bins - 2)) ? (int)fl : -1;
    }
    __device__ __forceinline__ int slot_of(long i) const { return slot_from(op.ts[i]); }
    // ROLE and TN (BMC_ENC_TNORM) are compile-time: the common event is ~40 instructions, one shared-memory atomic
    template <int ROLE, bool TN>
    __device__ __forceinline__ void event(unsigned* bins, int cur, float fcur, float fb, long i, float x, float y, float t, float p) const {
        const float tn = TN ? __fmul_rn(__fdiv_rn(__fsub_rn(t, op.t0), op.dt), fb) : __fmul_rn(t, fb);
        const float fl = floorf(tn);
        const bool in = (x < (float)op.W) & (x >= 0.f) & (y < (float)op.H) & (y >= 0.f);
        if (in & (fabsf(p) == 1.f) & (fl == fcur)) {                // (fl == fcur >= 0 implies tn >= 0)
            const int yy = (int)y, pix = (op.flip() ? op.H - 1 - yy : yy) * op.W + (int)x;
            const bool neg = p < 0.f;
            if (ROLE == 0) {
                const unsigned dhi = __float2uint_rn(__fsub_rn(tn, fl) * (float)(1 << kFixBits));
                const unsigned d = neg ? (1u << kFixBits) - dhi : dhi;
                const unsigned old = atomicAdd(&bins[pix], d);
                if (old + d < old) {                 // U wrapped: 2^32 units move from bin j to bin j + 1
                    const long plane = (long)op.H * op.W, gi = (long)cur * plane + pix;
                    atomicAdd(&g64[gi], (unsigned long long)(-(1ll << 32)));
                    atomicAdd(&g64[gi + plane], 1ull << 32);
                }
            } else {
                const int sh = neg ? 16 : 0;
                const unsigned old = atomicAdd(&bins[pix], 1u << sh);
                if (((old >> sh) & 0xFFFFu) == 0x7FFFu) {      // 16-bit count at 0x8000: take 32768 events out (Hist::add_int)
                    const long plane = (long)op.H * op.W, gi = (long)cur * plane + pix;
                    atomicSub(&bins[pix], 0x8000u << sh);
                    if (neg) atomicAdd(&g64[gi + plane], (unsigned long long)(-(1ll << (15 + kFixBits))));
                    else atomicAdd(&g64[gi], 1ull << (15 + kFixBits));
                }
            }
            return;
        }
        if (ROLE != 0) return;
        if (!in && mutate) oor.add(i);
        int slot; float lo, hi;
        if (op.pair(i, x, y, t, p, slot, lo, hi)) {
            if (lo != 0.f) atomicAdd(&g64[slot], (unsigned long long)__double2ll_rn((double)lo * (double)(1 << kFixBits)));
            if (hi != 0.f) atomicAdd(&g64[slot + op.H * op.W], (unsigned long long)__double2ll_rn((double)hi * (double)(1 << kFixBits)));
        }
    }
    __device__ __forceinline__ void flush(unsigned* bins, int role, int cur) const {
        const int plane = op.H * op.W;
        unsigned long long* g0 = g64 + (long)cur * plane;
        for (int k = threadIdx.x; k < plane; k += kRoleThreads) {
            const unsigned v = bins[k];
            if (!v) continue;
            bins[k] = 0u;
            if (role == 0) {
                atomicAdd(&g0[k], (unsigned long long)(-(long long)v));
                atomicAdd(&g0[plane + k], (unsigned long long)v);
            } else {
                if (v & 0xFFFFu) atomicAdd(&g0[k], (unsigned long long)(v & 0xFFFFu) << kFixBits);
                if (v >> 16) atomicAdd(&g0[plane + k], (unsigned long long)(-((long long)(v >> 16) << kFixBits)));
            }
        }
    }
};

// events_to_channels on grids up to 360x640: role 0 counts the positive events, role 1 the negative ones, each in a plane
// of 8-bit counters (230,400 B).  A field is advanced by a compare-and-swap: the transition 0xFF -> 0x00 is taken
// explicitly and credits 256 to the global grid, so the count is exact for any input (a native add would carry into the
// neighbouring pixel's field, and a threshold scheme has no head-room in 8 bits).  Half of a warp's lanes take the
// counting path in each role -- against a quarter with four spatial roles, which was measured slower than the L2 path.
struct CountsRole {
    static constexpr bool kNeedT = false;
    ChannelsOp op;                                // flags WITHOUT BMC_ENC_MUTATE
    int* g_cnt; float* g_ext;
    OorList oor;
    int mutate;
    __device__ __forceinline__ int words() const { return (op.H * op.W + 3) >> 2; }
    __device__ __forceinline__ int slot_from(float) const { return 0; }
    __device__ __forceinline__ int slot_of(long) const { return 0; }
    template <int ROLE, bool TN>
    __device__ __forceinline__ void event(unsigned* bins, int, float, float, long i, float x, float y, float t, float p) const {
        const bool in = (x < (float)op.W) & (x >= 0.f) & (y < (float)op.H) & (y >= 0.f);
        if (in & (p * p == 1.f)) {
            if ((p < 0.f) != (ROLE == 1)) return;
            const int pix = (op.H - 1 - (int)y) * op.W + (int)x, sh = (pix & 3) * 8;
            const uint32_t wa = smem_u32(bins) + (uint32_t)(pix >> 2) * 4u;
            unsigned old, assumed;
            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(old) : "r"(wa) : "memory");
            bool wrap;
            do {
                assumed = old;
                wrap = ((assumed >> sh) & 0xFFu) == 0xFFu;
                const unsigned nw = wrap ? assumed & ~(0xFFu << sh) : assumed + (1u << sh);
                asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(wa), "r"(assumed), "r"(nw) : "memory");
            } while (old != assumed);
            if (wrap) atomicAdd(&g_cnt[(ROLE ? op.H * op.W : 0) + pix], 256);
            return;
        }
        if (ROLE != 0) return;                       // out-of-range events (F9 leak, deferred zeroing) and non-unit weights: once
        if (!in && mutate) oor.add(i);
        Hist<kGlobal, false> h;
        h.g_cnt = g_cnt; h.g_ext = g_ext;
        op.run(h, i, x, y, t, p);
    }
    __device__ __forceinline__ void flush(unsigned* bins, int role, int) const {
        const int plane = op.H * op.W;
        int* g = g_cnt + (role ? plane : 0);
        for (int k = threadIdx.x; k < words(); k += kRoleThreads) {
            const unsigned v = bins[k];
            if (!v) continue;
            bins[k] = 0u;
#pragma unroll
            for (int f = 0; f < 4; ++f)
                if (((v >> (8 * f)) & 0xFFu) && 4 * k + f < plane) atomicAdd(&g[4 * k + f], (int)((v >> (8 * f)) & 0xFFu));
        }
    }
};

template <class R, int ROLE, bool TN>
__device__ __forceinline__ void role_body(R& pol, unsigned* bins, long n, int grp, int G, int vec_ok) {
    const int tid = threadIdx.x;
    const float fb = (float)(pol.op.bins - 1);
    int cur = -2;                                   // no slot yet
    float fcur = __int_as_float(0x7fc00000);
    // the events of this group: whole 4-event blocks (vector loads) or single events when unaligned
    const long units = vec_ok ? (n >> 2) : n;
    const long per = (units + G - 1) / G;
    const long u_lo = (long)grp * per, u_hi = min(units, u_lo + per);
    const int w = vec_ok ? 4 : 1;
    // Two units per thread and iteration, all loads issued before the first atomic; the time stamp that decides the NEXT
    // iteration's slot is requested now and looked at after this iteration's events (no dependent load on the path).
    float t_first = (R::kNeedT && u_lo < u_hi) ? pol.op.ts[u_lo * w] : 0.f;
    for (long base = u_lo; base < u_hi; base += 2 * kRoleThreads) {
        const int slot = pol.slot_from(t_first);     // uniform: the iteration's first event
        const long nb = base + 2 * kRoleThreads;
        if (R::kNeedT && nb < u_hi) t_first = ldg_stream1(pol.op.ts + nb * w);
        if (slot != cur) {
            __syncthreads();
            if (cur >= 0) pol.flush(bins, ROLE, cur);
            cur = slot; fcur = slot >= 0 ? (float)slot : __int_as_float(0x7fc00000);      // NaN: no event matches
            __syncthreads();
        }
        const long u = base + tid, u2 = u + kRoleThreads;
        if (vec_ok) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool one = u < u_hi, two = u2 < u_hi;
            const long i = u << 2, i2 = u2 << 2;
            const float4 x = one ? ldg_stream4(pol.op.xs + i) : z, y = one ? ldg_stream4(pol.op.ys + i) : z;
            const float4 p = one ? ldg_stream4(pol.op.ps + i) : z, t = (R::kNeedT && one) ? ldg_stream4(pol.op.ts + i) : z;
            const float4 x2 = two ? ldg_stream4(pol.op.xs + i2) : z, y2 = two ? ldg_stream4(pol.op.ys + i2) : z;
            const float4 p2 = two ? ldg_stream4(pol.op.ps + i2) : z, t2 = (R::kNeedT && two) ? ldg_stream4(pol.op.ts + i2) : z;
            if (one) {
                pol.template event<ROLE, TN>(bins, cur, fcur, fb, i + 0, x.x, y.x, t.x, p.x);
                pol.template event<ROLE, TN>(bins, cur, fcur, fb, i + 1, x.y, y.y, t.y, p.y);
                pol.template event<ROLE, TN>(bins, cur, fcur, fb, i + 2, x.z, y.z, t.z, p.z);
                pol.template event<ROLE, TN>(bins, cur, fcur, fb, i + 3, x.w, y.w, t.w, p.w);
            }
            if (two) {
                pol.template event<ROLE, TN>(bins, cur, fcur, fb, i2 + 0, x2.x, y2.x, t2.x, p2.x);
                pol.template event<ROLE, TN>(bins, cur, fcur, fb, i2 + 1, x2.y, y2.y, t2.y, p2.y);
                pol.template event<ROLE, TN>(bins, cur, fcur, fb, i2 + 2, x2.z, y2.z, t2.z, p2.z);
                pol.template event<ROLE, TN>(bins, cur, fcur, fb, i2 + 3, x2.w, y2.w, t2.w, p2.w);
            }
        } else {
            if (u < u_hi) pol.template event<ROLE, TN>(bins, cur, fcur, fb, u, pol.op.xs[u], pol.op.ys[u], R::kNeedT ? pol.op.ts[u] : 0.f, pol.op.ps[u]);
            if (u2 < u_hi) pol.template event<ROLE, TN>(bins, cur, fcur, fb, u2, pol.op.xs[u2], pol.op.ys[u2], R::kNeedT ? pol.op.ts[u2] : 0.f, pol.op.ps[u2]);
        }
    }
    if (vec_ok && grp == G - 1) {                    // the last n % 4 events
        const long i = (n & ~3L) + tid;
        if (i < n) {
            // (they may belong to another slot than the plane holds: event() then takes the global path in role 0)
            const bool same = pol.slot_of(i) == cur;
            pol.template event<ROLE, TN>(bins, cur, same ? fcur : __int_as_float(0x7fc00000), fb, i, pol.op.xs[i], pol.op.ys[i], R::kNeedT ? pol.op.ts[i] : 0.f, pol.op.ps[i]);
        }
    }
    __syncthreads();
    if (cur >= 0) pol.flush(bins, ROLE, cur);
}

template <class R>
__global__ void __launch_bounds__(kRoleThreads, 1) role_kernel(R pol, long n, int roles, int vec_ok) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* bins = reinterpret_cast<unsigned*>(smem_raw);
    const int role = blockIdx.x % roles, grp = blockIdx.x / roles, G = gridDim.x / roles;
    for (int k = threadIdx.x; k < pol.words(); k += kRoleThreads) bins[k] = 0u;
    pol.op.prepare(n);
    __syncthreads();
    const bool tn = pol.op.flags & BMC_ENC_TNORM;
    if (role == 0) { if (tn) role_body<R, 0, true>(pol, bins, n, grp, G, vec_ok); else role_body<R, 0, false>(pol, bins, n, grp, G, vec_ok); }
    else { if (tn) role_body<R, 1, true>(pol, bins, n, grp, G, vec_ok); else role_body<R, 1, false>(pol, bins, n, grp, G, vec_ok); }
}

// deferred in-place zeroing of out-of-range events (encodings.py:252-254) after role_kernel
__global__ void oor_zero_kernel(float* xs, float* ys, long n, int H, int W, const unsigned long long* count, const long* idx) {
    const unsigned long long c = *count;
    const long stride = (long)gridDim.x * blockDim.x, t0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= (unsigned long long)kOorCap) {
        for (long k = t0; k < (long)c; k += stride) { const long i = idx[k]; xs[i] = 0.f; ys[i] = 0.f; }
    } else {                                        // more than the list holds: look at every event
        for (long i = t0; i < n; i += stride)
            if (decode_xy(xs[i], ys[i], H, W, false).oor) { xs[i] = 0.f; ys[i] = 0.f; }
    }
}

// out = fp32(saturated count) + fp32 extras; the reference's serial `+= 1.0f` sticks at 2^24.
__global__ void finalize_kernel(const int* __restrict__ cnt, const float* __restrict__ ext,
                                float* __restrict__ out, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cnt[i];
    c = max(-(1 << 24), min(c, 1 << 24));
    out[i] = (float)c + ext[i];
}

// deterministic path: out = fp32(fixed-point sum * 2^-32), one rounding
__global__ void finalize64_kernel(const long long* __restrict__ g64, float* __restrict__ out, long n, double unit) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = (float)((double)g64[i] * unit);
}

// Bin boundaries of the stack encoders, evaluated exactly like encodings.py:172-178 + :75-97
// (float32, two roundings for ts[0] + delta_t*bi, any-equal binary search: every iteration tests ts[l], ts[r] and
// ts[mid] for equality, in that order, before it halves the range).
// One WARP per (bin, side).  The probe positions of the next five iterations form a binary tree that depends only
// on (l, r), not on the data: lane k < 31 replays the path to heap node k with integer arithmetic, loads the three
// stamps its node would test, and the warp then walks the tree through shuffles -- the reference's probe sequence
// and comparisons exactly (also on unsorted input), in ~log2(n)/5 round trips to DRAM instead of log2(n).
// `shift` re-bases the boundaries for a rank that holds events [shift, shift + n_local) of the recording.
// `zero_ends` (device int, always written): ts[0] == 0 && ts[n-1] == 0, the cheap necessary condition of the
// reference's `ts.sum() == 0` early-out; with `skip_zero_ends` such a call gets empty bins (all-zero output, no
// event touched) so that the host can decide the early-out AFTER the launch instead of synchronising before it.
constexpr int kBoundsThreads = 128;
__global__ void __launch_bounds__(kBoundsThreads) bin_bounds_kernel(
    const float* __restrict__ ts, long n, int bins, long* beg, long* end, long shift, int skip_zero_ends,
    int* zero_ends) {
    const int k = (blockIdx.x * kBoundsThreads + threadIdx.x) >> 5;      // warp = (bin, side)
    const int lane = threadIdx.x & 31;
    if (k >= 2 * bins) return;
    const int bi = k >> 1;
    const bool right = k & 1;
    const float t0 = ts[0];
    const bool ze = t0 == 0.f && ts[n - 1] == 0.f;
    if (k == 0 && lane == 0) *zero_ends = ze;
    if (ze && skip_zero_ends) {
        if (lane == 0) { if (right) end[bi] = 0; else beg[bi] = 0; }
        return;
    }
    const float dt = __fadd_rn(__fsub_rn(ts[n - 1], t0), 1e-6f);
    const float delta = __fdiv_rn(dt, (float)bins);
    const float tstart = __fadd_rn(t0, __fmul_rn(delta, (float)bi));
    const float x = right ? __fadd_rn(tstart, delta) : tstart;

    long l = 0, r = n - 1, res = 0;
    bool found = false, done = false;
    while (!done) {
        // lane -> heap node (lane 31 idles): replay the path from the round's root; child 2k+1 = "ts[mid] < x"
        long nl = l, nr = r;
        bool valid = lane < 31 && nl <= nr;
        const int idx = lane + 1;
        const int depth = 31 - __clz(idx);
        for (int d = depth - 1; d >= 0 && valid; --d) {
            const long mid = nl + (nr - nl) / 2;
            if ((idx >> d) & 1) nr = mid - 1; else nl = mid + 1;       // heap: left child 2k+1 has bit 0
            valid = nl <= nr;
        }
        const long nmid = nl + (nr - nl) / 2;
        float tl = 0.f, tr = 0.f, tm = 0.f;
        if (valid) { tl = ts[nl]; tr = ts[nr]; tm = ts[nmid]; }
        // walk five levels
        int cur = 0;
#pragma unroll 1
        for (int level = 0; level < 5; ++level) {
            const bool cv = __shfl_sync(0xffffffffu, (int)valid, cur);
            const long cl = __shfl_sync(0xffffffffu, nl, cur), cr = __shfl_sync(0xffffffffu, nr, cur);
            const long cm = __shfl_sync(0xffffffffu, nmid, cur);
            const float vl = __shfl_sync(0xffffffffu, tl, cur), vr = __shfl_sync(0xffffffffu, tr, cur);
            const float vm = __shfl_sync(0xffffffffu, tm, cur);
            l = cl; r = cr;
            if (!cv) { done = true; break; }                           // l > r: the search ended without a hit
            if (vl == x) { res = cl; found = done = true; break; }
            if (vr == x) { res = cr; found = done = true; break; }
            if (vm == x) { res = cm; found = done = true; break; }
            if (vm < x) { l = cm + 1; cur = 2 * cur + 1; } else { r = cm - 1; cur = 2 * cur + 2; }
        }
        if (!done && l > r) done = true;
    }
    if (!found) res = right ? r : l;
    if (lane == 0) { if (right) end[bi] = res + 1 - shift; else beg[bi] = res - shift; }
}

// One CTA per window (dataloader pattern: ~2048 events -> one [2,H,W] grid).  fp32 smem bins:
// sums of +1.0f below 2^24 are exact integers in any order, so this is bit-exact too.
__global__ void __launch_bounds__(256) channels_windows_kernel(
    float* xs, float* ys, const float* ps, const long* __restrict__ offsets, int H, int W,
    float* __restrict__ out, unsigned flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s = reinterpret_cast<float*>(smem_raw);
    const int nb = 2 * H * W;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s[k] = 0.f;
    __syncthreads();
    const long b = offsets[blockIdx.x], e = offsets[blockIdx.x + 1];
    const bool quirks = !(flags & BMC_ENC_NO_QUIRKS), mut = flags & BMC_ENC_MUTATE;
    for (long i = b + threadIdx.x; i < e; i += blockDim.x) {
        const float x = xs[i], y = ys[i], p = ps[i];
        Pix q = decode_xy(x, y, H, W, true);
        const float w = p * p;
        if (!q.oor) {
            if (w != 0.f) atomicAdd(&s[(p < 0.f ? H * W : 0) + q.y * W + q.x], w);
        } else {
            if (quirks && p < 0.f) atomicAdd(&s[H * W + (H - 1) * W], w);
            if (mut) { xs[i] = 0.f; ys[i] = 0.f; }
        }
    }
    __syncthreads();
    float* o = out + (long)blockIdx.x * nb;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) o[k] = s[k];
}

// events_to_channels on ONE small window (the reference's call pattern: h5dataset.py:518-526 encodes every 1024-2048 event
// window with its own call): zeroing, counting and the fp32 output in a single one-CTA launch instead of memset +
// scatter + finalize.  Same arithmetic as channels_windows_kernel (fp32 shared-memory bins: sums of +1.0f below 2^24 are
// exact integers in any order).
constexpr long kOneCtaMaxEvents = 32768;
__global__ void __launch_bounds__(1024) channels_one_kernel(float* xs, float* ys, const float* ps, long n, int H, int W,
                                                            float* __restrict__ out, unsigned flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s = reinterpret_cast<float*>(smem_raw);
    const int nb = 2 * H * W;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s[k] = 0.f;
    __syncthreads();
    const bool quirks = !(flags & BMC_ENC_NO_QUIRKS), mut = flags & BMC_ENC_MUTATE;
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
        const float x = xs[i], y = ys[i], p = ps[i];
        Pix q = decode_xy(x, y, H, W, true);
        const float w = p * p;
        if (!q.oor) {
            if (w != 0.f) atomicAdd(&s[(p < 0.f ? H * W : 0) + q.y * W + q.x], w);
        } else {
            if (quirks && p < 0.f) atomicAdd(&s[H * W + (H - 1) * W], w);
            if (mut) { xs[i] = 0.f; ys[i] = 0.f; }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nb; k += blockDim.x) out[k] = s[k];
}

// Window pipeline of the dataloader on raw recordings (h5dataset.py:197-210 compute_k_indices, :407-414
// get_events, base_dataset.py:24-31 event_formatting, h5dataset.py:518-526 create_cnt_encoding): window i is
// events [stride*i, min(stride*i + window, n_events - 1)) of the int16 / float64 arrays as stored in the
// HDF5 files (event_packagers.py:128-156), cast to float32 and counted like events_to_channels.  One CTA
// per window; 12 B/event (2 + 2 + 8), no intermediate float arrays.
__global__ void __launch_bounds__(256) channels_windows_raw_kernel(
    const short* __restrict__ xs, const short* __restrict__ ys, const double* __restrict__ ps, long n_events,
    long window, long stride, int H, int W, float* __restrict__ out, unsigned flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s = reinterpret_cast<float*>(smem_raw);
    const int nb = 2 * H * W;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s[k] = 0.f;
    __syncthreads();
    const long b = stride * blockIdx.x;
    long e = b + window;
    if (e > n_events - 1) e = n_events - 1;
    const bool quirks = !(flags & BMC_ENC_NO_QUIRKS);
    for (long i = b + threadIdx.x; i < e; i += blockDim.x) {
        const float x = (float)xs[i], y = (float)ys[i], p = (float)ps[i];
        Pix q = decode_xy(x, y, H, W, true);
        const float w = p * p;
        if (!q.oor) {
            if (w != 0.f) atomicAdd(&s[(p < 0.f ? H * W : 0) + q.y * W + q.x], w);
        } else if (quirks && p < 0.f) {
            atomicAdd(&s[H * W + (H - 1) * W], w);
        }
    }
    __syncthreads();
    float* o = out + (long)blockIdx.x * nb;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) o[k] = s[k];
}

// BaseDataset.event_formatting (base_dataset.py:24-31): float32 casts and ts = (ts - ts[0]) / (ts[-1] - ts[0] + 1e-6)
// evaluated in float32 like the reference's tensor arithmetic -> out [4][n].
__global__ void format_events_kernel(const short* __restrict__ xs, const short* __restrict__ ys,
                                     const double* __restrict__ ts, const double* __restrict__ ps, long n,
                                     float* __restrict__ out) {
    const float t0 = (float)ts[0];
    const float den = __fadd_rn(__fsub_rn((float)ts[n - 1], t0), 1e-6f);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        out[i] = (float)xs[i];
        out[n + i] = (float)ys[i];
        out[2 * n + i] = __fdiv_rn(__fsub_rn((float)ts[i], t0), den);
        out[3 * n + i] = (float)ps[i];
    }
}

// ---------------------------------------------------------------- host-side launch logic
struct Ws {
    int* cnt; float* ext; long* beg; long* end;
    unsigned long long* oor_count; long* oor_idx;      // role_kernel: deferred zeroing of out-of-range events
};

size_t ws_bytes(long out_elems) { return (size_t)out_elems * 8 + 2 * 64 * sizeof(long) + 256 + 16 + (size_t)kOorCap * sizeof(long); }
// the stack encoders' `zero_ends` word sits right after the two boundary arrays, inside the 256 spare bytes
size_t ws_flag_offset(long out_elems) { return (((size_t)out_elems * 8 + 15) & ~(size_t)15) + 2 * 64 * sizeof(long); }

int carve(void* ws, size_t ws_bytes_given, long out_elems, Ws& w) {
    if (!ws || ws_bytes_given < ws_bytes(out_elems)) {
        set_error("encoder workspace too small: need %zu bytes, got %zu", ws_bytes(out_elems),
                  ws_bytes_given);
        return BMC_ERR_WORKSPACE;
    }
    if ((uintptr_t)ws & 15) { set_error("encoder workspace must be 16-byte aligned"); return BMC_ERR_ARG; }
    char* p = static_cast<char*>(ws);
    w.cnt = reinterpret_cast<int*>(p);
    w.ext = reinterpret_cast<float*>(p + (size_t)out_elems * 4);
    size_t off = ((size_t)out_elems * 8 + 15) & ~(size_t)15;
    w.beg = reinterpret_cast<long*>(p + off);
    w.end = w.beg + 64;
    w.oor_count = reinterpret_cast<unsigned long long*>(p + off + 2 * 64 * sizeof(long) + 256);
    w.oor_idx = reinterpret_cast<long*>(w.oor_count + 2);
    return BMC_OK;
}

template <class Op, int MODE, int THREADS>
int launch_threads(const Op& op, long n, int nbins, const Ws& w, int vec_ok, size_t smem, int per_sm,
                   cudaStream_t st) {
    auto kern = scatter_kernel<Op, MODE, THREADS>;
    if (smem > 48 * 1024)
        BMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // Persistent grid: a multiple of the SM count, but never so many CTAs that flushing their bins outweighs the
    // events each CTA streams.  A CTA retires ~2 shared atomics per clock and flushes only its NON-ZERO bins, at
    // the ~140 G/s of the L2 atomic units: g CTAs cost n/g / 3.8e9 + g * min(nbins, n/g) / 140e9 seconds.  While
    // n/g < nbins the flush term is n / 140e9 whatever g is, so small streams on large grids (n < 37 nbins) take
    // every SM; beyond that the minimum is at g = sqrt(36.8 n / nbins).
    long want = (n + (long)THREADS * 4 * 8 - 1) / ((long)THREADS * 4 * 8);
    if (MODE != kGlobal && MODE != kGlobal64) {
        long by_flush = n / (4L * nbins) + 1;
        const double ratio = 36.8 * (double)n / (double)nbins;
        const long model = ratio < 36.8 * 37.0 ? (long)sm_count() * per_sm : (long)sqrt(ratio);
        if (by_flush < model) by_flush = model;
        want = (n + (long)THREADS * 4 - 1) / ((long)THREADS * 4);       // at least one 4-event group per thread
        if (want > by_flush) want = by_flush;
    }
    long grid = (long)sm_count() * per_sm;
    if (want < grid) grid = want < 1 ? 1 : want;
    kern<<<(unsigned)grid, THREADS, smem, st>>>(op, n, nbins, w.cnt, w.ext, vec_ok);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

template <class Op, int MODE>
int launch_mode(const Op& op, long n, int nbins, const Ws& w, int vec_ok, cudaStream_t st) {
    size_t smem = MODE == kSmem32 ? (size_t)nbins * 4 : (MODE == kSmem16 ? (size_t)((nbins + 1) / 2) * 4 : ((MODE == kSmem64 || MODE == kSmemFix) ? (size_t)nbins * 8 : 0));
    int per_sm = 1;
    BMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scatter_kernel<Op, MODE, kThreads>, kThreads, smem));
    // One 1024-thread CTA per SM keeps enough loads in flight (1024 x 2 groups x 3-4 arrays x 16 B >= 96 KB) and
    // flushes its bins once per SM instead of once per 512-thread CTA: with 4 CTAs per SM the flush of a 7200-bin
    // grid was 4.3 M global atomics, 17 % of a 1e8-event launch.  Small streams keep the finer grid.
    static int fat = -1;
    if (fat < 0) fat = measure_env("BMC_ENC_FAT", 1);
    if (per_sm < 2 || (fat && MODE != kGlobal && MODE != kGlobal64 && n >= (long)sm_count() * 1024 * 64))
        return launch_threads<Op, MODE, 1024>(op, n, nbins, w, vec_ok, smem, 1, st);
    return launch_threads<Op, MODE, kThreads>(op, n, nbins, w, vec_ok, smem, per_sm > 4 ? 4 : per_sm, st);
}

template <class Op>
int run_scatter(const Op& op, long n, long out_elems, float* out, void* ws, size_t wsb,
                cudaStream_t st) {
    Ws w;
    int rc = carve(ws, wsb, out_elems, w);
    if (rc) return rc;
    BMC_CUDA(cudaMemsetAsync(w.cnt, 0, (size_t)out_elems * 8, st));
    if (n > 0) {
        const int vec_ok = (((uintptr_t)op.xs | (uintptr_t)op.ys | (uintptr_t)op.ps |
                             (uintptr_t)(Op::kNeedT ? op.ts : nullptr)) & 15) == 0;
        const int nbins = (int)out_elems;
        if constexpr (Op::kFloat) {
            if (op.flags & BMC_ENC_DETERMINISTIC) {
                rc = out_elems <= kMaxBinsSmem64 ? launch_mode<Op, kSmem64>(op, n, nbins, w, vec_ok, st)
                                                 : launch_mode<Op, kGlobal64>(op, n, nbins, w, vec_ok, st);
                if (rc) return rc;
                finalize64_kernel<<<(unsigned)((out_elems + 255) / 256), 256, 0, st>>>(reinterpret_cast<const long long*>(w.cnt), out, out_elems, 1.0 / 4294967296.0);
                BMC_CUDA(cudaGetLastError());
                return BMC_OK;
            }
            if (out_elems <= kMaxBinsSmemFix) {
                rc = launch_mode<Op, kSmemFix>(op, n, nbins, w, vec_ok, st);
                if (rc) return rc;
                finalize64_kernel<<<(unsigned)((out_elems + 255) / 256), 256, 0, st>>>(reinterpret_cast<const long long*>(w.cnt), out, out_elems,
                                                                                      1.0 / (double)(1 << kFixBits));
                BMC_CUDA(cudaGetLastError());
                return BMC_OK;
            }
        }
        if (out_elems <= kMaxBinsSmem32) rc = launch_mode<Op, kSmem32>(op, n, nbins, w, vec_ok, st);
        else if (!Op::kFloat && !Op::kSigned && out_elems <= kMaxBinsSmem16)
            rc = launch_mode<Op, kSmem16>(op, n, nbins, w, vec_ok, st);
        else rc = launch_mode<Op, kGlobal>(op, n, nbins, w, vec_ok, st);
        if (rc) return rc;
    }
    const int thr = 256;
    finalize_kernel<<<(unsigned)((out_elems + thr - 1) / thr), thr, 0, st>>>(w.cnt, w.ext, out, out_elems);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

int run_voxel_pairs(const VoxelOp& op, long n, float* out, void* ws, size_t wsb, cudaStream_t st) {
    const int plane = op.H * op.W;
    const long out_elems = (long)op.bins * plane;
    Ws w;
    int rc = carve(ws, wsb, out_elems, w);
    if (rc) return rc;
    // the float2 slot grid [bins-1][plane] overlays the cnt + ext areas (2*bins*plane words)
    float2* pairs = reinterpret_cast<float2*>(w.cnt);
    BMC_CUDA(cudaMemsetAsync(pairs, 0, (size_t)(op.bins - 1) * plane * 8, st));
    if (n > 0) {
        const int vec_ok = (((uintptr_t)op.xs | (uintptr_t)op.ys | (uintptr_t)op.ps | (uintptr_t)op.ts) & 15) == 0;
        long grid = (long)sm_count() * 4;
        const long want = (n + (long)kPairThreads * 4 * 8 - 1) / ((long)kPairThreads * 4 * 8);
        if (grid > want) grid = want < 1 ? 1 : want;
        voxel_pair_kernel<<<(unsigned)grid, kPairThreads, 0, st>>>(op, n, pairs, vec_ok);
        BMC_CUDA(cudaGetLastError());
    }
    const int thr = 256;
    finalize_pairs_kernel<<<(unsigned)((out_elems + thr - 1) / thr), thr, 0, st>>>(pairs, out, op.bins, plane);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

// ---- bins split over CTAs that stream the same events (role_kernel)
constexpr long kRoleMinEvents = 1L << 24;     // below this the plane flushes (groups x plane atomics) outweigh the gain

bool roles_enabled() {
    static int on = -1;
    if (on < 0) on = measure_env("BMC_ENC_ROLES", 1);
    return on != 0;
}

template <class R>
int run_roles(R& pol, long n, int roles, long out_elems, void* ws, size_t wsb, Ws& w, cudaStream_t st) {
    int rc = carve(ws, wsb, out_elems, w);
    if (rc) return rc;
    BMC_CUDA(cudaMemsetAsync(w.cnt, 0, (size_t)out_elems * 8, st));
    BMC_CUDA(cudaMemsetAsync(w.oor_count, 0, 8, st));
    pol.oor.count = w.oor_count; pol.oor.idx = w.oor_idx;
    const int vec_ok = (((uintptr_t)pol.op.xs | (uintptr_t)pol.op.ys | (uintptr_t)pol.op.ps |
                         (uintptr_t)(R::kNeedT ? pol.op.ts : nullptr)) & 15) == 0;
    long groups = sm_count() / roles;
    const long want = n >> 20;                    // ~1M events per group amortise a plane flush
    if (groups > want) groups = want < 1 ? 1 : want;
    auto kern = role_kernel<R>;
    const size_t smem = (size_t)kRoleWords * 4;
    BMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)(groups * roles), kRoleThreads, smem, st>>>(pol, n, roles, vec_ok);
    BMC_CUDA(cudaGetLastError());
    if (pol.mutate) {
        oor_zero_kernel<<<sm_count(), 256, 0, st>>>(pol.op.xs, pol.op.ys, n, pol.op.H, pol.op.W, w.oor_count, w.oor_idx);
        BMC_CUDA(cudaGetLastError());
    }
    return BMC_OK;
}

int check_common(const void* xs, const void* ys, const void* ps, long n, int H, int W, const void* out) {
    BMC_REQUIRE(n >= 0 && H > 0 && W > 0, "encoder: bad sizes n=%ld H=%d W=%d", n, H, W);
    BMC_REQUIRE(out != nullptr, "encoder: out is NULL");
    BMC_REQUIRE(n == 0 || (xs && ys && ps), "encoder: NULL event array");
    BMC_REQUIRE((long)H * W * 2 < (1L << 30), "encoder: grid too large");
    return BMC_OK;
}

}  // namespace
}  // namespace bmc

using namespace bmc;

extern "C" BMC_EXPORT size_t bmc_encode_workspace_bytes(int64_t out_elems) { return ws_bytes(out_elems); }
extern "C" BMC_EXPORT size_t bmc_encode_stack_flag_offset(int64_t out_elems) { return ws_flag_offset(out_elems); }

extern "C" BMC_EXPORT int bmc_encode_channels(float* xs, float* ys, const float* ps, int64_t n, int H, int W,
                                   float* out, void* workspace, size_t workspace_bytes,
                                   unsigned flags, void* stream) {
    int rc = check_common(xs, ys, ps, n, H, W, out);
    if (rc) return rc;
    ChannelsOp op;
    op.xs = xs; op.ys = ys; op.ts = nullptr; op.ps = const_cast<float*>(ps);
    op.H = H; op.W = W; op.bins = 1; op.flags = flags | BMC_ENC_FLIP_Y;
    if (n > 0 && n <= kOneCtaMaxEvents && (size_t)2 * H * W * 4 <= (size_t)kSmemBudget && !(flags & BMC_ENC_SPLIT_BINS)) {
        const size_t smem = (size_t)2 * H * W * 4;             // one window: a single one-CTA launch
        if (smem > 48 * 1024)
            BMC_CUDA(cudaFuncSetAttribute(channels_one_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        channels_one_kernel<<<1, 1024, smem, as_stream(stream)>>>(xs, ys, ps, (long)n, H, W, out, flags);
        BMC_CUDA(cudaGetLastError());
        return BMC_OK;
    }
    if (2L * H * W > kMaxBinsSmem16 && (H * W + 3) / 4 <= kRoleWords && (n >= kRoleMinEvents || (flags & BMC_ENC_SPLIT_BINS)) && n > 0 &&
        roles_enabled()) {
        CountsRole pol;                     // up to 360x640: one CTA per polarity over the same events, 8-bit counter planes
        pol.op = op; pol.op.flags &= ~BMC_ENC_MUTATE; pol.mutate = (flags & BMC_ENC_MUTATE) ? 1 : 0;
        Ws w;
        cudaStream_t st = as_stream(stream);
        const long elems = 2L * H * W;
        rc = carve(workspace, workspace_bytes, elems, w);
        if (rc) return rc;
        pol.g_cnt = w.cnt; pol.g_ext = w.ext;
        rc = run_roles(pol, n, 2, elems, workspace, workspace_bytes, w, st);
        if (rc) return rc;
        finalize_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(w.cnt, w.ext, out, elems);
        BMC_CUDA(cudaGetLastError());
        return BMC_OK;
    }
    return run_scatter(op, n, 2L * H * W, out, workspace, workspace_bytes, as_stream(stream));
}

extern "C" BMC_EXPORT int bmc_encode_channels_windows(float* xs, float* ys, const float* ps,
                                           const int64_t* offsets, int n_windows, int H, int W,
                                           float* out, unsigned flags, void* stream) {
    BMC_REQUIRE(n_windows >= 0 && H > 0 && W > 0 && out && offsets, "encode_channels_windows: bad args");
    if (n_windows == 0) return BMC_OK;
    const size_t smem = (size_t)2 * H * W * 4;
    BMC_REQUIRE(smem <= (size_t)kSmemBudget,
                "encode_channels_windows: a [2,%d,%d] window grid exceeds 227 KB of shared memory; "
                "encode such windows one by one with bmc_encode_channels", H, W);
    if (smem > 48 * 1024)
        BMC_CUDA(cudaFuncSetAttribute(channels_windows_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    channels_windows_kernel<<<n_windows, 256, smem, as_stream(stream)>>>(
        xs, ys, ps, reinterpret_cast<const long*>(offsets), H, W, out, flags);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

extern "C" BMC_EXPORT int bmc_encode_channels_windows_raw(const int16_t* xs, const int16_t* ys, const double* ps,
                                                          int64_t n_events, int64_t window, int64_t stride, int n_windows,
                                                          int H, int W, float* out, unsigned flags, void* stream) {
    BMC_REQUIRE(n_windows >= 0 && H > 0 && W > 0, "encode_channels_windows_raw: bad sizes");
    if (n_windows == 0) return BMC_OK;
    BMC_REQUIRE(out && xs && ys && ps, "encode_channels_windows_raw: NULL argument");
    BMC_REQUIRE(window > 0 && stride > 0 && n_events >= 0, "encode_channels_windows_raw: window and stride must be positive");
    BMC_REQUIRE(n_windows == 0 || stride * (int64_t)(n_windows - 1) <= n_events - 1,
                "encode_channels_windows_raw: window %d starts past the recording (h5dataset.py:327-328)", n_windows - 1);
    if (n_windows == 0) return BMC_OK;
    const size_t smem = (size_t)2 * H * W * 4;
    BMC_REQUIRE(smem <= (size_t)kSmemBudget, "encode_channels_windows_raw: a [2,%d,%d] window grid exceeds 227 KB of shared memory", H, W);
    if (smem > 48 * 1024)
        BMC_CUDA(cudaFuncSetAttribute(channels_windows_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    channels_windows_raw_kernel<<<n_windows, 256, smem, as_stream(stream)>>>(
        reinterpret_cast<const short*>(xs), reinterpret_cast<const short*>(ys), ps, (long)n_events, (long)window, (long)stride,
        H, W, out, flags);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

extern "C" BMC_EXPORT int bmc_format_events(const int16_t* xs, const int16_t* ys, const double* ts, const double* ps,
                                            int64_t n, float* out, void* stream) {
    BMC_REQUIRE(n >= 1 && xs && ys && ts && ps && out, "format_events: needs n >= 1 events (the reference indexes ts[0], ts[-1])");
    long blocks = (n + 255) / 256;
    const long cap = (long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    format_events_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const short*>(xs),
                                                                        reinterpret_cast<const short*>(ys), ts, ps, (long)n, out);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

extern "C" BMC_EXPORT int bmc_encode_image(float* xs, float* ys, float* ps, int64_t n, int H, int W,
                                float* out, void* workspace, size_t workspace_bytes,
                                unsigned flags, void* stream) {
    int rc = check_common(xs, ys, ps, n, H, W, out);
    if (rc) return rc;
    ImageOp op;
    op.xs = xs; op.ys = ys; op.ts = nullptr; op.ps = ps;
    op.H = H; op.W = W; op.bins = 1; op.flags = flags;
    const long elems = (flags & BMC_ENC_BILINEAR) ? (long)(H + 1) * (W + 1) : (long)H * W;
    return run_scatter(op, n, elems, out, workspace, workspace_bytes, as_stream(stream));
}

extern "C" BMC_EXPORT int bmc_encode_voxel(float* xs, float* ys, const float* ts, const float* ps, int64_t n,
                                int bins, int H, int W, float* out, void* workspace,
                                size_t workspace_bytes, unsigned flags, void* stream) {
    int rc = check_common(xs, ys, ps, n, H, W, out);
    if (rc) return rc;
    BMC_REQUIRE(bins >= 1 && bins <= 64, "encode_voxel: bins must be in [1,64], got %d", bins);
    BMC_REQUIRE(n == 0 || ts, "encode_voxel: ts is NULL");
    VoxelOp op;
    op.xs = xs; op.ys = ys; op.ts = ts; op.ps = const_cast<float*>(ps);
    op.H = H; op.W = W; op.bins = bins; op.flags = flags;
    op.t0 = 0.f; op.dt = 1.f;     // BMC_ENC_TNORM: filled in on the device (VoxelOp::prepare)
    if (bins >= 2 && (long)bins * H * W > kMaxBinsSmemFix && H * W <= kRoleWords && (n >= kRoleMinEvents || (flags & BMC_ENC_SPLIT_BINS)) &&
        n > 0 && !(flags & BMC_ENC_DETERMINISTIC) && roles_enabled()) {
        VoxelRole pol;                      // up to 180x320: one time slot's state split over two CTAs (role_kernel)
        pol.op = op; pol.op.flags &= ~BMC_ENC_MUTATE; pol.mutate = (flags & BMC_ENC_MUTATE) ? 1 : 0;
        Ws w;
        cudaStream_t st = as_stream(stream);
        const long elems = (long)bins * H * W;
        rc = carve(workspace, workspace_bytes, elems, w);
        if (rc) return rc;
        pol.g64 = reinterpret_cast<unsigned long long*>(w.cnt);
        rc = run_roles(pol, n, 2, elems, workspace, workspace_bytes, w, st);
        if (rc) return rc;
        finalize64_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(reinterpret_cast<const long long*>(w.cnt), out, elems,
                                                                          1.0 / (double)(1 << kFixBits));
        BMC_CUDA(cudaGetLastError());
        return BMC_OK;
    }
    if (bins >= 2 && (long)bins * H * W > kMaxBinsSmem32 && !(flags & BMC_ENC_DETERMINISTIC))      // too large for shared-memory bins
        return run_voxel_pairs(op, n, out, workspace, workspace_bytes, as_stream(stream));
    return run_scatter(op, n, (long)bins * H * W, out, workspace, workspace_bytes, as_stream(stream));
}

extern "C" BMC_EXPORT int bmc_encode_stack_shard(float* xs, float* ys, float* ps, int64_t n_local,
                                const float* ts_all, int64_t n_total, int64_t first, int bins, int H, int W,
                                int polarity, float* out, void* workspace, size_t workspace_bytes,
                                unsigned flags, void* stream) {
    int rc = check_common(xs, ys, ps, n_local, H, W, out);
    if (rc) return rc;
    BMC_REQUIRE(bins >= 1 && bins <= 64, "encode_stack: bins must be in [1,64], got %d", bins);
    BMC_REQUIRE(n_total > 3 && ts_all, "encode_stack: n <= 3 is the reference's early-out (caller returns zeros)");
    BMC_REQUIRE(first >= 0 && first + n_local <= n_total, "encode_stack_shard: events [%ld, %ld) outside [0, %ld)",
                (long)first, (long)(first + n_local), (long)n_total);
    const long elems = (long)(polarity ? 2 : 1) * bins * H * W;
    Ws w;
    rc = carve(workspace, workspace_bytes, elems, w);
    if (rc) return rc;
    bin_bounds_kernel<<<(2 * bins * 32 + kBoundsThreads - 1) / kBoundsThreads, kBoundsThreads, 0, as_stream(stream)>>>(
        ts_all, n_total, bins, w.beg, w.end, first, (flags & BMC_ENC_SKIP_ZERO_ENDS) ? 1 : 0,
        reinterpret_cast<int*>(static_cast<char*>(workspace) + ws_flag_offset(elems)));
    BMC_CUDA(cudaGetLastError());
    StackOp op;
    op.xs = xs; op.ys = ys; op.ts = nullptr; op.ps = ps;
    op.H = H; op.W = W; op.bins = bins; op.flags = flags & ~BMC_ENC_FLIP_Y;
    op.beg = w.beg; op.end = w.end; op.polarity = polarity;
    static int by_bin = -1;
    if (by_bin < 0) by_bin = measure_env("BMC_ENC_STACK_BINS", 1);
    const size_t bin_smem = (size_t)H * W * 4;
    if (by_bin && elems > kMaxBinsSmem32 && bin_smem + 2048 <= (size_t)kSmemBudget && bins <= sm_count() &&
        n_local >= (long)bins * 65536) {
        // the planes of one time bin fit shared memory: one bin per CTA (stack_bins_kernel)
        cudaStream_t st = as_stream(stream);
        BMC_CUDA(cudaMemsetAsync(w.cnt, 0, (size_t)elems * 8, st));
        const int vec_ok = (((uintptr_t)xs | (uintptr_t)ys | (uintptr_t)ps) & 15) == 0;
        auto kern = polarity ? stack_bins_kernel<true> : stack_bins_kernel<false>;
        BMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_smem));
        kern<<<sm_count(), kBinThreads, bin_smem, st>>>(op, n_local, w.cnt, w.ext, vec_ok);
        BMC_CUDA(cudaGetLastError());
        if ((flags & BMC_ENC_MUTATE) && bins > 1) {
            if (polarity) stack_bins_mutate_kernel<true><<<bins - 1, 256, 0, st>>>(op, n_local);
            else stack_bins_mutate_kernel<false><<<bins - 1, 256, 0, st>>>(op, n_local);
            BMC_CUDA(cudaGetLastError());
        }
        const int thr = 256;
        finalize_kernel<<<(unsigned)((elems + thr - 1) / thr), thr, 0, st>>>(w.cnt, w.ext, out, elems);
        BMC_CUDA(cudaGetLastError());
        return BMC_OK;
    }
    return run_scatter(op, n_local, elems, out, workspace, workspace_bytes, as_stream(stream));
}

extern "C" BMC_EXPORT int bmc_encode_stack(float* xs, float* ys, const float* ts, float* ps, int64_t n,
                                int bins, int H, int W, int polarity, float* out, void* workspace,
                                size_t workspace_bytes, unsigned flags, void* stream) {
    BMC_REQUIRE(n > 3 && ts, "encode_stack: n <= 3 is the reference's early-out (caller returns zeros)");
    return bmc_encode_stack_shard(xs, ys, ps, n, ts, n, 0, bins, H, W, polarity, out, workspace, workspace_bytes,
                                  flags, stream);
}

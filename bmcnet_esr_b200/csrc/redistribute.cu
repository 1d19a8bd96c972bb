// Inverse encoders (reference dataloader/encodings.py:367-464 python_event_redistribute_{Polarity,NoPolarity}Stack and
// :653-671 stack2cnt): a stack of per-bin event counts back to an event cloud.
//
// Reference semantics, per batch entry: every voxel (nonzero order = row-major [p,] c, y, x) with rounded value v
// emits |v| events at (x, y) with polarity sign(v) and timestamps torch.linspace(t0, t1, |v|), t0 = c/C + 1/(100 C),
// t1 = (c+1)/C ('linear'), or rand * (t1 - t0) + t0 ('random'); the entry's events are then STABLY sorted by t and
// the batch is zero-padded to the longest entry.  The reference does this with Python loops over voxels and a
// Python `sorted` over 0-dim tensors (its Cython replacement `c_event_redistribute` was never shipped).
//
// Here: count -> per-entry exclusive scan -> expand (one thread per voxel) -> per-entry stable LSD radix sort of
// (timestamp bits, original index) -> gather.  Timestamps are positive floats, so their bit patterns sort like
// the values.  Not a bandwidth-critical path (the reference never calls it from infer / train): kernels are
// straightforward, one CTA per batch entry for the scan and the sort.
#include "common.cuh"

namespace bmc {
namespace {

constexpr int kSortThreads = 1024;

__device__ __forceinline__ int voxel_count(float v) { return (int)fabsf(rintf(v)); }      // torch.round: half to even

// counts per voxel, per-entry totals and the per-entry sum of the rounded values (`entry.sum() != 0`, :383,432)
__global__ void redis_count_kernel(const float* __restrict__ stack, long per_entry, int* __restrict__ cnt,
                                   unsigned long long* __restrict__ totals, long long* __restrict__ sums) {
    const int b = blockIdx.y;
    const float* s = stack + (long)b * per_entry;
    unsigned long long tot = 0;
    long long sum = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < per_entry; i += (long)gridDim.x * blockDim.x) {
        const float r = rintf(s[i]);
        const int n = (int)fabsf(r);
        cnt[(long)b * per_entry + i] = n;
        tot += (unsigned long long)n;
        sum += (long long)r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (tot) atomicAdd(&totals[b], tot);
        if (sum) atomicAdd(reinterpret_cast<unsigned long long*>(&sums[b]), (unsigned long long)sum);
    }
}

// exclusive scan of cnt over one entry (in place -> offsets); one CTA per entry, running carry over 1024-element chunks
__global__ void __launch_bounds__(kSortThreads) redis_scan_kernel(int* __restrict__ cnt, long per_entry) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    int* c = cnt + (long)blockIdx.x * per_entry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (long base = 0; base < per_entry; base += kSortThreads) {
        const long i = base + threadIdx.x;
        const int v = i < per_entry ? c[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;                       // inclusive over warps
        }
        __syncthreads();
        const int carry = carry_s;
        const int before = carry + (warp ? warp_sums[warp - 1] : 0) + incl - v;
        if (i < per_entry) c[i] = before;
        __syncthreads();
        if (threadIdx.x == kSortThreads - 1) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
}

// one thread per voxel: its events into the entry's unsorted list (x, y, t, p) + sort keys
__global__ void redis_expand_kernel(const float* __restrict__ stack, const int* __restrict__ offs, long per_entry,
                                    int P, int C, int Y, int X, long maxlen, const float* __restrict__ rnd,
                                    const long long* __restrict__ sums,
                                    float4* __restrict__ tmp, unsigned* __restrict__ keys, unsigned* __restrict__ idx) {
    const int b = blockIdx.y;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_entry || sums[b] == 0) return;        // an entry summing to zero is "empty" (its length is not part of maxlen)
    const float r = rintf(stack[(long)b * per_entry + i]);
    const int n = (int)fabsf(r);
    if (n == 0) return;
    const int x = (int)(i % X);
    const int y = (int)((i / X) % Y);
    const int c = (int)((i / ((long)X * Y)) % C);
    (void)P;
    const float t0 = __fadd_rn(__fdiv_rn((float)c, (float)C), (float)(1.0 / (100.0 * (double)C)));     // :387-388
    const float t1 = __fdiv_rn((float)(c + 1), (float)C);
    const float step = n > 1 ? __fdiv_rn(__fsub_rn(t1, t0), (float)(n - 1)) : 0.f;                      // torch.linspace, float32
    const int half = n / 2;
    const float pol = r > 0.f ? 1.f : -1.f;
    const long o = (long)b * maxlen + offs[(long)b * per_entry + i];
    for (int k = 0; k < n; ++k) {
        float t;
        if (rnd) t = __fadd_rn(__fmul_rn(rnd[o + k], __fsub_rn(t1, t0)), t0);                           // torch.rand * (t1-t0) + t0
        else if (n == 1) t = t0;
        else t = k < half ? fmaf(step, (float)k, t0) : fmaf(-step, (float)(n - 1 - k), t1);
        tmp[o + k] = make_float4((float)x, (float)y, t, pol);
        keys[o + k] = __float_as_uint(t);
        idx[o + k] = (unsigned)(offs[(long)b * per_entry + i] + k);
    }
}

// Stable LSD radix sort of one entry's (key, idx) pairs, 4 bits per pass, one CTA per entry.  Every thread owns a
// contiguous slice: it counts its 16 digits, a block-wide scan over (digit-major, thread-minor) gives its write
// positions, and it scatters its slice in order -- stable by construction.
__global__ void __launch_bounds__(kSortThreads) redis_sort_kernel(unsigned* keys_a, unsigned* idx_a, unsigned* keys_b,
                                                                  unsigned* idx_b, const unsigned long long* __restrict__ totals,
                                                                  const long long* __restrict__ sums, long maxlen) {
    extern __shared__ int hist[];                      // [16][kSortThreads]
    __shared__ int warp_sums[32];
    const int b = blockIdx.x;
    const long n = sums[b] == 0 ? 0 : (long)totals[b];
    if (n <= 1) return;
    unsigned* ka = keys_a + (long)b * maxlen; unsigned* ia = idx_a + (long)b * maxlen;
    unsigned* kb = keys_b + (long)b * maxlen; unsigned* ib = idx_b + (long)b * maxlen;
    const long per = (n + kSortThreads - 1) / kSortThreads;
    const long beg = min(n, (long)threadIdx.x * per), end = min(n, beg + per);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int pass = 0; pass < 8; ++pass) {
        const int sh = pass * 4;
        int c[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) c[d] = 0;
        for (long i = beg; i < end; ++i) {
            const int dg = (ka[i] >> sh) & 15;
#pragma unroll
            for (int d = 0; d < 16; ++d) c[d] += (dg == d);
        }
#pragma unroll
        for (int d = 0; d < 16; ++d) hist[d * kSortThreads + threadIdx.x] = c[d];
        __syncthreads();
        // exclusive scan over the 16 * 1024 counters in (digit, thread) order: each thread scans 16 consecutive entries
        int loc[16], s = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) { loc[j] = hist[threadIdx.x * 16 + j]; s += loc[j]; }
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        int run = (warp ? warp_sums[warp - 1] : 0) + incl - s;
#pragma unroll
        for (int j = 0; j < 16; ++j) { hist[threadIdx.x * 16 + j] = run; run += loc[j]; }
        __syncthreads();
        int pos[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) pos[d] = hist[d * kSortThreads + threadIdx.x];
        for (long i = beg; i < end; ++i) {
            const unsigned k = ka[i], v = ia[i];
            const int dg = (k >> sh) & 15;
            int p = 0;
#pragma unroll
            for (int d = 0; d < 16; ++d) if (dg == d) { p = pos[d]; pos[d] = p + 1; }
            kb[p] = k; ib[p] = v;
        }
        __syncthreads();
        unsigned* t1 = ka; ka = kb; kb = t1;
        unsigned* t2 = ia; ia = ib; ib = t2;
    }
    // 8 passes: the result is back in the *_a buffers
}

__global__ void redis_gather_kernel(const float4* __restrict__ tmp, const unsigned* __restrict__ idx,
                                    const unsigned long long* __restrict__ totals, const long long* __restrict__ sums,
                                    long maxlen, float4* __restrict__ out) {
    const int b = blockIdx.y;
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= maxlen) return;
    // an entry whose rounded values sum to zero is "empty" in the reference (a single zero row), whatever it holds
    const bool live = sums[b] != 0 && r < (long)totals[b];
    out[(long)b * maxlen + r] = live ? tmp[(long)b * maxlen + idx[(long)b * maxlen + r]] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// stack2cnt (encodings.py:653-671): [B,TB,H,W] -> [B,2,H,W], positive and negative parts of the rounded stack summed over bins
__global__ void stack2cnt_kernel(const float* __restrict__ stack, int TB, long plane, float* __restrict__ out) {
    const int b = blockIdx.y;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane) return;
    float pos = 0.f, neg = 0.f;
    for (int t = 0; t < TB; ++t) {
        const float r = rintf(stack[((long)b * TB + t) * plane + i]);
        if (r > 0.f) pos += r; else if (r < 0.f) neg += -r;
    }
    out[((long)b * 2) * plane + i] = pos;
    out[((long)b * 2 + 1) * plane + i] = neg;
}

struct RedisWs {
    int* cnt; unsigned long long* totals; long long* sums; float4* tmp; unsigned *keys_a, *idx_a, *keys_b, *idx_b;
};
size_t redis_ws_bytes(int B, long per_entry, long maxlen) {
    return (size_t)B * per_entry * 4 + 256 + (size_t)B * 16 + 256 + (size_t)B * maxlen * (16 + 16) + 1024;
}
void redis_carve(void* ws, int B, long per_entry, long maxlen, RedisWs& w) {
    char* p = static_cast<char*>(ws);
    auto take = [&](size_t bytes) { char* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
    w.totals = reinterpret_cast<unsigned long long*>(take((size_t)B * 8));
    w.sums = reinterpret_cast<long long*>(take((size_t)B * 8));
    w.cnt = reinterpret_cast<int*>(take((size_t)B * per_entry * 4));
    w.tmp = reinterpret_cast<float4*>(take((size_t)B * maxlen * 16));
    w.keys_a = reinterpret_cast<unsigned*>(take((size_t)B * maxlen * 4));
    w.idx_a = reinterpret_cast<unsigned*>(take((size_t)B * maxlen * 4));
    w.keys_b = reinterpret_cast<unsigned*>(take((size_t)B * maxlen * 4));
    w.idx_b = reinterpret_cast<unsigned*>(take((size_t)B * maxlen * 4));
}

}  // namespace
}  // namespace bmc

using namespace bmc;

extern "C" BMC_EXPORT size_t bmc_stack_to_events_workspace_bytes(int B, int64_t per_entry, int64_t maxlen) {
    return redis_ws_bytes(B, (long)per_entry, (long)maxlen) + 4096;
}

// Phase 1: per-entry event totals and rounded sums (the caller sizes the output from them, like the reference's maxlen).
extern "C" BMC_EXPORT int bmc_stack_event_counts(const float* stack, int B, int64_t per_entry, void* workspace,
                                                 size_t workspace_bytes, int64_t* totals_out, int64_t* sums_out, void* stream) {
    BMC_REQUIRE(stack && workspace && totals_out && sums_out && B > 0 && per_entry > 0, "stack_event_counts: bad argument");
    BMC_REQUIRE(per_entry < (1L << 31), "stack_event_counts: entry too large");
    BMC_REQUIRE(workspace_bytes >= bmc_stack_to_events_workspace_bytes(B, per_entry, 0), "stack_event_counts: workspace too small");
    BMC_REQUIRE(((uintptr_t)workspace & 255) == 0, "stack_event_counts: workspace must be 256-byte aligned");
    cudaStream_t st = as_stream(stream);
    RedisWs w;
    redis_carve(workspace, B, (long)per_entry, 0, w);
    BMC_CUDA(cudaMemsetAsync(w.totals, 0, (size_t)B * 8, st));
    BMC_CUDA(cudaMemsetAsync(w.sums, 0, (size_t)B * 8, st));
    dim3 grid((unsigned)std::min<long>((per_entry + 255) / 256, 1024), (unsigned)B);
    redis_count_kernel<<<grid, 256, 0, st>>>(stack, (long)per_entry, w.cnt, w.totals, w.sums);
    BMC_CUDA(cudaGetLastError());
    BMC_CUDA(cudaMemcpyAsync(totals_out, w.totals, (size_t)B * 8, cudaMemcpyDeviceToDevice, st));
    BMC_CUDA(cudaMemcpyAsync(sums_out, w.sums, (size_t)B * 8, cudaMemcpyDeviceToDevice, st));
    return BMC_OK;
}

// Phase 2: the event cloud [B][maxlen][4] = (x, y, t, p), sorted by t per entry, zero-padded.  `rnd` (device float
// [B][maxlen], uniform [0,1)) selects mode='random', NULL mode='linear'.  Must follow bmc_stack_event_counts on the
// same stack with a workspace sized for this maxlen.
extern "C" BMC_EXPORT int bmc_stack_to_events(const float* stack, int B, int P, int C, int Y, int X, int64_t maxlen,
                                              const float* rnd, float* out, void* workspace, size_t workspace_bytes,
                                              void* stream) {
    BMC_REQUIRE(stack && out && workspace && B > 0 && (P == 1 || P == 2) && C > 0 && Y > 0 && X > 0 && maxlen >= 1,
                "stack_to_events: bad argument");
    const long per_entry = (long)P * C * Y * X;
    BMC_REQUIRE(workspace_bytes >= bmc_stack_to_events_workspace_bytes(B, per_entry, maxlen), "stack_to_events: workspace too small");
    BMC_REQUIRE(((uintptr_t)workspace & 255) == 0, "stack_to_events: workspace must be 256-byte aligned");
    cudaStream_t st = as_stream(stream);
    RedisWs w;
    redis_carve(workspace, B, per_entry, (long)maxlen, w);
    // recount (the layout of the workspace depends on maxlen), scan, expand, sort, gather
    BMC_CUDA(cudaMemsetAsync(w.totals, 0, (size_t)B * 8, st));
    BMC_CUDA(cudaMemsetAsync(w.sums, 0, (size_t)B * 8, st));
    dim3 g1((unsigned)std::min<long>((per_entry + 255) / 256, 1024), (unsigned)B);
    redis_count_kernel<<<g1, 256, 0, st>>>(stack, per_entry, w.cnt, w.totals, w.sums);
    redis_scan_kernel<<<B, kSortThreads, 0, st>>>(w.cnt, per_entry);
    dim3 g2((unsigned)((per_entry + 255) / 256), (unsigned)B);
    redis_expand_kernel<<<g2, 256, 0, st>>>(stack, w.cnt, per_entry, P, C, Y, X, (long)maxlen, rnd, w.sums, w.tmp, w.keys_a, w.idx_a);
    const int smem = 16 * kSortThreads * 4;
    static PerDevice configured_dev;
    int& configured = configured_dev.cur();
    if (!configured) {
        BMC_CUDA(cudaFuncSetAttribute(redis_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = 1;
    }
    redis_sort_kernel<<<B, kSortThreads, smem, st>>>(w.keys_a, w.idx_a, w.keys_b, w.idx_b, w.totals, w.sums, (long)maxlen);
    dim3 g3((unsigned)((maxlen + 255) / 256), (unsigned)B);
    redis_gather_kernel<<<g3, 256, 0, st>>>(w.tmp, w.idx_a, w.totals, w.sums, (long)maxlen, reinterpret_cast<float4*>(out));
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

extern "C" BMC_EXPORT int bmc_stack2cnt(const float* stack, int B, int TB, int H, int W, float* out, void* stream) {
    BMC_REQUIRE(stack && out && B > 0 && TB > 0 && H > 0 && W > 0, "stack2cnt: bad argument");
    const long plane = (long)H * W;
    dim3 grid((unsigned)((plane + 255) / 256), (unsigned)B);
    stack2cnt_kernel<<<grid, 256, 0, as_stream(stream)>>>(stack, TB, plane, out);
    BMC_CUDA(cudaGetLastError());
    return BMC_OK;
}

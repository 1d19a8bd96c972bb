// Microbenchmark: throughput of distributed-shared-memory atomics in a cluster of 8 CTAs (is a cluster-privatised
// histogram for grids that do not fit one SM's shared memory worth building?).  Every thread adds 1 to a pseudo-random
// counter of a pseudo-random CTA of its cluster.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dsmem_atom_bench tools/dsmem_atom_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
constexpr int kWords = 28800;      // 115,200 B per CTA (x8 = 921,600 B: the 16-bit packed 2x360x640 grid)
template <int MODE>   // 0: red (no return) u32, 1: atom (returning) u32, 2: local only
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(1024) k(unsigned* out, long iters) {
    extern __shared__ unsigned s[];
    cg::cluster_group cl = cg::this_cluster();
    for (int i = threadIdx.x; i < kWords; i += 1024) s[i] = 0;
    cl.sync();
    unsigned x = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
    unsigned acc = 0;
    for (long it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        const unsigned w = (x >> 8) % kWords, c = MODE == 2 ? cl.block_rank() : (x >> 28) & 7u;
        unsigned* remote = cl.map_shared_rank(s, c);
        if (MODE == 1) acc += atomicAdd(remote + w, 1u);
        else atomicAdd(remote + w, 1u);
    }
    cl.sync();
    if (threadIdx.x == 0) out[blockIdx.x] = s[0] + acc;
}
template <int MODE>
void run(const char* name, unsigned* out) {
    auto kern = k<MODE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kWords * 4);
    const int grid = 144;
    const long iters = 4000;
    kern<<<grid, 1024, kWords * 4>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    kern<<<grid, 1024, kWords * 4>>>(out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s %8.3f ms  %7.1f G atomics/s  (%s)\n", name, ms, grid * 1024.0 * iters / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
// the same access pattern with fp32 adds (compile to ATOMS.CAST.SPIN loops) and with two adds per iteration
template <int MODE>   // 0: float atomicAdd, 1: int atomicAdd with overflow test on the returned value, 2: two int adds
__global__ void __launch_bounds__(1024) kl(unsigned* out, long iters) {
    extern __shared__ unsigned s[];
    float* sf = reinterpret_cast<float*>(s);
    for (int i = threadIdx.x; i < kWords; i += 1024) s[i] = 0;
    __syncthreads();
    unsigned x = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
    unsigned acc = 0;
    for (long it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        const unsigned w = (x >> 8) % kWords;
        if (MODE == 0) atomicAdd(sf + w, 0.25f);
        else if (MODE == 1) { const int old = atomicAdd(reinterpret_cast<int*>(s) + w, 37); if (old > (1 << 30)) acc += 1; }
        else { atomicAdd(s + w, 1u); atomicAdd(s + (w ^ 1), 1u); }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[0] + acc;
}
template <int MODE>
void runl(const char* name, unsigned* out, int per_iter) {
    auto kern = kl<MODE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kWords * 4);
    const int grid = 148;
    const long iters = 4000;
    kern<<<grid, 1024, kWords * 4>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    kern<<<grid, 1024, kWords * 4>>>(out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s %8.3f ms  %7.1f G atomics/s  (%s)\n", name, ms, grid * 1024.0 * iters * per_iter / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    unsigned* out; cudaMalloc(&out, 4096);
    runl<0>("local fp32 atomicAdd (CAS)", out, 1);
    runl<1>("local int add + ovf test", out, 1);
    runl<2>("local 2 int adds / iter", out, 2);
    run<2>("local shared atomicAdd", out);
    run<0>("cluster-8 red (no return)", out);
    run<1>("cluster-8 atom (returning)", out);
    return 0;
}

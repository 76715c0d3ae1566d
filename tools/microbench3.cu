// Micro-benchmark (round 2): cost of a 5-cell min-combine window per lane with DP-like addresses (lane i -> cell w0 + i + jitter,
// offers e = 0..4 to consecutive cells) for the primitives a 32-bit-key label ring could use, next to the 64-bit CAS pipeline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench3 tools/microbench3.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// MODE 0: 5 x red.shared.min.u32 (4 B cells)         1: 5 x atom.shared.min.u32, result used
// MODE 2: 5 x (ld.u64 pre-read, compare, CAS64 loop)  (8 B cells; the round-1 kernel's smem_min64)
// MODE 3: 5 x st.shared.u32                           4: 5 x red.min.u32 + 5 x red.max.u32 on an interleaved second word (8 B cells)
// MODE 5: 5 x red.shared.min.u32, predicated by a hash bit (about half of the offers suppressed)
template <int MODE>
__global__ void __launch_bounds__(512, 2) k(int iters, int jit, unsigned *out) {
    extern __shared__ unsigned long long sm[];      // 6400 x 8 B
    const int N = 6400;
    for (int i = threadIdx.x; i < N; i += blockDim.x) sm[i] = ~0ULL;
    __syncthreads();
    const unsigned base = smem_u32(sm);
    unsigned acc = 0, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int it = 0; it < iters; it++) {
        const unsigned h = hash32(it * 512 + tid);
        // smooth windows: lane i starts at w0 + i (+1 for a few lanes when jit): neighbours overlap in 4 of 5 cells
        unsigned cell = (warp * 397 + it * 61 + lane + ((jit && (h & 7) == 0) ? 1 : 0)) % (N - 48);
        const unsigned val = h | 1u;
#pragma unroll
        for (int e = 0; e < 5; e++) {
            const unsigned v = val + 977u * e;
            if (MODE == 0) asm volatile("red.shared.min.u32 [%0], %1;" :: "r"(base + 4u * (cell + e)), "r"(v) : "memory");
            if (MODE == 1) { unsigned o; asm volatile("atom.shared.min.u32 %0, [%1], %2;" : "=r"(o) : "r"(base + 4u * (cell + e)), "r"(v) : "memory"); acc += o; }
            if (MODE == 2) {
                const unsigned a = base + 8u * (cell + e); const unsigned long long vv = ((unsigned long long)v << 16) | lane;
                unsigned long long old; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(old) : "r"(a) : "memory");
                while (vv < old) { unsigned long long as = old; asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(a), "l"(as), "l"(vv) : "memory"); if (old == as) break; }
            }
            if (MODE == 3) asm volatile("st.shared.u32 [%0], %1;" :: "r"(base + 4u * (cell + e)), "r"(v) : "memory");
            if (MODE == 4) { asm volatile("red.shared.min.u32 [%0], %1;" :: "r"(base + 8u * (cell + e)), "r"(v) : "memory");
                             asm volatile("red.shared.max.u32 [%0], %1;" :: "r"(base + 8u * (cell + e) + 4u), "r"(~v) : "memory"); }
            if (MODE == 5) { if ((h >> (8 + e)) & 1u) asm volatile("red.shared.min.u32 [%0], %1;" :: "r"(base + 4u * (cell + e)), "r"(v) : "memory"); }
        }
        if ((it & 63) == 63) { __syncthreads(); for (int i = threadIdx.x; i < N; i += blockDim.x) sm[i] = ~0ULL; __syncthreads(); }
    }
    if (acc == 0x12345678) out[0] = acc;
}

template <int MODE> void run(const char *name, int jit) {
    int iters = 2048; unsigned *out; cudaMalloc(&out, 4);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6400 * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<296, 512, 6400 * 8>>>(iters, jit, out); cudaDeviceSynchronize();
    cudaEventRecord(a); k<MODE><<<296, 512, 6400 * 8>>>(iters, jit, out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double cyc = ms * 1e-3 * 1.9e9, wins = 32.0 * iters;        // per SM: 2 blocks x 16 warps x iters 5-cell windows
    printf("%-52s jitter=%d  %8.3f ms  %6.1f cycles per warp-window per SM\n", name, jit, ms, cyc / wins);
    cudaFree(out);
}
int main() {
    for (int jit : {0, 1}) {
        run<3>("5 x st.shared.u32", jit); run<0>("5 x red.shared.min.u32", jit); run<5>("5 x red.shared.min.u32, half predicated off", jit);
        run<1>("5 x atom.shared.min.u32 (result used)", jit); run<4>("5 x (red.min.u32 + red.max.u32)", jit);
        run<2>("5 x pre-read + CAS64 loop (round-1 smem_min64)", jit);
    }
    return 0;
}

// Micro-benchmark: throughput of the shared-memory primitives a min-combine can be built from (B200).
// Each warp issues `iters` warp-wide operations on DP-like addresses (lane i -> base + 2*i + small jitter).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// MODE 0: st.shared.u64   1: ld.shared.u64 (dependent use)   2: atom.cas.b64 with result   3: red.min.u32 (no result)
// MODE 4: atom.min.u32 with result   5: red.min.u32 x2 (two halves)   6: st.shared.u32   7: red.min.u64 (CAST.SPIN)
// MODE 8: pre-read + compare + CAS (the kernel's smem_min64)          9: atom.exch.b64 with result
template <int MODE>
__global__ void __launch_bounds__(512, 2) k(int iters, int dup, unsigned *out) {
    extern __shared__ unsigned long long sm[];      // 6400 words
    const int N = 6400;
    for (int i = threadIdx.x; i < N; i += blockDim.x) sm[i] = ~0ULL;
    __syncthreads();
    const unsigned base = smem_u32(sm);
    unsigned acc = 0, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int it = 0; it < iters; it++) {
        // 32 lanes -> mostly distinct cells, `dup` of 32 lanes share a cell with their neighbour
        unsigned cell = (warp * 397 + it * 61 + lane * 2 - ((lane < (unsigned)dup) ? (lane & 1) * 2 : 0)) % (N - 8);
        unsigned a = base + 8u * cell;
        unsigned long long val = ((unsigned long long)hash32(it * 512 + tid) << 32) | tid;
        if (MODE == 0) asm volatile("st.shared.u64 [%0], %1;" :: "r"(a), "l"(val) : "memory");
        if (MODE == 1) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); acc += (unsigned)v; }
        if (MODE == 2) { unsigned long long o; asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(o) : "r"(a), "l"(~0ULL), "l"(val) : "memory"); acc += (unsigned)o; }
        if (MODE == 3) asm volatile("red.shared.min.u32 [%0], %1;" :: "r"(a + 4), "r"((unsigned)(val >> 32)) : "memory");
        if (MODE == 4) { unsigned o; asm volatile("atom.shared.min.u32 %0, [%1], %2;" : "=r"(o) : "r"(a + 4), "r"((unsigned)(val >> 32)) : "memory"); acc += o; }
        if (MODE == 5) { asm volatile("red.shared.min.u32 [%0], %1;" :: "r"(a + 4), "r"((unsigned)(val >> 32)) : "memory");
                         asm volatile("red.shared.min.u32 [%0], %1;" :: "r"(a), "r"((unsigned)val) : "memory"); }
        if (MODE == 6) asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"((unsigned)val) : "memory");
        if (MODE == 7) asm volatile("red.shared.min.u64 [%0], %1;" :: "r"(a), "l"(val) : "memory");
        if (MODE == 8) {
            unsigned long long old; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(old) : "r"(a) : "memory");
            while (val < old) { unsigned long long as = old; asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(a), "l"(as), "l"(val) : "memory"); if (old == as) break; }
        }
        if (MODE == 9) { unsigned long long o; asm volatile("atom.shared.exch.b64 %0, [%1], %2;" : "=l"(o) : "r"(a), "l"(val) : "memory"); acc += (unsigned)o; }
        if ((it & 63) == 63) { __syncthreads(); for (int i = threadIdx.x; i < N; i += blockDim.x) sm[i] = ~0ULL; __syncthreads(); }   // keep cells "fresh"
    }
    if (acc == 0x12345678) out[0] = acc;
}

template <int MODE> void run(const char *name, int dup) {
    int iters = 4096; unsigned *out; cudaMalloc(&out, 4);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6400 * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<296, 512, 6400 * 8>>>(iters, dup, out); cudaDeviceSynchronize();
    cudaEventRecord(a); k<MODE><<<296, 512, 6400 * 8>>>(iters, dup, out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // per SM: 2 blocks x 16 warps x iters warp-ops in ms
    double cyc = ms * 1e-3 * 1.9e9, ops = 32.0 * iters;
    printf("%-44s dup=%2d  %8.3f ms  %6.1f cycles per warp-op per SM (incl. refresh)\n", name, dup, ms, cyc / ops);
    cudaFree(out);
}
int main() {
    for (int dup : {0, 4}) {
        run<0>("st.shared.u64", dup); run<6>("st.shared.u32", dup); run<1>("ld.shared.u64 (used)", dup);
        run<2>("atom.shared.cas.b64 (result used)", dup); run<9>("atom.shared.exch.b64 (result used)", dup);
        run<3>("red.shared.min.u32", dup); run<4>("atom.shared.min.u32 (result used)", dup);
        run<5>("2 x red.shared.min.u32", dup); run<7>("red.shared.min.u64 (CAST.SPIN loop)", dup);
        run<8>("pre-read + CAS loop (kernel's smem_min64)", dup);
    }
    return 0;
}

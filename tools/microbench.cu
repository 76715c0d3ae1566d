// Design micro-benchmarks for the DP kernels (B200): shared-memory atomics, fp64 vs fp32 issue rate,
// __syncthreads cost.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void atom_kernel(int iters, unsigned *out) {
    __shared__ unsigned sm[4096];
    __shared__ unsigned long long sm64[2048];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm64[i] = ~0ULL;
    __syncthreads();
    unsigned acc = 0, tid = threadIdx.x;
    for (int it = 0; it < iters; it++) {
        unsigned a;
        if (MODE == 0) a = (tid + it * 37) & 4095;                       // conflict-free, consecutive lanes
        else if (MODE == 1) a = hash32(tid * 977 + it) & 4095;           // random spread
        else if (MODE == 2) a = ((tid >> 1) * 3 + (it & 3)) & 4095;      // pairs of lanes collide (DP-like)
        else if (MODE == 3) a = (tid + it * 37) & 2047;                  // 64-bit min, consecutive
        else if (MODE == 4) a = ((tid >> 1) * 3 + (it & 3)) & 2047;      // 64-bit min, pairs collide
        else a = (tid + it * 37) & 4095;                                 // plain store baseline
        if (MODE <= 2) acc += atomicAdd(&sm[a], 1u);
        else if (MODE <= 4) atomicMin(&sm64[a], ((unsigned long long)hash32(it + tid) << 32) | tid);
        else { sm[a] = it; acc += sm[(a + 1) & 4095]; }
    }
    __syncthreads();
    if (acc == 0x12345678) out[0] = acc + sm[tid] + (unsigned)sm64[tid & 2047];
}

template <int MODE>
__global__ void alu_kernel(int iters, double *out) {
    double x = threadIdx.x * 1e-3 + 1.0, y = 1.000001, z = 0.5;
    float fx = threadIdx.x * 1e-3f + 1.0f, fy = 1.000001f, fz = 0.5f;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { x = fma(x, y, z); y = fma(y, 0.999999, 1e-7); z = fma(z, 1.0000001, -1e-9); x = fma(x, 0.99, y); }     // 4 dep-light DFMA
        if (MODE == 1) { fx = fmaf(fx, fy, fz); fy = fmaf(fy, 0.999999f, 1e-7f); fz = fmaf(fz, 1.0000001f, -1e-9f); fx = fmaf(fx, 0.99f, fy); }
        if (MODE == 2) { x = x / y + 1.0; }                                // fp64 division
        if (MODE == 3) { fx = __fdividef(fx, fy) + 1.0f; }
        if (MODE == 4) { x = fabs(x - y) < z ? x + 1.0 : x * 0.5; y += 1e-9; }   // compare/select mix
    }
    if (x + y + z + fx + fy + fz == 123.456) out[0] = x;
}

__global__ void sync_kernel(int iters, int *out) {
    __shared__ int s;
    int acc = 0;
    for (int it = 0; it < iters; it++) { if (threadIdx.x == (it & 31)) s = it; __syncthreads(); acc += s; }
    if (acc == 12345) out[0] = acc;
}

template <class F>
float time_it(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s SMs %d clock %d kHz smem/block optin %zu\n", p.name, sms, clk_khz, p.sharedMemPerBlockOptin);
    unsigned *dout; CK(cudaMalloc(&dout, 1024)); double *ddout; CK(cudaMalloc(&ddout, 1024));
    const int iters = 20000;
    const char *names[] = {"atomicAdd u32 consecutive", "atomicAdd u32 random", "atomicAdd u32 pair-collide", "atomicMin u64 consecutive",
                           "atomicMin u64 pair-collide", "plain STS+LDS"};
    for (int threads : {256, 512}) {
        for (int blocks_per_sm : {1, 2}) {
            int grid = sms * blocks_per_sm;
            float ms[6];
            ms[0] = time_it([&] { atom_kernel<0><<<grid, threads>>>(iters, dout); });
            ms[1] = time_it([&] { atom_kernel<1><<<grid, threads>>>(iters, dout); });
            ms[2] = time_it([&] { atom_kernel<2><<<grid, threads>>>(iters, dout); });
            ms[3] = time_it([&] { atom_kernel<3><<<grid, threads>>>(iters, dout); });
            ms[4] = time_it([&] { atom_kernel<4><<<grid, threads>>>(iters, dout); });
            ms[5] = time_it([&] { atom_kernel<5><<<grid, threads>>>(iters, dout); });
            for (int m = 0; m < 6; m++) {
                double ops = (double)iters * threads * blocks_per_sm;           // per SM
                double cyc = ms[m] * 1e-3 * clk_khz * 1e3;
                printf("threads %d blocks/SM %d  %-28s %8.3f ms  %.2f lane-ops/clk/SM\n", threads, blocks_per_sm, names[m], ms[m], ops / cyc);
            }
        }
    }
    const char *anames[] = {"DFMA x4", "FFMA x4", "fp64 div", "fp32 fast div", "fp64 cmp/select mix"};
    int opsper[] = {4, 4, 1, 1, 1};
    for (int m = 0; m < 5; m++) {
        int threads = 512, grid = sms * 2, it2 = 20000;
        float ms = 0;
        if (m == 0) ms = time_it([&] { alu_kernel<0><<<grid, threads>>>(it2, ddout); });
        if (m == 1) ms = time_it([&] { alu_kernel<1><<<grid, threads>>>(it2, ddout); });
        if (m == 2) ms = time_it([&] { alu_kernel<2><<<grid, threads>>>(it2, ddout); });
        if (m == 3) ms = time_it([&] { alu_kernel<3><<<grid, threads>>>(it2, ddout); });
        if (m == 4) ms = time_it([&] { alu_kernel<4><<<grid, threads>>>(it2, ddout); });
        double ops = (double)it2 * threads * 2 * opsper[m];
        double cyc = ms * 1e-3 * clk_khz * 1e3;
        printf("%-22s %8.3f ms  %.2f lane-ops/clk/SM\n", anames[m], ms, ops / cyc);
    }
    for (int threads : {128, 256, 512, 1024}) {
        int it2 = 20000;
        float ms = time_it([&] { sync_kernel<<<sms, threads>>>(it2, (int *)dout); });
        printf("__syncthreads %4d threads: %.1f cycles each (nominal clock)\n", threads, ms * 1e-3 * clk_khz * 1e3 / it2);
    }
    return 0;
}

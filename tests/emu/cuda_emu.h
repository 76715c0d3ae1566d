// Minimal CPU emulation of the CUDA execution model -- TEST INFRASTRUCTURE ONLY.
//
// Lets the .cu sources of libmpcb200 (rl_mpc_lanemerging_b200/csrc) be compiled by g++ and EXECUTED on the host, so that
// kernel logic written without GPU access can be run against the CPU oracle / model (tests/test_kernel_emulation_cpu.py).
// It is not a product path and not a performance model: one OS thread, every CUDA thread is a fiber (ucontext), blocks run
// one after another, a barrier or warp collective yields to a round-robin scheduler.  Consequences:
//   * deterministic, no data races: atomics are plain operations, a CAS never loses (retry paths are not exercised);
//   * only full-mask warp collectives (0xffffffff) on converged code are supported -- all the library uses;
//   * __shared__ variables are function-local statics (one block at a time); shared-window addresses (the fast kernel's
//     32-bit ld/st/atom.shared operands) are byte offsets from a fixed base inside this module's data segment;
//   * fp64 intrinsics map to plain IEEE operations: compile with -ffp-contract=off.
// What it cannot show: anything about the hardware (inline PTX semantics, memory model, occupancy, timing).
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#ifndef MPC_HOST_EMU
#define MPC_HOST_EMU 1
#endif
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

// ---- vector types / runtime stubs -------------------------------------------------------------------------
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct alignas(16) int4 { int x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
typedef int cudaError_t;
typedef void *cudaStream_t;
static const cudaError_t cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }

// ---- runtime stubs: "device" memory is host memory, streams and events do nothing -------------------------------
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; };
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)calloc(n ? n : 1, 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->multiProcessorCount = 2; p->sharedMemPerBlockOptin = 232448; return 0; }   // B200's opt-in limit, two "SMs"
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (void *)1; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)2; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (void *)1; return 0; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return 0; }
// resident blocks per SM as the shared-memory budget of a B200 SM allows (228 KB, 1 KB reserved per block, ~11 KB static)
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int threads, size_t smem) {
    int by_smem = (int)(233472 / (smem + 11264 + 1024)), by_threads = 2048 / (threads > 0 ? threads : 1);
    *n = by_smem < by_threads ? by_smem : by_threads;
    return 0;
}

namespace emu {
struct Fiber { ucontext_t ctx; char *stack = nullptr; bool done = false; };
struct State {
    emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
    ucontext_t main_ctx;
    std::vector<Fiber> fibers;
    int cur = -1;
    // barriers: generation counters; a waiting fiber yields until its generation moves on
    std::vector<int> warp_arrived, warp_gen, warp_live;
    int block_arrived = 0, block_gen = 0, block_live = 0;
    std::vector<unsigned long long> slot;      // [2][nthreads] exchange buffer of the warp collectives
    std::vector<int> warp_phase;
    std::function<void()> body;
    unsigned char *dyn_smem = nullptr;
    long progress = 0;                          // barrier releases + finished fibers (deadlock watchdog)
    long long nodes = 0;                        // nodes finalised by the fast kernel (MPC_EMU_COUNT_NODE)
    long long cas_issued = 0, cell_reads = 0;   // shared-memory traffic of the fast kernel's label ring (accessor counters)
    long cas_lost = 0;                          // compare-and-swap operations that found another value (only under preemption)
};
inline State &S() { static State s; return s; }
inline void yield() { State &s = S(); swapcontext(&s.fibers[s.cur].ctx, &s.main_ctx); }
inline void trampoline() { State &s = S(); s.body(); s.fibers[s.cur].done = true; swapcontext(&s.fibers[s.cur].ctx, &s.main_ctx); }

// Optional random preemption at shared-memory accesses (emu::set_preempt(seed)): threads then interleave inside a phase, so a
// CAS can lose and the retry / redo paths of the lock-free min-combine are exercised.  Results must not depend on it.
inline unsigned long long &preempt_state() { static unsigned long long x = 0; return x; }
inline void set_preempt(unsigned long long seed) { preempt_state() = seed; }
inline void preempt_point() {
    unsigned long long &x = preempt_state();
    if (!x || S().cur < 0) return;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    if ((x & 7) == 0) { S().progress++; yield(); }
}

inline void warp_barrier() {
    State &s = S();
    const int w = s.cur >> 5, g = s.warp_gen[w];
    if (++s.warp_arrived[w] == s.warp_live[w]) { s.warp_arrived[w] = 0; s.warp_gen[w]++; s.progress++; return; }
    while (s.warp_gen[w] == g) yield();
}
inline void block_barrier() {
    State &s = S();
    const int g = s.block_gen;
    if (++s.block_arrived == s.block_live) { s.block_arrived = 0; s.block_gen++; s.progress++; return; }
    while (s.block_gen == g) yield();
}
// deposit a value, wait for the warp, hand out the buffer of this collective
inline const unsigned long long *exchange(unsigned long long v) {
    State &s = S();
    const int w = s.cur >> 5, n = (int)s.fibers.size();
    const int ph = s.warp_phase[s.cur] & 1;
    s.warp_phase[s.cur]++;
    s.slot[(size_t)ph * n + s.cur] = v;
    warp_barrier();
    return &s.slot[(size_t)ph * n + (size_t)w * 32];
}

// run `body` as grid x block CUDA threads; blocks sequentially
inline void launch(unsigned grid, unsigned block, size_t dyn_smem_bytes, std::function<void()> body) {
    State &s = S();
    alignas(16) static unsigned char smem[232448];       // in the data segment, like every other "shared" object (227 KB opt-in limit)
    if (dyn_smem_bytes > sizeof(smem)) { fprintf(stderr, "cuda_emu: %zu bytes of dynamic shared memory requested\n", dyn_smem_bytes); abort(); }
    memset(smem, 0xcd, sizeof(smem));                    // (uninitialised on a device)
    s.dyn_smem = smem;
    s.body = std::move(body);
    s.gridDim.x = grid; s.blockDim.x = block;
    const size_t STACK = 256 << 10;
    for (unsigned b = 0; b < grid; b++) {
        s.blockIdx.x = b;
        s.fibers.assign(block, Fiber());
        const int nw = (int)(block + 31) / 32;
        s.warp_arrived.assign(nw, 0); s.warp_gen.assign(nw, 0); s.warp_live.assign(nw, 0);
        for (unsigned t = 0; t < block; t++) s.warp_live[t >> 5]++;
        s.block_arrived = 0; s.block_gen = 0; s.block_live = (int)block;
        s.slot.assign(2 * (size_t)block, 0); s.warp_phase.assign(block, 0);
        for (unsigned t = 0; t < block; t++) {
            Fiber &f = s.fibers[t];
            f.stack = (char *)malloc(STACK);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK; f.ctx.uc_link = &s.main_ctx;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        unsigned live = block;
        while (live) {
            const long before = s.progress;
            for (unsigned t = 0; t < block; t++) {
                Fiber &f = s.fibers[t];
                if (f.done) continue;
                s.cur = (int)t; s.threadIdx.x = t;
                swapcontext(&s.main_ctx, &f.ctx);
                if (f.done) {                      // an exited thread no longer takes part in barriers
                    live--; s.progress++;
                    const int w = t >> 5;
                    if (--s.warp_live[w] > 0 && s.warp_arrived[w] == s.warp_live[w]) { s.warp_arrived[w] = 0; s.warp_gen[w]++; }
                    if (--s.block_live > 0 && s.block_arrived == s.block_live) { s.block_arrived = 0; s.block_gen++; }
                }
            }
            if (live && s.progress == before) { fprintf(stderr, "cuda_emu: deadlock (divergent barrier or collective) in block %u\n", b); abort(); }
        }
        for (Fiber &f : s.fibers) free(f.stack);
    }
    s.cur = -1;
}

// shared-window addresses: byte offsets from a base 2 GiB below an anchor in this module's data segment
inline uintptr_t smem_base() { static char anchor; return (uintptr_t)&anchor - (1ull << 31); }
inline unsigned to_shared(const void *p) {
    const uintptr_t d = (uintptr_t)p - smem_base();
    if (d >> 32) { fprintf(stderr, "cuda_emu: shared object outside the 4 GiB window\n"); abort(); }
    return (unsigned)d;
}
template <class T> inline T *from_shared(unsigned a) { return reinterpret_cast<T *>(smem_base() + a); }
}  // namespace emu

#define threadIdx (emu::S().threadIdx)
#define blockIdx (emu::S().blockIdx)
#define blockDim (emu::S().blockDim)
#define gridDim (emu::S().gridDim)

// ---- synchronisation / warp collectives (full mask, converged) -----------------------------------------------
static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline int emu_lane() { return emu::S().cur & 31; }
template <class T> static inline unsigned long long emu_bits(T v) { unsigned long long b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T emu_unbits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_unbits<T>(emu::exchange(emu_bits(v))[src & 31]); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    const unsigned long long *b = emu::exchange(emu_bits(v));
    const int l = emu_lane();
    return l >= (int)d ? emu_unbits<T>(b[l - (int)d]) : v;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    const unsigned long long *b = emu::exchange(emu_bits(v));
    const int l = emu_lane();
    return l + (int)d < 32 ? emu_unbits<T>(b[l + (int)d]) : v;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_unbits<T>(emu::exchange(emu_bits(v))[(emu_lane() ^ m) & 31]); }
static inline unsigned __ballot_sync(unsigned, bool p) {
    const unsigned long long *b = emu::exchange(p ? 1ull : 0ull);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (unsigned)(b[i] & 1ull) << i;
    const emu::State &s = emu::S();                 // lanes that have exited do not vote
    (void)s;
    return r;
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
static inline int __reduce_min_sync(unsigned, int v) { const unsigned long long *b = emu::exchange(emu_bits(v)); int r = INT_MAX; for (int i = 0; i < 32; i++) r = std::min(r, emu_unbits<int>(b[i])); return r; }
static inline int __reduce_max_sync(unsigned, int v) { const unsigned long long *b = emu::exchange(emu_bits(v)); int r = INT_MIN; for (int i = 0; i < 32; i++) r = std::max(r, emu_unbits<int>(b[i])); return r; }

// ---- atomics (single OS thread: plain operations) -----------------------------------------------------------
static inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
static inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline unsigned atomicMin(unsigned *p, unsigned v) { unsigned o = *p; if (v < o) *p = v; return o; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }

// ---- arithmetic intrinsics --------------------------------------------------------------------------------------
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline unsigned __float2uint_rn(float a) { return a <= 0.0f ? 0u : (unsigned)llrintf(a); }
static inline long long __double2ll_rn(double a) { return std::isnan(a) ? (long long)0x8000000000000000ull : llrint(a); }
static inline long long __double_as_longlong(double a) { return (long long)emu_bits(a); }
static inline double __longlong_as_double(long long a) { return emu_unbits<double>((unsigned long long)a); }
static inline double __hiloint2double(int hi, int lo) { return emu_unbits<double>(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (sh & 31)); }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void *p) { return emu::to_shared(p); }
using std::max;
using std::min;

// fast_pull_kernel (MPC_MODE_FAST): the throughput kernel of the fused gap-evaluation.
//
// From layer 2 on all three history points of a node are grid cells, so speed v, acceleration a
// and jerk are small integers (cells per step^n), the successor window is integer arithmetic
// (plus an exact fp64 check where the reference's speed clamps sit on an integer cell boundary),
// and the kinematic edge cost is a function of three small integers: it is tabulated once per
// Settings snapshot in 2^-18 fixed point.  Layers 0 and 1 carry the off-grid start history
// (st_cy.pyx:329-330) and are handled in exact fp64, then quantised.
//
// Integer DP: a cell's state is ONE 64-bit word  [label : 48 bits fixed point][255 - v : 8][a + 128 : 8].
// One pass over the layers; every cell k of layer t is handled by one thread:
//   1. read + clear its word (the buffer is the destination buffer of layer t+2);
//   2. distance penalty through the layer's sorted search structure (bucket table + sentinel edges), label += penalty;
//   3. integer successor window [w, w+n); for each successor k' the candidate word
//      (label + V[v'] + A[a'] + J[j'], v', a') is min-combined into the next layer's buffer: plain pre-read, CAS only
//      where the candidate beats what was read.
// The unsigned 64-bit order of the word IS the reference's heap order: smaller label first, ties to the
// larger v' = smaller predecessor index (st_cy.pyx:388).  Integer addition is exact and associative, so the
// result is independent of thread scheduling and is reproduced bit for bit by the CPU model
// orc_solve_fast_model (oracle/mpc_oracle.c).  Quantisation (<= 1.9e-6 per edge) keeps the cost within
// ~1e-7 rel of the reference; fp32 labels were tried first and rejected.
//
// Two main loops share the prologue (layers 0..2) and the epilogue (back-track, crash test):
//   * the LEAN bounded pass (descriptor-fed problems, first attempt): cells in a band or a penalty zone are blocked through a
//     bit array built one layer ahead, one barrier per layer, software-pipelined branch-free min-combine;
//   * the general loop (dense grids, no bound / retry after a bounded attempt that did not reach the horizon): obstacle bands
//     clipped at push time, exact fp64 threshold test of the penalty per node, best label tracked per layer.
// Design history, ablations and what bounds the kernel: profiles/README.md.
#include "mpc_solve_common.cuh"
#ifndef MPC_ABLATE
#define MPC_ABLATE 0      // dev-only timing ablations (wrong results when != 0)
#endif

#define EMPTY_LAB 0x7ff0000000000000ULL      // +inf
#ifndef MPC_EMU_COUNT_NODE                    // tests/emu counts the nodes each pass finalises (compared with the CPU model)
#ifdef MPC_HOST_EMU
#define MPC_EMU_COUNT_NODE() (emu::S().nodes++)
#else
#define MPC_EMU_COUNT_NODE()
#endif
#endif

struct __align__(16) FastShared {
    LayerSearch layer[4];        // first: staged with 16-byte vector stores; layer t lives in slot t & 3
    BlockShared S;
};
static_assert(sizeof(LayerSearch) % 16 == 0, "LayerSearch buffers must stay 16-byte aligned");

// cell providers for the fast kernel ------------------------------------------------------------------
struct FastDescProv {
    const LayerDesc *base;      // desc + b*num_t
    LayerSearch *sm;            // four staging buffers in shared memory
    int4 pre;                   // prefetch register
    // a LayerSearch is a contiguous tail of LayerDesc behind its 32-byte header (counts, lower sentinel edge)
    static constexpr int kHead = (int)offsetof(LayerSearch, edge);
    static constexpr int kTail = (int)(sizeof(LayerSearch) - kHead) / 16;
    __device__ __forceinline__ void load(int t) {          // issue the global loads of layer t (no wait)
        const LayerDesc *src = base + t;
        if (threadIdx.x < kTail) pre = reinterpret_cast<const int4 *>(reinterpret_cast<const char *>(src) + offsetof(LayerDesc, edge))[threadIdx.x];
        else if (threadIdx.x == kTail) pre = make_int4(src->n_edge, src->n_band, src->n_blk, 0);
    }
    __device__ __forceinline__ void store(int t) {         // park them in slot t & 3
        LayerSearch *dst = sm + (t & 3);
        if (threadIdx.x < kTail) reinterpret_cast<int4 *>(reinterpret_cast<char *>(dst) + kHead)[threadIdx.x] = pre;
        else if (threadIdx.x == kTail) { *reinterpret_cast<int4 *>(dst) = pre; dst->edge_lo = -1e300; }
    }
    static constexpr bool kClipAtPush = true;               // successors are tested against the next layer's bands before the push
    static constexpr bool kDense = false;
    __device__ __forceinline__ double eval_staged(int t, int k, double s, bool &ob) const { return cell_distance_sorted(sm[t & 3], s, k, ob); }
    // distance only (the caller knows the cell is not in a band)
    __device__ __forceinline__ double distance_staged(int t, int k, double s) const {
        const LayerSearch &L = sm[t & 3];
        int e = L.bucket_edge[k >> MPC_BUCKET_SHIFT], M = L.n_edge;
        while (e < M && L.edge[e] < s) e++;
        double d = 1E10;
        if (e > 0) { double x = __dsub_rn(s, L.edge[e - 1]); d = x < d ? x : d; }
        if (e < M) { double x = fabs(__dsub_rn(s, L.edge[e])); d = x < d ? x : d; }
        return d;
    }
    // the (at most two) merged bands of layer t that can intersect the short cell window starting at k
    __device__ __forceinline__ void bands_near(int t, int k, int2 &b0, int2 &b1) const {
        const LayerSearch &L = sm[t & 3];
        int i = L.bucket_band[k >> MPC_BUCKET_SHIFT], m = L.n_band;
        while (i < m && L.mband[i].y <= k) i++;
        b0 = i < m ? L.mband[i] : make_int2(INT_MAX, INT_MAX);
        b1 = i + 1 < m ? L.mband[i + 1] : make_int2(INT_MAX, INT_MAX);
    }
    __device__ __forceinline__ bool is_obstacle(int t, int k, int j) const {
        const LayerSearch &L = sm[t & 3];
        int i = L.bucket_band[j], m = L.n_band;
        while (i < m && L.mband[i].y <= k) i++;
        return i < m && L.mband[i].x <= k;
    }
    // in a band or inside a penalty zone (LayerDesc::blk); a handful of calls per problem
    __device__ __forceinline__ bool is_blocked(int t, int k) const {
        const LayerSearch &L = sm[t & 3];
        bool b = false;
        for (int i = 0; i < L.n_blk; i++) b = b || (k >= L.blk[i].x && k < L.blk[i].y);
        return b;
    }
    __device__ __forceinline__ double eval_global(int t, int k, double s, bool &ob) const { return cell_distance(base[t], s, k, ob); }
};
static_assert(offsetof(LayerDesc, edge) % 16 == 0 && sizeof(LayerSearch) % 16 == 0, "LayerSearch must be int4-copyable");
static_assert(offsetof(LayerDesc, bucket_band) - offsetof(LayerDesc, edge) == offsetof(LayerSearch, bucket_band) - offsetof(LayerSearch, edge), "layout mismatch");
static_assert(offsetof(LayerDesc, blk) - offsetof(LayerDesc, edge) == offsetof(LayerSearch, blk) - offsetof(LayerSearch, edge), "layout mismatch");
static_assert(offsetof(LayerSearch, edge) % 16 == 0 && offsetof(LayerSearch, edge_lo) + 8 == offsetof(LayerSearch, edge), "edge[-1] must be the sentinel");

template <typename DT>
struct FastDenseProv {
    const uint8_t *ob_base; const DT *d_base; int stride;
    __device__ __forceinline__ void load(int) {}
    __device__ __forceinline__ void store(int) {}
    __device__ __forceinline__ double eval_staged(int t, int k, double s, bool &ob) const {
        size_t o = (size_t)t * stride + k;
        ob = ob_base[o] != 0;
        return ob ? 0.0 : (double)d_base[o];
    }
    __device__ __forceinline__ double eval_global(int t, int k, double s, bool &ob) const {
        size_t o = (size_t)t * stride + k;
        ob = ob_base[o] != 0;
        return (double)d_base[o];
    }
    static constexpr bool kClipAtPush = false;              // dense grids: the obstacle test stays at the destination
    static constexpr bool kDense = true;
    // (fast32_kernel) in a band, or inside a penalty zone: the cells the bounded pass never enters
    __device__ __forceinline__ bool is_blocked_zone(int t, int k, double min_allowed) const {
        const size_t o = (size_t)t * stride + k;
        return ob_base[o] != 0 || (double)d_base[o] < min_allowed;
    }
    __device__ __forceinline__ double distance_at(int t, int k) const { return (double)d_base[(size_t)t * stride + k]; }
    // the same test for the four cells k4 .. k4+3 (k4 % 4 == 0) with one vector load per array; bit j = cell k4 + j.  Needs vec_ok().
    __device__ __forceinline__ bool vec_ok() const {
        return (stride & 3) == 0 && ((uintptr_t)ob_base & 3) == 0 && ((uintptr_t)d_base & 15) == 0;
    }
    __device__ __forceinline__ unsigned blocked4(int t, int k4, double min_allowed, int num_s) const {
        if (k4 >= stride) return 0xfu;
        const size_t o = (size_t)t * stride + k4;
        const uchar4 m = *reinterpret_cast<const uchar4 *>(ob_base + o);
        double d0, d1, d2, d3;
        if constexpr (sizeof(DT) == 4) {
            const float4 d = *reinterpret_cast<const float4 *>(d_base + o);
            d0 = (double)d.x; d1 = (double)d.y; d2 = (double)d.z; d3 = (double)d.w;
        } else {
            const double2 a = *reinterpret_cast<const double2 *>(d_base + o), b = *reinterpret_cast<const double2 *>(d_base + o + 2);
            d0 = a.x; d1 = a.y; d2 = b.x; d3 = b.y;
        }
        unsigned r = 0;
        r |= (k4 + 0 >= num_s || m.x != 0 || d0 < min_allowed) ? 1u : 0u;
        r |= (k4 + 1 >= num_s || m.y != 0 || d1 < min_allowed) ? 2u : 0u;
        r |= (k4 + 2 >= num_s || m.z != 0 || d2 < min_allowed) ? 4u : 0u;
        r |= (k4 + 3 >= num_s || m.w != 0 || d3 < min_allowed) ? 8u : 0u;
        return r;
    }
    // Bulk L2 prefetch (cp.async.bulk.prefetch.L2, one instruction per array) of the mask bytes and the distances of cells
    // [klo, khi] of layer t: issued by ONE thread a layer before build_blocked_bits_dense reads them, so that those reads -- the
    // first touch of every byte of the grid the DP uses -- find the lines in L2 instead of waiting for HBM.
    __device__ __forceinline__ void prefetch_span(int t, int klo, int khi) const {
#ifndef MPC_HOST_EMU
        if (khi < klo) return;
        const size_t o = (size_t)t * stride;
        const uintptr_t a0 = (uintptr_t)(ob_base + o + klo) & ~(uintptr_t)15, a1 = ((uintptr_t)(ob_base + o + khi + 1) + 15) & ~(uintptr_t)15;
        const uintptr_t b0 = (uintptr_t)(d_base + o + klo) & ~(uintptr_t)15, b1 = ((uintptr_t)(d_base + o + khi + 1) + 15) & ~(uintptr_t)15;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0), "r"((unsigned)(a1 - a0)) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(b0), "r"((unsigned)(b1 - b0)) : "memory");
#endif
    }
    __device__ __forceinline__ double distance_staged(int t, int k, double s) const { bool ob; return eval_staged(t, k, s, ob); }
    __device__ __forceinline__ void bands_near(int, int, int2 &b0, int2 &b1) const { b0 = make_int2(INT_MAX, INT_MAX); b1 = b0; }
    __device__ __forceinline__ bool is_obstacle(int t, int k, int) const { return ob_base[(size_t)t * stride + k] != 0; }
    __device__ __forceinline__ bool is_blocked(int t, int k) const { return is_obstacle(t, k, 0); }
};

#define FX_EMPTY 0xffffffffffffffffULL
#define FX_LABEL_LIMIT ((1ULL << 48) - (1ULL << 36))   // labels above this (1.07e9) hand the problem to the exact kernel

// The two places where the reference's successor window depends on fp64 rounding (SURVEY.md §7 "hard part 1"):
//   lower clamp at speed 0:   first index = ceil((s_k - s0)/ds)            in {k, k+1}
//   upper clamp at MAX_SPEED: last index  = int((s_k + v_max*dt - s0)/ds)  in {k+vmax_c-1, k+vmax_c}   (when v_max*dt/ds is an integer)
// Both are functions of the cell index only, so they are evaluated once per problem for every cell (exact fp64, the
// reference's expression) and kept as two bit arrays; the per-node window is then pure integer arithmetic.
struct ClampBits { unsigned *lo, *hi; };      // (num_s_max + 31) / 32 words each, in dynamic shared memory

__device__ __forceinline__ void build_clamp_bits(const DevParams &P, const SGrid &g, const ClampBits &C) {
    const int words = (g.num_s + 31) >> 5, lane = threadIdx.x & 31;
    for (int wd = threadIdx.x >> 5; wd < words; wd += blockDim.x >> 5) {           // one warp per 32 cells, one cell per lane
        const int k = (wd << 5) + lane;
        const double s = g.sval(k);
        const double me = __ddiv_rn(__dsub_rn(s, g.s0), g.ds);
        int mi = (int)me; if ((double)mi < me) mi += 1;
        bool hi_short = false;
        if (P.vmax_is_int) hi_short = ((int)__ddiv_rn(__dsub_rn(__dadd_rn(s, __dmul_rn(P.p.max_speed, P.p.t_disc)), g.s0), g.ds) - k) != P.vmax_c;
        const unsigned lo = __ballot_sync(0xffffffffu, mi != k);                    // own cell excluded: window starts at k+1
        const unsigned hi = __ballot_sync(0xffffffffu, hi_short);                   // last index one short of k + vmax_c
        if (lane == 0) { C.lo[wd] = lo; C.hi[wd] = hi; }
    }
}

// integer successor window of a node (k, v, a): cells [wlo, wlo+n).  Mirrors st_cy.pyx:65-93 for on-grid history.
__device__ __forceinline__ void int_window(const DevParams &P, const SGrid &g, const ClampBits &C, int k, int v, int a, int &wlo, int &n) {
    int alo = max(a + P.jlo_c, P.alo_c), ahi = min(a + P.jhi_c, P.ahi_c);
    int vlo = v + alo, vhi = v + ahi;
    if (vlo <= 0) vlo = (C.lo[k >> 5] >> (k & 31)) & 1;
    bool clamp_hi = P.vmax_is_int ? (vhi >= P.vmax_c) : ((double)v + fmin((double)a + P.jhi_r, P.ahi_r) > P.vmax_r);
    if (clamp_hi) vhi = P.vmax_c - (P.vmax_is_int ? (int)((C.hi[k >> 5] >> (k & 31)) & 1) : 0);
    wlo = k + vlo;
    int whi = min(k + vhi, g.num_s - 1);
    n = whi - wlo + 1; n = n < 0 ? 0 : n;
}

// Shared-memory words are addressed by their 32-bit shared-window address and accessed with explicit
// ld/st/atom.shared: through a generic pointer that nvcc cannot prove to be shared (the two label buffers
// are selected by layer parity) it emits generic LD.E / ATOM.E.CAS plus the software fall-back of generic
// atomics on shared memory, several times the cost of LDS / ATOMS.CAS (seen in the SASS of the v8 kernel).
#ifdef MPC_HOST_EMU     // tests/emu: the same accessors on the emulated shared window (plain memory operations)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return emu::to_shared(p); }
__device__ __forceinline__ unsigned long long lds_u64(unsigned a) { emu::preempt_point(); emu::S().cell_reads++; return *emu::from_shared<unsigned long long>(a); }
__device__ __forceinline__ void sts_u64(unsigned a, unsigned long long v) { emu::preempt_point(); *emu::from_shared<unsigned long long>(a) = v; }
__device__ __forceinline__ unsigned long long atoms_cas_u64(unsigned a, unsigned long long cmp, unsigned long long val) {
    emu::preempt_point();
    emu::S().cas_issued++;
    unsigned long long *p = emu::from_shared<unsigned long long>(a), old = *p; if (old == cmp) *p = val; else emu::S().cas_lost++; return old;
}
__device__ __forceinline__ unsigned long long lds_u64_nc(unsigned a) { return lds_u64(a); }
__device__ __forceinline__ unsigned lds_u8_nc(unsigned a) { return *emu::from_shared<unsigned char>(a); }
__device__ __forceinline__ double lds_f64_nc(unsigned a) { return *emu::from_shared<double>(a); }
__device__ __forceinline__ unsigned long long atoms_cas_u64_if(unsigned a, unsigned long long cmp, unsigned long long val, bool doit) {
    return doit ? atoms_cas_u64(a, cmp, val) : cmp;
}
__device__ __forceinline__ unsigned lds_u32_nc(unsigned a) { return *emu::from_shared<unsigned>(a); }
__device__ __forceinline__ unsigned long long atoms_cas_u64_nc(unsigned a, unsigned long long cmp, unsigned long long val) { return atoms_cas_u64(a, cmp, val); }
#else
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long lds_u64(unsigned a) {
    unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void sts_u64(unsigned a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long atoms_cas_u64(unsigned a, unsigned long long cmp, unsigned long long val) {
    unsigned long long old; asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(a), "l"(cmp), "l"(val) : "memory"); return old;
}

// same, without the compiler-level memory barrier: several of them may be scheduled back to back (the caller orders them
// against the other accesses of the same words with the barrier variants / __syncthreads)
__device__ __forceinline__ unsigned long long lds_u64_nc(unsigned a) {
    unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v;
}
__device__ __forceinline__ unsigned lds_u8_nc(unsigned a) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds_f64_nc(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
// predicated CAS: returns `cmp` when `doit` is false (no branch around the atomic)
__device__ __forceinline__ unsigned long long atoms_cas_u64_if(unsigned a, unsigned long long cmp, unsigned long long val, bool doit) {
    unsigned long long old;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\tmov.b64 %0, %2;\n\t@p atom.shared.cas.b64 %0, [%1], %2, %3;\n\t}"
                 : "=&l"(old) : "r"(a), "l"(cmp), "l"(val), "r"((unsigned)doit));
    return old;
}
__device__ __forceinline__ unsigned lds_u32_nc(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned long long atoms_cas_u64_nc(unsigned a, unsigned long long cmp, unsigned long long val) {
    unsigned long long old; asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(a), "l"(cmp), "l"(val)); return old;
}
#endif

// int_window with the clamp bit arrays addressed in the shared window (lo at cba, hi at cba + 4 * NW)
__device__ __forceinline__ void int_window_sa(const DevParams &P, int num_s, unsigned cba, int NW, int k, int v, int a, int &wlo, int &n) {
    int alo = max(a + P.jlo_c, P.alo_c), ahi = min(a + P.jhi_c, P.ahi_c);
    int vlo = v + alo, vhi = v + ahi;
    if (vlo <= 0) vlo = (lds_u32_nc(cba + 4u * (unsigned)(k >> 5)) >> (k & 31)) & 1;
    bool clamp_hi = P.vmax_is_int ? (vhi >= P.vmax_c) : ((double)v + fmin((double)a + P.jhi_r, P.ahi_r) > P.vmax_r);
    if (clamp_hi) vhi = P.vmax_c - (P.vmax_is_int ? (int)((lds_u32_nc(cba + 4u * (unsigned)(NW + (k >> 5))) >> (k & 31)) & 1) : 0);
    wlo = k + vlo;
    int whi = min(k + vhi, num_s - 1);
    n = whi - wlo + 1; n = n < 0 ? 0 : n;
}

// min-combine into shared memory; the CAS is only issued when the candidate beats the stored word.
__device__ __forceinline__ void smem_min64(unsigned addr, unsigned long long val) {
    unsigned long long old = lds_u64(addr);
    while (val < old) {
        unsigned long long assumed = old;
        old = atoms_cas_u64(addr, assumed, val);
        if (old == assumed) return;
    }
}

__device__ __forceinline__ unsigned long long fx_from_double(double x) { return (unsigned long long)__double2ll_rn(__dmul_rn(x, MPC_FX_ONE)); }

// distance penalty d_w * pen(d) in label units (st_cy.pyx:34-38,50); the threshold test is exact fp64.  Outside the
// penalty zone d_w / d is evaluated in fp32 -- rint(kw * (1.0f / (float)d)), kw = (float)(d_w * 2^18) -- (relative error
// 1e-7 of a term <= d_w / m); the CPU model orc_solve_fast_model does the same three IEEE operations.
__device__ __forceinline__ unsigned fx_inv_penalty(float kw, double d) { return __float2uint_rn(__fmul_rn(kw, __frcp_rn((float)d))); }
__device__ __forceinline__ unsigned long long fx_penalty(const DevParams &P, double d) {
    if (d < P.p.min_allowed_distance) return fx_from_double(__dmul_rn(P.p.d_weight, __ddiv_rn(1000000.0, d > 1.0 ? d : 1.0)));
    return (unsigned long long)fx_inv_penalty(P.kw, d);
}

// Lean bounded pass: bit w*32+j of the array = cell is blocked in the layer (LayerDesc::blk).  Only the words that
// the layer's nodes can occupy, [klo, khi], are rebuilt.
__device__ __forceinline__ void build_blocked_bits(const LayerSearch &L, unsigned *bits, int klo, int khi, int tid, int nth) {
    const int w1 = khi >> 5, nb = L.n_blk;
    for (int w = (klo >> 5) + tid; w <= w1; w += nth) {
        const int c0 = w << 5;
        unsigned m = 0;
        for (int i = 0; i < nb; i++) {
            const int2 b = L.blk[i];
            if (b.x >= c0 + 32) break;
            const int lo = max(b.x - c0, 0), hi = min(b.y - c0, 32);
            if (hi > lo) m |= (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
        }
        bits[w] = m;
    }
}

// Cost bound.  Labels only grow along a path (every term of st_cy.pyx:46-50 is >= 0) and a cell keeps the arrival with
// the smallest label, so dropping every node whose label exceeds a bound U leaves all nodes with label <= U -- label,
// speed/acceleration code and back-pointer -- exactly as the unbounded pass computes them (induction over the layers:
// the winner of such a node has a smaller label and survives too).  If the bounded pass reaches the horizon its answer
// IS the unbounded answer; if it does not, the same block solves the problem again without the bound.  The first pass uses
// U = d_weight * 1e6 / min_allowed_distance (DevParams::bound_fx): the cheapest possible label of a path that spends one
// step inside a penalty zone (st_cy.pyx:34-38).  Ordinary plans never do, and the sub-trees behind those zones are
// ~30 % of all nodes at H=50 (oracle model counts).
struct FxTables { unsigned v[256], aj[32 * 16]; };     // aj[(a'+16)*16 + (j'+8)] = A[a'] + J[j']: one lookup, stride 17 along a window

template <class Prov, bool DESC, bool WRAP, int MAXT, bool HINT>
__global__ void __launch_bounds__(MAXT, (MAXT <= 192 ? 5 : MAXT <= 256 ? 4 : MAXT <= 384 ? 3 : MAXT <= 512 ? 2 : 1)) fast_pull_kernel(DevParams P, int B, SolveIO io, const LayerDesc *desc,
                                                                              const uint8_t *dense_ob, const void *dense_d, int dense_stride, int Wc,
                                                                              unsigned long long bound) {
#ifdef MPC_HOST_EMU
    unsigned char *const smem_raw = emu::S().dyn_smem;
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    __shared__ FastShared FS;
    __shared__ FxTables TB;
    __shared__ unsigned long long s_layer_best[2];
    __shared__ int s_chunk[3];          // dense traversal: next 32-cell chunk of a pass
    BlockShared &S = FS.S;
    // two label buffers of Wc 64-bit words (layer parity), addressed in the shared window
    const unsigned sb0 = smem_u32(smem_raw), sb1 = sb0 + 8u * (unsigned)Wc;
    ClampBits CB;
    const int NW = (P.num_s_max + 31) >> 5;
    CB.lo = reinterpret_cast<unsigned *>(smem_raw + (size_t)16 * Wc); CB.hi = CB.lo + NW;
    unsigned *const blkbits[2] = {CB.hi + NW, CB.hi + 2 * NW + 2};       // lean pass: blocked cells of layer t in blkbits[t & 1] (NW + 2 words each)
    uint16_t *bp = io.bp + (size_t)blockIdx.x * P.num_t * io.bp_stride;
    const int T = P.num_t, tid = threadIdx.x, nth = blockDim.x;
    auto ring = [Wc](int k) -> int { return WRAP ? (k >= Wc ? k - Wc : k) : k; };
    for (int i = tid; i < 256; i += nth) TB.v[i] = P.vtab[i];
    for (int i = tid; i < 512; i += nth) TB.aj[i] = P.atab[i >> 4] + P.jtab[i & 15];
    for (int k = tid; k < 2 * Wc; k += nth) sts_u64(sb0 + 8u * k, FX_EMPTY);       // both buffers; every pass leaves them empty again
    if (io.B_dev) B = *io.B_dev;
    if (io.n_hi != 0 && (B < io.n_lo || B > io.n_hi)) return;         // (uniform) the other launch shape has this list
    for (;;) {
        if (tid == 0) S.b = atomicAdd(io.work_counter, 1);
        __syncthreads();
        int wi = S.b;
        if (wi >= B) break;
        int b = io.subset ? io.subset[wi] : wi;
        // bit 30 of a list entry (set by fast32_kernel): the bounded attempt has already failed for this problem
        if (io.only_flagged && !((b >> 30) & 1)) { __syncthreads(); continue; }      // (with fast32_kernel's second shape, running concurrently)
        const unsigned long long bound_b = ((b >> 30) & 1) ? FX_EMPTY : bound;
        b &= 0x3fffffff;
        SGrid g;
        double v0, a0;
        if (DESC) { g = make_sgrid(P, io.ego[4 * b], io.ego[4 * b + 1]); v0 = io.ego[4 * b + 2]; a0 = io.ego[4 * b + 3]; }
        else { g.s0 = io.s0[b]; g.ds = io.ds[b]; g.num_s = io.num_s[b]; v0 = io.v0[b]; a0 = io.a0[b]; }
        Prov prov;
        if constexpr (DESC) { prov.base = desc + (size_t)b * T; prov.sm = FS.layer; }
        else { prov.ob_base = dense_ob + (size_t)b * T * dense_stride;
               prov.d_base = reinterpret_cast<decltype(prov.d_base)>(dense_d) + (size_t)b * T * dense_stride;
               prov.stride = dense_stride; }
        build_clamp_bits(P, g, CB);
        const double est_prev = __dsub_rn(g.s0, __dmul_rn(v0, P.p.t_disc));
        const double est_second = __dsub_rn(est_prev, __dmul_rn(P.p.t_disc, __dsub_rn(v0, __dmul_rn(a0, P.p.t_disc))));
        int bt = 0; unsigned long long best_word = 0ULL;      // deepest non-empty layer and its (label << 16 | k)
        int dlo = 0, dhi = -1;
        // first attempt under the cost bound (and with its penalty zones closed), second attempt without -- see the note above
        unsigned long long bnd = bound_b;
        int zone = (Prov::kClipAtPush && bound_b != FX_EMPTY) ? P.zone_cells : 0;
        // lean first attempt (descriptor-fed, bounded): see the note above the kernel
        bool lean = DESC && Prov::kClipAtPush && bound_b != FX_EMPTY && P.zone_ok;
        // Per-problem cost hint (mpc_plan_hinted: a coarse probe plan, or the previous tick's plan in closed loop): an
        // ESTIMATE of the plan's cost, used as a tighter first bound.  Any bound is exact when the pass reaches the
        // horizon (note above the kernel), so a hint can only cost time: too low -> the attempt is repeated under the
        // standard bound; at or above the standard bound (the plan is expected to cross a penalty zone) -> the
        // zone-closed attempts, which could not succeed, are skipped.
        // (HINT is a template parameter so that the code of the un-hinted kernel does not change.)
        if (HINT && DESC && bound_b != FX_EMPTY && (io.hint_reached == nullptr || io.hint_reached[b] == io.hint_full_t)) {
            const double hc = __dmul_rn(io.hint_cost[b], io.hint_scale);
            if (hc > 0.0 && hc < 1.0e9) {                      // (a NaN fails both tests)
                bnd = fx_from_double(hc);
                if (bnd >= bound_b) { zone = 0; lean = false; }
            }
        }
      for (;;) {
        prov.load(1);
        bt = 0; best_word = 0ULL; dlo = 0; dhi = -1;
        if (tid == 0) { S.need_fallback = 0; S.bound_hit = 0; for (int i = 0; i < 3; i++) { S.nlo[i] = INT_MAX; S.nhi[i] = -1; } s_layer_best[0] = FX_EMPTY; s_layer_best[1] = FX_EMPTY; s_chunk[0] = 0; s_chunk[1] = 0; s_chunk[2] = 0; }
        // ---- prologue: layers 0 -> 1 -> 2 have off-grid history (st_cy.pyx:329-330): exact fp64, then quantised ----
        int imin0, imax0;
        exact_window(P, g.s0, g.ds, g.s0, est_prev, est_second, imin0, imax0);
        if (imax0 > g.num_s) imax0 = g.num_s;
        prov.store(1);
        prov.load(2);
        __syncthreads();
        // layer 1 (buffer 1): full label = kinematics + penalty; v = k, a unused
        if (tid < imax0 - imin0) {
            int kk = imin0 + tid;
            double sn = g.sval(kk);
            bool ob; double d = prov.eval_staged(1, kk, sn, ob);
            if (!ob && kk >= 256) S.need_fallback = 1;
            else if (!ob) {
                unsigned long long l1 = fx_from_double(exact_cost(P, sn, g.s0, est_prev, est_second, d));
                sts_u64(sb1 + 8u * ring(kk), (l1 << 16) | ((unsigned long long)(255 - kk) << 8) | 128ULL);
                bp[(size_t)1 * io.bp_stride + kk] = 0;
                atomicMin(&s_layer_best[1], (l1 << 16) | (unsigned long long)kk);
                atomicMin(&S.nlo[1], kk); atomicMax(&S.nhi[1], kk);
            }
        }
        prov.store(2);
        if (T > 3) prov.load(3);
        __syncthreads();
        if (T > 3) prov.store(3);
        if (T > 4) prov.load(4);
        bool done = false;
        if (S.nhi[1] < 0) done = true;                        // nothing reachable at layer 1: best node is (0,0)
        else {
            const int lo1 = S.nlo[1], hi1 = S.nhi[1];
            bt = 1; best_word = s_layer_best[1];
            // layer 1 -> 2 edges with the exact window / exact kinematic cost; one thread per (node, successor)
            const int lme = P.lmax_exact;
            int mylo = INT_MAX, myhi = -1;
            for (int e = tid; e < (hi1 - lo1 + 1) * lme; e += nth) {
                int k1 = lo1 + e / lme, j = e % lme;
                unsigned long long w1 = lds_u64(sb1 + 8u * ring(k1));
                if (w1 == FX_EMPTY) continue;
                double s = g.sval(k1);
                int imin, imax;
                exact_window(P, g.s0, g.ds, s, g.s0, est_prev, imin, imax);
                int kk = imin + j;
                if (kk >= imax || kk >= g.num_s) continue;
                if (lean ? prov.is_blocked(2, kk) : prov.is_obstacle(2, kk, kk >> MPC_BUCKET_SHIFT)) {               // st_cy.pyx:383-384
                    if (lean) S.bound_hit = 1;                // (a zone cell: if the plan ends here it is repeated without the bound)
                    continue;
                }
                int vn = kk - k1, an = vn - k1;
                if (vn > 255 || an < -16 || an > 15) { S.need_fallback = 1; continue; }
                unsigned long long tot = (w1 >> 16) + fx_from_double(exact_kin(P, g.sval(kk), s, g.s0, est_prev));
                smem_min64(sb0 + 8u * ring(kk), (tot << 16) | ((unsigned long long)(255 - vn) << 8) | (unsigned long long)(an + 128));
                mylo = min(mylo, kk); myhi = max(myhi, kk);
            }
            mylo = warp_min_i(mylo); myhi = warp_max_i(myhi);
            if ((tid & 31) == 0 && myhi >= 0) { atomicMin(&S.nlo[2], mylo); atomicMax(&S.nhi[2], myhi); }
            __syncthreads();
            for (int k1 = lo1 + tid; k1 <= hi1; k1 += nth) sts_u64(sb1 + 8u * ring(k1), FX_EMPTY);     // layer 1 is consumed
            dlo = S.nlo[2]; dhi = S.nhi[2];
            if (dhi < 0) done = true;                         // layer-1 nodes have no successors
        }
        if (T > 4) prov.store(4);
        const int lane = tid & 31;
        if (lean && !done) {
            if (T > 5) prov.load(5);
            // ---- lean bounded pass.  Every cell that is inside an obstacle band or a penalty zone of layer t is marked in a bit
            // array one iteration ahead; successors are tested against it at push time, so every node that is finalised has
            // d >= MIN_ALLOWED_DISTANCE: its penalty is the 1/d branch, and (being the complement of what the bound drops) the
            // surviving nodes are exactly those of the unbounded pass.  One barrier per layer; the layer's best label is only
            // tracked at the horizon -- a pass that does not get there is repeated without the bound. ----
            if (T > 3) build_blocked_bits(FS.layer[3], blkbits[1], dlo, min(dhi + P.vmax_c, g.num_s - 1), tid, nth);
            if (tid == 0) { s_layer_best[0] = FX_EMPTY; s_layer_best[1] = FX_EMPTY; }      // (layer 1's best is in registers by now)
            const float kw = P.kw;
            const unsigned tbv = smem_u32(TB.v), tbaj = smem_u32(TB.aj), cba = smem_u32(CB.lo);
            for (int t = 2; t < T; t++) {
                const int par = t & 1, s3 = t % 3, n3 = (t + 1) % 3;
                const unsigned cur = par ? sb1 : sb0, nxt = par ? sb0 : sb1;
                __syncthreads();                              // pushes into layer t, its span, staging of layer t+2 and bits of t+1 are complete
                dlo = S.nlo[s3]; dhi = S.nhi[s3];
                if (dhi < 0) break;                           // no successors
                if (WRAP && dhi - dlo + 1 + (t == T - 1 ? 0 : P.vmax_c + 8) > Wc) { if (tid == 0) S.need_fallback = 1; break; }   // frontier (and, but for the last layer, its successors) may outgrow the ring
                if (tid == 0) { S.nlo[(t + 2) % 3] = INT_MAX; S.nhi[(t + 2) % 3] = -1; s_chunk[n3] = 0; }    // what iteration t+1 accumulates into
                if (t + 3 < T) prov.store(t + 3);             // loaded during iteration t-1; first read in iteration t+1
                if (t + 4 < T) prov.load(t + 4);
                const bool last = (t == T - 1);
                // hinted solves: per-bucket reachability caps of layer t (mpc_reach.cu), nrem = steps to the horizon
                const unsigned short *caprow = (HINT && io.capb != nullptr) ? io.capb + ((size_t)b * T + t) * io.cap_stride : nullptr;
                const int nrem = T - 1 - t, dmax = P.vstar_c * nrem;
                const float rcpn = 1.0f / (float)(nrem > 0 ? nrem : 1);
                if (t + 2 < T) build_blocked_bits(FS.layer[(t + 2) & 3], blkbits[par], dlo, min(dhi + 2 * P.vmax_c, g.num_s - 1), tid, nth);
                const unsigned edge0 = smem_u32(FS.layer[t & 3].edge), bucket0 = smem_u32(FS.layer[t & 3].bucket_edge);
                const unsigned bwa = smem_u32(blkbits[0]) + (par ? 0u : 4u * (unsigned)(NW + 2));      // blocked bits of layer t+1
                uint16_t *bp_row = bp + (size_t)t * io.bp_stride;
#ifndef MPC_HOST_EMU
                asm volatile("" : "+l"(bp_row));              // keep the row pointer in registers (else it is rebuilt per node)
#endif
                unsigned long long mybest = FX_EMPTY;
                int mylo = INT_MAX, myhi = -1;
                // 32-cell chunks of the layer's span: the first one per warp is static, the rest come from a shared counter
                const int nwarp = nth >> 5;
                int c = tid >> 5;
                for (;;) {
                    const int base = dlo + (c << 5);
                    if (base > dhi) break;
                    int cn = 0;                               // next chunk: fetched before this one is processed
                    if (lane == 0) cn = nwarp + atomicAdd(&s_chunk[s3], 1);
                    const int k = base + lane, rk = ring(k);
                    unsigned long long w = FX_EMPTY;
                    if (k <= dhi) w = lds_u64(cur + 8u * rk);
                    if (w != FX_EMPTY) {
                        sts_u64(cur + 8u * rk, FX_EMPTY);     // this buffer receives layer t+2
                        unsigned pen = 0;
                        if (MPC_ABLATE != 3) {                // nearest distance-field edge on either side; edge[-1] / edge[n_edge] are -/+1e300
                            const double sv = g.sval(k);
                            unsigned ea = edge0 + 8u * lds_u8_nc(bucket0 + (k >> MPC_BUCKET_SHIFT));
                            double hi = lds_f64_nc(ea);
                            while (hi < sv) { ea += 8u; hi = lds_f64_nc(ea); }
                            const double dl = __dsub_rn(sv, lds_f64_nc(ea - 8u)), dr = __dsub_rn(hi, sv);
                            pen = fx_inv_penalty(kw, dr < dl ? dr : dl);
                        }
                        const unsigned long long label = (w >> 16) + pen;
                        bool keep = label <= bnd;             // (un-hinted: only a layer-1 node inside a zone can push a label above the bound)
                        if (HINT && caprow != nullptr && !last) {
                            // exact A*-style pruning (mpc_reach.cu): the rest of the plan costs at least h = (n - r) V[q] + r V[q+1],
                            // q = D div n, r = D mod n, D = cells the ego can still advance; drop the node when label + h > bound
                            const unsigned capv = __ldg(caprow + (k >> MPC_BUCKET_SHIFT));
                            int D = (int)capv - k; D = D < 0 ? 0 : D; D = D > dmax ? dmax : D;
                            int q = (int)((float)D * rcpn), r = D - q * nrem;              // D < 2^16: the estimate is off by at most one
                            if (r < 0) { q--; r += nrem; } else if (r >= nrem) { q++; r -= nrem; }
                            const unsigned long long h = (unsigned long long)(unsigned)(nrem - r) * lds_u32_nc(tbv + 4u * q) +
                                                         (unsigned long long)(unsigned)r * lds_u32_nc(tbv + 4u * q + 4u);
                            keep = capv != 0xffffu && label + h <= bnd;
                        }
                        if (keep) {
                            MPC_EMU_COUNT_NODE();
                            const int v = 255 - (int)((w >> 8) & 0xff), a = (int)(w & 0xff) - 128;
                            if (MPC_ABLATE != 4) bp_row[k] = (uint16_t)(k - v);
                            if (last) {
                                const unsigned long long key = (label << 16) | (unsigned long long)k;
                                mybest = key < mybest ? key : mybest;
                            } else {
                                int wlo, n;
                                int_window_sa(P, g.num_s, cba, NW, k, v, a, wlo, n);
                                if (n > 0) {
                                    const int vn = wlo - k, an = vn - v, jn = an - a;
                                    const unsigned wa = bwa + 4u * (unsigned)(wlo >> 5);
                                    const unsigned open = ~__funnelshift_r(lds_u32_nc(wa), lds_u32_nc(wa + 4u), wlo & 31) & ((1u << n) - 1u);      // bits >= n are 0
                                    mylo = min(mylo, wlo); myhi = max(myhi, wlo + n - 1);
                                    const unsigned long long word = (label << 16) | ((unsigned long long)(255 - vn) << 8) | (unsigned long long)(an + 128);
                                    const unsigned tva = tbv + 4u * vn, taja = tbaj + 4u * ((an + 16) * 16 + (jn + 8));    // V[v'], A[a'] + J[j'] of successor 0
                                    const int r0 = ring(wlo);
                                    const unsigned ra = nxt + 8u * r0;
                                    if (!WRAP || r0 + n <= Wc) {                // the window does not cross the end of the ring
                                        // Five independent min-combines (five different cells), offered best-first (zero jerk, then outwards) and
                                        // software-pipelined: the pre-read of the next cell is in flight while this one's CAS is, and a CAS
                                        // that lost a race is not retried in place but noted and redone after the window (no loop, no branch).
                                        unsigned redo = 0;
                                        unsigned long long oldA, oldB, resA, resB, valA, valB;
#define MPC_PREREAD(E, OLD) OLD = lds_u64_nc(ra + 8u * (E))
#define MPC_OFFER(E, OLD, RES, VAL)                                                                                         \
                                        VAL = word - 255ULL * (E) + ((unsigned long long)(lds_u32_nc(tva + 4u * (E)) + lds_u32_nc(taja + 68u * (E))) << 16); \
                                        RES = atoms_cas_u64_if(ra + 8u * (E), OLD, VAL, ((open >> (E)) & 1u) && VAL < OLD)
#define MPC_CHECK(E, OLD, RES, VAL) redo |= (RES != OLD && VAL < RES) ? (1u << (E)) : 0u
                                        // Cells offered while another pre-read is in flight are 3 apart: neighbouring nodes have windows
                                        // shifted by 1-2 cells, so their CAS rarely lands on a cell this lane has already pre-read.
                                        MPC_PREREAD(2, oldA);
                                        MPC_OFFER(2, oldA, resA, valA);
                                        MPC_PREREAD(3, oldB); MPC_CHECK(2, oldA, resA, valA); MPC_PREREAD(0, oldA);
                                        MPC_OFFER(3, oldB, resB, valB);
                                        MPC_OFFER(0, oldA, resA, valA);
                                        MPC_CHECK(3, oldB, resB, valB); MPC_PREREAD(1, oldB);
                                        MPC_CHECK(0, oldA, resA, valA); MPC_PREREAD(4, oldA);
                                        MPC_OFFER(1, oldB, resB, valB);
                                        MPC_OFFER(4, oldA, resA, valA);
                                        MPC_CHECK(1, oldB, resB, valB);
                                        MPC_CHECK(4, oldA, resA, valA);
#undef MPC_PREREAD
#undef MPC_OFFER
#undef MPC_CHECK
                                        for (int e = 5; e < n; e++)                 // (windows longer than 5 cells: other Settings)
                                            if ((open >> e) & 1u) smem_min64(ra + 8u * e, word - 255ULL * e + ((unsigned long long)(lds_u32_nc(tva + 4u * e) + lds_u32_nc(taja + 68u * e)) << 16));
                                        while (redo) {                              // a CAS lost a race (two nodes of the warp with the same window, or another warp)
                                            const int e = __ffs(redo) - 1; redo &= redo - 1;
                                            smem_min64(ra + 8u * e, word - 255ULL * e + ((unsigned long long)(lds_u32_nc(tva + 4u * e) + lds_u32_nc(taja + 68u * e)) << 16));
                                        }
                                    } else {
                                        for (int e = 0; e < n; e++)
                                            if ((open >> e) & 1u) {
                                                const unsigned r = nxt + 8u * (r0 + e >= Wc ? r0 + e - Wc : r0 + e);
                                                smem_min64(r, word - 255ULL * e + ((unsigned long long)(lds_u32_nc(tva + 4u * e) + lds_u32_nc(taja + 68u * e)) << 16));
                                            }
                                    }
                                }
                            }
                        }
                    }
                    c = __shfl_sync(FULL, cn, 0);
                }
                mylo = __reduce_min_sync(FULL, mylo); myhi = __reduce_max_sync(FULL, myhi);
                if (last) for (int o = 16; o; o >>= 1) { unsigned long long x = __shfl_xor_sync(FULL, mybest, o); mybest = x < mybest ? x : mybest; }
                if (lane == 0) {
                    if (mybest != FX_EMPTY) atomicMin(&s_layer_best[par], mybest);
                    if (myhi >= 0) { atomicMin(&S.nlo[n3], mylo); atomicMax(&S.nhi[n3], myhi); }
                }
                if (last) {
                    __syncthreads();
                    if (s_layer_best[par] != FX_EMPTY) { bt = t; best_word = s_layer_best[par]; }
                }
            }
            __syncthreads();
            if (bt < T - 1 && !S.need_fallback) {             // the bound (or a blocked zone) cut the plan short: repeat without
                for (int k = tid; k < 2 * Wc; k += nth) sts_u64(sb0 + 8u * k, FX_EMPTY);
                // a cost hint that was too low: once more under hint_retry x the hint (CPU model: -2..-9 % nodes against going
                // straight to the standard bound), then under the standard bound, then without
                unsigned long long mid = FX_EMPTY;
                if (HINT && bnd < bound_b && io.hint_retry > 1.0) mid = fx_from_double(__dmul_rn(__dmul_rn(io.hint_cost[b], io.hint_scale), io.hint_retry));
                if (HINT && bnd < mid && mid < bound_b) bnd = mid;
                else if (HINT && bnd < bound_b) bnd = bound_b;
                else { bnd = FX_EMPTY; zone = 0; lean = false; }
                __syncthreads();
                continue;
            }
            break;
        }
        // ---- main loop: pass t finalises the nodes of layer t (buffer t&1) and pushes their successors ----
        for (int t = 2; !done && t < T; t++) {
            const int par = t & 1, s3 = t % 3, n3 = (t + 1) % 3;
            const unsigned cur = par ? sb1 : sb0, nxt = par ? sb0 : sb1;
            if (WRAP && dhi - dlo + 1 + (t == T - 1 ? 0 : P.vmax_c + 8) > Wc) { if (tid == 0) S.need_fallback = 1; break; }   // frontier may outgrow the ring
            if (tid == 0) { S.nlo[n3] = INT_MAX; S.nhi[n3] = -1; s_layer_best[par] = FX_EMPTY; s_chunk[n3] = 0; }
            __syncthreads();                                  // pushes into layer t complete; staging of layers t, t+1 visible
            if (t + 2 < T) prov.load(t + 2);                  // prefetch the search structure of layer t+2
            uint16_t *bp_row = bp + (size_t)t * io.bp_stride;
            unsigned long long mybest = FX_EMPTY;
            int mylo = INT_MAX, myhi = -1;
            bool hit = false;                                 // this thread dropped a node / closed a zone cell under the bound
            const bool last = (t == T - 1);
            auto process = [&](const int k, const int rk, const unsigned long long w) {
                sts_u64(cur + 8u * rk, FX_EMPTY);             // this buffer receives layer t+2
                double s = g.sval(k);
                double d;
                if (Prov::kClipAtPush) d = prov.distance_staged(t, k, s);      // pushes never land in a band
                else { bool ob; d = prov.eval_staged(t, k, s, ob); if (ob) return; }   // st_cy.pyx:383-384
                unsigned long long label = (w >> 16) + fx_penalty(P, d);
                if (label > bnd) { hit = true; return; }                      // cost bound: see the note above the kernel
                if (label >= FX_LABEL_LIMIT) { S.need_fallback = 1; return; }
                const int v = 255 - (int)((w >> 8) & 0xff), a = (int)(w & 0xff) - 128;
                MPC_EMU_COUNT_NODE();
                bp_row[k] = (uint16_t)(k - v);
                unsigned long long key = (label << 16) | (unsigned long long)k;
                mybest = key < mybest ? key : mybest;
                if (last) return;
                int wlo, n;
                int_window(P, g, CB, k, v, a, wlo, n);
                if (n <= 0) return;
                const int vn = wlo - k, an = vn - v, jn = an - a;      // table indices stay in range: derive_params() validated the limits
                int2 b0, b1;
                prov.bands_near(t + 1, max(wlo - zone, 0), b0, b1);
                // bit e of `open` = successor e is not inside an obstacle band of layer t+1 (st_cy.pyx:383-384) -- nor, under
                // the cost bound, within `zone` cells of one: every such cell is closer than MIN_ALLOWED_DISTANCE to the band's
                // metric edge, its penalty alone exceeds the bound, the node would be dropped when it is finalised
                unsigned open = (1u << n) - 1;
                {
                    const int l0 = max(b0.x - zone - wlo, 0), h0 = b0.x == INT_MAX ? 0 : min(b0.y + zone - wlo, n);
                    const int l1 = max(b1.x - zone - wlo, 0), h1 = b1.x == INT_MAX ? 0 : min(b1.y + zone - wlo, n);
                    if (h0 > l0) open &= ~(((1u << h0) - 1) & ~((1u << l0) - 1));
                    if (h1 > l1) open &= ~(((1u << h1) - 1) & ~((1u << l1) - 1));
                    if (zone && open != (1u << n) - 1) hit = true;          // (counts real band cells too: only costs a spare retry)
                }
                mylo = min(mylo, wlo); myhi = max(myhi, wlo + n - 1);
                unsigned ra = nxt + 8u * ring(wlo);
                const unsigned ra_end = nxt + 8u * (unsigned)Wc;
                unsigned long long word = (label << 16) | ((unsigned long long)(255 - vn) << 8) | (unsigned long long)(an + 128);
                const unsigned *tv = TB.v + vn, *taj = TB.aj + (an + 16) * 16 + (jn + 8);
#pragma unroll
                for (int e = 0; e < 5; e++) {                               // the common window has 5 cells: unrolled, predicated
                    if ((open >> e) & 1u) smem_min64(ra, word + ((unsigned long long)(tv[e] + taj[17 * e]) << 16));
                    word = word - 255ULL;                                   // v' + 1 (bits 8..15 hold 255 - v'), a' + 1 (bits 0..7)
                    ra += 8u; if (WRAP && ra == ra_end) ra = nxt;
                }
                for (int e = 5; e < n; e++) {                               // longer windows (other Settings)
                    if ((open >> e) & 1u) smem_min64(ra, word + ((unsigned long long)(tv[e] + taj[17 * e]) << 16));
                    word = word - 255ULL;
                    ra += 8u; if (WRAP && ra == ra_end) ra = nxt;
                }
            };
            // dense traversal of the layer's cell span: warps take 32-cell chunks from a shared counter (obstacle bands leave
            // long empty runs, static striding would idle whole warps).  A node-list traversal was measured 1.6-1.9x slower:
            // arrival order scatters neighbouring cells over warps (bank conflicts, uncoalesced back-pointer stores).
            for (;;) {
                int c = 0;
                if (lane == 0) c = atomicAdd(&s_chunk[s3], 1);
                c = __shfl_sync(FULL, c, 0);
                const int base = dlo + (c << 5);
                if (base > dhi) break;
                const int k = base + lane;
                if (k <= dhi) { const int rk = ring(k); const unsigned long long w = lds_u64(cur + 8u * rk); if (w != FX_EMPTY) process(k, rk, w); }
            }
            for (int o = 16; o; o >>= 1) { unsigned long long x = __shfl_xor_sync(FULL, mybest, o); mybest = x < mybest ? x : mybest; }
            mylo = warp_min_i(mylo); myhi = warp_max_i(myhi);
            hit = __any_sync(FULL, hit);
            if (lane == 0) {
                if (hit) S.bound_hit = 1;
                if (mybest != FX_EMPTY) atomicMin(&s_layer_best[par], mybest);
                if (myhi >= 0) { atomicMin(&S.nlo[n3], mylo); atomicMax(&S.nhi[n3], myhi); }
            }
            if (t + 2 < T) prov.store(t + 2);
            __syncthreads();
            if (s_layer_best[par] == FX_EMPTY) break;         // no node of layer t survived: layer t-1 is deepest
            bt = t; best_word = s_layer_best[par];
            dlo = S.nlo[n3]; dhi = S.nhi[n3];
            if (dhi < 0) break;                               // no successors (or last layer)
        }
        __syncthreads();
        // nodes were dropped by the cost bound and the horizon was not reached: the bound was too low for this problem
        // (its best path crosses a penalty zone, or it has no full-horizon path at all) -> solve it again without
        if (bnd != FX_EMPTY && S.bound_hit && bt < T - 1 && !S.need_fallback) {
            for (int k = tid; k < 2 * Wc; k += nth) sts_u64(sb0 + 8u * k, FX_EMPTY);
            bnd = FX_EMPTY; zone = 0; lean = false;
            __syncthreads();
            continue;
        }
        break;
      }
        if (S.need_fallback) {        // saturated label / out-of-range code / ring too small / bound too low: hand the problem on
            for (int k = tid; k < 2 * Wc; k += nth) sts_u64(sb0 + 8u * k, FX_EMPTY);
            if (tid == 0) { int p = atomicAdd(io.fallback_count, 1); io.fallback_list[p] = b; }
            __syncthreads();
            continue;
        }
        // an early exit leaves pushed-but-unprocessed words of layer bt+1 behind: clear them
        if (bt < T - 1 && bt >= 1) {
            const unsigned nb = ((bt + 1) & 1) ? sb1 : sb0;
            if (dhi >= 0) { for (int k = dlo + tid; k <= dhi; k += nth) sts_u64(nb + 8u * ring(k), FX_EMPTY); }
        }
        int fbk = (int)(best_word & 0xffff);
        double best_cost = (double)(best_word >> 16) * (1.0 / MPC_FX_ONE);
        finish_problem(P, io, prov, &S, b, g, bt, fbk, best_cost, bp, DESC || io.crash != nullptr);
    }
}

// ------------------------------------------------------------------------------------------------
template <class K>
static cudaError_t set_smem(K kernel, size_t smem) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

#include "mpc_fast32.cuh"

// five launch shapes are compiled: <= 192 threads x 5 blocks/SM, <= 256 x 4, <= 384 x 3, <= 512 x 2, <= 1024 x 1
template <class Prov, bool DESC, bool HINT>
static cudaError_t launch_fast_t(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, const uint8_t *ob,
                                 const void *dist, int stride, cudaStream_t st) {
    cudaError_t e;
#define MPC_LAUNCH_FAST(WRAPV, MAXTV)                                                                      \
    do {                                                                                                   \
        auto k = fast_pull_kernel<Prov, DESC, WRAPV, MAXTV, HINT>;                                         \
        if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;                                            \
        MPC_LAUNCH(k, L.grid, L.threads, L.smem, st, P, L.B, io, desc, ob, dist, stride, L.W, L.bound);                     \
    } while (0)
    if (L.threads <= 192) { if (L.wrap) MPC_LAUNCH_FAST(true, 192); else MPC_LAUNCH_FAST(false, 192); }
    else if (L.threads <= 256) { if (L.wrap) MPC_LAUNCH_FAST(true, 256); else MPC_LAUNCH_FAST(false, 256); }
    else if (L.threads <= 384) { if (L.wrap) MPC_LAUNCH_FAST(true, 384); else MPC_LAUNCH_FAST(false, 384); }
    else if (L.threads <= 512) { if (L.wrap) MPC_LAUNCH_FAST(true, 512); else MPC_LAUNCH_FAST(false, 512); }
    else { if (L.wrap) MPC_LAUNCH_FAST(true, 1024); else MPC_LAUNCH_FAST(false, 1024); }
#undef MPC_LAUNCH_FAST
    return cudaGetLastError();
}

cudaError_t launch_fast_desc(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, cudaStream_t st) {
    if (L.B <= 0) return cudaSuccess;
    if (io.hint_cost != nullptr) return launch_fast_t<FastDescProv, true, true>(P, L, io, desc, nullptr, nullptr, 0, st);
    return launch_fast_t<FastDescProv, true, false>(P, L, io, desc, nullptr, nullptr, 0, st);
}

cudaError_t launch_fast_dense(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const uint8_t *ob, const void *dist,
                              int dist_f32, int stride, cudaStream_t st) {
    if (L.B <= 0) return cudaSuccess;
    return dist_f32 ? launch_fast_t<FastDenseProv<float>, false, false>(P, L, io, nullptr, ob, dist, stride, st)
                    : launch_fast_t<FastDenseProv<double>, false, false>(P, L, io, nullptr, ob, dist, stride, st);
}

cudaError_t launch_fast32_desc(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, cudaStream_t st) {
    if (L.B <= 0) return cudaSuccess;
    return launch_fast32_desc_impl(P, L, io, desc, st);
}

cudaError_t launch_fast32_dense(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const uint8_t *ob, const void *dist,
                                int dist_f32, int stride, cudaStream_t st) {
    if (L.B <= 0) return cudaSuccess;
    return launch_fast32_dense_impl(P, L, io, ob, dist, dist_f32, stride, st);
}

size_t fast32_head_bytes(int num_s_max) { return fast32_smem_head(num_s_max); }

int fast32_occupancy(int threads, size_t smem, int wrap) {
    int n = 0;
    cudaError_t e;
#define MPC_OCC(WRAPV, MAXTV)                                                                              \
    do {                                                                                                   \
        auto k = fast32_kernel<FastDescProv, WRAPV, MAXTV>;                                                \
        if (set_smem(k, smem) != cudaSuccess) { cudaGetLastError(); return 0; }                            \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, smem);                           \
    } while (0)
    if (threads <= 192) { if (wrap) MPC_OCC(true, 192); else MPC_OCC(false, 192); }
    else if (threads <= 256) { if (wrap) MPC_OCC(true, 256); else MPC_OCC(false, 256); }
    else if (threads <= 384) { if (wrap) MPC_OCC(true, 384); else MPC_OCC(false, 384); }
    else if (threads <= 512) { if (wrap) MPC_OCC(true, 512); else MPC_OCC(false, 512); }
    else { if (wrap) MPC_OCC(true, 1024); else MPC_OCC(false, 1024); }
#undef MPC_OCC
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int fast_occupancy(int threads, size_t smem, int wrap) {
    int n = 0;
    cudaError_t e;
#define MPC_OCC(WRAPV, MAXTV)                                                                              \
    do {                                                                                                   \
        auto k = fast_pull_kernel<FastDescProv, true, WRAPV, MAXTV, false>;                                \
        if (set_smem(k, smem) != cudaSuccess) { cudaGetLastError(); return 0; }                            \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, smem);                           \
    } while (0)
    if (threads <= 192) { if (wrap) MPC_OCC(true, 192); else MPC_OCC(false, 192); }
    else if (threads <= 256) { if (wrap) MPC_OCC(true, 256); else MPC_OCC(false, 256); }
    else if (threads <= 384) { if (wrap) MPC_OCC(true, 384); else MPC_OCC(false, 384); }
    else if (threads <= 512) { if (wrap) MPC_OCC(true, 512); else MPC_OCC(false, 512); }
    else { if (wrap) MPC_OCC(true, 1024); else MPC_OCC(false, 1024); }
#undef MPC_OCC
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

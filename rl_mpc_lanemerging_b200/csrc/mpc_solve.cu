// DP solve kernels: the B200 replacement for st_cy.solve_s_t_path_fast (st_cy.pyx:315-399).
//
// The reference runs a Dijkstra over (t, s) cells.  Every edge goes from layer t to t+1 and has a
// strictly positive cost, so the first pop of a cell carries the minimum over all predecessors of
// (label + edge cost) with ties going to the smaller predecessor index (heap tuple order,
// st_cy.pyx:388): the search is a forward layered DP.  One thread block owns one
// problem; its label arrays live in shared memory; blocks are persistent and pull problems from a
// global work counter.
//
//  * exact_push_kernel  (MPC_MODE_EXACT): fp64, the reference's operation order -> index-identical
//    to st_cy.  Each node pushes its successors with shared-memory atomicMin (label), a second
//    pass resolves the winning predecessor (ties -> smaller index).
//  * fast_pull_kernel   (MPC_MODE_FAST): from layer 2 on every history point is a grid cell, so
//    speed/acceleration/jerk are small integers (cells per step^n) and the edge cost needs no
//    positions at all.  Nodes register their successor window in a shared-memory multimap keyed
//    by window start (one 32-bit atomic per node); every destination cell then *pulls* the minimum
//    over the few windows that cover it with plain loads -- no atomics on the hot edges.  The
//    distance penalty depends on the destination only and is added after the min.  fp32 labels.
//
// Both kernels get obstacle/distance information either from the compact per-layer descriptors of
// the traffic predictor (fused path, no grid in HBM) or from dense grids (drop-in path).
#include "mpc_common.cuh"
#include <limits.h>

#define FULL 0xffffffffu
#define INF_BITS 0x7ff0000000000000ULL
#define EMPTY64 0xffffffffffffffffULL
#define OVF_CAP 256

// ------------------------------------------------------------------------------------------------
// cell providers
// ------------------------------------------------------------------------------------------------
struct DescProv {
    const LayerDesc *base;   // desc + b*num_t (global)
    LayerDesc *sm;           // staging buffer (shared)
    __device__ __forceinline__ void stage(int t) {
        const LayerDesc *src = base + t;
        int n_act = src->n_act;
        if (threadIdx.x == 0) sm->n_act = n_act;
        if (threadIdx.x < n_act) { int i = threadIdx.x; sm->ef[i] = src->ef[i]; sm->eb[i] = src->eb[i]; sm->band[i] = src->band[i]; }
    }
    __device__ __forceinline__ double eval_staged(int k, double s, bool &ob) const { return cell_distance(*sm, s, k, ob); }
    __device__ __forceinline__ double eval_global(int t, int k, double s, bool &ob) const { return cell_distance(base[t], s, k, ob); }
};

template <typename DT>
struct DenseProv {
    const uint8_t *ob_base;  // obstacles + b*num_t*stride
    const DT *d_base;
    int stride, t_cur;
    __device__ __forceinline__ void stage(int t) { t_cur = t; }
    __device__ __forceinline__ double eval_staged(int k, double s, bool &ob) const {
        size_t o = (size_t)t_cur * stride + k;
        ob = ob_base[o] != 0;
        return ob ? 0.0 : (double)d_base[o];
    }
    __device__ __forceinline__ double eval_global(int t, int k, double s, bool &ob) const {
        size_t o = (size_t)t * stride + k;
        ob = ob_base[o] != 0;
        return (double)d_base[o];
    }
};

// exact successor window: st_cy.pyx:65-75 + 78-93
__device__ __forceinline__ void exact_window(const DevParams &P, double s0, double ds, double s, double s1, double s2,
                                             int &imin, int &imax_excl) {
    double dt = P.p.t_disc;
    double prev_v = __ddiv_rn(__dsub_rn(s1, s2), dt);
    double v = __ddiv_rn(__dsub_rn(s, s1), dt);
    double a = __ddiv_rn(__dsub_rn(v, prev_v), dt);
    double min_a = __dadd_rn(a, __dmul_rn(P.p.j_min, dt)); if (P.p.a_min > min_a) min_a = P.p.a_min;
    double max_a = __dadd_rn(a, __dmul_rn(P.p.j_max, dt)); if (P.p.a_max < max_a) max_a = P.p.a_max;
    double min_v = __dadd_rn(v, __dmul_rn(min_a, dt)); if (0.0 > min_v) min_v = 0.0;
    double max_v = __dadd_rn(v, __dmul_rn(max_a, dt)); if (P.p.max_speed < max_v) max_v = P.p.max_speed;
    double min_s = __dadd_rn(s, __dmul_rn(min_v, dt)), max_s = __dadd_rn(s, __dmul_rn(max_v, dt));
    double min_exact = __ddiv_rn(__dsub_rn(min_s, s0), ds);
    int mi = (int)min_exact;
    int ma = (int)__ddiv_rn(__dsub_rn(max_s, s0), ds);
    if ((double)mi < min_exact) mi += 1;
    imin = mi; imax_excl = ma + 1;
}

// kinematic part of st_cy.pyx:46-50 in the reference's operation order
__device__ __forceinline__ double exact_kin(const DevParams &P, double sn, double s, double s1, double s2) {
    double v = __ddiv_rn(__dsub_rn(sn, s), P.p.t_disc);
    double a = __ddiv_rn(__dadd_rn(__dsub_rn(sn, __dmul_rn(2.0, s)), s1), P.dt2);
    double j = __ddiv_rn(__dsub_rn(__dadd_rn(__dsub_rn(sn, __dmul_rn(3.0, s)), __dmul_rn(3.0, s1)), s2), P.dt3);
    double dv = __dsub_rn(v, P.p.desired_speed);
    return __dadd_rn(__dadd_rn(__dmul_rn(P.p.v_weight, __dmul_rn(dv, dv)), __dmul_rn(P.p.a_weight, __dmul_rn(a, a))),
                     __dmul_rn(P.p.j_weight, __dmul_rn(j, j)));
}
__device__ __forceinline__ double exact_cost(const DevParams &P, double sn, double s, double s1, double s2, double d) {
    return __dadd_rn(exact_kin(P, sn, s, s1, s2), __dmul_rn(P.p.d_weight, distance_penalty_f64(d, P.p.min_allowed_distance)));
}

__device__ __forceinline__ int warp_min_i(int v) { for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o)); return v; }
__device__ __forceinline__ int warp_max_i(int v) { for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o)); return v; }

struct BlockShared {
    int b;                  // current problem
    int nlo[2], nhi[2];     // next-layer span (double buffered by layer parity)
    int any[2];
    unsigned long long best_bits;
    int best_k;
    unsigned long long mind_bits;
    int crash;
    int ovf_cnt[2];
    int need_fallback;
    unsigned ovf[OVF_CAP];
};

// Write outputs for a finished DP: back-track from (bt, bk), then the crash test of st.py:790-802.
template <class Prov>
__device__ void finish_problem(const DevParams &P, const SolveIO &io, Prov &prov, BlockShared *S, int b, const SGrid &g,
                               int bt, int bk, double best_cost, const uint16_t *bp, bool want_crash) {
    int T = P.num_t;
    __shared__ int s_path[MPC_MAX_T];
    if (threadIdx.x == 0) {
        int k = bk;
        for (int t = bt; t > 0; t--) { s_path[t] = k; k = bp[(size_t)t * io.bp_stride + k]; }
        s_path[0] = k;
        S->mind_bits = INF_BITS; S->crash = 0;
        if (io.cost) io.cost[b] = best_cost;
        if (io.reached) io.reached[b] = bt;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        bool have = t <= bt;
        int k = have ? s_path[t] : -1;
        double sv = have ? g.sval(k) : 0.0;
        if (io.idx) io.idx[(size_t)b * T + t] = k;
        if (io.s_seq) io.s_seq[(size_t)b * T + t] = sv;
        if (want_crash && bt == T - 1) {                      // st.py:797-801
            int si = (int)__ddiv_rn(__dsub_rn(sv, g.s0), g.ds);
            bool ob;
            double d = prov.eval_global(t, si, g.sval(si), ob);
            atomicMin(&S->mind_bits, (unsigned long long)__double_as_longlong(d));
            if (d < P.crash_thresh) S->crash = 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && want_crash) {
        // trailing zeros == incomplete plan (st.py:792-796)
        double last = (bt == T - 1) ? g.sval(s_path[T - 1]) : 0.0;
        bool incomplete = (last == 0.0);
        if (io.crash) io.crash[b] = (incomplete || S->crash) ? 1 : 0;
        if (io.min_dist) io.min_dist[b] = (bt == T - 1) ? __longlong_as_double((long long)S->mind_bits)
                                                          : __longlong_as_double((long long)INF_BITS);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// exact fp64 push kernel
// ------------------------------------------------------------------------------------------------
template <class Prov, bool DESC>
__global__ void __launch_bounds__(512) exact_push_kernel(DevParams P, int B, SolveIO io, const LayerDesc *desc,
                                                          const uint8_t *dense_ob, const void *dense_d, int dense_stride,
                                                          int W, unsigned long long *glab, unsigned *ghist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ BlockShared S;
    __shared__ LayerDesc s_layer;
    unsigned long long *lab[2];
    unsigned *hist[2];
    if (glab) {   // labels too large for shared memory: per-block global scratch
        lab[0] = glab + (size_t)blockIdx.x * 2 * W; lab[1] = lab[0] + W;
        hist[0] = ghist + (size_t)blockIdx.x * 2 * W; hist[1] = hist[0] + W;
    } else {
        lab[0] = reinterpret_cast<unsigned long long *>(smem_raw); lab[1] = lab[0] + W;
        hist[0] = reinterpret_cast<unsigned *>(lab[1] + W); hist[1] = hist[0] + W;
    }
    uint16_t *bp = io.bp + (size_t)blockIdx.x * P.num_t * io.bp_stride;
    const int T = P.num_t, tid = threadIdx.x, nth = blockDim.x;
    if (io.B_dev) B = *io.B_dev;
    for (;;) {
        if (tid == 0) S.b = atomicAdd(io.work_counter, 1);
        __syncthreads();
        int wi = S.b;
        if (wi >= B) break;
        int b = io.subset ? io.subset[wi] : wi;
        SGrid g;
        double v0, a0;
        if (DESC) { g = make_sgrid(P, io.ego[4 * b], io.ego[4 * b + 1]); v0 = io.ego[4 * b + 2]; a0 = io.ego[4 * b + 3]; }
        else { g.s0 = io.s0[b]; g.ds = io.ds[b]; g.num_s = io.num_s[b]; v0 = io.v0[b]; a0 = io.a0[b]; }
        Prov prov;
        if constexpr (DESC) { prov.base = desc + (size_t)b * T; prov.sm = &s_layer; }
        else { prov.ob_base = dense_ob + (size_t)b * T * dense_stride;
               prov.d_base = reinterpret_cast<decltype(prov.d_base)>(dense_d) + (size_t)b * T * dense_stride;
               prov.stride = dense_stride; prov.t_cur = 0; }
        double est_prev = __dsub_rn(g.s0, __dmul_rn(v0, P.p.t_disc));                                   // st_cy.pyx:329
        double est_second = __dsub_rn(est_prev, __dmul_rn(P.p.t_disc, __dsub_rn(v0, __dmul_rn(a0, P.p.t_disc))));  // 330
        for (int k = tid; k < g.num_s; k += nth) { lab[0][k] = INF_BITS; lab[1][k] = INF_BITS; hist[0][k] = 0xffffffffu; hist[1][k] = 0xffffffffu; }
        __syncthreads();
        if (tid == 0) { lab[0][0] = 0ULL; hist[0][0] = 0u; }
        int lo = 0, hi = 0, bt = 0;
        for (int t = 0; t < T - 1; t++) {
            int cur = t & 1, nxt = cur ^ 1;
            prov.stage(t + 1);
            if (tid == 0) { S.nlo[cur] = INT_MAX; S.nhi[cur] = -1; }
            __syncthreads();
            int mylo = INT_MAX, myhi = -1;
            for (int pass = 0; pass < 2; pass++) {
                for (int k = lo + tid; k <= hi; k += nth) {
                    unsigned long long lb = lab[cur][k];
                    if (lb == INF_BITS) continue;
                    unsigned h = hist[cur][k];
                    int k1 = h >> 16, k2 = h & 0xffff;
                    double s = g.sval(k);
                    double s1 = t >= 1 ? g.sval(k1) : est_prev;
                    double s2 = t >= 2 ? g.sval(k2) : (t == 1 ? est_prev : est_second);
                    if (pass == 0) bp[(size_t)t * io.bp_stride + k] = (uint16_t)k1;
                    int imin, imax;
                    exact_window(P, g.s0, g.ds, s, s1, s2, imin, imax);
                    double label = __longlong_as_double((long long)lb);
                    for (int kk = imin; kk < imax && kk < g.num_s; kk++) {
                        double sn = g.sval(kk);
                        bool ob;
                        double d = prov.eval_staged(kk, sn, ob);
                        if (ob) continue;                                                   // st_cy.pyx:383-384
                        double tot = __dadd_rn(label, exact_cost(P, sn, s, s1, s2, d));      // 387-388
                        unsigned long long tb = (unsigned long long)__double_as_longlong(tot);
                        if (pass == 0) { atomicMin(&lab[nxt][kk], tb); mylo = min(mylo, kk); myhi = max(myhi, kk); }
                        else if (lab[nxt][kk] == tb) atomicMin(&hist[nxt][kk], ((unsigned)k << 16) | (unsigned)k1);
                    }
                    if (pass == 1) { lab[cur][k] = INF_BITS; hist[cur][k] = 0xffffffffu; }
                }
                if (pass == 0) {
                    mylo = warp_min_i(mylo); myhi = warp_max_i(myhi);
                    if ((tid & 31) == 0 && myhi >= 0) { atomicMin(&S.nlo[cur], mylo); atomicMax(&S.nhi[cur], myhi); }
                }
                __syncthreads();
                if (pass == 0 && S.nhi[cur] < 0) break;
            }
            if (S.nhi[cur] < 0) break;          // frontier died: layer t is the deepest, its labels are intact
            lo = S.nlo[cur]; hi = S.nhi[cur]; bt = t + 1;
        }
        // arg-min of the deepest layer (ties -> smaller index) == first pop of that layer (st_cy.pyx:365-367)
        int fb = bt & 1;
        if (tid == 0) { S.best_bits = EMPTY64; S.best_k = INT_MAX; }
        __syncthreads();
        unsigned long long mb = EMPTY64;
        for (int k = lo + tid; k <= hi; k += nth) { unsigned long long lb = lab[fb][k]; if (lb != INF_BITS && lb < mb) mb = lb; }
        for (int o = 16; o; o >>= 1) { unsigned long long x = __shfl_xor_sync(FULL, mb, o); mb = x < mb ? x : mb; }
        if ((tid & 31) == 0) atomicMin(&S.best_bits, mb);
        __syncthreads();
        unsigned long long bb = S.best_bits;
        for (int k = lo + tid; k <= hi; k += nth) if (lab[fb][k] == bb) { atomicMin(&S.best_k, k); bp[(size_t)bt * io.bp_stride + k] = (uint16_t)(hist[fb][k] >> 16); }
        __syncthreads();
        finish_problem(P, io, prov, &S, b, g, bt, S.best_k, __longlong_as_double((long long)bb), bp, DESC || io.crash != nullptr);
    }
}

// ------------------------------------------------------------------------------------------------
// fast integer/fp32 pull kernel
// ------------------------------------------------------------------------------------------------
// node word   : [label f32 : 32][v : 8][a+128 : 8][wofs : 8][n : 8]
// dest word   : [cost  f32 : 32][pred k : 16][0 : 16]
// multimap    : [slot2 : 8][slot1 : 8][slot0 : 8][count : 8]   slot = window_start - node_index
__device__ __forceinline__ float penalty_f32(double d, double min_allowed) {
    float df = (float)d;
    return (d < min_allowed) ? __fdividef(1000000.0f, fmaxf(df, 1.0f)) : __fdividef(1.0f, df);
}

// one predecessor window [w, w+n) covering destination kk: candidate = label + kinematic edge cost
__device__ __forceinline__ unsigned long long pull_candidate(const DevParams &P, unsigned long long nd, int k, int w, int kk,
                                                             unsigned long long best) {
    int n = (int)(nd & 0xff);
    if (kk >= w + n) return best;
    int v = (int)((nd >> 24) & 0xff), a = (int)((nd >> 16) & 0xff) - 128;
    int vn = kk - k, an = vn - v, jn = an - a;
    float fv = (float)vn - P.vdes_c, fa = (float)an, fj = (float)jn;
    float tot = __uint_as_float((unsigned)(nd >> 32)) + fmaf(P.cv * fv, fv, fmaf(P.ca * fa, fa, P.cj * fj * fj));
    unsigned long long cand = ((unsigned long long)__float_as_uint(tot) << 32) | ((unsigned long long)k << 16);
    return cand < best ? cand : best;
}

template <class Prov, bool DESC>
__global__ void __launch_bounds__(512) fast_pull_kernel(DevParams P, int B, SolveIO io, const LayerDesc *desc,
                                                         const uint8_t *dense_ob, const void *dense_d, int dense_stride, int W) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ BlockShared S;
    __shared__ LayerDesc s_layer;
    unsigned long long *N[2];
    unsigned *mm[2];
    N[0] = reinterpret_cast<unsigned long long *>(smem_raw); N[1] = N[0] + W;
    mm[0] = reinterpret_cast<unsigned *>(N[1] + W); mm[1] = mm[0] + W;
    uint16_t *bp = io.bp + (size_t)blockIdx.x * P.num_t * io.bp_stride;
    const int T = P.num_t, tid = threadIdx.x, nth = blockDim.x, lmax = P.lmax;
    const double min_allowed = P.p.min_allowed_distance;
    if (io.B_dev) B = *io.B_dev;
    for (;;) {
        if (tid == 0) S.b = atomicAdd(io.work_counter, 1);
        __syncthreads();
        int wi = S.b;
        if (wi >= B) break;
        int b = io.subset ? io.subset[wi] : wi;
        SGrid g;
        double v0, a0;
        if (DESC) { g = make_sgrid(P, io.ego[4 * b], io.ego[4 * b + 1]); v0 = io.ego[4 * b + 2]; a0 = io.ego[4 * b + 3]; }
        else { g.s0 = io.s0[b]; g.ds = io.ds[b]; g.num_s = io.num_s[b]; v0 = io.v0[b]; a0 = io.a0[b]; }
        Prov prov;
        if constexpr (DESC) { prov.base = desc + (size_t)b * T; prov.sm = &s_layer; }
        else { prov.ob_base = dense_ob + (size_t)b * T * dense_stride;
               prov.d_base = reinterpret_cast<decltype(prov.d_base)>(dense_d) + (size_t)b * T * dense_stride;
               prov.stride = dense_stride; prov.t_cur = 0; }
        double est_prev = __dsub_rn(g.s0, __dmul_rn(v0, P.p.t_disc));
        double est_second = __dsub_rn(est_prev, __dmul_rn(P.p.t_disc, __dsub_rn(v0, __dmul_rn(a0, P.p.t_disc))));
        for (int k = tid; k < g.num_s; k += nth) { N[0][k] = EMPTY64; N[1][k] = EMPTY64; mm[0][k] = 0u; mm[1][k] = 0u; }
        if (tid == 0) { S.need_fallback = 0; S.ovf_cnt[0] = 0; S.ovf_cnt[1] = 0; S.nlo[0] = INT_MAX; S.nhi[0] = -1; S.nlo[1] = INT_MAX; S.nhi[1] = -1; S.any[0] = 0; S.any[1] = 0; }
        // ---- prologue: layers 0 -> 1 -> 2 have off-grid history (st_cy.pyx:329-330): exact fp64 ----
        int imin0, imax0;
        exact_window(P, g.s0, g.ds, g.s0, est_prev, est_second, imin0, imax0);
        if (imax0 > g.num_s) imax0 = g.num_s;
        prov.stage(1);
        __syncthreads();
        int bt = 0, lo = 0, hi = 0;
        bool dead = false;
        if (tid < imax0 - imin0) {           // layer 1 nodes: full label (kinematics + penalty)
            int kk = imin0 + tid;
            double sn = g.sval(kk);
            bool ob; double d = prov.eval_staged(kk, sn, ob);
            if (!ob) {
                float lab1 = (float)exact_cost(P, sn, g.s0, est_prev, est_second, d);
                N[0][kk] = ((unsigned long long)__float_as_uint(lab1) << 32);
                bp[(size_t)1 * io.bp_stride + kk] = 0;
                atomicMin(&S.nlo[1], kk); atomicMax(&S.nhi[1], kk);
            }
        }
        __syncthreads();
        if (S.nhi[1] < 0) { dead = true; }       // nothing reachable at layer 1: best node is (0,0)
        else {
            lo = S.nlo[1]; hi = S.nhi[1]; bt = 1;
            int n1 = hi - lo + 1;
            for (int e = tid; e < n1 * lmax; e += nth) {      // layer 1 -> 2 edges, kinematic part only
                int k1 = lo + e / lmax, j = e % lmax;
                unsigned long long w1 = N[0][k1];
                if (w1 == EMPTY64) continue;
                double s = g.sval(k1);
                int imin, imax;
                exact_window(P, g.s0, g.ds, s, g.s0, est_prev, imin, imax);
                int kk = imin + j;
                if (kk >= imax || kk >= g.num_s) continue;
                float tot = __uint_as_float((unsigned)(w1 >> 32)) + (float)exact_kin(P, g.sval(kk), s, g.s0, est_prev);
                atomicMin(&N[1][kk], ((unsigned long long)__float_as_uint(tot) << 32) | ((unsigned long long)k1 << 16));
                atomicMin(&S.nlo[0], kk); atomicMax(&S.nhi[0], kk);
            }
            __syncthreads();
            if (S.nhi[0] < 0) dead = true;       // layer 1 nodes had no successors
        }
        int cur = 0;                             // buffer holding the deepest finalised layer (span [lo,hi])
        if (!dead) {
            int dlo = S.nlo[0], dhi = S.nhi[0];  // span of the layer-2 destination words (in N[1])
            cur = 1;
            __syncthreads();
            if (tid == 0) { S.nlo[0] = INT_MAX; S.nhi[0] = -1; S.nlo[1] = INT_MAX; S.nhi[1] = -1; }
            for (int t = 2; t < T; t++) {
                // here: N[cur] = destination words of layer t over [dlo,dhi]; N[cur^1] = nodes of layer t-1 over [lo,hi]
                int prv = cur ^ 1, par = t & 1;
                unsigned *mmA = mm[par], *mmB = mm[par ^ 1];
                prov.stage(t);
                __syncthreads();
                // every thread has finished B(t-1) and therefore consumed the parity-(t-1) slots: recycle them
                if (tid == 0) { S.ovf_cnt[par ^ 1] = 0; S.nlo[par ^ 1] = INT_MAX; S.nhi[par ^ 1] = -1; S.any[par ^ 1] = 0; }
                // ---- F(t): finalise the nodes of layer t, register their successor windows ----
                int mylo = INT_MAX, myhi = -1, myany = 0;
                for (int k = dlo + tid; k <= dhi; k += nth) {
                    mmB[k] = 0u;
                    unsigned long long w = N[cur][k];
                    if (w == EMPTY64) continue;
                    double s = g.sval(k);
                    bool ob; double d = prov.eval_staged(k, s, ob);
                    if (ob) { N[cur][k] = EMPTY64; continue; }
                    float label = fmaf(P.dw, penalty_f32(d, min_allowed), __uint_as_float((unsigned)(w >> 32)));
                    int pred = (int)((w >> 16) & 0xffff);
                    int v = k - pred;
                    int vp = (t == 2) ? pred : (int)((N[prv][pred] >> 24) & 0xff);
                    int a = v - vp;
                    bp[(size_t)t * io.bp_stride + k] = (uint16_t)pred;
                    myany = 1;
                    int alo = max(a + P.jlo_c, P.alo_c), ahi = min(a + P.jhi_c, P.ahi_c);
                    int vlo = v + alo, vhi = v + ahi;
                    if (vlo <= 0) {            // speed clamp at 0: the reference's index sits on an integer -> exact check
                        double me = __ddiv_rn(__dsub_rn(s, g.s0), g.ds);
                        int mi = (int)me; if ((double)mi < me) mi += 1;
                        vlo = mi - k;
                    }
                    bool clamp_hi = P.vmax_is_int ? (vhi >= P.vmax_c)
                                                  : ((double)v + fmin((double)a + P.jhi_r, P.ahi_r) > P.vmax_r);
                    if (clamp_hi) {
                        if (P.vmax_is_int) vhi = (int)__ddiv_rn(__dsub_rn(__dadd_rn(s, __dmul_rn(P.p.max_speed, P.p.t_disc)), g.s0), g.ds) - k;
                        else vhi = P.vmax_c;
                    }
                    int wlo = k + vlo, whi = min(k + vhi, g.num_s - 1);
                    int n = whi - wlo + 1; n = n < 0 ? 0 : n;
                    if (n > 255 || vlo > 255 || v > 255 || a < -128 || a > 127) { S.need_fallback = 1; n = 0; }
                    N[cur][k] = ((unsigned long long)__float_as_uint(label) << 32) | ((unsigned long long)(v & 0xff) << 24) |
                                ((unsigned long long)((a + 128) & 0xff) << 16) | ((unsigned long long)(vlo & 0xff) << 8) | (unsigned long long)n;
                    if (n > 0 && t < T - 1) {
                        unsigned old = atomicAdd(&mmA[wlo], 1u);
                        unsigned rank = old & 0xff;
                        if (rank < 3) reinterpret_cast<unsigned char *>(&mmA[wlo])[1 + rank] = (unsigned char)vlo;
                        else {
                            int pos = atomicAdd(&S.ovf_cnt[par], 1);
                            if (pos < OVF_CAP) S.ovf[pos] = ((unsigned)wlo << 16) | (unsigned)k; else S.need_fallback = 1;
                        }
                        mylo = min(mylo, wlo); myhi = max(myhi, whi);
                    }
                }
                mylo = warp_min_i(mylo); myhi = warp_max_i(myhi); myany = __any_sync(FULL, myany);
                if ((tid & 31) == 0) { if (myhi >= 0) { atomicMin(&S.nlo[par], mylo); atomicMax(&S.nhi[par], myhi); } if (myany) S.any[par] = 1; }
                __syncthreads();
                if (!S.any[par]) { cur = prv; break; }          // every reachable cell of layer t is an obstacle: layer t-1 is deepest
                bt = t; lo = dlo; hi = dhi;
                int nlo = S.nlo[par], nhi = S.nhi[par];
                if (t == T - 1 || nhi < 0) break;              // last layer, or no successors
                // ---- B(t): every cell of layer t+1 pulls its best predecessor ----
                int novf = min(S.ovf_cnt[par], OVF_CAP);
                for (int kk = nlo + tid; kk <= nhi; kk += nth) {
                    unsigned long long best = EMPTY64;
                    int wstart = max(kk - lmax + 1, nlo);
                    for (int w = wstart; w <= kk; w++) {
                        unsigned m = mmA[w];
                        int cnt = m & 0xff;
                        if (!cnt) continue;
                        int c3 = min(cnt, 3);
                        for (int i = 0; i < c3; i++) {
                            int k = w - (int)((m >> (8 * (i + 1))) & 0xff);
                            best = pull_candidate(P, N[cur][k], k, w, kk, best);
                        }
                        if (cnt > 3) {
                            for (int i = 0; i < novf; i++) {
                                unsigned e = S.ovf[i];
                                if ((int)(e >> 16) != w) continue;
                                int k = e & 0xffff;
                                best = pull_candidate(P, N[cur][k], k, w, kk, best);
                            }
                        }
                    }
                    N[prv][kk] = best;
                }
                dlo = nlo; dhi = nhi; cur = prv;
                // (the __syncthreads at the top of the next iteration orders B(t) before F(t+1))
            }
        }
        __syncthreads();
        if (S.need_fallback) {        // bucket overflow / out-of-range window: hand the problem to the exact kernel
            if (tid == 0) { int p = atomicAdd(io.fallback_count, 1); io.fallback_list[p] = b; }
            __syncthreads();
            continue;
        }
        // ---- arg-min over the deepest layer (labels are final node words in N[cur] over [lo,hi]) ----
        int fbk;
        unsigned long long bb;
        if (bt == 0) { fbk = 0; bb = 0ULL; }
        else {
            if (tid == 0) { S.best_bits = EMPTY64; }
            __syncthreads();
            unsigned long long mb = EMPTY64;
            for (int k = lo + tid; k <= hi; k += nth) {
                unsigned long long w = N[cur][k];
                if (w == EMPTY64) continue;
                unsigned long long key = (w & 0xffffffff00000000ULL) | (unsigned long long)k;
                mb = key < mb ? key : mb;
            }
            for (int o = 16; o; o >>= 1) { unsigned long long x = __shfl_xor_sync(FULL, mb, o); mb = x < mb ? x : mb; }
            if ((tid & 31) == 0) atomicMin(&S.best_bits, mb);
            __syncthreads();
            bb = S.best_bits;
            fbk = (int)(bb & 0xffffffffULL);
        }
        double best_cost = (double)__uint_as_float((unsigned)(bb >> 32));
        finish_problem(P, io, prov, &S, b, g, bt, fbk, best_cost, bp, DESC || io.crash != nullptr);
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <class K>
static cudaError_t set_smem(K kernel, size_t smem) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t launch_solve_desc(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, cudaStream_t st) {
    if (L.B <= 0) return cudaSuccess;
    cudaError_t e;
    if (L.mode == MPC_MODE_FAST) {
        auto k = fast_pull_kernel<DescProv, true>;
        if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;
        k<<<L.grid, L.threads, L.smem, st>>>(P, L.B, io, desc, nullptr, nullptr, 0, L.W);
    } else {
        auto k = exact_push_kernel<DescProv, true>;
        if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;
        k<<<L.grid, L.threads, L.smem, st>>>(P, L.B, io, desc, nullptr, nullptr, 0, L.W, L.glab, L.ghist);
    }
    return cudaGetLastError();
}

cudaError_t launch_solve_dense(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const uint8_t *ob, const void *dist,
                               int dist_f32, int stride, cudaStream_t st) {
    if (L.B <= 0) return cudaSuccess;
    cudaError_t e;
    if (L.mode == MPC_MODE_FAST) {
        if (dist_f32) {
            auto k = fast_pull_kernel<DenseProv<float>, false>;
            if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;
            k<<<L.grid, L.threads, L.smem, st>>>(P, L.B, io, nullptr, ob, dist, stride, L.W);
        } else {
            auto k = fast_pull_kernel<DenseProv<double>, false>;
            if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;
            k<<<L.grid, L.threads, L.smem, st>>>(P, L.B, io, nullptr, ob, dist, stride, L.W);
        }
    } else {
        if (dist_f32) {
            auto k = exact_push_kernel<DenseProv<float>, false>;
            if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;
            k<<<L.grid, L.threads, L.smem, st>>>(P, L.B, io, nullptr, ob, dist, stride, L.W, L.glab, L.ghist);
        } else {
            auto k = exact_push_kernel<DenseProv<double>, false>;
            if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;
            k<<<L.grid, L.threads, L.smem, st>>>(P, L.B, io, nullptr, ob, dist, stride, L.W, L.glab, L.ghist);
        }
    }
    return cudaGetLastError();
}

// resident blocks per SM of the descriptor-fed kernel of a mode (used to size the persistent grid)
int solve_occupancy(int mode, int desc, int threads, size_t smem) {
    int n = 0;
    cudaError_t e;
    if (mode == MPC_MODE_FAST) {
        auto k = fast_pull_kernel<DescProv, true>;
        if (set_smem(k, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, smem);
    } else {
        auto k = exact_push_kernel<DescProv, true>;
        if (set_smem(k, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, smem);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    (void)desc;
    return n;
}

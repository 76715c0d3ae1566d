// Device helpers shared by the DP kernels (exact push kernel, fast pull kernel).
#pragma once
#include "mpc_common.cuh"
#include <limits.h>

#define FULL 0xffffffffu
#define INF_BITS 0x7ff0000000000000ULL
#define EMPTY64 0xffffffffffffffffULL
#define OVF_CAP 128

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
    int nlo[3], nhi[3];     // next-layer span (buffered by layer index mod 2 or 3)
    int any[3];
    unsigned long long best_bits;
    int best_k;
    unsigned long long mind_bits;
    int crash;
    int need_fallback;
    int bound_hit;          // fast kernel: the cost bound dropped at least one node of this problem
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


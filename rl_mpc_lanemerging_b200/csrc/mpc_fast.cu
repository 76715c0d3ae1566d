// fast_pull_kernel (MPC_MODE_FAST): the throughput kernel of the fused gap-evaluation.
//
// From layer 2 on all three history points of a node are grid cells, so speed v, acceleration a
// and jerk are small integers (cells per step^n), the successor window is integer arithmetic
// (plus an exact fp64 check where the reference's speed clamps sit on an integer cell boundary),
// and the kinematic edge cost needs no positions.  Layers 0 and 1 carry the off-grid start history
// (st_cy.pyx:329-330) and are handled in exact fp64.
//
// One phase per layer (one block = one problem, arrays in shared memory behind a ring window, ONE
// __syncthreads per layer).  Phase t turns the nodes of layer t into the nodes of layer t+1; every
// destination cell k' of layer t+1 is handled by one thread:
//   1. obstacle test through the layer's sorted band table (obstacle cells stop here);
//   2. pull: the nodes of layer t registered their successor window [w, w+n) in a multimap keyed by w
//      (count + 3 inline slots per key in one 32-bit word, overflow list for a rare 4th entry); k'
//      gathers the nodes under keys k'-lmax+1 .. k' into a packed register list and takes
//      min(label + kinematic cost) with plain loads -- ties go to the smaller predecessor index
//      exactly like the reference's heap order; no atomics on the ~4 edges per cell;
//   3. finalise: distance penalty through the sorted edge table (O(1) lookup, exact fp64 threshold
//      test), label = pulled + d_w * penalty, integer successor window, back-pointer, and the new
//      node registers itself for phase t+1 with ONE 32-bit shared atomic.
// Three multimaps rotate (read / written / being cleared).
// Labels are fp64 (the reference's heuristic keeps one history per cell, so near-tie flips change
// the reachable set: fp32 labels were measured to diverge by percents on penalty-dominated states);
// the kinematic edge cost is evaluated in fp32 and accumulated in fp64.  oracle/mpc_oracle.c holds a
// CPU model of exactly this arithmetic (orc_solve_fast_model) that the GPU result is tested against
// bit for bit.
#include "mpc_solve_common.cuh"

#define EMPTY_LAB 0x7ff0000000000000ULL      // +inf

struct __align__(16) FastShared {
    LayerSearch layer[2];        // first: staged with 16-byte vector stores
    BlockShared S;
};
static_assert(sizeof(LayerSearch) % 16 == 0, "LayerSearch buffers must stay 16-byte aligned");

// cell providers for the fast kernel ------------------------------------------------------------------
struct FastDescProv {
    const LayerDesc *base;      // desc + b*num_t
    LayerSearch *sm;            // two staging buffers in shared memory
    int4 pre;                   // prefetch register
    // a LayerSearch is a contiguous tail of LayerDesc except for its 16-byte header
    static constexpr int kTail = (int)(sizeof(LayerSearch) - 16) / 16;
    __device__ __forceinline__ void load(int t) {          // issue the global loads of layer t (no wait)
        const LayerDesc *src = base + t;
        if (threadIdx.x < kTail) pre = reinterpret_cast<const int4 *>(reinterpret_cast<const char *>(src) + offsetof(LayerDesc, edge))[threadIdx.x];
        else if (threadIdx.x == kTail) pre = make_int4(src->n_edge, src->n_band, 0, 0);
    }
    __device__ __forceinline__ void store(int t) {         // park them in buffer t&1
        LayerSearch *dst = sm + (t & 1);
        if (threadIdx.x < kTail) reinterpret_cast<int4 *>(reinterpret_cast<char *>(dst) + 16)[threadIdx.x] = pre;
        else if (threadIdx.x == kTail) *reinterpret_cast<int4 *>(dst) = pre;
    }
    __device__ __forceinline__ double eval_staged(int t, int k, double s, bool &ob) const { return cell_distance_sorted(sm[t & 1], s, k, ob); }
    __device__ __forceinline__ bool is_obstacle(int t, int k, int j) const {
        const LayerSearch &L = sm[t & 1];
        int i = L.bucket_band[j], m = L.n_band;
        while (i < m && L.mband[i].y <= k) i++;
        return i < m && L.mband[i].x <= k;
    }
    __device__ __forceinline__ double eval_global(int t, int k, double s, bool &ob) const { return cell_distance(base[t], s, k, ob); }
};
static_assert(offsetof(LayerDesc, edge) % 16 == 0 && sizeof(LayerSearch) % 16 == 0, "LayerSearch must be int4-copyable");
static_assert(offsetof(LayerDesc, bucket_band) - offsetof(LayerDesc, edge) == offsetof(LayerSearch, bucket_band) - 16, "layout mismatch");

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
    __device__ __forceinline__ bool is_obstacle(int t, int k, int) const { return ob_base[(size_t)t * stride + k] != 0; }
};

// node meta (u16): [n:3][a+16:5][v:8]
__device__ __forceinline__ unsigned pack_meta(int v, int a, int n) { return (unsigned)v | ((unsigned)(a + 16) << 8) | ((unsigned)n << 13); }

// integer successor window of a node (k, v, a): cells [wlo, wlo+n).  Mirrors st_cy.pyx:65-93 for on-grid history.
__device__ __forceinline__ void int_window(const DevParams &P, const SGrid &g, int k, int v, int a, double s, int &wlo, int &n) {
    int alo = max(a + P.jlo_c, P.alo_c), ahi = min(a + P.jhi_c, P.ahi_c);
    int vlo = v + alo, vhi = v + ahi;
    if (vlo <= 0) {                        // clamp at speed 0: the reference's index sits on an integer -> exact check
        double me = __ddiv_rn(__dsub_rn(s, g.s0), g.ds);
        int mi = (int)me; if ((double)mi < me) mi += 1;
        vlo = mi - k;
    }
    bool clamp_hi = P.vmax_is_int ? (vhi >= P.vmax_c) : ((double)v + fmin((double)a + P.jhi_r, P.ahi_r) > P.vmax_r);
    if (clamp_hi) {
        if (P.vmax_is_int) vhi = (int)__ddiv_rn(__dsub_rn(__dadd_rn(s, __dmul_rn(P.p.max_speed, P.p.t_disc)), g.s0), g.ds) - k;
        else vhi = P.vmax_c;
    }
    wlo = k + vlo;
    int whi = min(k + vhi, g.num_s - 1);
    n = whi - wlo + 1; n = n < 0 ? 0 : n;
}

// per-warp reduction state of one phase
struct PhaseAcc { int lo, hi, any; };

// Finalise destination cell kk of layer tn (= t+1): penalty, label, integer window, multimap registration.
// `pulled` is min over predecessors of (label + kinematic cost); (bk, bv) the winning predecessor and its speed.
template <class Prov, bool WRAP>
__device__ __forceinline__ void finalize_node(const DevParams &P, const SGrid &g, Prov &prov, BlockShared &S, int tn, int T, int kk,
                                              double pulled, int bk, int bv, bool layer2, double *labN, unsigned short *metaN,
                                              unsigned *mmN, int ovf_par, uint16_t *bp_row, int Wc, PhaseAcc &acc) {
    const int rkk = WRAP ? (kk >= Wc ? kk - Wc : kk) : kk;
    double s = g.sval(kk);
    bool ob; double d = prov.eval_staged(tn, kk, s, ob);      // obstacle cells were filtered before the pull; d only
    double pen;
    if (d < P.p.min_allowed_distance) pen = __ddiv_rn(1000000.0, d > 1.0 ? d : 1.0);
    else pen = (double)__fdiv_rn(1.0f, (float)d);
    double label = __dadd_rn(__dmul_rn((double)P.dw, pen), pulled);
    int v = kk - bk, a = v - (layer2 ? bk : bv);
    int wlo, n;
    int_window(P, g, kk, v, a, s, wlo, n);
    if (n > 7 || wlo - kk > 255 || v > 255 || a < -16 || a > 15) { S.need_fallback = 1; n = 0; }
    labN[rkk] = label;
    metaN[rkk] = (unsigned short)pack_meta(v, a, n);
    bp_row[kk] = (uint16_t)bk;
    acc.any = 1;
    if (n > 0 && tn < T - 1) {
        const int rw = WRAP ? (wlo >= Wc ? wlo - Wc : wlo) : wlo;
        unsigned *key = &mmN[rw];
        unsigned rank = atomicAdd(key, 1u) & 0xff;
        if (rank < 3) reinterpret_cast<unsigned char *>(key)[1 + rank] = (unsigned char)(wlo - kk);
        else {
            int pos = atomicAdd(&S.ovf_cnt[ovf_par], 1);
            if (pos < OVF_CAP) S.ovf[ovf_par * OVF_CAP + pos] = ((unsigned)wlo << 16) | (unsigned)kk; else S.need_fallback = 1;
        }
        acc.lo = min(acc.lo, wlo); acc.hi = max(acc.hi, wlo + n - 1);
    }
}

template <class Prov, bool DESC, bool WRAP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) fast_pull_kernel(DevParams P, int B, SolveIO io, const LayerDesc *desc,
                                                            const uint8_t *dense_ob, const void *dense_d, int dense_stride, int Wc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ FastShared FS;
    BlockShared &S = FS.S;
    double *lab[2]; unsigned *mm[3]; unsigned short *meta[2];
    lab[0] = reinterpret_cast<double *>(smem_raw); lab[1] = lab[0] + Wc;
    mm[0] = reinterpret_cast<unsigned *>(lab[1] + Wc); mm[1] = mm[0] + Wc; mm[2] = mm[1] + Wc;
    meta[0] = reinterpret_cast<unsigned short *>(mm[2] + Wc); meta[1] = meta[0] + Wc;
    uint16_t *bp = io.bp + (size_t)blockIdx.x * P.num_t * io.bp_stride;
    const int T = P.num_t, tid = threadIdx.x, nth = blockDim.x, lmax = P.lmax;
    const double INF = __longlong_as_double((long long)EMPTY_LAB);
    auto ring = [Wc](int k) -> int { return WRAP ? (k >= Wc ? k - Wc : k) : k; };
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
        if constexpr (DESC) { prov.base = desc + (size_t)b * T; prov.sm = FS.layer; }
        else { prov.ob_base = dense_ob + (size_t)b * T * dense_stride;
               prov.d_base = reinterpret_cast<decltype(prov.d_base)>(dense_d) + (size_t)b * T * dense_stride;
               prov.stride = dense_stride; }
        prov.load(1);
        double est_prev = __dsub_rn(g.s0, __dmul_rn(v0, P.p.t_disc));
        double est_second = __dsub_rn(est_prev, __dmul_rn(P.p.t_disc, __dsub_rn(v0, __dmul_rn(a0, P.p.t_disc))));
        for (int k = tid; k < Wc; k += nth) { mm[0][k] = 0u; mm[1][k] = 0u; mm[2][k] = 0u; }
        if (tid == 0) { S.need_fallback = 0; for (int i = 0; i < 3; i++) { S.ovf_cnt[i] = 0; S.nlo[i] = INT_MAX; S.nhi[i] = -1; S.any[i] = 0; } }
        // ---- prologue: layers 0 -> 1 -> 2 have off-grid history (st_cy.pyx:329-330): exact fp64 ----
        int imin0, imax0;
        exact_window(P, g.s0, g.ds, g.s0, est_prev, est_second, imin0, imax0);
        if (imax0 > g.num_s) imax0 = g.num_s;
        prov.store(1);
        prov.load(2);
        __syncthreads();
        int bt = 0, lo = 0, hi = 0;          // deepest finalised layer, its span; its nodes live in buffer `cur`
        int cur = 0;
        bool done = false;
        // layer 1 nodes (buffer 0): full label = kinematics + penalty
        if (tid < imax0 - imin0) {
            int kk = imin0 + tid;
            double sn = g.sval(kk);
            bool ob; double d = prov.eval_staged(1, kk, sn, ob);
            double l1 = INF;
            if (!ob) {
                l1 = exact_cost(P, sn, g.s0, est_prev, est_second, d);
                bp[(size_t)1 * io.bp_stride + kk] = 0;
                atomicMin(&S.nlo[1], kk); atomicMax(&S.nhi[1], kk);
            }
            lab[0][ring(kk)] = l1;
        }
        prov.store(2);
        if (T > 3) prov.load(3);
        __syncthreads();
        int dlo = 0, dhi = -1;               // destination span of the next phase
        if (S.nhi[1] < 0) done = true;       // nothing reachable at layer 1: best node is (0,0)
        else {
            lo = S.nlo[1]; hi = S.nhi[1]; bt = 1;
            // layer 2: destinations = union of the layer-1 windows; each one pulls (ascending k1, strict <) and is finalised
            int w_lo = INT_MAX, w_hi = -1;
            for (int k1 = lo; k1 <= hi; k1++) {
                if (lab[0][ring(k1)] == INF) continue;
                int imin, imax;
                exact_window(P, g.s0, g.ds, g.sval(k1), g.s0, est_prev, imin, imax);
                if (imax > g.num_s) imax = g.num_s;
                if (imin < imax) { w_lo = min(w_lo, imin); w_hi = max(w_hi, imax - 1); }
            }
            if (w_hi < 0) done = true;       // layer-1 nodes have no successors
            else {
                PhaseAcc acc = {INT_MAX, -1, 0};
                for (int kk = w_lo + tid; kk <= w_hi; kk += nth) {
                    double sn = g.sval(kk);
                    bool ob; prov.eval_staged(2, kk, sn, ob);
                    double best = INF; int bk = 0;
                    if (!ob) {
                        for (int k1 = lo; k1 <= hi; k1++) {
                            double l1 = lab[0][ring(k1)];
                            if (l1 == INF) continue;
                            double s = g.sval(k1);
                            int imin, imax;
                            exact_window(P, g.s0, g.ds, s, g.s0, est_prev, imin, imax);
                            if (kk < imin || kk >= imax) continue;
                            double tot = __dadd_rn(l1, exact_kin(P, sn, s, g.s0, est_prev));
                            if (tot < best) { best = tot; bk = k1; }
                        }
                    }
                    if (best == INF) lab[1][ring(kk)] = INF;
                    else finalize_node<Prov, WRAP>(P, g, prov, S, 2, T, kk, best, bk, 0, true, lab[1], meta[1], mm[2], 2, bp + (size_t)2 * io.bp_stride, Wc, acc);
                }
                acc.lo = warp_min_i(acc.lo); acc.hi = warp_max_i(acc.hi); acc.any = __any_sync(FULL, acc.any);
                if ((tid & 31) == 0) { if (acc.hi >= 0) { atomicMin(&S.nlo[2], acc.lo); atomicMax(&S.nhi[2], acc.hi); } if (acc.any) S.any[2] = 1; }
                if (T > 3) prov.store(3);
                __syncthreads();
                if (!S.any[2]) done = true;  // every reachable cell of layer 2 is an obstacle: layer 1 is deepest
                else {
                    bt = 2; cur = 1; lo = w_lo; hi = w_hi;
                    dlo = S.nlo[2]; dhi = S.nhi[2];
                    if (dhi < 0 || T == 3) done = true;
                }
            }
        }
        // ---- main loop: phase t computes the nodes of layer t+1 from the nodes of layer t ----
        for (int t = 2; !done && t < T - 1; t++) {
            const int prv = cur ^ 1, p3 = t % 3, n3 = (t + 1) % 3, c3 = (t + 2) % 3, tn = t + 1;
            if (WRAP && dhi - dlo + 1 > Wc) { if (tid == 0) S.need_fallback = 1; break; }   // frontier wider than the ring
            const unsigned *mmA = mm[p3]; unsigned *mmN = mm[n3], *mmC = mm[c3];
            // recycle the slots / multimap of two phases ago (every thread passed the barrier that ended phase t-1)
            if (tid == 0) { S.ovf_cnt[c3] = 0; S.nlo[c3] = INT_MAX; S.nhi[c3] = -1; S.any[c3] = 0; }
            for (int k = lo + tid; k <= hi; k += nth) mmC[ring(k)] = 0u;
            if (tn + 1 < T) prov.load(tn + 1);                       // prefetch the search structure of layer t+2
            const int novf = min(S.ovf_cnt[p3], OVF_CAP);
            const unsigned *ovf = S.ovf + p3 * OVF_CAP;
            const double *labC = lab[cur]; const unsigned short *metaC = meta[cur];
            double *labN = lab[prv]; unsigned short *metaN = meta[prv];
            uint16_t *bp_row = bp + (size_t)tn * io.bp_stride;
            PhaseAcc acc = {INT_MAX, -1, 0};
            for (int kk = dlo + tid; kk <= dhi; kk += nth) {
                const int rkk = ring(kk);
                bool ob;
                { int j = kk >> MPC_BUCKET_SHIFT; ob = prov.is_obstacle(tn, kk, j); }
                if (ob) { labN[rkk] = INF; continue; }
                // gather the candidate predecessors into a packed register list (12 bits each = (j << 8) | (kk - k)),
                // then evaluate the list in one loop: lanes stay converged although every cell has its own candidates
                unsigned long long c0 = 0, c1 = 0; int nc = 0;
                double best = INF; int bk = INT_MAX, bv = 0;
                auto evaluate = [&]() {
                    for (int c = 0; c < nc; c++) {
                        const unsigned code = (unsigned)((c < 5 ? c0 >> (12 * c) : c1 >> (12 * (c - 5))) & 0xfff);
                        const int vn = code & 0xff, j = code >> 8;
                        const int k = kk - vn, rk = ring(k);
                        const unsigned mt = metaC[rk];
                        if (j >= (int)(mt >> 13)) continue;               // kk outside this node's window
                        const int v = mt & 0xff, a = (int)((mt >> 8) & 31) - 16;
                        const int an = vn - v, jn = an - a;
                        const float fv = (float)vn - P.vdes_c, fa = (float)an, fj = (float)jn;
                        const float kin = fmaf(__fmul_rn(P.cv, fv), fv, fmaf(__fmul_rn(P.ca, fa), fa, __fmul_rn(__fmul_rn(P.cj, fj), fj)));
                        const double tot = __dadd_rn(labC[rk], (double)kin);
                        if (tot < best || (tot == best && k < bk)) { best = tot; bk = k; bv = v; }
                    }
                    c0 = 0; c1 = 0; nc = 0;
                };
                auto append = [&](unsigned code) {
                    if (nc < 5) c0 |= (unsigned long long)code << (12 * nc);
                    else c1 |= (unsigned long long)code << (12 * (nc - 5));
                    if (++nc == 10) evaluate();                           // list full (rare): flush and keep gathering
                };
                const int jmax = min(lmax - 1, kk - dlo);
                for (int j = 0; j <= jmax; j++) {
                    const unsigned m = mmA[ring(kk - j)];
                    const int cnt = m & 0xff;
                    const int c3n = cnt < 3 ? cnt : 3;
                    for (int i = 0; i < c3n; i++) append((((m >> (8 * (i + 1))) & 0xff) + j) | (j << 8));
                    if (cnt > 3) {                                        // rare: 4th+ node of this key sits in the overflow list
                        for (int i = 0; i < novf; i++) {
                            unsigned e = ovf[i];
                            if ((int)(e >> 16) == kk - j) append((unsigned)(kk - (int)(e & 0xffff)) | (j << 8));
                        }
                    }
                }
                evaluate();
                if (best == INF) { labN[rkk] = INF; continue; }
                finalize_node<Prov, WRAP>(P, g, prov, S, tn, T, kk, best, bk, bv, false, labN, metaN, mmN, n3, bp_row, Wc, acc);
            }
            acc.lo = warp_min_i(acc.lo); acc.hi = warp_max_i(acc.hi); acc.any = __any_sync(FULL, acc.any);
            if ((tid & 31) == 0) { if (acc.hi >= 0) { atomicMin(&S.nlo[n3], acc.lo); atomicMax(&S.nhi[n3], acc.hi); } if (acc.any) S.any[n3] = 1; }
            if (tn + 1 < T) prov.store(tn + 1);
            __syncthreads();
            if (!S.any[n3]) break;                                    // layer t+1 is empty: layer t (buffer cur) is the deepest
            bt = tn; cur = prv; lo = dlo; hi = dhi;
            dlo = S.nlo[n3]; dhi = S.nhi[n3];
            if (dhi < 0) break;                                       // no successors
        }
        __syncthreads();
        if (S.need_fallback) {        // bucket overflow / frontier wider than the ring: hand the problem to the exact kernel
            if (tid == 0) { int p = atomicAdd(io.fallback_count, 1); io.fallback_list[p] = b; }
            __syncthreads();
            continue;
        }
        // ---- arg-min over the deepest layer (ties -> smaller index, st_cy.pyx:365-367) ----
        int fbk = 0; double best_cost = 0.0;
        if (bt > 0) {
            if (tid == 0) { S.best_bits = EMPTY64; }
            __syncthreads();
            unsigned long long mb = EMPTY64;
            for (int k = lo + tid; k <= hi; k += nth) {
                unsigned long long l = (unsigned long long)__double_as_longlong(lab[cur][ring(k)]);
                if (l < mb) mb = l;                        // +inf (empty) never wins over a finite label
            }
            for (int o = 16; o; o >>= 1) { unsigned long long x = __shfl_xor_sync(FULL, mb, o); mb = x < mb ? x : mb; }
            if ((tid & 31) == 0) atomicMin(&S.best_bits, mb);
            __syncthreads();
            unsigned long long bb = S.best_bits;
            if (tid == 0) S.best_k = INT_MAX;
            __syncthreads();
            for (int k = lo + tid; k <= hi; k += nth)
                if ((unsigned long long)__double_as_longlong(lab[cur][ring(k)]) == bb) atomicMin(&S.best_k, k);
            __syncthreads();
            fbk = S.best_k; best_cost = __longlong_as_double((long long)bb);
        }
        finish_problem(P, io, prov, &S, b, g, bt, fbk, best_cost, bp, DESC || io.crash != nullptr);
    }
}

// ------------------------------------------------------------------------------------------------
template <class K>
static cudaError_t set_smem(K kernel, size_t smem) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// two register budgets are compiled: <= 512 threads/block (128 registers) and <= 1024 (64 registers)
template <class Prov, bool DESC>
static cudaError_t launch_fast_t(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, const uint8_t *ob,
                                 const void *dist, int stride, cudaStream_t st) {
    cudaError_t e;
#define MPC_LAUNCH_FAST(WRAPV, MAXTV)                                                                      \
    do {                                                                                                   \
        auto k = fast_pull_kernel<Prov, DESC, WRAPV, MAXTV>;                                               \
        if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;                                            \
        k<<<L.grid, L.threads, L.smem, st>>>(P, L.B, io, desc, ob, dist, stride, L.W);                     \
    } while (0)
    if (L.threads <= 512) { if (L.wrap) MPC_LAUNCH_FAST(true, 512); else MPC_LAUNCH_FAST(false, 512); }
    else { if (L.wrap) MPC_LAUNCH_FAST(true, 1024); else MPC_LAUNCH_FAST(false, 1024); }
#undef MPC_LAUNCH_FAST
    return cudaGetLastError();
}

cudaError_t launch_fast_desc(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, cudaStream_t st) {
    if (L.B <= 0) return cudaSuccess;
    return launch_fast_t<FastDescProv, true>(P, L, io, desc, nullptr, nullptr, 0, st);
}

cudaError_t launch_fast_dense(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const uint8_t *ob, const void *dist,
                              int dist_f32, int stride, cudaStream_t st) {
    if (L.B <= 0) return cudaSuccess;
    return dist_f32 ? launch_fast_t<FastDenseProv<float>, false>(P, L, io, nullptr, ob, dist, stride, st)
                    : launch_fast_t<FastDenseProv<double>, false>(P, L, io, nullptr, ob, dist, stride, st);
}

int fast_occupancy(int threads, size_t smem, int wrap) {
    int n = 0;
    cudaError_t e;
#define MPC_OCC(WRAPV, MAXTV)                                                                              \
    do {                                                                                                   \
        auto k = fast_pull_kernel<FastDescProv, true, WRAPV, MAXTV>;                                       \
        if (set_smem(k, smem) != cudaSuccess) { cudaGetLastError(); return 0; }                            \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, smem);                           \
    } while (0)
    if (threads <= 512) { if (wrap) MPC_OCC(true, 512); else MPC_OCC(false, 512); }
    else { if (wrap) MPC_OCC(true, 1024); else MPC_OCC(false, 1024); }
#undef MPC_OCC
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

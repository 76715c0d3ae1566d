// fast_pull_kernel (MPC_MODE_FAST): the throughput kernel of the fused gap-evaluation.
//
// From layer 2 on all three history points of a node are grid cells, so speed v, acceleration a
// and jerk are small integers (cells per step^n), the successor window is integer arithmetic
// (plus an exact fp64 check where the reference's speed clamps sit on an integer cell boundary),
// and the kinematic edge cost needs no positions.  Layers 0 and 1 carry the off-grid start history
// (st_cy.pyx:329-330) and are handled in exact fp64.
//
// Per layer t (one block = one problem, arrays in shared memory, indexed through a ring window):
//   F(t)  every reachable cell of layer t: obstacle test + distance penalty through the sorted
//         per-layer search structure (O(1) lookups), label = pulled cost + penalty (fp64), integer
//         successor window [w, w+n); the node registers in a multimap keyed by w with ONE 32-bit
//         shared atomic (count + 3 inline slots per key, overflow list for the rare 4th entry).
//   B(t)  every cell k' of layer t+1 pulls min over the <= lmax keys w in (k'-lmax, k'] of
//         label + kinematic cost with plain loads; ties go to the smaller predecessor index
//         exactly like the reference's heap order.  No atomics on the ~4 edges per cell.
// Labels are fp64 (the reference's heuristic keeps one history per cell, so near-tie flips change
// the reachable set: fp32 labels were measured to diverge by percents on penalty-dominated states);
// the kinematic edge cost is evaluated in fp32 and accumulated in fp64.  oracle/mpc_oracle.c holds a
// CPU model of exactly this arithmetic (orc_solve_fast_model) that the GPU result is tested against
// bit for bit.
#include "mpc_solve_common.cuh"

#define EMPTY_LAB 0x7ff0000000000000ULL      // +inf

struct FastShared {
    BlockShared S;
    LayerSearch layer[2];
    double best_lab; int best_k;
};

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
};

// node meta (u16): [n:3][a+16:5][v:8]     destination meta (u16): v' = k' - predecessor
__device__ __forceinline__ unsigned pack_meta(int v, int a, int n) { return (unsigned)v | ((unsigned)(a + 16) << 8) | ((unsigned)n << 13); }

template <class Prov, bool DESC, bool WRAP>
__global__ void __launch_bounds__(1024, 1) fast_pull_kernel(DevParams P, int B, SolveIO io, const LayerDesc *desc,
                                                            const uint8_t *dense_ob, const void *dense_d, int dense_stride, int Wc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ FastShared FS;
    BlockShared &S = FS.S;
    double *lab[2]; unsigned *mm[2]; unsigned short *meta[2];
    lab[0] = reinterpret_cast<double *>(smem_raw); lab[1] = lab[0] + Wc;
    mm[0] = reinterpret_cast<unsigned *>(lab[1] + Wc); mm[1] = mm[0] + Wc;
    meta[0] = reinterpret_cast<unsigned short *>(mm[1] + Wc); meta[1] = meta[0] + Wc;
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
        for (int k = tid; k < Wc; k += nth) { mm[0][k] = 0u; mm[1][k] = 0u; }
        if (tid == 0) { S.need_fallback = 0; S.ovf_cnt[0] = 0; S.ovf_cnt[1] = 0; S.nlo[0] = INT_MAX; S.nhi[0] = -1; S.nlo[1] = INT_MAX; S.nhi[1] = -1; S.any[0] = 0; S.any[1] = 0; }
        // ---- prologue: layers 0 -> 1 -> 2 have off-grid history (st_cy.pyx:329-330): exact fp64 ----
        int imin0, imax0;
        exact_window(P, g.s0, g.ds, g.s0, est_prev, est_second, imin0, imax0);
        if (imax0 > g.num_s) imax0 = g.num_s;
        prov.store(1);
        prov.load(2);
        __syncthreads();
        int bt = 0, lo = 0, hi = 0;
        bool dead = false;
        // layer 1 nodes (buffer 0, not ringed: indices < lmax_exact): full label = kinematics + penalty
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
        __syncthreads();
        int dlo = 0, dhi = -1;
        if (S.nhi[1] < 0) dead = true;                 // nothing reachable at layer 1: best node is (0,0)
        else {
            lo = S.nlo[1]; hi = S.nhi[1]; bt = 1;
            // layer 2 destination range = union of the layer-1 windows; each destination pulls (ascending k1, strict <)
            int w_lo = INT_MAX, w_hi = -1;
            for (int k1 = lo; k1 <= hi; k1++) {
                if (lab[0][ring(k1)] == INF) continue;
                int imin, imax;
                exact_window(P, g.s0, g.ds, g.sval(k1), g.s0, est_prev, imin, imax);
                if (imax > g.num_s) imax = g.num_s;
                if (imin < imax) { w_lo = min(w_lo, imin); w_hi = max(w_hi, imax - 1); }
            }
            if (w_hi < 0) dead = true;                 // layer-1 nodes have no successors
            else {
                dlo = w_lo; dhi = w_hi;
                for (int kk = dlo + tid; kk <= dhi; kk += nth) {
                    double best = INF; int bk = 0;
                    double sn = g.sval(kk);
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
                    lab[1][ring(kk)] = best;
                    meta[1][ring(kk)] = (unsigned short)(kk - bk);
                }
            }
        }
        int cur = 0;                                   // buffer holding the deepest finalised layer, span [lo,hi]
        if (!dead) {
            cur = 1;
            __syncthreads();
            if (tid == 0) { S.nlo[0] = INT_MAX; S.nhi[0] = -1; S.nlo[1] = INT_MAX; S.nhi[1] = -1; }
            for (int t = 2; t < T; t++) {
                // here: buffer cur = destination words of layer t over [dlo,dhi]; buffer cur^1 = nodes of layer t-1 over [lo,hi]
                const int prv = cur ^ 1, par = t & 1;
                unsigned *mmA = mm[par], *mmB = mm[par ^ 1];
                __syncthreads();                       // B(t-1) and the staging of layer t are complete
                if (tid == 0) { S.ovf_cnt[par ^ 1] = 0; S.nlo[par ^ 1] = INT_MAX; S.nhi[par ^ 1] = -1; S.any[par ^ 1] = 0; }
                if (t + 1 < T) prov.load(t + 1);       // prefetch the next layer's search structure
                // ---- F(t) ----
                int mylo = INT_MAX, myhi = -1, myany = 0;
                for (int k = dlo + tid; k <= dhi; k += nth) {
                    const int rk = ring(k);
                    mmB[rk] = 0u;
                    double c = lab[cur][rk];
                    if (c == INF) continue;
                    double s = g.sval(k);
                    bool ob; double d = prov.eval_staged(t, k, s, ob);
                    if (ob) { lab[cur][rk] = INF; continue; }
                    double pen;
                    if (d < P.p.min_allowed_distance) pen = __ddiv_rn(1000000.0, d > 1.0 ? d : 1.0);
                    else pen = (double)__fdiv_rn(1.0f, (float)d);
                    double label = __dadd_rn(__dmul_rn((double)P.dw, pen), c);
                    int vq = meta[cur][rk];            // v' = k - predecessor
                    int pred = k - vq;
                    int v = vq;
                    int vp = (t == 2) ? pred : (int)(meta[prv][ring(pred)] & 0xff);
                    int a = v - vp;
                    bp[(size_t)t * io.bp_stride + k] = (uint16_t)pred;
                    myany = 1;
                    int alo = max(a + P.jlo_c, P.alo_c), ahi = min(a + P.jhi_c, P.ahi_c);
                    int vlo = v + alo, vhi = v + ahi;
                    if (vlo <= 0) {                    // clamp at speed 0: the reference's index sits on an integer -> exact check
                        double me = __ddiv_rn(__dsub_rn(s, g.s0), g.ds);
                        int mi = (int)me; if ((double)mi < me) mi += 1;
                        vlo = mi - k;
                    }
                    bool clamp_hi = P.vmax_is_int ? (vhi >= P.vmax_c) : ((double)v + fmin((double)a + P.jhi_r, P.ahi_r) > P.vmax_r);
                    if (clamp_hi) {
                        if (P.vmax_is_int) vhi = (int)__ddiv_rn(__dsub_rn(__dadd_rn(s, __dmul_rn(P.p.max_speed, P.p.t_disc)), g.s0), g.ds) - k;
                        else vhi = P.vmax_c;
                    }
                    int wlo = k + vlo, whi = min(k + vhi, g.num_s - 1);
                    int n = whi - wlo + 1; n = n < 0 ? 0 : n;
                    if (n > 7 || vlo > 255 || v > 255 || a < -16 || a > 15) { S.need_fallback = 1; n = 0; }
                    lab[cur][rk] = label;
                    meta[cur][rk] = (unsigned short)pack_meta(v, a, n);
                    if (n > 0 && t < T - 1) {
                        unsigned *key = &mmA[ring(wlo)];
                        unsigned rank = atomicAdd(key, 1u) & 0xff;
                        if (rank < 3) reinterpret_cast<unsigned char *>(key)[1 + rank] = (unsigned char)vlo;
                        else {
                            int pos = atomicAdd(&S.ovf_cnt[par], 1);
                            if (pos < OVF_CAP) S.ovf[pos] = ((unsigned)wlo << 16) | (unsigned)k; else S.need_fallback = 1;
                        }
                        mylo = min(mylo, wlo); myhi = max(myhi, whi);
                    }
                }
                mylo = warp_min_i(mylo); myhi = warp_max_i(myhi); myany = __any_sync(FULL, myany);
                if ((tid & 31) == 0) { if (myhi >= 0) { atomicMin(&S.nlo[par], mylo); atomicMax(&S.nhi[par], myhi); } if (myany) S.any[par] = 1; }
                if (t + 1 < T) prov.store(t + 1);
                __syncthreads();
                if (!S.any[par]) { cur = prv; break; }          // every reachable cell of layer t is an obstacle: layer t-1 is deepest
                bt = t; lo = dlo; hi = dhi;
                const int nlo = S.nlo[par], nhi = S.nhi[par];
                if (t == T - 1 || nhi < 0) break;              // last layer, or no successors
                if (WRAP && nhi - nlo + 1 > Wc) { if (tid == 0) S.need_fallback = 1; break; }   // frontier wider than the ring
                // ---- B(t) ----
                const int novf = min(S.ovf_cnt[par], OVF_CAP);
                const double *labC = lab[cur]; const unsigned short *metaC = meta[cur];
                for (int kk = nlo + tid; kk <= nhi; kk += nth) {
                    double best = INF; int bk = INT_MAX;
                    const int wstart = max(kk - lmax + 1, nlo);
                    for (int w = wstart; w <= kk; w++) {
                        const unsigned m = mmA[ring(w)];
                        const int cnt = m & 0xff;
                        if (!cnt) continue;
                        const int total = cnt <= 3 ? cnt : 3 + novf;
                        for (int i = 0; i < total; i++) {
                            int k;
                            if (i < 3) k = w - (int)((m >> (8 * (i + 1))) & 0xff);
                            else { unsigned e = S.ovf[i - 3]; if ((int)(e >> 16) != w) continue; k = e & 0xffff; }
                            const int rk = ring(k);
                            const unsigned mt = metaC[rk];
                            if (kk >= w + (int)(mt >> 13)) continue;
                            const int v = mt & 0xff, a = (int)((mt >> 8) & 31) - 16;
                            const int vn = kk - k, an = vn - v, jn = an - a;
                            const float fv = (float)vn - P.vdes_c, fa = (float)an, fj = (float)jn;
                            const float kin = fmaf(__fmul_rn(P.cv, fv), fv, fmaf(__fmul_rn(P.ca, fa), fa, __fmul_rn(__fmul_rn(P.cj, fj), fj)));
                            const double tot = __dadd_rn(labC[rk], (double)kin);
                            if (tot < best || (tot == best && k < bk)) { best = tot; bk = k; }
                        }
                    }
                    const int rkk = ring(kk);
                    lab[prv][rkk] = best;
                    meta[prv][rkk] = (unsigned short)(kk - bk);
                }
                dlo = nlo; dhi = nhi; cur = prv;
            }
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

cudaError_t launch_fast_desc(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, cudaStream_t st) {
    if (L.B <= 0) return cudaSuccess;
    cudaError_t e;
    if (L.wrap) {
        auto k = fast_pull_kernel<FastDescProv, true, true>;
        if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;
        k<<<L.grid, L.threads, L.smem, st>>>(P, L.B, io, desc, nullptr, nullptr, 0, L.W);
    } else {
        auto k = fast_pull_kernel<FastDescProv, true, false>;
        if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;
        k<<<L.grid, L.threads, L.smem, st>>>(P, L.B, io, desc, nullptr, nullptr, 0, L.W);
    }
    return cudaGetLastError();
}

template <typename DT>
static cudaError_t launch_fast_dense_t(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const uint8_t *ob, const void *dist,
                                       int stride, cudaStream_t st) {
    cudaError_t e;
    if (L.wrap) {
        auto k = fast_pull_kernel<FastDenseProv<DT>, false, true>;
        if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;
        k<<<L.grid, L.threads, L.smem, st>>>(P, L.B, io, nullptr, ob, dist, stride, L.W);
    } else {
        auto k = fast_pull_kernel<FastDenseProv<DT>, false, false>;
        if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;
        k<<<L.grid, L.threads, L.smem, st>>>(P, L.B, io, nullptr, ob, dist, stride, L.W);
    }
    return cudaGetLastError();
}

cudaError_t launch_fast_dense(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const uint8_t *ob, const void *dist,
                              int dist_f32, int stride, cudaStream_t st) {
    if (L.B <= 0) return cudaSuccess;
    return dist_f32 ? launch_fast_dense_t<float>(P, L, io, ob, dist, stride, st) : launch_fast_dense_t<double>(P, L, io, ob, dist, stride, st);
}

int fast_occupancy(int threads, size_t smem, int wrap) {
    int n = 0;
    cudaError_t e;
    if (wrap) {
        auto k = fast_pull_kernel<FastDescProv, true, true>;
        if (set_smem(k, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, smem);
    } else {
        auto k = fast_pull_kernel<FastDescProv, true, false>;
        if (set_smem(k, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, smem);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

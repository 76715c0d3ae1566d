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
#include "mpc_solve_common.cuh"

// ------------------------------------------------------------------------------------------------
// exact fp64 push kernel
// ------------------------------------------------------------------------------------------------
template <class Prov, bool DESC>
__global__ void __launch_bounds__(512) exact_push_kernel(DevParams P, int B, SolveIO io, const LayerDesc *desc,
                                                          const uint8_t *dense_ob, const void *dense_d, int dense_stride,
                                                          int W, unsigned long long *glab, unsigned *ghist) {
#ifdef MPC_HOST_EMU
    unsigned char *const smem_raw = emu::S().dyn_smem;
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
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
// host launchers
// ------------------------------------------------------------------------------------------------
template <class K>
static cudaError_t set_smem(K kernel, size_t smem) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t launch_exact_desc(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, cudaStream_t st) {
    if (L.B <= 0) return cudaSuccess;
    auto k = exact_push_kernel<DescProv, true>;
    cudaError_t e = set_smem(k, L.smem);
    if (e != cudaSuccess) return e;
    MPC_LAUNCH(k, L.grid, L.threads, L.smem, st, P, L.B, io, desc, nullptr, nullptr, 0, L.W, L.glab, L.ghist);
    return cudaGetLastError();
}

cudaError_t launch_exact_dense(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const uint8_t *ob, const void *dist,
                               int dist_f32, int stride, cudaStream_t st) {
    if (L.B <= 0) return cudaSuccess;
    cudaError_t e;
    if (dist_f32) {
        auto k = exact_push_kernel<DenseProv<float>, false>;
        if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;
        MPC_LAUNCH(k, L.grid, L.threads, L.smem, st, P, L.B, io, nullptr, ob, dist, stride, L.W, L.glab, L.ghist);
    } else {
        auto k = exact_push_kernel<DenseProv<double>, false>;
        if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;
        MPC_LAUNCH(k, L.grid, L.threads, L.smem, st, P, L.B, io, nullptr, ob, dist, stride, L.W, L.glab, L.ghist);
    }
    return cudaGetLastError();
}

// resident blocks per SM (sizes the persistent grid)
int exact_occupancy(int threads, size_t smem) {
    int n = 0;
    auto k = exact_push_kernel<DescProv, true>;
    if (set_smem(k, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// reach_caps_kernel: per-bucket reachability caps for the fast kernel's exact A*-style pruning (hinted solves only).
//
// Not part of the reference (its Dijkstra, st_cy.pyx:315-399, has neither a bound nor a heuristic): this is a device
// that lets the layered DP expand fewer nodes without changing its answer.  The argument (DESIGN.md §3 "Cost hints",
// CPU model + exactness test: oracle/bound_model.py, tests/test_bound_model_cpu.py):
//   * a node may be dropped when label + h > U if h depends on the CELL only and is consistent
//     (h(p) <= edge(p -> c) + h(c) on every edge, h = 0 on the last layer);
//   * such an h: the ego cannot step over a blocked interval (LayerDesc::blk: obstacle band + penalty zones of a car), so
//     from cell k of layer t it advances at most D = cap - k cells in the n = T-1-t remaining steps, where cap bounds
//     the highest cell of the LAST layer any cell path from k can reach; the speed term of the edge cost
//     (st_cy.pyx:47,50) is convex in the step length, so the rest of the plan costs at least n * V(D / n).
// cap is kept per bucket of 64 cells (MPC_BUCKET_SHIFT):
//   cap[T-1][j] = highest free cell of bucket j                                  (-1: none)
//   cap[t][j]   = max cap[t+1][j'] over j' = j .. j + reach, reach = (63 + vmax_c) >> 6   (-1 when j has no free cell)
// which makes cap[t][bucket(k)] >= cap[t+1][bucket(k')] on every edge k -> k' (0 <= k' - k <= vmax_c): all the proof needs.
//
// One warp per episode, lanes over the buckets, layers backwards.  Output: u16 [B][T][stride], 0xffff = no path from here.
#include "mpc_common.cuh"

#define RC_WARPS 4
#define RC_PER_LANE ((MPC_MAX_BUCKETS + 31) / 32)

__global__ void __launch_bounds__(32 * RC_WARPS) reach_caps_kernel(DevParams P, int B, const LayerDesc *__restrict__ desc,
                                                                  const int32_t *__restrict__ num_s,
                                                                  unsigned short *__restrict__ capb, int stride) {
    __shared__ int s_cap[RC_WARPS][MPC_MAX_BUCKETS + 8];          // caps of layer t+1 (-1 padded behind the last bucket)
    __shared__ int2 s_blk[RC_WARPS][MPC_NMAX];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * RC_WARPS + w;
    if (b >= B) return;                                           // (warp-uniform; only __syncwarp below)
    const int T = P.num_t, ns = num_s[b];
    const int NB = min((ns + 63) >> MPC_BUCKET_SHIFT, min(stride, MPC_MAX_BUCKETS));
    const int reach = (63 + P.vmax_c) >> MPC_BUCKET_SHIFT;
    int *cap = s_cap[w];
    int2 *blk = s_blk[w];
    for (int j = lane; j < MPC_MAX_BUCKETS + 8; j += 32) cap[j] = -1;
    __syncwarp();
    for (int t = T - 1; t >= 0; t--) {
        const LayerDesc &L = desc[(size_t)b * T + t];
        const int nb = (t == 0) ? 0 : min(L.n_blk, MPC_NMAX);     // the start cell is never tested (st_cy.pyx:383-384 act on successors)
        if (lane < nb) blk[lane] = L.blk[lane];
        __syncwarp();
        int nv[RC_PER_LANE];
#pragma unroll
        for (int i = 0; i < RC_PER_LANE; i++) {
            const int j = lane + 32 * i;
            int v = -1;
            if (j < NB) {
                int c = min((j << MPC_BUCKET_SHIFT) + 63, ns - 1);                  // highest free cell of the bucket
                for (int q = nb - 1; q >= 0; q--) { const int2 iv = blk[q]; if (c >= iv.x && c < iv.y) c = iv.x - 1; }
                if (c >= (j << MPC_BUCKET_SHIFT)) {
                    if (t == T - 1) v = c;
                    else for (int r = 0; r <= reach && r < 8; r++) v = max(v, cap[j + r]);
                }
            }
            nv[i] = v;
        }
        __syncwarp();                                             // every lane has read cap[] of layer t+1
#pragma unroll
        for (int i = 0; i < RC_PER_LANE; i++) {
            const int j = lane + 32 * i;
            if (j < NB) {
                cap[j] = nv[i];
                capb[((size_t)b * T + t) * stride + j] = nv[i] < 0 ? (unsigned short)0xffff : (unsigned short)nv[i];
            }
        }
        __syncwarp();
    }
}

// Ordered compaction of a byte mask: subset[0 .. *count) = the indices b with mask[b] != 0, ascending.  One block.
__global__ void __launch_bounds__(1024) compact_mask_kernel(const uint8_t *__restrict__ mask, int B, int32_t *__restrict__ subset,
                                                           int *__restrict__ count) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < B; start += blockDim.x) {
        const int b = start + tid;
        const bool on = b < B && mask[b] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if (lane == 0) s_warp[w] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int i = 0; i < w; i++) off += s_warp[i];
        if (on) subset[off + __popc(bal & ((1u << lane) - 1u))] = b;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int i = 0; i < nw; i++) t += s_warp[i]; s_base += t; }
        __syncthreads();
    }
    if (tid == 0) *count = s_base;
}

cudaError_t launch_compact_mask(const uint8_t *mask, int B, int32_t *subset, int *count, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    MPC_LAUNCH(compact_mask_kernel, 1, 1024, 0, st, mask, B, subset, count);
    return cudaGetLastError();
}

cudaError_t launch_reach_caps(const DevParams &P, int B, const LayerDesc *desc, const int32_t *num_s, unsigned short *capb,
                              int stride, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    MPC_LAUNCH(reach_caps_kernel, (B + RC_WARPS - 1) / RC_WARPS, 32 * RC_WARPS, 0, st, P, B, desc, num_s, capb, stride);
    return cudaGetLastError();
}

// Shared device-side definitions for libmpcb200 (sm_100a).
// All fp64 arithmetic that must be bit-identical to the reference is written with explicit
// round-to-nearest intrinsics (__dadd_rn/__dmul_rn/__ddiv_rn) so that nvcc never contracts a
// multiply-add into an FMA; the library is also compiled with -fmad=false and uses fmaf()
// explicitly where a fused fp32 multiply-add is wanted.
#pragma once
#ifdef MPC_HOST_EMU               // tests/emu: the kernels compiled by g++ and run on fibers (test infrastructure only)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include "../../include/mpcb200.h"

// kernel launch: the CUDA launch syntax for nvcc; the fiber scheduler of tests/emu when the sources are compiled by g++
#ifdef MPC_HOST_EMU
#define MPC_LAUNCH(kernel, grid, block, smem, st, ...) emu::launch((grid), (block), (smem), [=] { kernel(__VA_ARGS__); })
#else
#define MPC_LAUNCH(kernel, grid, block, smem, st, ...) kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__)
#endif

#define MPC_NMAX 32            // cars per problem handled by one warp (lane == car)
#define MPC_MAX_T 128          // time layers supported (H <= 127)

// Device copy of the parameters plus everything the host derives once per mpc_set_params.
struct DevParams {
    mpc_params p;
    int num_t;                 // len(arange(0, FUTURE_T + dt, dt))            (st.py:32)
    int num_s_max;             // upper bound of len(arange(s0, s0+FUTURE_S+ds, ds)) over s0
    int discrete_length;       // int(CAR_LENGTH / delta_s)                     (st.py:37)
    int lmax;                  // max successor-window length in cells, on-grid layers (fast kernel)
    int lmax_exact;            // same for the off-grid first layers (+1 safety)
    double dt2, dt3;           // pow(dt,2), pow(dt,3) as libm computes them     (st_cy.pyx:48-49)
    double obs_min_s;          // CRASH_MIN_S - MIN_ALLOWED_DISTANCE            (st.py:46)
    double crash_thresh;       // COMBINATION_MIN_DISTANCE - CAR_LENGTH         (st.py:800)
    // ---- fast mode (integer-cell kinematics); valid only when fast_ok != 0 --------------------
    int fast_ok;
    int jlo_c, jhi_c;          // jerk window in cells/step^3:   a' in [a+jlo_c, a+jhi_c]
    int alo_c, ahi_c;          // accel bounds in cells/step^2
    int vmax_c;                // floor(max_speed*dt/ds)
    int vmax_is_int;           // max_speed*dt/ds sits on an integer -> per-cell fp64 check needed
    double jhi_r, ahi_r, vmax_r; // real-valued (cells) jerk/accel/speed upper bounds for the non-integer clamp test
    float cv, ca, cj;          // v_w*(ds/dt)^2, a_w*(ds/dt^2)^2, j_w*(ds/dt^3)^2
    float vdes_c;              // desired speed in cells/step
    float dw;                  // d_weight
    // fixed-point (2^-MPC_FX_FRAC) kinematic edge-cost tables of the fast kernel, indexed by the integer
    // speed v' in [0,255], acceleration a'+16 in [0,31], jerk j'+8 in [0,15] (cells per step^n)
    unsigned vtab[256], atab[32], jtab[16];
    unsigned long long bound_fx;   // first-pass cost bound of the fast kernel in label units (0 = none), see mpc_fast.cu
    int zone_cells;                // cells next to a band that are certainly inside its penalty zone (>= 0)
    int zone_ok;                   // LayerDesc::blk (band + exact penalty zone per car) is one interval per car: lean bounded pass allowed
    float kw;                      // (float)(d_weight * 2^MPC_FX_FRAC): 1/d penalty in label units = rint(kw * (1.0f / (float)d))
    int vstar_c;                   // cheapest speed of vtab on [0, vmax_c] (cells/step): vtab is convex and non-increasing up to it
    // ---- 32-bit-key kernel (mpc_fast32.cuh): labels in 2^-f32_frac fixed point that fit 32 bits; valid only when f32_ok != 0 ----
    int f32_ok;
    int f32_frac;                  // fraction bits of its labels (<= MPC_FX_FRAC: the longer the horizon, the fewer)
    unsigned f32_bound;            // its cost bound in label units: label + any edge + any penalty stays below 0xff000000
    float kw32;                    // (float)(d_weight * 2^f32_frac)
    double f32_one;                // 2^f32_frac
    unsigned vtab32[256], atab32[32], jtab32[16];   // the kinematic tables in 2^-f32_frac units
};
#define MPC_FX_FRAC 18
#define MPC_FX_ONE 262144.0

// ---- geometry: control.py:366-380 ---------------------------------------------------------------
__device__ __forceinline__ double dist2d(double x0, double y0, double x1, double y1) {
    double dx = __dsub_rn(x0, x1), dy = __dsub_rn(y0, y1);
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}
__device__ __forceinline__ double get_ego_s(double x, double y) {
    if (x < -50.9) return -dist2d(x, y, -50.9, 1.72);
    else if (x < 1.5) return dist2d(x, y, -50.9, 1.72);
    else return __dadd_rn(__dsub_rn(x, 1.5), 52.5);
}

// numpy.arange(start, start + future_s + ds, ds): length and fill rule (st.py:31)
__device__ __forceinline__ int arange_len(double start, double stop, double step) {
    double c = ceil(__ddiv_rn(__dsub_rn(stop, start), step));
    return c < 0.0 ? 0 : (int)c;
}
struct SGrid {
    double s0, ds;     // s_values[0], s_values[1]-s_values[0]
    int num_s;
    __device__ __forceinline__ double sval(int k) const { return __dadd_rn(s0, __dmul_rn((double)k, ds)); }
};
__device__ __forceinline__ SGrid make_sgrid(const DevParams &P, double ex, double ey) {
    SGrid g;
    g.s0 = get_ego_s(ex, ey);
    g.num_s = arange_len(g.s0, __dadd_rn(__dadd_rn(g.s0, P.p.future_s), P.p.s_disc), P.p.s_disc);
    g.ds = __dsub_rn(__dadd_rn(g.s0, P.p.s_disc), g.s0);
    return g;
}

// st_cy.pyx:34-38
__device__ __forceinline__ double distance_penalty_f64(double d, double min_allowed) {
    if (d < min_allowed) return __ddiv_rn(1000000.0, d > 1.0 ? d : 1.0);
    return __ddiv_rn(1.0, d);
}

// One layer of the compact obstacle description produced by the traffic predictor (K1a) and
// consumed by the rasteriser (K1b) and the fused planner (K3).  For every car that passes the
// reference's range filters (st.py:46-49): the two edges the distance field is measured to
// (st.py:52-53) and the obstacle band [imin, imax) in cells (st.py:60-65).
#define MPC_BUCKET_SHIFT 6       // lookup buckets of 64 cells (3.2 m at the published discretisation)
#define MPC_MAX_BUCKETS 288      // supports num_s <= 18432 (H = 100 needs 18001)
struct LayerDesc {
    int n_act;                   // cars in reference order (exact kernel, rasteriser)
    int n_edge;                  // sorted distance-field edges (= 2 * n_act)
    int n_band;                  // obstacle bands merged into disjoint intervals
    int n_blk;                   // blocked intervals (see blk)
    double ef[MPC_NMAX];
    double eb[MPC_NMAX];
    int2 band[MPC_NMAX];
    // ---- search structure for the fast kernel (same information, sorted) ----
    double edge[2 * MPC_NMAX + 2];           // ascending; edge[n_edge] = +1e300 (sentinel for the branch-free lookup)
    int2 mband[MPC_NMAX];                    // disjoint [x, y), ascending
    unsigned char bucket_edge[MPC_MAX_BUCKETS];   // #edges  <  s_values[64*j]
    unsigned char bucket_band[MPC_MAX_BUCKETS];   // #merged bands with y <= 64*j
    // cells a bounded plan can never use: in an obstacle band, or closer than MIN_ALLOWED_DISTANCE to a distance-field
    // edge (penalty 1e6/max(d,1), st_cy.pyx:34-38).  Per car one interval [first cell with |s-ef| < m, last cell with
    // |s-eb| < m] (exact fp64 tests) joined with its band; merged into disjoint [x, y), ascending.
    int2 blk[MPC_NMAX];
};

// the part of a LayerDesc the fast kernel stages in shared memory
struct LayerSearch {
    int n_edge, n_band, n_blk, pad1;
    double pad2, edge_lo;        // edge_lo = edge[-1] = -1e300 (sentinel)
    double edge[2 * MPC_NMAX + 2];
    int2 mband[MPC_NMAX];
    unsigned char bucket_edge[MPC_MAX_BUCKETS];
    unsigned char bucket_band[MPC_MAX_BUCKETS];
    int2 blk[MPC_NMAX];
};

// distance-field value / obstacle flag through the sorted structure: bit-identical to cell_distance()
// (the nearest edge on either side decides the min; |s-e| is monotone in e on each side).
template <class LS>
__device__ __forceinline__ double cell_distance_sorted(const LS &L, double s, int k, bool &obstacle) {
    int j = k >> MPC_BUCKET_SHIFT;
    int e = L.bucket_edge[j], M = L.n_edge;
    while (e < M && L.edge[e] < s) e++;
    double d = 1E10;
    if (e > 0) { double x = __dsub_rn(s, L.edge[e - 1]); d = x < d ? x : d; }
    if (e < M) { double x = fabs(__dsub_rn(s, L.edge[e])); d = x < d ? x : d; }
    int i = L.bucket_band[j], m = L.n_band;
    while (i < m && L.mband[i].y <= k) i++;
    bool ob = i < m && L.mband[i].x <= k;
    obstacle = ob;
    return ob ? 0.0 : d;
}

// distance-field value and obstacle flag of cell k in a layer (st.py:52-65), exact fp64
__device__ __forceinline__ double cell_distance(const LayerDesc &L, double s, int k, bool &obstacle) {
    double d = 1E10;
    bool ob = false;
    for (int c = 0; c < L.n_act; c++) {
        double df = fabs(__dsub_rn(s, L.ef[c])), db = fabs(__dsub_rn(s, L.eb[c]));
        d = df < d ? df : d;
        d = db < d ? db : d;
        int2 b = L.band[c];
        ob = ob || (k >= b.x && k < b.y);
    }
    obstacle = ob;
    return ob ? 0.0 : d;
}

// ---- solver I/O shared by the kernel file and the API file ---------------------------------------
struct SolveIO {
    // per-problem inputs
    const double *s0, *ds;       // s_values[0], s_values[1]-s_values[0]
    const int32_t *num_s;
    const double *v0, *a0;       // ego speed / acceleration;  v0 == NULL -> read ego[4b+2], ego[4b+3]
    const double *ego;
    // outputs (any may be NULL)
    int32_t *idx; double *s_seq; double *cost; int32_t *reached; uint8_t *crash; double *min_dist;
    // scratch
    uint16_t *bp;                // [gridDim.x][num_t][bp_stride]
    int bp_stride;
    int *work_counter;
    const int32_t *subset;       // optional list of problem ids (fallback re-solve); NULL = 0..B-1
    const int *B_dev;            // optional: number of problems read on the device (overrides B)
    int32_t *fallback_list; int *fallback_count;   // fast kernel: problems that need the exact kernel
    int *overflow_count;         // 32-bit-key kernel (optional): how many of its hand-overs were frontiers wider than its ring
    int *flagged_count;          // 32-bit-key kernel (optional): how many of its hand-overs carry bit 30 (bounded attempt failed)
    int n_lo, n_hi;              // 64-bit kernel: when n_hi != 0 the launch only works if its list holds n_lo <= n <= n_hi entries (two
                                 // launch shapes are enqueued for the hand-over list; its length, known on the device only, picks one)
    int skip_flagged;            // 32-bit-key kernel, second shape: list entries with bit 30 belong to the 64-bit kernel's concurrent launch
    int only_flagged;            // 64-bit kernel: take only the list entries with bit 30 (the others are with the second shape)
    // optional per-problem cost hint of the fast kernel (mpc_plan_hinted): first bound = hint_scale * hint_cost[b], used when
    // hint_reached == NULL or hint_reached[b] == hint_full_t
    const double *hint_cost; const int32_t *hint_reached; int hint_full_t; double hint_scale;
    double hint_retry;           // second attempt under hint_retry * first bound (<= 1: straight to the standard bound)
    // hinted solves: per-bucket reachability caps u16 [B][num_t][cap_stride] (mpc_reach.cu), NULL = no heuristic pruning
    const unsigned short *capb; int cap_stride;
};

struct SolveLaunch {
    int B;
    int grid, threads;
    size_t smem;
    int W;                  // label-array length in cells (exact: full row; fast: ring capacity)
    int wrap;               // fast kernel: ring shorter than the row
    unsigned long long bound;   // fast kernel: cost bound in label units (~0ULL = none)
    unsigned long long *glab; unsigned *ghist;    // exact kernel global label scratch (or NULL)
};

#define MPC_CUDA_OK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return mpc_set_cuda_error(e__, #call); } while (0)
int mpc_set_cuda_error(cudaError_t e, const char *what);
int mpc_set_error(int code, const char *msg);

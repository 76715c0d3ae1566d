// K1a traffic predictor -> per-layer obstacle descriptors, K1b dense S-T rasteriser,
// K4 rollout-tick pieces (predict_step_with_ego, observation vector, jerk->speed).
//
// One warp per problem, lane == car (front -> back).  The reference updates the cars strictly
// sequentially because a follower reads its leader's *already updated* position and speed
// (prediction.py:93-94).  That chain is a fixed point new_i = F(old_i, new_{i-1}); we iterate all
// lanes in parallel (Jacobi) until nothing changes, which yields exactly the sequential result
// (after m sweeps the first m cars are final) in 1-3 sweeps for ordinary traffic.
#include "mpc_common.cuh"
#include <limits.h>

#define FULL 0xffffffffu

struct EgoState { double x, y, v, a; };

// prediction.py:46-105 for one warp.  (x,v) are this lane's car (lane < n valid).
// Returns crashed; writes the predicted ego and this lane's new car state.
__device__ __forceinline__ bool warp_predict_with_ego(const DevParams &P, int lane, int n, const EgoState &ego,
                                                      double x, double v, double sel, double dt,
                                                      double min_crash_distance, EgoState &ego_out,
                                                      double &nx, double &nv, double &na) {
    double px, py;
    if (ego.x < 1.5) {                                              // 48-56
        double dx = __dsub_rn(1.5, ego.x), dy = __dsub_rn(-1.5, ego.y);
        double nrm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        dx = __ddiv_rn(dx, nrm); dy = __ddiv_rn(dy, nrm);
        double step = __dmul_rn(sel, dt);
        dx = __dmul_rn(dx, step); dy = __dmul_rn(dy, step);
        px = __dadd_rn(ego.x, dx); py = __dadd_rn(ego.y, dy);
        if (py < -1.6) py = -1.6;
    } else {                                                        // 57-59
        py = ego.y; px = __dadd_rn(ego.x, __dmul_rn(sel, dt));
    }
    ego_out.a = __ddiv_rn(__dsub_rn(sel, ego.v), dt);               // 61
    ego_out.x = px; ego_out.y = py; ego_out.v = sel;
    double pes = get_ego_s(px, py);
    bool can_crash = pes > 11.0, merged = pes > 8.0;                // 64, 66
    bool valid = lane < n;
    unsigned behind = __ballot_sync(FULL, valid && x < px);         // 78: first car behind the predicted ego
    int enc = behind ? __ffs(behind) - 1 : -1;
    bool ego_leads = merged && lane == enc;

    nv = v; na = 0.0; nx = __dadd_rn(x, __dmul_rn(v, dt));         // free-flow guess
    for (int it = 0; it <= n; it++) {
        double lx = __shfl_up_sync(FULL, nx, 1), lv = __shfl_up_sync(FULL, nv, 1);
        if (lane == 0) { lx = __longlong_as_double(0x7ff0000000000000LL); lv = 0.0; }   // 72-73
        if (ego_leads) { lx = px; lv = sel; }                       // 80-82
        double speed_diff = __dsub_rn(lv, v), x_diff = __dsub_rn(lx, x), a2, v2;
        if (speed_diff < 0.0 && x_diff < 30.0) {                    // 85-87
            a2 = (P.p.max_predicted_decel > speed_diff) ? P.p.max_predicted_decel : speed_diff;
            v2 = __dadd_rn(v, __dmul_rn(a2, dt));
        } else { a2 = 0.0; v2 = v; }                                // 88-90
        double x2 = __dadd_rn(x, __dmul_rn(v2, dt));                // 91
        bool changed = valid && (x2 != nx || v2 != nv || a2 != na);
        nx = x2; nv = v2; na = a2;
        if (!__any_sync(FULL, changed)) break;
    }
    double cdd = min_crash_distance > P.p.car_length ? min_crash_distance : P.p.car_length;   // 100
    bool hit = valid && can_crash && fabs(__dsub_rn(nx, px)) < cdd;                           // 101-103
    return __any_sync(FULL, hit);
}

// The SUMO-free world with the car-following model of the reference's traffic (merge_impossible.rou.xml:3: vType "normal",
// carFollowModel Krauss, accel 4.5, decel 6.0, minGap 1, sigma 0, tau 0.5; maxSpeed = OTHER_CAR_SPEED, sumo.py:60; Euler update with
// step-length TICK_LENGTH, ramp.sumocfg:8-20).  Every car computes its safe speed from its leader's state at the BEGINNING of the
// step, like SUMO's planMove, so there is no sequential chain:
//     v_safe = -b tau + sqrt((b tau)^2 + v_lead^2 + 2 b gap),  gap = x_lead - length - x - minGap      (Krauss 1998, SUMO's "Krauss")
//     v'     = max(0, v - b dt, min(v + a dt, v_safe, v_max)),   x' = x + v' dt
// The leader is the next car ahead, or the ego once it is on the junction's merging lane / the highway (ego_s >= 0; on the
// junction its position along the lane is ego_s - 51, control.py:366-380) and sits between the two.  The ego itself moves as in
// predict_step_with_ego and the crash test is that function's too (prediction.py:46-61, 97-103).  Parity against SUMO is not
// pinned (no SUMO here): the world is judged on the aggregate statistics of whole episodes (tools/closed_loop_stats.py).
struct KraussParams { double accel, decel, tau, min_gap, max_speed; };
__device__ __forceinline__ bool warp_krauss_with_ego(const DevParams &P, const KraussParams &K, int lane, int n, const EgoState &ego,
                                                     double x, double v, double sel, double dt, double min_crash_distance,
                                                     EgoState &ego_out, double &nx, double &nv, double &na) {
    double px, py;
    if (ego.x < 1.5) {
        double dx = __dsub_rn(1.5, ego.x), dy = __dsub_rn(-1.5, ego.y);
        double nrm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        dx = __ddiv_rn(dx, nrm); dy = __ddiv_rn(dy, nrm);
        double step = __dmul_rn(sel, dt);
        px = __dadd_rn(ego.x, __dmul_rn(dx, step)); py = __dadd_rn(ego.y, __dmul_rn(dy, step));
        if (py < -1.6) py = -1.6;
    } else { py = ego.y; px = __dadd_rn(ego.x, __dmul_rn(sel, dt)); }
    ego_out.a = __ddiv_rn(__dsub_rn(sel, ego.v), dt);
    ego_out.x = px; ego_out.y = py; ego_out.v = sel;
    const bool can_crash = get_ego_s(px, py) > 11.0;
    const bool valid = lane < n;
    const double es = get_ego_s(ego.x, ego.y);
    const bool ego_on_lane = es >= 0.0;
    const double ex = ego.x >= 1.5 ? ego.x : es - 51.0;
    double lx = __shfl_up_sync(FULL, x, 1), lv = __shfl_up_sync(FULL, v, 1);
    if (lane == 0) { lx = 1.0e300; lv = 0.0; }
    if (ego_on_lane && ex > x && ex < lx) { lx = ex; lv = ego.v; }
    const double gap = lx - P.p.car_length - x - K.min_gap, bt = K.decel * K.tau;
    const double vsafe = gap <= 0.0 ? 0.0 : (gap > 1.0e200 ? 1.0e300 : -bt + sqrt(bt * bt + lv * lv + 2.0 * K.decel * gap));
    double v2 = fmin(fmin(v + K.accel * dt, vsafe), K.max_speed);
    v2 = fmax(v2, fmax(0.0, v - K.decel * dt));
    nv = v2; nx = x + v2 * dt; na = (v2 - v) / dt;
    const double cdd = min_crash_distance > P.p.car_length ? min_crash_distance : P.p.car_length;
    return __any_sync(FULL, valid && can_crash && fabs(nx - px) < cdd);
}

__global__ void __launch_bounds__(128) krauss_step_kernel(DevParams P, KraussParams K, int B, int nmax, double *ego, double *cars_x,
                                                          double *cars_v, double *cars_a, const int32_t *n_cars,
                                                          const double *sel_speed, double dt, double mcd, uint8_t *crashed) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const int b = warp;
    int n = n_cars[b]; n = n < 0 ? 0 : (n > nmax ? nmax : n);
    EgoState e = {ego[4 * b], ego[4 * b + 1], ego[4 * b + 2], ego[4 * b + 3]}, eo;
    const size_t o = (size_t)b * nmax + lane;
    const double x = lane < n ? cars_x[o] : 0.0, v = lane < n ? cars_v[o] : 0.0;
    double nx, nv, na;
    const bool cr = warp_krauss_with_ego(P, K, lane, n, e, x, v, sel_speed[b], dt, mcd, eo, nx, nv, na);
    __syncwarp();
    if (lane == 0) { ego[4 * b] = eo.x; ego[4 * b + 1] = eo.y; ego[4 * b + 2] = eo.v; ego[4 * b + 3] = eo.a; crashed[b] = cr ? 1 : 0; }
    if (lane < n) { cars_x[o] = nx; cars_v[o] = nv; cars_a[o] = na; }
}

// prediction.py:22-44: choose the pseudo-ego, then step.
__device__ __forceinline__ void warp_predict_without_ego(const DevParams &P, int lane, int n, EgoState &ego,
                                                         double &x, double &v, double dt) {
    double ego_s = get_ego_s(ego.x, ego.y);
    EgoState e = ego;
    double sel;
    double x0 = __shfl_sync(FULL, x, 0);
    if (ego_s < 8.0 || n == 0) { sel = 0.0; }                       // 26-27
    else if (x0 < ego.x) { e.x = -20.0; e.y = -10.0; e.v = 0.0; e.a = 0.0; sel = 0.0; }   // 28-31
    else {                                                          // 33-44
        unsigned behind = __ballot_sync(FULL, lane < n && x < ego.x);
        int first = behind ? __ffs(behind) - 1 : n;                 // first >= 1 here
        double lx = __shfl_sync(FULL, x, first - 1), lv = __shfl_sync(FULL, v, first - 1);
        sel = lv;
        if (behind) { e.x = __dsub_rn(__dsub_rn(lx, P.p.car_length), 5.0); e.v = lv; e.a = 0.0; }
    }
    EgoState out; double nx, nv, na;
    warp_predict_with_ego(P, lane, n, e, x, v, sel, dt, 5.0, out, nx, nv, na);
    ego = out; x = nx; v = nv;
}

// Merge intervals sorted by start (lane r reads sorted interval r of sb[0..nb)) into disjoint ones, overlapping or
// touching intervals joined; writes them back to sb[0..n) and returns n.  One warp.
__device__ __forceinline__ int merge_sorted_intervals(int lane, int nb, int2 *sb) {
    int2 mine = lane < nb ? sb[lane] : make_int2(INT_MAX, INT_MIN);
    int pm = lane < nb ? mine.y : INT_MIN;                                          // inclusive prefix max of interval ends
    for (int o = 1; o < 32; o <<= 1) { int x = __shfl_up_sync(FULL, pm, o); if (lane >= o) pm = max(pm, x); }
    int pm_excl = __shfl_up_sync(FULL, pm, 1);
    bool start = lane < nb && (lane == 0 || mine.x > pm_excl);
    unsigned sm = __ballot_sync(FULL, start);
    int gid = __popc(sm & ((2u << lane) - 1)) - 1;                                  // group of this interval
    unsigned below = sm & ((2u << lane) - 1);
    int first = below ? 31 - __clz(below) : 0;                                      // lane that opened my group
    int gstart = __shfl_sync(FULL, mine.x, first);
    bool last = lane < nb && (lane == nb - 1 || ((sm >> (lane + 1)) & 1u));
    __syncwarp();                                                                   // every lane has read sb[lane]
    if (last) sb[gid] = make_int2(gstart, pm);
    __syncwarp();
    return __popc(sm);
}

// ---- K1a -----------------------------------------------------------------------------------------
// desc[B][num_t]; hdr_s0/ds/num_s per problem.
// SUBSET: only the episodes listed in subset[0 .. *count) are predicted (masked plans; the default instance is unchanged)
template <bool SUBSET>
__global__ void __launch_bounds__(128, 7) predict_layers_kernel(DevParams P, int B, const double *__restrict__ ego,
                                                            const double *__restrict__ cars_x,
                                                            const double *__restrict__ cars_v,
                                                            const int32_t *__restrict__ n_cars, int nmax,
                                                            LayerDesc *__restrict__ desc, double *__restrict__ o_s0,
                                                            double *__restrict__ o_ds, int32_t *__restrict__ o_num_s,
                                                            const int32_t *__restrict__ subset, const int *__restrict__ count) {
    __shared__ double s_edge[4][2 * MPC_NMAX];
    __shared__ int2 s_band[4][MPC_NMAX], s_blk[4][MPC_NMAX];
    __shared__ int s_hist[4][2][MPC_MAX_BUCKETS];
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (warp >= B) return;
    int b = warp;
    if (SUBSET) { if (warp >= *count) return; b = subset[warp]; }
    int n = n_cars[b]; n = n < 0 ? 0 : (n > nmax ? nmax : n);
    EgoState e = {ego[4 * b], ego[4 * b + 1], ego[4 * b + 2], ego[4 * b + 3]};
    double x = lane < n ? cars_x[(size_t)b * nmax + lane] : 0.0;
    double v = lane < n ? cars_v[(size_t)b * nmax + lane] : 0.0;
    SGrid g = make_sgrid(P, e.x, e.y);
    if (lane == 0) { if (o_s0) o_s0[b] = g.s0; if (o_ds) o_ds[b] = g.ds; if (o_num_s) o_num_s[b] = g.num_s; }
    double s_last = g.sval(g.num_s - 1);
    double far_lim = __dadd_rn(s_last, P.p.car_length);
    for (int t = 0; t < P.num_t; t++) {
        if (t) warp_predict_without_ego(P, lane, n, e, x, v, P.p.t_disc);         // st.py:42-43
        double tt = __dmul_rn((double)t, P.p.t_disc);
        double unc = __dadd_rn(P.p.start_uncertainty, __dmul_rn(P.p.uncertainty_per_second, tt));   // 40
        int du = (int)__ddiv_rn(unc, P.p.s_disc);                                  // 41
        double obs_s = __dadd_rn(x, 51.0);                                          // control.py:388-389
        bool valid = lane < n;
        unsigned stop = __ballot_sync(FULL, valid && obs_s < P.obs_min_s);          // 46-47: break
        int first_stop = stop ? __ffs(stop) - 1 : 32;
        bool act = valid && lane < first_stop && !(obs_s > far_lim);                // 48-49: continue
        unsigned am = __ballot_sync(FULL, act);
        LayerDesc *L = desc + (size_t)b * P.num_t + t;
        int n_act = __popc(am);
        double ef = 0.0, eb = 0.0; int imin = 0, imax = 0;
        if (act) {
            int slot = __popc(am & ((1u << lane) - 1));
            ef = __dsub_rn(__dsub_rn(obs_s, P.p.car_length), unc);                  // 52
            eb = __dadd_rn(__dadd_rn(obs_s, P.p.car_length), unc);                  // 53
            int si = (int)__ddiv_rn(__dsub_rn(obs_s, g.s0), P.p.s_disc);            // 60
            imin = si - P.discrete_length - du; imin = imin < 0 ? 0 : imin;         // 61
            imax = si + P.discrete_length + du; imax = imax > g.num_s ? g.num_s : imax;   // 62
            if (!(imin < g.num_s && imax > 0)) { imin = 0; imax = 0; }               // 63
            L->ef[slot] = ef; L->eb[slot] = eb; L->band[slot] = make_int2(imin, imax);
        }
        // ---- blocked interval of this car: its band joined with the cells closer than MIN_ALLOWED_DISTANCE to its two
        //      distance-field edges, the tests being the reference's own fp64 expression |s_k - edge| < m (st.py:52-53,
        //      st_cy.pyx:34-38).  s_k is increasing in k, so each end is a monotone predicate: estimate, then fix up. ----
        int zlo = 0, zhi = 0;                                                        // [zlo, zhi)
        if (act) {
            const double m = P.p.min_allowed_distance;
            int k = (int)floor(__ddiv_rn(__dsub_rn(__dsub_rn(ef, m), g.s0), g.ds));
            k = k < 0 ? 0 : (k > g.num_s ? g.num_s : k);
            // first k with s_k >= ef or ef - s_k < m
            while (k > 0 && !(fabs(__dsub_rn(g.sval(k - 1), ef)) >= m && g.sval(k - 1) < ef)) k--;
            while (k < g.num_s && (fabs(__dsub_rn(g.sval(k), ef)) >= m && g.sval(k) < ef)) k++;
            zlo = k;
            k = (int)floor(__ddiv_rn(__dsub_rn(__dadd_rn(eb, m), g.s0), g.ds)) + 1;
            k = k < 0 ? 0 : (k > g.num_s ? g.num_s : k);
            // one past the last k with s_k <= eb or s_k - eb < m
            while (k < g.num_s && !(fabs(__dsub_rn(g.sval(k), eb)) >= m && g.sval(k) > eb)) k++;
            while (k > 0 && (fabs(__dsub_rn(g.sval(k - 1), eb)) >= m && g.sval(k - 1) > eb)) k--;
            zhi = k;
            if (imin < imax) { zlo = min(zlo, imin); zhi = max(zhi, imax); }
            if (zhi < zlo) zhi = zlo;
        }
        // ---- the same information sorted, for the fast kernel's O(1) lookups ----
        // rank of each edge among all 2*n_act edges (ties by edge id), of each band by start cell, of each blocked interval
        int r_f = 0, r_b = 0, r_band = 0, r_blk = 0;
        bool hasband = act && imin < imax, hasblk = act && zlo < zhi;
        unsigned bm = __ballot_sync(FULL, hasband), zm = __ballot_sync(FULL, hasblk);
        // Fast path.  The cars come front -> back, so with ordinary spacing (gap > 2 (CAR_LENGTH + uncertainty)) the ascending
        // order is simply "rearmost car's front edge, its back edge, the next car's front edge, ...": rank = 2 x (active cars
        // behind me) (+1), and bands / blocked intervals are ordered like the cars.  The guess is checked below on the sorted
        // arrays themselves; when it does not hold (perturbed states, large uncertainty) the ranks are counted.
        const unsigned above = lane == 31 ? 0u : ~((2u << lane) - 1u);                // lanes behind me
        r_f = 2 * __popc(am & above); r_b = r_f + 1;
        r_band = __popc(bm & above); r_blk = __popc(zm & above);
        {
            // sorted <=> every edge is <= the next car-in-front's front edge ... checked pairwise between neighbouring active cars
            const int nxt = (am & ((1u << lane) - 1u)) ? 31 - __clz(am & ((1u << lane) - 1u)) : -1;     // next active car in front of me
            const double nef = __shfl_sync(FULL, ef, nxt < 0 ? 0 : nxt);
            const int nimin = __shfl_sync(FULL, imin, nxt < 0 ? 0 : nxt), nzlo = __shfl_sync(FULL, zlo, nxt < 0 ? 0 : nxt);
            const int nband = __shfl_sync(FULL, (int)hasband, nxt < 0 ? 0 : nxt), nblk = __shfl_sync(FULL, (int)hasblk, nxt < 0 ? 0 : nxt);
            bool bad = act && (ef > eb);
            if (act && nxt >= 0) bad = bad || eb > nef || (hasband && nband && imin > nimin) || (hasblk && nblk && zlo > nzlo);
            // (a car without a band / interval between two that have one: compare with the next one that has; rare -> count)
            if (act && nxt >= 0 && ((hasband && !nband) || (hasblk && !nblk))) bad = true;
            if (__any_sync(FULL, bad)) {
                r_f = 0; r_b = 0; r_band = 0; r_blk = 0;
                for (int j = 0; j < 32; j++) {
                    if (!((am >> j) & 1u)) continue;                                        // warp-uniform
                    double fj = __shfl_sync(FULL, ef, j), bj = __shfl_sync(FULL, eb, j);
                    int ij = __shfl_sync(FULL, imin, j), zj = __shfl_sync(FULL, zlo, j);
                    r_f += (fj < ef || (fj == ef && 2 * j < 2 * lane)) + (bj < ef || (bj == ef && 2 * j + 1 < 2 * lane));
                    r_b += (fj < eb || (fj == eb && 2 * j < 2 * lane + 1)) + (bj < eb || (bj == eb && 2 * j + 1 < 2 * lane + 1));
                    if ((bm >> j) & 1u) r_band += (ij < imin || (ij == imin && j < lane));
                    if ((zm >> j) & 1u) r_blk += (zj < zlo || (zj == zlo && j < lane));
                }
            }
        }
        double *se = s_edge[wib]; int2 *sb = s_band[wib], *sz = s_blk[wib];
        if (act) { se[r_f] = ef; se[r_b] = eb; }
        if (hasband) sb[r_band] = make_int2(imin, imax);
        if (hasblk) sz[r_blk] = make_int2(zlo, zhi);
        __syncwarp();
        int n_edge = 2 * n_act;
        const int n_band = merge_sorted_intervals(lane, __popc(bm), sb);             // sb / sz now hold disjoint intervals
        const int n_blk = merge_sorted_intervals(lane, __popc(zm), sz);
        for (int i = lane; i < n_edge; i += 32) L->edge[i] = se[i];
        if (lane < n_band) L->mband[lane] = sb[lane];
        if (lane < n_blk) L->blk[lane] = sz[lane];
        if (lane == 0) { L->n_act = n_act; L->n_edge = n_edge; L->n_band = n_band; L->n_blk = n_blk; L->edge[n_edge] = 1e300; }
        __syncwarp();
        // bucket tables: bucket j starts at cell 64*j; bucket_edge[j] = #edges < s_(64 j), bucket_band[j] = #bands ending <= 64 j.
        // Every edge / band adds one to the first bucket it counts for (estimated by division, fixed up with the exact
        // comparison), an inclusive prefix sum over the buckets does the rest: ~130 warp instructions per layer where one
        // binary search per bucket took ~350.
        int nbuck = (g.num_s + (1 << MPC_BUCKET_SHIFT) - 1) >> MPC_BUCKET_SHIFT;
        int *he = s_hist[wib][0], *hb = s_hist[wib][1];
        for (int j = lane; j < nbuck; j += 32) { he[j] = 0; hb[j] = 0; }
        __syncwarp();
        for (int i = lane; i < n_edge; i += 32) {
            const double ev = se[i];
            int j = (int)floor(__ddiv_rn(__dsub_rn(ev, g.s0), __dmul_rn(g.ds, 64.0))) + 1;
            j = j < 0 ? 0 : (j > nbuck ? nbuck : j);
            while (j > 0 && ev < g.sval((j - 1) << MPC_BUCKET_SHIFT)) j--;            // first j with ev < s_(64 j)
            while (j < nbuck && !(ev < g.sval(j << MPC_BUCKET_SHIFT))) j++;
            if (j < nbuck) atomicAdd(&he[j], 1);
        }
        if (lane < n_band) { const int j = (sb[lane].y + (1 << MPC_BUCKET_SHIFT) - 1) >> MPC_BUCKET_SHIFT; if (j < nbuck) atomicAdd(&hb[j], 1); }
        __syncwarp();
        {
            const int per = (nbuck + 31) >> 5, j0 = lane * per;                       // consecutive buckets per lane
            int se_ = 0, sb_ = 0;
            for (int q = 0; q < per; q++) if (j0 + q < nbuck) { se_ += he[j0 + q]; sb_ += hb[j0 + q]; }
            int pe = se_, pb = sb_;
            for (int o = 1; o < 32; o <<= 1) { int a_ = __shfl_up_sync(FULL, pe, o), b_ = __shfl_up_sync(FULL, pb, o); if (lane >= o) { pe += a_; pb += b_; } }
            pe -= se_; pb -= sb_;                                                     // exclusive
            for (int q = 0; q < per; q++) if (j0 + q < nbuck) {
                pe += he[j0 + q]; pb += hb[j0 + q];
                L->bucket_edge[j0 + q] = (unsigned char)pe; L->bucket_band[j0 + q] = (unsigned char)pb;
            }
        }
        __syncwarp();
    }
}

// ---- K1b: dense grids in the layout st_cy consumes ---------------------------------------------
// One block per (episode, layer); the layer's sorted search structure is staged in shared memory, every thread
// evaluates 4 consecutive cells (O(1) lookups) and writes them with one 32-bit mask store and one/two 128-bit distance
// stores.  This is the one kernel of the path that is bound by HBM writes (1 + 8 or 1 + 4 bytes per cell).
template <typename DT>
__global__ void __launch_bounds__(256) rasterise_kernel(DevParams P, int B, int stride_s, const LayerDesc *__restrict__ desc,
                                                        const double *__restrict__ s0v, const double *__restrict__ dsv,
                                                        const int32_t *__restrict__ nsv, uint8_t *__restrict__ obstacles,
                                                        DT *__restrict__ distances, int vec_ok) {
    __shared__ __align__(16) LayerSearch L;
    const int bt = blockIdx.x, b = bt / P.num_t;
    {
        const char *src = reinterpret_cast<const char *>(desc + bt);
        constexpr int kTail = (int)(sizeof(LayerSearch) - offsetof(LayerSearch, edge)) / 16;
        if (threadIdx.x < kTail) reinterpret_cast<int4 *>(reinterpret_cast<char *>(&L) + offsetof(LayerSearch, edge))[threadIdx.x] =
            reinterpret_cast<const int4 *>(src + offsetof(LayerDesc, edge))[threadIdx.x];
        else if (threadIdx.x == kTail) *reinterpret_cast<int4 *>(&L) = make_int4(desc[bt].n_edge, desc[bt].n_band, 0, 0);
    }
    __syncthreads();
    SGrid g; g.s0 = s0v[b]; g.ds = dsv[b]; g.num_s = nsv[b];
    const size_t row = (size_t)bt * stride_s;
    const int M = L.n_edge, m = L.n_band;
    // 8 consecutive cells per thread: the bucket is looked up once, the two distance-field edges that bracket the cell and the
    // current band stay in registers and are advanced only when a cell passes them (edges are >= 10 m = 200 cells apart), the cell
    // index becomes a double with one add (2^52 trick) instead of a conversion; one 64-bit mask store, two 128-bit distance stores.
    for (int k8 = threadIdx.x * 8; k8 < stride_s; k8 += blockDim.x * 8) {
        const int jb = min(k8, g.num_s - 1) >> MPC_BUCKET_SHIFT;
        int e = L.bucket_edge[jb], i = L.bucket_band[jb];
        {   // position the cursors on the first cell (at most a few steps: the bucket starts <= 63 cells earlier)
            const double sv0 = g.sval(min(k8, g.num_s - 1));
            while (e < M && L.edge[e] < sv0) e++;
            while (i < m && L.mband[i].y <= k8) i++;
        }
        double lo = e > 0 ? L.edge[e - 1] : -1.0e300, hi = e < M ? L.edge[e] : 1.0e300;
        int2 band = i < m ? L.mband[i] : make_int2(INT_MAX, INT_MAX);
        unsigned ob_lo = 0, ob_hi = 0;
        DT d8[8];
        // Straight-line path: at most ONE edge and ONE band boundary inside the 8 cells (edges and bands of a layer are metres
        // apart; 8 cells are 0.4 m) -- the bracketing pair is then a select between two register pairs, no loop, no branch.
        const double hi2 = e + 1 < M ? L.edge[e + 1] : 1.0e300;
        const int2 band2 = i + 1 < m ? L.mband[i + 1] : make_int2(INT_MAX, INT_MAX);
        const int klast = min(k8 + 7, g.num_s - 1);
        if (M > 0 && !(hi2 < g.sval(klast)) && band2.y > klast) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const int k = k8 + c;
                const double kd = __dsub_rn(__hiloint2double(0x43300000, k), 4503599627370496.0);      // (double)k, exactly
                const double sv = __dadd_rn(g.s0, __dmul_rn(kd, g.ds));
                const bool past = hi < sv;
                const double lo_c = past ? hi : lo, hi_c = past ? hi2 : hi;
                const double dl = __dsub_rn(sv, lo_c), dr = __dsub_rn(hi_c, sv);        // (== fabs(sv - hi_c): hi_c >= sv)
                double d = dl < dr ? dl : dr;
                const int2 bc = band.y <= k ? band2 : band;
                bool ob = bc.x <= k;                                                     // (k < bc.y holds: bc is the first band ending behind k)
                if (k >= g.num_s) { ob = true; }
                if (ob) d = 0.0;
                if (c < 4) ob_lo |= (ob ? 1u : 0u) << (8 * c); else ob_hi |= (ob ? 1u : 0u) << (8 * (c - 4));
                d8[c] = (DT)d;
            }
        } else {
#pragma unroll 1
            for (int c = 0; c < 8; c++) {
                const int k = k8 + c;
                double d = 0.0; bool ob = true;
                if (k < g.num_s) {
                    const double sv = g.sval(k);
                    while (hi < sv) { e++; lo = hi; hi = e < M ? L.edge[e] : 1.0e300; }
                    const double dl = __dsub_rn(sv, lo), dr = fabs(__dsub_rn(sv, hi));
                    d = 1E10; d = dl < d ? dl : d; d = dr < d ? dr : d;
                    while (band.y <= k) { i++; band = i < m ? L.mband[i] : make_int2(INT_MAX, INT_MAX); }
                    ob = band.x <= k;
                    if (ob) d = 0.0;
                }
                if (c < 4) ob_lo |= (ob ? 1u : 0u) << (8 * c); else ob_hi |= (ob ? 1u : 0u) << (8 * (c - 4));
                d8[c] = (DT)d;
            }
        }
        if (vec_ok && k8 + 7 < stride_s) {
            *reinterpret_cast<uint2 *>(obstacles + row + k8) = make_uint2(ob_lo, ob_hi);
            if (sizeof(DT) == 4) {
                float4 *q = reinterpret_cast<float4 *>(distances + row + k8);
                q[0] = make_float4((float)d8[0], (float)d8[1], (float)d8[2], (float)d8[3]);
                q[1] = make_float4((float)d8[4], (float)d8[5], (float)d8[6], (float)d8[7]);
            } else {
                double2 *q = reinterpret_cast<double2 *>(distances + row + k8);
                q[0] = make_double2((double)d8[0], (double)d8[1]); q[1] = make_double2((double)d8[2], (double)d8[3]);
                q[2] = make_double2((double)d8[4], (double)d8[5]); q[3] = make_double2((double)d8[6], (double)d8[7]);
            }
        } else {
            for (int c = 0; c < 8 && k8 + c < stride_s; c++) {
                obstacles[row + k8 + c] = (unsigned char)(((c < 4 ? ob_lo : ob_hi) >> (8 * (c & 3))) & 1u);
                distances[row + k8 + c] = d8[c];
            }
        }
    }
}

// ---- K1b, segment version (round 2).  The <= 64 distance-field edges cut the row into segments with ONE bracketing pair (lo, hi)
// each; the first cell of every segment is found once per edge (estimate by division, fix up with the reference's own
// comparison).  A warp owns a contiguous run of 32-cell chunks (lane = cell) and walks it with warp-uniform cursors: up to the next
// EVENT (a segment boundary, a band boundary, the end of the episode's grid) every chunk has the same (lo, hi) and the same
// blocked / free state for all lanes, and a free cell costs: s = s0 + k ds, two differences, a min, one conversion, two coalesced
// stores; a blocked chunk is two stores.  Only the chunk an event falls into takes the per-lane path (cursor loops), after which
// the uniform state is derived again.  Bit-identical to rasterise_kernel and to the oracle (tests).
// (A variant that assembled the row in shared memory and wrote it with cp.async.bulk -- UBLKCP -- was measured 40 % slower than
// the old kernel: assembly and store serialise inside a block.  The first segment version interleaved the chunks of a row over
// the warps and re-derived the cursors for every chunk: 93 instructions per cell against 44, ncu capture of round 2.)
template <typename DT>
__global__ void __launch_bounds__(256) rasterise_rows_kernel(DevParams P, int B, int stride_s, const LayerDesc *__restrict__ desc,
                                                             const double *__restrict__ s0v, const double *__restrict__ dsv,
                                                             const int32_t *__restrict__ nsv, uint8_t *__restrict__ obstacles,
                                                             DT *__restrict__ distances) {
    __shared__ __align__(16) LayerSearch L;
    __shared__ int kx[2 * MPC_NMAX + 4];                                             // kx[j] = first cell k with edge[j] < s_k
    const int bt = blockIdx.x, b = bt / P.num_t, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    SGrid g; g.s0 = s0v[b]; g.ds = dsv[b]; g.num_s = nsv[b];                         // (in flight together with the descriptor)
    {
        const char *src = reinterpret_cast<const char *>(desc + bt);
        constexpr int kTail = (int)(offsetof(LayerSearch, bucket_edge) - offsetof(LayerSearch, edge)) / 16;    // edges + bands
        static_assert((offsetof(LayerSearch, bucket_edge) - offsetof(LayerSearch, edge)) % 16 == 0, "copied as int4");
        for (int t = tid; t < kTail; t += blockDim.x)
            reinterpret_cast<int4 *>(reinterpret_cast<char *>(&L) + offsetof(LayerSearch, edge))[t] =
                reinterpret_cast<const int4 *>(src + offsetof(LayerDesc, edge))[t];
        if (tid == blockDim.x - 1) { *reinterpret_cast<int4 *>(&L) = make_int4(desc[bt].n_edge, desc[bt].n_band, 0, 0); L.edge_lo = -1.0e300; }
    }
    __syncthreads();
    const int M = L.n_edge, m = L.n_band, ns = g.num_s;
    for (int j = tid; j < M; j += blockDim.x) {
        const double ev = L.edge[j];
        int k = (int)floor(__ddiv_rn(__dsub_rn(ev, g.s0), g.ds)) + 1;
        k = k < 0 ? 0 : (k > ns ? ns : k);
        while (k > 0 && ev < g.sval(k - 1)) k--;
        while (k < ns && !(ev < g.sval(k))) k++;
        kx[j] = k;
    }
    if (tid == 0) kx[M] = INT_MAX;
    __syncthreads();
    const double *E = &L.edge_lo;                                                    // E[i + 1] = edge[i]; E[0] = -1e300, edge[M] = +1e300
    const int chunks = (stride_s + 31) >> 5, per_warp = (chunks + nwarps - 1) / nwarps;
    const int full = stride_s & ~31;                                                 // chunks from here on are partial: per-lane path
    int kb = warp * per_warp * 32;
    const int kend = min(kb + per_warp * 32, stride_s);
    if (kb >= kend) return;
    DT *dp = distances + (size_t)bt * stride_s + kb + lane;                          // this lane's cell of the current chunk
    uint8_t *op = obstacles + (size_t)bt * stride_s + kb + lane;
    double kd = __dsub_rn(__hiloint2double(0x43300000, kb + lane), 4503599627370496.0);         // (double)k, exactly (k < 2^31)
    const double s0 = g.s0, ds = g.ds;
    const int lim0 = min(full, kend);                                                // whole chunks of this warp end here
    // (uniform) segment / band cursors of the warp at its first cell: counted by the lanes (kx and the band ends ascend)
    int j = __popc(__ballot_sync(0xffffffffu, lane < M && kx[lane] <= kb)) + __popc(__ballot_sync(0xffffffffu, lane + 32 < M && kx[lane + 32] <= kb));
    int i = __popc(__ballot_sync(0xffffffffu, lane < m && L.mband[lane].y <= kb));
    while (kb < kend) {
        // uniform state at cell kb: bracketing pair, blocked or free, and the next EVENT (segment / band boundary, end of the grid)
        while (kx[j] <= kb) j++;
        while (i < m && L.mband[i].y <= kb) i++;
        const int2 bd = i < m ? L.mband[i] : make_int2(INT_MAX, INT_MAX);
        const bool blocked = bd.x <= kb || kb >= ns;
        const int ev = min(kx[j], kb >= ns ? INT_MAX : (bd.x <= kb ? bd.y : min(bd.x, ns)));   // > kb
        int r = (min(ev, lim0) - kb) >> 5;                                           // whole chunks before it: no per-lane decisions
        kb += r << 5;
        if (!blocked && M > 0) {
            const double lo = E[j], hi = E[j + 1];
            for (; r >= 2; r -= 2, dp += 64, op += 64, kd = __dadd_rn(kd, 64.0)) {
                const double sa = __dadd_rn(s0, __dmul_rn(kd, ds)), sb = __dadd_rn(s0, __dmul_rn(__dadd_rn(kd, 32.0), ds));
                const double la = __dsub_rn(sa, lo), ra = __dsub_rn(hi, sa), lb = __dsub_rn(sb, lo), rb = __dsub_rn(hi, sb);
                dp[0] = (DT)(la < ra ? la : ra); dp[32] = (DT)(lb < rb ? lb : rb);
                op[0] = 0; op[32] = 0;
            }
            if (r) {
                const double sa = __dadd_rn(s0, __dmul_rn(kd, ds));
                const double la = __dsub_rn(sa, lo), ra = __dsub_rn(hi, sa);
                dp[0] = (DT)(la < ra ? la : ra); op[0] = 0;
                dp += 32; op += 32; kd = __dadd_rn(kd, 32.0);
            }
        } else {
            const DT dv = blocked ? (DT)0.0 : (DT)1E10;
            const uint8_t ov = blocked ? 1 : 0;
            kd = __dadd_rn(kd, (double)(r << 5));
            for (; r > 0; r--, dp += 32, op += 32) { dp[0] = dv; op[0] = ov; }
        }
        if (kb >= kend) break;
        if (kb == ev) continue;                                                      // the event sits on a chunk boundary
        {                                                                            // the chunk with the event (or the row's partial
            const int k = kb + lane;                                                 // last chunk): per-lane cursors
            const double sv = __dadd_rn(s0, __dmul_rn(kd, ds));
            int jj = j;
            while (kx[jj] <= k) jj++;
            const double dl = __dsub_rn(sv, E[jj]), dr = __dsub_rn(E[jj + 1], sv);
            double d = dl < dr ? dl : dr;
            if (M == 0) d = 1E10;
            int ii = i;
            while (ii < m && L.mband[ii].y <= k) ii++;
            bool ob = ii < m && L.mband[ii].x <= k;
            if (k >= ns) ob = true;
            if (ob) d = 0.0;
            if (k < stride_s) { dp[0] = (DT)d; op[0] = ob ? 1 : 0; }
            kb += 32; dp += 32; op += 32; kd = __dadd_rn(kd, 32.0);
        }
    }
}

// ---- self-test: the sorted search structure must reproduce the reference-order evaluation bit for bit ----
__global__ void __launch_bounds__(256) check_sorted_kernel(DevParams P, int B, const LayerDesc *__restrict__ desc,
                                                           const double *__restrict__ s0v, const double *__restrict__ dsv,
                                                           const int32_t *__restrict__ nsv, unsigned long long *mismatches) {
    int bt = blockIdx.x, b = bt / P.num_t;
    const LayerDesc &L = desc[bt];
    SGrid g; g.s0 = s0v[b]; g.ds = dsv[b]; g.num_s = nsv[b];
    unsigned long long bad = 0;
    for (int k = threadIdx.x; k < g.num_s; k += blockDim.x) {
        bool o1, o2;
        double s = g.sval(k);
        double d1 = cell_distance(L, s, k, o1), d2 = cell_distance_sorted(L, s, k, o2);
        if (o1 != o2 || d1 != d2) bad++;
        // blocked intervals (lean bounded pass) == in a band, or closer than MIN_ALLOWED_DISTANCE to a distance-field edge
        double dn = 1E10;
        for (int c = 0; c < L.n_act; c++) { dn = fmin(dn, fabs(__dsub_rn(s, L.ef[c]))); dn = fmin(dn, fabs(__dsub_rn(s, L.eb[c]))); }
        bool blocked = false;
        for (int i = 0; i < L.n_blk; i++) blocked = blocked || (k >= L.blk[i].x && k < L.blk[i].y);
        if (P.zone_ok && blocked != (o1 || dn < P.p.min_allowed_distance)) bad++;
        if (k == 0 && L.edge[L.n_edge] != 1e300) bad++;                     // upper sentinel of the sorted edge list
    }
    if (bad) atomicAdd(mismatches, bad);
}

cudaError_t launch_check_sorted(const DevParams &P, int B, const LayerDesc *desc, const double *s0, const double *ds,
                                const int32_t *ns, unsigned long long *mismatches, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    MPC_LAUNCH(check_sorted_kernel, B * P.num_t, 256, 0, st, P, B, desc, s0, ds, ns, mismatches);
    return cudaGetLastError();
}

// ---- K4 -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) predict_step_with_ego_kernel(DevParams P, int B, int nmax, const double *ego,
                                                                   const double *cars_x, const double *cars_v,
                                                                   const double *cars_a, const int32_t *n_cars,
                                                                   const double *sel_speed, double dt, double mcd,
                                                                   double *ego_out, double *ox, double *ov, double *oa,
                                                                   uint8_t *crashed) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    int b = warp;
    int n = n_cars[b]; n = n < 0 ? 0 : (n > nmax ? nmax : n);
    EgoState e = {ego[4 * b], ego[4 * b + 1], ego[4 * b + 2], ego[4 * b + 3]}, eo;
    double x = lane < n ? cars_x[(size_t)b * nmax + lane] : 0.0;
    double v = lane < n ? cars_v[(size_t)b * nmax + lane] : 0.0;
    double nx, nv, na;
    bool cr = warp_predict_with_ego(P, lane, n, e, x, v, sel_speed[b], dt, mcd, eo, nx, nv, na);
    if (lane == 0) {
        ego_out[4 * b] = eo.x; ego_out[4 * b + 1] = eo.y; ego_out[4 * b + 2] = eo.v; ego_out[4 * b + 3] = eo.a;
        if (crashed) crashed[b] = cr ? 1 : 0;
    }
    if (lane < nmax) {
        size_t o = (size_t)b * nmax + lane;
        bool ok = lane < n;
        ox[o] = ok ? nx : (cars_x == ox ? cars_x[o] : 0.0);
        ov[o] = ok ? nv : (cars_v == ov ? cars_v[o] : 0.0);
        if (oa) oa[o] = ok ? na : ((cars_a && cars_a == oa) ? cars_a[o] : 0.0);
    }
}

// HighwayState.predict_step_without_ego (prediction.py:22-44) as a public step: the pseudo-ego choice of
// warp_predict_without_ego, but with the caller's min_crash_distance and with everything predict_step_with_ego returns
// (new ego incl. its acceleration, car accelerations, crash flag).
__global__ void __launch_bounds__(128) predict_step_without_ego_kernel(DevParams P, int B, int nmax, const double *ego,
                                                                      const double *cars_x, const double *cars_v,
                                                                      const double *cars_a, const int32_t *n_cars, double dt,
                                                                      double mcd, double *ego_out, double *ox, double *ov,
                                                                      double *oa, uint8_t *crashed) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    int b = warp;
    int n = n_cars[b]; n = n < 0 ? 0 : (n > nmax ? nmax : n);
    EgoState e = {ego[4 * b], ego[4 * b + 1], ego[4 * b + 2], ego[4 * b + 3]}, eo;
    double x = lane < n ? cars_x[(size_t)b * nmax + lane] : 0.0;
    double v = lane < n ? cars_v[(size_t)b * nmax + lane] : 0.0;
    const double ego_s = get_ego_s(e.x, e.y);
    double sel;
    const double x0 = __shfl_sync(FULL, x, 0);
    if (ego_s < 8.0 || n == 0) { sel = 0.0; }                                                 // 26-27: the ego stays put
    else if (x0 < e.x) { e.x = -20.0; e.y = -10.0; e.v = 0.0; e.a = 0.0; sel = 0.0; }         // 28-31: ego in front of everybody
    else {                                                                                    // 33-44: ego follows the car ahead
        const unsigned behind = __ballot_sync(FULL, lane < n && x < e.x);
        const int first = behind ? __ffs(behind) - 1 : n;                                     // >= 1 here
        const double lx = __shfl_sync(FULL, x, first - 1), lv = __shfl_sync(FULL, v, first - 1);
        sel = lv;
        if (behind) { e.x = __dsub_rn(__dsub_rn(lx, P.p.car_length), 5.0); e.v = lv; e.a = 0.0; }
    }
    double nx, nv, na;
    bool cr = warp_predict_with_ego(P, lane, n, e, x, v, sel, dt, mcd, eo, nx, nv, na);
    if (lane == 0) {
        ego_out[4 * b] = eo.x; ego_out[4 * b + 1] = eo.y; ego_out[4 * b + 2] = eo.v; ego_out[4 * b + 3] = eo.a;
        if (crashed) crashed[b] = cr ? 1 : 0;
    }
    if (lane < nmax) {
        size_t o = (size_t)b * nmax + lane;
        bool ok = lane < n;
        ox[o] = ok ? nx : (cars_x == ox ? cars_x[o] : 0.0);
        ov[o] = ok ? nv : (cars_v == ov ? cars_v[o] : 0.0);
        if (oa) oa[o] = ok ? na : ((cars_a && cars_a == oa) ? cars_a[o] : 0.0);
    }
}

// dqn.py:389-446 with CARS_AHEAD = CARS_BEHIND = 2, acceleration + speed difference, normalised.
__global__ void __launch_bounds__(128) state_vector_kernel(DevParams P, int B, int nmax, const double *ego,
                                                          const double *cars_x, const double *cars_v,
                                                          const double *cars_a, const int32_t *n_cars, float *out,
                                                          int out_stride) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    int b = warp;
    int n = n_cars[b]; n = n < 0 ? 0 : (n > nmax ? nmax : n);
    double ex = ego[4 * b], ey = ego[4 * b + 1], ev = ego[4 * b + 2], ea = ego[4 * b + 3];
    bool valid = lane < n;
    size_t o = (size_t)b * nmax + lane;
    double x = valid ? cars_x[o] : 0.0, v = valid ? cars_v[o] : 0.0, a = valid ? cars_a[o] : 0.0;
    unsigned front = __ballot_sync(FULL, valid && x > ex);          // 411-414
    unsigned back = __ballot_sync(FULL, valid && !(x > ex));
    // front list is reversed (nearest first): the highest set lane is slot 0, next highest slot 1
    int slot = -1;
    if (valid && x > ex) { int above = __popc(front >> lane) - 1; if (above < 2) slot = above; }
    if (valid && !(x > ex)) { int below = __popc(back & ((1u << lane) - 1)); if (below < 2) slot = 2 + below; }
    float *row = out + (size_t)b * out_stride;
    if (slot >= 0) {
        row[4 * slot + 0] = (float)__ddiv_rn(a, 9.0);
        row[4 * slot + 1] = (float)__ddiv_rn(__dsub_rn(v, ev), P.p.max_speed);
        row[4 * slot + 2] = (float)__ddiv_rn(__dsub_rn(x, ex), P.p.sensor_radius);
        row[4 * slot + 3] = 1.0f;
    }
    int nf = __popc(front), nb = __popc(back);
    if (lane < 4) {                       // zero the absent slots
        int s = lane; bool present = s < 2 ? (s < nf) : (s - 2 < nb);
        if (!present) { row[4 * s] = 0.f; row[4 * s + 1] = 0.f; row[4 * s + 2] = 0.f; row[4 * s + 3] = 0.f; }
    }
    if (lane == 0) {
        row[16] = (float)__ddiv_rn(ev, P.p.max_speed); row[17] = (float)__ddiv_rn(ea, 9.0);
        row[18] = (float)__ddiv_rn(ex, 300.0); row[19] = (float)__ddiv_rn(ey, 100.0);
    }
}

// control.py:160-171
__global__ void speed_from_jerk_kernel(DevParams P, int B, const double *ego, const double *jerk, double *speed) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double v = ego[4 * b + 2], a = ego[4 * b + 3];
    double na = __dadd_rn(a, __dmul_rn(jerk[b], P.p.tick_length));
    if (na > P.p.a_max) na = P.p.a_max;
    if (na < P.p.a_min) na = P.p.a_min;
    double nv = __dadd_rn(v, __dmul_rn(na, P.p.tick_length));
    if (nv > P.p.max_speed) nv = P.p.max_speed;
    if (nv < 0.0) nv = 0.0;
    speed[b] = nv;
}

// One step of the combined controller's policy rollout (dqn.py:129-141), masked and in place: jerk -> speed
// (control.py:160-171), predict_step_with_ego, and the per-episode bookkeeping the loop carries.
__global__ void __launch_bounds__(128) rollout_step_kernel(DevParams P, int B, int nmax, double *ego, double *cars_x,
                                                          double *cars_v, double *cars_a, const int32_t *n_cars,
                                                          const double *jerk, double dt, double mcd, double stop_x,
                                                          int step, uint8_t *alive, double *sel_speed, double *roll_s,
                                                          int roll_stride, int32_t *roll_len, uint8_t *crash_pred) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    int b = warp;
    double *rs = roll_s + (size_t)b * roll_stride;
    if (!alive[b]) {                                                 // the episode's rollout ended earlier (dqn.py:140-141)
        if (lane == 0) rs[step] = rs[step - 1];
        return;
    }
    int n = n_cars[b]; n = n < 0 ? 0 : (n > nmax ? nmax : n);
    EgoState e = {ego[4 * b], ego[4 * b + 1], ego[4 * b + 2], ego[4 * b + 3]}, eo;
    double na = __dadd_rn(e.a, __dmul_rn(jerk[b], P.p.tick_length));    // control.py:160-171
    na = na > P.p.a_max ? P.p.a_max : na; na = na < P.p.a_min ? P.p.a_min : na;
    double sel = __dadd_rn(e.v, __dmul_rn(na, P.p.tick_length));
    sel = sel > P.p.max_speed ? P.p.max_speed : sel; sel = sel < 0.0 ? 0.0 : sel;
    size_t o = (size_t)b * nmax + lane;
    double x = lane < n ? cars_x[o] : 0.0, v = lane < n ? cars_v[o] : 0.0, nx, nv, nacc;
    bool cr = warp_predict_with_ego(P, lane, n, e, x, v, sel, dt, mcd, eo, nx, nv, nacc);
    if (lane < n) { cars_x[o] = nx; cars_v[o] = nv; cars_a[o] = nacc; }
    if (lane == 0) {
        ego[4 * b] = eo.x; ego[4 * b + 1] = eo.y; ego[4 * b + 2] = eo.v; ego[4 * b + 3] = eo.a;
        sel_speed[b] = sel;
        rs[step] = get_ego_s(eo.x, eo.y);                            // rollout_s_history (dqn.py:139)
        roll_len[b] += 1;
        if (cr) crash_pred[b] = 1;
        if (cr || eo.x > stop_x) alive[b] = 0;
    }
}

// ---- host launchers (called from mpc_api.cu) -----------------------------------------------------
// ---- fused environment tick (mpc_env_step): see include/mpcb200.h.  One warp per episode, lane == car slot; registers and
// shuffles only.  Every fp64 operation is written in the order the tensor expressions of merge_gym.MergeEnv.step evaluate
// them, so the fused tick is bit-identical to the unfused one given the same random numbers. ----
__global__ void __launch_bounds__(128) env_step_kernel(DevParams P, mpc_env_params E, int B, int nmax, double *ego, double *cars_x,
                                                      double *cars_v, double *cars_a, int32_t *n_cars, double *prev_acc,
                                                      double *delay, int32_t *ticks, const double *jerk, const double *u_spawn,
                                                      const double *gap_u, const double *first_u, const double *speed_z,
                                                      const double *delay_u, double *reward, uint8_t *flags, double *proj_jerk) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const int b = warp;
    int n = n_cars[b]; n = n < 0 ? 0 : (n > nmax ? nmax : n);
    EgoState e = {ego[4 * b], ego[4 * b + 1], ego[4 * b + 2], ego[4 * b + 3]}, eo;
    const size_t o = (size_t)b * nmax + lane;
    const bool slot_ok = lane < nmax;
    double x = slot_ok ? cars_x[o] : 0.0, v = slot_ok ? cars_v[o] : 0.0, a = slot_ok ? cars_a[o] : 0.0;
    // merge_gym.py:83-96: clip the projected acceleration / speed, remember the realised jerk
    // every per-episode scalar is read here, before the first collective: lane 0 overwrites them at the end of the kernel
    const double pa = prev_acc[b], delay_in = delay[b];
    const int ticks_in = ticks[b];
    // _handle_jerk (merge_gym.py:83-96): the projected acceleration is clipped, ELSE the projected speed (computed with the
    // unclipped acceleration) is clipped and the acceleration follows from it; either way the action counts as invalid
    double acc_p = __dadd_rn(pa, __dmul_rn(jerk[b], E.tick));
    double spd_p = __dadd_rn(e.v, __dmul_rn(acc_p, E.tick));
    const double acc_c = acc_p < E.a_min ? E.a_min : (acc_p > E.a_max ? E.a_max : acc_p);
    bool invalid = false;
    if (acc_p > E.a_max || acc_p < E.a_min) { invalid = true; acc_p = acc_c; }
    else if (spd_p > E.max_speed || spd_p < 0.0) {
        invalid = true;
        spd_p = spd_p < 0.0 ? 0.0 : E.max_speed;
        acc_p = __ddiv_rn(__dsub_rn(spd_p, e.v), E.tick);
    }
    const double pj = __ddiv_rn(__dsub_rn(acc_p, pa), E.tick);
    // the command itself: control.set_ego_jerk -> get_ego_speed_from_jerk (control.py:160-176), both clamps in sequence
    double spd = __dadd_rn(e.v, __dmul_rn(acc_c, E.tick));
    spd = spd > E.max_speed ? E.max_speed : (spd < 0.0 ? 0.0 : spd);
    // world step: the reference predictor as dynamics (cars beyond n keep their values, like the in-place K4 call)
    double nx, nv, na;
    const KraussParams KP = {E.krauss_accel, E.krauss_decel, E.krauss_tau, E.krauss_min_gap, E.other_speed};
    const bool crashed = E.world == 1 ? warp_krauss_with_ego(P, KP, lane, n, e, x, v, spd, E.tick, E.min_crash_distance, eo, nx, nv, na)
                                      : warp_predict_with_ego(P, lane, n, e, x, v, spd, E.tick, E.min_crash_distance, eo, nx, nv, na);
    if (lane < n) { x = nx; v = nv; a = na; }
    double new_prev_acc = eo.a;
    // recycle the front car once it is out of sensor range ahead (all slots shift), enter a new car at the back
    const double x0 = __shfl_sync(FULL, x, 0);
    const bool gone = n > 0 && __dsub_rn(x0, eo.x) > E.sensor_radius;
    {
        const double sx = __shfl_down_sync(FULL, x, 1), sv = __shfl_down_sync(FULL, v, 1), sa = __shfl_down_sync(FULL, a, 1);
        if (gone) { const bool last = lane >= nmax - 1; x = last ? 0.0 : sx; v = last ? 0.0 : sv; a = last ? 0.0 : sa; }
    }
    n -= gone ? 1 : 0;
    double dl = __dsub_rn(delay_in, E.tick);
    const bool spawn = dl <= 0.0 && n < nmax;
    if (spawn && lane == n) { x = E.spawn_x; v = E.other_speed; a = 0.0; }
    n += spawn ? 1 : 0;
    if (spawn) dl = __dadd_rn(u_spawn ? u_spawn[b] : 0.0, E.interval);
    int tk = ticks_in + 1;
    const bool arrived = eo.x > E.arrival_x && !crashed;
    const bool timeout = tk >= E.max_ticks && !crashed && !arrived;
    const bool done = crashed || arrived || timeout;
    // merge_gym.py:102-140 + dqn.py:557-563: a tick that ends in a crash / an arrival pays the terminal reward, every other tick
    // the time + jerk reward of the MEASURED jerk (new acceleration against the previous one); a clipped action adds its penalty
    const double jm = __ddiv_rn(__dsub_rn(eo.a, pa), E.tick);
    double r = __dsub_rn(E.time_reward_step, __dmul_rn(__dmul_rn(E.jerk_weight, __dmul_rn(jm, jm)), E.tick));
    if (arrived) r = E.success_reward;
    if (crashed) r = E.crash_reward;
    if (invalid) r = __dadd_rn(r, E.invalid_action_step);
    if (E.auto_reset && done) {
        // fresh initial conditions: spawner-spaced traffic (control.py:215-226), ego at the ramp start (control.py:41-44, 198-204)
        double g = slot_ok ? __dmul_rn(E.other_speed, __dadd_rn(E.interval, gap_u[o])) : 0.0;
        if (lane == 0) g = __dmul_rn(__dmul_rn(first_u[b], E.other_speed), __dadd_rn(E.interval, 0.5));
        double cs = 0.0;                                       // cumulative sum, left to right like a sequential scan
        for (int j = 0; j < 32; j++) { const double gj = __shfl_sync(FULL, g, j); if (j <= lane) cs = j == 0 ? gj : __dadd_rn(cs, gj); }
        const double xs = __dsub_rn(__dadd_rn(E.ego_start_x, E.sensor_radius), cs);
        const bool keep = slot_ok && xs >= E.spawn_x;
        n = __popc(__ballot_sync(FULL, keep));
        x = keep ? xs : 0.0; v = keep ? E.other_speed : 0.0; a = 0.0;
        double v0 = E.start_speed;
        if (speed_z) { v0 = __dadd_rn(E.start_speed, __dmul_rn(E.start_speed_var, speed_z[b])); v0 = v0 < E.min_start_speed ? E.min_start_speed : (v0 > E.max_start_speed ? E.max_start_speed : v0); }
        eo.x = E.ego_start_x; eo.y = E.ego_start_y; eo.v = v0; eo.a = 0.0;
        dl = __dadd_rn(E.interval, delay_u[b]);
        tk = 0; new_prev_acc = 0.0;
    }
    __syncwarp();                                              // all lanes are done reading the episode's inputs
    if (slot_ok) { cars_x[o] = x; cars_v[o] = v; cars_a[o] = a; }
    if (lane == 0) {
        ego[4 * b] = eo.x; ego[4 * b + 1] = eo.y; ego[4 * b + 2] = eo.v; ego[4 * b + 3] = eo.a;
        n_cars[b] = n; prev_acc[b] = new_prev_acc; delay[b] = dl; ticks[b] = tk;
        reward[b] = r; proj_jerk[b] = pj;
        flags[b] = done ? 1 : 0; flags[(size_t)B + b] = crashed ? 1 : 0; flags[2 * (size_t)B + b] = arrived ? 1 : 0;
        flags[3 * (size_t)B + b] = timeout ? 1 : 0;
    }
}

cudaError_t launch_env_step(const DevParams &P, const mpc_env_params &E, int B, int nmax, double *ego, double *cx, double *cv, double *ca,
                            int32_t *n, double *prev_acc, double *delay, int32_t *ticks, const double *jerk, const double *u_spawn,
                            const double *gap_u, const double *first_u, const double *speed_z, const double *delay_u, double *reward,
                            uint8_t *flags, double *proj_jerk, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    MPC_LAUNCH(env_step_kernel, (B + 3) / 4, 128, 0, st, P, E, B, nmax, ego, cx, cv, ca, n, prev_acc, delay, ticks, jerk, u_spawn, gap_u,
               first_u, speed_z, delay_u, reward, flags, proj_jerk);
    return cudaGetLastError();
}

cudaError_t launch_predict_layers(const DevParams &P, int B, int nmax, const double *ego, const double *cx,
                                  const double *cv, const int32_t *n, LayerDesc *desc, double *s0, double *ds,
                                  int32_t *ns, cudaStream_t st, const int32_t *subset, const int *count) {
    if (B <= 0) return cudaSuccess;
    int wpb = 4;
    if (subset) MPC_LAUNCH(predict_layers_kernel<true>, (B + wpb - 1) / wpb, wpb * 32, 0, st, P, B, ego, cx, cv, n, nmax, desc, s0, ds, ns, subset, count);
    else MPC_LAUNCH(predict_layers_kernel<false>, (B + wpb - 1) / wpb, wpb * 32, 0, st, P, B, ego, cx, cv, n, nmax, desc, s0, ds, ns, subset, count);
    return cudaGetLastError();
}

cudaError_t launch_rasterise(const DevParams &P, int B, int stride_s, const LayerDesc *desc, const double *s0,
                             const double *ds, const int32_t *ns, uint8_t *obstacles, void *distances, int dist_f32,
                             cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    // Measured on a B200 (256 episodes, H=50, 51 x 9008 cells; HBM copy peak 6551 GB/s): the segment kernel writes the fp32 grid in
    // 0.116 ms (5.08 TB/s, 0.78 of the peak; 128 threads per row) and the fp64 grid in 0.185 ms (5.71 TB/s, 0.87; 256 threads);
    // the 8-cells-per-thread kernel below needs 0.167 / 0.380 ms.  MPC_RASTER_ROWS=0 selects the latter (kept for the A/B tests and
    // for rows whose descriptor the segment kernel cannot take -- none today), MPC_RASTER_THREADS overrides the block size (dev).
    const char *e_rows = getenv("MPC_RASTER_ROWS"), *e_thr = getenv("MPC_RASTER_THREADS");
    if (!(e_rows && e_rows[0] == '0')) {
        const int v = e_thr ? atoi(e_thr) : 0;
        const int rows_threads = (v >= 32 && v <= 256 && v % 32 == 0) ? v : (dist_f32 ? 128 : 256);
        if (dist_f32) MPC_LAUNCH(rasterise_rows_kernel<float>, B * P.num_t, rows_threads, 0, st, P, B, stride_s, desc, s0, ds, ns, obstacles, (float *)distances);
        else MPC_LAUNCH(rasterise_rows_kernel<double>, B * P.num_t, rows_threads, 0, st, P, B, stride_s, desc, s0, ds, ns, obstacles, (double *)distances);
        return cudaGetLastError();
    }
    const int vec_ok = (stride_s % 8 == 0) && ((uintptr_t)obstacles % 8 == 0) && ((uintptr_t)distances % 16 == 0);
    if (dist_f32) MPC_LAUNCH(rasterise_kernel<float>, B * P.num_t, 256, 0, st, P, B, stride_s, desc, s0, ds, ns, obstacles, (float *)distances, vec_ok);
    else MPC_LAUNCH(rasterise_kernel<double>, B * P.num_t, 256, 0, st, P, B, stride_s, desc, s0, ds, ns, obstacles, (double *)distances, vec_ok);
    return cudaGetLastError();
}

cudaError_t launch_predict_step(const DevParams &P, int B, int nmax, const double *ego, const double *cx, const double *cv,
                                const double *ca, const int32_t *n, const double *sel, double dt, double mcd,
                                double *ego_out, double *ox, double *ov, double *oa, uint8_t *crashed, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    int wpb = 4;
    MPC_LAUNCH(predict_step_with_ego_kernel, (B + wpb - 1) / wpb, wpb * 32, 0, st, P, B, nmax, ego, cx, cv, ca, n, sel, dt, mcd, ego_out, ox, ov, oa, crashed);
    return cudaGetLastError();
}

cudaError_t launch_krauss_step(const DevParams &P, int B, int nmax, double *ego, double *cx, double *cv, double *ca, const int32_t *n,
                               const double *sel, double dt, double mcd, double accel, double decel, double tau, double min_gap,
                               double max_speed, uint8_t *crashed, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    KraussParams K = {accel, decel, tau, min_gap, max_speed};
    MPC_LAUNCH(krauss_step_kernel, (B + 3) / 4, 128, 0, st, P, K, B, nmax, ego, cx, cv, ca, n, sel, dt, mcd, crashed);
    return cudaGetLastError();
}

cudaError_t launch_predict_step_without_ego(const DevParams &P, int B, int nmax, const double *ego, const double *cx,
                                            const double *cv, const double *ca, const int32_t *n, double dt, double mcd,
                                            double *ego_out, double *ox, double *ov, double *oa, uint8_t *crashed, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    int wpb = 4;
    MPC_LAUNCH(predict_step_without_ego_kernel, (B + wpb - 1) / wpb, wpb * 32, 0, st, P, B, nmax, ego, cx, cv, ca, n, dt, mcd, ego_out, ox, ov, oa, crashed);
    return cudaGetLastError();
}

cudaError_t launch_state_vector(const DevParams &P, int B, int nmax, const double *ego, const double *cx, const double *cv,
                                const double *ca, const int32_t *n, float *out, int stride, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    int wpb = 4;
    MPC_LAUNCH(state_vector_kernel, (B + wpb - 1) / wpb, wpb * 32, 0, st, P, B, nmax, ego, cx, cv, ca, n, out, stride);
    return cudaGetLastError();
}

cudaError_t launch_speed_from_jerk(const DevParams &P, int B, const double *ego, const double *jerk, double *speed, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    MPC_LAUNCH(speed_from_jerk_kernel, (B + 127) / 128, 128, 0, st, P, B, ego, jerk, speed);
    return cudaGetLastError();
}

cudaError_t launch_rollout_step(const DevParams &P, int B, int nmax, double *ego, double *cx, double *cv, double *ca,
                                const int32_t *n, const double *jerk, double dt, double mcd, double stop_x, int step,
                                uint8_t *alive, double *sel_speed, double *roll_s, int roll_stride, int32_t *roll_len,
                                uint8_t *crash_pred, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    MPC_LAUNCH(rollout_step_kernel, (B + 3) / 4, 128, 0, st, P, B, nmax, ego, cx, cv, ca, n, jerk, dt, mcd, stop_x, step, alive,
                                                     sel_speed, roll_s, roll_stride, roll_len, crash_pred);
    return cudaGetLastError();
}

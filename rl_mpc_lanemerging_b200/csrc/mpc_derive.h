// Host-side derivation of DevParams from an mpc_params snapshot (once per mpc_create / mpc_set_params).
// A header so that the kernel-emulation harness (tests/emu, test infrastructure) derives exactly what the library derives.
#pragma once
#include <math.h>
#include <string.h>
#include "mpc_common.cuh"

static int host_arange_len(double start, double stop, double step) {
    double c = ceil((stop - start) / step);
    return c < 0 ? 0 : (int)c;
}
static bool near_int(double x, double eps) { return fabs(x - nearbyint(x)) < eps; }

static int derive_params(const mpc_params *p, DevParams *D) {
    memset(D, 0, sizeof(*D));
    D->p = *p;
    if (!(p->s_disc > 0) || !(p->t_disc > 0) || !(p->future_s > 0) || !(p->future_t >= 0))
        return mpc_set_error(MPC_E_INVALID, "discretisation / horizon must be positive");
    D->num_t = host_arange_len(0.0, p->future_t + p->t_disc, p->t_disc);
    if (D->num_t < 3 || D->num_t > MPC_MAX_T) return mpc_set_error(MPC_E_INVALID, "num_t must be in [3,128]");
    D->num_s_max = host_arange_len(0.0, p->future_s + p->s_disc, p->s_disc) + 1;
    if (D->num_s_max > 65000) return mpc_set_error(MPC_E_INVALID, "num_s too large for 16-bit back-pointers");
    if (D->num_s_max > (MPC_MAX_BUCKETS << MPC_BUCKET_SHIFT))        // LayerDesc::bucket_edge / bucket_band hold one entry per 64 cells
        return mpc_set_error(MPC_E_CAPACITY, "FUTURE_S / S_DISCRETIZATION above 18432 cells: the per-layer lookup tables are too small");
    D->discrete_length = (int)(p->car_length / p->s_disc);
    D->dt2 = pow(p->t_disc, 2.0);
    D->dt3 = pow(p->t_disc, 3.0);
    D->obs_min_s = p->crash_min_s - p->min_allowed_distance;
    D->crash_thresh = p->combination_min_distance - p->car_length;
    // jerk / acceleration / speed limits in cells per step^n
    double ds = p->s_disc, dt = p->t_disc;
    double jlo = p->j_min * dt * dt * dt / ds, jhi = p->j_max * dt * dt * dt / ds;
    double alo = p->a_min * dt * dt / ds, ahi = p->a_max * dt * dt / ds;
    double vmax = p->max_speed * dt / ds;
    D->lmax_exact = (int)floor(jhi - jlo) + 2;
    const double eps = 1e-6;
    D->jlo_c = (int)ceil(jlo); D->jhi_c = (int)floor(jhi);
    D->alo_c = (int)ceil(alo); D->ahi_c = (int)floor(ahi);
    D->vmax_is_int = near_int(vmax, 1e-9) ? 1 : 0;
    D->vmax_c = D->vmax_is_int ? (int)nearbyint(vmax) : (int)floor(vmax);
    D->jhi_r = jhi; D->ahi_r = ahi; D->vmax_r = vmax;
    bool ok = !near_int(jlo, eps) && !near_int(jhi, eps) && !near_int(alo, eps) && !near_int(ahi, eps);
    if (!D->vmax_is_int) ok = ok && !near_int(vmax, eps) && !near_int(vmax - jhi, eps) && !near_int(vmax - ahi, eps);
    D->lmax = D->jhi_c - D->jlo_c + 1;              // longest on-grid successor window
    ok = ok && D->vmax_c <= 250 && D->alo_c >= -16 && D->ahi_c <= 15 && D->lmax >= 1 && D->lmax <= 7 && alo < 0 && ahi > 0;
    D->fast_ok = ok ? 1 : 0;
    D->cv = (float)(p->v_weight * (ds / dt) * (ds / dt));
    D->ca = (float)(p->a_weight * (ds / (dt * dt)) * (ds / (dt * dt)));
    D->cj = (float)(p->j_weight * (ds / (dt * dt * dt)) * (ds / (dt * dt * dt)));
    D->vdes_c = (float)(p->desired_speed * dt / ds);
    D->dw = (float)p->d_weight;
    // fixed-point cost tables (fast kernel): every entry must fit 32 bits
    double mx = 0.0;
    // (speeds above vmax_c are never looked up: the window clamp of st_cy.pyx:72 keeps v' <= vmax_c; a coarse probe grid has v = 255 far beyond MAX_SPEED)
    for (int v = 0; v < 256; v++) { double x = p->v_weight * (v * ds / dt - p->desired_speed) * (v * ds / dt - p->desired_speed); if (v <= D->vmax_c + 1) mx = fmax(mx, x); D->vtab[v] = (unsigned)llrint(fmin(x, 16000.0) * MPC_FX_ONE); }
    for (int i = 0; i < 32; i++) { double acc = (i - 16) * ds / (dt * dt), x = p->a_weight * acc * acc; mx = fmax(mx, x); D->atab[i] = (unsigned)llrint(fmin(x, 16000.0) * MPC_FX_ONE); }
    for (int i = 0; i < 16; i++) { double jk = (i - 8) * ds / (dt * dt * dt), x = p->j_weight * jk * jk; mx = fmax(mx, x); D->jtab[i] = (unsigned)llrint(fmin(x, 16000.0) * MPC_FX_ONE); }
    if (mx >= 16000.0 || D->jlo_c < -8 || D->jhi_c > 7 || !(p->d_weight >= 0) || p->d_weight > 1e4) D->fast_ok = 0;
    // cheapest label of a path with one step inside a penalty zone: d_w * 1e6 / max(d,1) with d < min_allowed (st_cy.pyx:34-38)
    D->bound_fx = 0;
    if (p->d_weight > 0 && p->min_allowed_distance > 0) {
        double zone = p->d_weight * 1000000.0 / fmax(p->min_allowed_distance, 1.0);
        if (zone > 64.0) D->bound_fx = (unsigned long long)llrint(zone * MPC_FX_ONE) - 2;
        // band cell range vs metric band edge differ by < 3 cells (int() of the car position, of CAR_LENGTH/ds and of the
        // uncertainty, st.py:52-66): cells within floor(min_allowed/ds) - 4 of a band are strictly inside the zone
        int zc = (int)floor(p->min_allowed_distance / ds) - 4;
        D->zone_cells = (D->bound_fx && zc > 0) ? zc : 0;
        // LayerDesc::blk joins a car's band with the two penalty zones next to it into ONE interval; that is the exact set of
        // blocked cells as long as the zones (m wide) cover the few cells by which int() rounding lets the band miss its edges
        D->zone_ok = (D->bound_fx && p->min_allowed_distance >= 4.0 * ds && p->min_allowed_distance >= 1.0) ? 1 : 0;
    }
    D->kw = (float)(p->d_weight * MPC_FX_ONE);
    D->vstar_c = 0;
    for (int v = 1; v <= D->vmax_c && v < 256; v++) if (D->vtab[v] < D->vtab[D->vstar_c]) D->vstar_c = v;
    // the reachability heuristic of hinted solves (mpc_reach.cu) needs the ROUNDED table to be convex up to vstar_c
    for (int v = 1; v < D->vstar_c; v++)
        if ((long long)D->vtab[v - 1] - 2LL * D->vtab[v] + (long long)D->vtab[v + 1] < 0) { D->vstar_c = 0; break; }
    // 32-bit-key kernel (mpc_fast32.cuh): the whole label must stay below 0xff000000.  A plan that never enters a penalty zone
    // costs at most ~(num_t - 1) * (max V + d_w / m) (standing still all the way; measured maxima are within 5 % of it), so the
    // label precision is the largest 2^-q (q <= 18) that holds 1.25 x that plus one edge and one penalty.  The bound only decides
    // which problems the kernel can finish (the others go to the 64-bit kernel), never what the answer is.
    D->f32_ok = 0;
    if (D->fast_ok && D->zone_ok && D->bound_fx) {
        double maxv = 0.0, maxa = 0.0, maxj = 0.0;
        for (int v = 0; v <= D->vmax_c + 1 && v < 256; v++) maxv = fmax(maxv, p->v_weight * (v * ds / dt - p->desired_speed) * (v * ds / dt - p->desired_speed));
        // (on-grid edges only use a' in [alo_c, ahi_c] and j' in [jlo_c, jhi_c]: int_window_sa)
        for (int i = D->alo_c; i <= D->ahi_c; i++) { double acc = i * ds / (dt * dt); maxa = fmax(maxa, p->a_weight * acc * acc); }
        for (int i = D->jlo_c; i <= D->jhi_c; i++) { double jk = i * ds / (dt * dt * dt); maxj = fmax(maxj, p->j_weight * jk * jk); }
        const double maxpen = p->d_weight / p->min_allowed_distance, edge = maxv + maxa + maxj;
        const double need = 1.25 * (D->num_t - 1) * (maxv + maxpen) + 2.0 * (edge + maxpen) + 16.0;
        int q = MPC_FX_FRAC;
        while (q >= 12 && ldexp(need, q) >= 4278190080.0) q--;
        if (q >= 12) {
            D->f32_ok = 1; D->f32_frac = q; D->f32_one = ldexp(1.0, q);
            D->kw32 = (float)(p->d_weight * D->f32_one);
            unsigned mv = 0, ma = 0, mj = 0;
            for (int v = 0; v < 256; v++) { double x = p->v_weight * (v * ds / dt - p->desired_speed) * (v * ds / dt - p->desired_speed); D->vtab32[v] = (unsigned)llrint(fmin(x, 16000.0) * D->f32_one); if (v <= D->vmax_c + 1 && D->vtab32[v] > mv) mv = D->vtab32[v]; }
            for (int i = 0; i < 32; i++) { double acc = (i - 16) * ds / (dt * dt), x = p->a_weight * acc * acc; D->atab32[i] = (unsigned)llrint(fmin(x, 16000.0) * D->f32_one); if (i - 16 >= D->alo_c && i - 16 <= D->ahi_c && D->atab32[i] > ma) ma = D->atab32[i]; }
            for (int i = 0; i < 16; i++) { double jk = (i - 8) * ds / (dt * dt * dt), x = p->j_weight * jk * jk; D->jtab32[i] = (unsigned)llrint(fmin(x, 16000.0) * D->f32_one); if (i - 8 >= D->jlo_c && i - 8 <= D->jhi_c && D->jtab32[i] > mj) mj = D->jtab32[i]; }
            const double room = 4278190080.0 - 8.0 - (double)mv - (double)ma - (double)mj - ceil(ldexp(maxpen, q) * 1.001);
            D->f32_bound = (unsigned)room;
        }
    }
    return MPC_OK;
}

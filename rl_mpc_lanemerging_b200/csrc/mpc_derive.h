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
    return MPC_OK;
}

/* CPU oracle for the S-T MPC hot path -- TEST INFRASTRUCTURE, NOT THE PRODUCT.
 * See mpc_oracle.h.  fp64 everywhere, operation order follows the reference so that results are
 * bit-identical to the reference's Cython/Python code (verified by tests/golden/make_golden.py).
 * Compile: gcc -O2 -ffp-contract=off -fPIC -shared -pthread mpc_oracle.c -lm
 */
#include "mpc_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- Settings snapshot (config.py:30-37,94-110,143,153 and configs/st_*.json) ------------- */
void orc_default_params(orc_params *p) {
    p->s_disc = 0.05; p->t_disc = 0.30; p->future_s = 150.0; p->future_t = 5.0;
    p->start_uncertainty = 0.0; p->uncertainty_per_second = 0.0;
    p->d_weight = 10.0; p->v_weight = 0.5; p->a_weight = 10.0; p->j_weight = 10.0;
    p->desired_speed = 30.0; p->max_speed = 30.0;
    p->a_min = -6.0; p->a_max = 4.5; p->j_min = -5.0; p->j_max = 5.0;
    p->min_allowed_distance = 5.0; p->crash_min_s = 20.0; p->car_length = 5.0;
    p->max_predicted_decel = -4.0; p->tick_length = 0.2; p->sensor_radius = 125.0;
    p->combination_min_distance = 5.1;
}

/* ---- geometry: control.py:366-380 ---------------------------------------------------------- */
static const double MP_X = -50.9, MP_Y = 1.72;     /* merge_point  */
static const double MP2_X = 1.5, MP2_Y = -1.5;     /* merge_point2 */
static const double MP3_X = -51.0;                 /* merge_point3[0] */

static double dist2d(double x0, double y0, double x1, double y1) { /* control.py:37-38 */
    double dx = x0 - x1, dy = y0 - y1;
    return sqrt(dx * dx + dy * dy);
}

double orc_get_ego_s(double x, double y) {
    if (x < MP_X) return -dist2d(x, y, MP_X, MP_Y);
    else if (x < MP2_X) return dist2d(x, y, MP_X, MP_Y);
    else return x - MP2_X + (MP2_X - MP3_X);   /* common_s = 1.5 - (-51) */
}

/* ---- numpy.arange(start, stop, step) length: ceil((stop-start)/step) in fp64 ---------------- */
int orc_arange_len(double start, double stop, double step) {
    double v = (stop - start) / step;
    double c = ceil(v);
    if (c < 0) c = 0;
    return (int)c;
}

/* st.py:31-32: s_values = arange(s0, s0+future_s+ds, ds); t_values = arange(0, T+dt, dt).
 * numpy fills element i as start + i*(second-first); delta_s_eff is that (second-first). */
void orc_grid_dims(const orc_params *p, double start_s, int *num_t, int *num_s, double *delta_s_eff) {
    if (num_s) *num_s = orc_arange_len(start_s, start_s + p->future_s + p->s_disc, p->s_disc);
    if (num_t) *num_t = orc_arange_len(0.0, p->future_t + p->t_disc, p->t_disc);
    if (delta_s_eff) *delta_s_eff = (start_s + p->s_disc) - start_s;
}

static void fill_s_values(double start_s, double s_disc, int num_s, double *s_values) {
    double delta = (start_s + s_disc) - start_s;
    for (int i = 0; i < num_s; i++) s_values[i] = start_s + (double)i * delta;
    if (num_s > 1) s_values[1] = start_s + s_disc;  /* numpy writes first two explicitly */
}

/* ---- prediction.py:46-105 ------------------------------------------------------------------ */
int orc_predict_step_with_ego(const orc_params *p, const orc_state *in, double selected_speed,
                              double dt, double min_crash_distance, orc_state *out) {
    double cx = in->ego_x, cy = in->ego_y, px, py;
    if (cx < MP2_X) {                                   /* 48-56 */
        double dx = MP2_X - cx, dy = MP2_Y - cy;
        double nrm = sqrt(dx * dx + dy * dy);
        dx /= nrm; dy /= nrm;
        double step = selected_speed * dt;
        dx *= step; dy *= step;
        px = cx + dx; py = cy + dy;
        if (py < -1.6) py = -1.6;
    } else {                                            /* 57-59 */
        py = cy; px = cx + selected_speed * dt;
    }
    double next_acc = (selected_speed - in->ego_v) / dt; /* 61 */
    double pes = orc_get_ego_s(px, py);
    int ego_can_crash = pes > 11.0;                      /* 64, ego_crash_threshold */
    int ego_has_merged = pes > 8.0;                      /* 66, ego_reaction_threshold */

    double last_x = INFINITY, last_speed = 0.0;          /* 72-73 */
    int ego_encountered = 0, n = in->n;
    double nx[ORC_MAX_CARS], nv[ORC_MAX_CARS], na[ORC_MAX_CARS];
    for (int i = 0; i < n; i++) {                        /* 75-97 */
        double ov = in->v[i], ox = in->x[i];
        if (ox < px && !ego_encountered) {
            ego_encountered = 1;
            if (ego_has_merged) { last_x = px; last_speed = selected_speed; }
        }
        double speed_diff = last_speed - ov, x_diff = last_x - ox, nacc, nspd;
        if (speed_diff < 0 && x_diff < 30) {
            nacc = (p->max_predicted_decel > speed_diff) ? p->max_predicted_decel : speed_diff;
            nspd = ov + nacc * dt;
        } else { nacc = 0; nspd = ov; }
        double npos = ox + nspd * dt;
        last_x = npos; last_speed = nspd;
        nx[i] = npos; nv[i] = nspd; na[i] = nacc;
    }
    int crashed = 0;                                     /* 99-103 */
    double cdd = (min_crash_distance > p->car_length) ? min_crash_distance : p->car_length;
    for (int i = 0; i < n; i++)
        if (fabs(nx[i] - px) < cdd && ego_can_crash) crashed = 1;
    out->ego_x = px; out->ego_y = py; out->ego_v = selected_speed; out->ego_a = next_acc;
    out->n = n;
    memcpy(out->x, nx, sizeof(double) * n); memcpy(out->v, nv, sizeof(double) * n);
    memcpy(out->a, na, sizeof(double) * n);
    return crashed;
}

/* ---- prediction.py:22-44 ------------------------------------------------------------------- */
int orc_predict_step_without_ego(const orc_params *p, const orc_state *in, double dt,
                                 double min_crash_distance, orc_state *out) {
    double ego_s = orc_get_ego_s(in->ego_x, in->ego_y), ego_x = in->ego_x;
    if (ego_s < 8.0 || in->n == 0)                                   /* 26-27 */
        return orc_predict_step_with_ego(p, in, 0.0, dt, min_crash_distance, out);
    orc_state m = *in;
    if (in->x[0] < ego_x) {                                          /* 28-31 */
        m.ego_x = -20; m.ego_y = -10; m.ego_v = 0; m.ego_a = 0;
        return orc_predict_step_with_ego(p, &m, 0.0, dt, min_crash_distance, out);
    }
    double last_speed = 0, last_x = 0;                               /* 33-44 */
    for (int i = 0; i < in->n; i++) {
        if (in->x[i] < ego_x) {
            m.ego_x = last_x - p->car_length - 5; m.ego_v = last_speed; m.ego_a = 0;
            return orc_predict_step_with_ego(p, &m, last_speed, dt, min_crash_distance, out);
        }
        last_speed = in->v[i]; last_x = in->x[i];
    }
    return orc_predict_step_with_ego(p, in, last_speed, dt, min_crash_distance, out);
}

/* ---- st.py:20-22,25-70 --------------------------------------------------------------------- */
void orc_build_grid(const orc_params *p, const orc_state *st, int num_t, int num_s,
                    uint8_t *obstacles, double *distances, double *s_values, double *obs_s_out) {
    double start_s = orc_get_ego_s(st->ego_x, st->ego_y);
    fill_s_values(start_s, p->s_disc, num_s, s_values);
    memset(obstacles, 0, (size_t)num_t * num_s);
    for (size_t i = 0; i < (size_t)num_t * num_s; i++) distances[i] = 0.0 + 1E10;
    int discrete_length = (int)(p->car_length / p->s_disc);
    orc_state cur = *st, nxt;
    double s_last = s_values[num_s - 1];
    for (int ti = 0; ti < num_t; ti++) {
        double t = 0.0 + (double)ti * p->t_disc;            /* arange(0, ..)[ti] */
        if (ti == 1) t = p->t_disc;
        double uncertainty = p->start_uncertainty + p->uncertainty_per_second * t;
        int du = (int)(uncertainty / p->s_disc);
        if (ti != 0) { orc_predict_step_without_ego(p, &cur, p->t_disc, 5.0, &nxt); cur = nxt; }
        uint8_t *orow = obstacles + (size_t)ti * num_s;
        double *drow = distances + (size_t)ti * num_s;
        if (obs_s_out)
            for (int c = 0; c < ORC_MAX_CARS; c++)
                obs_s_out[ti * ORC_MAX_CARS + c] = (c < cur.n) ? cur.x[c] - MP3_X : NAN;
        for (int c = 0; c < cur.n; c++) {
            double obs_s = cur.x[c] - MP3_X;                               /* control.py:388-389 */
            if (obs_s < p->crash_min_s - p->min_allowed_distance) break;   /* st.py:46-47 */
            else if (obs_s > s_last + p->car_length) continue;             /* 48-49 */
            double ef = obs_s - p->car_length - uncertainty;               /* 52 */
            double eb = obs_s + p->car_length + uncertainty;               /* 53 */
            for (int k = 0; k < num_s; k++) {
                double df = fabs(s_values[k] - ef), db = fabs(s_values[k] - eb);
                double d = drow[k];
                d = (df < d) ? df : d;                                      /* 56 */
                d = (db < d) ? db : d;                                      /* 57 */
                drow[k] = d;
            }
            int si = (int)((obs_s - start_s) / p->s_disc);                  /* 60, get_range_index */
            int imin = si - discrete_length - du; if (imin < 0) imin = 0;   /* 61 */
            int imax = si + discrete_length + du; if (imax > num_s) imax = num_s; /* 62 */
            if (imin < num_s && imax > 0)                                   /* 63-65 */
                for (int k = imin; k < imax; k++) { orow[k] = 1; drow[k] = 0; }
        }
    }
}

/* ---- solver helpers: st_cy.pyx:34-38, 46-50, 65-75, 78-93 ----------------------------------- */
static double distance_penalty(double d, double min_allowed) {
    if (d < min_allowed) return 1000000.0 / (d > 1.0 ? d : 1.0);
    return 1 / d;
}

static double cost_with_jerk(const orc_params *p, double s, double s1, double s2, double s3,
                             double dt, double min_distance) {
    double v = (s - s1) / dt;
    double a = (s - 2 * s1 + s2) / pow(dt, 2.0);
    double j = (s - 3 * s1 + 3 * s2 - s3) / pow(dt, 3.0);
    return p->v_weight * ((v - p->desired_speed) * (v - p->desired_speed)) + p->a_weight * (a * a) +
           p->j_weight * (j * j) + p->d_weight * distance_penalty(min_distance, p->min_allowed_distance);
}

static void next_index_range(const orc_params *p, double start_s, double delta_s, double s,
                             double s1, double s2, double dt, int *imin, int *imax_excl) {
    double prev_v = (s1 - s2) / dt;                                      /* 66 */
    double v = (s - s1) / dt;
    double a = (v - prev_v) / dt;
    double min_a = a + p->j_min * dt; if (p->a_min > min_a) min_a = p->a_min;     /* 69 */
    double max_a = a + p->j_max * dt; if (p->a_max < max_a) max_a = p->a_max;     /* 70 */
    double min_v = v + min_a * dt; if (0 > min_v) min_v = 0;                      /* 71 */
    double max_v = v + max_a * dt; if (p->max_speed < max_v) max_v = p->max_speed;/* 72 */
    double min_s = s + min_v * dt, max_s = s + max_v * dt;                        /* 73-74 */
    double min_exact = (min_s - start_s) / delta_s;                               /* 88 */
    int mi = (int)min_exact;
    int ma = (int)((max_s - start_s) / delta_s);
    if (mi < min_exact) mi += 1;
    *imin = mi; *imax_excl = ma + 1;
}

/* ---- Dijkstra: st_cy.pyx:315-399 ------------------------------------------------------------ */
typedef struct hnode {
    double total; int t, k; int k1; double s1; int k2; double s2;
} hnode;

static int hless(const hnode *a, const hnode *b) {   /* python tuple order of the 8-tuple, st_cy.pyx:388 */
    if (a->total != b->total) return a->total < b->total;
    if (a->t != b->t) return a->t < b->t;
    if (a->k != b->k) return a->k < b->k;            /* s_value is a function of k */
    if (a->k1 != b->k1) return a->k1 < b->k1;
    if (a->s1 != b->s1) return a->s1 < b->s1;
    if (a->k2 != b->k2) return a->k2 < b->k2;
    return a->s2 < b->s2;
}

typedef struct heap { hnode *a; long n, cap; } heap;
static void hpush(heap *h, hnode x) {
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 1024; h->a = (hnode *)realloc(h->a, h->cap * sizeof(hnode)); }
    long i = h->n++;
    while (i > 0) { long par = (i - 1) / 2; if (!hless(&x, &h->a[par])) break; h->a[i] = h->a[par]; i = par; }
    h->a[i] = x;
}
static hnode hpop(heap *h) {
    hnode top = h->a[0], x = h->a[--h->n];
    long i = 0;
    for (;;) {
        long c = 2 * i + 1; if (c >= h->n) break;
        if (c + 1 < h->n && hless(&h->a[c + 1], &h->a[c])) c++;
        if (!hless(&h->a[c], &x)) break;
        h->a[i] = h->a[c]; i = c;
    }
    if (h->n > 0) h->a[i] = x;
    return top;
}

static int backtrack(int num_t, int num_s, const int *previous, const double *s_values, int best_t,
                     int best_k, int *idx_out, double *s_seq_out) {
    for (int t = 0; t < num_t; t++) { if (idx_out) idx_out[t] = -1; if (s_seq_out) s_seq_out[t] = 0.0; }
    int k = best_k;
    for (int t = best_t; t > 0; t--) {                                  /* 394-396 */
        if (idx_out) idx_out[t] = k; if (s_seq_out) s_seq_out[t] = s_values[k];
        k = previous[(size_t)t * num_s + k];
    }
    if (idx_out) idx_out[0] = k; if (s_seq_out) s_seq_out[0] = s_values[k];
    return best_t;
}

int orc_solve_dijkstra(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                       const double *distances, const double *s_values, double delta_t,
                       double v0, double a0, int *idx_out, double *s_seq_out, double *cost_out,
                       orc_solve_stats *stats) {
    double delta_s = s_values[1] - s_values[0], start_s = s_values[0];      /* 318-320 */
    uint8_t *enc = (uint8_t *)calloc((size_t)num_t * num_s, 1);
    int *previous = (int *)calloc((size_t)num_t * num_s, sizeof(int));
    double est_prev = start_s - v0 * delta_t;                               /* 329 */
    double est_second = est_prev - delta_t * (v0 - a0 * delta_t);           /* 330 */
    heap h = {0, 0, 0};
    hnode first = {0, 0, 0, 0, est_prev, 0, est_second};
    hpush(&h, first);
    int best_last = 0, best_t = 0; double best_total = 0; long pops = 0, pushes = 1;
    while (h.n > 0) {
        hnode n = hpop(&h); pops++;
        size_t id = (size_t)n.t * num_s + n.k;
        if (enc[id]) continue;
        enc[id] = 1; previous[id] = n.k1;
        if (n.t > best_t) { best_t = n.t; best_last = n.k; best_total = n.total; }
        if (n.t == num_t - 1) break;
        double sv = s_values[n.k];
        int imin, imax;
        next_index_range(p, start_s, delta_s, sv, n.s1, n.s2, delta_t, &imin, &imax);
        int nt = n.t + 1;
        for (int k = imin; k < imax; k++) {
            if (k >= num_s) break;
            size_t nid = (size_t)nt * num_s + k;
            if (!enc[nid]) {
                if (obstacles[nid]) continue;
                double c = cost_with_jerk(p, s_values[k], sv, n.s1, n.s2, delta_t, distances[nid]);
                hnode m = {n.total + c, nt, k, n.k, sv, n.k1, n.s1};
                hpush(&h, m); pushes++;
            }
        }
    }
    if (cost_out) *cost_out = best_total;
    if (stats) { stats->pops = pops; stats->pushes = pushes; }
    int r = backtrack(num_t, num_s, previous, s_values, best_t, best_last, idx_out, s_seq_out);
    free(enc); free(previous); free(h.a);
    return r;
}

/* ---- forward layered DP giving the same answer ---------------------------------------------- */
int orc_solve_layered(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                      const double *distances, const double *s_values, double delta_t,
                      double v0, double a0, int *idx_out, double *s_seq_out, double *cost_out,
                      orc_solve_stats *stats) {
    double delta_s = s_values[1] - s_values[0], start_s = s_values[0];
    int *previous = (int *)calloc((size_t)num_t * num_s, sizeof(int));
    double *lab[2]; int *k1[2], *k2[2]; double *s1[2], *s2[2];
    for (int b = 0; b < 2; b++) {
        lab[b] = (double *)malloc(sizeof(double) * num_s);
        k1[b] = (int *)malloc(sizeof(int) * num_s); k2[b] = (int *)malloc(sizeof(int) * num_s);
        s1[b] = (double *)malloc(sizeof(double) * num_s); s2[b] = (double *)malloc(sizeof(double) * num_s);
        for (int k = 0; k < num_s; k++) lab[b][k] = INFINITY;
    }
    int *wcount = (int *)calloc((size_t)num_s + 1, sizeof(int)); long wh[8] = {0};
    int *coast_mark = (int *)malloc(sizeof(int) * (3 * (size_t)num_s + 8));
    for (size_t i = 0; i < 3 * (size_t)num_s + 8; i++) coast_mark[i] = -1;
    double est_prev = start_s - v0 * delta_t;
    double est_second = est_prev - delta_t * (v0 - a0 * delta_t);
    lab[0][0] = 0; k1[0][0] = 0; s1[0][0] = est_prev; k2[0][0] = 0; s2[0][0] = est_second;
    int lo = 0, hi = 0;              /* span of the current layer's finite labels */
    int best_t = 0, best_k = 0; double best_total = 0;
    long nodes = 0, edges = 0, coll = 0; int max_span = 1, max_width = 1;
    for (int t = 0; t < num_t; t++) {
        int cur = t & 1, nxt = cur ^ 1;
        /* arg-min of this layer (ties -> smaller index) = first pop of the layer in the Dijkstra */
        int amin = -1, width = 0;
        for (int k = lo; k <= hi; k++) if (lab[cur][k] < INFINITY) {
            width++;
            if (amin < 0 || lab[cur][k] < lab[cur][amin]) amin = k;
        }
        if (amin < 0) break;
        best_t = t; best_k = amin; best_total = lab[cur][amin];
        nodes += width; if (width > max_width) max_width = width;
        if (hi - lo + 1 > max_span) max_span = hi - lo + 1;
        for (int k = lo; k <= hi; k++) if (lab[cur][k] < INFINITY) previous[(size_t)t * num_s + k] = k1[cur][k];
        if (t == num_t - 1) break;
        int nlo = num_s, nhi = -1;
        for (int k = lo; k <= hi; k++) {
            if (!(lab[cur][k] < INFINITY)) continue;
            double sv = s_values[k];
            int imin, imax;
            next_index_range(p, start_s, delta_s, sv, s1[cur][k], s2[cur][k], delta_t, &imin, &imax);
            if (imin >= 0 && imin < num_s && imax > imin) wcount[imin]++;
            if (t >= 2) {   /* statistics only: coast cell of this node */
                int c = k + (k - k1[cur][k]) + ((k - k1[cur][k]) - (k1[cur][k] - k2[cur][k]));
                if (c >= 0 && c < 3 * num_s + 8) { if (coast_mark[c] == t) coll++; coast_mark[c] = t; }
            }
            for (int kk = imin; kk < imax; kk++) {
                if (kk >= num_s) break;
                size_t nid = (size_t)(t + 1) * num_s + kk;
                if (obstacles[nid]) continue;
                edges++;
                double c = cost_with_jerk(p, s_values[kk], sv, s1[cur][k], s2[cur][k], delta_t, distances[nid]);
                double tot = lab[cur][k] + c;
                if (tot < lab[nxt][kk]) {   /* strict: equal totals keep the smaller predecessor index */
                    lab[nxt][kk] = tot; k1[nxt][kk] = k; s1[nxt][kk] = sv;
                    k2[nxt][kk] = k1[cur][k]; s2[nxt][kk] = s1[cur][k];
                    if (kk < nlo) nlo = kk; if (kk > nhi) nhi = kk;
                }
            }
        }
        for (int k = lo; k <= hi; k++) lab[cur][k] = INFINITY;
        for (int k = 0; k < num_s; k++) if (wcount[k]) { wh[wcount[k] > 8 ? 7 : wcount[k] - 1]++; wcount[k] = 0; }
        lo = nlo; hi = nhi;
        if (nhi < 0) break;
    }
    if (cost_out) *cost_out = best_total;
    if (stats) { stats->nodes = nodes; stats->edges = edges; stats->coast_collisions = coll;
                 stats->max_span = max_span; stats->max_width = max_width;
                 for (int i = 0; i < 8; i++) stats->wmult_hist[i] = wh[i]; }
    int r = backtrack(num_t, num_s, previous, s_values, best_t, best_k, idx_out, s_seq_out);
    for (int b = 0; b < 2; b++) { free(lab[b]); free(k1[b]); free(k2[b]); free(s1[b]); free(s2[b]); }
    free(previous); free(coast_mark); free(wcount);
    return r;
}


/* ---- CPU model of the CUDA fast kernel's arithmetic.  NOT part of the reference: it exists so that the
 * fast mode's tolerance claim can be checked on CPU over many states, and so that the GPU kernel can be
 * tested bit for bit.  Mirrors fast_pull_kernel (rl_mpc_lanemerging_b200/csrc/mpc_fast.cu): layers 0/1 in
 * exact fp64 then quantised to 2^-18 fixed point, integer-cell kinematics with tabulated edge costs from
 * layer 2 on, 48-bit integer labels, ties to the larger v' (= smaller predecessor index).
 * label_bits: 0 = the kernel's arithmetic; 32 = (rejected design) fp32 labels, kept to document why. */
#include <float.h>
#define FXONE 262144.0


static int64_t g_last_collisions = 0; /* pushes that met a candidate of the same (label >> coll_shift) bucket in their cell (32-bit-key kernel: overflow-list entries) */
static int g_coll_shift = 8;
int64_t orc_last_collisions(void) { return g_last_collisions; }
void orc_set_coll_shift(int s) { g_coll_shift = s; }
static int g_last_max_span = 0;      /* widest layer span (cells) the kernel's ring would have had to hold in the last model run */
int orc_last_max_span(void) { return g_last_max_span; }

static int fast_model_impl(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                            const double *distances, const double *s_values, double delta_t,
                            double v0, double a0, int f32_labels, uint64_t prune_fx, const uint64_t *hfx, uint64_t *fmin_out,
                            int *idx_out, double *s_seq_out, double *cost_out, int64_t *counts, int frac_bits, int key_shift) {
    /* frac_bits: fixed-point precision of the labels (18 = the 64-bit kernel); key_shift: low label bits that do not take part in
     * the min-combine (candidates whose labels agree above them are ordered by v' alone) -- the 32-bit-key kernel, mpc_fast32.cu */
    const double FX1 = ldexp(1.0, frac_bits);
#define FXQ(x) ((uint64_t)llrint((x) * FX1))
    /* prune_fx != 0: nodes whose label exceeds it are dropped (the fast kernel's cost bound, mpc_fast.cu);
     * counts[0..1] = nodes expanded, pushes */
    int64_t n_nodes = 0, n_push = 0, n_coll = 0; int max_span_ = 0;
    double dsn = p->s_disc, dt = p->t_disc;
    double jlo = p->j_min * dt * dt * dt / dsn, jhi = p->j_max * dt * dt * dt / dsn;
    double alo_r = p->a_min * dt * dt / dsn, ahi_r = p->a_max * dt * dt / dsn, vmax_r = p->max_speed * dt / dsn;
    int jlo_c = (int)ceil(jlo), jhi_c = (int)floor(jhi), alo_c = (int)ceil(alo_r), ahi_c = (int)floor(ahi_r);
    int vmax_is_int = fabs(vmax_r - nearbyint(vmax_r)) < 1e-9;
    int vmax_c = vmax_is_int ? (int)nearbyint(vmax_r) : (int)floor(vmax_r);
    uint32_t vtab[256], atab[32], jtab[16];
    for (int v = 0; v < 256; v++) { double x = p->v_weight * (v * dsn / dt - p->desired_speed) * (v * dsn / dt - p->desired_speed); vtab[v] = (uint32_t)llrint(fmin(x, 16000.0) * FX1); }
    for (int i = 0; i < 32; i++) { double acc = (i - 16) * dsn / (dt * dt), x = p->a_weight * acc * acc; atab[i] = (uint32_t)llrint(fmin(x, 16000.0) * FX1); }
    for (int i = 0; i < 16; i++) { double jk = (i - 8) * dsn / (dt * dt * dt), x = p->j_weight * jk * jk; jtab[i] = (uint32_t)llrint(fmin(x, 16000.0) * FX1); }
    const float kwf = (float)(p->d_weight * FX1);
    float cvf = (float)(p->v_weight * (dsn / dt) * (dsn / dt)), caf = (float)(p->a_weight * (dsn / (dt * dt)) * (dsn / (dt * dt)));
    float cjf = (float)(p->j_weight * (dsn / (dt * dt * dt)) * (dsn / (dt * dt * dt))), vdes = (float)(p->desired_speed * dt / dsn);
    double delta_s = s_values[1] - s_values[0], start_s = s_values[0];
    int *previous = (int *)calloc((size_t)num_t * num_s, sizeof(int));
    /* per cell: label (integer fixed point, or float bits when f32_labels), v, a */
    uint64_t *lab[2]; float *labf[2]; int *vv[2], *aa[2]; uint8_t *has[2];
    for (int b = 0; b < 2; b++) { lab[b] = (uint64_t *)malloc(8 * num_s); labf[b] = (float *)malloc(4 * num_s); vv[b] = (int *)malloc(4 * num_s);
                                  aa[b] = (int *)malloc(4 * num_s); has[b] = (uint8_t *)calloc(num_s, 1); }
    double est_prev = start_s - v0 * delta_t, est_second = est_prev - delta_t * (v0 - a0 * delta_t);
    int best_t = 0, best_k = 0; uint64_t best_lab = 0; float best_labf = 0;
    int imin, imax;
    next_index_range(p, start_s, delta_s, start_s, est_prev, est_second, delta_t, &imin, &imax);
    int lo = num_s, hi = -1;
    /* layer 1: complete label */
    for (int k = imin; k < imax && k < num_s; k++) {
        if (obstacles[(size_t)num_s + k]) continue;
        double c = cost_with_jerk(p, s_values[k], start_s, est_prev, est_second, delta_t, distances[(size_t)num_s + k]);
        lab[1][k] = FXQ(c); labf[1][k] = (float)c; vv[1][k] = k; aa[1][k] = 0; has[1][k] = 1; previous[(size_t)num_s + k] = 0;
        if (k < lo) lo = k; if (k > hi) hi = k;
    }
#define BETTER(nl, nlf, nv, kk, buf) (!has[buf][kk] || (f32_labels ? ((nlf) < labf[buf][kk]) : (((nl) >> key_shift) < (lab[buf][kk] >> key_shift) || (((nl) >> key_shift) == (lab[buf][kk] >> key_shift) && (nv) > vv[buf][kk]))))
    if (hi >= 0) {
        best_t = 1; best_k = -1;
        for (int k = lo; k <= hi; k++) if (has[1][k]) {
            int better = best_k < 0 || (f32_labels ? labf[1][k] < best_labf : lab[1][k] < best_lab);
            if (better) { best_k = k; best_lab = lab[1][k]; best_labf = labf[1][k]; }
        }
        /* layer 1 -> 2: exact windows, exact kinematic cost quantised; penalty of layer 2 added when layer 2 is finalised */
        int dlo = num_s, dhi = -1;
        for (int k1 = lo; k1 <= hi; k1++) if (has[1][k1]) {
            double s = s_values[k1];
            next_index_range(p, start_s, delta_s, s, start_s, est_prev, delta_t, &imin, &imax);
            for (int kk = imin; kk < imax && kk < num_s; kk++) {
                double sn = s_values[kk];
                double v = (sn - s) / delta_t, a = (sn - 2 * s + start_s) / pow(delta_t, 2.0), j = (sn - 3 * s + 3 * start_s - est_prev) / pow(delta_t, 3.0);
                double kin = p->v_weight * ((v - p->desired_speed) * (v - p->desired_speed)) + p->a_weight * (a * a) + p->j_weight * (j * j);
                uint64_t tot = lab[1][k1] + FXQ(kin); float totf = labf[1][k1] + (float)kin;
                int vn = kk - k1;
                if (prune_fx && tot > prune_fx) continue;
                n_push++;
                if (BETTER(tot, totf, vn, kk, 0)) { lab[0][kk] = tot; labf[0][kk] = totf; vv[0][kk] = vn; aa[0][kk] = vn - k1; has[0][kk] = 1; }
                if (kk < dlo) dlo = kk; if (kk > dhi) dhi = kk;
            }
        }
        for (int k = lo; k <= hi; k++) has[1][k] = 0;
        /* pass t: finalise layer t (buffer t&1), push into layer t+1 */
        for (int t = 2; t < num_t && dhi >= 0; t++) {
            int cur = t & 1, nxt = cur ^ 1, any = 0, nlo = num_s, nhi = -1, bk = -1; uint64_t bl = 0; float blf = 0;
            int klo_ = num_s, khi_ = -1;
            for (int k = dlo; k <= dhi; k++) {
                if (!has[cur][k]) continue;
                has[cur][k] = 0;
                size_t id = (size_t)t * num_s + k;
                if (obstacles[id]) continue;
                double d = distances[id];
                float df = (float)d;
                /* zone: fp64 as the reference; elsewhere d_w/d in fp32: rint(kw * (1.0f / (float)d)), the kernel's fx_inv_penalty */
                uint64_t penfx = (d < p->min_allowed_distance) ? FXQ(p->d_weight * (1000000.0 / (d > 1.0 ? d : 1.0)))
                                                               : (uint64_t)lrintf(kwf * (1.0f / df));
                uint64_t label = lab[cur][k] + penfx;
                float penf = (d < p->min_allowed_distance) ? 1000000.0f / fmaxf(df, 1.0f) : 1.0f / df;
                float labelf = fmaf((float)p->d_weight, penf, labf[cur][k]);
                int v = vv[cur][k], a = aa[cur][k];
                if (prune_fx && label + (hfx ? hfx[id] : 0) > prune_fx) continue;
                previous[id] = k - v; any = 1; n_nodes++;
                if (fmin_out) { uint64_t f = label + (hfx ? hfx[id] : 0); if (f < fmin_out[t]) fmin_out[t] = f; }
                int better = bk < 0 || (f32_labels ? labelf < blf : label < bl);
                if (better) { bk = k; bl = label; blf = labelf; }
                if (t == num_t - 1) continue;
                int al = a + jlo_c > alo_c ? a + jlo_c : alo_c, ah = a + jhi_c < ahi_c ? a + jhi_c : ahi_c;
                int vlo = v + al, vhi = v + ah;
                double s = s_values[k];
                if (vlo <= 0) { double me = (s - start_s) / delta_s; int mi = (int)me; if (mi < me) mi++; vlo = mi - k; }
                int clamp = vmax_is_int ? (vhi >= vmax_c) : ((double)v + fmin((double)a + jhi, ahi_r) > vmax_r);
                if (clamp) vhi = vmax_is_int ? (int)((s + p->max_speed * dt - start_s) / delta_s) - k : vmax_c;
                int wlo = k + vlo, whi = k + vhi; if (whi > num_s - 1) whi = num_s - 1;
                if (whi >= wlo) { if (wlo < klo_) klo_ = wlo; if (whi > khi_) khi_ = whi; }   /* the kernel's span: windows of kept nodes */
                for (int kk = wlo; kk <= whi; kk++) {
                    int vn = kk - k, an = vn - v, jn = an - a;
                    uint64_t tot = label + vtab[vn] + atab[an + 16] + jtab[jn + 8];
                    if (prune_fx && tot > prune_fx) continue;
                    n_push++;
                    if (has[nxt][kk] && (tot >> g_coll_shift) == (lab[nxt][kk] >> g_coll_shift)) n_coll++;
                    float fv = (float)vn - vdes, fa = (float)an, fj = (float)jn;
                    float totf = labelf + fmaf(cvf * fv, fv, fmaf(caf * fa, fa, cjf * fj * fj));
                    if (BETTER(tot, totf, vn, kk, nxt)) { lab[nxt][kk] = tot; labf[nxt][kk] = totf; vv[nxt][kk] = vn; aa[nxt][kk] = an; has[nxt][kk] = 1; }
                    if (kk < nlo) nlo = kk; if (kk > nhi) nhi = kk;
                }
            }
            if (khi_ - klo_ + 1 > max_span_) max_span_ = khi_ - klo_ + 1;
            if (!any) break;
            best_t = t; best_k = bk; best_lab = bl; best_labf = blf;
            dlo = nlo; dhi = nhi;
        }
    }
#undef BETTER
#undef FXQ
    if (cost_out) *cost_out = f32_labels ? (double)best_labf : (double)best_lab * (1.0 / FX1);
    if (counts) { counts[0] = n_nodes; counts[1] = n_push; }
    g_last_max_span = max_span_; g_last_collisions = n_coll;
    int r = backtrack(num_t, num_s, previous, s_values, best_t, best_k, idx_out, s_seq_out);
    for (int b = 0; b < 2; b++) { free(lab[b]); free(labf[b]); free(vv[b]); free(aa[b]); free(has[b]); }
    free(previous);
    return r;
}

int orc_solve_fast_model_ex(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                            const double *distances, const double *s_values, double delta_t,
                            double v0, double a0, int f32_labels, uint64_t prune_fx,
                            int *idx_out, double *s_seq_out, double *cost_out, int64_t *counts) {
    return fast_model_impl(p, num_t, num_s, obstacles, distances, s_values, delta_t, v0, a0, f32_labels, prune_fx, NULL, NULL,
                           idx_out, s_seq_out, cost_out, counts, 18, 0);
}

int orc_solve_fast_model_h(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                           const double *distances, const double *s_values, double delta_t,
                           double v0, double a0, uint64_t prune_fx, const uint64_t *hfx, uint64_t *fmin_out,
                           int *idx_out, double *s_seq_out, double *cost_out, int64_t *counts) {
    if (fmin_out) for (int t = 0; t < num_t; t++) fmin_out[t] = UINT64_MAX;
    return fast_model_impl(p, num_t, num_s, obstacles, distances, s_values, delta_t, v0, a0, 0, prune_fx, hfx, fmin_out,
                           idx_out, s_seq_out, cost_out, counts, 18, 0);
}

/* the 32-bit-key kernel's arithmetic: labels in 2^-frac_bits fixed point, min-combine on (label >> key_shift, larger v') */
int orc_solve_fast_model_q(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                           const double *distances, const double *s_values, double delta_t,
                           double v0, double a0, int frac_bits, int key_shift, uint64_t prune_fx,
                           int *idx_out, double *s_seq_out, double *cost_out, int64_t *counts) {
    return fast_model_impl(p, num_t, num_s, obstacles, distances, s_values, delta_t, v0, a0, 0, prune_fx, NULL, NULL,
                           idx_out, s_seq_out, cost_out, counts, frac_bits, key_shift);
}

int orc_solve_fast_model(const orc_params *p, int num_t, int num_s, const uint8_t *obstacles,
                         const double *distances, const double *s_values, double delta_t,
                         double v0, double a0, int f32_labels, int *idx_out, double *s_seq_out, double *cost_out) {
    return orc_solve_fast_model_ex(p, num_t, num_s, obstacles, distances, s_values, delta_t, v0, a0, f32_labels, 0,
                                   idx_out, s_seq_out, cost_out, NULL);
}

/* ---- cost of a given index path (the solver's own history convention) ----------------------- */
double orc_path_cost(const orc_params *p, int n, const int *idx, const double *s_values,
                     const double *distances, int num_s, double delta_t, double v0, double a0) {
    double start_s = s_values[0];
    double s1 = start_s - v0 * delta_t;
    double s2 = s1 - delta_t * (v0 - a0 * delta_t);
    double s = s_values[idx[0]], total = 0;
    for (int t = 1; t < n; t++) {
        if (idx[t] < 0) break;
        double ns = s_values[idx[t]];
        total = total + cost_with_jerk(p, ns, s, s1, s2, delta_t, distances[(size_t)t * num_s + idx[t]]);
        s2 = s1; s1 = s; s = ns;
    }
    return total;
}

/* ---- st.py:726-754 + 790-802 ---------------------------------------------------------------- */
int orc_plan(const orc_params *p, const orc_state *st, int use_layered, int *idx_out,
             double *s_seq_out, double *cost_out, int *guaranteed_crash, double *min_path_distance,
             double *start_s_out, double *delta_s_out, int *num_s_out, orc_solve_stats *stats) {
    double start_s = orc_get_ego_s(st->ego_x, st->ego_y);
    int num_t, num_s; double dse;
    orc_grid_dims(p, start_s, &num_t, &num_s, &dse);
    uint8_t *obs = (uint8_t *)malloc((size_t)num_t * num_s);
    double *dist = (double *)malloc(sizeof(double) * (size_t)num_t * num_s);
    double *sv = (double *)malloc(sizeof(double) * num_s);
    orc_build_grid(p, st, num_t, num_s, obs, dist, sv, NULL);
    double delta_t = p->t_disc;      /* t_values[1]-t_values[0] == t_disc (arange from 0) */
    int *idx = idx_out ? idx_out : (int *)malloc(sizeof(int) * num_t);
    double *seq = s_seq_out ? s_seq_out : (double *)malloc(sizeof(double) * num_t);
    int r = (use_layered ? orc_solve_layered : orc_solve_dijkstra)(
        p, num_t, num_s, obs, dist, sv, delta_t, st->ego_v, st->ego_a, idx, seq, cost_out, stats);
    /* st.py:790-802 */
    int end_point = num_t;
    while (end_point > 0 && seq[end_point - 1] == 0) end_point--;
    int crash = (end_point != num_t);
    double mind = INFINITY, ds = sv[1] - sv[0];
    if (!crash) {
        for (int i = 0; i < num_t; i++) {
            int si = (int)((seq[i] - sv[0]) / ds);
            double d = dist[(size_t)i * num_s + si];
            if (d < mind) mind = d;
            if (d < p->combination_min_distance - p->car_length) crash = 1;
        }
    }
    if (guaranteed_crash) *guaranteed_crash = crash;
    if (min_path_distance) *min_path_distance = mind;
    if (start_s_out) *start_s_out = start_s;
    if (delta_s_out) *delta_s_out = ds;
    if (num_s_out) *num_s_out = num_s;
    if (!idx_out) free(idx); if (!s_seq_out) free(seq);
    free(obs); free(dist); free(sv);
    return r;
}

/* ---- batch driver (CPU baseline) ------------------------------------------------------------ */
typedef struct batch_job {
    const orc_params *p; int B, nmax, lo, hi, use_layered, num_t;
    const double *ego, *cx, *cv, *ca; const int *n;
    int *idx; double *seq, *cost; int *reached, *crash; double *mind;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    for (int b = j->lo; b < j->hi; b++) {
        orc_state st;
        st.ego_x = j->ego[4 * b]; st.ego_y = j->ego[4 * b + 1]; st.ego_v = j->ego[4 * b + 2]; st.ego_a = j->ego[4 * b + 3];
        st.n = j->n[b] > ORC_MAX_CARS ? ORC_MAX_CARS : j->n[b];
        for (int c = 0; c < st.n; c++) {
            st.x[c] = j->cx[(size_t)b * j->nmax + c]; st.v[c] = j->cv[(size_t)b * j->nmax + c];
            st.a[c] = j->ca[(size_t)b * j->nmax + c];
        }
        double cost, mind; int crash;
        int r = orc_plan(j->p, &st, j->use_layered, j->idx ? j->idx + (size_t)b * j->num_t : NULL,
                         j->seq ? j->seq + (size_t)b * j->num_t : NULL, &cost, &crash, &mind, NULL, NULL, NULL, NULL);
        if (j->cost) j->cost[b] = cost; if (j->reached) j->reached[b] = r;
        if (j->crash) j->crash[b] = crash; if (j->mind) j->mind[b] = mind;
    }
    return NULL;
}

void orc_plan_batch(const orc_params *p, int B, int nmax, const double *ego, const double *cars_x,
                    const double *cars_v, const double *cars_a, const int *n_cars, int use_layered,
                    int nthreads, int num_t, int *idx_out, double *s_seq_out, double *cost_out,
                    int *reached_out, int *crash_out, double *min_dist_out) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    batch_job *jobs = (batch_job *)malloc(sizeof(batch_job) * nthreads);
    for (int i = 0; i < nthreads; i++) {
        batch_job j = {p, B, nmax, (int)((long)B * i / nthreads), (int)((long)B * (i + 1) / nthreads),
                       use_layered, num_t, ego, cars_x, cars_v, cars_a, n_cars,
                       idx_out, s_seq_out, cost_out, reached_out, crash_out, min_dist_out};
        jobs[i] = j;
        if (nthreads == 1) batch_worker(&jobs[i]);
        else pthread_create(&th[i], NULL, batch_worker, &jobs[i]);
    }
    if (nthreads > 1) for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    free(th); free(jobs);
}

/* ---- control.py:160-171 --------------------------------------------------------------------- */
double orc_speed_from_jerk(const orc_params *p, double v, double a, double jerk) {
    double na = a + jerk * p->tick_length;
    if (na > p->a_max) na = p->a_max;
    if (na < p->a_min) na = p->a_min;
    double nv = v + na * p->tick_length;
    if (nv > p->max_speed) nv = p->max_speed;
    if (nv < 0) nv = 0;
    return nv;
}

/* ---- dqn.py:389-446 (CARS_AHEAD=CARS_BEHIND=2, accel+speed-difference+normalised) ----------- */
void orc_state_vector(const orc_params *p, const orc_state *st, double *out) {
    for (int i = 0; i < 20; i++) out[i] = 0.0;
    int nf = 0, first_back = st->n;
    /* cars with x > ego_x are "front" (list order is far->near, reversed to near->far); else back */
    int fidx[ORC_MAX_CARS], bidx[ORC_MAX_CARS], nb = 0;
    for (int i = 0; i < st->n; i++) { if (st->x[i] > st->ego_x) fidx[nf++] = i; else bidx[nb++] = i; }
    (void)first_back;
    for (int s = 0; s < 2; s++) {
        if (s < nf) {
            int i = fidx[nf - 1 - s];
            out[4 * s + 0] = st->a[i] / 9; out[4 * s + 1] = (st->v[i] - st->ego_v) / p->max_speed;
            out[4 * s + 2] = (st->x[i] - st->ego_x) / p->sensor_radius; out[4 * s + 3] = 1;
        }
        if (s < nb) {
            int i = bidx[s];
            out[8 + 4 * s + 0] = st->a[i] / 9; out[8 + 4 * s + 1] = (st->v[i] - st->ego_v) / p->max_speed;
            out[8 + 4 * s + 2] = (st->x[i] - st->ego_x) / p->sensor_radius; out[8 + 4 * s + 3] = 1;
        }
    }
    out[16] = st->ego_v / p->max_speed; out[17] = st->ego_a / 9;
    out[18] = st->ego_x / 300; out[19] = st->ego_y / 100;
}

/* ---- st.py:274-288 -------------------------------------------------------------------------- */
double orc_path_mean_abs_jerk(const double *s, int n, double v0, double a0, double dt) {
    double prev_a = a0, prev_v = v0, acc = 0;
    for (int i = 1; i < n; i++) {
        double v = (s[i] - s[i - 1]) / dt, a = (v - prev_v) / dt, j = (a - prev_a) / dt;
        prev_v = v; prev_a = a; acc += fabs(j);
    }
    return acc / (n - 1);
}

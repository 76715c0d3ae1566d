// Batched `finer_fit` (reference st.py:584-723): re-sample the coarse plan (T_DISCRETIZATION) onto the control tick
// (TICK_LENGTH) by linear interpolation, then project it onto the speed / acceleration / jerk limits:
//
//     minimise ||x - b||^2   s.t.  0 <= v_i <= v_max,  a_min <= a_i <= a_max,  j_min <= j_i <= j_max,  x_0 fixed
//
// where v, a, j are first/second/third backward differences of x whose first entries use the current ego speed and
// acceleration (st.py:626-668).  With the virtual history x_{-1} = x_0 - v0*dt, x_{-2} = x_{-1} - (v0 - a0*dt)*dt all
// three constraint families are plain finite differences of the extended sequence, i.e. l <= A x <= u with A = [D1;
// D2; D3] (rows scaled to position units) over the m = n-1 unknown positions.
//
// The reference hands this QP to cvxopt's interior-point solver (maxiters = 10).  cvxopt is not available and its
// iterates are not reproducible bit for bit; here the same convex QP is solved by a Mehrotra predictor-corrector
// primal-dual interior-point method (the same family, ~12 iterations to 1e-9): with G = [A; -A] the Newton system
// reduces to (I + A^T diag(w) A) dx = g, which is banded (half-bandwidth 3) and is factorised in place per iteration.
// One thread per episode (m <= 25 unknowns at the published discretisation).  Parity vs cvxopt is UNPINNED (SURVEY.md
// §8c); the tests check the KKT conditions and agreement with an independent CPU solve (scipy SLSQP).
#include "mpc_common.cuh"

#define QP_MAXN 40                 // fine points per plan supported (published config: 26)
#define QP_M (QP_MAXN - 1)
#define QP_R (3 * QP_M)

// rows of A over the unknowns u[0..m-1] (u[i] = x_{i+1}):  r1[i] = u[i]-u[i-1],  r2[i] = u[i]-2u[i-1]+u[i-2],
// r3[i] = u[i]-3u[i-1]+3u[i-2]-u[i-3]; entries with a negative index belong to the fixed history and are constants.
__device__ __forceinline__ void qp_apply_A(int m, const double *u, double *r) {
    for (int i = 0; i < m; i++) {
        double u0 = u[i], u1 = i >= 1 ? u[i - 1] : 0.0, u2 = i >= 2 ? u[i - 2] : 0.0, u3 = i >= 3 ? u[i - 3] : 0.0;
        r[i] = u0 - u1; r[m + i] = u0 - 2.0 * u1 + u2; r[2 * m + i] = u0 - 3.0 * u1 + 3.0 * u2 - u3;
    }
}
__device__ __forceinline__ void qp_apply_AT(int m, const double *w, double *g) {      // g += A^T w
    for (int i = 0; i < m; i++) {
        double w1 = w[i], w2 = w[m + i], w3 = w[2 * m + i];
        g[i] += w1 + w2 + w3;
        if (i >= 1) g[i - 1] += -w1 - 2.0 * w2 - 3.0 * w3;
        if (i >= 2) g[i - 2] += w2 + 3.0 * w3;
        if (i >= 3) g[i - 3] += -w3;
    }
}
// banded SPD solve: M = I + A^T diag(w) A  (Mb[i][d] = M(i, i-d), d = 0..3); factorised in place, then L L^T x = g
__device__ __forceinline__ void qp_factor(int m, const double *w, double (*Mb)[4]) {
    for (int i = 0; i < m; i++) { Mb[i][0] = 1.0; Mb[i][1] = 0.0; Mb[i][2] = 0.0; Mb[i][3] = 0.0; }
    const double c1[2] = {1, -1}, c2[3] = {1, -2, 1}, c3[4] = {1, -3, 3, -1};
    for (int i = 0; i < m; i++) {
        double w1 = w[i], w2 = w[m + i], w3 = w[2 * m + i];
        for (int p = 0; p < 4; p++) for (int q = p; q < 4; q++) {          // column i-p (row index), column i-q: i-p >= i-q
            if (i - q < 0) continue;
            double v = w3 * c3[p] * c3[q];
            if (q < 3) v += w2 * c2[p] * c2[q];
            if (q < 2) v += w1 * c1[p] * c1[q];
            Mb[i - p][q - p] += v;                                        // M(i-p, i-q), band offset (i-p)-(i-q) = q-p
        }
    }
    for (int i = 0; i < m; i++) {
        for (int d = 3; d >= 0; d--) {
            int j = i - d;
            if (j < 0) { Mb[i][d] = 0.0; continue; }
            double s = Mb[i][d];
            for (int k = (i - 3 > 0 ? i - 3 : 0); k < j; k++) s -= Mb[i][i - k] * Mb[j][j - k];
            Mb[i][d] = (d == 0) ? sqrt(s) : s / Mb[j][0];
        }
    }
}
__device__ __forceinline__ void qp_solve(int m, const double (*Lb)[4], double *g) {
    for (int i = 0; i < m; i++) {
        double s = g[i];
        for (int d = 1; d < 4 && i - d >= 0; d++) s -= Lb[i][d] * g[i - d];
        g[i] = s / Lb[i][0];
    }
    for (int i = m - 1; i >= 0; i--) {
        double s = g[i];
        for (int d = 1; d < 4 && i + d < m; d++) s -= Lb[i + d][d] * g[i + d];
        g[i] = s / Lb[i][0];
    }
}

__global__ void __launch_bounds__(64) finer_fit_kernel(DevParams P, int B, int T, const double *__restrict__ s_seq,
                                                       const int32_t *__restrict__ reached, const double *__restrict__ ego,
                                                       int max_iter, double tol, double *__restrict__ fine, int fine_stride,
                                                       int32_t *__restrict__ n_fine, double *__restrict__ speed,
                                                       int32_t *__restrict__ iters_out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double dt = P.p.tick_length, dtc = P.p.t_disc;
    const int Lc = reached[b] + 1;                               // coarse points that exist (st.py:762-768 trims the 0.0 tail)
    const double *sc = s_seq + (size_t)b * T;
    const double v0 = ego[4 * b + 2], a0 = ego[4 * b + 3];
    double *out = fine + (size_t)b * fine_stride;
    int n = 1;
    if (Lc > 1) {                                                // st.py:590-594
        const double t_last = (double)(Lc - 1) * dtc;
        n = (int)rint(t_last / dt + 1.0);
        if ((double)(n - 1) * dt > t_last) n -= 1;
        if (n > QP_MAXN) n = QP_MAXN;
        if (n > fine_stride) n = fine_stride;
    }
    const int m = n - 1;
    if (m <= 0) {                                                // st.py:587-588, 775-777: keep the current speed
        out[0] = sc[0]; n_fine[b] = 1; if (speed) speed[b] = v0; if (iters_out) iters_out[b] = 0;
        return;
    }
    // st.py:597-598: linear interpolation (scipy interp1d: slope * (x - x_lo) + y_lo on the bracketing interval)
    double bv[QP_M];
    for (int i = 1; i < n; i++) {
        double tf = (double)i * dt;
        int hi = 1;
        while (hi < Lc - 1 && (double)hi * dtc < tf) hi++;
        int lo = hi - 1;
        double xl = (double)lo * dtc, xh = (double)hi * dtc;
        bv[i - 1] = (sc[hi] - sc[lo]) / (xh - xl) * (tf - xl) + sc[lo];
    }
    // bounds of the rows with the fixed history (x_{-2}, x_{-1}, x_0; st.py:626-668) folded in
    const double f2 = sc[0], f1 = f2 - v0 * dt, f0 = f1 - (v0 - a0 * dt) * dt;
    double lb[QP_R], ub[QP_R];
    {
        const double lo1 = 0.0, hi1 = P.p.max_speed * dt, lo2 = P.p.a_min * dt * dt, hi2 = P.p.a_max * dt * dt;
        const double lo3 = P.p.j_min * dt * dt * dt, hi3 = P.p.j_max * dt * dt * dt;
        for (int i = 0; i < m; i++) {
            double k1 = (i == 0) ? -f2 : 0.0;
            double k2 = (i == 0) ? (-2.0 * f2 + f1) : (i == 1 ? f2 : 0.0);
            double k3 = (i == 0) ? (-3.0 * f2 + 3.0 * f1 - f0) : (i == 1 ? (3.0 * f2 - f1) : (i == 2 ? -f2 : 0.0));
            lb[i] = lo1 - k1; ub[i] = hi1 - k1; lb[m + i] = lo2 - k2; ub[m + i] = hi2 - k2; lb[2 * m + i] = lo3 - k3; ub[2 * m + i] = hi3 - k3;
        }
    }
    const int R = 3 * m;
    double x[QP_M], r[QP_R], su[QP_R], sl[QP_R], zu[QP_R], zl[QP_R], w[QP_R], dsu[QP_R], dsl[QP_R], dzu[QP_R], dzl[QP_R], g[QP_M], tcu[QP_R], tcl[QP_R];
    double Mb[QP_M][4];
    for (int i = 0; i < m; i++) x[i] = bv[i];
    qp_apply_A(m, x, r);
    for (int i = 0; i < R; i++) { su[i] = fmax(ub[i] - r[i], 0.1); sl[i] = fmax(r[i] - lb[i], 0.1); zu[i] = 1.0; zl[i] = 1.0; }
    int it = 0;
    for (; it < max_iter; it++) {
        qp_apply_A(m, x, r);
        double mu = 0.0, res = 0.0;
        for (int i = 0; i < R; i++) { mu += su[i] * zu[i] + sl[i] * zl[i]; res = fmax(res, fmax(fabs(r[i] + su[i] - ub[i]), fabs(-r[i] + sl[i] + lb[i]))); }
        mu /= (double)(2 * R);
        // rd = x - b + A^T (zu - zl)
        double rd[QP_M];
        for (int i = 0; i < m; i++) rd[i] = x[i] - bv[i];
        for (int i = 0; i < R; i++) w[i] = zu[i] - zl[i];
        qp_apply_AT(m, w, rd);
        for (int i = 0; i < m; i++) res = fmax(res, fabs(rd[i]));
        if ((res < tol && mu < tol) || mu < 1e-13) break;
        for (int i = 0; i < R; i++) w[i] = zu[i] / su[i] + zl[i] / sl[i];
        qp_factor(m, w, Mb);
        double sigma_mu = 0.0;
        for (int pass = 0; pass < 2; pass++) {
            // complementarity targets: predictor su*zu ; corrector su*zu + dsu*dzu - sigma*mu
            for (int i = 0; i < m; i++) g[i] = rd[i];
            for (int i = 0; i < R; i++) {
                double rpu = r[i] + su[i] - ub[i], rpl = -r[i] + sl[i] + lb[i];
                double rcu = su[i] * zu[i], rcl = sl[i] * zl[i];
                if (pass) { rcu += dsu[i] * dzu[i] - sigma_mu; rcl += dsl[i] * dzl[i] - sigma_mu; }
                w[i] = (-rcu + zu[i] * rpu) / su[i] - (-rcl + zl[i] * rpl) / sl[i];
                tcu[i] = rcu; tcl[i] = rcl;
            }
            qp_apply_AT(m, w, g);
            for (int i = 0; i < m; i++) g[i] = -g[i];
            qp_solve(m, Mb, g);                                // g = dx
            double ax_tmp[QP_R];
            qp_apply_A(m, g, ax_tmp);
            double ap = 1.0, ad = 1.0;
            for (int i = 0; i < R; i++) {
                double rpu = r[i] + su[i] - ub[i], rpl = -r[i] + sl[i] + lb[i];
                double d_su = -rpu - ax_tmp[i], d_sl = -rpl + ax_tmp[i];
                double d_zu = -(tcu[i] + zu[i] * d_su) / su[i], d_zl = -(tcl[i] + zl[i] * d_sl) / sl[i];
                dsu[i] = d_su; dsl[i] = d_sl; dzu[i] = d_zu; dzl[i] = d_zl;
                if (d_su < 0.0) ap = fmin(ap, -su[i] / d_su);
                if (d_sl < 0.0) ap = fmin(ap, -sl[i] / d_sl);
                if (d_zu < 0.0) ad = fmin(ad, -zu[i] / d_zu);
                if (d_zl < 0.0) ad = fmin(ad, -zl[i] / d_zl);
            }
            if (pass == 0) {
                double mu_aff = 0.0;
                for (int i = 0; i < R; i++) mu_aff += (su[i] + ap * dsu[i]) * (zu[i] + ad * dzu[i]) + (sl[i] + ap * dsl[i]) * (zl[i] + ad * dzl[i]);
                mu_aff /= (double)(2 * R);
                double sg = mu_aff / mu; sigma_mu = sg * sg * sg * mu;
            } else {
                // fraction to the boundary 0.995 (the unit step is kept when no slack / multiplier blocks it)
                double apf = 1.0e9, adf = 1.0e9;
                for (int i = 0; i < R; i++) {
                    if (dsu[i] < 0.0) apf = fmin(apf, -su[i] / dsu[i]);
                    if (dsl[i] < 0.0) apf = fmin(apf, -sl[i] / dsl[i]);
                    if (dzu[i] < 0.0) adf = fmin(adf, -zu[i] / dzu[i]);
                    if (dzl[i] < 0.0) adf = fmin(adf, -zl[i] / dzl[i]);
                }
                ap = fmin(1.0, 0.995 * apf); ad = fmin(1.0, 0.995 * adf);
                for (int i = 0; i < m; i++) x[i] += ap * g[i];
                for (int i = 0; i < R; i++) { su[i] += ap * dsu[i]; sl[i] += ap * dsl[i]; zu[i] += ad * dzu[i]; zl[i] += ad * dzl[i]; }
            }
        }
    }
    out[0] = sc[0];
    for (int i = 0; i < m; i++) out[i + 1] = x[i];
    n_fine[b] = n;
    if (speed) speed[b] = (out[1] - out[0]) / dt;                  // st.py:780-781
    if (iters_out) iters_out[b] = it;
}

cudaError_t launch_finer_fit(const DevParams &P, int B, int T, const double *s_seq, const int32_t *reached, const double *ego,
                             int max_iter, double tol, double *fine, int fine_stride, int32_t *n_fine, double *speed,
                             int32_t *iters, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    finer_fit_kernel<<<(B + 63) / 64, 64, 0, st>>>(P, B, T, s_seq, reached, ego, max_iter, tol, fine, fine_stride, n_fine, speed, iters);
    return cudaGetLastError();
}

int qp_max_fine() { return QP_MAXN; }

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
// One warp per episode (m <= 25 unknowns at the published discretisation).  Parity vs cvxopt is UNPINNED (SURVEY.md
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

// ---- one WARP per episode ------------------------------------------------------------------------------------------
// The rows of A (<= 117) are spread over the lanes, the vectors live in shared memory; only the banded Cholesky factorisation
// and the two triangular solves (m <= 39 steps, half-bandwidth 3) are sequential and done by lane 0.  (The first version ran one
// THREAD per episode with ~13 KB of local-memory arrays: 7.2 ms for the ~1200 take-over episodes of a closed-loop tick, 70 % of
// the tick; this one takes tens of microseconds.)
#define QP_WARPS 2
struct QpWarpShared {
    double x[QP_MAXN], g[QP_MAXN], rd[QP_MAXN], bv[QP_MAXN];
    double r[QP_R], w[QP_R], ax[QP_R], su[QP_R], sl[QP_R], zu[QP_R], zl[QP_R], lb[QP_R], ub[QP_R];
    double dsu[QP_R], dsl[QP_R], dzu[QP_R], dzl[QP_R], tcu[QP_R], tcl[QP_R];
    double Mb[QP_M][4];
};
__device__ __forceinline__ double wsum(double v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
__device__ __forceinline__ double wmax(double v) { for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ double wmin(double v) { for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }

// row rr = (family f, position i) of A applied to u (see qp_apply_A)
__device__ __forceinline__ double qp_row(int m, int rr, const double *u) {
    const int f = rr / m, i = rr - f * m;
    const double u0 = u[i], u1 = i >= 1 ? u[i - 1] : 0.0, u2 = i >= 2 ? u[i - 2] : 0.0, u3 = i >= 3 ? u[i - 3] : 0.0;
    return f == 0 ? u0 - u1 : (f == 1 ? u0 - 2.0 * u1 + u2 : u0 - 3.0 * u1 + 3.0 * u2 - u3);
}
// (A^T w)[i]  (see qp_apply_AT)
__device__ __forceinline__ double qp_col(int m, int i, const double *w) {
    const double *w1 = w, *w2 = w + m, *w3 = w + 2 * m;
    double s = w1[i] + w2[i] + w3[i];
    if (i + 1 < m) s += -w1[i + 1] - 2.0 * w2[i + 1] - 3.0 * w3[i + 1];
    if (i + 2 < m) s += w2[i + 2] + 3.0 * w3[i + 2];
    if (i + 3 < m) s += -w3[i + 3];
    return s;
}

// SUBSET: only the episodes listed in subset[0 .. *count) are fitted (masked calls; the default instance is unchanged)
template <bool SUBSET>
__global__ void __launch_bounds__(32 * QP_WARPS) finer_fit_kernel(DevParams P, int B, int T, const double *__restrict__ s_seq,
                                                                 const int32_t *__restrict__ reached, const double *__restrict__ ego,
                                                                 int max_iter, double tol, double *__restrict__ fine, int fine_stride,
                                                                 int32_t *__restrict__ n_fine, double *__restrict__ speed,
                                                                 int32_t *__restrict__ iters_out, const int32_t *__restrict__ subset,
                                                                 const int *__restrict__ count) {
    __shared__ QpWarpShared SH[QP_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * QP_WARPS + wib;
    if (w >= B) return;                                          // (whole warp)
    if (SUBSET && w >= *count) return;
    const int b = SUBSET ? subset[w] : w;
    QpWarpShared &S = SH[wib];
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
        if (lane == 0) { out[0] = sc[0]; n_fine[b] = 1; if (speed) speed[b] = v0; if (iters_out) iters_out[b] = 0; }
        return;
    }
    const int R = 3 * m;
    // st.py:597-598: linear interpolation (scipy interp1d: slope * (x - x_lo) + y_lo on the bracketing interval)
    for (int i = 1 + lane; i < n; i += 32) {
        double tf = (double)i * dt;
        int hi = 1;
        while (hi < Lc - 1 && (double)hi * dtc < tf) hi++;
        int lo = hi - 1;
        double xl = (double)lo * dtc, xh = (double)hi * dtc;
        S.bv[i - 1] = (sc[hi] - sc[lo]) / (xh - xl) * (tf - xl) + sc[lo];
        S.x[i - 1] = S.bv[i - 1];
    }
    // bounds of the rows with the fixed history (x_{-2}, x_{-1}, x_0; st.py:626-668) folded in
    const double f2 = sc[0], f1 = f2 - v0 * dt, f0 = f1 - (v0 - a0 * dt) * dt;
    {
        const double lo1 = 0.0, hi1 = P.p.max_speed * dt, lo2 = P.p.a_min * dt * dt, hi2 = P.p.a_max * dt * dt;
        const double lo3 = P.p.j_min * dt * dt * dt, hi3 = P.p.j_max * dt * dt * dt;
        for (int rr = lane; rr < R; rr += 32) {
            const int f = rr / m, i = rr - f * m;
            double k, lo, hi;
            if (f == 0) { k = (i == 0) ? -f2 : 0.0; lo = lo1; hi = hi1; }
            else if (f == 1) { k = (i == 0) ? (-2.0 * f2 + f1) : (i == 1 ? f2 : 0.0); lo = lo2; hi = hi2; }
            else { k = (i == 0) ? (-3.0 * f2 + 3.0 * f1 - f0) : (i == 1 ? (3.0 * f2 - f1) : (i == 2 ? -f2 : 0.0)); lo = lo3; hi = hi3; }
            S.lb[rr] = lo - k; S.ub[rr] = hi - k;
        }
    }
    __syncwarp();
    for (int rr = lane; rr < R; rr += 32) {
        const double rv = qp_row(m, rr, S.x);
        S.su[rr] = fmax(S.ub[rr] - rv, 0.1); S.sl[rr] = fmax(rv - S.lb[rr], 0.1); S.zu[rr] = 1.0; S.zl[rr] = 1.0;
    }
    __syncwarp();
    int it = 0;
    for (; it < max_iter; it++) {
        double mu = 0.0, res = 0.0;
        for (int rr = lane; rr < R; rr += 32) {
            const double rv = qp_row(m, rr, S.x);
            S.r[rr] = rv;
            mu += S.su[rr] * S.zu[rr] + S.sl[rr] * S.zl[rr];
            res = fmax(res, fmax(fabs(rv + S.su[rr] - S.ub[rr]), fabs(-rv + S.sl[rr] + S.lb[rr])));
            S.w[rr] = S.zu[rr] - S.zl[rr];
        }
        __syncwarp();
        mu = wsum(mu) / (double)(2 * R);
        for (int i = lane; i < m; i += 32) {                    // rd = x - b + A^T (zu - zl)
            const double v = S.x[i] - S.bv[i] + qp_col(m, i, S.w);
            S.rd[i] = v; res = fmax(res, fabs(v));
        }
        res = wmax(res);
        if ((res < tol && mu < tol) || mu < 1e-13) break;        // (uniform)
        __syncwarp();
        for (int rr = lane; rr < R; rr += 32) S.w[rr] = S.zu[rr] / S.su[rr] + S.zl[rr] / S.sl[rr];
        __syncwarp();
        // M = I + A^T diag(w) A, band storage Mb[i][d] = M(i, i-d)
        for (int i = lane; i < m; i += 32) {
            const double c1[2] = {1, -1}, c2[3] = {1, -2, 1}, c3[4] = {1, -3, 3, -1};
            for (int d = 0; d < 4; d++) {
                double v = (d == 0) ? 1.0 : 0.0;
                if (i - d >= 0)
                    for (int p = 0; p + d < 4 && i + p < m; p++) {
                        v += S.w[2 * m + i + p] * c3[p] * c3[p + d];
                        if (p + d < 3) v += S.w[m + i + p] * c2[p] * c2[p + d];
                        if (p + d < 2) v += S.w[i + p] * c1[p] * c1[p + d];
                    }
                S.Mb[i][d] = v;
            }
        }
        __syncwarp();
        if (lane == 0) {                                        // banded Cholesky in place (sequential)
            for (int i = 0; i < m; i++)
                for (int d = 3; d >= 0; d--) {
                    const int j = i - d;
                    if (j < 0) { S.Mb[i][d] = 0.0; continue; }
                    double sv = S.Mb[i][d];
                    for (int k = (i - 3 > 0 ? i - 3 : 0); k < j; k++) sv -= S.Mb[i][i - k] * S.Mb[j][j - k];
                    S.Mb[i][d] = (d == 0) ? sqrt(sv) : sv / S.Mb[j][0];
                }
        }
        __syncwarp();
        double sigma_mu = 0.0;
        for (int pass = 0; pass < 2; pass++) {
            // complementarity targets: predictor su*zu ; corrector su*zu + dsu*dzu - sigma*mu
            for (int rr = lane; rr < R; rr += 32) {
                const double rv = S.r[rr];
                const double rpu = rv + S.su[rr] - S.ub[rr], rpl = -rv + S.sl[rr] + S.lb[rr];
                double rcu = S.su[rr] * S.zu[rr], rcl = S.sl[rr] * S.zl[rr];
                if (pass) { rcu += S.dsu[rr] * S.dzu[rr] - sigma_mu; rcl += S.dsl[rr] * S.dzl[rr] - sigma_mu; }
                S.w[rr] = (-rcu + S.zu[rr] * rpu) / S.su[rr] - (-rcl + S.zl[rr] * rpl) / S.sl[rr];
                S.tcu[rr] = rcu; S.tcl[rr] = rcl;
            }
            __syncwarp();
            for (int i = lane; i < m; i += 32) S.g[i] = -(S.rd[i] + qp_col(m, i, S.w));
            __syncwarp();
            if (lane == 0) qp_solve(m, S.Mb, S.g);              // g = dx
            __syncwarp();
            double ap = 1.0, ad = 1.0;
            for (int rr = lane; rr < R; rr += 32) {
                const double rv = S.r[rr], axv = qp_row(m, rr, S.g);
                const double rpu = rv + S.su[rr] - S.ub[rr], rpl = -rv + S.sl[rr] + S.lb[rr];
                const double d_su = -rpu - axv, d_sl = -rpl + axv;
                const double d_zu = -(S.tcu[rr] + S.zu[rr] * d_su) / S.su[rr], d_zl = -(S.tcl[rr] + S.zl[rr] * d_sl) / S.sl[rr];
                S.dsu[rr] = d_su; S.dsl[rr] = d_sl; S.dzu[rr] = d_zu; S.dzl[rr] = d_zl;
                if (d_su < 0.0) ap = fmin(ap, -S.su[rr] / d_su);
                if (d_sl < 0.0) ap = fmin(ap, -S.sl[rr] / d_sl);
                if (d_zu < 0.0) ad = fmin(ad, -S.zu[rr] / d_zu);
                if (d_zl < 0.0) ad = fmin(ad, -S.zl[rr] / d_zl);
            }
            ap = wmin(ap); ad = wmin(ad);
            if (pass == 0) {
                double mu_aff = 0.0;
                for (int rr = lane; rr < R; rr += 32)
                    mu_aff += (S.su[rr] + ap * S.dsu[rr]) * (S.zu[rr] + ad * S.dzu[rr]) + (S.sl[rr] + ap * S.dsl[rr]) * (S.zl[rr] + ad * S.dzl[rr]);
                mu_aff = wsum(mu_aff) / (double)(2 * R);
                const double sg = mu_aff / mu; sigma_mu = sg * sg * sg * mu;
            } else {
                // fraction to the boundary 0.995 (the unit step is kept when no slack / multiplier blocks it)
                double apf = 1.0e9, adf = 1.0e9;
                for (int rr = lane; rr < R; rr += 32) {
                    if (S.dsu[rr] < 0.0) apf = fmin(apf, -S.su[rr] / S.dsu[rr]);
                    if (S.dsl[rr] < 0.0) apf = fmin(apf, -S.sl[rr] / S.dsl[rr]);
                    if (S.dzu[rr] < 0.0) adf = fmin(adf, -S.zu[rr] / S.dzu[rr]);
                    if (S.dzl[rr] < 0.0) adf = fmin(adf, -S.zl[rr] / S.dzl[rr]);
                }
                apf = wmin(apf); adf = wmin(adf);
                ap = fmin(1.0, 0.995 * apf); ad = fmin(1.0, 0.995 * adf);
                __syncwarp();
                for (int i = lane; i < m; i += 32) S.x[i] += ap * S.g[i];
                for (int rr = lane; rr < R; rr += 32) { S.su[rr] += ap * S.dsu[rr]; S.sl[rr] += ap * S.dsl[rr]; S.zu[rr] += ad * S.dzu[rr]; S.zl[rr] += ad * S.dzl[rr]; }
            }
            __syncwarp();
        }
    }
    __syncwarp();
    // Guard.  The QP is infeasible when the fixed history already breaks a limit that the first rows cannot repair -- an ego that
    // has just braked to a standstill with a0 < ~-1 (speed, acceleration and jerk rows of step 0 contradict each other), or one
    // near MAX_SPEED with a0 > 0; both are reachable in closed loop.  The slacks then collapse and the iterate turns into NaN
    // (cvxopt in the reference stops after maxiters = 10 with a finite, slightly infeasible iterate; st.py:16-17).  An iterate
    // that is not finite or breaks a limit by more than the tolerance is replaced by the interpolated plan tracked step by step
    // through the limits in the order jerk, acceleration, speed (control.py:160-171 applied to every step): always finite,
    // always inside the limits the environment enforces.  iters_out = max_iter + 1 marks such an episode.
    {
        double viol = 0.0;
        for (int rr = lane; rr < R; rr += 32) {
            const double rv = qp_row(m, rr, S.x);
            const double e = fmax(rv - S.ub[rr], S.lb[rr] - rv);
            viol = (e == e) ? fmax(viol, e) : 1.0e300;            // (NaN -> violation)
        }
        viol = wmax(viol);
        if (!(viol <= 10.0 * tol + 1.0e-9)) {
            if (lane == 0) {
                double prev = sc[0], v = v0, a = a0;
                for (int i = 0; i < m; i++) {
                    const double vt = (S.bv[i] - prev) / dt;
                    double jk = ((vt - v) / dt - a) / dt;
                    jk = jk < P.p.j_min ? P.p.j_min : (jk > P.p.j_max ? P.p.j_max : jk);
                    double an = a + jk * dt;
                    an = an < P.p.a_min ? P.p.a_min : (an > P.p.a_max ? P.p.a_max : an);
                    double vn = v + an * dt;
                    if (vn < 0.0) { vn = 0.0; an = (vn - v) / dt; }
                    else if (vn > P.p.max_speed) { vn = P.p.max_speed; an = (vn - v) / dt; }
                    prev += vn * dt; S.x[i] = prev; v = vn; a = an;
                }
            }
            it = max_iter + 1;
            __syncwarp();
        }
    }
    if (lane == 0) out[0] = sc[0];
    for (int i = lane; i < m; i += 32) out[i + 1] = S.x[i];
    if (lane == 0) {
        n_fine[b] = n;
        if (speed) speed[b] = (S.x[0] - sc[0]) / dt;              // st.py:780-781
        if (iters_out) iters_out[b] = it;
    }
}

cudaError_t launch_finer_fit(const DevParams &P, int B, int T, const double *s_seq, const int32_t *reached, const double *ego,
                             int max_iter, double tol, double *fine, int fine_stride, int32_t *n_fine, double *speed,
                             int32_t *iters, cudaStream_t st, const int32_t *subset, const int *count) {
    if (B <= 0) return cudaSuccess;
    if (subset) MPC_LAUNCH(finer_fit_kernel<true>, (B + QP_WARPS - 1) / QP_WARPS, 32 * QP_WARPS, 0, st, P, B, T, s_seq, reached, ego, max_iter, tol, fine, fine_stride, n_fine, speed, iters, subset, count);
    else MPC_LAUNCH(finer_fit_kernel<false>, (B + QP_WARPS - 1) / QP_WARPS, 32 * QP_WARPS, 0, st, P, B, T, s_seq, reached, ego, max_iter, tol, fine, fine_stride, n_fine, speed, iters, subset, count);
    return cudaGetLastError();
}

int qp_max_fine() { return QP_MAXN; }

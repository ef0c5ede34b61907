/*
 * oracle/pmv_oracle_ba.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).   *** PARITY UNPINNED ***
 *
 * Restates the bundle adjustment the reference runs through Ceres:
 *   residual          /root/reference/include/ProjectionResidual.h:38-58 (templated functor) evaluated
 *                     with forward-mode dual numbers ("Jets") exactly as
 *                     ceres::AutoDiffCostFunction<ProjectionResidual,2,6,3> does
 *                     (/root/reference/ProjectionResidual.cpp:3-8)
 *   problem / solver  /root/reference/CeresBundleAdjustment.cpp:26-61: pose block [rodrigues(R^T), -t],
 *                     point block X, HuberLoss(1.0), SPARSE_SCHUR, max_num_iterations, all else default
 *   minimiser         Ceres >= 1.13 (README.md:9; NOT vendored, NOT installable here): trust-region
 *                     Levenberg-Marquardt with Jacobi scaling, Schur elimination of the points,
 *                     Cholesky of the reduced camera system (SURVEY.md Appendix C restates the
 *                     published algorithm and defaults).
 * No Ceres binary, source or golden vector exists in the image and the reference has no tests, so this
 * oracle cannot be pinned against Ceres: "parity unpinned" (DESIGN.md).  What pins it instead
 * (tests/test_oracle_ba.py): Jacobians vs torch.autograd (fp64) and finite differences of the same
 * expression; Schur LM vs dense normal-equation LM (two independent linear-algebra paths) to 1e-10;
 * cost decrease / convergence on problems with known minima.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ Jets (9 partials) ---- */
#define NJ 9
typedef struct { double a; double v[NJ]; } jet;

static inline jet jconst(double a) { jet r; r.a = a; memset(r.v, 0, sizeof r.v); return r; }
static inline jet jvar(double a, int k) { jet r = jconst(a); r.v[k] = 1.0; return r; }
static inline jet jadd(jet x, jet y) { jet r; r.a = x.a + y.a; for (int i = 0; i < NJ; i++) r.v[i] = x.v[i] + y.v[i]; return r; }
static inline jet jsub(jet x, jet y) { jet r; r.a = x.a - y.a; for (int i = 0; i < NJ; i++) r.v[i] = x.v[i] - y.v[i]; return r; }
static inline jet jmul(jet x, jet y) { jet r; r.a = x.a * y.a; for (int i = 0; i < NJ; i++) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
static inline jet jdiv(jet x, jet y)
{   /* ceres/jet.h operator/: g.a_inverse = 1/g.a; f_a_by_g_a = f.a*g_a_inverse; v = (f.v - f_a_by_g_a*g.v)*g_a_inverse */
    jet r; double gi = 1.0 / y.a; double fg = x.a * gi; r.a = fg;
    for (int i = 0; i < NJ; i++) r.v[i] = (x.v[i] - fg * y.v[i]) * gi;
    return r;
}
static inline jet jmuls(jet x, double s) { jet r; r.a = x.a * s; for (int i = 0; i < NJ; i++) r.v[i] = x.v[i] * s; return r; }
static inline jet jadds(jet x, double s) { x.a += s; return x; }
static inline jet jsqrt(jet x) { jet r; double t = sqrt(x.a); r.a = t; double d = 1.0 / (2.0 * t); for (int i = 0; i < NJ; i++) r.v[i] = x.v[i] * d; return r; }
static inline jet jsin(jet x) { jet r; r.a = sin(x.a); double c = cos(x.a); for (int i = 0; i < NJ; i++) r.v[i] = c * x.v[i]; return r; }
static inline jet jcos(jet x) { jet r; r.a = cos(x.a); double s = -sin(x.a); for (int i = 0; i < NJ; i++) r.v[i] = s * x.v[i]; return r; }

/* ceres::AngleAxisRotatePoint<Jet> (ceres/rotation.h) */
static void angle_axis_rotate_point(const jet aa[3], const jet pt[3], jet out[3])
{
    jet theta2 = jadd(jadd(jmul(aa[0], aa[0]), jmul(aa[1], aa[1])), jmul(aa[2], aa[2]));
    if (theta2.a > DBL_EPSILON) {
        jet theta = jsqrt(theta2), ct = jcos(theta), st = jsin(theta);
        jet ti = jdiv(jconst(1.0), theta);
        jet w[3] = {jmul(aa[0], ti), jmul(aa[1], ti), jmul(aa[2], ti)};
        jet wxp[3] = {jsub(jmul(w[1], pt[2]), jmul(w[2], pt[1])), jsub(jmul(w[2], pt[0]), jmul(w[0], pt[2])),
                      jsub(jmul(w[0], pt[1]), jmul(w[1], pt[0]))};
        jet tmp = jmul(jadd(jadd(jmul(w[0], pt[0]), jmul(w[1], pt[1])), jmul(w[2], pt[2])), jsub(jconst(1.0), ct));
        for (int i = 0; i < 3; i++) out[i] = jadd(jadd(jmul(pt[i], ct), jmul(wxp[i], st)), jmul(w[i], tmp));
    } else {
        jet wxp[3] = {jsub(jmul(aa[1], pt[2]), jmul(aa[2], pt[1])), jsub(jmul(aa[2], pt[0]), jmul(aa[0], pt[2])),
                      jsub(jmul(aa[0], pt[1]), jmul(aa[1], pt[0]))};
        for (int i = 0; i < 3; i++) out[i] = jadd(pt[i], wxp[i]);
    }
}

/* Optional evaluation hook (oracle/_ref): when set, every residual / Jacobian evaluation of the minimiser below goes
 * to the callback instead of the restated functor -- oracle/ref_shim's ceres::Solve points it at the REFERENCE's own
 * ceres::CostFunction::Evaluate (ProjectionResidual compiled unchanged under real Jets); `obs` is then an opaque
 * per-observation payload (the residual-block id) that the minimiser only passes through. */
typedef void (*orc_residual_hook_t)(const double pose[6], const double pt[3], const double obs[2], const double K[9],
                                    double r[2], double *Jc, double *Jp);
static orc_residual_hook_t g_residual_hook = 0;
ORC_API void orc_ba_set_residual_hook(orc_residual_hook_t h) { g_residual_hook = h; }

/* ProjectionResidual::operator()<Jet> (ProjectionResidual.h:38-58): r[2], J_pose 2x6, J_pt 2x3 (row major) */
ORC_API void orc_ba_residual(const double pose[6], const double pt[3], const double obs[2], const double K[9],
                             double r[2], double Jc[12], double Jp[6])
{
    if (g_residual_hook) { g_residual_hook(pose, pt, obs, K, r, Jc, Jp); return; }
    jet tr[6], X[3];
    for (int i = 0; i < 6; i++) tr[i] = jvar(pose[i], i);
    for (int i = 0; i < 3; i++) X[i] = jvar(pt[i], 6 + i);
    jet p2[3] = {jadd(X[0], tr[3]), jadd(X[1], tr[4]), jadd(X[2], tr[5])}, p[3];
    angle_axis_rotate_point(tr, p2, p);
    p[2] = jmuls(p[2], -1.0);
    p[0] = jadds(jmuls(jdiv(p[0], p[2]), K[0]), K[2]);
    p[1] = jadds(jmuls(jdiv(p[1], p[2]), K[4]), K[5]);
    jet r0 = jsub(jconst(obs[0]), p[0]), r1 = jsub(jconst(obs[1]), p[1]);
    r[0] = r0.a; r[1] = r1.a;
    if (Jc) for (int k = 0; k < 6; k++) { Jc[k] = r0.v[k]; Jc[6 + k] = r1.v[k]; }
    if (Jp) for (int k = 0; k < 3; k++) { Jp[k] = r0.v[6 + k]; Jp[3 + k] = r1.v[6 + k]; }
}

/* scalar-only residual (cost evaluation of the candidate point) */
static void residual_only(const double pose[6], const double pt[3], const double obs[2], const double K[9], double r[2])
{
    if (g_residual_hook) { g_residual_hook(pose, pt, obs, K, r, 0, 0); return; }
    double q[3] = {pt[0] + pose[3], pt[1] + pose[4], pt[2] + pose[5]}, p[3];
    double t2 = pose[0] * pose[0] + pose[1] * pose[1] + pose[2] * pose[2];
    if (t2 > DBL_EPSILON) {
        double th = sqrt(t2), ct = cos(th), st = sin(th), ti = 1.0 / th;
        double w[3] = {pose[0] * ti, pose[1] * ti, pose[2] * ti};
        double wx[3] = {w[1] * q[2] - w[2] * q[1], w[2] * q[0] - w[0] * q[2], w[0] * q[1] - w[1] * q[0]};
        double tmp = (w[0] * q[0] + w[1] * q[1] + w[2] * q[2]) * (1.0 - ct);
        for (int i = 0; i < 3; i++) p[i] = q[i] * ct + wx[i] * st + w[i] * tmp;
    } else {
        double wx[3] = {pose[1] * q[2] - pose[2] * q[1], pose[2] * q[0] - pose[0] * q[2], pose[0] * q[1] - pose[1] * q[0]};
        for (int i = 0; i < 3; i++) p[i] = q[i] + wx[i];
    }
    p[2] = p[2] * -1.0;
    r[0] = obs[0] - (p[0] / p[2] * K[0] + K[2]);
    r[1] = obs[1] - (p[1] / p[2] * K[4] + K[5]);
}

/* ceres::HuberLoss(a)::Evaluate */
static inline void huber(double a, double s, double rho[3])
{
    double b = a * a;
    if (s > b) {
        double r = sqrt(s);
        rho[0] = 2.0 * a * r - b;
        rho[1] = a / r; if (rho[1] < DBL_MIN) rho[1] = DBL_MIN;
        rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

/* cost = 1/2 sum rho(|r|^2); raw r / J out (what ProjectionResidual + AutoDiff return); huber_delta<=0: no loss */
ORC_API double orc_ba_eval(const double *poses, const double *points, const double *obs, const int *cam_idx,
                           const int *pt_idx, int No, const double *K, double huber_delta,
                           double *r, double *Jc, double *Jp)
{
    double cost = 0;
    for (int i = 0; i < No; i++) {
        double ri[2], jc[12], jp[6];
        orc_ba_residual(poses + 6 * cam_idx[i], points + 3 * pt_idx[i], obs + 2 * i, K, ri, jc, jp);
        double s = ri[0] * ri[0] + ri[1] * ri[1];
        if (huber_delta > 0) { double rho[3]; huber(huber_delta, s, rho); cost += 0.5 * rho[0]; }
        else cost += 0.5 * s;
        if (r) { r[2 * i] = ri[0]; r[2 * i + 1] = ri[1]; }
        if (Jc) memcpy(Jc + 12 * i, jc, sizeof jc);
        if (Jp) memcpy(Jp + 6 * i, jp, sizeof jp);
    }
    return cost;
}

/* ------------------------------------------------------------------ LM + Schur ------------ */
typedef struct {
    double initial_cost, final_cost;
    int iterations;            /* LM iterations executed (accepted + rejected + invalid) */
    int successful_steps;
    int termination;           /* 0 max iterations, 1 function tol, 2 parameter tol, 3 gradient tol, 4 failure, 5 min radius */
    double final_radius;
    double cost_log[128];      /* cost of the iterate after each iteration (index 0 = initial) */
    double radius_log[128];
    int accepted_log[128];
} orc_ba_summary;

typedef struct {
    int Nc, Np, No;
    const int *cam, *pt;       /* per observation, sorted by point */
    const double *obs;
    int *pt_off;               /* CSR over points */
    double K[9], delta;
} ba_prob;

typedef struct { double r[2], jc[12], jp[6]; } obs_lin;  /* corrected + column-scaled */

static double eval_cost(const ba_prob *P, const double *x_c, const double *x_p)
{
    double cost = 0;
    #pragma omp parallel for reduction(+:cost) schedule(static)
    for (int i = 0; i < P->No; i++) {
        double r[2];
        residual_only(x_c + 6 * P->cam[i], x_p + 3 * P->pt[i], P->obs + 2 * i, P->K, r);
        double s = r[0] * r[0] + r[1] * r[1];
        if (P->delta > 0) { double rho[3]; huber(P->delta, s, rho); cost += 0.5 * rho[0]; } else cost += 0.5 * s;
    }
    return cost;
}

/* residuals + Jacobians at x, Corrector applied (rho'' <= 0 branch for Huber: pure sqrt(rho') scaling) */
static double linearize(const ba_prob *P, const double *x_c, const double *x_p, obs_lin *L)
{
    double cost = 0;
    #pragma omp parallel for reduction(+:cost) schedule(static)
    for (int i = 0; i < P->No; i++) {
        obs_lin *l = &L[i];
        orc_ba_residual(x_c + 6 * P->cam[i], x_p + 3 * P->pt[i], P->obs + 2 * i, P->K, l->r, l->jc, l->jp);
        double s = l->r[0] * l->r[0] + l->r[1] * l->r[1];
        if (P->delta > 0) {
            double rho[3]; huber(P->delta, s, rho); cost += 0.5 * rho[0];
            double sr = sqrt(rho[1]);
            /* Corrector: sq_norm == 0 || rho[2] <= 0  ->  residual_scaling = sqrt(rho'), alpha = 0 */
            for (int k = 0; k < 12; k++) l->jc[k] *= sr;
            for (int k = 0; k < 6; k++) l->jp[k] *= sr;
            l->r[0] *= sr; l->r[1] *= sr;
        } else cost += 0.5 * s;
    }
    return cost;
}

static int chol3_inverse(const double V[9], double inv[9])
{   /* 3x3 SPD inverse through LL^T (Eigen's llt().solve(I) in InvertPSDMatrix) */
    double l00 = V[0]; if (!(l00 > 0)) return 0; l00 = sqrt(l00);
    double l10 = V[3] / l00, l20 = V[6] / l00;
    double l11 = V[4] - l10 * l10; if (!(l11 > 0)) return 0; l11 = sqrt(l11);
    double l21 = (V[7] - l20 * l10) / l11;
    double l22 = V[8] - l20 * l20 - l21 * l21; if (!(l22 > 0)) return 0; l22 = sqrt(l22);
    for (int c = 0; c < 3; c++) {
        double b[3] = {c == 0, c == 1, c == 2};
        double y0 = b[0] / l00, y1 = (b[1] - l10 * y0) / l11, y2 = (b[2] - l20 * y0 - l21 * y1) / l22;
        double x2 = y2 / l22, x1 = (y1 - l21 * x2) / l11, x0 = (y0 - l10 * x1 - l20 * x2) / l00;
        inv[c] = x0; inv[3 + c] = x1; inv[6 + c] = x2;
    }
    return 1;
}

/* envelope (profile) Cholesky of the reduced camera matrix: exploits the band a sparse Cholesky would */
static int envelope_cholesky_solve(double *S, int n, double *b)
{
    int *first = (int *)malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++) { int j = 0; while (j < i && S[(size_t)i * n + j] == 0.0) j++; first[i] = j; }
    int ok = 1;
    for (int i = 0; i < n && ok; i++) {
        for (int j = first[i]; j <= i; j++) {
            int k0 = first[i] > first[j] ? first[i] : first[j];
            double s = S[(size_t)i * n + j];
            const double *li = S + (size_t)i * n, *lj = S + (size_t)j * n;
            for (int k = k0; k < j; k++) s -= li[k] * lj[k];
            if (j < i) S[(size_t)i * n + j] = s / lj[j];
            else { if (!(s > 0) || !isfinite(s)) { ok = 0; break; } S[(size_t)i * n + i] = sqrt(s); }
        }
    }
    if (ok) {
        for (int i = 0; i < n; i++) { double s = b[i]; for (int k = first[i]; k < i; k++) s -= S[(size_t)i * n + k] * b[k]; b[i] = s / S[(size_t)i * n + i]; }
        for (int i = n - 1; i >= 0; i--) {
            b[i] /= S[(size_t)i * n + i];
            double bi = b[i];
            for (int k = first[i]; k < i; k++) b[k] -= S[(size_t)i * n + k] * bi;
        }
    }
    free(first);
    return ok;
}

/* Solve (J^T J + D^2) y = J^T r by eliminating the points (SchurEliminator + BackSubstitute). */
static int schur_solve(const ba_prob *P, const obs_lin *L, const double *D2c, const double *D2p,
                       double *yc, double *yp)
{
    const int n = 6 * P->Nc;
    double *S = (double *)calloc((size_t)n * n, sizeof(double));
    double *rhs = (double *)calloc(n, sizeof(double));
    double *Vinv = (double *)malloc(sizeof(double) * 9 * P->Np), *gp = (double *)malloc(sizeof(double) * 3 * P->Np);
    int ok = 1;
    /* diagonal blocks U_c + D_c^2 and J_c^T r */
    for (int i = 0; i < P->No; i++) {
        const obs_lin *l = &L[i]; int c = P->cam[i];
        for (int a = 0; a < 6; a++) {
            for (int b = 0; b < 6; b++) S[(size_t)(6 * c + a) * n + 6 * c + b] += l->jc[a] * l->jc[b] + l->jc[6 + a] * l->jc[6 + b];
            rhs[6 * c + a] += l->jc[a] * l->r[0] + l->jc[6 + a] * l->r[1];
        }
    }
    for (int j = 0; j < n; j++) S[(size_t)j * n + j] += D2c[j];
    for (int p = 0; p < P->Np; p++) {
        int o0 = P->pt_off[p], o1 = P->pt_off[p + 1];
        double V[9] = {0}, g[3] = {0};
        for (int i = o0; i < o1; i++) {
            const obs_lin *l = &L[i];
            for (int a = 0; a < 3; a++) {
                for (int b = 0; b < 3; b++) V[3 * a + b] += l->jp[a] * l->jp[b] + l->jp[3 + a] * l->jp[3 + b];
                g[a] += l->jp[a] * l->r[0] + l->jp[3 + a] * l->r[1];
            }
        }
        for (int a = 0; a < 3; a++) V[4 * a] += D2p[3 * p + a];
        double *Vi = Vinv + 9 * p;
        if (!chol3_inverse(V, Vi)) { ok = 0; break; }
        memcpy(gp + 3 * p, g, sizeof g);
        for (int i = o0; i < o1; i++) {
            const obs_lin *li = &L[i]; int ci = P->cam[i];
            double W[18], Y[18]; /* W = Jc^T Jp (6x3), Y = W V^-1 */
            for (int a = 0; a < 6; a++) for (int b = 0; b < 3; b++) W[3 * a + b] = li->jc[a] * li->jp[b] + li->jc[6 + a] * li->jp[3 + b];
            for (int a = 0; a < 6; a++) for (int b = 0; b < 3; b++) Y[3 * a + b] = W[3 * a] * Vi[b] + W[3 * a + 1] * Vi[3 + b] + W[3 * a + 2] * Vi[6 + b];
            for (int a = 0; a < 6; a++) rhs[6 * ci + a] -= Y[3 * a] * g[0] + Y[3 * a + 1] * g[1] + Y[3 * a + 2] * g[2];
            for (int k = o0; k < o1; k++) {
                const obs_lin *lk = &L[k]; int ck = P->cam[k];
                double Wk[18];
                for (int a = 0; a < 6; a++) for (int b = 0; b < 3; b++) Wk[3 * a + b] = lk->jc[a] * lk->jp[b] + lk->jc[6 + a] * lk->jp[3 + b];
                for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++)
                    S[(size_t)(6 * ci + a) * n + 6 * ck + b] -= Y[3 * a] * Wk[3 * b] + Y[3 * a + 1] * Wk[3 * b + 1] + Y[3 * a + 2] * Wk[3 * b + 2];
            }
        }
    }
    if (ok) {
        memcpy(yc, rhs, sizeof(double) * n);
        ok = envelope_cholesky_solve(S, n, yc);
    }
    if (ok) {
        for (int p = 0; p < P->Np; p++) {
            double t[3] = {gp[3 * p], gp[3 * p + 1], gp[3 * p + 2]};
            for (int i = P->pt_off[p]; i < P->pt_off[p + 1]; i++) {
                const obs_lin *l = &L[i]; const double *y = yc + 6 * P->cam[i];
                double jy0 = 0, jy1 = 0;
                for (int a = 0; a < 6; a++) { jy0 += l->jc[a] * y[a]; jy1 += l->jc[6 + a] * y[a]; }
                for (int b = 0; b < 3; b++) t[b] -= l->jp[b] * jy0 + l->jp[3 + b] * jy1; /* W^T y_c */
            }
            const double *Vi = Vinv + 9 * p;
            for (int a = 0; a < 3; a++) yp[3 * p + a] = Vi[3 * a] * t[0] + Vi[3 * a + 1] * t[1] + Vi[3 * a + 2] * t[2];
        }
    }
    free(S); free(rhs); free(Vinv); free(gp);
    return ok;
}

/* dense normal equations: independent linear-algebra path used to cross-check schur_solve */
static int dense_solve(const ba_prob *P, const obs_lin *L, const double *D2c, const double *D2p, double *yc, double *yp)
{
    const int nc = 6 * P->Nc, n = nc + 3 * P->Np;
    double *A = (double *)calloc((size_t)n * n, sizeof(double)), *b = (double *)calloc(n, sizeof(double));
    for (int i = 0; i < P->No; i++) {
        const obs_lin *l = &L[i];
        int idx[9];
        for (int a = 0; a < 6; a++) idx[a] = 6 * P->cam[i] + a;
        for (int a = 0; a < 3; a++) idx[6 + a] = nc + 3 * P->pt[i] + a;
        for (int row = 0; row < 2; row++) {
            double j[9];
            for (int a = 0; a < 6; a++) j[a] = l->jc[6 * row + a];
            for (int a = 0; a < 3; a++) j[6 + a] = l->jp[3 * row + a];
            for (int a = 0; a < 9; a++) {
                b[idx[a]] += j[a] * l->r[row];
                for (int c = 0; c < 9; c++) A[(size_t)idx[a] * n + idx[c]] += j[a] * j[c];
            }
        }
    }
    for (int j = 0; j < nc; j++) A[(size_t)j * n + j] += D2c[j];
    for (int j = 0; j < 3 * P->Np; j++) A[(size_t)(nc + j) * n + nc + j] += D2p[j];
    int ok = envelope_cholesky_solve(A, n, b);
    if (ok) { memcpy(yc, b, sizeof(double) * nc); memcpy(yp, b + nc, sizeof(double) * 3 * P->Np); }
    free(A); free(b);
    return ok;
}

static int all_finite(const double *v, int n) { for (int i = 0; i < n; i++) if (!isfinite(v[i])) return 0; return 1; }

/* Ceres TrustRegionMinimizer + LevenbergMarquardtStrategy, defaults of Ceres >= 1.13 (SURVEY Appx C.3).
 * cam_idx/pt_idx need not be sorted.  use_dense != 0 solves the normal equations densely. */
ORC_API int orc_ba_solve(double *poses, double *points, const double *obs_in, const int *cam_idx, const int *pt_idx,
                         int Nc, int Np, int No, const double *K, double huber_delta, int max_iters,
                         int use_dense, orc_ba_summary *out)
{
    orc_ba_summary sum; memset(&sum, 0, sizeof sum);
    /* ---- order observations by point (stable counting sort): the e-blocks Ceres eliminates */
    int *off = (int *)calloc(Np + 1, sizeof(int));
    for (int i = 0; i < No; i++) off[pt_idx[i] + 1]++;
    for (int p = 0; p < Np; p++) off[p + 1] += off[p];
    int *pos = (int *)malloc(sizeof(int) * (Np + 1)); memcpy(pos, off, sizeof(int) * (Np + 1));
    int *cam = (int *)malloc(sizeof(int) * (No ? No : 1)), *pt = (int *)malloc(sizeof(int) * (No ? No : 1));
    double *obs = (double *)malloc(sizeof(double) * 2 * (No ? No : 1));
    for (int i = 0; i < No; i++) { int d = pos[pt_idx[i]]++; cam[d] = cam_idx[i]; pt[d] = pt_idx[i]; obs[2 * d] = obs_in[2 * i]; obs[2 * d + 1] = obs_in[2 * i + 1]; }
    free(pos);
    ba_prob P = {Nc, Np, No, cam, pt, obs, off, {0}, huber_delta};
    memcpy(P.K, K, sizeof P.K);
    /* parameter blocks exist only if they have a residual (Ceres adds blocks with AddResidualBlock) */
    char *act_c = (char *)calloc(Nc ? Nc : 1, 1), *act_p = (char *)calloc(Np ? Np : 1, 1);
    for (int i = 0; i < No; i++) { act_c[cam[i]] = 1; act_p[pt[i]] = 1; }

    const int nc = 6 * Nc, np = 3 * Np;
    obs_lin *L = (obs_lin *)malloc(sizeof(obs_lin) * (No ? No : 1));
    double *sc_c = (double *)calloc(nc ? nc : 1, sizeof(double)), *sc_p = (double *)calloc(np ? np : 1, sizeof(double));
    double *dg_c = (double *)calloc(nc ? nc : 1, sizeof(double)), *dg_p = (double *)calloc(np ? np : 1, sizeof(double));
    double *D2c = (double *)malloc(sizeof(double) * (nc ? nc : 1)), *D2p = (double *)malloc(sizeof(double) * (np ? np : 1));
    double *yc = (double *)malloc(sizeof(double) * (nc ? nc : 1)), *yp = (double *)malloc(sizeof(double) * (np ? np : 1));
    double *cand_c = (double *)malloc(sizeof(double) * (nc ? nc : 1)), *cand_p = (double *)malloc(sizeof(double) * (np ? np : 1));

    double radius = 1e4, decrease_factor = 2.0;
    const double max_radius = 1e16, min_radius = 1e-32, min_rel_dec = 1e-3, min_diag = 1e-6, max_diag = 1e32;
    const double f_tol = 1e-6, g_tol = 1e-10, p_tol = 1e-8;
    int reuse_diagonal = 0, invalid_run = 0;

    /* ---- iteration 0 */
    double cost = linearize(&P, poses, points, L);
    sum.initial_cost = cost; sum.cost_log[0] = cost; sum.radius_log[0] = radius; sum.accepted_log[0] = 1;
    /* gradient (unscaled J) max norm */
    #define GRADIENT_MAX(gm) do { \
        double *gc_ = (double *)calloc(nc + np + 1, sizeof(double)); \
        for (int i_ = 0; i_ < No; i_++) { const obs_lin *l_ = &L[i_]; \
            for (int a_ = 0; a_ < 6; a_++) gc_[6 * cam[i_] + a_] += l_->jc[a_] * l_->r[0] + l_->jc[6 + a_] * l_->r[1]; \
            for (int a_ = 0; a_ < 3; a_++) gc_[nc + 3 * pt[i_] + a_] += l_->jp[a_] * l_->r[0] + l_->jp[3 + a_] * l_->r[1]; } \
        gm = 0; for (int j_ = 0; j_ < nc + np; j_++) if (fabs(gc_[j_]) > gm) gm = fabs(gc_[j_]); free(gc_); } while (0)
    double gmax; GRADIENT_MAX(gmax);
    /* Jacobi scaling from the initial (corrected) Jacobian, kept for the whole solve */
    for (int i = 0; i < No; i++) {
        for (int a = 0; a < 6; a++) sc_c[6 * cam[i] + a] += L[i].jc[a] * L[i].jc[a] + L[i].jc[6 + a] * L[i].jc[6 + a];
        for (int a = 0; a < 3; a++) sc_p[3 * pt[i] + a] += L[i].jp[a] * L[i].jp[a] + L[i].jp[3 + a] * L[i].jp[3 + a];
    }
    for (int j = 0; j < nc; j++) sc_c[j] = 1.0 / (1.0 + sqrt(sc_c[j]));
    for (int j = 0; j < np; j++) sc_p[j] = 1.0 / (1.0 + sqrt(sc_p[j]));
    #define SCALE_COLUMNS() do { \
        for (int i_ = 0; i_ < No; i_++) { obs_lin *l_ = &L[i_]; \
            for (int a_ = 0; a_ < 6; a_++) { l_->jc[a_] *= sc_c[6 * cam[i_] + a_]; l_->jc[6 + a_] *= sc_c[6 * cam[i_] + a_]; } \
            for (int a_ = 0; a_ < 3; a_++) { l_->jp[a_] *= sc_p[3 * pt[i_] + a_]; l_->jp[3 + a_] *= sc_p[3 * pt[i_] + a_]; } } } while (0)
    SCALE_COLUMNS();
    double x_norm = 0;
    #define X_NORM(xc_, xp_, res) do { double s_ = 0; \
        for (int c_ = 0; c_ < Nc; c_++) if (act_c[c_]) for (int a_ = 0; a_ < 6; a_++) s_ += (xc_)[6 * c_ + a_] * (xc_)[6 * c_ + a_]; \
        for (int p_ = 0; p_ < Np; p_++) if (act_p[p_]) for (int a_ = 0; a_ < 3; a_++) s_ += (xp_)[3 * p_ + a_] * (xp_)[3 * p_ + a_]; \
        res = sqrt(s_); } while (0)
    X_NORM(poses, points, x_norm);

    int iter = 0;
    sum.termination = 0;
    if (No == 0) { sum.final_cost = cost; goto done; }
    if (gmax <= g_tol) { sum.termination = 3; goto done; }
    while (iter < max_iters) {
        iter++;
        int idx = iter < 128 ? iter : 127;
        /* ---- LevenbergMarquardtStrategy::ComputeStep */
        if (!reuse_diagonal) {
            memset(dg_c, 0, sizeof(double) * nc); memset(dg_p, 0, sizeof(double) * np);
            for (int i = 0; i < No; i++) {
                for (int a = 0; a < 6; a++) dg_c[6 * cam[i] + a] += L[i].jc[a] * L[i].jc[a] + L[i].jc[6 + a] * L[i].jc[6 + a];
                for (int a = 0; a < 3; a++) dg_p[3 * pt[i] + a] += L[i].jp[a] * L[i].jp[a] + L[i].jp[3 + a] * L[i].jp[3 + a];
            }
            for (int j = 0; j < nc; j++) dg_c[j] = fmin(fmax(dg_c[j], min_diag), max_diag);
            for (int j = 0; j < np; j++) dg_p[j] = fmin(fmax(dg_p[j], min_diag), max_diag);
        }
        for (int j = 0; j < nc; j++) { double d = sqrt(dg_c[j] / radius); D2c[j] = d * d; }
        for (int j = 0; j < np; j++) { double d = sqrt(dg_p[j] / radius); D2p[j] = d * d; }
        int ok = use_dense ? dense_solve(&P, L, D2c, D2p, yc, yp) : schur_solve(&P, L, D2c, D2p, yc, yp);
        if (ok) ok = all_finite(yc, nc) && all_finite(yp, np);
        reuse_diagonal = 1;
        int step_valid = 0;
        double model_cost_change = 0;
        if (ok) {
            /* step = -y ; model_cost_change = -(J step)^T (r + J step / 2) */
            for (int i = 0; i < No; i++) {
                const obs_lin *l = &L[i]; const double *a = yc + 6 * cam[i], *b = yp + 3 * pt[i];
                for (int row = 0; row < 2; row++) {
                    double m = 0;
                    for (int k = 0; k < 6; k++) m -= l->jc[6 * row + k] * a[k];
                    for (int k = 0; k < 3; k++) m -= l->jp[3 * row + k] * b[k];
                    model_cost_change -= m * (l->r[row] + m / 2.0);
                }
            }
            step_valid = model_cost_change > 0.0;
        }
        if (!step_valid) {
            /* HandleInvalidStep */
            if (++invalid_run >= 5) { sum.termination = 4; sum.cost_log[idx] = cost; sum.radius_log[idx] = radius; break; }
            radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;
            sum.cost_log[idx] = cost; sum.radius_log[idx] = radius; sum.accepted_log[idx] = 0;
            if (radius < min_radius) { sum.termination = 5; break; }
            continue;
        }
        invalid_run = 0;
        /* delta = step .* scale ; candidate = x + delta */
        double step_norm2 = 0;
        for (int c = 0; c < Nc; c++) for (int a = 0; a < 6; a++) {
            double d = -yc[6 * c + a] * sc_c[6 * c + a];
            cand_c[6 * c + a] = poses[6 * c + a] + (act_c[c] ? d : 0.0);
            if (act_c[c]) { double dd = poses[6 * c + a] - cand_c[6 * c + a]; step_norm2 += dd * dd; }
        }
        for (int p = 0; p < Np; p++) for (int a = 0; a < 3; a++) {
            double d = -yp[3 * p + a] * sc_p[3 * p + a];
            cand_p[3 * p + a] = points[3 * p + a] + (act_p[p] ? d : 0.0);
            if (act_p[p]) { double dd = points[3 * p + a] - cand_p[3 * p + a]; step_norm2 += dd * dd; }
        }
        double cand_cost = eval_cost(&P, cand_c, cand_p);
        if (!isfinite(cand_cost)) cand_cost = DBL_MAX;
        /* ParameterToleranceReached */
        if (sqrt(step_norm2) <= p_tol * (x_norm + p_tol)) { sum.termination = 2; sum.cost_log[idx] = cost; sum.radius_log[idx] = radius; break; }
        /* FunctionToleranceReached */
        if (fabs(cost - cand_cost) <= f_tol * cost) { sum.termination = 1; sum.cost_log[idx] = cost; sum.radius_log[idx] = radius; break; }
        double rel = (cost - cand_cost) / model_cost_change;
        if (rel > min_rel_dec) {
            memcpy(poses, cand_c, sizeof(double) * nc); memcpy(points, cand_p, sizeof(double) * np);
            X_NORM(poses, points, x_norm);
            cost = linearize(&P, poses, points, L);
            GRADIENT_MAX(gmax);
            SCALE_COLUMNS();
            radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rel - 1.0, 3));
            radius = fmin(max_radius, radius);
            decrease_factor = 2.0; reuse_diagonal = 0;
            sum.successful_steps++;
            sum.cost_log[idx] = cost; sum.radius_log[idx] = radius; sum.accepted_log[idx] = 1;
            if (gmax <= g_tol) { sum.termination = 3; break; }
        } else {
            radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;
            sum.cost_log[idx] = cost; sum.radius_log[idx] = radius; sum.accepted_log[idx] = 0;
        }
        if (radius < min_radius) { sum.termination = 5; break; }
    }
done:
    sum.iterations = iter; sum.final_cost = cost; sum.final_radius = radius;
    if (out) *out = sum;
    free(off); free(cam); free(pt); free(obs); free(act_c); free(act_p); free(L);
    free(sc_c); free(sc_p); free(dg_c); free(dg_p); free(D2c); free(D2p); free(yc); free(yp); free(cand_c); free(cand_p);
    return sum.termination;
}

/* Thread count of the OpenMP regions (bench.py times the port with the reference's 4 threads,
 * CeresBundleAdjustment.cpp:58, and with every host core). */
ORC_API void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* W independent windows (BASELINE config 4): window w uses poses[w], points[w], and the observation
 * slice [obs_off[w], obs_off[w+1]).  OpenMP over windows. */
ORC_API void orc_ba_solve_batched(double *poses, double *points, const double *obs, const int *cam_idx,
                                  const int *pt_idx, const int *obs_off, int W, int Nc, int Np,
                                  const double *K, double huber_delta, int max_iters, int nthreads,
                                  orc_ba_summary *out)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for schedule(dynamic, 1)
    for (int w = 0; w < W; w++) {
        int o0 = obs_off[w], n = obs_off[w + 1] - o0;
        orc_ba_solve(poses + (size_t)w * 6 * Nc, points + (size_t)w * 3 * Np, obs + 2 * (size_t)o0, cam_idx + o0,
                     pt_idx + o0, Nc, Np, n, K, huber_delta, max_iters, 0, out ? &out[w] : NULL);
    }
}

// ba_kernels.cuh -- device code of the bundle adjuster: K8 residual/Jacobian, K9 Schur assembly,
// K10 small reduced-system Cholesky, K11 back-substitution + candidate cost, LM bookkeeping.
//
// Replaces what ceres::Solve does for CeresBundleAdjustment::apply (reference
// CeresBundleAdjustment.cpp:54-61) on residual blocks built from ProjectionResidual
// (include/ProjectionResidual.h:38-58, ProjectionResidual.cpp:3-8): AutoDiff Jacobians, Huber
// Corrector, Jacobi scaling, SchurEliminator, Cholesky, TrustRegionMinimizer (SURVEY Appx C).
// fp64 throughout (final cost must agree to 1e-6 relative).
#pragma once
#include <float.h>
#include <math.h>

#include "ba.cuh"

namespace {

// ---- ProjectionResidual: residual + analytic 2x6 / 2x3 Jacobians (== what Jets produce) ------------
// The angle-axis terms of a pose that do not depend on the point: cos / sin of the angle, its reciprocal, the unit axis.
// ba_project() derives them in place; the kernels of one LARGE problem read them from a per-camera table that
// ba_cam_trig_kernel fills once per iteration (5 observations per point and ~5 000 per camera: the square root, the
// sine / cosine pair and the division were ~40 % of the instructions of every kernel that recomputes a Jacobian).
// Same expressions either way, so both routes give the same bits.
struct CamTrig {
    double ct, st, ti, w0, w1, w2;   // small: w0..w2 hold the raw angle-axis vector (first-order rotation)
    bool small;
};

__device__ __forceinline__ CamTrig ba_cam_trig(const double *pose)
{
    CamTrig T;
    const double a0 = pose[0], a1 = pose[1], a2 = pose[2];
    const double th2 = a0 * a0 + a1 * a1 + a2 * a2;
    T.small = !(th2 > DBL_EPSILON);
    if (!T.small) {
        const double th = sqrt(th2);
        T.ct = cos(th); T.st = sin(th); T.ti = 1.0 / th;
        T.w0 = a0 * T.ti; T.w1 = a1 * T.ti; T.w2 = a2 * T.ti;
    } else {
        T.ct = 1.0; T.st = 0.0; T.ti = 0.0; T.w0 = a0; T.w1 = a1; T.w2 = a2;
    }
    return T;
}

// table layout: 8 doubles per camera {ct, st, ti, w0, w1, w2, small ? 1 : 0, 0}
__device__ __forceinline__ CamTrig ba_cam_trig_load(const double *__restrict__ t)
{
    const double2 a = *reinterpret_cast<const double2 *>(t), b = *reinterpret_cast<const double2 *>(t + 2),
                  c = *reinterpret_cast<const double2 *>(t + 4);
    CamTrig T;
    T.ct = a.x; T.st = a.y; T.ti = b.x; T.w0 = b.y; T.w1 = c.x; T.w2 = c.y;
    T.small = t[6] != 0.0;
    return T;
}

__device__ __forceinline__ void ba_project_t(const CamTrig &T, const double *tr /* pose + 3 */, const double *X, double p[3], double R[9],
                                             double dpa[9] /* dpa[3*k+i] = d p_i / d a_k */, bool want_jac)
{
    const double q0 = X[0] + tr[0], q1 = X[1] + tr[1], q2 = X[2] + tr[2];
    if (!T.small) {
        const double ct = T.ct, st = T.st, ti = T.ti;
        const double w0 = T.w0, w1 = T.w1, w2 = T.w2;
        const double x0 = w1 * q2 - w2 * q1, x1 = w2 * q0 - w0 * q2, x2 = w0 * q1 - w1 * q0;  // w x q
        const double d = w0 * q0 + w1 * q1 + w2 * q2, omc = 1.0 - ct;
        const double tmp = d * omc;
        p[0] = q0 * ct + x0 * st + w0 * tmp;
        p[1] = q1 * ct + x1 * st + w1 * tmp;
        p[2] = q2 * ct + x2 * st + w2 * tmp;
        if (want_jac) {
            // R = ct I + st [w]x + omc w w^T  (d p / d X = d p / d c)
            R[0] = ct + omc * w0 * w0; R[1] = -st * w2 + omc * w0 * w1; R[2] = st * w1 + omc * w0 * w2;
            R[3] = st * w2 + omc * w1 * w0; R[4] = ct + omc * w1 * w1; R[5] = -st * w0 + omc * w1 * w2;
            R[6] = -st * w1 + omc * w2 * w0; R[7] = st * w0 + omc * w2 * w1; R[8] = ct + omc * w2 * w2;
            const double w[3] = {w0, w1, w2}, q[3] = {q0, q1, q2}, x[3] = {x0, x1, x2};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                // d theta / d a_k = w_k ;  d w / d a_k = (e_k - w w_k) / theta
                double dw[3] = {-w0 * w[k] * ti, -w1 * w[k] * ti, -w2 * w[k] * ti};
                dw[k] += ti;
                const double dx0 = dw[1] * q2 - dw[2] * q1, dx1 = dw[2] * q0 - dw[0] * q2, dx2 = dw[0] * q1 - dw[1] * q0;
                const double dd = dw[0] * q0 + dw[1] * q1 + dw[2] * q2;
                const double dxs[3] = {dx0, dx1, dx2};
#pragma unroll
                for (int i = 0; i < 3; i++)
                    dpa[3 * k + i] = -q[i] * st * w[k] + dxs[i] * st + x[i] * ct * w[k] + dw[i] * tmp +
                                     w[i] * (dd * omc + d * st * w[k]);
            }
        }
    } else {
        const double a0 = T.w0, a1 = T.w1, a2 = T.w2;
        p[0] = q0 + (a1 * q2 - a2 * q1);
        p[1] = q1 + (a2 * q0 - a0 * q2);
        p[2] = q2 + (a0 * q1 - a1 * q0);
        if (want_jac) {
            R[0] = 1; R[1] = -a2; R[2] = a1; R[3] = a2; R[4] = 1; R[5] = -a0; R[6] = -a1; R[7] = a0; R[8] = 1;
            // d(a x q)/d a_k = e_k x q
            dpa[0] = 0; dpa[1] = -q2; dpa[2] = q1;
            dpa[3] = q2; dpa[4] = 0; dpa[5] = -q0;
            dpa[6] = -q1; dpa[7] = q0; dpa[8] = 0;
        }
    }
}

__device__ __forceinline__ void ba_project(const double *pose, const double *X, double p[3], double R[9],
                                           double dpa[9] /* dpa[3*k+i] = d p_i / d a_k */, bool want_jac)
{
    ba_project_t(ba_cam_trig(pose), pose + 3, X, p, R, dpa, want_jac);
}

// ---- fp64 reciprocal / reciprocal square root for the kernels of one large problem -------------------------------------
// The hardware seed (MUFU.RCP64H / RSQ64H on the high word, measured relative error 1e-6 = 2^-20) refined by one cubic
// step: relative error 1 - 2 ulp for normal arguments (tools/approx_probe.cu measures seed and result), without the special-case
// branches and the correctly-rounded last step of x / y and sqrt() -- a third of the instructions and of the dependent
// latency.  Used where the oracle's value is reproduced to ~1 ulp anyway (depth division, V^-1).
__device__ __forceinline__ double ba_rcp_fast(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);     // error e -> e^3: the 2^-20 seed becomes 2^-60, below the rounding of the last fma
}
__device__ __forceinline__ double ba_rsqrt_fast(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);           // 1 - x y^2
    return fma(y, e * fma(0.375, e, 0.5), y);       // y (1 + e / 2 + 3 e^2 / 8): error -> O(e^3)
}

__device__ __forceinline__ void ba_residual_jac(const double *pose, const double *X, double ox, double oy,
                                                double fx, double cx, double fy, double cy,
                                                double r[2], double Jc[12], double Jp[6])
{
    double p[3], R[9], dpa[9];
    ba_project(pose, X, p, R, dpa, true);
    const double pz = p[2] * -1.0;
    r[0] = ox - (p[0] / pz * fx + cx);
    r[1] = oy - (p[1] / pz * fy + cy);
    const double iz = 1.0 / p[2];
    const double a0 = fx * iz, a2 = -fx * p[0] * iz * iz;   // d r0 / d p0, d r0 / d p2
    const double b1 = fy * iz, b2 = -fy * p[1] * iz * iz;   // d r1 / d p1, d r1 / d p2
#pragma unroll
    for (int k = 0; k < 3; k++) {
        Jc[k] = a0 * dpa[3 * k] + a2 * dpa[3 * k + 2];
        Jc[6 + k] = b1 * dpa[3 * k + 1] + b2 * dpa[3 * k + 2];
        const double j0 = a0 * R[k] + a2 * R[6 + k], j1 = b1 * R[3 + k] + b2 * R[6 + k];
        Jc[3 + k] = j0; Jc[9 + k] = j1;
        Jp[k] = j0; Jp[3 + k] = j1;
    }
}

// the same residual / Jacobians from the per-camera table, one refined reciprocal instead of three divisions
__device__ __forceinline__ void ba_residual_jac_t(const CamTrig &T, const double *tr, const double *X, double ox, double oy,
                                                  double fx, double cx, double fy, double cy,
                                                  double r[2], double Jc[12], double Jp[6])
{
    double p[3], R[9], dpa[9];
    ba_project_t(T, tr, X, p, R, dpa, true);
    const double iz = ba_rcp_fast(p[2]);
    r[0] = ox - (-(p[0] * iz) * fx + cx);
    r[1] = oy - (-(p[1] * iz) * fy + cy);
    const double a0 = fx * iz, a2 = -fx * p[0] * iz * iz;   // d r0 / d p0, d r0 / d p2
    const double b1 = fy * iz, b2 = -fy * p[1] * iz * iz;   // d r1 / d p1, d r1 / d p2
#pragma unroll
    for (int k = 0; k < 3; k++) {
        Jc[k] = a0 * dpa[3 * k] + a2 * dpa[3 * k + 2];
        Jc[6 + k] = b1 * dpa[3 * k + 1] + b2 * dpa[3 * k + 2];
        const double j0 = a0 * R[k] + a2 * R[6 + k], j1 = b1 * R[3 + k] + b2 * R[6 + k];
        Jc[3 + k] = j0; Jc[9 + k] = j1;
        Jp[k] = j0; Jp[3 + k] = j1;
    }
}

__device__ __forceinline__ void ba_residual_only_t(const CamTrig &T, const double *tr, const double *X, double ox, double oy,
                                                   double fx, double cx, double fy, double cy, double r[2])
{
    double p[3];
    ba_project_t(T, tr, X, p, nullptr, nullptr, false);
    const double iz = ba_rcp_fast(p[2]);
    r[0] = ox - (-(p[0] * iz) * fx + cx);
    r[1] = oy - (-(p[1] * iz) * fy + cy);
}

// Corrector scale sqrt(rho') -- 1 for every inlier (rho' == 1 exactly): the square root only runs for outliers
__device__ __forceinline__ double ba_sqrt_rho1(double rho1) { return rho1 == 1.0 ? 1.0 : sqrt(rho1); }

// per-camera table of the kernels of one large problem (see CamTrig)
__global__ void __launch_bounds__(128) ba_cam_trig_kernel(const double *__restrict__ poses, double *__restrict__ trig, int ncam,
                                                          const BAState *st, int need_chol)
{
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= ncam || st->done || (need_chol && !st->chol_ok)) return;
    const CamTrig T = ba_cam_trig(poses + 6 * (size_t)c);
    double *t = trig + 8 * (size_t)c;
    t[0] = T.ct; t[1] = T.st; t[2] = T.ti; t[3] = T.w0; t[4] = T.w1; t[5] = T.w2; t[6] = T.small ? 1.0 : 0.0; t[7] = 0.0;
}

__device__ __forceinline__ void ba_residual_only(const double *pose, const double *X, double ox, double oy,
                                                 double fx, double cx, double fy, double cy, double r[2])
{
    double p[3];
    ba_project(pose, X, p, nullptr, nullptr, false);
    const double pz = p[2] * -1.0;
    r[0] = ox - (p[0] / pz * fx + cx);
    r[1] = oy - (p[1] / pz * fy + cy);
}

// ceres::HuberLoss(a): rho(s), rho'(s)
__device__ __forceinline__ void ba_huber(double a, double s, double &rho0, double &rho1)
{
    const double b = a * a;
    if (a > 0 && s > b) {
        const double r = sqrt(s);
        rho0 = 2.0 * a * r - b;
        rho1 = fmax(a / r, DBL_MIN);
    } else { rho0 = s; rho1 = 1.0; }
}

__device__ __forceinline__ double warp_sum_d(double v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void atomic_max_pos_double(double *addr, double v)
{
    if (v > 0) atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// ---- K8 materialising evaluation for pmv_ba_eval (raw r / J as the cost function returns them) ------
__global__ void __launch_bounds__(128)
ba_eval_kernel(const double *__restrict__ poses, const double *__restrict__ points, const double *__restrict__ obs,
               const int *__restrict__ cam, const int *__restrict__ pt, int No, double fx, double cx, double fy,
               double cy, double delta, double *__restrict__ r_out, double *__restrict__ jc_out,
               double *__restrict__ jp_out, double *__restrict__ cost)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    double c = 0;
    if (i < No) {
        double r[2], jc[12], jp[6];
        ba_residual_jac(poses + 6 * cam[i], points + 3 * pt[i], obs[2 * i], obs[2 * i + 1], fx, cx, fy, cy, r, jc, jp);
        double rho0, rho1;
        ba_huber(delta, r[0] * r[0] + r[1] * r[1], rho0, rho1);
        c = 0.5 * rho0;
        if (r_out) { r_out[2 * i] = r[0]; r_out[2 * i + 1] = r[1]; }
        if (jc_out) for (int k = 0; k < 12; k++) jc_out[12 * (size_t)i + k] = jc[k];
        if (jp_out) for (int k = 0; k < 6; k++) jp_out[6 * (size_t)i + k] = jp[k];
    }
    c = warp_sum_d(c);
    if ((threadIdx.x & 31) == 0 && c != 0) atomicAdd(cost, c);
}

// ---- K8 linearisation at x (Corrector applied): one thread per observation ----------------------------
__global__ void __launch_bounds__(128) ba_linearize_kernel(const BADev D)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    const bool in = i < D.No;
    const int w = in ? D.obs_win[i] : -1;
    double c = 0;
    bool act = false;
    if (in) {
        const BAState *st = &D.st[w];
        act = !st->done && st->need_linearize;
        if (act) {
            double r[2], jc[12], jp[6];
            ba_residual_jac(D.poses + 6 * ((size_t)w * D.Nc + D.obs_cam[i]), D.points + 3 * ((size_t)w * D.Np + D.obs_pt[i]),
                            D.obs_xy[2 * i], D.obs_xy[2 * i + 1], D.fx, D.cx, D.fy, D.cy, r, jc, jp);
            double rho0, rho1;
            ba_huber(D.delta, r[0] * r[0] + r[1] * r[1], rho0, rho1);
            c = 0.5 * rho0;
            const double sr = sqrt(rho1);  // Corrector, rho'' <= 0 branch: residual and Jacobian scale by sqrt(rho')
            D.Lr[2 * (size_t)i] = r[0] * sr; D.Lr[2 * (size_t)i + 1] = r[1] * sr;
#pragma unroll
            for (int k = 0; k < 12; k++) D.Ljc[12 * (size_t)i + k] = jc[k] * sr;
#pragma unroll
            for (int k = 0; k < 6; k++) D.Ljp[6 * (size_t)i + k] = jp[k] * sr;
        }
    }
    // warp-aggregated cost: one atomic per warp when the whole warp belongs to one window
    const int w0 = __shfl_sync(0xffffffffu, w, 0);
    const bool uniform = __all_sync(0xffffffffu, w == w0 || !in);
    if (uniform) {
        c = warp_sum_d(c);
        if ((threadIdx.x & 31) == 0 && w0 >= 0 && c != 0) atomicAdd(&D.st[w0].new_cost, c);
    } else if (act) {
        atomicAdd(&D.st[w].new_cost, c);
    }
}

// ---- K9a per-camera blocks: raw J_c^T J_c (21 unique) and J_c^T r (6), one warp per (window, camera) ----
// out: Uraw[27 * (w*Nc + c)] = 21 upper-triangular entries (row major) followed by 6 gradient entries
__global__ void __launch_bounds__(128) ba_cam_accumulate_kernel(const BADev D, double *__restrict__ Uraw)
{
    const int wc = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (wc >= D.W * D.Nc) return;
    const int w = wc / D.Nc, lane = threadIdx.x & 31;
    const BAState *st = &D.st[w];
    if (st->done || !st->need_linearize) return;
    double acc[27];
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0;
    for (int j = D.cam_off[wc] + lane; j < D.cam_off[wc + 1]; j += 32) {
        const size_t i = D.cam_obs[j];
        double jc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) jc[k] = D.Ljc[12 * i + k];
        const double r0 = D.Lr[2 * i], r1 = D.Lr[2 * i + 1];
        int t = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) {
#pragma unroll
            for (int b = a; b < 6; b++) acc[t++] += jc[a] * jc[b] + jc[6 + a] * jc[6 + b];
        }
#pragma unroll
        for (int a = 0; a < 6; a++) acc[21 + a] += jc[a] * r0 + jc[6 + a] * r1;
    }
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = warp_sum_d(acc[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 27; k++) Uraw[27 * (size_t)wc + k] = acc[k];
    }
}

// Same sums for cameras with thousands of observations (BAL scale: 5 000 per camera): one CTA of 8 warps per
// (window, camera) and a shared-memory reduction -- a single warp per camera leaves the GPU at 7 warps per SM
// walking 156 dependent gathers each (0.69 ms at 1 000 cameras x 5 M observations).
__global__ void __launch_bounds__(256) ba_cam_accumulate_wide_kernel(const BADev D, double *__restrict__ Uraw)
{
    __shared__ double part[8][27];
    const int wc = blockIdx.x;
    const int w = wc / D.Nc, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const BAState *st = &D.st[w];
    if (st->done || !st->need_linearize) return;
    double acc[27];
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0;
    for (int j = D.cam_off[wc] + threadIdx.x; j < D.cam_off[wc + 1]; j += 256) {
        const size_t i = D.cam_obs[j];
        double jc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) jc[k] = D.Ljc[12 * i + k];
        const double r0 = D.Lr[2 * i], r1 = D.Lr[2 * i + 1];
        int t = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) {
#pragma unroll
            for (int b = a; b < 6; b++) acc[t++] += jc[a] * jc[b] + jc[6 + a] * jc[6 + b];
        }
#pragma unroll
        for (int a = 0; a < 6; a++) acc[21 + a] += jc[a] * r0 + jc[6 + a] * r1;
    }
#pragma unroll
    for (int k = 0; k < 27; k++) {
        acc[k] = warp_sum_d(acc[k]);
        if (lane == 0) part[warp][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 27) {
        double v = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) v += part[q][threadIdx.x];
        Uraw[27 * (size_t)wc + threadIdx.x] = v;
    }
}

// finalize per camera (after the optional cross-rank reduction of Uraw): Jacobi scale, LM diagonal,
// scaled U (36) and gradient.  One thread per (window, camera).
__global__ void __launch_bounds__(128) ba_cam_finalize_kernel(const BADev D, const double *__restrict__ Uraw)
{
    const int wc = blockIdx.x * 128 + threadIdx.x;
    if (wc >= D.W * D.Nc) return;
    const int w = wc / D.Nc;
    BAState *st = &D.st[w];
    if (st->done || !st->need_linearize) return;
    const double *u = Uraw + 27 * (size_t)wc;
    double s[6];
    const int dg[6] = {0, 6, 11, 15, 18, 20};
#pragma unroll
    for (int a = 0; a < 6; a++) {
        if (!st->scale_ready) D.scale_c[6 * (size_t)wc + a] = 1.0 / (1.0 + sqrt(u[dg[a]]));
        s[a] = D.scale_c[6 * (size_t)wc + a];
    }
    int t = 0;
    double gm = 0;
#pragma unroll
    for (int a = 0; a < 6; a++) {
#pragma unroll
        for (int b = a; b < 6; b++) {
            const double v = u[t++] * s[a] * s[b];
            D.U[36 * (size_t)wc + 6 * a + b] = v;
            D.U[36 * (size_t)wc + 6 * b + a] = v;
            if (a == b) D.diag_c[6 * (size_t)wc + a] = fmin(fmax(v, 1e-6), 1e32);
        }
        D.gc[6 * (size_t)wc + a] = u[21 + a] * s[a];
        gm = fmax(gm, fabs(u[21 + a]));
    }
    atomic_max_pos_double(&st->gmax, gm);
}

// ---- assemble: S = 0, rhs = 0 ; latch the cost of the new linearisation --------------------------------
__global__ void __launch_bounds__(256) ba_clear_system_kernel(const BADev D)
{
    const int w = blockIdx.y;
    BAState *st = &D.st[w];
    if (st->done) return;
    const size_t nn = (size_t)D.n * D.n;
    if (D.W == 1 && D.chol_lim != nullptr && D.n > 160) {
        // One large problem with a known envelope: S was zeroed once at creation and nothing ever writes outside the
        // envelope, so only that part is cleared (12 MB instead of 288 MB at BASELINE config 5).  Row r: from the first
        // column of its camera's diagonal block to the end of the envelope rounded up to the factorisation's tiles + one.
        const int n = D.n;
        for (int r = blockIdx.x; r < n; r += gridDim.x) {
            const int c0 = r / 6 * 6;
            int c1 = (D.chol_lim[r / PMV_CHOL_NB] + PMV_CHOL_NB - 1) / PMV_CHOL_NB * PMV_CHOL_NB + PMV_CHOL_NB;
            c1 = c1 < n ? c1 : n;
            for (int c = c0 + threadIdx.x; c < c1; c += 256) D.S[(size_t)r * n + c] = 0.0;
        }
    } else {
        for (size_t i = blockIdx.x * 256 + threadIdx.x; i < nn; i += (size_t)gridDim.x * 256) D.S[(size_t)w * nn + i] = 0.0;
    }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < D.n; i += 256) D.rhs[(size_t)w * D.n + i] = 0.0;
    }
}

// latch cost (after the optional cross-rank reduction of new_cost)
__global__ void ba_latch_cost_kernel(const BADev D)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= D.W) return;
    BAState *st = &D.st[w];
    if (st->done || !st->need_linearize) return;
    st->cost = st->new_cost;
    if (st->iter == 0) st->initial_cost = st->cost;
}

// ---- K9b point elimination: one warp per (window, point) -----------------------------------------------
// V = sum J_p^T J_p + D_p^2, g = sum J_p^T r; S(ci,ck) -= W_i V^-1 W_k^T (ci <= ck); rhs_ci -= W_i V^-1 g.
__global__ void __launch_bounds__(128) ba_point_schur_kernel(const BADev D, const int pass2 /* 0: V^-1 / g only, S comes from ba_pair_schur_kernel */)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * 4;
    double gm_acc = 0.0;   // max |gradient| of this warp's points, flushed with ONE atomic per window
    int gm_w = -1;
    for (int wp = blockIdx.x * 4 + (threadIdx.x >> 5); wp < D.W * D.Np; wp += nwarps) {
    const int w = wp / D.Np;
    BAState *st = &D.st[w];
    if (st->done) continue;
    const int o0 = D.pt_off[wp], o1 = D.pt_off[wp + 1];
    if (o1 == o0) continue;
    const bool lin = st->need_linearize != 0;
    if (w != gm_w) {
        if (gm_w >= 0 && lane == 0) atomic_max_pos_double(&D.st[gm_w].gmax, gm_acc);
        gm_w = w; gm_acc = 0.0;
    }
    // pass 1: raw column sums of J_p (3x3 upper + gradient 3)
    double a[9];
#pragma unroll
    for (int k = 0; k < 9; k++) a[k] = 0;
    for (int i = o0 + lane; i < o1; i += 32) {
        double jp[6];
#pragma unroll
        for (int k = 0; k < 6; k++) jp[k] = D.Ljp[6 * (size_t)i + k];
        const double r0 = D.Lr[2 * (size_t)i], r1 = D.Lr[2 * (size_t)i + 1];
        a[0] += jp[0] * jp[0] + jp[3] * jp[3]; a[1] += jp[0] * jp[1] + jp[3] * jp[4]; a[2] += jp[0] * jp[2] + jp[3] * jp[5];
        a[3] += jp[1] * jp[1] + jp[4] * jp[4]; a[4] += jp[1] * jp[2] + jp[4] * jp[5]; a[5] += jp[2] * jp[2] + jp[5] * jp[5];
        a[6] += jp[0] * r0 + jp[3] * r1; a[7] += jp[1] * r0 + jp[4] * r1; a[8] += jp[2] * r0 + jp[5] * r1;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) a[k] = warp_sum_d(a[k]);
    double sp[3];
    if (lin && !st->scale_ready && lane == 0) {
        D.scale_p[3 * (size_t)wp] = 1.0 / (1.0 + sqrt(a[0]));
        D.scale_p[3 * (size_t)wp + 1] = 1.0 / (1.0 + sqrt(a[3]));
        D.scale_p[3 * (size_t)wp + 2] = 1.0 / (1.0 + sqrt(a[5]));
    }
    __syncwarp();
    sp[0] = D.scale_p[3 * (size_t)wp]; sp[1] = D.scale_p[3 * (size_t)wp + 1]; sp[2] = D.scale_p[3 * (size_t)wp + 2];
    // scaled V and g
    double V[6] = {a[0] * sp[0] * sp[0], a[1] * sp[0] * sp[1], a[2] * sp[0] * sp[2],
                   a[3] * sp[1] * sp[1], a[4] * sp[1] * sp[2], a[5] * sp[2] * sp[2]};
    const double g[3] = {a[6] * sp[0], a[7] * sp[1], a[8] * sp[2]};
    if (lin && lane == 0) {
        D.diag_p[3 * (size_t)wp] = fmin(fmax(V[0], 1e-6), 1e32);
        D.diag_p[3 * (size_t)wp + 1] = fmin(fmax(V[3], 1e-6), 1e32);
        D.diag_p[3 * (size_t)wp + 2] = fmin(fmax(V[5], 1e-6), 1e32);
        gm_acc = fmax(gm_acc, fmax(fabs(a[6]), fmax(fabs(a[7]), fabs(a[8]))));
    }
    __syncwarp();
    const double radius = st->radius;
    {
        double d0 = sqrt(D.diag_p[3 * (size_t)wp] / radius), d1 = sqrt(D.diag_p[3 * (size_t)wp + 1] / radius),
               d2 = sqrt(D.diag_p[3 * (size_t)wp + 2] / radius);
        V[0] += d0 * d0; V[3] += d1 * d1; V[5] += d2 * d2;
    }
    // (V)^-1 through LL^T, like Eigen's llt().solve(I)
    double Vi[6];
    {
        const double l00 = sqrt(V[0]), l10 = V[1] / l00, l20 = V[2] / l00;
        const double l11 = sqrt(V[3] - l10 * l10), l21 = (V[4] - l20 * l10) / l11;
        const double l22 = sqrt(V[5] - l20 * l20 - l21 * l21);
        // inverse of L (lower), then Vi = L^-T L^-1
        const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
        const double i10 = -l10 * i00 * i11, i21 = -l21 * i11 * i22, i20 = -(l20 * i00 + l21 * i10) * i22;
        Vi[0] = i00 * i00 + i10 * i10 + i20 * i20; Vi[1] = i10 * i11 + i20 * i21; Vi[2] = i20 * i22;
        Vi[3] = i11 * i11 + i21 * i21; Vi[4] = i21 * i22; Vi[5] = i22 * i22;
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 6; k++) D.Vinv[6 * (size_t)wp + k] = Vi[k];
        D.gp[3 * (size_t)wp] = g[0]; D.gp[3 * (size_t)wp + 1] = g[1]; D.gp[3 * (size_t)wp + 2] = g[2];
    }
    if (!pass2) continue;
    const double vg[3] = {Vi[0] * g[0] + Vi[1] * g[1] + Vi[2] * g[2], Vi[1] * g[0] + Vi[3] * g[1] + Vi[4] * g[2],
                          Vi[2] * g[0] + Vi[4] * g[1] + Vi[5] * g[2]};
    double *S = D.S + (size_t)w * D.n * D.n;
    double *rhs = D.rhs + (size_t)w * D.n;
    // pass 2: lane-strided over observations i; all lanes walk k together (uniform, broadcast loads)
    for (int ib = o0; ib < o1; ib += 32) {
        const int i = ib + lane;
        const bool vi = i < o1;
        int ci = 0;
        double Y[18];
        if (vi) {
            ci = D.obs_cam[i];
            const double *sc = D.scale_c + 6 * ((size_t)w * D.Nc + ci);
            double jc[12], jp[6];
#pragma unroll
            for (int k = 0; k < 6; k++) { jc[k] = D.Ljc[12 * (size_t)i + k] * sc[k]; jc[6 + k] = D.Ljc[12 * (size_t)i + 6 + k] * sc[k]; }
#pragma unroll
            for (int k = 0; k < 3; k++) { jp[k] = D.Ljp[6 * (size_t)i + k] * sp[k]; jp[3 + k] = D.Ljp[6 * (size_t)i + 3 + k] * sp[k]; }
#pragma unroll
            for (int r = 0; r < 6; r++) {
                const double w0 = jc[r] * jp[0] + jc[6 + r] * jp[3], w1 = jc[r] * jp[1] + jc[6 + r] * jp[4],
                             w2 = jc[r] * jp[2] + jc[6 + r] * jp[5];
                Y[3 * r] = w0 * Vi[0] + w1 * Vi[1] + w2 * Vi[2];
                Y[3 * r + 1] = w0 * Vi[1] + w1 * Vi[3] + w2 * Vi[4];
                Y[3 * r + 2] = w0 * Vi[2] + w1 * Vi[4] + w2 * Vi[5];
                atomicAdd(&rhs[6 * ci + r], -(w0 * vg[0] + w1 * vg[1] + w2 * vg[2]));
            }
        }
        for (int k = o0; k < o1; k++) {
            const int ck = D.obs_cam[k];
            if (!vi || ci > ck) continue;
            const double *sc = D.scale_c + 6 * ((size_t)w * D.Nc + ck);
            double jp[6];
#pragma unroll
            for (int q = 0; q < 3; q++) { jp[q] = D.Ljp[6 * (size_t)k + q] * sp[q]; jp[3 + q] = D.Ljp[6 * (size_t)k + 3 + q] * sp[q]; }
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const double j0 = D.Ljc[12 * (size_t)k + c] * sc[c], j1 = D.Ljc[12 * (size_t)k + 6 + c] * sc[c];
                const double w0 = j0 * jp[0] + j1 * jp[3], w1 = j0 * jp[1] + j1 * jp[4], w2 = j0 * jp[2] + j1 * jp[5];
#pragma unroll
                for (int r = 0; r < 6; r++)
                    atomicAdd(&S[(size_t)(6 * ci + r) * D.n + 6 * ck + c], -(Y[3 * r] * w0 + Y[3 * r + 1] * w1 + Y[3 * r + 2] * w2));
            }
        }
    }
    }   // grid-stride loop over points
    if (gm_w >= 0 && lane == 0) atomic_max_pos_double(&D.st[gm_w].gmax, gm_acc);
}

// ---- single large problem (W == 1), few observations per point: G lanes per point instead of a whole warp ----
// BAL-scale problems see every point from ~5 cameras; a warp per point leaves 27 of 32 lanes idle in the two
// per-point kernels (1.5 ms each at 1 M points).  These variants put 32 / G points in a warp.
template <int G>
__device__ __forceinline__ double group_sum_d(double v)
{
#pragma unroll
    for (int o = G / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// V^-1, g, diag_p, scale_p and max |gradient| of every point (what ba_point_schur_kernel does before its pass 2)
template <int G>
__global__ void __launch_bounds__(128) ba_point_vinv_w1_kernel(const BADev D)
{
    BAState *st = &D.st[0];
    if (st->done) return;
    const bool lin = st->need_linearize != 0, scale_ready = st->scale_ready != 0;
    const double radius = st->radius;
    const int lane = threadIdx.x & 31, gl = lane % G, ppw = 32 / G;
    const int nwarps = gridDim.x * 4;
    double gm_acc = 0.0;
    for (int base = (blockIdx.x * 4 + (threadIdx.x >> 5)) * ppw; base < D.Np; base += nwarps * ppw) {
        const int wp = base + lane / G;
        const bool valid = wp < D.Np;
        const int o0 = valid ? D.pt_off[wp] : 0, o1 = valid ? D.pt_off[wp + 1] : 0;
        const bool active = o1 > o0;
        double a[9];
#pragma unroll
        for (int k = 0; k < 9; k++) a[k] = 0;
        for (int i = o0 + gl; i < o1; i += G) {
            double jp[6];
#pragma unroll
            for (int k = 0; k < 6; k++) jp[k] = D.Ljp[6 * (size_t)i + k];
            const double r0 = D.Lr[2 * (size_t)i], r1 = D.Lr[2 * (size_t)i + 1];
            a[0] += jp[0] * jp[0] + jp[3] * jp[3]; a[1] += jp[0] * jp[1] + jp[3] * jp[4]; a[2] += jp[0] * jp[2] + jp[3] * jp[5];
            a[3] += jp[1] * jp[1] + jp[4] * jp[4]; a[4] += jp[1] * jp[2] + jp[4] * jp[5]; a[5] += jp[2] * jp[2] + jp[5] * jp[5];
            a[6] += jp[0] * r0 + jp[3] * r1; a[7] += jp[1] * r0 + jp[4] * r1; a[8] += jp[2] * r0 + jp[5] * r1;
        }
#pragma unroll
        for (int k = 0; k < 9; k++) a[k] = group_sum_d<G>(a[k]);
        if (!active) continue;   // no shuffles below: groups of a warp may part here
        double sp[3];
        if (lin && !scale_ready) {
            sp[0] = 1.0 / (1.0 + sqrt(a[0])); sp[1] = 1.0 / (1.0 + sqrt(a[3])); sp[2] = 1.0 / (1.0 + sqrt(a[5]));
            if (gl == 0) { D.scale_p[3 * (size_t)wp] = sp[0]; D.scale_p[3 * (size_t)wp + 1] = sp[1]; D.scale_p[3 * (size_t)wp + 2] = sp[2]; }
        } else {
            sp[0] = D.scale_p[3 * (size_t)wp]; sp[1] = D.scale_p[3 * (size_t)wp + 1]; sp[2] = D.scale_p[3 * (size_t)wp + 2];
        }
        if (gl != 0) continue;
        double V[6] = {a[0] * sp[0] * sp[0], a[1] * sp[0] * sp[1], a[2] * sp[0] * sp[2],
                       a[3] * sp[1] * sp[1], a[4] * sp[1] * sp[2], a[5] * sp[2] * sp[2]};
        const double g[3] = {a[6] * sp[0], a[7] * sp[1], a[8] * sp[2]};
        double dp[3];
        if (lin) {
            dp[0] = fmin(fmax(V[0], 1e-6), 1e32); dp[1] = fmin(fmax(V[3], 1e-6), 1e32); dp[2] = fmin(fmax(V[5], 1e-6), 1e32);
            D.diag_p[3 * (size_t)wp] = dp[0]; D.diag_p[3 * (size_t)wp + 1] = dp[1]; D.diag_p[3 * (size_t)wp + 2] = dp[2];
            gm_acc = fmax(gm_acc, fmax(fabs(a[6]), fmax(fabs(a[7]), fabs(a[8]))));
        } else {
            dp[0] = D.diag_p[3 * (size_t)wp]; dp[1] = D.diag_p[3 * (size_t)wp + 1]; dp[2] = D.diag_p[3 * (size_t)wp + 2];
        }
        {
            const double d0 = sqrt(dp[0] / radius), d1 = sqrt(dp[1] / radius), d2 = sqrt(dp[2] / radius);
            V[0] += d0 * d0; V[3] += d1 * d1; V[5] += d2 * d2;
        }
        double Vi[6];
        {
            const double l00 = sqrt(V[0]), l10 = V[1] / l00, l20 = V[2] / l00;
            const double l11 = sqrt(V[3] - l10 * l10), l21 = (V[4] - l20 * l10) / l11;
            const double l22 = sqrt(V[5] - l20 * l20 - l21 * l21);
            const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
            const double i10 = -l10 * i00 * i11, i21 = -l21 * i11 * i22, i20 = -(l20 * i00 + l21 * i10) * i22;
            Vi[0] = i00 * i00 + i10 * i10 + i20 * i20; Vi[1] = i10 * i11 + i20 * i21; Vi[2] = i20 * i22;
            Vi[3] = i11 * i11 + i21 * i21; Vi[4] = i21 * i22; Vi[5] = i22 * i22;
        }
#pragma unroll
        for (int k = 0; k < 6; k++) D.Vinv[6 * (size_t)wp + k] = Vi[k];
        D.gp[3 * (size_t)wp] = g[0]; D.gp[3 * (size_t)wp + 1] = g[1]; D.gp[3 * (size_t)wp + 2] = g[2];
    }
    // one atomic per warp
    __syncwarp();
    for (int o = 16; o; o >>= 1) gm_acc = fmax(gm_acc, __shfl_xor_sync(0xffffffffu, gm_acc, o));
    if (lane == 0 && gm_acc > 0.0) atomic_max_pos_double(&st->gmax, gm_acc);
}

// ---- K9b': the same elimination organised by BLOCK of S instead of by point (one problem, W == 1) ---------
// The host lists, for every co-observed camera pair (ci <= ck), the observation pairs (i, k) of the points
// that see both (ba.cu: pair list).  One warp sums a segment of <= 512 such entries of ONE pair in registers,
// Y_i W_k^T = W_i V^-1 W_k^T (6x6), reduces over its lanes and adds the block to S once: 36 atomics per
// segment instead of 36 per entry -- the per-point kernel above spends its time (4.5 ms at 1 M points) in
// 570 M fp64 atomics on ~700 k addresses of S.  Entries with i == k also carry the right-hand side
// rhs_ci -= W_i V^-1 g.  Needs V^-1 and g of every point (ba_point_schur_kernel with pass2 = 0).
struct BAPairSeg { int ci, ck, begin, end; };

__global__ void __launch_bounds__(128, 3) ba_pair_schur_kernel(const BADev D, const BAPairSeg *__restrict__ segs, int nsegs,
                                                            const int2 *__restrict__ entries)
{
    const int lane = threadIdx.x & 31;
    const int sidx = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (sidx >= nsegs) return;
    const BAState *st = &D.st[0];
    if (st->done) return;
    const BAPairSeg sg = segs[sidx];
    double sci[6], sck[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { sci[k] = D.scale_c[6 * (size_t)sg.ci + k]; sck[k] = D.scale_c[6 * (size_t)sg.ck + k]; }
    double acc[36], racc[6];
#pragma unroll
    for (int k = 0; k < 36; k++) acc[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) racc[k] = 0.0;
    for (int e = sg.begin + lane; e < sg.end; e += 32) {
        const int2 en = entries[e];
        const int oi = en.x, ok = en.y;
        const int pt = D.obs_pt[oi];
        double sp[3], Vi[6];
#pragma unroll
        for (int k = 0; k < 3; k++) sp[k] = D.scale_p[3 * (size_t)pt + k];
#pragma unroll
        for (int k = 0; k < 6; k++) Vi[k] = D.Vinv[6 * (size_t)pt + k];
        double Wi[18], Wk[18];
        {
            double jp[6];
#pragma unroll
            for (int k = 0; k < 3; k++) { jp[k] = D.Ljp[6 * (size_t)oi + k] * sp[k]; jp[3 + k] = D.Ljp[6 * (size_t)oi + 3 + k] * sp[k]; }
#pragma unroll
            for (int r = 0; r < 6; r++) {
                const double j0 = D.Ljc[12 * (size_t)oi + r] * sci[r], j1 = D.Ljc[12 * (size_t)oi + 6 + r] * sci[r];
                Wi[3 * r] = j0 * jp[0] + j1 * jp[3]; Wi[3 * r + 1] = j0 * jp[1] + j1 * jp[4]; Wi[3 * r + 2] = j0 * jp[2] + j1 * jp[5];
            }
        }
        if (ok == oi) {
#pragma unroll
            for (int k = 0; k < 18; k++) Wk[k] = Wi[k];
            const double g0 = D.gp[3 * (size_t)pt], g1 = D.gp[3 * (size_t)pt + 1], g2 = D.gp[3 * (size_t)pt + 2];
            const double vg0 = Vi[0] * g0 + Vi[1] * g1 + Vi[2] * g2, vg1 = Vi[1] * g0 + Vi[3] * g1 + Vi[4] * g2,
                         vg2 = Vi[2] * g0 + Vi[4] * g1 + Vi[5] * g2;
#pragma unroll
            for (int r = 0; r < 6; r++) racc[r] += Wi[3 * r] * vg0 + Wi[3 * r + 1] * vg1 + Wi[3 * r + 2] * vg2;
        } else {
            double jp[6];
#pragma unroll
            for (int k = 0; k < 3; k++) { jp[k] = D.Ljp[6 * (size_t)ok + k] * sp[k]; jp[3 + k] = D.Ljp[6 * (size_t)ok + 3 + k] * sp[k]; }
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const double j0 = D.Ljc[12 * (size_t)ok + c] * sck[c], j1 = D.Ljc[12 * (size_t)ok + 6 + c] * sck[c];
                Wk[3 * c] = j0 * jp[0] + j1 * jp[3]; Wk[3 * c + 1] = j0 * jp[1] + j1 * jp[4]; Wk[3 * c + 2] = j0 * jp[2] + j1 * jp[5];
            }
        }
#pragma unroll
        for (int r = 0; r < 6; r++) {
            const double y0 = Wi[3 * r] * Vi[0] + Wi[3 * r + 1] * Vi[1] + Wi[3 * r + 2] * Vi[2];
            const double y1 = Wi[3 * r] * Vi[1] + Wi[3 * r + 1] * Vi[3] + Wi[3 * r + 2] * Vi[4];
            const double y2 = Wi[3 * r] * Vi[2] + Wi[3 * r + 1] * Vi[4] + Wi[3 * r + 2] * Vi[5];
#pragma unroll
            for (int c = 0; c < 6; c++) acc[6 * r + c] += y0 * Wk[3 * c] + y1 * Wk[3 * c + 1] + y2 * Wk[3 * c + 2];
        }
    }
#pragma unroll
    for (int k = 0; k < 36; k++) acc[k] = warp_sum_d(acc[k]);
    double *S = D.S;
    // lane l adds elements l and l + 32 of the 6x6 block (row major)
#pragma unroll
    for (int k = 0; k < 36; k++)
        if (lane == (k & 31)) atomicAdd(&S[(size_t)(6 * sg.ci + k / 6) * D.n + 6 * sg.ck + k % 6], -acc[k]);
    if (sg.ci == sg.ck) {
#pragma unroll
        for (int k = 0; k < 6; k++) racc[k] = warp_sum_d(racc[k]);
#pragma unroll
        for (int k = 0; k < 6; k++)
            if (lane == k) atomicAdd(&D.rhs[6 * sg.ci + k], -racc[k]);
    }
}

// after the optional cross-rank reduction of S / rhs: add the camera diagonal blocks U + D_c^2 and J_c^T r
__global__ void __launch_bounds__(128) ba_add_cam_blocks_kernel(const BADev D)
{
    const int wc = blockIdx.x * 128 + threadIdx.x;
    if (wc >= D.W * D.Nc) return;
    const int w = wc / D.Nc, c = wc - w * D.Nc;
    const BAState *st = &D.st[w];
    if (st->done) return;
    double *S = D.S + (size_t)w * D.n * D.n;
    const double radius = st->radius;
    for (int a = 0; a < 6; a++) {
        for (int b = 0; b < 6; b++) {
            double v = D.U[36 * (size_t)wc + 6 * a + b];
            if (a == b) { const double d = sqrt(D.diag_c[6 * (size_t)wc + a] / radius); v += d * d; }
            S[(size_t)(6 * c + a) * D.n + 6 * c + b] += v;
        }
        D.rhs[(size_t)w * D.n + 6 * c + a] += D.gc[6 * (size_t)wc + a];
    }
}

// ---- K10 (small n): Cholesky of the reduced camera system in shared memory, one CTA per window ------------
__global__ void __launch_bounds__(256) ba_cholesky_small_kernel(const BADev D)
{
    extern __shared__ double sm[];
    const int w = blockIdx.x, n = D.n, tid = threadIdx.x;
    BAState *st = &D.st[w];
    if (st->done) return;
    __shared__ int s_ok;
    if (tid == 0) {
        s_ok = 1;
        // gradient tolerance is checked once the new linearisation is complete (Ceres: before the next iteration)
        if (st->need_linearize && !(st->gmax > 1e-10)) { st->done = 1; st->termination = 3; s_ok = -1; }
    }
    __syncthreads();
    if (s_ok < 0) return;
    double *A = sm;            // n x (n+1) padded, lower triangle holds L
    double *b = sm + (size_t)n * (n + 1);
    const double *S = D.S + (size_t)w * n * n;
    const int ld = n + 1;
    for (int i = tid; i < n * n; i += 256) {
        int r = i / n, c = i - r * n;
        if (c >= r) { double v = S[(size_t)r * n + c]; A[c * ld + r] = v; }   // store lower: A[c][r] = S[r][c]
    }
    for (int i = tid; i < n; i += 256) b[i] = D.rhs[(size_t)w * n + i];
    __syncthreads();
    for (int j = 0; j < n; j++) {
        if (tid == 0) {
            double d = A[j * ld + j];
            if (!(d > 0) || !isfinite(d)) s_ok = 0; else A[j * ld + j] = sqrt(d);
        }
        __syncthreads();
        if (!s_ok) break;
        const double dj = A[j * ld + j];
        for (int i = j + 1 + tid; i < n; i += 256) A[i * ld + j] /= dj;
        __syncthreads();
        // trailing update of the lower triangle: A[i][k] -= A[i][j] * A[k][j], j < k <= i
        const int m = n - j - 1;
        for (int t = tid; t < m * m; t += 256) {
            int ii = t / m, kk = t - ii * m;
            if (kk <= ii) { int i = j + 1 + ii, k = j + 1 + kk; A[i * ld + k] -= A[i * ld + j] * A[k * ld + j]; }
        }
        __syncthreads();
    }
    if (!s_ok) { if (tid == 0) st->chol_ok = 0; return; }
    // forward substitution L z = b (column oriented)
    for (int j = 0; j < n; j++) {
        if (tid == 0) b[j] /= A[j * ld + j];
        __syncthreads();
        const double bj = b[j];
        for (int i = j + 1 + tid; i < n; i += 256) b[i] -= A[i * ld + j] * bj;
        __syncthreads();
    }
    // back substitution L^T y = z
    for (int j = n - 1; j >= 0; j--) {
        if (tid == 0) b[j] /= A[j * ld + j];
        __syncthreads();
        const double bj = b[j];
        for (int i = tid; i < j; i += 256) b[i] -= A[j * ld + i] * bj;
        __syncthreads();
    }
    bool fin = true;
    for (int i = tid; i < n; i += 256) { D.yc[(size_t)w * n + i] = b[i]; fin = fin && isfinite(b[i]); }
    if (!__syncthreads_and(fin)) { if (tid == 0) st->chol_ok = 0; }
    else if (tid == 0) st->chol_ok = 1;
}

// ---- candidate poses: cand = x - y .* scale; |delta|^2, |x|^2 over active cameras -----------------------
__global__ void __launch_bounds__(128) ba_cam_candidate_kernel(const BADev D, int count_norms)
{
    const int wc = blockIdx.x * 128 + threadIdx.x;
    if (wc >= D.W * D.Nc) return;
    const int w = wc / D.Nc;
    BAState *st = &D.st[w];
    if (st->done || !st->chol_ok) return;
    const bool active = D.cam_active[wc] != 0;   // global over ranks: every rank must move the same replicated poses
    double sn = 0, xn = 0;
    for (int a = 0; a < 6; a++) {
        const size_t j = 6 * (size_t)wc + a;
        const double x = D.poses[j];
        const double cnd = active ? x + (-D.yc[j] * D.scale_c[j]) : x;
        D.cand_poses[j] = cnd;
        if (active) { const double d = x - cnd; sn += d * d; xn += x * x; }
    }
    if (active && count_norms) { atomicAdd(&st->step_norm2, sn); atomicAdd(&st->x_norm2, xn); }
}

// ---- K11 back-substitution + model cost change + candidate cost: one warp per (window, point) ----------
__global__ void __launch_bounds__(128) ba_backsub_kernel(const BADev D)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * 4;
    // per-warp partial sums, flushed with one set of atomics per window (a same-address atomic per point
    // serialises in L2: 1M points -> milliseconds)
    double a_mc = 0, a_cc = 0, a_sn = 0, a_xn = 0;
    int a_w = -1;
    auto flush = [&]() {
        if (a_w >= 0 && lane == 0) {
            BAState *fs = &D.st[a_w];
            atomicAdd(&fs->model_change, a_mc); atomicAdd(&fs->cand_cost, a_cc);
            atomicAdd(&fs->step_norm2, a_sn); atomicAdd(&fs->x_norm2, a_xn);
        }
        a_mc = a_cc = a_sn = a_xn = 0;
    };
    for (int wp = blockIdx.x * 4 + (threadIdx.x >> 5); wp < D.W * D.Np; wp += nwarps) {
    const int w = wp / D.Np;
    BAState *st = &D.st[w];
    if (st->done || !st->chol_ok) continue;
    const int o0 = D.pt_off[wp], o1 = D.pt_off[wp + 1];
    if (o1 == o0) {  // inactive point: candidate = x
        if (lane < 3) D.cand_points[3 * (size_t)wp + lane] = D.points[3 * (size_t)wp + lane];
        continue;
    }
    if (w != a_w) { flush(); a_w = w; }
    const double sp[3] = {D.scale_p[3 * (size_t)wp], D.scale_p[3 * (size_t)wp + 1], D.scale_p[3 * (size_t)wp + 2]};
    const double *yc = D.yc + (size_t)w * D.n;
    // t = g_p - sum W_i^T y_c(i) = g_p - sum J_p^T (J_c y_c)
    double t0 = 0, t1 = 0, t2 = 0;
    for (int i = o0 + lane; i < o1; i += 32) {
        const int ci = D.obs_cam[i];
        const double *sc = D.scale_c + 6 * ((size_t)w * D.Nc + ci);
        double jy0 = 0, jy1 = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const double y = yc[6 * ci + k] * sc[k];
            jy0 += D.Ljc[12 * (size_t)i + k] * y; jy1 += D.Ljc[12 * (size_t)i + 6 + k] * y;
        }
        t0 -= D.Ljp[6 * (size_t)i] * sp[0] * jy0 + D.Ljp[6 * (size_t)i + 3] * sp[0] * jy1;
        t1 -= D.Ljp[6 * (size_t)i + 1] * sp[1] * jy0 + D.Ljp[6 * (size_t)i + 4] * sp[1] * jy1;
        t2 -= D.Ljp[6 * (size_t)i + 2] * sp[2] * jy0 + D.Ljp[6 * (size_t)i + 5] * sp[2] * jy1;
    }
    t0 = warp_sum_d(t0) + D.gp[3 * (size_t)wp];
    t1 = warp_sum_d(t1) + D.gp[3 * (size_t)wp + 1];
    t2 = warp_sum_d(t2) + D.gp[3 * (size_t)wp + 2];
    const double *Vi = D.Vinv + 6 * (size_t)wp;
    const double yp[3] = {Vi[0] * t0 + Vi[1] * t1 + Vi[2] * t2, Vi[1] * t0 + Vi[3] * t1 + Vi[4] * t2,
                          Vi[2] * t0 + Vi[4] * t1 + Vi[5] * t2};
    double cand[3], sn = 0, xn = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double x = D.points[3 * (size_t)wp + k];
        cand[k] = x + (-yp[k] * sp[k]);
        const double d = x - cand[k];
        sn += d * d; xn += x * x;
    }
    if (lane < 3) D.cand_points[3 * (size_t)wp + lane] = cand[lane];
    double mc = 0, cc = 0;
    for (int i = o0 + lane; i < o1; i += 32) {
        const int ci = D.obs_cam[i];
        const double *sc = D.scale_c + 6 * ((size_t)w * D.Nc + ci);
        double m0 = 0, m1 = 0;  // J * step, step = -y
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const double y = yc[6 * ci + k] * sc[k];
            m0 -= D.Ljc[12 * (size_t)i + k] * y; m1 -= D.Ljc[12 * (size_t)i + 6 + k] * y;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double y = yp[k] * sp[k];
            m0 -= D.Ljp[6 * (size_t)i + k] * y; m1 -= D.Ljp[6 * (size_t)i + 3 + k] * y;
        }
        mc -= m0 * (D.Lr[2 * (size_t)i] + m0 / 2.0) + m1 * (D.Lr[2 * (size_t)i + 1] + m1 / 2.0);
        double r[2];
        ba_residual_only(D.cand_poses + 6 * ((size_t)w * D.Nc + ci), cand, D.obs_xy[2 * (size_t)i], D.obs_xy[2 * (size_t)i + 1],
                         D.fx, D.cx, D.fy, D.cy, r);
        double rho0, rho1;
        ba_huber(D.delta, r[0] * r[0] + r[1] * r[1], rho0, rho1);
        cc += 0.5 * rho0;
    }
    mc = warp_sum_d(mc); cc = warp_sum_d(cc);
    a_mc += mc; a_cc += cc; a_sn += sn; a_xn += xn;
    }   // grid-stride loop over points
    flush();
}

// back-substitution, W == 1, G lanes per point (see ba_point_vinv_w1_kernel).  On the run-organised path (ba_runs.cuh,
// D.Ljc == nullptr) the linearisation of an observation is recomputed here instead of read back (24 B instead of
// 160 B per observation).
__device__ __forceinline__ void backsub_fetch(const BADev &D, int i, int ci, const double *X, double r[2], double jc[12], double jp[6])
{
    if (D.Ljc) {
        r[0] = D.Lr[2 * (size_t)i]; r[1] = D.Lr[2 * (size_t)i + 1];
#pragma unroll
        for (int k = 0; k < 12; k++) jc[k] = D.Ljc[12 * (size_t)i + k];
#pragma unroll
        for (int k = 0; k < 6; k++) jp[k] = D.Ljp[6 * (size_t)i + k];
    } else {
        const double2 o = *reinterpret_cast<const double2 *>(D.obs_xy + 2 * (size_t)i);
        ba_residual_jac_t(ba_cam_trig_load(D.trig + 8 * (size_t)ci), D.poses + 6 * (size_t)ci + 3, X, o.x, o.y, D.fx, D.cx, D.fy, D.cy, r, jc, jp);
        double rho0, rho1;
        ba_huber(D.delta, r[0] * r[0] + r[1] * r[1], rho0, rho1);
        const double sr = ba_sqrt_rho1(rho1);
        r[0] = r[0] * sr; r[1] = r[1] * sr;
#pragma unroll
        for (int k = 0; k < 12; k++) jc[k] = jc[k] * sr;
#pragma unroll
        for (int k = 0; k < 6; k++) jp[k] = jp[k] * sr;
    }
}

template <int G, int MINB>
__global__ void __launch_bounds__(128, MINB) ba_backsub_w1_kernel(const BADev D)
{
    BAState *st = &D.st[0];
    if (st->done || !st->chol_ok) return;
    const int lane = threadIdx.x & 31, gl = lane % G, ppw = 32 / G;
    const int nwarps = gridDim.x * 4;
    const double *yc = D.yc;
    double a_mc = 0, a_cc = 0, a_sn = 0, a_xn = 0;
    for (int base = (blockIdx.x * 4 + (threadIdx.x >> 5)) * ppw; base < D.Np; base += nwarps * ppw) {
        const int wp = base + lane / G;
        const bool valid = wp < D.Np;
        const int o0 = valid ? D.pt_off[wp] : 0, o1 = valid ? D.pt_off[wp + 1] : 0;
        const bool active = o1 > o0;
        double sp[3] = {1.0, 1.0, 1.0};
        if (active) { sp[0] = D.scale_p[3 * (size_t)wp]; sp[1] = D.scale_p[3 * (size_t)wp + 1]; sp[2] = D.scale_p[3 * (size_t)wp + 2]; }
        double t0 = 0, t1 = 0, t2 = 0;
        double Lr[2] = {0, 0}, Ljc[12], Ljp[6];   // linearisation of this lane's observation (kept for the second loop)
        const bool single = o1 - o0 <= G;
        for (int i = o0 + gl; i < o1; i += G) {
            const int ci = D.obs_cam[i];
            const double *sc = D.scale_c + 6 * (size_t)ci;
            backsub_fetch(D, i, ci, D.points + 3 * (size_t)wp, Lr, Ljc, Ljp);
            double jy0 = 0, jy1 = 0;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const double y = yc[6 * ci + k] * sc[k];
                jy0 += Ljc[k] * y; jy1 += Ljc[6 + k] * y;
            }
            t0 -= Ljp[0] * sp[0] * jy0 + Ljp[3] * sp[0] * jy1;
            t1 -= Ljp[1] * sp[1] * jy0 + Ljp[4] * sp[1] * jy1;
            t2 -= Ljp[2] * sp[2] * jy0 + Ljp[5] * sp[2] * jy1;
        }
        t0 = group_sum_d<G>(t0); t1 = group_sum_d<G>(t1); t2 = group_sum_d<G>(t2);
        double yp[3] = {0, 0, 0}, cand[3] = {0, 0, 0}, sn = 0, xn = 0;
        if (active) {
            t0 += D.gp[3 * (size_t)wp]; t1 += D.gp[3 * (size_t)wp + 1]; t2 += D.gp[3 * (size_t)wp + 2];
            const double *Vi = D.Vinv + 6 * (size_t)wp;
            yp[0] = Vi[0] * t0 + Vi[1] * t1 + Vi[2] * t2; yp[1] = Vi[1] * t0 + Vi[3] * t1 + Vi[4] * t2;
            yp[2] = Vi[2] * t0 + Vi[4] * t1 + Vi[5] * t2;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const double x = D.points[3 * (size_t)wp + k];
                cand[k] = x + (-yp[k] * sp[k]);
                const double d = x - cand[k];
                sn += d * d; xn += x * x;
            }
            if (gl < 3) D.cand_points[3 * (size_t)wp + gl] = cand[gl];
        } else if (valid && gl < 3) {
            D.cand_points[3 * (size_t)wp + gl] = D.points[3 * (size_t)wp + gl];   // inactive point: candidate = x
        }
        double mc = 0, cc = 0;
        for (int i = o0 + gl; i < o1; i += G) {
            const int ci = D.obs_cam[i];
            const double *sc = D.scale_c + 6 * (size_t)ci;
            if (!single) backsub_fetch(D, i, ci, D.points + 3 * (size_t)wp, Lr, Ljc, Ljp);
            double m0 = 0, m1 = 0;  // J * step, step = -y
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const double y = yc[6 * ci + k] * sc[k];
                m0 -= Ljc[k] * y; m1 -= Ljc[6 + k] * y;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const double y = yp[k] * sp[k];
                m0 -= Ljp[k] * y; m1 -= Ljp[3 + k] * y;
            }
            mc -= m0 * (Lr[0] + m0 / 2.0) + m1 * (Lr[1] + m1 / 2.0);
            double r[2];
            if (D.cand_trig)
                ba_residual_only_t(ba_cam_trig_load(D.cand_trig + 8 * (size_t)ci), D.cand_poses + 6 * (size_t)ci + 3, cand,
                                   D.obs_xy[2 * (size_t)i], D.obs_xy[2 * (size_t)i + 1], D.fx, D.cx, D.fy, D.cy, r);
            else
                ba_residual_only(D.cand_poses + 6 * (size_t)ci, cand, D.obs_xy[2 * (size_t)i], D.obs_xy[2 * (size_t)i + 1],
                                 D.fx, D.cx, D.fy, D.cy, r);
            double rho0, rho1;
            ba_huber(D.delta, r[0] * r[0] + r[1] * r[1], rho0, rho1);
            cc += 0.5 * rho0;
        }
        a_mc += mc; a_cc += cc;                      // per-lane partial sums, reduced once at the end
        if (gl == 0) { a_sn += sn; a_xn += xn; }     // once per point
    }
    __syncwarp();
    a_mc = warp_sum_d(a_mc); a_cc = warp_sum_d(a_cc); a_sn = warp_sum_d(a_sn); a_xn = warp_sum_d(a_xn);
    if (lane == 0) {
        atomicAdd(&st->model_change, a_mc); atomicAdd(&st->cand_cost, a_cc);
        atomicAdd(&st->step_norm2, a_sn); atomicAdd(&st->x_norm2, a_xn);
    }
}

// ---- trust-region bookkeeping: one thread per window (Ceres TrustRegionMinimizer / LM strategy) --------
__global__ void ba_lm_update_kernel(const BADev D)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= D.W) return;
    BAState *st = &D.st[w];
    if (st->done) { st->accepted = 0; return; }
    if (D.max_iters == 0) {   // Ceres with max_num_iterations = 0: iteration zero evaluates the cost, no step is taken
        st->done = 1; st->termination = 0; st->accepted = 0; st->need_linearize = 0;
        st->new_cost = 0; st->cand_cost = 0; st->model_change = 0; st->step_norm2 = 0; st->x_norm2 = 0;
        return;
    }
    st->iter++;
    st->scale_ready = 1;
    const bool valid = st->chol_ok && isfinite(st->model_change) && st->model_change > 0.0;
    st->accepted = 0;
    if (!valid) {
        st->need_linearize = 0;
        if (++st->invalid_run >= 5) { st->done = 1; st->termination = 4; }
        else { st->radius /= st->decrease_factor; st->decrease_factor *= 2.0; st->reuse_diagonal = 1; }
    } else {
        st->invalid_run = 0;
        const double cand = isfinite(st->cand_cost) ? st->cand_cost : DBL_MAX;
        const double x_norm = sqrt(st->x_norm2), step_norm = sqrt(st->step_norm2);
        if (step_norm <= 1e-8 * (x_norm + 1e-8)) { st->done = 1; st->termination = 2; st->need_linearize = 0; }
        else if (fabs(st->cost - cand) <= 1e-6 * st->cost) { st->done = 1; st->termination = 1; st->need_linearize = 0; }
        else {
            const double rel = (st->cost - cand) / st->model_change;
            if (rel > 1e-3) {
                st->accepted = 1; st->need_linearize = 1; st->successful_steps++;
                st->cost = cand;
                const double t = 2.0 * rel - 1.0;
                st->radius = fmin(1e16, st->radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
                st->decrease_factor = 2.0; st->reuse_diagonal = 0;
                st->gmax = 0.0;
            } else {
                st->need_linearize = 0;
                st->radius /= st->decrease_factor; st->decrease_factor *= 2.0; st->reuse_diagonal = 1;
            }
        }
    }
    if (!st->done && st->radius < 1e-32) { st->done = 1; st->termination = 5; }
    if (!st->done && st->iter >= D.max_iters) { st->done = 1; st->termination = 0; }
    st->new_cost = 0; st->cand_cost = 0; st->model_change = 0; st->step_norm2 = 0; st->x_norm2 = 0;
}

__global__ void __launch_bounds__(256) ba_accept_kernel(const BADev D)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t ncp = (size_t)D.W * D.Nc * 6, npp = (size_t)D.W * D.Np * 3;
    if (i < ncp) {
        if (D.st[i / (6 * (size_t)D.Nc)].accepted) D.poses[i] = D.cand_poses[i];
    } else if (i < ncp + npp) {
        const size_t j = i - ncp;
        if (D.st[j / (3 * (size_t)D.Np)].accepted) D.points[j] = D.cand_points[j];
    }
}

}  // namespace

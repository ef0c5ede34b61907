// fivept_math.cuh -- fp64 arithmetic of the two-view initialiser behind BaseTriangulator.
//
// Replaces what OpenCVFivePointTri::triangulate asks of OpenCV (reference OpenCVFivePointTri.cpp:25-27):
//   E = cv::findEssentialMat(p1, p2, camera, cv::RANSAC, 0.99, 1, mask)       -- five-point RANSAC, 1000 iterations
//   cv::recoverPose(E, p1, p2, camera, R, t, HUGE_VAL, mask, tri)              -- cheirality vote + triangulation
// OpenCV's calib3d is un-vendored; this restates its published algorithms (Nister's five-point solver in the
// Gauss-Jordan / degree-10 polynomial form, Durand-Kerner roots, Sampson distance, Hartley's linear triangulation) in
// the operation order the library uses where the order decides something visible: the null-space basis (one-sided
// Jacobi SVD with its deterministic completion vectors), the order of the real roots (and so of the candidate models
// of one sample), the float inlier test, the SVD sign conventions of decomposeEssentialMat.
//
// Everything is __host__ __device__: tests/ compiles this header with g++ and pins it to cv2.findEssentialMat /
// cv2.recoverPose without a GPU; the product path is fivept.cu (device only).
#pragma once
#include "pnp_math.cuh"

namespace fivept {

// ---- null space of the 5 x 9 epipolar constraint matrix -----------------------------------------------------------
// cv::SVD::compute(Q, W, U, Vt, MODIFY_A | FULL_UV) on a wide matrix works on Q^T: the five ROWS of Q are rotated
// pairwise until orthogonal (JacobiSVDImpl_, m = 9, n = 5), sorted by norm, normalised; the four missing right
// singular vectors are then COMPLETED from pseudo-random +-1/9 vectors (cv::RNG(0x12345678), bit 8 of each draw)
// orthogonalised twice against everything before them.  EE = rows 5..8 of that Vt.
PNP_HD void null_space_5x9(double At[9][9], double EE[4][9])
{
    const int m = 9, n = 5;
    double W[5];
    for (int i = 0; i < n; i++) {
        double sd = 0.0;
        for (int k = 0; k < m; k++) sd += At[i][k] * At[i][k];
        W[i] = sd;
    }
    const double eps = 2.220446049250313e-16 * 10;
    for (int iter = 0; iter < 30; iter++) {
        bool changed = false;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                double *Ai = At[i], *Aj = At[j];
                double a = W[i], p = 0.0, b = W[j];
                for (int k = 0; k < m; k++) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = hypot(p, beta);
                double c, s;
                if (beta < 0) { const double delta = (gamma - beta) * 0.5; s = sqrt(delta / gamma); c = p / (gamma * s * 2); }
                else { c = sqrt((gamma + beta) / (gamma * 2)); s = p / (gamma * c * 2); }
                a = b = 0.0;
                for (int k = 0; k < m; k++) {
                    const double t0 = c * Ai[k] + s * Aj[k], t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
            }
        if (!changed) break;
    }
    for (int i = 0; i < n; i++) {
        double sd = 0.0;
        for (int k = 0; k < m; k++) sd += At[i][k] * At[i][k];
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < n - 1; i++) {
        int j = i;
        for (int k = i + 1; k < n; k++) if (W[j] < W[k]) j = k;
        if (i != j) {
            const double tw = W[i]; W[i] = W[j]; W[j] = tw;
            for (int k = 0; k < m; k++) { const double t = At[i][k]; At[i][k] = At[j][k]; At[j][k] = t; }
        }
    }
    pnp::CvRng rng;
    rng.state = 0x12345678ull;
    const double minval = 2.2250738585072014e-308;
    for (int i = 0; i < 9; i++) {
        double sd = i < n ? W[i] : 0.0;
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            const double val0 = 1.0 / m;
            for (int k = 0; k < m; k++) At[i][k] = (rng.next() & 256) != 0 ? val0 : -val0;
            for (int iter = 0; iter < 2; iter++)
                for (int j = 0; j < i; j++) {
                    sd = 0.0;
                    for (int k = 0; k < m; k++) sd += At[i][k] * At[j][k];
                    double asum = 0.0;
                    for (int k = 0; k < m; k++) {
                        const double t = At[i][k] - sd * At[j][k];
                        At[i][k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1.0 / asum : 0.0;
                    for (int k = 0; k < m; k++) At[i][k] *= asum;
                }
            sd = 0.0;
            for (int k = 0; k < m; k++) sd += At[i][k] * At[i][k];
            sd = sqrt(sd);
        }
        const double s = sd > minval ? 1.0 / sd : 0.0;
        for (int k = 0; k < m; k++) At[i][k] *= s;
    }
    for (int e = 0; e < 4; e++)
        for (int k = 0; k < 9; k++) EE[e][k] = At[5 + e][k];
}

// ---- the ten cubic constraints on E(x, y, z) = x E0 + y E1 + z E2 + E3 ---------------------------------------------
// polynomials in (x, y, z): linear = 4 coefficients (x, y, z, 1); quadratic = 10 (x2 y2 z2 xy xz yz x y z 1);
// cubic = 20 in Nister's column order, the first ten being eliminated:
//   x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1
PNP_HD int quad_index(int i, int j)     // product of two linear terms (0:x 1:y 2:z 3:1)
{
    const int a = i < j ? i : j, b = i < j ? j : i;
    if (a == b) return a == 3 ? 9 : a;
    if (b == 3) return 6 + a;
    return a == 0 ? (b == 1 ? 3 : 4) : 5;
}

PNP_HD int cubic_index(int q, int l)    // quadratic term q times linear term l
{
    // exponents of the quadratic terms, then add the linear one
    int ex, ey, ez;
    switch (q) {
        case 0: ex = 2; ey = 0; ez = 0; break;
        case 1: ex = 0; ey = 2; ez = 0; break;
        case 2: ex = 0; ey = 0; ez = 2; break;
        case 3: ex = 1; ey = 1; ez = 0; break;
        case 4: ex = 1; ey = 0; ez = 1; break;
        case 5: ex = 0; ey = 1; ez = 1; break;
        case 6: ex = 1; ey = 0; ez = 0; break;
        case 7: ex = 0; ey = 1; ez = 0; break;
        case 8: ex = 0; ey = 0; ez = 1; break;
        default: ex = 0; ey = 0; ez = 0; break;
    }
    if (l == 0) ex++; else if (l == 1) ey++; else if (l == 2) ez++;
    switch (ex * 16 + ey * 4 + ez) {
        case 3 * 16: return 0;             // x3
        case 3 * 4: return 1;              // y3
        case 2 * 16 + 4: return 2;         // x2y
        case 16 + 2 * 4: return 3;         // xy2
        case 2 * 16 + 1: return 4;         // x2z
        case 2 * 16: return 5;             // x2
        case 2 * 4 + 1: return 6;          // y2z
        case 2 * 4: return 7;              // y2
        case 16 + 4 + 1: return 8;         // xyz
        case 16 + 4: return 9;             // xy
        case 16 + 2: return 10;            // xz2
        case 16 + 1: return 11;            // xz
        case 16: return 12;                // x
        case 4 + 2: return 13;             // yz2
        case 4 + 1: return 14;             // yz
        case 4: return 15;                 // y
        case 3: return 16;                 // z3
        case 2: return 17;                 // z2
        case 1: return 18;                 // z
        default: return 19;                // 1
    }
}

PNP_HD void lin_mul(const double a[4], const double b[4], double q[10], double sign)
{
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) q[quad_index(i, j)] += sign * a[i] * b[j];
}

PNP_HD void quad_lin_mul(const double q[10], const double l[4], double c[20], double scale)
{
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 4; j++) c[cubic_index(i, j)] += scale * q[i] * l[j];
}

// the 10 x 20 coefficient matrix of  det E = 0  and  2 E E^T E - tr(E E^T) E = 0; EE[k] is basis matrix k, row major
PNP_HD void constraint_matrix(const double EE[4][9], double A[10][20])
{
    for (int r = 0; r < 10; r++)
        for (int k = 0; k < 20; k++) A[r][k] = 0.0;
    double e[9][4];                                       // entry (i, j) of E as a linear polynomial
    for (int k = 0; k < 9; k++)
        for (int b = 0; b < 4; b++) e[k][b] = EE[b][k];
    // det E by the first row
    {
        double m0[10] = {0}, m1[10] = {0}, m2[10] = {0};
        lin_mul(e[4], e[8], m0, 1.0); lin_mul(e[5], e[7], m0, -1.0);
        lin_mul(e[3], e[8], m1, 1.0); lin_mul(e[5], e[6], m1, -1.0);
        lin_mul(e[3], e[7], m2, 1.0); lin_mul(e[4], e[6], m2, -1.0);
        quad_lin_mul(m0, e[0], A[0], 1.0);
        quad_lin_mul(m1, e[1], A[0], -1.0);
        quad_lin_mul(m2, e[2], A[0], 1.0);
    }
    // G = E E^T (symmetric, quadratic entries), T = G - tr(G)/2 I, rows 1..9 = T E
    double G[6][10];                                      // 00 01 02 11 12 22
    for (int g = 0; g < 6; g++)
        for (int k = 0; k < 10; k++) G[g][k] = 0.0;
    {
        int g = 0;
        for (int i = 0; i < 3; i++)
            for (int j = i; j < 3; j++, g++)
                for (int k = 0; k < 3; k++) lin_mul(e[3 * i + k], e[3 * j + k], G[g], 1.0);
    }
    double half_tr[10];
    for (int k = 0; k < 10; k++) half_tr[k] = 0.5 * (G[0][k] + G[3][k] + G[5][k]);
    for (int k = 0; k < 10; k++) { G[0][k] -= half_tr[k]; G[3][k] -= half_tr[k]; G[5][k] -= half_tr[k]; }
    const int gi[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) quad_lin_mul(G[gi[i][k]], e[3 * k + j], A[1 + 3 * i + j], 1.0);
}

// A[:, 10:20] <- A[:, 0:10]^-1 A[:, 10:20] (Gauss-Jordan, partial pivoting); false when singular
PNP_HD bool reduce_constraints(double A[10][20])
{
    for (int c = 0; c < 10; c++) {
        int piv = c;
        for (int r = c + 1; r < 10; r++) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (fabs(A[piv][c]) < 2.2250738585072014e-308) return false;
        if (piv != c) for (int k = 0; k < 20; k++) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
        const double inv = 1.0 / A[c][c];
        for (int k = c; k < 20; k++) A[c][k] *= inv;
        for (int r = 0; r < 10; r++) {
            if (r == c) continue;
            const double f = A[r][c];
            if (f == 0.0) continue;
            for (int k = c; k < 20; k++) A[r][k] -= f * A[c][k];
        }
    }
    return true;
}

// ---- cv::solvePoly: Durand-Kerner with in-place (Gauss-Seidel) updates, start values (1 + i)^k ---------------------
// OpenCV iterates 300 times unless every correction is exactly zero; here the loop also stops once the corrections
// have been below 4 ulp of the root scale twice in a row (later iterations only move the last bits).
struct Cx { double re, im; };
PNP_HD Cx cmul(Cx a, Cx b) { Cx r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }
PNP_HD Cx cdiv(Cx a, Cx b)
{
    const double t = 1.0 / (b.re * b.re + b.im * b.im);
    Cx r; r.re = (a.re * b.re + a.im * b.im) * t; r.im = (-a.re * b.im + a.im * b.re) * t; return r;
}

PNP_HD int solve_poly10(const double coeffs[11], Cx roots[10])
{
    int n = 10;
    for (; n > 1; n--)
        if (fabs(coeffs[n]) > 2.220446049250313e-16) break;
    Cx p = {1.0, 0.0};
    const Cx r = {1.0, 1.0};
    for (int i = 0; i < n; i++) { roots[i] = p; p = cmul(p, r); }
    int calm = 0;
    for (int iter = 0; iter < 300; iter++) {
        double max_diff = 0.0, scale = 0.0;
        for (int i = 0; i < n; i++) {
            p = roots[i];
            Cx num = {coeffs[n], 0.0}, denom = {coeffs[n], 0.0};
            for (int j = 0; j < n; j++) {
                num = cmul(num, p); num.re += coeffs[n - j - 1];
                if (j != i) {
                    const Cx d = {p.re - roots[j].re, p.im - roots[j].im};
                    if (d.re != 0 || d.im != 0) denom = cmul(denom, d);
                }
            }
            num = cdiv(num, denom);
            roots[i].re = p.re - num.re; roots[i].im = p.im - num.im;
            const double an = sqrt(num.re * num.re + num.im * num.im), ar = fabs(p.re) + fabs(p.im);
            max_diff = an > max_diff ? an : max_diff;
            scale = ar > scale ? ar : scale;
        }
        if (max_diff <= 0) break;
        calm = max_diff <= 8.881784197001252e-16 * (1.0 + scale) ? calm + 1 : 0;
        if (calm >= 2) break;
    }
    for (int i = 0; i < n; i++)
        if (fabs(roots[i].im) < 1e-100) roots[i].im = 0;
    return n;
}

// ---- EMEstimatorCallback::runKernel: up to ten essential matrices through five correspondences --------------------
// x1, x2: normalised image points ((u - cx) / fx, (v - cy) / fy) of the first / second view; E row major, unit
// Frobenius norm, x2^T E x1 = 0.  Returns the number of models (real roots), in OpenCV's order.
PNP_HD int models_from_sample(const double x1[5][2], const double x2[5][2], double E[10][9])
{
    double At[9][9];
    for (int i = 0; i < 5; i++) {
        const double a0 = x1[i][0], a1 = x1[i][1], b0 = x2[i][0], b1 = x2[i][1];
        At[i][0] = a0 * b0; At[i][1] = a1 * b0; At[i][2] = b0;
        At[i][3] = a0 * b1; At[i][4] = a1 * b1; At[i][5] = b1;
        At[i][6] = a0;      At[i][7] = a1;      At[i][8] = 1.0;
    }
    double EE[4][9];
    null_space_5x9(At, EE);
    double A[10][20];
    constraint_matrix(EE, A);
    if (!reduce_constraints(A)) return 0;
    // rows <e> - z<f>, <g> - z<h>, <i> - z<j> of Nister's elimination: B(z) [x y 1]^T = 0 with
    // B row = [x z3 z2 z 1 | y z3 z2 z 1 | z4 z3 z2 z 1]
    double b[3][13];
    for (int i = 0; i < 3; i++) {
        const double *r1 = &A[i * 2 + 4][10], *r2 = &A[i * 2 + 5][10];
        double row1[13] = {0}, row2[13] = {0};
        for (int k = 0; k < 3; k++) { row1[1 + k] = r1[k]; row1[5 + k] = r1[3 + k]; row2[k] = r2[k]; row2[4 + k] = r2[3 + k]; }
        for (int k = 0; k < 4; k++) { row1[9 + k] = r1[6 + k]; row2[8 + k] = r2[6 + k]; }
        for (int k = 0; k < 13; k++) b[i][k] = row1[k] - row2[k];
    }
    // det B(z): degree 10; coefficient arrays in ASCENDING powers
    double c[11];
    for (int k = 0; k < 11; k++) c[k] = 0.0;
    {
        // column polynomials of row j: p0 (deg 3) = b[j][0..3], p1 (deg 3) = b[j][4..7], p2 (deg 4) = b[j][8..12], stored descending
        auto minor_times = [&](int ra, int rb, int ca, int cb, int rc, int cc, double sign) {
            const int off[3] = {0, 4, 8}, deg[3] = {3, 3, 4};
            double q[9];
            for (int k = 0; k < 9; k++) q[k] = 0.0;
            // (b[ra][ca] * b[rb][cb] - b[ra][cb] * b[rb][ca]) ascending
            for (int u = 0; u <= deg[ca]; u++)
                for (int v = 0; v <= deg[cb]; v++) {
                    q[u + v] += b[ra][off[ca] + deg[ca] - u] * b[rb][off[cb] + deg[cb] - v];
                    q[u + v] -= b[rb][off[ca] + deg[ca] - u] * b[ra][off[cb] + deg[cb] - v];
                }
            for (int u = 0; u <= deg[ca] + deg[cb]; u++)
                for (int v = 0; v <= deg[cc]; v++) c[u + v] += sign * q[u] * b[rc][off[cc] + deg[cc] - v];
        };
        // expansion along row 0: b00 (b11 b22 - b12 b21) - b01 (b10 b22 - b12 b20) + b02 (b10 b21 - b11 b20)
        minor_times(1, 2, 1, 2, 0, 0, 1.0);
        minor_times(1, 2, 0, 2, 0, 1, -1.0);
        minor_times(1, 2, 0, 1, 0, 2, 1.0);
    }
    Cx roots[10];
    const int nroots = solve_poly10(c, roots);
    int count = 0;
    for (int i = 0; i < nroots; i++) {
        if (fabs(roots[i].im) > 1e-10) continue;
        const double z1 = roots[i].re, z2 = z1 * z1, z3 = z2 * z1, z4 = z3 * z1;
        double bz[9], w[3], Ut[9], Vt[9];
        for (int j = 0; j < 3; j++) {
            const double *br = b[j];
            bz[3 * j] = br[0] * z3 + br[1] * z2 + br[2] * z1 + br[3];
            bz[3 * j + 1] = br[4] * z3 + br[5] * z2 + br[6] * z1 + br[7];
            bz[3 * j + 2] = br[8] * z4 + br[9] * z3 + br[10] * z2 + br[11] * z1 + br[12];
        }
        pnp::cv_svd<3>(bz, w, Ut, Vt);                     // SVD::solveZ: the right singular vector of the smallest value
        if (fabs(Vt[8]) < 1e-10) continue;
        const double x = Vt[6] / Vt[8], y = Vt[7] / Vt[8];
        double nrm = 0.0;
        for (int k = 0; k < 9; k++) {
            const double v = EE[0][k] * x + EE[1][k] * y + EE[2][k] * z1 + EE[3][k];
            E[count][k] = v;
            nrm += v * v;
        }
        nrm = sqrt(nrm);
        for (int k = 0; k < 9; k++) E[count][k] /= nrm;
        count++;
    }
    return count;
}

// EMEstimatorCallback::computeError + findInliers: Sampson distance in double, stored as float, compared as float
PNP_HD bool pair_is_inlier(const double E[9], const double x1[2], const double x2[2], float thr2)
{
    const double a0 = x1[0], a1 = x1[1], b0 = x2[0], b1 = x2[1];
    const double Ex0 = E[0] * a0 + E[1] * a1 + E[2], Ex1 = E[3] * a0 + E[4] * a1 + E[5], Ex2 = E[6] * a0 + E[7] * a1 + E[8];
    const double Et0 = E[0] * b0 + E[3] * b1 + E[6], Et1 = E[1] * b0 + E[4] * b1 + E[7];
    const double x2tEx1 = b0 * Ex0 + b1 * Ex1 + Ex2;
    const double a = Ex0 * Ex0, b = Ex1 * Ex1, c = Et0 * Et0, d = Et1 * Et1;
    const float err = (float)(x2tEx1 * x2tEx1 / (a + b + c + d));
    return err <= thr2;
}

// ---- cv::decomposeEssentialMat -------------------------------------------------------------------------------------
PNP_HD double det3(const double *M)
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

PNP_HD void decompose_essential(const double E[9], double R1[9], double R2[9], double t[3])
{
    double w[3], Ut[9], Vt[9];
    pnp::cv_svd<3>(E, w, Ut, Vt);
    if (det3(Ut) < 0) for (int k = 0; k < 9; k++) Ut[k] = -Ut[k];
    if (det3(Vt) < 0) for (int k = 0; k < 9; k++) Vt[k] = -Vt[k];
    // U W Vt with W = [0 1 0; -1 0 0; 0 0 1]: (U W) = [-u1 | u0 | u2] by columns
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const double u0 = Ut[i], u1 = Ut[3 + i], u2 = Ut[6 + i];
            R1[3 * i + j] = -u1 * Vt[j] + u0 * Vt[3 + j] + u2 * Vt[6 + j];
            R2[3 * i + j] = u1 * Vt[j] - u0 * Vt[3 + j] + u2 * Vt[6 + j];
        }
    for (int i = 0; i < 3; i++) t[i] = Ut[6 + i];
}

// ---- cv::triangulatePoints for P0 = [I | 0], P1 = [R | t]: homogeneous point = last right singular vector ----------
PNP_HD void triangulate_pair(const double R[9], const double t[3], const double x1[2], const double x2[2], double X[4])
{
    double A[16], w[4], Ut[16], Vt[16];
    const double P0[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    const double P1[12] = {R[0], R[1], R[2], t[0], R[3], R[4], R[5], t[1], R[6], R[7], R[8], t[2]};
    for (int k = 0; k < 4; k++) {
        A[k] = x1[0] * P0[8 + k] - P0[k];
        A[4 + k] = x1[1] * P0[8 + k] - P0[4 + k];
        A[8 + k] = x2[0] * P1[8 + k] - P1[k];
        A[12 + k] = x2[1] * P1[8 + k] - P1[4 + k];
    }
    pnp::cv_svd<4>(A, w, Ut, Vt);
    for (int k = 0; k < 4; k++) X[k] = Vt[12 + k];
}

// the cheirality vote of cv::recoverPose for one pose candidate (distance threshold dist, HUGE_VAL = none)
PNP_HD bool in_front_of_both(const double R[9], const double t[3], const double X[4], double dist)
{
    bool ok = X[2] * X[3] > 0;
    const double q0 = X[0] / X[3], q1 = X[1] / X[3], q2 = X[2] / X[3];
    ok = ok && q2 < dist;
    const double z = R[6] * q0 + R[7] * q1 + R[8] * q2 + t[2];
    return ok && z > 0 && z < dist;
}

}  // namespace fivept

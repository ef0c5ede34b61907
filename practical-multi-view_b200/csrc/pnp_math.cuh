// pnp_math.cuh -- fp64 arithmetic of the RANSAC pose solver behind BasePnPSolver.
//
// Replaces what cv::solvePnPRansac does for OpenCVEPnPSolver::solvePnP (reference OpenCVEPnPSolver.cpp:34-35:
// iterationsCount 100, reprojectionError 8, confidence .99, useExtrinsicGuess true, default flags): OpenCV draws
// 5-point subsets with cv::RNG((uint64)-1), solves each with EPnP (Lepetit, Moreno-Noguer, Fua 2009 -- un-vendored
// OpenCV calib3d, restated here from the published algorithm), counts the points whose reprojection error is within the
// threshold, shortens the loop by the usual confidence rule, and finally minimises the reprojection error over the
// inliers of the best hypothesis starting from the caller's pose.
//
// Everything is __host__ __device__ so that the CPU test suite can compile the very same arithmetic (tests/ builds it
// with g++) and pin it to cv2.solvePnP(SOLVEPNP_EPNP) / cv2.solvePnPRansac without a GPU; the product path is
// pnp.cu (device only).
#pragma once
#include <math.h>

#ifndef PNP_HD
#ifdef __CUDACC__
#define PNP_HD __host__ __device__ __forceinline__
#else
#define PNP_HD inline
#endif
#endif

namespace pnp {

// ---- small dense helpers ---------------------------------------------------------------------------------------
// cyclic Jacobi eigen-decomposition of a symmetric N x N matrix (row major, destroyed): A = V diag(w) V^T, the
// columns of V are the eigenvectors; sorted by DESCENDING eigenvalue (cv::SVD order for a PSD matrix)
template <int N>
PNP_HD void jacobi_eigen(double *A, double *w, double *V)
{
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) V[i * N + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < N; i++) {
            diag += A[i * N + i] * A[i * N + i];
            for (int j = i + 1; j < N; j++) off += A[i * N + j] * A[i * N + j];
        }
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < N - 1; p++)
            for (int q = p + 1; q < N; q++) {
                const double apq = A[p * N + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < N; k++) {   // A <- J^T A J
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq; A[k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; k++) {
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk; A[q * N + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; k++) {
                    const double vkp = V[k * N + p], vkq = V[k * N + q];
                    V[k * N + p] = c * vkp - s * vkq; V[k * N + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < N; i++) w[i] = A[i * N + i];
    for (int i = 0; i < N - 1; i++) {   // selection sort, descending
        int m = i;
        for (int j = i + 1; j < N; j++) if (w[j] > w[m]) m = j;
        if (m != i) {
            const double tw = w[i]; w[i] = w[m]; w[m] = tw;
            for (int k = 0; k < N; k++) { const double tv = V[k * N + i]; V[k * N + i] = V[k * N + m]; V[k * N + m] = tv; }
        }
    }
}

// One-sided (Hestenes) Jacobi SVD of a square matrix in the operation order of cv::SVD (JacobiSVDImpl_ in OpenCV's
// core/lapack.cpp: rows of A^T rotated pairwise until orthogonal, eps = 10 DBL_EPSILON, at most max(N, 30) sweeps,
// singular values sorted descending).  A = U diag(w) V^T; Ut rows are the left, Vt rows the right singular vectors --
// with the SIGNS OpenCV returns, which EPnP's control points and null-space basis depend on.
template <int N>
PNP_HD void cv_svd(const double *A, double *w, double *Ut, double *Vt)
{
    double W[N];
    for (int i = 0; i < N; i++) {
        double sd = 0.0;
        for (int k = 0; k < N; k++) { const double t = A[k * N + i]; Ut[i * N + k] = t; sd += t * t; }   // Ut starts as A^T
        W[i] = sd;
        for (int k = 0; k < N; k++) Vt[i * N + k] = i == k ? 1.0 : 0.0;
    }
    const double eps = 2.220446049250313e-16 * 10;
    const int max_iter = N > 30 ? N : 30;
    for (int iter = 0; iter < max_iter; iter++) {
        bool changed = false;
        for (int i = 0; i < N - 1; i++)
            for (int j = i + 1; j < N; j++) {
                double *Ai = Ut + i * N, *Aj = Ut + j * N;
                double a = W[i], p = 0.0, b = W[j];
                for (int k = 0; k < N; k++) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = hypot(p, beta);
                double c, sn;
                if (beta < 0) { const double delta = (gamma - beta) * 0.5; sn = sqrt(delta / gamma); c = p / (gamma * sn * 2); }
                else { c = sqrt((gamma + beta) / (gamma * 2)); sn = p / (gamma * c * 2); }
                a = b = 0.0;
                for (int k = 0; k < N; k++) {
                    const double t0 = c * Ai[k] + sn * Aj[k], t1 = -sn * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                double *Vi = Vt + i * N, *Vj = Vt + j * N;
                for (int k = 0; k < N; k++) {
                    const double t0 = c * Vi[k] + sn * Vj[k], t1 = -sn * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < N; i++) {
        double sd = 0.0;
        for (int k = 0; k < N; k++) sd += Ut[i * N + k] * Ut[i * N + k];
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < N - 1; i++) {
        int j = i;
        for (int k = i + 1; k < N; k++) if (W[j] < W[k]) j = k;
        if (i != j) {
            const double tw = W[i]; W[i] = W[j]; W[j] = tw;
            for (int k = 0; k < N; k++) {
                double t = Ut[i * N + k]; Ut[i * N + k] = Ut[j * N + k]; Ut[j * N + k] = t;
                t = Vt[i * N + k]; Vt[i * N + k] = Vt[j * N + k]; Vt[j * N + k] = t;
            }
        }
    }
    for (int i = 0; i < N; i++) {
        w[i] = W[i];
        const double sc = W[i] > 2.2250738585072014e-308 ? 1.0 / W[i] : 0.0;
        for (int k = 0; k < N; k++) Ut[i * N + k] *= sc;
    }
}

// least squares x = argmin |A x - b| for an M x N system (M >= N) through the eigen-decomposition of A^T A with the
// small eigenvalues dropped (what an SVD solve returns for a rank-deficient system)
template <int M, int N>
PNP_HD void lstsq(const double *A, const double *b, double *x)
{
    double AtA[N * N], Atb[N], w[N], V[N * N];
    for (int i = 0; i < N; i++) {
        double s = 0.0;
        for (int k = 0; k < M; k++) s += A[k * N + i] * b[k];
        Atb[i] = s;
        for (int j = 0; j < N; j++) {
            double t = 0.0;
            for (int k = 0; k < M; k++) t += A[k * N + i] * A[k * N + j];
            AtA[i * N + j] = t;
        }
    }
    jacobi_eigen<N>(AtA, w, V);
    for (int i = 0; i < N; i++) x[i] = 0.0;
    for (int e = 0; e < N; e++) {
        if (!(w[e] > 1e-24 * w[0])) continue;   // squared singular values: 1e-12 relative on the singular value
        double c = 0.0;
        for (int i = 0; i < N; i++) c += V[i * N + e] * Atb[i];
        c /= w[e];
        for (int i = 0; i < N; i++) x[i] += c * V[i * N + e];
    }
}

// SVD of a 3 x 3 matrix, A = U diag(s) V^T (row major)
PNP_HD void svd3(const double *A, double *U, double *s, double *V)
{
    double AtA[9], w[3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) AtA[3 * i + j] = A[i] * A[j] + A[3 + i] * A[3 + j] + A[6 + i] * A[6 + j];
    jacobi_eigen<3>(AtA, w, V);
    for (int e = 0; e < 3; e++) {
        s[e] = sqrt(w[e] > 0 ? w[e] : 0.0);
        for (int i = 0; i < 3; i++) U[3 * i + e] = A[3 * i] * V[e] + A[3 * i + 1] * V[3 + e] + A[3 * i + 2] * V[6 + e];
    }
    // normalise the columns of U; a (near-)zero singular value takes the cross product of the other two
    for (int e = 0; e < 2; e++) {
        const double nrm = sqrt(U[e] * U[e] + U[3 + e] * U[3 + e] + U[6 + e] * U[6 + e]);
        if (nrm > 0) { U[e] /= nrm; U[3 + e] /= nrm; U[6 + e] /= nrm; }
    }
    if (s[2] > 1e-12 * s[0]) {
        const double nrm = sqrt(U[2] * U[2] + U[5] * U[5] + U[8] * U[8]);
        U[2] /= nrm; U[5] /= nrm; U[8] /= nrm;
    } else {
        U[2] = U[3] * U[7] - U[6] * U[4]; U[5] = U[6] * U[1] - U[0] * U[7]; U[8] = U[0] * U[4] - U[3] * U[1];
    }
}

PNP_HD void rodrigues_to_matrix(const double r[3], double R[9])
{
    const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < 2.220446049250313e-16) { R[0] = R[4] = R[8] = 1; R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0; return; }
    const double c = cos(th), s = sin(th), c1 = 1.0 - c, x = r[0] / th, y = r[1] / th, z = r[2] / th;
    R[0] = c + c1 * x * x; R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y; R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

PNP_HD void matrix_to_rodrigues(const double R[9], double r[3])
{
    // cv::Rodrigues (matrix -> vector): axis from the antisymmetric part, angle from atan2
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    const double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0.0; return; }
        double t;
        t = (R[0] + 1) * 0.5; rx = sqrt(t > 0 ? t : 0.0);
        t = (R[4] + 1) * 0.5; ry = sqrt(t > 0 ? t : 0.0) * (R[1] < 0 ? -1.0 : 1.0);
        t = (R[8] + 1) * 0.5; rz = sqrt(t > 0 ? t : 0.0) * (R[2] < 0 ? -1.0 : 1.0);
        if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
        const double k = theta / sqrt(rx * rx + ry * ry + rz * rz);
        r[0] = rx * k; r[1] = ry * k; r[2] = rz * k;
    } else {
        const double vth = 1.0 / (2.0 * s) * theta;
        r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
    }
}

// ---- EPnP on n points (n <= PNP_MAXPTS; the RANSAC kernel uses 5) ------------------------------------------------
constexpr int PNP_MAXPTS = 8;

struct EPnP {
    int n;
    double fu, fv, uc, vc;
    double pws[PNP_MAXPTS][3], us[PNP_MAXPTS][2], alphas[PNP_MAXPTS][4], pcs[PNP_MAXPTS][3];
    double cws[4][3], ccs[4][3];
    int dbg_which = 0, dbg_gn = 5;

    PNP_HD void choose_control_points()
    {
        cws[0][0] = cws[0][1] = cws[0][2] = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < 3; j++) cws[0][j] += pws[i][j];
        for (int j = 0; j < 3; j++) cws[0][j] /= n;
        double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, w[3], Ut[9], Vt[9];
        for (int i = 0; i < n; i++)
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) C[3 * a + b] += (pws[i][a] - cws[0][a]) * (pws[i][b] - cws[0][b]);
        cv_svd<3>(C, w, Ut, Vt);   // cvSVD(PW0tPW0, DC, UCt, 0, MODIFY_A | U_T)
        for (int i = 1; i < 4; i++) {
            const double k = sqrt(w[i - 1] / n);
            for (int j = 0; j < 3; j++) cws[i][j] = cws[0][j] + k * Ut[3 * (i - 1) + j];
        }
    }

    PNP_HD void compute_barycentric_coordinates()
    {
        double cc[9], ci[9];
        for (int i = 0; i < 3; i++)
            for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
        // inverse of the 3 x 3 matrix (adjugate)
        const double det = cc[0] * (cc[4] * cc[8] - cc[5] * cc[7]) - cc[1] * (cc[3] * cc[8] - cc[5] * cc[6]) + cc[2] * (cc[3] * cc[7] - cc[4] * cc[6]);
        const double id = 1.0 / det;
        ci[0] = (cc[4] * cc[8] - cc[5] * cc[7]) * id; ci[1] = (cc[2] * cc[7] - cc[1] * cc[8]) * id; ci[2] = (cc[1] * cc[5] - cc[2] * cc[4]) * id;
        ci[3] = (cc[5] * cc[6] - cc[3] * cc[8]) * id; ci[4] = (cc[0] * cc[8] - cc[2] * cc[6]) * id; ci[5] = (cc[2] * cc[3] - cc[0] * cc[5]) * id;
        ci[6] = (cc[3] * cc[7] - cc[4] * cc[6]) * id; ci[7] = (cc[1] * cc[6] - cc[0] * cc[7]) * id; ci[8] = (cc[0] * cc[4] - cc[1] * cc[3]) * id;
        for (int i = 0; i < n; i++) {
            double *a = alphas[i];
            for (int j = 0; j < 3; j++)
                a[1 + j] = ci[3 * j] * (pws[i][0] - cws[0][0]) + ci[3 * j + 1] * (pws[i][1] - cws[0][1]) + ci[3 * j + 2] * (pws[i][2] - cws[0][2]);
            a[0] = 1.0 - a[1] - a[2] - a[3];
        }
    }

    PNP_HD static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
    PNP_HD static double dist2(const double *a, const double *b)
    {
        return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
    }

    PNP_HD void compute_ccs_pcs(const double *betas, const double *const v[4])
    {
        for (int i = 0; i < 4; i++)
            for (int k = 0; k < 3; k++) ccs[i][k] = betas[0] * v[0][3 * i + k] + betas[1] * v[1][3 * i + k] + betas[2] * v[2][3 * i + k] + betas[3] * v[3][3 * i + k];
        for (int i = 0; i < n; i++)
            for (int k = 0; k < 3; k++)
                pcs[i][k] = alphas[i][0] * ccs[0][k] + alphas[i][1] * ccs[1][k] + alphas[i][2] * ccs[2][k] + alphas[i][3] * ccs[3][k];
        if (pcs[0][2] < 0.0) {   // solve_for_sign
            for (int i = 0; i < 4; i++)
                for (int k = 0; k < 3; k++) ccs[i][k] = -ccs[i][k];
            for (int i = 0; i < n; i++)
                for (int k = 0; k < 3; k++) pcs[i][k] = -pcs[i][k];
        }
    }

    PNP_HD double estimate_R_and_t(double R[9], double t[3])
    {
        double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
        for (int i = 0; i < n; i++)
            for (int j = 0; j < 3; j++) { pc0[j] += pcs[i][j]; pw0[j] += pws[i][j]; }
        for (int j = 0; j < 3; j++) { pc0[j] /= n; pw0[j] /= n; }
        double ABt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, U[9], s[3], V[9];
        for (int i = 0; i < n; i++)
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 3; k++) ABt[3 * j + k] += (pcs[i][j] - pc0[j]) * (pws[i][k] - pw0[k]);
        cv_svd<3>(ABt, s, U, V);   // U, V hold U^T and V^T here: cvSVD(ABt, D, U, V, MODIFY_A), R = U V^T
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) R[3 * i + j] = U[i] * V[j] + U[3 + i] * V[3 + j] + U[6 + i] * V[6 + j];
        const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
        if (det < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
        for (int i = 0; i < 3; i++) t[i] = pc0[i] - (R[3 * i] * pw0[0] + R[3 * i + 1] * pw0[1] + R[3 * i + 2] * pw0[2]);
        double sum2 = 0.0;   // reprojection_error
        for (int i = 0; i < n; i++) {
            const double Xc = R[0] * pws[i][0] + R[1] * pws[i][1] + R[2] * pws[i][2] + t[0];
            const double Yc = R[3] * pws[i][0] + R[4] * pws[i][1] + R[5] * pws[i][2] + t[1];
            const double iz = 1.0 / (R[6] * pws[i][0] + R[7] * pws[i][1] + R[8] * pws[i][2] + t[2]);
            const double ue = uc + fu * Xc * iz, ve = vc + fv * Yc * iz;
            sum2 += sqrt((us[i][0] - ue) * (us[i][0] - ue) + (us[i][1] - ve) * (us[i][1] - ve));
        }
        return sum2 / n;
    }

    PNP_HD static void gauss_newton(const double *L, const double *rho, double betas[4], int gn_iters = 5)
    {
        for (int it = 0; it < gn_iters; it++) {
            double A[24], b[6], x[4];
            for (int i = 0; i < 6; i++) {
                const double *l = L + 10 * i;
                double *a = A + 4 * i;
                a[0] = 2 * l[0] * betas[0] + l[1] * betas[1] + l[3] * betas[2] + l[6] * betas[3];
                a[1] = l[1] * betas[0] + 2 * l[2] * betas[1] + l[4] * betas[2] + l[7] * betas[3];
                a[2] = l[3] * betas[0] + l[4] * betas[1] + 2 * l[5] * betas[2] + l[8] * betas[3];
                a[3] = l[6] * betas[0] + l[7] * betas[1] + l[8] * betas[2] + 2 * l[9] * betas[3];
                b[i] = rho[i] - (l[0] * betas[0] * betas[0] + l[1] * betas[0] * betas[1] + l[2] * betas[1] * betas[1] + l[3] * betas[0] * betas[2] +
                                 l[4] * betas[1] * betas[2] + l[5] * betas[2] * betas[2] + l[6] * betas[0] * betas[3] + l[7] * betas[1] * betas[3] +
                                 l[8] * betas[2] * betas[3] + l[9] * betas[3] * betas[3]);
            }
            lstsq<6, 4>(A, b, x);
            for (int i = 0; i < 4; i++) betas[i] += x[i];
        }
    }

    // returns the mean reprojection error of the chosen solution; R (row major), t: p_cam = R p_world + t
    PNP_HD double compute_pose(double R[9], double t[3])
    {
        choose_control_points();
        compute_barycentric_coordinates();
        double MtM[144], w[12], V[144];
        for (int i = 0; i < 144; i++) MtM[i] = 0.0;
        for (int i = 0; i < n; i++) {
            double m1[12], m2[12];
            for (int j = 0; j < 4; j++) {
                m1[3 * j] = alphas[i][j] * fu; m1[3 * j + 1] = 0.0; m1[3 * j + 2] = alphas[i][j] * (uc - us[i][0]);
                m2[3 * j] = 0.0; m2[3 * j + 1] = alphas[i][j] * fv; m2[3 * j + 2] = alphas[i][j] * (vc - us[i][1]);
            }
            for (int a = 0; a < 12; a++)
                for (int b = 0; b < 12; b++) MtM[12 * a + b] += m1[a] * m1[b] + m2[a] * m2[b];
        }
        double Vt12[144];
        cv_svd<12>(MtM, w, V, Vt12);   // cvSVD(MtM, D, Ut, 0, MODIFY_A | U_T): V holds U^T
        double vbuf[4][12];   // rows 11, 10, 9, 8 of U^T: the singular vectors of the smallest singular values
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 12; k++) vbuf[j][k] = V[12 * (11 - j) + k];
        const double *v[4] = {vbuf[0], vbuf[1], vbuf[2], vbuf[3]};
        double L[60], rho[6];
        {
            double dv[4][6][3];
            for (int i = 0; i < 4; i++) {
                int a = 0, b = 1;
                for (int j = 0; j < 6; j++) {
                    for (int k = 0; k < 3; k++) dv[i][j][k] = v[i][3 * a + k] - v[i][3 * b + k];
                    b++;
                    if (b > 3) { a++; b = a + 1; }
                }
            }
            for (int i = 0; i < 6; i++) {
                double *row = L + 10 * i;
                row[0] = dot3(dv[0][i], dv[0][i]); row[1] = 2.0 * dot3(dv[0][i], dv[1][i]); row[2] = dot3(dv[1][i], dv[1][i]);
                row[3] = 2.0 * dot3(dv[0][i], dv[2][i]); row[4] = 2.0 * dot3(dv[1][i], dv[2][i]); row[5] = dot3(dv[2][i], dv[2][i]);
                row[6] = 2.0 * dot3(dv[0][i], dv[3][i]); row[7] = 2.0 * dot3(dv[1][i], dv[3][i]); row[8] = 2.0 * dot3(dv[2][i], dv[3][i]);
                row[9] = dot3(dv[3][i], dv[3][i]);
            }
            rho[0] = dist2(cws[0], cws[1]); rho[1] = dist2(cws[0], cws[2]); rho[2] = dist2(cws[0], cws[3]);
            rho[3] = dist2(cws[1], cws[2]); rho[4] = dist2(cws[1], cws[3]); rho[5] = dist2(cws[2], cws[3]);
        }
        double best = 1e300;
        for (int N = 1; N <= 3; N++) {
            double betas[4] = {0, 0, 0, 0};
            if (N == 1) {        // [B11 B12 B13 B14]
                double A[24], b4[4];
                for (int i = 0; i < 6; i++) { A[4 * i] = L[10 * i]; A[4 * i + 1] = L[10 * i + 1]; A[4 * i + 2] = L[10 * i + 3]; A[4 * i + 3] = L[10 * i + 6]; }
                lstsq<6, 4>(A, rho, b4);
                if (b4[0] < 0) { betas[0] = sqrt(-b4[0]); betas[1] = -b4[1] / betas[0]; betas[2] = -b4[2] / betas[0]; betas[3] = -b4[3] / betas[0]; }
                else { betas[0] = sqrt(b4[0]); betas[1] = b4[1] / betas[0]; betas[2] = b4[2] / betas[0]; betas[3] = b4[3] / betas[0]; }
            } else if (N == 2) { // [B11 B12 B22]
                double A[18], b3[3];
                for (int i = 0; i < 6; i++) { A[3 * i] = L[10 * i]; A[3 * i + 1] = L[10 * i + 1]; A[3 * i + 2] = L[10 * i + 2]; }
                lstsq<6, 3>(A, rho, b3);
                if (b3[0] < 0) { betas[0] = sqrt(-b3[0]); betas[1] = b3[2] < 0 ? sqrt(-b3[2]) : 0.0; }
                else { betas[0] = sqrt(b3[0]); betas[1] = b3[2] > 0 ? sqrt(b3[2]) : 0.0; }
                if (b3[1] < 0) betas[0] = -betas[0];
            } else {             // [B11 B12 B22 B13 B23]
                double A[30], b5[5];
                for (int i = 0; i < 6; i++) for (int j = 0; j < 5; j++) A[5 * i + j] = L[10 * i + j];
                lstsq<6, 5>(A, rho, b5);
                if (b5[0] < 0) { betas[0] = sqrt(-b5[0]); betas[1] = b5[2] < 0 ? sqrt(-b5[2]) : 0.0; }
                else { betas[0] = sqrt(b5[0]); betas[1] = b5[2] > 0 ? sqrt(b5[2]) : 0.0; }
                if (b5[1] < 0) betas[0] = -betas[0];
                betas[2] = b5[3] / betas[0];
            }
            if (dbg_which && dbg_which != N) continue;
            gauss_newton(L, rho, betas, dbg_gn);
            double Rn[9], tn[3];
            compute_ccs_pcs(betas, v);
            const double err = estimate_R_and_t(Rn, tn);
            if (err < best) {
                best = err;
                for (int i = 0; i < 9; i++) R[i] = Rn[i];
                for (int i = 0; i < 3; i++) t[i] = tn[i];
            }
        }
        return best;
    }
};

// PnPRansacCallback::computeError: the points are float, cv::projectPoints runs in double and stores float projections;
// the squared error is accumulated in float and compared with the squared threshold.  (Rodrigues round trip of the
// hypothesis, as the callback does: model = [rvec | tvec].)
PNP_HD bool point_is_inlier(const float *X, const float *uv, const double R[9], const double t[3], double fu, double fv, double uc, double vc,
                            float thr2)
{
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0], y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1],
                 z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    const double iz = z ? 1.0 / z : 1.0;
    const float pu = (float)(x * iz * fu + uc), pv = (float)(y * iz * fv + vc);
    const float dx = pu - uv[0], dy = pv - uv[1];
    return dx * dx + dy * dy <= thr2;
}

PNP_HD int count_inliers(const float *X, const float *uv, int n, const double Rin[9], const double t[3], double fu, double fv, double uc,
                         double vc, float thr2, unsigned char *mask)
{
    double r[3], R[9];
    matrix_to_rodrigues(Rin, r);     // the hypothesis travels as a rotation VECTOR through the registrator
    rodrigues_to_matrix(r, R);
    int c = 0;
    for (int i = 0; i < n; i++) {
        const bool in = point_is_inlier(X + 3 * i, uv + 2 * i, R, t, fu, fv, uc, vc, thr2);
        if (mask) mask[i] = in;
        c += in;
    }
    return c;
}

// ---- final refinement: minimise the reprojection error over the inliers (solvePnP, SOLVEPNP_ITERATIVE) ----------
// residual (projected - observed) and its Jacobian with respect to a LOCAL perturbation (dw, dt):
// R <- exp([dw]x) R, t <- t + dt.  The minimiser is the same point whatever the parametrisation.
PNP_HD void reproj_jac(const float *X, const float *uv, const double R[9], const double t[3], double fu, double fv, double uc, double vc,
                       double r[2], double J[12])
{
    const double q0 = R[0] * X[0] + R[1] * X[1] + R[2] * X[2], q1 = R[3] * X[0] + R[4] * X[1] + R[5] * X[2],
                 q2 = R[6] * X[0] + R[7] * X[1] + R[8] * X[2];
    const double p0 = q0 + t[0], p1 = q1 + t[1], p2 = q2 + t[2];
    const double iz = 1.0 / p2;
    r[0] = fu * p0 * iz + uc - uv[0];
    r[1] = fv * p1 * iz + vc - uv[1];
    const double a0 = fu * iz, a2 = -fu * p0 * iz * iz, b1 = fv * iz, b2 = -fv * p1 * iz * iz;
    // d p / d w_k = e_k x q
    J[0] = a2 * q1;            J[1] = a0 * q2 - a2 * q0;  J[2] = -a0 * q1;
    J[3] = a0; J[4] = 0.0; J[5] = a2;
    J[6] = -b1 * q2 + b2 * q1; J[7] = -b2 * q0;           J[8] = b1 * q0;
    J[9] = 0.0; J[10] = b1; J[11] = b2;
}

// one accumulation of the normal equations: H (21 upper entries, row major), g (6), cost
PNP_HD void lm_accumulate(const double r[2], const double J[12], double *H, double *g, double *cost)
{
    int k = 0;
    for (int a = 0; a < 6; a++)
        for (int b = a; b < 6; b++) H[k++] += J[a] * J[b] + J[6 + a] * J[6 + b];
    for (int a = 0; a < 6; a++) g[a] += J[a] * r[0] + J[6 + a] * r[1];
    *cost += r[0] * r[0] + r[1] * r[1];
}

// (H + lambda diag(H)) d = -g by Gaussian elimination with partial pivoting; false when singular
PNP_HD bool lm_solve(const double *H21, const double *g, double lambda, double d[6])
{
    double A[6][7];
    int k = 0;
    for (int a = 0; a < 6; a++)
        for (int b = a; b < 6; b++) { A[a][b] = A[b][a] = H21[k++]; }
    for (int a = 0; a < 6; a++) { A[a][a] += lambda * A[a][a]; A[a][6] = -g[a]; }
    for (int c = 0; c < 6; c++) {
        int piv = c;
        for (int r = c + 1; r < 6; r++) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (fabs(A[piv][c]) < 1e-300) return false;
        if (piv != c) for (int j = 0; j < 7; j++) { const double tmp = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = tmp; }
        for (int r = c + 1; r < 6; r++) {
            const double f = A[r][c] / A[c][c];
            for (int j = c; j < 7; j++) A[r][j] -= f * A[c][j];
        }
    }
    for (int c = 5; c >= 0; c--) {
        double v = A[c][6];
        for (int j = c + 1; j < 6; j++) v -= A[c][j] * d[j];
        d[c] = v / A[c][c];
    }
    return true;
}

// candidate pose of a step
PNP_HD void lm_apply(const double R[9], const double t[3], const double d[6], double Rn[9], double tn[3])
{
    double E[9];
    rodrigues_to_matrix(d, E);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rn[3 * i + j] = E[3 * i] * R[j] + E[3 * i + 1] * R[3 + j] + E[3 * i + 2] * R[6 + j];
    for (int i = 0; i < 3; i++) tn[i] = t[i] + d[3 + i];
}

// cv::RNG (multiply-with-carry), as RANSACPointSetRegistrator seeds it: RNG((uint64)-1)
struct CvRng {
    unsigned long long state;
    PNP_HD CvRng() : state(0xffffffffffffffffull) {}
    PNP_HD unsigned next()
    {
        state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    PNP_HD int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// cv::RANSACUpdateNumIters
PNP_HD int ransac_update_num_iters(double p, double ep, int model_points, int max_iters)
{
    p = p < 0 ? 0 : (p > 1 ? 1 : p);
    ep = ep < 0 ? 0 : (ep > 1 ? 1 : ep);
    double num = 1.0 - p;
    if (num < 2.2250738585072014e-308) num = 2.2250738585072014e-308;
    double denom = 1.0 - pow(1.0 - ep, model_points);
    if (denom < 2.2250738585072014e-308) return 0;
    num = log(num); denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

}  // namespace pnp

// ba.cuh -- shared device structures of the bundle adjuster (K8-K11).
#pragma once
#include "common.cuh"

constexpr int PMV_CHOL_NB = 32;   // block-row height of the blocked HBM Cholesky (ba_chol.cu)

// Per-problem (per-window) Levenberg-Marquardt state, resident on the device so the whole solve runs
// without host synchronisation.  Mirrors Ceres' TrustRegionMinimizer + LevenbergMarquardtStrategy
// members (SURVEY Appx C.3).
struct BAState {
    double cost;            // cost at the current iterate x
    double new_cost;        // accumulator: cost of the latest linearisation
    double cand_cost;       // accumulator: cost at the candidate point
    double model_change;    // accumulator: -(J step)^T (r + J step / 2)
    double step_norm2;      // accumulator: |delta|^2 over active parameters
    double x_norm2;         // accumulator: |x|^2 over active parameters
    double radius;
    double decrease_factor;
    double initial_cost;
    double gmax;            // accumulator: max |gradient| (unscaled), as bits of a non-negative double
    int iter;               // LM iterations executed
    int done;               // termination reached (see termination)
    int termination;        // 0 max iterations, 1 function tol, 2 parameter tol, 3 gradient tol, 4 failure, 5 min radius
    int need_linearize;     // 1 -> x changed (or first iteration): re-evaluate r, J
    int reuse_diagonal;
    int invalid_run;
    int chol_ok;            // reduced system factorised
    int accepted;           // last step accepted (candidate must be copied into x)
    int successful_steps;
    int scale_ready;        // Jacobi scaling computed (first linearisation)
    int pad0, pad1;
};

struct BADev {
    int W, Nc, Np, n, No;
    // topology, observations sorted by (window, point)
    const int *obs_cam, *obs_pt, *obs_win;
    const double *obs_xy;
    const int *pt_off;    // W*Np + 1
    const int *cam_off;   // W*Nc + 1
    const int *cam_obs;   // No, observation indices grouped by (window, camera)
    const int *cam_active; // W*Nc: the camera has an observation on ANY rank (a sharded rank may hold none of them)
    double fx, cx, fy, cy, delta;
    // parameters
    double *poses, *points, *cand_poses, *cand_points;
    // linearisation (Corrector applied, columns NOT yet scaled): 2 + 12 + 6 doubles per observation
    double *Lr, *Ljc, *Ljp;
    double *scale_c, *scale_p;   // Jacobi scaling, fixed after the first linearisation
    double *diag_c, *diag_p;     // clamp(|J_:j|^2) of the scaled Jacobian (LM diagonal before / radius)
    double *U, *gc;              // per camera: scaled J_c^T J_c (36) and J_c^T r (6)
    double *S, *rhs, *yc;        // reduced camera system per window (n x n, upper blocks valid), solution
    double *Vinv, *gp;           // per point: (V + D^2)^-1 (6 unique) and J_p^T r (3)
    double *trig, *cand_trig;    // run-organised path: per-camera CamTrig table (8 doubles) of poses / cand_poses, else nullptr
    BAState *st;
    const int *chol_lim;         // per PMV_CHOL_NB-row block of S: end column of its envelope (blocked Cholesky)
    int max_iters;
};

// Two-sided elimination of a banded reduced camera system (ba_chol.cu: split solve).  The band is cut at a separator M
// of w columns (w >= the envelope width) in the middle: the part above and the part below do not couple, so the
// leading system [top | M] is factorised top-down and the trailing system [bottom | M], index-reversed, also
// top-down -- concurrently, by two clusters -- and the separator's Schur complement is assembled from the two
// partial factors and solved last.  Halves the chain of dependent block steps of the factorisation and of the
// back-substitution.
struct BASplit {
    int enabled = 0;
    int a = 0, w = 0, h1 = 0, h2 = 0;            // separator [a, a + w); h1 = a + w rows of the leading system, h2 = n - a of the reversed one
    double *S2 = nullptr, *b2 = nullptr, *y2 = nullptr;      // reversed trailing system (h2 x h2, own leading dimension), rhs, solution
    double *AM = nullptr, *bM = nullptr;                     // saved separator block (w x w) and rhs
    double *S3 = nullptr, *b3 = nullptr, *y3 = nullptr;      // separator Schur complement, rhs, solution
    int *lim1 = nullptr, *lim2 = nullptr, *lim3 = nullptr;   // envelopes of the three systems (device)
    const int *lim_orig = nullptr;
    BAState *st3 = nullptr;                                  // [3]: private status of the three factorisations
    int *lim1_h = nullptr, *lim2_h = nullptr, *lim3_h = nullptr;   // host copies (owned by the problem)
    cudaStream_t s2 = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

// Partitioned elimination of a long banded reduced camera system (ba_chol.cu: part solve) -- nested dissection with P
// segments.  The band is cut by P - 1 separators M_j of w columns (w >= the envelope width, so neighbouring segments
// do not couple): seg_1 M_1 seg_2 M_2 ... M_{P-1} seg_P.  The extended systems X_i = [seg_i | M_i] (X_P = seg_P) are
// factorised IN PLACE in S, concurrently, each by its own cluster; a segment's coupling to the separator on its LEFT is
// carried as w extra right-hand sides through that factor (the "spike" G_i = U_{X_i}^-T B_i, part_spike_kernel); the
// separators' Schur complement -- block tridiagonal, D_j = U_M(j)^T U_M(j) - G_s(j+1)^T G_s(j+1),
// E_j = H(j+1)^T U_M(j+1), r_j = U_M(j)^T z_M(j) - G_s(j+1)^T z_s(j+1) with G = [G_s; H] split at the segment / separator
// boundary -- is assembled without atomics (every rank of a sharded solve must produce the same bits) and eliminated
// separator by separator (dense w x w factorisations by the same cluster kernel, F_j = U_j^-T E_j by the spike kernel,
// D_{j+1} -= F_j^T F_j), and the segments are back-substituted concurrently.  Chain of dependent 32-row block steps:
// n / 32 -> (n / P + w) / 32 + (P - 1) w / 32.
constexpr int PMV_PART_MAX = 12;
struct BAPart {
    int enabled = 0;
    int P = 0, w = 0, nR = 0;                    // segments, separator width, order of the separator system (P - 1) w
    int a[PMV_PART_MAX] = {0};                   // first row of segment i
    int ns[PMV_PART_MAX] = {0};                  // rows of segment i (without its separator)
    int nx[PMV_PART_MAX] = {0};                  // rows of the extended system X_i
    int *limX[PMV_PART_MAX] = {nullptr};         // envelopes of the extended systems, relative (device)
    int *limX_h[PMV_PART_MAX] = {nullptr};       // host copies (owned by the problem)
    double *G[PMV_PART_MAX] = {nullptr};         // spikes, nx[i] x w row major (i >= 1)
    // separator system (block tridiagonal): diagonal blocks D_j, couplings E_j, F_j = U_j^-T E_j, each w x w (ld = w)
    double *Dsep = nullptr, *E = nullptr, *F = nullptr, *bR = nullptr, *yR = nullptr;
    int *limD = nullptr, *limD_h = nullptr;      // dense envelope of one separator block (w in every entry)
    BAState *stp = nullptr;                      // [2 P - 1]: private status of the P segment and P - 1 separator factorisations
    cudaStream_t str[PMV_PART_MAX] = {nullptr};  // streams of segments 1 .. P-1 (segment 0 runs on the caller's)
    cudaEvent_t ev_fork[2] = {nullptr, nullptr};
    cudaEvent_t ev_join[2][PMV_PART_MAX] = {{nullptr}, {nullptr}};
};

// ba_runs.cuh -- K8 + K9 of ONE large problem (BAL scale, W == 1) organised by RUNS of points that are seen by the
// same tuple of cameras.
//
// Replaces, like ba_kernels.cuh, the evaluator + SchurEliminator of ceres::Solve reached from
// CeresBundleAdjustment.cpp:54-61 (ProjectionResidual.h:38-58 under Jets, Huber Corrector, Jacobi scaling,
// S -= W V^-1 W^T).
//
// Why runs: the reduced camera system has ~15 blocks per point; adding them to S point by point costs 540 fp64 atomics
// per point on a few hundred thousand hot addresses (4.5 ms per iteration at 1 M points), and the per-camera-pair entry
// lists that replaced that (ba_pair_schur_kernel) read every observation's Jacobian ~5 times through gathers (5 GB per
// iteration).  Points that share their camera tuple contribute to the SAME blocks: a warp walks one run, recomputes
// the residual and Jacobian of each observation once from 24 B of observation data (nothing is materialised -- the
// 160 B / observation linearisation buffers are not even allocated on this path), keeps the 6 x 6 blocks of the
// tuple in registers, and touches S once per run.  Structure-from-motion point lists have long runs (a stretch of
// track is seen by the same few poses); when they do not, ba.cu keeps the pair-list path.
#pragma once
#include "ba_kernels.cuh"

namespace {

constexpr int RUN_MAXK = 8;       // observations per point this path handles (lanes = points x observations)
constexpr int RUN_MAXLEN = 48;    // points per run: bounds the imbalance between warps, keeps > 20 k warps at 1 M points

// What a lane keeps for the whole run: slot s of every point of a run is an observation by the SAME camera (points of a
// run share their camera tuple, observations of a point are ordered by camera), so the camera index, its CamTrig entry
// and its translation are loaded once per run instead of once per point -- and the chain of dependent loads of an
// iteration (run_pt -> pt_off -> obs_cam -> pose) shrinks to {point index, first observation} -> {point, observation},
// the first of which is fetched one iteration ahead.  run_pt holds (point, pt_off[point]) pairs.
struct RunCam {
    int cam;
    CamTrig T;
    double tr[3];
};

__device__ __forceinline__ RunCam run_cam_load(const BADev &D, const double *__restrict__ trig, const double *__restrict__ poses,
                                               const int2 *__restrict__ run_pt, int p0, int slot, bool lane_on)
{
    RunCam C;
    C.cam = -1;
    C.T.ct = 1.0; C.T.st = 0.0; C.T.ti = 0.0; C.T.w0 = C.T.w1 = C.T.w2 = 0.0; C.T.small = true;
    C.tr[0] = C.tr[1] = C.tr[2] = 0.0;
    if (lane_on) {
        C.cam = D.obs_cam[run_pt[p0].y + slot];
        C.T = ba_cam_trig_load(trig + 8 * (size_t)C.cam);
        C.tr[0] = poses[6 * (size_t)C.cam + 3]; C.tr[1] = poses[6 * (size_t)C.cam + 4]; C.tr[2] = poses[6 * (size_t)C.cam + 5];
    }
    return C;
}

// linearisation of one observation at x with the Corrector applied: exactly what ba_linearize_kernel stores
__device__ __forceinline__ double run_lin_obs(const BADev &D, const RunCam &C, int i, const double *X, double r[2], double jc[12], double jp[6])
{
    const double2 o = *reinterpret_cast<const double2 *>(D.obs_xy + 2 * (size_t)i);
    ba_residual_jac_t(C.T, C.tr, X, o.x, o.y, D.fx, D.cx, D.fy, D.cy, r, jc, jp);
    double rho0, rho1;
    ba_huber(D.delta, r[0] * r[0] + r[1] * r[1], rho0, rho1);
    const double sr = ba_sqrt_rho1(rho1);
    r[0] = r[0] * sr; r[1] = r[1] * sr;
#pragma unroll
    for (int k = 0; k < 12; k++) jc[k] = jc[k] * sr;
#pragma unroll
    for (int k = 0; k < 6; k++) jp[k] = jp[k] * sr;
    return 0.5 * rho0;
}

// ---- pass A: cost of the linearisation + raw per-camera blocks J_c^T J_c (21) | J_c^T r (6) -----------------------
// Uraw must be zero on entry.  One warp per run; lane = (point of the warp iteration, observation slot).
template <int MINB>
__global__ void __launch_bounds__(128, MINB) ba_run_cam_kernel(const BADev D, const int *__restrict__ run_off, const int2 *__restrict__ run_pt,
                                                               int nruns, double *__restrict__ Uraw)
{
    __shared__ double red[4][32][27];
    __shared__ int cams[4][RUN_MAXK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int run = blockIdx.x * 4 + warp;
    BAState *st = &D.st[0];
    if (run >= nruns || st->done || !st->need_linearize) return;   // warp-uniform; only warp-level sync below
    const int p0 = run_off[run], p1 = run_off[run + 1];
    const int2 first = run_pt[p0];
    const int k = D.pt_off[first.x + 1] - first.y;
    const int ppw = 32 / k, pl = lane / k, slot = lane - pl * k;
    const bool lane_on = pl < ppw;
    const RunCam C = run_cam_load(D, D.trig, D.poses, run_pt, p0, slot, lane_on);
    double acc[27];
#pragma unroll
    for (int v = 0; v < 27; v++) acc[v] = 0.0;
    double cost = 0.0;
    int2 e_next = make_int2(0, 0);
    if (lane_on && p0 + pl < p1) e_next = run_pt[p0 + pl];
    for (int base = p0; base < p1; base += ppw) {
        const int idx = base + pl;
        const int2 e = e_next;
        if (lane_on && idx + ppw < p1) e_next = run_pt[idx + ppw];
        if (lane_on && idx < p1) {
            double r[2], jc[12], jp[6];
            cost += run_lin_obs(D, C, e.y + slot, D.points + 3 * (size_t)e.x, r, jc, jp);
            int t = 0;
#pragma unroll
            for (int a = 0; a < 6; a++) {
#pragma unroll
                for (int b = a; b < 6; b++, t++) acc[t] = fma(jc[a], jc[b], fma(jc[6 + a], jc[6 + b], acc[t]));
            }
#pragma unroll
            for (int a = 0; a < 6; a++) acc[21 + a] = fma(jc[a], r[0], fma(jc[6 + a], r[1], acc[21 + a]));
        }
    }
#pragma unroll
    for (int v = 0; v < 27; v++) red[warp][lane][v] = acc[v];
    if (lane < k) cams[warp][lane] = C.cam;   // lanes 0 .. k-1 are the slots of the first point of the run
    __syncwarp();
    for (int e = lane; e < k * 27; e += 32) {
        const int s = e / 27, v = e - s * 27;
        double sum = 0.0;
        for (int q = 0; q < ppw; q++) sum += red[warp][q * k + s][v];
        atomicAdd(&Uraw[27 * (size_t)cams[warp][s] + v], sum);
    }
    cost = warp_sum_d(cost);
    if (lane == 0 && cost != 0.0) atomicAdd(&st->new_cost, cost);
}

// ---- pass B: point elimination of one run --------------------------------------------------------------------------
// Per point (K lanes): V = sum J_p^T J_p (+ LM diagonal), g = sum J_p^T r, V^-1 (stored for the back-substitution);
// per observation W = J_c^T J_p (6 x 3, Jacobi-scaled), Y = W V^-1.  The warp then adds Y_i W_j^T of every slot pair
// i <= j to its register copy of block (i, j): lane <-> (block, row) items, operands broadcast from shared memory.
template <int K>
__device__ __forceinline__ void run_schur_body(const BADev &D, const int2 *__restrict__ run_pt, int p0, int p1, double (*sm)[36],
                                               int *cams_s, int lane)
{
    constexpr int PPW = 32 / K, NBLK = K * (K + 1) / 2, NITEM = NBLK * 6, IPL = (NITEM + 31) / 32;
    const int pl = lane / K, slot = lane - pl * K;
    const bool lane_on = pl < PPW;
    BAState *st = &D.st[0];
    const bool lin = st->need_linearize != 0, scale_ready = st->scale_ready != 0;
    const double inv_radius = 1.0 / st->radius;
    int it_i[IPL], it_j[IPL], it_r[IPL];
#pragma unroll
    for (int m = 0; m < IPL; m++) {
        const int item = lane + 32 * m;
        int b = item / 6, i = 0;
        it_r[m] = item - b * 6;
        while (i < K - 1 && b >= K - i) { b -= K - i; i++; }
        it_i[m] = item < NITEM ? i : -1;
        it_j[m] = i + b;
    }
    double acc[IPL][6], racc[6];
#pragma unroll
    for (int m = 0; m < IPL; m++)
#pragma unroll
        for (int c = 0; c < 6; c++) acc[m][c] = 0.0;
#pragma unroll
    for (int c = 0; c < 6; c++) racc[c] = 0.0;
    double gm_acc = 0.0;
    const RunCam C = run_cam_load(D, D.trig, D.poses, run_pt, p0, slot, lane_on);
    double sc[6] = {1, 1, 1, 1, 1, 1};
    if (lane_on) {
#pragma unroll
        for (int q = 0; q < 6; q++) sc[q] = D.scale_c[6 * (size_t)C.cam + q];
    }
    int2 e_next = make_int2(0, 0);
    if (lane_on && p0 + pl < p1) e_next = run_pt[p0 + pl];
    for (int base = p0; base < p1; base += PPW) {
        const int idx = base + pl;
        const bool on = lane_on && idx < p1;
        const int2 e = e_next;
        if (lane_on && idx + PPW < p1) e_next = run_pt[idx + PPW];
        const int pt = e.x;
        double r[2] = {0, 0}, jc[12], jp[6] = {0, 0, 0, 0, 0, 0};
        if (on) run_lin_obs(D, C, e.y + slot, D.points + 3 * (size_t)pt, r, jc, jp);
        // V (6 unique) and g (3) of the point: sum over its K lanes
        double a[9];
        a[0] = jp[0] * jp[0] + jp[3] * jp[3]; a[1] = jp[0] * jp[1] + jp[3] * jp[4]; a[2] = jp[0] * jp[2] + jp[3] * jp[5];
        a[3] = jp[1] * jp[1] + jp[4] * jp[4]; a[4] = jp[1] * jp[2] + jp[4] * jp[5]; a[5] = jp[2] * jp[2] + jp[5] * jp[5];
        a[6] = jp[0] * r[0] + jp[3] * r[1]; a[7] = jp[1] * r[0] + jp[4] * r[1]; a[8] = jp[2] * r[0] + jp[5] * r[1];
#pragma unroll
        for (int v = 0; v < 9; v++) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < K; q++) s += __shfl_sync(0xffffffffu, a[v], (pl * K + q) & 31);
            a[v] = s;
        }
        if (on) {
            double sp[3];
            if (lin && !scale_ready) {
                sp[0] = 1.0 / (1.0 + sqrt(a[0])); sp[1] = 1.0 / (1.0 + sqrt(a[3])); sp[2] = 1.0 / (1.0 + sqrt(a[5]));
                if (slot == 0) { D.scale_p[3 * (size_t)pt] = sp[0]; D.scale_p[3 * (size_t)pt + 1] = sp[1]; D.scale_p[3 * (size_t)pt + 2] = sp[2]; }
            } else {
                sp[0] = D.scale_p[3 * (size_t)pt]; sp[1] = D.scale_p[3 * (size_t)pt + 1]; sp[2] = D.scale_p[3 * (size_t)pt + 2];
            }
            double V[6] = {a[0] * sp[0] * sp[0], a[1] * sp[0] * sp[1], a[2] * sp[0] * sp[2],
                           a[3] * sp[1] * sp[1], a[4] * sp[1] * sp[2], a[5] * sp[2] * sp[2]};
            const double g[3] = {a[6] * sp[0], a[7] * sp[1], a[8] * sp[2]};
            double dp[3];
            if (lin) {
                dp[0] = fmin(fmax(V[0], 1e-6), 1e32); dp[1] = fmin(fmax(V[3], 1e-6), 1e32); dp[2] = fmin(fmax(V[5], 1e-6), 1e32);
                if (slot == 0) {
                    D.diag_p[3 * (size_t)pt] = dp[0]; D.diag_p[3 * (size_t)pt + 1] = dp[1]; D.diag_p[3 * (size_t)pt + 2] = dp[2];
                    gm_acc = fmax(gm_acc, fmax(fabs(a[6]), fmax(fabs(a[7]), fabs(a[8]))));
                }
            } else {
                dp[0] = D.diag_p[3 * (size_t)pt]; dp[1] = D.diag_p[3 * (size_t)pt + 1]; dp[2] = D.diag_p[3 * (size_t)pt + 2];
            }
            // LM diagonal D^2 = diag / radius (Ceres forms sqrt(diag / radius) and squares it again: the same value to an ulp)
            V[0] = fma(dp[0], inv_radius, V[0]); V[3] = fma(dp[1], inv_radius, V[3]); V[5] = fma(dp[2], inv_radius, V[5]);
            double Vi[6];
            {
                // Cholesky V = L L^T through the reciprocal pivots (three refined rsqrt seeds, no division, no square root)
                const double i00 = ba_rsqrt_fast(V[0]), l10 = V[1] * i00, l20 = V[2] * i00;
                const double i11 = ba_rsqrt_fast(V[3] - l10 * l10), l21 = (V[4] - l20 * l10) * i11;
                const double i22 = ba_rsqrt_fast(V[5] - l20 * l20 - l21 * l21);
                const double i10 = -l10 * i00 * i11, i21 = -l21 * i11 * i22, i20 = -(l20 * i00 + l21 * i10) * i22;
                Vi[0] = i00 * i00 + i10 * i10 + i20 * i20; Vi[1] = i10 * i11 + i20 * i21; Vi[2] = i20 * i22;
                Vi[3] = i11 * i11 + i21 * i21; Vi[4] = i21 * i22; Vi[5] = i22 * i22;
            }
            if (slot == 0) {
#pragma unroll
                for (int q = 0; q < 6; q++) D.Vinv[6 * (size_t)pt + q] = Vi[q];
                D.gp[3 * (size_t)pt] = g[0]; D.gp[3 * (size_t)pt + 1] = g[1]; D.gp[3 * (size_t)pt + 2] = g[2];
            }
            const double jps[6] = {jp[0] * sp[0], jp[1] * sp[1], jp[2] * sp[2], jp[3] * sp[0], jp[4] * sp[1], jp[5] * sp[2]};
            const double vg0 = Vi[0] * g[0] + Vi[1] * g[1] + Vi[2] * g[2], vg1 = Vi[1] * g[0] + Vi[3] * g[1] + Vi[4] * g[2],
                         vg2 = Vi[2] * g[0] + Vi[4] * g[1] + Vi[5] * g[2];
            // W row by row, straight into shared memory with its Y = W V^-1 row (the rows of lanes past the last point of the
            // run are never read)
#pragma unroll
            for (int q = 0; q < 6; q++) {
                const double j0 = jc[q] * sc[q], j1 = jc[6 + q] * sc[q];
                const double w0 = fma(j1, jps[3], j0 * jps[0]), w1 = fma(j1, jps[4], j0 * jps[1]), w2 = fma(j1, jps[5], j0 * jps[2]);
                racc[q] = fma(w2, vg2, fma(w1, vg1, fma(w0, vg0, racc[q])));
                sm[lane][3 * q] = w0; sm[lane][3 * q + 1] = w1; sm[lane][3 * q + 2] = w2;
                sm[lane][18 + 3 * q] = fma(w2, Vi[2], fma(w1, Vi[1], w0 * Vi[0]));
                sm[lane][18 + 3 * q + 1] = fma(w2, Vi[4], fma(w1, Vi[3], w0 * Vi[1]));
                sm[lane][18 + 3 * q + 2] = fma(w2, Vi[5], fma(w1, Vi[4], w0 * Vi[2]));
            }
        }
        __syncwarp();
        const int npts = p1 - base < PPW ? p1 - base : PPW;
#pragma unroll
        for (int m = 0; m < IPL; m++) {
            if (it_i[m] < 0) continue;
            for (int q = 0; q < npts; q++) {
                const double *Yr = &sm[q * K + it_i[m]][18 + 3 * it_r[m]];
                const double *Wj = &sm[q * K + it_j[m]][0];
                const double y0 = Yr[0], y1 = Yr[1], y2 = Yr[2];
#pragma unroll
                for (int c = 0; c < 6; c++) acc[m][c] = fma(y2, Wj[3 * c + 2], fma(y1, Wj[3 * c + 1], fma(y0, Wj[3 * c], acc[m][c])));
            }
        }
        __syncwarp();   // readers done before the next iteration's rows are written
    }
    // ---- flush: the tuple's cameras, then S and rhs once per run
    __syncwarp();
    if (lane < K) cams_s[lane] = C.cam;
#pragma unroll
    for (int c = 0; c < 6; c++) sm[lane][c] = racc[c];
    __syncwarp();
    double *S = D.S;
    const size_t n = (size_t)D.n;
#pragma unroll
    for (int m = 0; m < IPL; m++) {
        if (it_i[m] < 0) continue;
        const int ci = cams_s[it_i[m]], cj = cams_s[it_j[m]], r = it_r[m];
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const double v = -acc[m][c];
            if (ci <= cj) atomicAdd(&S[(6 * (size_t)ci + r) * n + 6 * (size_t)cj + c], v);
            else atomicAdd(&S[(6 * (size_t)cj + c) * n + 6 * (size_t)ci + r], v);           // upper triangle holds the transposed block
            if (ci == cj && it_i[m] != it_j[m]) atomicAdd(&S[(6 * (size_t)ci + c) * n + 6 * (size_t)ci + r], v);   // one camera twice in a tuple
        }
    }
    for (int e = lane; e < K * 6; e += 32) {
        const int s = e / 6, r = e - s * 6;
        double sum = 0.0;
        for (int q = 0; q < PPW; q++) sum += sm[q * K + s][r];
        atomicAdd(&D.rhs[6 * (size_t)cams_s[s] + r], -sum);
    }
    for (int o = 16; o; o >>= 1) gm_acc = fmax(gm_acc, __shfl_xor_sync(0xffffffffu, gm_acc, o));
    if (lane == 0 && gm_acc > 0.0) atomic_max_pos_double(&st->gmax, gm_acc);
}

// one launch per tuple size K (the host orders the runs by K): registers sized for that K's blocks
template <int K, int MINB>
__global__ void __launch_bounds__(128, MINB) ba_run_schur_kernel(const BADev D, const int *__restrict__ run_off, const int2 *__restrict__ run_pt,
                                                                 int run_begin, int run_end)
{
    __shared__ double sm[4][32][36];
    __shared__ int cams[4][RUN_MAXK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int run = run_begin + blockIdx.x * 4 + warp;
    if (run >= run_end || D.st[0].done) return;   // warp-uniform; only warp-level sync below
    run_schur_body<K>(D, run_pt, run_off[run], run_off[run + 1], sm[warp], cams[warp], lane);
}

// resident CTAs per SM the run kernels are compiled for (register cap 255 / 168 / 128), one digit per kernel:
// elimination, camera sums, back-substitution.  Tuning knob (PMV_RUN_MINB=abc); the default is the measured best.
inline int run_minb(int which)
{
    static const int v = getenv("PMV_RUN_MINB") ? atoi(getenv("PMV_RUN_MINB")) : 333;
    const int d = which == 0 ? v / 100 : which == 1 ? (v / 10) % 10 : v % 10;
    return d < 2 ? 2 : d > 4 ? 4 : d;
}

// The launches of the tuple sizes with few runs are latency (a handful of warps, each walking its run alone): they go to
// a side stream beside the launch of the dominant tuple size (fork / join by events, also under graph capture).
struct RunSide {
    cudaStream_t s2 = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
inline int run_dominant_k(const int *kbegin)
{
    int best = 1;
    for (int k = 2; k <= RUN_MAXK; k++) if (kbegin[k + 1] - kbegin[k] > kbegin[best + 1] - kbegin[best]) best = k;
    return best;
}

template <int K>
int launch_run_schur(pmv_ctx *ctx, const BADev &D, const int *run_off, const int2 *run_pt, const int *kbegin, cudaStream_t s)
{
    const int r0 = kbegin[K], r1 = kbegin[K + 1];
    if (r1 > r0) {
        const int mb = run_minb(0);
        if (mb == 4) ba_run_schur_kernel<K, 4><<<(r1 - r0 + 3) / 4, 128, 0, s>>>(D, run_off, run_pt, r0, r1);
        else if (mb == 3) ba_run_schur_kernel<K, 3><<<(r1 - r0 + 3) / 4, 128, 0, s>>>(D, run_off, run_pt, r0, r1);
        else ba_run_schur_kernel<K, 2><<<(r1 - r0 + 3) / 4, 128, 0, s>>>(D, run_off, run_pt, r0, r1);
        PMV_LAUNCH_CHECK(ctx, "ba_run_schur_kernel");
    }
    return PMV_OK;
}

// kbegin[k] .. kbegin[k + 1]: the runs whose points have k observations (k = 1 .. RUN_MAXK)
template <int K>
int launch_run_schur_k(pmv_ctx *ctx, const BADev &D, const int *run_off, const int2 *run_pt, const int *kbegin, int dom, cudaStream_t s,
                       cudaStream_t side)
{
    int rc = launch_run_schur<K>(ctx, D, run_off, run_pt, kbegin, K == dom ? s : side);
    if constexpr (K < RUN_MAXK) { if (!rc) rc = launch_run_schur_k<K + 1>(ctx, D, run_off, run_pt, kbegin, dom, s, side); }
    return rc;
}
inline int launch_run_schur_all(pmv_ctx *ctx, const BADev &D, const int *run_off, const int2 *run_pt, const int *kbegin, cudaStream_t s,
                                const RunSide &side)
{
    const int dom = run_dominant_k(kbegin);
    const bool fork = side.s2 != nullptr;
    if (fork) { PMV_CUDA_TRY(ctx, cudaEventRecord(side.fork, s)); PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(side.s2, side.fork, 0)); }
    int rc = launch_run_schur_k<1>(ctx, D, run_off, run_pt, kbegin, dom, s, fork ? side.s2 : s);
    if (rc) return rc;
    if (fork) { PMV_CUDA_TRY(ctx, cudaEventRecord(side.join, side.s2)); PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(s, side.join, 0)); }
    return PMV_OK;
}

// ---- pass C: back-substitution + model cost change + candidate cost of one run ------------------------------------
// What ba_backsub_w1_kernel does per point, with the lanes of ba_run_schur_kernel (30 of 32 busy at five observations per
// point instead of 20) and the per-camera operands -- CamTrig of the pose and of the candidate pose, translation,
// y_c .* scale_c -- fetched once per run.  Unobserved points are in no run: their candidate stays equal to x
// (cand_points is initialised from points and only observed points are ever written).
template <int K, int MINB>
__global__ void __launch_bounds__(128, MINB) ba_run_backsub_kernel(const BADev D, const int *__restrict__ run_off, const int2 *__restrict__ run_pt,
                                                                   int run_begin, int run_end)
{
    constexpr int PPW = 32 / K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int run = run_begin + blockIdx.x * 4 + warp;
    BAState *st = &D.st[0];
    if (run >= run_end || st->done || !st->chol_ok) return;   // warp-uniform
    const int p0 = run_off[run], p1 = run_off[run + 1];
    const int pl = lane / K, slot = lane - pl * K;
    const bool lane_on = pl < PPW;
    const RunCam C = run_cam_load(D, D.trig, D.poses, run_pt, p0, slot, lane_on);
    CamTrig Tc = C.T;
    double trc[3] = {0, 0, 0}, ysc[6] = {0, 0, 0, 0, 0, 0};
    if (lane_on) {
        Tc = ba_cam_trig_load(D.cand_trig + 8 * (size_t)C.cam);
#pragma unroll
        for (int q = 0; q < 3; q++) trc[q] = D.cand_poses[6 * (size_t)C.cam + 3 + q];
#pragma unroll
        for (int q = 0; q < 6; q++) ysc[q] = D.yc[6 * (size_t)C.cam + q] * D.scale_c[6 * (size_t)C.cam + q];
    }
    double a_mc = 0, a_cc = 0, a_sn = 0, a_xn = 0;
    int2 e_next = make_int2(0, 0);
    if (lane_on && p0 + pl < p1) e_next = run_pt[p0 + pl];
    for (int base = p0; base < p1; base += PPW) {
        const int idx = base + pl;
        const bool on = lane_on && idx < p1;
        const int2 e = e_next;
        if (lane_on && idx + PPW < p1) e_next = run_pt[idx + PPW];
        const int pt = e.x, i = e.y + slot;
        double Lr[2] = {0, 0}, Ljc[12], Ljp[6];
        double X[3] = {0, 0, 0}, sp[3] = {1, 1, 1}, t[3] = {0, 0, 0};
        if (on) {
#pragma unroll
            for (int q = 0; q < 3; q++) { X[q] = D.points[3 * (size_t)pt + q]; sp[q] = D.scale_p[3 * (size_t)pt + q]; }
            run_lin_obs(D, C, i, X, Lr, Ljc, Ljp);
            double jy0 = 0, jy1 = 0;
#pragma unroll
            for (int q = 0; q < 6; q++) { jy0 = fma(Ljc[q], ysc[q], jy0); jy1 = fma(Ljc[6 + q], ysc[q], jy1); }
            t[0] = -(Ljp[0] * sp[0] * jy0 + Ljp[3] * sp[0] * jy1);
            t[1] = -(Ljp[1] * sp[1] * jy0 + Ljp[4] * sp[1] * jy1);
            t[2] = -(Ljp[2] * sp[2] * jy0 + Ljp[5] * sp[2] * jy1);
        }
#pragma unroll
        for (int v = 0; v < 3; v++) {
            double sum = 0.0;
#pragma unroll
            for (int q = 0; q < K; q++) sum += __shfl_sync(0xffffffffu, t[v], (pl * K + q) & 31);
            t[v] = sum;
        }
        if (on) {
            t[0] += D.gp[3 * (size_t)pt]; t[1] += D.gp[3 * (size_t)pt + 1]; t[2] += D.gp[3 * (size_t)pt + 2];
            const double *Vi = D.Vinv + 6 * (size_t)pt;
            const double yp[3] = {Vi[0] * t[0] + Vi[1] * t[1] + Vi[2] * t[2], Vi[1] * t[0] + Vi[3] * t[1] + Vi[4] * t[2],
                                  Vi[2] * t[0] + Vi[4] * t[1] + Vi[5] * t[2]};
            double cand[3], sn = 0, xn = 0;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                cand[q] = X[q] + (-yp[q] * sp[q]);
                const double d = X[q] - cand[q];
                sn += d * d; xn += X[q] * X[q];
            }
            if (slot == 0) {
                D.cand_points[3 * (size_t)pt] = cand[0]; D.cand_points[3 * (size_t)pt + 1] = cand[1]; D.cand_points[3 * (size_t)pt + 2] = cand[2];
                a_sn += sn; a_xn += xn;
            }
            double m0 = 0, m1 = 0;  // J * step, step = -y
#pragma unroll
            for (int q = 0; q < 6; q++) { m0 = fma(-Ljc[q], ysc[q], m0); m1 = fma(-Ljc[6 + q], ysc[q], m1); }
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const double y = yp[q] * sp[q];
                m0 = fma(-Ljp[q], y, m0); m1 = fma(-Ljp[3 + q], y, m1);
            }
            a_mc -= m0 * (Lr[0] + m0 / 2.0) + m1 * (Lr[1] + m1 / 2.0);
            double r[2];
            const double2 o = *reinterpret_cast<const double2 *>(D.obs_xy + 2 * (size_t)i);
            ba_residual_only_t(Tc, trc, cand, o.x, o.y, D.fx, D.cx, D.fy, D.cy, r);
            double rho0, rho1;
            ba_huber(D.delta, r[0] * r[0] + r[1] * r[1], rho0, rho1);
            a_cc += 0.5 * rho0;
        }
    }
    a_mc = warp_sum_d(a_mc); a_cc = warp_sum_d(a_cc); a_sn = warp_sum_d(a_sn); a_xn = warp_sum_d(a_xn);
    if (lane == 0) {
        atomicAdd(&st->model_change, a_mc); atomicAdd(&st->cand_cost, a_cc);
        atomicAdd(&st->step_norm2, a_sn); atomicAdd(&st->x_norm2, a_xn);
    }
}

template <int K>
int launch_run_backsub(pmv_ctx *ctx, const BADev &D, const int *run_off, const int2 *run_pt, const int *kbegin, cudaStream_t s)
{
    const int r0 = kbegin[K], r1 = kbegin[K + 1];
    if (r1 > r0) {
        if (run_minb(2) >= 4) ba_run_backsub_kernel<K, 4><<<(r1 - r0 + 3) / 4, 128, 0, s>>>(D, run_off, run_pt, r0, r1);
        else ba_run_backsub_kernel<K, 3><<<(r1 - r0 + 3) / 4, 128, 0, s>>>(D, run_off, run_pt, r0, r1);
        PMV_LAUNCH_CHECK(ctx, "ba_run_backsub_kernel");
    }
    return PMV_OK;
}

template <int K>
int launch_run_backsub_k(pmv_ctx *ctx, const BADev &D, const int *run_off, const int2 *run_pt, const int *kbegin, int dom, cudaStream_t s,
                         cudaStream_t side)
{
    int rc = launch_run_backsub<K>(ctx, D, run_off, run_pt, kbegin, K == dom ? s : side);
    if constexpr (K < RUN_MAXK) { if (!rc) rc = launch_run_backsub_k<K + 1>(ctx, D, run_off, run_pt, kbegin, dom, s, side); }
    return rc;
}
inline int launch_run_backsub_all(pmv_ctx *ctx, const BADev &D, const int *run_off, const int2 *run_pt, const int *kbegin, cudaStream_t s,
                                  const RunSide &side)
{
    const int dom = run_dominant_k(kbegin);
    const bool fork = side.s2 != nullptr;
    if (fork) { PMV_CUDA_TRY(ctx, cudaEventRecord(side.fork, s)); PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(side.s2, side.fork, 0)); }
    int rc = launch_run_backsub_k<1>(ctx, D, run_off, run_pt, kbegin, dom, s, fork ? side.s2 : s);
    if (rc) return rc;
    if (fork) { PMV_CUDA_TRY(ctx, cudaEventRecord(side.join, side.s2)); PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(s, side.join, 0)); }
    return PMV_OK;
}

}  // namespace

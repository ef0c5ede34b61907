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

// linearisation of one observation at x with the Corrector applied: exactly what ba_linearize_kernel stores
__device__ __forceinline__ double run_lin_obs(const BADev &D, int i, int cam, const double *X, double r[2], double jc[12], double jp[6])
{
    ba_residual_jac(D.poses + 6 * (size_t)cam, X, D.obs_xy[2 * (size_t)i], D.obs_xy[2 * (size_t)i + 1], D.fx, D.cx, D.fy, D.cy, r, jc, jp);
    double rho0, rho1;
    ba_huber(D.delta, r[0] * r[0] + r[1] * r[1], rho0, rho1);
    const double sr = sqrt(rho1);
    r[0] = r[0] * sr; r[1] = r[1] * sr;
#pragma unroll
    for (int k = 0; k < 12; k++) jc[k] = jc[k] * sr;
#pragma unroll
    for (int k = 0; k < 6; k++) jp[k] = jp[k] * sr;
    return 0.5 * rho0;
}

// ---- pass A: cost of the linearisation + raw per-camera blocks J_c^T J_c (21) | J_c^T r (6) -----------------------
// Uraw must be zero on entry.  One warp per run; lane = (point of the warp iteration, observation slot).
__global__ void __launch_bounds__(128) ba_run_cam_kernel(const BADev D, const int *__restrict__ run_off, const int *__restrict__ run_pt,
                                                         int nruns, double *__restrict__ Uraw)
{
    __shared__ double red[4][32][27];
    __shared__ int cams[4][RUN_MAXK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int run = blockIdx.x * 4 + warp;
    BAState *st = &D.st[0];
    if (run >= nruns || st->done || !st->need_linearize) return;   // warp-uniform; only warp-level sync below
    const int p0 = run_off[run], p1 = run_off[run + 1];
    const int first = run_pt[p0];
    const int k = D.pt_off[first + 1] - D.pt_off[first];
    const int ppw = 32 / k, pl = lane / k, slot = lane - pl * k;
    double acc[27];
#pragma unroll
    for (int v = 0; v < 27; v++) acc[v] = 0.0;
    double cost = 0.0;
    int mycam = -1;
    for (int base = p0; base < p1; base += ppw) {
        const int idx = base + pl;
        if (pl < ppw && idx < p1) {
            const int pt = run_pt[idx];
            const int i = D.pt_off[pt] + slot;
            const int cam = D.obs_cam[i];
            mycam = cam;
            double r[2], jc[12], jp[6];
            cost += run_lin_obs(D, i, cam, D.points + 3 * (size_t)pt, r, jc, jp);
            int t = 0;
#pragma unroll
            for (int a = 0; a < 6; a++) {
#pragma unroll
                for (int b = a; b < 6; b++) acc[t++] += jc[a] * jc[b] + jc[6 + a] * jc[6 + b];
            }
#pragma unroll
            for (int a = 0; a < 6; a++) acc[21 + a] += jc[a] * r[0] + jc[6 + a] * r[1];
        }
    }
#pragma unroll
    for (int v = 0; v < 27; v++) red[warp][lane][v] = acc[v];
    if (lane < k) cams[warp][lane] = mycam;   // lanes 0 .. k-1 are the slots of the first point of the run
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
__device__ __forceinline__ void run_schur_body(const BADev &D, const int *__restrict__ run_pt, int p0, int p1, double (*sm)[36],
                                               int *cams_s, int lane)
{
    constexpr int PPW = 32 / K, NBLK = K * (K + 1) / 2, NITEM = NBLK * 6, IPL = (NITEM + 31) / 32;
    const int pl = lane / K, slot = lane - pl * K;
    const bool lane_on = pl < PPW;
    BAState *st = &D.st[0];
    const bool lin = st->need_linearize != 0, scale_ready = st->scale_ready != 0;
    const double radius = st->radius;
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
    int mycam = -1;
    double sc[6] = {1, 1, 1, 1, 1, 1};
    for (int base = p0; base < p1; base += PPW) {
        const int idx = base + pl;
        const bool on = lane_on && idx < p1;
        int pt = 0;
        double r[2] = {0, 0}, jc[12], jp[6] = {0, 0, 0, 0, 0, 0};
        if (on) {
            pt = run_pt[idx];
            const int i = D.pt_off[pt] + slot;
            const int cam = D.obs_cam[i];
            if (mycam < 0) {
                mycam = cam;
#pragma unroll
                for (int q = 0; q < 6; q++) sc[q] = D.scale_c[6 * (size_t)cam + q];
            }
            run_lin_obs(D, i, cam, D.points + 3 * (size_t)pt, r, jc, jp);
        }
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
        double W[18], Y[18];
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
            if (slot == 0) {
#pragma unroll
                for (int q = 0; q < 6; q++) D.Vinv[6 * (size_t)pt + q] = Vi[q];
                D.gp[3 * (size_t)pt] = g[0]; D.gp[3 * (size_t)pt + 1] = g[1]; D.gp[3 * (size_t)pt + 2] = g[2];
            }
            const double jps[6] = {jp[0] * sp[0], jp[1] * sp[1], jp[2] * sp[2], jp[3] * sp[0], jp[4] * sp[1], jp[5] * sp[2]};
#pragma unroll
            for (int q = 0; q < 6; q++) {
                const double j0 = jc[q] * sc[q], j1 = jc[6 + q] * sc[q];
                W[3 * q] = j0 * jps[0] + j1 * jps[3]; W[3 * q + 1] = j0 * jps[1] + j1 * jps[4]; W[3 * q + 2] = j0 * jps[2] + j1 * jps[5];
            }
            const double vg0 = Vi[0] * g[0] + Vi[1] * g[1] + Vi[2] * g[2], vg1 = Vi[1] * g[0] + Vi[3] * g[1] + Vi[4] * g[2],
                         vg2 = Vi[2] * g[0] + Vi[4] * g[1] + Vi[5] * g[2];
#pragma unroll
            for (int q = 0; q < 6; q++) {
                racc[q] += W[3 * q] * vg0 + W[3 * q + 1] * vg1 + W[3 * q + 2] * vg2;
                Y[3 * q] = W[3 * q] * Vi[0] + W[3 * q + 1] * Vi[1] + W[3 * q + 2] * Vi[2];
                Y[3 * q + 1] = W[3 * q] * Vi[1] + W[3 * q + 1] * Vi[3] + W[3 * q + 2] * Vi[4];
                Y[3 * q + 2] = W[3 * q] * Vi[2] + W[3 * q + 1] * Vi[4] + W[3 * q + 2] * Vi[5];
            }
        } else {
#pragma unroll
            for (int q = 0; q < 18; q++) { W[q] = 0.0; Y[q] = 0.0; }
        }
        __syncwarp();   // the previous iteration's readers are done
#pragma unroll
        for (int q = 0; q < 18; q++) { sm[lane][q] = W[q]; sm[lane][18 + q] = Y[q]; }
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
                for (int c = 0; c < 6; c++) acc[m][c] += y0 * Wj[3 * c] + y1 * Wj[3 * c + 1] + y2 * Wj[3 * c + 2];
            }
        }
    }
    // ---- flush: the tuple's cameras, then S and rhs once per run
    __syncwarp();
    if (lane < K) cams_s[lane] = mycam;
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
template <int K>
__global__ void __launch_bounds__(128) ba_run_schur_kernel(const BADev D, const int *__restrict__ run_off, const int *__restrict__ run_pt,
                                                           int run_begin, int run_end)
{
    __shared__ double sm[4][32][36];
    __shared__ int cams[4][RUN_MAXK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int run = run_begin + blockIdx.x * 4 + warp;
    if (run >= run_end || D.st[0].done) return;   // warp-uniform; only warp-level sync below
    run_schur_body<K>(D, run_pt, run_off[run], run_off[run + 1], sm[warp], cams[warp], lane);
}

template <int K>
int launch_run_schur(pmv_ctx *ctx, const BADev &D, const int *run_off, const int *run_pt, const int *kbegin, cudaStream_t s)
{
    const int r0 = kbegin[K], r1 = kbegin[K + 1];
    if (r1 > r0) {
        ba_run_schur_kernel<K><<<(r1 - r0 + 3) / 4, 128, 0, s>>>(D, run_off, run_pt, r0, r1);
        PMV_LAUNCH_CHECK(ctx, "ba_run_schur_kernel");
    }
    return PMV_OK;
}

// kbegin[k] .. kbegin[k + 1]: the runs whose points have k observations (k = 1 .. RUN_MAXK)
inline int launch_run_schur_all(pmv_ctx *ctx, const BADev &D, const int *run_off, const int *run_pt, const int *kbegin, cudaStream_t s)
{
    int rc = launch_run_schur<1>(ctx, D, run_off, run_pt, kbegin, s);
    if (!rc) rc = launch_run_schur<2>(ctx, D, run_off, run_pt, kbegin, s);
    if (!rc) rc = launch_run_schur<3>(ctx, D, run_off, run_pt, kbegin, s);
    if (!rc) rc = launch_run_schur<4>(ctx, D, run_off, run_pt, kbegin, s);
    if (!rc) rc = launch_run_schur<5>(ctx, D, run_off, run_pt, kbegin, s);
    if (!rc) rc = launch_run_schur<6>(ctx, D, run_off, run_pt, kbegin, s);
    if (!rc) rc = launch_run_schur<7>(ctx, D, run_off, run_pt, kbegin, s);
    if (!rc) rc = launch_run_schur<8>(ctx, D, run_off, run_pt, kbegin, s);
    return rc;
}

}  // namespace

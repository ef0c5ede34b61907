// ba_chol.cu -- K10 (large n): blocked dense Cholesky of the reduced camera system S (n = 6 Nc) in
// HBM, for camera counts where it is a real contraction (BASELINE config 5: Nc = 1000, n = 6000,
// n^3/3 = 72 GFLOP).  Replaces the sparse Cholesky Ceres runs on the reduced system
// (SchurComplementSolver, reached from reference CeresBundleAdjustment.cpp:61).
//
// Right-looking U^T U factorisation on the upper triangle (row major), block size 64:
//   chol_diag_kernel    factor S_kk = U_kk^T U_kk in shared memory; z_k = U_kk^-T b_k
//   chol_panel_kernel   U_kj = U_kk^-T S_kj for the block row to the right
//   chol_update_kernel  S_ij -= U_ki^T U_kj on the trailing upper triangle -- fp64 TENSOR-CORE
//                       tiles (mma.sync.m8n8k4.f64 / DMMA; tcgen05 has no fp64 kind), plus
//                       b_j -= U_kj^T z_k (forward substitution fused as an extra column)
//   chol_backsub_kernel U y = z, one CTA walking the block rows from the bottom
// The trailing update is the n^3/3 term and the only place the tensor pipe is used in this library.
#include <algorithm>

#include "ba.cuh"

namespace {

constexpr int NB = 64;
constexpr int LDS_ = 68;  // shared-memory row stride (doubles): conflict-free DMMA fragment loads

__global__ void __launch_bounds__(256)
chol_diag_kernel(double *S, double *b, int n, int k0, BAState *st)
{
    __shared__ double A[NB][NB + 1];
    __shared__ double z[NB];
    __shared__ int ok;
    if (st->done) return;
    const int tid = threadIdx.x, nb = min(NB, n - k0);
    if (tid == 0) ok = (k0 == 0) ? 1 : st->chol_ok;
    for (int i = tid; i < NB * NB; i += 256) {
        int r = i / NB, c = i % NB;
        A[r][c] = (r < nb && c < nb && c >= r) ? S[(size_t)(k0 + r) * n + k0 + c] : (r == c ? 1.0 : 0.0);
    }
    if (tid < NB) z[tid] = tid < nb ? b[k0 + tid] : 0.0;
    __syncthreads();
    if (!ok) return;
    for (int j = 0; j < nb; j++) {
        if (tid == 0) {
            double d = A[j][j];
            if (!(d > 0) || !isfinite(d)) ok = 0; else A[j][j] = sqrt(d);
        }
        __syncthreads();
        if (!ok) break;
        const double dj = A[j][j];
        for (int c = j + 1 + tid; c < nb; c += 256) A[j][c] /= dj;   // row j of U
        __syncthreads();
        const int m = nb - j - 1;
        for (int t = tid; t < m * m; t += 256) {
            int rr = t / m, cc = t - rr * m;
            if (cc >= rr) A[j + 1 + rr][j + 1 + cc] -= A[j][j + 1 + rr] * A[j][j + 1 + cc];
        }
        __syncthreads();
    }
    if (!ok) { if (tid == 0) st->chol_ok = 0; return; }
    // z_k = U_kk^-T b_k (forward substitution with the lower-triangular U^T)
    for (int j = 0; j < nb; j++) {
        if (tid == 0) z[j] /= A[j][j];
        __syncthreads();
        const double zj = z[j];
        for (int c = j + 1 + tid; c < nb; c += 256) z[c] -= A[j][c] * zj;
        __syncthreads();
    }
    for (int i = tid; i < nb * nb; i += 256) {
        int r = i / nb, c = i - r * nb;
        if (c >= r) S[(size_t)(k0 + r) * n + k0 + c] = A[r][c];
    }
    if (tid < nb) b[k0 + tid] = z[tid];
    if (tid == 0) st->chol_ok = 1;
}

// U_kj = U_kk^-T S_kj : one thread per column of the block row (columns k0+NB .. n-1)
__global__ void __launch_bounds__(128)
chol_panel_kernel(double *S, int n, int k0, int nlim, const BAState *st)
{
    __shared__ double Ukk[NB][NB + 1];
    if (st->done || !st->chol_ok) return;
    const int nb = min(NB, n - k0);
    for (int i = threadIdx.x; i < NB * NB; i += 128) {
        int r = i / NB, c = i % NB;
        Ukk[r][c] = (r < nb && c < nb && c >= r) ? S[(size_t)(k0 + r) * n + k0 + c] : 0.0;
    }
    __syncthreads();
    const int col = k0 + nb + blockIdx.x * 128 + threadIdx.x;
    if (col >= nlim) return;   // columns beyond the envelope of this block row are structurally zero
    double x[NB];
#pragma unroll 8
    for (int r = 0; r < NB; r++) x[r] = r < nb ? S[(size_t)(k0 + r) * n + col] : 0.0;
    // solve U_kk^T x = s  (U_kk^T lower): x_r = (s_r - sum_{t<r} U[t][r] x_t) / U[r][r]
    for (int r = 0; r < nb; r++) {
        double s = x[r];
        for (int t = 0; t < r; t++) s -= Ukk[t][r] * x[t];
        x[r] = s / Ukk[r][r];
    }
    for (int r = 0; r < nb; r++) S[(size_t)(k0 + r) * n + col] = x[r];
}

__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// trailing update: tile (bi, bj), bi <= bj, of the upper triangle right/below block row k.
// S[i0+r][j0+c] -= sum_t P[t][i0+r] * P[t][j0+c],  P = block row k (nb x n).
__global__ void __launch_bounds__(256)
chol_update_kernel(double *S, double *b, int n, int k0, int t0 /* first trailing column */, int nlim, const BAState *st)
{
    constexpr int KH = 32;   // the 64-row block row is staged in two halves (48 KB static smem limit)
    __shared__ double Pi[KH * LDS_];
    __shared__ double Pj[KH * LDS_];
    if (st->done || !st->chol_ok) return;
    // linear tile index -> (bi, bj) with bi <= bj
    const int nt = (nlim - t0 + NB - 1) / NB;   // only tiles inside the envelope [t0, nlim) of block row k
    int bi = 0, rem = blockIdx.x;
    while (rem >= nt - bi) { rem -= nt - bi; bi++; }
    const int bj = bi + rem;
    const int i0 = t0 + bi * NB, j0 = t0 + bj * NB;
    const int nb = min(NB, n - k0);
    const int tid = threadIdx.x;
    // 8 warps: warp (wr, wc) owns rows wr*16..+15 (2 mma tiles), cols wc*32..+31 (4 mma tiles)
    const int warp = tid >> 5, lane = tid & 31;
    const int wr = warp >> 1, wc = warp & 1;
    const int g = lane >> 2, q = lane & 3;   // fragment row group / k index
    double acc[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c][0] = acc[a][c][1] = 0.0;
    double bsum = 0.0;  // fused forward substitution of the right-hand side (diagonal tiles)
    for (int kh = 0; kh < NB; kh += KH) {
        __syncthreads();
        for (int i = tid; i < KH * NB; i += 256) {
            int t = i / NB, c = i % NB;
            Pi[t * LDS_ + c] = (kh + t < nb && i0 + c < n) ? S[(size_t)(k0 + kh + t) * n + i0 + c] : 0.0;
            Pj[t * LDS_ + c] = (kh + t < nb && j0 + c < n) ? S[(size_t)(k0 + kh + t) * n + j0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < KH; kk += 4) {
            double af[2], bf[4];
#pragma unroll
            for (int a = 0; a < 2; a++) af[a] = Pi[(kk + q) * LDS_ + wr * 16 + a * 8 + g];    // A[row g][k q] = P[k][i0+row]
#pragma unroll
            for (int c = 0; c < 4; c++) bf[c] = Pj[(kk + q) * LDS_ + wc * 32 + c * 8 + g];    // B[k q][col g]
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int c = 0; c < 4; c++) dmma_8x8x4(acc[a][c][0], acc[a][c][1], af[a], bf[c]);
        }
        if (bi == bj && tid < NB && i0 + tid < n)
            for (int t = 0; t < KH && kh + t < nb; t++) bsum += Pi[t * LDS_ + tid] * b[k0 + kh + t];
    }
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int r = i0 + wr * 16 + a * 8 + g;
            const int cc = j0 + wc * 32 + c * 8 + 2 * q;
            if (r < n) {
                if (cc < n && cc >= r) S[(size_t)r * n + cc] -= acc[a][c][0];
                if (cc + 1 < n && cc + 1 >= r) S[(size_t)r * n + cc + 1] -= acc[a][c][1];
            }
        }
    // forward substitution of the right-hand side, fused: b_j -= P[:, j]^T z_k  (diagonal tiles only)
    if (bi == bj && tid < NB && i0 + tid < n) b[i0 + tid] -= bsum;
}

// U y = z from the bottom block row upwards; one CTA (1024 threads)
__global__ void __launch_bounds__(1024)
chol_backsub_kernel(const double *S, const double *z, double *y, int n, const int *__restrict__ lim, BAState *st)
{
    __shared__ double part[32][NB + 1];
    __shared__ double yk[NB];
    if (st->done || !st->chol_ok) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = (n + NB - 1) / NB;
    __shared__ int fin;
    if (tid == 0) fin = 1;
    for (int kb = nblk - 1; kb >= 0; kb--) {
        const int k0 = kb * NB, nb = min(NB, n - k0);
        // dot products of the block row with the already-known tail of y: 32 warps x 2 rows each
        for (int r = warp; r < nb; r += 32) {
            double s = 0;
            const int cend = lim[kb];   // U_kj == 0 beyond the envelope
            for (int c = k0 + nb + lane; c < cend; c += 32) s += S[(size_t)(k0 + r) * n + c] * y[c];
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) yk[r] = z[k0 + r] - s;
        }
        __syncthreads();
        // small triangular solve U_kk y_k = rhs, serial in one warp
        if (warp == 0) {
            for (int r = nb - 1; r >= 0; r--) {
                double s = 0;
                for (int c = r + 1 + lane; c < nb; c += 32) s += S[(size_t)(k0 + r) * n + k0 + c] * yk[c];
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) {
                    double v = (yk[r] - s) / S[(size_t)(k0 + r) * n + k0 + r];
                    yk[r] = v;
                    if (!isfinite(v)) fin = 0;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (tid < nb) y[k0 + tid] = yk[tid];
        __threadfence_block();
        __syncthreads();
    }
    (void)part;
    if (tid == 0 && !fin) st->chol_ok = 0;
}

__global__ void chol_gradient_check_kernel(BAState *st)
{
    // gradient tolerance is checked once the new linearisation is complete (Ceres: before the next iteration)
    if (!st->done && st->need_linearize && !(st->gmax > 1e-10)) { st->done = 1; st->termination = 3; }
}

}  // namespace

// lim_host[kb] / D.chol_lim[kb]: end column (exclusive, <= n) of the envelope of block row kb -- cumulative
// maximum of the camera co-visibility reach, so fill-in stays inside it.  A banded reduced camera system
// (BASELINE config 5: every point seen by 5 of the <= 40 nearest poses) is factorised in O(n * band^2).
int pmv_internal_ba_cholesky_large(pmv_ctx *ctx, const BADev &D, const int *lim_host, cudaStream_t s)
{
    const int n = D.n;
    for (int w = 0; w < D.W; w++) {
        double *S = D.S + (size_t)w * n * n, *b = D.rhs + (size_t)w * n, *y = D.yc + (size_t)w * n;
        BAState *st = D.st + w;
        chol_gradient_check_kernel<<<1, 1, 0, s>>>(st);
        PMV_LAUNCH_CHECK(ctx, "chol_gradient_check_kernel");
        for (int k0 = 0, kb = 0; k0 < n; k0 += NB, kb++) {
            chol_diag_kernel<<<1, 256, 0, s>>>(S, b, n, k0, st);
            PMV_LAUNCH_CHECK(ctx, "chol_diag_kernel");
            const int t0 = k0 + NB;
            const int nlim = lim_host ? std::min(n, lim_host[kb]) : n;
            if (t0 < nlim) {
                chol_panel_kernel<<<(nlim - t0 + 127) / 128, 128, 0, s>>>(S, n, k0, nlim, st);
                PMV_LAUNCH_CHECK(ctx, "chol_panel_kernel");
                const int nt = (nlim - t0 + NB - 1) / NB;
                chol_update_kernel<<<nt * (nt + 1) / 2, 256, 0, s>>>(S, b, n, k0, t0, nlim, st);
                PMV_LAUNCH_CHECK(ctx, "chol_update_kernel");
            }
        }
        chol_backsub_kernel<<<1, 1024, 0, s>>>(S, b, y, n, D.chol_lim, st);
        PMV_LAUNCH_CHECK(ctx, "chol_backsub_kernel");
    }
    return PMV_OK;
}

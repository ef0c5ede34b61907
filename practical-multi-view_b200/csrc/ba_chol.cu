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
#include <cstdlib>

#include "ba.cuh"

namespace {

constexpr int NB = PMV_CHOL_NB;   // 32 rows per block step: panel columns keep their 32 unknowns in registers
constexpr int TW = 64;            // trailing-update tile width
constexpr int LDS_ = 68;          // shared-memory row stride (doubles): conflict-free DMMA fragment loads

// Factor the diagonal block S_kk = U_kk^T U_kk and solve z_k = U_kk^-T b_k -- warp-synchronous: lane j of
// warp 0 holds column j in registers, pivots and row entries travel by shuffles (no barriers, ~2 us).
// With fuse_nlim > 0 (banded system) the whole CTA then solves the panel U_kj = U_kk^-T S_kj.
__global__ void __launch_bounds__(256)
chol_diag_kernel(double *S, double *b, int n, int k0, int fuse_nlim /* > 0: also solve the panel columns [k0+NB, fuse_nlim) */, BAState *st)
{
    __shared__ double U[NB][NB + 1];
    __shared__ double invd[NB];
    __shared__ int ok;
    if (st->done) return;
    const int tid = threadIdx.x, nb = min(NB, n - k0);
    if (tid == 0) ok = (k0 == 0) ? 1 : st->chol_ok;
    __syncthreads();
    if (!ok) return;
    if (tid < 32) {
        const int j = tid;
        double col[NB];
#pragma unroll
        for (int i = 0; i < NB; i++)
            col[i] = (i < nb && j < nb && i <= j) ? S[(size_t)(k0 + i) * n + k0 + j] : (i == j ? 1.0 : 0.0);
        double bj = j < nb ? b[k0 + j] : 0.0;
        bool good = true;
#pragma unroll
        for (int k = 0; k < NB; k++) {
            const double piv = __shfl_sync(0xffffffffu, col[k], k);
            if (!(piv > 0) || !isfinite(piv)) good = false;      // warp-uniform
            const double inv = 1.0 / sqrt(good ? piv : 1.0);
            const double ukj = col[k] * inv;                      // U[k][j] (meaningful for j >= k)
            col[k] = ukj;
            if (j == k) invd[k] = inv;
#pragma unroll
            for (int i = k + 1; i < NB; i++) {
                const double uki = __shfl_sync(0xffffffffu, ukj, i);
                if (i <= j) col[i] -= uki * ukj;
            }
            // forward substitution of the right-hand side rides along: z_k = b_k / U_kk, b_j -= U_kj z_k
            const double zk = __shfl_sync(0xffffffffu, bj, k) * inv;
            if (j == k) bj = zk; else if (j > k) bj -= ukj * zk;
        }
        if (!good) { if (tid == 0) { ok = 0; st->chol_ok = 0; } }
        else {
#pragma unroll
            for (int i = 0; i < NB; i++) {
                U[i][j] = (i <= j) ? col[i] : 0.0;
                if (i < nb && j < nb && i <= j) S[(size_t)(k0 + i) * n + k0 + j] = col[i];
            }
            if (j < nb) b[k0 + j] = bj;
            if (tid == 0) st->chol_ok = 1;
        }
    }
    __syncthreads();
    if (!ok) return;
    // panel: one column per thread, its NB unknowns in registers, reciprocal diagonal from the factorisation
    for (int colx = k0 + nb + tid; colx < fuse_nlim; colx += 256) {
        double x[NB];
#pragma unroll
        for (int r = 0; r < NB; r++) x[r] = r < nb ? S[(size_t)(k0 + r) * n + colx] : 0.0;
#pragma unroll
        for (int r = 0; r < NB; r++) {
            double sacc = x[r];
#pragma unroll
            for (int t = 0; t < r; t++) sacc -= U[t][r] * x[t];
            x[r] = sacc * invd[r];
        }
#pragma unroll
        for (int r = 0; r < NB; r++)
            if (r < nb) S[(size_t)(k0 + r) * n + colx] = x[r];
    }
}

// U_kj = U_kk^-T S_kj : one thread per column of the block row (columns k0+NB .. nlim-1)
__global__ void __launch_bounds__(128)
chol_panel_kernel(double *S, int n, int k0, int nlim, const BAState *st)
{
    __shared__ double Ukk[NB][NB + 1];
    if (st->done || !st->chol_ok) return;
    const int nb = min(NB, n - k0);
    for (int i = threadIdx.x; i < NB * NB; i += 128) {
        int r = i / NB, c = i % NB;
        Ukk[r][c] = (r < nb && c < nb && c >= r) ? S[(size_t)(k0 + r) * n + k0 + c] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    const int col = k0 + nb + blockIdx.x * 128 + threadIdx.x;
    if (col >= nlim) return;   // columns beyond the envelope of this block row are structurally zero
    double x[NB];
#pragma unroll
    for (int r = 0; r < NB; r++) x[r] = r < nb ? S[(size_t)(k0 + r) * n + col] : 0.0;
    // solve U_kk^T x = s  (U_kk^T lower): x_r = (s_r - sum_{t<r} U[t][r] x_t) / U[r][r]
#pragma unroll
    for (int r = 0; r < NB; r++) {
        double sacc = x[r];
#pragma unroll
        for (int t = 0; t < r; t++) sacc -= Ukk[t][r] * x[t];
        x[r] = sacc / Ukk[r][r];
    }
#pragma unroll
    for (int r = 0; r < NB; r++)
        if (r < nb) S[(size_t)(k0 + r) * n + col] = x[r];
}

__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// trailing update: 64x64 tile (bi, bj), bi <= bj, of the upper triangle right/below block row k.
// S[i0+r][j0+c] -= sum_t P[t][i0+r] * P[t][j0+c],  P = block row k (nb <= 32 rows).  fp64 tensor-core tiles.
__global__ void __launch_bounds__(256)
chol_update_kernel(double *S, double *b, int n, int k0, int t0 /* first trailing column */, int nlim, const BAState *st)
{
    __shared__ double Pi[NB * LDS_];
    __shared__ double Pj[NB * LDS_];
    if (st->done || !st->chol_ok) return;
    // linear tile index -> (bi, bj) with bi <= bj
    const int nt = (nlim - t0 + TW - 1) / TW;   // only tiles inside the envelope [t0, nlim) of block row k
    int bi = 0, rem = blockIdx.x;
    while (rem >= nt - bi) { rem -= nt - bi; bi++; }
    const int bj = bi + rem;
    const int i0 = t0 + bi * TW, j0 = t0 + bj * TW;
    const int nb = min(NB, n - k0);
    const int tid = threadIdx.x;
    for (int i = tid; i < NB * TW; i += 256) {
        int t = i / TW, c = i % TW;
        Pi[t * LDS_ + c] = (t < nb && i0 + c < n) ? S[(size_t)(k0 + t) * n + i0 + c] : 0.0;
        Pj[t * LDS_ + c] = (t < nb && j0 + c < n) ? S[(size_t)(k0 + t) * n + j0 + c] : 0.0;
    }
    __syncthreads();
    // 8 warps: warp (wr, wc) owns rows wr*16..+15 (2 mma tiles), cols wc*32..+31 (4 mma tiles)
    const int warp = tid >> 5, lane = tid & 31;
    const int wr = warp >> 1, wc = warp & 1;
    const int g = lane >> 2, q = lane & 3;   // fragment row group / k index
    double acc[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c][0] = acc[a][c][1] = 0.0;
#pragma unroll
    for (int kk = 0; kk < NB; kk += 4) {
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
    if (bi == bj && tid < TW && i0 + tid < n) {
        double bsum = 0.0;
        for (int t = 0; t < nb; t++) bsum += Pi[t * LDS_ + tid] * b[k0 + t];
        b[i0 + tid] -= bsum;
    }
}

// U y = z from the bottom block row upwards; one CTA (512 threads).  Per block row: 16 warps form the dot
// products with the known tail of y (inside the envelope), then one warp solves the 32x32 triangle.
// For narrow envelopes (banded systems) the NB x (envelope) block row is copied into shared memory with
// cp.async one step AHEAD (it does not depend on y), so the 188 dependent steps see no HBM/L2 latency.
constexpr int BS_MAXW = 321;   // widest envelope (columns from k0) staged in shared memory; odd stride: no bank conflicts down a column

constexpr int BS_THREADS = 512;

__global__ void __launch_bounds__(BS_THREADS)
chol_backsub_kernel(const double *S, const double *z, double *y, int n, size_t ld, const int *__restrict__ lim, BAState *st, int staged)
{
    extern __shared__ double sbuf[];            // staged: 2 x NB x BS_MAXW
    __shared__ double yk[NB];
    __shared__ int fin;
    if (st->done || !st->chol_ok) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = (n + NB - 1) / NB;
    if (tid == 0) fin = 1;
    double *zs_base = staged ? sbuf + 2 * NB * BS_MAXW + 512 : nullptr;   // [2][NB] staged right-hand sides
    auto stage = [&](int kb, double *dst) {
        const int k0 = kb * NB, nb = min(NB, n - k0), wdt = min(n, lim[kb]) - k0;
        // threads 32..511 (warps 1..15), 16 per row, 30 rows per pass; no index arithmetic beyond an add
        const int t = tid - 32;
        for (int r = t >> 4; r < nb; r += (BS_THREADS - 32) / 16) {
            const double *src = S + (size_t)(k0 + r) * ld + k0;
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + r * BS_MAXW);
            for (int c = r + (t & 15); c < wdt; c += 16)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa + 8u * c), "l"(src + c));
        }
        if (zs_base && t < nb) {   // the right-hand side of the block row rides along (no L2 round trip per step)
            const unsigned za = (unsigned)__cvta_generic_to_shared(zs_base + (kb & 1) * NB + t);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(za), "l"(z + k0 + t));
        }
        asm volatile("cp.async.commit_group;");
    };
    if (staged) {
        // Banded system.  Block rows are staged two steps ahead by warps 1..15, the known tail of y lives in a
        // shared-memory ring, and the dot products of a block row are split: the part against unknowns that are
        // two or more blocks old is formed by warps 1..15 WHILE warp 0 solves the triangle of the block above
        // (branch-free, row in registers, reciprocal diagonal, ~46 cycles per unknown); only the 32 columns of
        // the block just solved remain on the critical path.
        double *ysm = sbuf + 2 * NB * BS_MAXW;   // ring of the last 512 unknowns, indexed by column & 511
        double *sP = zs_base + 2 * NB;           // [NB] old-part sums of the block row about to be solved
        auto ubuf = [&](int kb) { return sbuf + (kb & 1 ? NB * BS_MAXW : 0); };
        if (tid < NB) sP[tid] = 0.0;             // the top block row has no columns beyond its triangle
        if (warp > 0) {
            stage(nblk - 1, ubuf(nblk - 1));
            if (nblk > 1) stage(nblk - 2, ubuf(nblk - 2)); else asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 1;");
        }
        __syncthreads();   // block row nblk - 1 landed
        for (int kb = nblk - 1; kb >= 0; kb--) {
            const int k0 = kb * NB, nb = min(NB, n - k0);
            const int wdt = min(n, lim[kb]) - k0;   // U_kj == 0 beyond the envelope
            const double *U = ubuf(kb);
            // warp 0 takes its triangle rows into registers BEFORE barrier A: afterwards this buffer is refilled
            double urow[NB], myinv = 0.0;
            if (warp == 0) {
#pragma unroll
                for (int r = 0; r < NB; r++) urow[r] = (lane < nb && r < nb && r > lane) ? U[lane * BS_MAXW + r] : 0.0;
                myinv = lane < nb ? 1.0 / U[lane * BS_MAXW + lane] : 0.0;
            }
            // (1) the 32 columns of the block solved in the previous step, one per lane
            for (int row = warp; row < nb; row += BS_THREADS / 32) {
                const int c = NB + lane;
                double sacc = (c < wdt) ? U[row * BS_MAXW + c] * ysm[(k0 + c) & 511] : 0.0;
                for (int o = 16; o; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
                if (lane == 0) yk[row] = zs_base[(kb & 1) * NB + row] - sP[row] - sacc;
            }
            __syncthreads();   // A: right-hand side of the triangle complete; block row kb no longer needed
            if (warp == 0) {
                // lane l keeps unknown l; rows >= nb of a partial last block are inert
                double mine = lane < nb ? yk[lane] : 0.0;
#pragma unroll
                for (int r = NB - 1; r >= 0; r--) {
                    const double v = __shfl_sync(0xffffffffu, mine * myinv, r);
                    const double upd = fma(-urow[r], v, mine);
                    mine = (lane == r) ? v : ((lane < r) ? upd : mine);
                }
                if (!isfinite(mine)) fin = 0;
                if (lane < nb) { y[k0 + lane] = mine; ysm[(k0 + lane) & 511] = mine; }
            } else {
                // (2) refill the buffer of block row kb with block row kb - 2, make block row kb - 1 visible, and
                //     form its sums against the unknowns of blocks >= kb + 1 (all known)
                if (kb >= 2) stage(kb - 2, ubuf(kb - 2)); else asm volatile("cp.async.commit_group;");
                asm volatile("cp.async.wait_group 1;");
                asm volatile("bar.sync 1, %0;" ::"n"(BS_THREADS - 32) : "memory");
                if (kb >= 1) {
                    const int k1 = k0 - NB, w1 = min(n, lim[kb - 1]) - k1;
                    const double *U1 = ubuf(kb - 1);
                    for (int row = warp - 1; row < NB; row += BS_THREADS / 32 - 1) {
                        double sacc = 0;
                        for (int c = 2 * NB + lane; c < w1; c += 32) sacc += U1[row * BS_MAXW + c] * ysm[(k1 + c) & 511];
                        for (int o = 16; o; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
                        if (lane == 0) sP[row] = sacc;
                    }
                }
            }
            __syncthreads();   // B: unknowns of block kb published, sums of block row kb - 1 ready
        }
        if (tid == 0 && !fin) st->chol_ok = 0;
        return;
    }
    for (int kb = nblk - 1; kb >= 0; kb--) {
        const int k0 = kb * NB, nb = min(NB, n - k0);
        const int cend = min(n, lim[kb]);   // U_kj == 0 beyond the envelope
        if (warp < nb) {
            double sacc = 0;
            for (int c = k0 + nb + lane; c < cend; c += 32) sacc += S[(size_t)(k0 + warp) * ld + c] * y[c];
            for (int o = 16; o; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (lane == 0) yk[warp] = z[k0 + warp] - sacc;
        }
        __syncthreads();
        if (warp == 0) {
            // lane r keeps row r's unknown; column r of the triangle is read as needed
            double mine = lane < nb ? yk[lane] : 0.0;
            for (int r = nb - 1; r >= 0; r--) {
                const double urr = S[(size_t)(k0 + r) * ld + k0 + r];
                const double v = __shfl_sync(0xffffffffu, mine, r) / urr;
                if (lane == r) { mine = v; if (!isfinite(v)) fin = 0; }
                else if (lane < r) mine -= S[(size_t)(k0 + lane) * ld + k0 + r] * v;
            }
            if (lane < nb) y[k0 + lane] = mine;
        }
        __syncthreads();
    }
    if (tid == 0 && !fin) st->chol_ok = 0;
}

__global__ void chol_gradient_check_kernel(BAState *st)
{
    // gradient tolerance is checked once the new linearisation is complete (Ceres: before the next iteration)
    if (!st->done && st->need_linearize && !(st->gmax > 1e-10)) { st->done = 1; st->termination = 3; }
}

// ---- split solve (BASplit): small kernels around the two concurrent half factorisations -----------------------
// 1. private status blocks, saved separator block / rhs, and the index-reversed copy of the trailing system.
//    S2[i'][j'] = S[n-1-j'][n-1-i'] inside the reversed envelope (whole 32-column tiles, zero where S has no entry).
__global__ void __launch_bounds__(256)
split_prepare_kernel(const double *__restrict__ S, const double *__restrict__ b, int n, const int *__restrict__ lim, BASplit P, const BAState *st)
{
    if (st->done) return;
    const int a = P.a, w = P.w, h2 = P.h2;
    const int row = blockIdx.x;
    if (row < h2) {
        const int ip = row, r_or = n - 1 - ip;                         // reversed row i' <-> original column index r_or
        int cend = min(h2, P.lim2[ip / PMV_CHOL_NB]);
        cend = min(h2, (cend + PMV_CHOL_NB - 1) / PMV_CHOL_NB * PMV_CHOL_NB);
        for (int jp = ip + threadIdx.x; jp < cend; jp += 256) {
            const int c_or = n - 1 - jp;                               // original row (c_or <= r_or)
            const bool inside = r_or < min(n, lim[c_or / PMV_CHOL_NB]);
            P.S2[(size_t)ip * h2 + jp] = inside ? S[(size_t)c_or * n + r_or] : 0.0;
        }
        if (threadIdx.x == 0) P.b2[ip] = b[r_or];
    } else if (row < h2 + w) {
        const int i = row - h2;
        for (int j = threadIdx.x; j < w; j += 256) P.AM[(size_t)i * w + j] = j >= i ? S[(size_t)(a + i) * n + a + j] : 0.0;
        if (threadIdx.x == 0) P.bM[i] = b[a + i];
    } else if (threadIdx.x < 3) {
        BAState t = *st;
        t.chol_ok = 0;
        P.st3[threadIdx.x] = t;
    }
}

// 2. separator: A3 = U1^T U1 + flip(U2^T U2) - A_M and b3 = U1^T z1 + flip(U2^T z2) - b_M, where U1 / U2 are the
//    trailing w x w blocks of the two partial factors (each equals A_M minus that side's Schur update, factorised).
__global__ void __launch_bounds__(256)
split_middle_kernel(const double *__restrict__ S, const double *__restrict__ z1, int n, BASplit P)
{
    if (!P.st3[0].chol_ok || !P.st3[1].chol_ok || P.st3[0].done) return;
    __shared__ double colU1[512], colU2[512];
    __shared__ double red[8];
    const int i = blockIdx.x, w = P.w, a = P.a, h2 = P.h2, o2 = h2 - w;
    const int ip = w - 1 - i;
    for (int k = threadIdx.x; k < w; k += 256) {
        colU1[k] = k <= i ? S[(size_t)(a + k) * n + a + i] : 0.0;                  // column i of U1
        colU2[k] = k <= ip ? P.S2[(size_t)(o2 + k) * h2 + o2 + ip] : 0.0;           // column i' of U2
    }
    __syncthreads();
    for (int j = threadIdx.x; j < w; j += 256) {
        double v = 0.0;
        if (j >= i) {
            const int jp = w - 1 - j;                                               // jp <= ip
            double s1 = 0.0, s2 = 0.0;
            for (int k = 0; k <= i; k++) s1 = fma(colU1[k], S[(size_t)(a + k) * n + a + j], s1);
            for (int k = 0; k <= jp; k++) s2 = fma(P.S2[(size_t)(o2 + k) * h2 + o2 + jp], colU2[k], s2);
            v = (s1 + s2) - P.AM[(size_t)i * w + j];
        }
        P.S3[(size_t)i * w + j] = v;
    }
    double t = 0.0;
    for (int k = threadIdx.x; k < w; k += 256) t += colU1[k] * z1[a + k] + colU2[k] * P.b2[o2 + k];
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int q = 0; q < 8; q++) v += red[q];
        P.b3[i] = v - P.bM[i];
    }
}

// 3. with the separator solved (y3): the separator rows of the two forward-substituted right-hand sides become
//    U y3, so the ordinary back-substitutions of the two half systems continue upwards from it.
__global__ void __launch_bounds__(256)
split_fix_kernel(const double *__restrict__ S, double *__restrict__ z1, int n, BASplit P)
{
    if (!P.st3[2].chol_ok || P.st3[2].done) return;
    __shared__ double red[8];
    const int w = P.w, a = P.a, h2 = P.h2, o2 = h2 - w;
    const int i = blockIdx.x % w, side = blockIdx.x / w;
    double t = 0.0;
    if (side == 0) { for (int j = i + threadIdx.x; j < w; j += 256) t += S[(size_t)(a + i) * n + a + j] * P.y3[j]; }
    else { for (int j = i + threadIdx.x; j < w; j += 256) t += P.S2[(size_t)(o2 + i) * h2 + o2 + j] * P.y3[w - 1 - j]; }
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int q = 0; q < 8; q++) v += red[q];
        if (side == 0) z1[a + i] = v; else P.b2[o2 + i] = v;
    }
}

// 4. solution of the bottom part back into original order; overall status
__global__ void __launch_bounds__(256) split_finish_kernel(double *__restrict__ y, int n, BASplit P, BAState *st)
{
    if (st->done) return;
    const int nb = P.h2 - P.w;
    for (int ip = blockIdx.x * 256 + threadIdx.x; ip < nb; ip += gridDim.x * 256) y[n - 1 - ip] = P.y2[ip];
    if (blockIdx.x == 0 && threadIdx.x == 0) st->chol_ok = P.st3[0].chol_ok && P.st3[1].chol_ok && P.st3[2].chol_ok;
}

// ---- part solve (BAPart): kernels around the P concurrent segment factorisations ----------------------------------
// 0. private status blocks
__global__ void part_prepare_kernel(BAPart Q, const BAState *st)
{
    if (threadIdx.x < 2 * Q.P - 1) {
        BAState t = *st;
        t.chol_ok = 0;
        Q.stp[threadIdx.x] = t;
    }
}

// 1. spike of segment i >= 1: G = U_X^-T B, forward substitution of w right-hand sides through the factor of the
//    extended system X_i (in place in S at (a_i, a_i)).  B[r][c] = S[mprev + c][a_i + r] inside the envelope of the
//    separator rows mprev + c (non-zero only in the first < w rows).  One CTA per 32 right-hand sides: the current and the
//    next PS_T block rows of the right-hand side live in a shared-memory ring; per block row warp 0 solves the 32 x 32
//    triangle (lane = column of the right-hand side, the column in registers) while warp j stages the factor's tile
//    U_{k,k+j}, then warp j applies U_{k,k+j}^T Y to the j-th tile below.  The NEXT diagonal tile arrives by cp.async.
constexpr int PS_T = 10;                 // off-diagonal tiles per block row of the factor (envelope up to 11 tiles of 32 columns)
constexpr int PS_WARPS = PS_T + 1;       // warp 0: triangle; warp j = 1 .. PS_T: the tile j below
constexpr size_t PS_SMEM = sizeof(double) * ((size_t)(2 + PS_T) * NB * NB + (size_t)(PS_T + 2) * NB * 16);   // U_kk x 2 | U_{k,k+j} | ring | Y

__device__ __forceinline__ void ps_cp_async_16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

struct SpikeArgs {
    const double *Sx;        // factor (upper triangular rows inside the envelope), element (r, c) at Sx[r * ld + c]
    size_t ld;
    int nx;                  // order of the system
    const int *lim;          // its envelope per 32-row block, relative (device)
    int w;                   // right-hand sides
    // right-hand sides: mode 0 -- B[row][col] = Bsrc[col * bld + row] where row < blim[(brow0 + col) / 32] - bcol0 (rows of a
    // banded matrix read as columns); mode 1 -- B[row][col] = Bsrc[row * bld + col], row < nx
    int bmode;
    const double *Bsrc;
    size_t bld;
    const int *blim;
    int brow0, bcol0;
    double *G;               // out: nx x w, row major
    const BAState *st;       // status of the factorisation the factor comes from
};

// One CTA per PS_CC = 16 right-hand sides; lane = (column c = lane & 15, half h = lane >> 4): a lane owns rows
// h, h + 2, h + 4, ... of its column (interleaved, so both halves stay busy down the triangle).
constexpr int PS_CC = 16;

__global__ void __launch_bounds__(PS_WARPS * 32)
part_spike_kernel(const SpikeArgs A)
{
    const BAState *st = A.st;
    if (st->done || !st->chol_ok) return;
    extern __shared__ __align__(16) double psm[];
    double *Ukk = psm;                                   // [2][NB][NB]  diagonal tile of the current / next block row
    double *Uoff = psm + 2 * NB * NB;                    // [PS_T][NB][NB] tile j of the current block row, staged by warp j
    double *Bring = Uoff + PS_T * NB * NB;               // [PS_T + 1][NB][PS_CC] right-hand-side rows kb*32 .., slot = block row % (PS_T + 1)
    double *Ys = Bring + (PS_T + 1) * NB * PS_CC;        // [NB][PS_CC]
    __shared__ double rdiag[NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = lane & 15, h = lane >> 4;
    const int nx = A.nx, w = A.w;
    const int *lim = A.lim;
    const size_t ld = A.ld;
    const int nblk = (nx + NB - 1) / NB;
    const int c0 = blockIdx.x * PS_CC;                      // first right-hand-side column of this CTA
    double *G = A.G;
    const double *Sx = A.Sx;
    const bool even = (ld & 1) == 0 && ((size_t)Sx & 15) == 0;   // tile rows start on 16-byte boundaries
    auto tiles_of = [&](int kb) { return min(PS_T + 1, (min(nx, lim[kb]) - kb * NB + NB - 1) / NB); };   // diagonal tile included
    // one tile of the factor, rows k0 .. k0 + 32, columns cb .. cb + 32 of X, into dst[NB][NB] by ONE warp (zero beyond X)
    auto load_tile = [&](double *dst, int k0, int cb, bool async) {
        for (int i = lane; i < NB * (NB / 2); i += 32) {
            const int r = i / (NB / 2), p2 = i - r * (NB / 2), row = k0 + r, col = cb + 2 * p2;
            double *d = dst + r * NB + 2 * p2;
            if (row < nx && col + 1 < nx && even) {
                const double *g = Sx + (size_t)row * ld + col;
                if (async) ps_cp_async_16(d, g); else *reinterpret_cast<double2 *>(d) = *reinterpret_cast<const double2 *>(g);
            } else {
                d[0] = (row < nx && col < nx) ? Sx[(size_t)row * ld + col] : 0.0;
                d[1] = (row < nx && col + 1 < nx) ? Sx[(size_t)row * ld + col + 1] : 0.0;
            }
        }
    };
    // right-hand-side tile jb enters the ring: rows jb*32 .. of B (zero beyond the reach of the separator rows)
    auto enter = [&](int jb) {
        double *dst = Bring + (size_t)(jb % (PS_T + 1)) * NB * PS_CC;
        for (int i = tid; i < NB * PS_CC; i += PS_WARPS * 32) {
            int r, cc;
            if (A.bmode == 0) { cc = i / NB; r = i - cc * NB; } else { r = i / PS_CC; cc = i - r * PS_CC; }   // consecutive threads along the source rows
            const int row = jb * NB + r, col = c0 + cc;
            double v = 0.0;
            if (row < nx && col < w) {
                if (A.bmode == 0) { if (row < A.blim[(A.brow0 + col) / NB] - A.bcol0) v = A.Bsrc[(size_t)col * A.bld + row]; }
                else v = A.Bsrc[(size_t)row * A.bld + col];
            }
            dst[r * PS_CC + cc] = v;
        }
    };
    for (int jb = 0; jb < min(nblk, PS_T + 1); jb++) enter(jb);
    if (warp == 0) load_tile(Ukk, 0, 0, false);
    __syncthreads();
    for (int kb = 0; kb < nblk; kb++) {
        const int k0 = kb * NB, rows = min(NB, nx - k0), nt = tiles_of(kb);
        const double *Uk = Ukk + (size_t)(kb & 1) * NB * NB;
        double *Bk = Bring + (size_t)(kb % (PS_T + 1)) * NB * PS_CC;
        if (warp == 0) {
            if (kb + 1 < nblk) load_tile(Ukk + (size_t)((kb + 1) & 1) * NB * NB, k0 + NB, k0 + NB, true);   // next diagonal tile, asynchronously
            asm volatile("cp.async.commit_group;" ::: "memory");
            rdiag[lane] = lane < rows ? 1.0 / Uk[lane * NB + lane] : 0.0;
            __syncwarp();
            // U_kk^T Y = B_k: lane (c, h) keeps rows h, h + 2, ... of column c; the row being solved is broadcast by its owner
            double col[NB / 2];
#pragma unroll
            for (int i = 0; i < NB / 2; i++) col[i] = Bk[(2 * i + h) * PS_CC + c];
#pragma unroll
            for (int r = 0; r < NB; r++) {
                const double mine = col[r >> 1] * rdiag[r];
                const double y = __shfl_sync(0xffffffffu, mine, c + 16 * (r & 1));
                if (h == (r & 1)) col[r >> 1] = y;
#pragma unroll
                for (int i = (r >> 1); i < NB / 2; i++) {
                    const int q = 2 * i + h;               // rows below r only (q > r); the compare folds away for most i
                    if (q > r) col[i] = fma(-Uk[r * NB + q], y, col[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < NB / 2; i++) {
                const int r = 2 * i + h;
                Ys[r * PS_CC + c] = col[i];
                if (r < rows && c0 + c < w) G[(size_t)(k0 + r) * w + c0 + c] = col[i];
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else if (warp < nt) {
            load_tile(Uoff + (size_t)(warp - 1) * NB * NB, k0, k0 + warp * NB, false);   // overlaps with warp 0's triangle
        }
        __syncthreads();
        if (warp >= 1 && warp < nt) {
            // B_{k+j} -= U_{k,k+j}^T Y, j = warp: lane (c, h) updates rows 16 h .. 16 h + 15 of column c
            const double *Uj = Uoff + (size_t)(warp - 1) * NB * NB + 16 * h;
            double *Bj = Bring + (size_t)((kb + warp) % (PS_T + 1)) * NB * PS_CC + (16 * h) * PS_CC + c;
            double acc[16];
#pragma unroll
            for (int r = 0; r < 16; r++) acc[r] = Bj[r * PS_CC];
#pragma unroll 4
            for (int k = 0; k < NB; k++) {
                const double y = Ys[k * PS_CC + c];
#pragma unroll
                for (int r = 0; r < 16; r += 2) {
                    const double2 u = *reinterpret_cast<const double2 *>(Uj + k * NB + r);
                    acc[r] = fma(-u.x, y, acc[r]); acc[r + 1] = fma(-u.y, y, acc[r + 1]);
                }
            }
#pragma unroll
            for (int r = 0; r < 16; r++) Bj[r * PS_CC] = acc[r];
        }
        __syncthreads();
        if (kb + PS_T + 1 < nblk) enter(kb + PS_T + 1);   // the slot of block row kb is free again (read next after a barrier)
    }
}

// 2. separator system (block tridiagonal, upper part): one CTA per 32 x 32 tile, no atomics.
//    blockIdx.y = separator j; blockIdx.x enumerates: D tiles (ti <= tj), E tiles (all, only j < P - 2), then one CTA for r_j.
//    D_j goes to Dsep + j w^2, E_j to E + j w^2 (both w x w, ld = w), r_j to bR + j w.
__global__ void __launch_bounds__(256)
part_reduce_kernel(const double *__restrict__ S, const double *__restrict__ z, size_t ld, BAPart Q)
{
    if (Q.stp[0].done) return;
    for (int i = 0; i < Q.P; i++) if (!Q.stp[i].chol_ok) return;
    __shared__ double As[NB][NB + 1], Bs[NB][NB + 1];
    const int w = Q.w, wt = w / NB, j = blockIdx.y;
    const int nD = wt * (wt + 1) / 2, nE = wt * wt;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;       // thread computes rows ty, ty+8, ty+16, ty+24 of column tx
    const int mj = Q.a[j] + Q.ns[j];                                 // first row of separator j
    const double *Gn = Q.G[j + 1];                                   // spike of the segment right of separator j
    const int nsn = Q.ns[j + 1];
    int id = blockIdx.x;
    if (id < nD) {
        // D_j tile (ti, tj): U_M^T U_M - G_s^T G_s
        int ti = 0;
        while (id >= wt - ti) { id -= wt - ti; ti++; }
        const int tj = ti + id;
        double acc[4] = {0, 0, 0, 0};
        // U_M(j): rows k <= column; k runs over block rows 0 .. tj (entries above the diagonal block of column tile ti vanish for k > ti*32+31)
        for (int kb = 0; kb <= ti; kb++) {
            for (int q = ty; q < NB; q += 8) {
                const int k = kb * NB + q;
                As[q][tx] = (k <= ti * NB + tx) ? S[(size_t)(mj + k) * ld + mj + ti * NB + tx] : 0.0;
                Bs[q][tx] = (k <= tj * NB + tx) ? S[(size_t)(mj + k) * ld + mj + tj * NB + tx] : 0.0;
            }
            __syncthreads();
#pragma unroll 8
            for (int q = 0; q < NB; q++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[e] = fma(As[q][ty + 8 * e], Bs[q][tx], acc[e]);
            __syncthreads();
        }
        double sub[4] = {0, 0, 0, 0};
        for (int r0 = 0; r0 < nsn; r0 += NB) {
            for (int q = ty; q < NB; q += 8) {
                const int r = r0 + q;
                As[q][tx] = r < nsn ? Gn[(size_t)r * w + ti * NB + tx] : 0.0;
                Bs[q][tx] = r < nsn ? Gn[(size_t)r * w + tj * NB + tx] : 0.0;
            }
            __syncthreads();
#pragma unroll 8
            for (int q = 0; q < NB; q++)
#pragma unroll
                for (int e = 0; e < 4; e++) sub[e] = fma(As[q][ty + 8 * e], Bs[q][tx], sub[e]);
            __syncthreads();
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int i = ti * NB + ty + 8 * e, c = tj * NB + tx;
            Q.Dsep[(size_t)j * w * w + (size_t)i * w + c] = c >= i ? acc[e] - sub[e] : 0.0;
        }
    } else if (id < nD + nE) {
        if (j + 2 >= Q.P) return;                                     // the last separator has no neighbour on its right
        // E_j tile (ti, tj) = H^T U_M(j+1): H = rows ns .. ns + w of G(j+1), U_M(j+1) upper triangular
        id -= nD;
        const int ti = id / wt, tj = id - ti * wt;
        const int mn = Q.a[j + 1] + Q.ns[j + 1];                      // first row of separator j + 1
        double acc[4] = {0, 0, 0, 0};
        for (int kb = 0; kb <= tj; kb++) {
            for (int q = ty; q < NB; q += 8) {
                const int k = kb * NB + q;
                As[q][tx] = Gn[(size_t)(nsn + k) * w + ti * NB + tx];
                Bs[q][tx] = (k <= tj * NB + tx) ? S[(size_t)(mn + k) * ld + mn + tj * NB + tx] : 0.0;
            }
            __syncthreads();
#pragma unroll 8
            for (int q = 0; q < NB; q++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[e] = fma(As[q][ty + 8 * e], Bs[q][tx], acc[e]);
            __syncthreads();
        }
#pragma unroll
        for (int e = 0; e < 4; e++)
            Q.E[(size_t)j * w * w + (size_t)(ti * NB + ty + 8 * e) * w + tj * NB + tx] = acc[e];
    } else {
        // r_j = U_M(j)^T z_M(j) - G_s(j+1)^T z_s(j+1): thread i of the first w
        const int an = Q.a[j + 1];
        for (int i = tid; i < w; i += 256) {
            double v = 0.0;
            for (int k = 0; k <= i; k++) v = fma(S[(size_t)(mj + k) * ld + mj + i], z[mj + k], v);
            double u = 0.0;
            for (int r = 0; r < nsn; r++) u = fma(Gn[(size_t)r * w + i], z[an + r], u);
            Q.bR[j * w + i] = v - u;
        }
    }
}

// 2b. separator chain, forward: D_j -= F_{j-1}^T F_{j-1} (upper tiles), r_j -= F_{j-1}^T z_{j-1} (z_{j-1}: the forward-substituted
//     right-hand side of separator j - 1, in place in bR).  One CTA per tile + one for the right-hand side.
__global__ void __launch_bounds__(256)
part_sep_update_kernel(BAPart Q, int j)
{
    const BAState *stq = Q.stp + Q.P + j - 1;
    if (stq->done || !stq->chol_ok) return;
    __shared__ double As[NB][NB + 1], Bs[NB][NB + 1];
    const int w = Q.w, wt = w / NB, nD = wt * (wt + 1) / 2;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const double *F = Q.F + (size_t)(j - 1) * w * w;
    double *D = Q.Dsep + (size_t)j * w * w;
    int id = blockIdx.x;
    if (id < nD) {
        int ti = 0;
        while (id >= wt - ti) { id -= wt - ti; ti++; }
        const int tj = ti + id;
        double sub[4] = {0, 0, 0, 0};
        for (int r0 = 0; r0 < w; r0 += NB) {
            for (int q = ty; q < NB; q += 8) {
                As[q][tx] = F[(size_t)(r0 + q) * w + ti * NB + tx];
                Bs[q][tx] = F[(size_t)(r0 + q) * w + tj * NB + tx];
            }
            __syncthreads();
#pragma unroll 8
            for (int q = 0; q < NB; q++)
#pragma unroll
                for (int e = 0; e < 4; e++) sub[e] = fma(As[q][ty + 8 * e], Bs[q][tx], sub[e]);
            __syncthreads();
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int i = ti * NB + ty + 8 * e, c = tj * NB + tx;
            if (c >= i) D[(size_t)i * w + c] -= sub[e];
        }
    } else {
        const double *zp = Q.bR + (size_t)(j - 1) * w;
        for (int i = tid; i < w; i += 256) {
            double u = 0.0;
            for (int r = 0; r < w; r++) u = fma(F[(size_t)r * w + i], zp[r], u);
            Q.bR[(size_t)j * w + i] -= u;
        }
    }
}

// 2c. separator chain, backward: z_j -= F_j x_{j+1} before the back-substitution of separator j (one warp per row)
__global__ void __launch_bounds__(256)
part_sep_back_kernel(BAPart Q, int j)
{
    const BAState *stq = Q.stp + Q.P + j;
    if (stq->done || !stq->chol_ok) return;
    const int w = Q.w, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *F = Q.F + (size_t)j * w * w, *x = Q.yR + (size_t)(j + 1) * w;
    for (int r = blockIdx.x * 8 + warp; r < w; r += gridDim.x * 8) {
        double v = 0.0;
        for (int c = lane; c < w; c += 32) v = fma(F[(size_t)r * w + c], x[c], v);
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) Q.bR[(size_t)j * w + r] -= v;
    }
}

// 3. with the separators solved (yR): right-hand sides of the segments' back-substitutions.
//    z_s(i) -= G_s(i) x_M(i-1) (i >= 1); separator rows of X_i (i < P - 1) become U_M(i) x_M(i) so that the ordinary
//    back-substitution of the extended system reproduces x_M(i) and continues upwards from it.
__global__ void __launch_bounds__(256)
part_fix_kernel(const double *__restrict__ S, double *__restrict__ z, size_t ld, BAPart Q)
{
    if (Q.stp[0].done) return;
    for (int i = Q.P; i < 2 * Q.P - 1; i++) if (!Q.stp[i].chol_ok) return;
    const int w = Q.w, seg = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a = Q.a[seg], ns = Q.ns[seg];
    for (int r = blockIdx.x * 8 + warp; r < Q.nx[seg]; r += gridDim.x * 8) {
        double v = 0.0;
        if (r < ns) {
            if (seg == 0) continue;
            const double *g = Q.G[seg] + (size_t)r * w, *x = Q.yR + (seg - 1) * w;
            for (int c = lane; c < w; c += 32) v = fma(g[c], x[c], v);
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) z[a + r] -= v;
        } else {
            const int k = r - ns, m = a + ns;
            const double *x = Q.yR + seg * w;
            for (int c = k + lane; c < w; c += 32) v = fma(S[(size_t)(m + k) * ld + m + c], x[c], v);
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) z[a + r] = v;
        }
    }
}

// 4. overall status
__global__ void part_finish_kernel(BAPart Q, BAState *st)
{
    if (st->done) return;
    int ok = 1;
    for (int i = 0; i < 2 * Q.P - 1; i++) ok = ok && Q.stp[i].chol_ok;
    st->chol_ok = ok;
}

}  // namespace

int pmv_internal_ba_cholesky_band(pmv_ctx *ctx, double *S, double *b, int n, size_t ld, const int *lim_host, const int *lim_dev,
                                  BAState *st, cudaStream_t s);   // ba_chol_band.cu

// lim_host[kb] / D.chol_lim[kb]: end column (exclusive, <= n) of the envelope of block row kb -- cumulative
// maximum of the camera co-visibility reach, so fill-in stays inside it.  A banded reduced camera system
// (BASELINE config 5: every point seen by 5 of the <= 40 nearest poses) is factorised in O(n * band^2).
static int launch_backsub(pmv_ctx *ctx, const double *S, const double *z, double *y, int n, size_t ld, const int *lim_host, const int *lim_dev,
                          BAState *st, cudaStream_t s)
{
    int maxw = 0;
    for (int k0 = 0, kb = 0; k0 < n; k0 += NB, kb++) maxw = std::max(maxw, (lim_host ? std::min(n, lim_host[kb]) : n) - k0);
    const int staged = (maxw <= BS_MAXW) && !getenv("PMV_CHOL_NO_STAGE");
    const size_t smem = staged ? sizeof(double) * (2 * NB * BS_MAXW + 512 + 3 * NB) : 0;
    if (ctx->attr_first(PMV_ATTR_CHOL_BACKSUB))
        cudaFuncSetAttribute(chol_backsub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    chol_backsub_kernel<<<1, BS_THREADS, smem, s>>>(S, z, y, n, ld, lim_dev, st, staged);
    PMV_LAUNCH_CHECK(ctx, "chol_backsub_kernel");
    return PMV_OK;
}

// Two-sided solve of a banded system (BASplit, ba.cuh).  Returns 1 when done, 0 when a cluster launch is not possible
// for one of the three systems (caller falls back to the one-sided path; nothing has been modified), < 0 on error.
static int split_solve(pmv_ctx *ctx, const BADev &D, const BASplit &P, cudaStream_t s)
{
    const int n = D.n;
    double *S = D.S, *b = D.rhs, *y = D.yc;
    BAState *st = D.st;
    split_prepare_kernel<<<P.h2 + P.w + 1, 256, 0, s>>>(S, b, n, P.lim_orig, P, st);
    PMV_LAUNCH_CHECK(ctx, "split_prepare_kernel");
    // fork: leading system on s (in place in S), reversed trailing system on s2
    PMV_CUDA_TRY(ctx, cudaEventRecord(P.ev[0], s));
    PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(P.s2, P.ev[0], 0));
    int r1 = pmv_internal_ba_cholesky_band(ctx, S, b, P.h1, (size_t)n, P.lim1_h, P.lim1, P.st3 + 0, s);
    int r2 = pmv_internal_ba_cholesky_band(ctx, P.S2, P.b2, P.h2, (size_t)P.h2, P.lim2_h, P.lim2, P.st3 + 1, P.s2);
    PMV_CUDA_TRY(ctx, cudaEventRecord(P.ev[1], P.s2));
    PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(s, P.ev[1], 0));
    if (r1 < 0 || r2 < 0) return r1 < 0 ? r1 : r2;
    if (r1 == 0 || r2 == 0) return ctx->fail(PMV_ERR_UNSUPPORTED, "split solve: half system not eligible for the cluster kernel");
    split_middle_kernel<<<P.w, 256, 0, s>>>(S, b, n, P);
    PMV_LAUNCH_CHECK(ctx, "split_middle_kernel");
    int r3 = pmv_internal_ba_cholesky_band(ctx, P.S3, P.b3, P.w, (size_t)P.w, P.lim3_h, P.lim3, P.st3 + 2, s);
    if (r3 <= 0) return r3 < 0 ? r3 : ctx->fail(PMV_ERR_UNSUPPORTED, "split solve: separator not eligible for the cluster kernel");
    int rc = launch_backsub(ctx, P.S3, P.b3, P.y3, P.w, (size_t)P.w, P.lim3_h, P.lim3, P.st3 + 2, s);
    if (rc) return rc;
    split_fix_kernel<<<2 * P.w, 256, 0, s>>>(S, b, n, P);
    PMV_LAUNCH_CHECK(ctx, "split_fix_kernel");
    PMV_CUDA_TRY(ctx, cudaEventRecord(P.ev[2], s));
    PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(P.s2, P.ev[2], 0));
    rc = launch_backsub(ctx, S, b, y, P.h1, (size_t)n, P.lim1_h, P.lim1, P.st3 + 0, s);
    if (rc) return rc;
    rc = launch_backsub(ctx, P.S2, P.b2, P.y2, P.h2, (size_t)P.h2, P.lim2_h, P.lim2, P.st3 + 1, P.s2);
    if (rc) return rc;
    PMV_CUDA_TRY(ctx, cudaEventRecord(P.ev[3], P.s2));
    PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(s, P.ev[3], 0));
    split_finish_kernel<<<8, 256, 0, s>>>(y, n, P, st);
    PMV_LAUNCH_CHECK(ctx, "split_finish_kernel");
    return 1;
}

// Partitioned solve of a banded system (BAPart, ba.cuh).  Returns 1 when done, < 0 on error.
static int part_solve(pmv_ctx *ctx, const BADev &D, const BAPart &Q, cudaStream_t s)
{
    const int n = D.n, P = Q.P, w = Q.w, wt = w / NB;
    double *S = D.S, *b = D.rhs, *y = D.yc;
    const size_t ld = (size_t)n, ww = (size_t)w * w;
    part_prepare_kernel<<<1, 32, 0, s>>>(Q, D.st);
    PMV_LAUNCH_CHECK(ctx, "part_prepare_kernel");
    if (ctx->attr_first(PMV_ATTR_CHOL_SPIKE))
        cudaFuncSetAttribute(part_spike_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PS_SMEM);
    // fork: extended system i on its own stream (i = 0 on s): factorisation (+ forward substitution of b), then the spike
    PMV_CUDA_TRY(ctx, cudaEventRecord(Q.ev_fork[0], s));
    for (int i = 0; i < P; i++) {
        cudaStream_t si = i == 0 ? s : Q.str[i];
        if (i > 0) PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(si, Q.ev_fork[0], 0));
        const int r = pmv_internal_ba_cholesky_band(ctx, S + (size_t)Q.a[i] * ld + Q.a[i], b + Q.a[i], Q.nx[i], ld, Q.limX_h[i], Q.limX[i], Q.stp + i, si);
        if (r < 0) return r;
        if (r == 0) return ctx->fail(PMV_ERR_UNSUPPORTED, "part solve: segment not eligible for the cluster kernel");
        if (i > 0) {
            SpikeArgs A;
            A.Sx = S + (size_t)Q.a[i] * ld + Q.a[i]; A.ld = ld; A.nx = Q.nx[i]; A.lim = Q.limX[i]; A.w = w;
            A.bmode = 0; A.Bsrc = S + (size_t)(Q.a[i] - w) * ld + Q.a[i]; A.bld = ld; A.blim = D.chol_lim; A.brow0 = Q.a[i] - w; A.bcol0 = Q.a[i];
            A.G = Q.G[i]; A.st = Q.stp + i;
            part_spike_kernel<<<w / PS_CC, PS_WARPS * 32, PS_SMEM, si>>>(A);
            PMV_LAUNCH_CHECK(ctx, "part_spike_kernel");
            PMV_CUDA_TRY(ctx, cudaEventRecord(Q.ev_join[0][i], si));
            PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(s, Q.ev_join[0][i], 0));
        }
    }
    part_reduce_kernel<<<dim3(wt * (wt + 1) / 2 + wt * wt + 1, P - 1), 256, 0, s>>>(S, b, ld, Q);
    PMV_LAUNCH_CHECK(ctx, "part_reduce_kernel");
    // separator chain, forward
    for (int j = 0; j < P - 1; j++) {
        if (j > 0) {
            part_sep_update_kernel<<<wt * (wt + 1) / 2 + 1, 256, 0, s>>>(Q, j);
            PMV_LAUNCH_CHECK(ctx, "part_sep_update_kernel");
        }
        const int r = pmv_internal_ba_cholesky_band(ctx, Q.Dsep + j * ww, Q.bR + (size_t)j * w, w, (size_t)w, Q.limD_h, Q.limD, Q.stp + P + j, s);
        if (r <= 0) return r < 0 ? r : ctx->fail(PMV_ERR_UNSUPPORTED, "part solve: separator block not eligible for the cluster kernel");
        if (j < P - 2) {
            SpikeArgs A;
            A.Sx = Q.Dsep + j * ww; A.ld = (size_t)w; A.nx = w; A.lim = Q.limD; A.w = w;
            A.bmode = 1; A.Bsrc = Q.E + j * ww; A.bld = (size_t)w; A.blim = nullptr; A.brow0 = 0; A.bcol0 = 0;
            A.G = Q.F + j * ww; A.st = Q.stp + P + j;
            part_spike_kernel<<<w / PS_CC, PS_WARPS * 32, PS_SMEM, s>>>(A);
            PMV_LAUNCH_CHECK(ctx, "part_spike_kernel");
        }
    }
    // separator chain, backward
    for (int j = P - 2; j >= 0; j--) {
        if (j < P - 2) {
            part_sep_back_kernel<<<8, 256, 0, s>>>(Q, j);
            PMV_LAUNCH_CHECK(ctx, "part_sep_back_kernel");
        }
        int rc = launch_backsub(ctx, Q.Dsep + j * ww, Q.bR + (size_t)j * w, Q.yR + (size_t)j * w, w, (size_t)w, Q.limD_h, Q.limD, Q.stp + P + j, s);
        if (rc) return rc;
    }
    part_fix_kernel<<<dim3(32, P), 256, 0, s>>>(S, b, ld, Q);
    PMV_LAUNCH_CHECK(ctx, "part_fix_kernel");
    PMV_CUDA_TRY(ctx, cudaEventRecord(Q.ev_fork[1], s));
    for (int i = 0; i < P; i++) {
        cudaStream_t si = i == 0 ? s : Q.str[i];
        if (i > 0) PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(si, Q.ev_fork[1], 0));
        int rc = launch_backsub(ctx, S + (size_t)Q.a[i] * ld + Q.a[i], b + Q.a[i], y + Q.a[i], Q.nx[i], ld, Q.limX_h[i], Q.limX[i], Q.stp + i, si);
        if (rc) return rc;
        if (i > 0) {
            PMV_CUDA_TRY(ctx, cudaEventRecord(Q.ev_join[1][i], si));
            PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(s, Q.ev_join[1][i], 0));
        }
    }
    part_finish_kernel<<<1, 1, 0, s>>>(Q, D.st);
    PMV_LAUNCH_CHECK(ctx, "part_finish_kernel");
    return 1;
}

int pmv_internal_ba_cholesky_large(pmv_ctx *ctx, const BADev &D, const int *lim_host, const BASplit *split, const BAPart *part,
                                   cudaStream_t s)
{
    const int n = D.n;
    if (part && part->enabled && D.W == 1) {
        chol_gradient_check_kernel<<<1, 1, 0, s>>>(D.st);
        PMV_LAUNCH_CHECK(ctx, "chol_gradient_check_kernel");
        const int r = part_solve(ctx, D, *part, s);
        return r < 0 ? r : PMV_OK;
    }
    if (split && split->enabled && D.W == 1) {
        chol_gradient_check_kernel<<<1, 1, 0, s>>>(D.st);
        PMV_LAUNCH_CHECK(ctx, "chol_gradient_check_kernel");
        const int r = split_solve(ctx, D, *split, s);
        return r < 0 ? r : PMV_OK;
    }
    for (int w = 0; w < D.W; w++) {
        double *S = D.S + (size_t)w * n * n, *b = D.rhs + (size_t)w * n, *y = D.yc + (size_t)w * n;
        BAState *st = D.st + w;
        chol_gradient_check_kernel<<<1, 1, 0, s>>>(st);
        PMV_LAUNCH_CHECK(ctx, "chol_gradient_check_kernel");
        // banded systems: the whole factorisation in one cluster launch (ba_chol_band.cu)
        const int band = pmv_internal_ba_cholesky_band(ctx, S, b, n, (size_t)n, lim_host, D.chol_lim, st, s);
        if (band < 0) return band;
        for (int k0 = 0, kb = 0; k0 < n && !band; k0 += NB, kb++) {
            const int t0 = k0 + NB;
            const int nlim = lim_host ? std::min(n, lim_host[kb]) : n;
            const bool fuse = t0 < nlim && nlim - t0 <= 512;
            chol_diag_kernel<<<1, 256, 0, s>>>(S, b, n, k0, fuse ? nlim : 0, st);
            PMV_LAUNCH_CHECK(ctx, "chol_diag_kernel");
            if (t0 < nlim) {
                if (!fuse) {
                    chol_panel_kernel<<<(nlim - t0 + 127) / 128, 128, 0, s>>>(S, n, k0, nlim, st);
                    PMV_LAUNCH_CHECK(ctx, "chol_panel_kernel");
                }
                const int nt = (nlim - t0 + TW - 1) / TW;
                chol_update_kernel<<<nt * (nt + 1) / 2, 256, 0, s>>>(S, b, n, k0, t0, nlim, st);
                PMV_LAUNCH_CHECK(ctx, "chol_update_kernel");
            }
        }
        {
            int maxw = 0;
            for (int k0 = 0, kb = 0; k0 < n; k0 += NB, kb++) maxw = std::max(maxw, (lim_host ? std::min(n, lim_host[kb]) : n) - k0);
            const int staged = (maxw <= BS_MAXW) && !getenv("PMV_CHOL_NO_STAGE");
            const size_t smem = staged ? sizeof(double) * (2 * NB * BS_MAXW + 512 + 3 * NB) : 0;
            if (ctx->attr_first(PMV_ATTR_CHOL_BACKSUB))
                cudaFuncSetAttribute(chol_backsub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            chol_backsub_kernel<<<1, BS_THREADS, smem, s>>>(S, b, y, n, (size_t)n, D.chol_lim, st, staged);
            PMV_LAUNCH_CHECK(ctx, "chol_backsub_kernel");
        }
    }
    return PMV_OK;
}

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

int pmv_internal_ba_cholesky_large(pmv_ctx *ctx, const BADev &D, const int *lim_host, const BASplit *split, cudaStream_t s)
{
    const int n = D.n;
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

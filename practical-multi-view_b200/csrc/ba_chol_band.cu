// ba_chol_band.cu -- K10 for BANDED reduced camera systems (BASELINE config 5: n = 6000, envelope ~150
// columns): the whole U^T U factorisation and the forward substitution in ONE launch by one thread-block
// cluster whose distributed shared memory holds the sliding window of the band.
//
// Why: the banded factorisation is a chain of n/32 = 188 dependent block steps with ~1 MFLOP each -- pure
// latency.  As separate launches (ba_chol.cu) every step pays 2-3 kernel boundaries and a round trip through
// L2 (~32 us per step, 6 ms per factorisation).  Here the live part of the matrix -- the T x T tile window
// right/below the pivot block, upper triangle -- stays in the shared memory of the 8 CTAs of a cluster, the
// CTAs exchange panel tiles through DSMEM, and a step costs two cluster barriers.
//
// Replaces (like ba_chol.cu) the sparse Cholesky Ceres runs on the reduced camera system
// (SchurComplementSolver, reached from reference CeresBundleAdjustment.cpp:61).
//
// Layout.  Tiles are 32 x 32 doubles, tile (I, J), I <= J, holds S[32 I + r][32 J + c].  The window is a ring
// of T' = T + 1 tile indices (T = widest envelope in tiles): tile (I, J) lives in the slot of the UNORDERED
// residue pair {I mod T', J mod T'}; while the pivot block kb is processed the slots of residue (kb - 1) mod T'
// (retired the step before) are refilled from HBM with the tiles of index kb + T by dedicated loader warps,
// one step ahead of their first use.  Slot {a, b} is owned by CTA (a + b) mod 8.
//
// Warp roles in every CTA (12 warps): 8 COMPUTE warps (panel solves, trailing updates), 3 LOADER warps (HBM
// traffic, right-hand side) and 1 FACTOR warp (the diagonal tile), so the 32 column registers of the
// factorisation are not shared with other code.
//
// One step (pivot block kb), two cluster barriers:
//   A  (end of step kb - 1) the factor warp of the CTA owning tile (kb, kb) subtracts U_(kb-1,kb)^T U_(kb-1,kb)
//      from it and factors it in registers: lane j = column j, 32 pivots as shuffle -> rcp -> FMA chains in four
//      sub-blocks of 8, the rows of the later sub-blocks updated on the fp64 tensor pipe; publishes U_kk and
//      1/diag in its CTA's shared memory
//   -- cluster barrier --
//   B  compute warps owning a panel tile (kb, J) copy U_kk through DSMEM into a private buffer (17 coalesced
//      requests) and solve U_kJ = U_kk^-T S_kJ (one column per lane); the tile stays in its slot with its columns
//      permuted so that a lane's tensor-core fragment is one 32 B piece.  Meanwhile the loader warps write the
//      previous block row to HBM, start the loads of the tiles of index kb + T, and one of them solves
//      z_k = U_kk^-T b_k
//   -- cluster barrier --
//   C  compute warps subtract U_kI^T U_kJ from their trailing tiles (I, J) with fp64 tensor-core tiles
//      (mma.sync.m8n8k4.f64 -> DMMA; tcgen05 has no fp64 kind), operands read from the panel owners' shared
//      memory; loader warps store the incoming tiles and fold b_J -= U_kJ^T z_k; the factor warp runs phase A of
//      step kb + 1.
// Measured on B200 (tools/lat_probe.cu, PMV_CHOL_TRACE=1): a step is 7.2 us = panel solve 2.7 + diagonal-tile
// update 1.8 + factor 2.3 + barriers; remote shared-memory access is request-bound, hence the coalesced copies.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ba.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NB = PMV_CHOL_NB;   // 32
constexpr int CL = 8;             // CTAs per cluster (portable maximum)
constexpr int TLD = 36;           // tile row stride in doubles: conflict-free DMMA fragment loads, 16 B aligned rows
constexpr int TSZ = NB * TLD;
constexpr int ULD = NB + 2;        // even: pairs of U_kk entries are fetched as one 16 B broadcast load
constexpr int CW = 8;             // compute warps per CTA
constexpr int LW = 3;             // loader warps per CTA (12 warps in all: 13 x 144 registers did not fit one SM)
constexpr int NW = CW + LW + 1;    // + the factor warp (phase A), a role of its own so its 32 column registers are not shared with other code
constexpr int FW = CW + LW;
constexpr int THREADS = NW * 32;
constexpr int CSLD = 34;          // row stride of the scratch the tensor-core products of phase A return through
constexpr int UTLD = 10;          // row stride of the transposed sub-block copy of phase A (16 B aligned rows)
constexpr int UBSZ = NB * ULD + 2 * NB;   // U_kk (row stride ULD) | 1/diag | z_k
constexpr int UMAX = 24, PMAX = 4, LMAX = 4;   // per-CTA work-list capacities (trailing / panel / incoming tiles)
constexpr int MAXTP = 17;         // ring size limit (shared memory: <= 21 slots of 9 KB per CTA)

__host__ __device__ __forceinline__ int owner_of(int a, int b) { return (a + b) & (CL - 1); }

__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct BandSmem {
    double *tiles, *Ubuf, *dU, *dinv, *dz, *bwin, *UT;   // Ubuf: [PMAX] private copies of U_kk | 1/diag | z_k for the panel warps
    unsigned char *slot_of;
    uchar4 *ulist;            // [T'][UMAX] trailing tiles owned by this CTA when the pivot residue is r0: {di, dj, ri, rj}
    uchar2 *plist, *llist;    // [T'][PMAX] panel tiles {dj, rj}; [T'][LMAX] incoming tiles {di, ri}
    unsigned char *ucount, *pcount, *lcount;
    unsigned short *jend;     // [nblk] end tile of the envelope of block row kb
};

__device__ __forceinline__ BandSmem carve(double *base, int nslots, int Tp, int nblk)
{
    BandSmem m;
    m.tiles = base;
    m.Ubuf = m.tiles + (size_t)nslots * TSZ;
    m.dU = m.Ubuf + PMAX * UBSZ;
    m.dinv = m.dU + 2 * NB * ULD;
    m.dz = m.dinv + 2 * NB;
    m.bwin = m.dz + 2 * NB;
    m.UT = m.bwin + Tp * NB;
    m.ulist = reinterpret_cast<uchar4 *>(m.UT + NB * UTLD + 24 * CSLD);
    m.plist = reinterpret_cast<uchar2 *>(m.ulist + Tp * UMAX);
    m.llist = m.plist + Tp * PMAX;
    m.jend = reinterpret_cast<unsigned short *>(m.llist + Tp * LMAX);
    m.slot_of = reinterpret_cast<unsigned char *>(m.jend + nblk);
    m.ucount = m.slot_of + Tp * Tp;
    m.pcount = m.ucount + Tp;
    m.lcount = m.pcount + Tp;
    return m;
}

// HBM -> registers -> shared: tile (I, J) of S, zero outside the matrix, identity on the padded diagonal.
// Split in two so the loader warps keep the loads in flight across a cluster barrier.
__device__ __forceinline__ void tile_fetch(const double *__restrict__ S, int n, size_t ld, int I, int J, double2 (&v)[16], int lane)
{
#pragma unroll
    for (int it = 0; it < 16; it++) {
        const int chunk = it * 32 + lane, r = chunk >> 4, c2 = (chunk & 15) * 2;
        const int gr = I * NB + r, gc = J * NB + c2;
        double2 x = make_double2(0.0, 0.0);
        if (gr < n && gc < n) x = *reinterpret_cast<const double2 *>(S + (size_t)gr * ld + gc);
        else if (I == J) { if (r == c2) x.x = 1.0; if (r == c2 + 1) x.y = 1.0; }
        v[it] = x;
    }
}
__device__ __forceinline__ void tile_store(double *dst, const double2 (&v)[16], int lane)
{
#pragma unroll
    for (int it = 0; it < 16; it++) {
        const int chunk = it * 32 + lane, r = chunk >> 4, c2 = (chunk & 15) * 2;
        *reinterpret_cast<double2 *>(dst + r * TLD + c2) = v[it];
    }
}
__device__ __forceinline__ void load_tile(const double *__restrict__ S, int n, size_t ld, int I, int J, double *dst, int lane)
{
    double2 v[16];
    tile_fetch(S, n, ld, I, J, v, lane);
    tile_store(dst, v, lane);
}

// Phase A: factor the diagonal tile of pivot block kb (one warp; lane j keeps column j in registers).
// The chain of 32 dependent pivots (shuffle -> rsqrt -> scale -> shuffle -> FMA) is the critical path of the
// whole factorisation, so the block is processed as 4 sub-blocks of 8 pivots: inside a sub-block only the
// <= 7 rows below the pivot are updated by shuffles; the rows of the later sub-blocks get the 8 rank-1
// updates at once from a transposed copy in shared memory (4 broadcast LDS.128 + 8 FMA per row).
// __noinline__: keeps the 32 column registers out of the caller's allocation (inlined they spilled).
// Column permutation of a solved panel tile in its slot: logical column 8 a + g is stored at 4 g + a, so the four
// columns a lane feeds to the tensor core (g, 8 + g, 16 + g, 24 + g) are one contiguous 32 B piece and an
// operand tile is pulled through DSMEM with 16 requests of 1 KB instead of 64 of 256 B (remote shared-memory
// access is request-bound: ~13-25 cycles per warp-wide request whatever its size).
__device__ __forceinline__ int pcol(int c) { return ((c & 7) << 2) | (c >> 3); }

// 1 / x and 1 / sqrt(x) for the pivots: hardware seed (~20 bits) + one third-order correction (the same
// scheme the CUDA math library uses, minus its range special-casing: pivots of the Jacobi-scaled, damped
// reduced camera system are O(1); a non-positive or non-finite pivot is caught by the `good` flag).
__device__ __forceinline__ double fast_rcp(double x)
{
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    const double e = fma(-x, r0, 1.0);
    const double t = fma(e, e, e);
    return fma(r0, t, r0);
}
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double h = x * y0;
    const double e = fma(-h, y0, 1.0);
    const double p = fma(e, 0.375, 0.5);
    const double ye = y0 * e;
    return fma(p, ye, y0);
}

// Lanes j < i carry don't-care values in col[i] (below the diagonal): nothing reads them, so the updates are
// not predicated (a predicated fp64 update costs two extra selects on a path that is instruction-issue bound).
template <int SB>
__device__ __forceinline__ void phase_a_sub(double (&col)[NB], bool &good, double *dinv, double *UT, int j)
{
#pragma unroll
    for (int k = SB; k < SB + 8; k++) {
        // raw pivot row (before scaling) to every lane: all shuffles issue before the reciprocal is known
        double raw[8];
#pragma unroll
        for (int i = k; i < SB + 8; i++) raw[i - SB] = __shfl_sync(0xffffffffu, col[k], i);
        const double piv = raw[k - SB];
        good = good && (piv > 0.0) && (piv < 1e300);         // warp-uniform; off the dependent chain
        const double rcp = fast_rcp(piv);                    // the chain: shuffle -> rcp -> FMA -> next shuffle
        const double mr = -col[k] * rcp;
#pragma unroll
        for (int i = k + 1; i < SB + 8; i++) col[i] = fma(raw[i - SB], mr, col[i]);   // col[i] - U[k][i] U[k][j]
        const double inv = fast_rsqrt(piv);                  // scaling of the finished row: not on the chain
        col[k] *= inv;                                       // U[k][j] (meaningful for j >= k)
        if (j == k) dinv[k] = inv;
    }
    if (SB + 8 < NB) {
        // Rows of the later sub-blocks get the 8 rank-1 updates of this sub-block at once, on the fp64 tensor
        // pipe: with UT[c][kk] = U[SB + kk][c] (transposed copy in shared memory) the update of the 8x8 block
        // (ib, jb), jb >= ib, is UT_ib UT_jb^T -- two DMMA k-steps whose A and B fragments are the same
        // gather from UT.  The products return through a shared-memory scratch because the factorisation
        // keeps "lane = column" while the tensor core returns "lane = (row, column pair)".
        // (As scalar code this was 384 FMAs that ptxas serialised into load -> FMA chains: 5 800 cycles.)
#pragma unroll
        for (int kk = 0; kk < 8; kk++) UT[j * UTLD + kk] = col[SB + kk];
        __syncwarp();
        const int g = j >> 2, q = j & 3;
        constexpr int B0 = SB / 8 + 1;   // first later 8-block
        double f[4][2];
#pragma unroll
        for (int bb = B0; bb < 4; bb++)
#pragma unroll
            for (int ks = 0; ks < 2; ks++) f[bb][ks] = UT[(8 * bb + g) * UTLD + 4 * ks + q];
        double *Cs = UT + NB * UTLD;     // [24][CSLD] scratch, rows relative to row 8
#pragma unroll
        for (int ib = B0; ib < 4; ib++)
#pragma unroll
            for (int jb = ib; jb < 4; jb++) {
                double c0 = 0.0, c1 = 0.0;
                dmma_8x8x4(c0, c1, f[ib][0], f[jb][0]);
                dmma_8x8x4(c0, c1, f[ib][1], f[jb][1]);
                *reinterpret_cast<double2 *>(Cs + (8 * (ib - 1) + g) * CSLD + 8 * jb + 2 * q) = make_double2(c0, c1);
            }
        __syncwarp();
#pragma unroll
        for (int i = SB + 8; i < NB; i++) col[i] -= Cs[(i - 8) * CSLD + j];   // lanes j < 8 (i / 8): don't care
        __syncwarp();
    }
}

// The forward substitution z_k = U_kk^-T b_k does NOT ride along: interleaved with the pivots it cost 3 000 -
// 7 000 of this function's cycles (a second dependent chain through the same in-order warp); a loader warp
// does it during phase B instead (z_solve).
__device__ __forceinline__ void phase_a(const double *tile, double *dU, double *dinv, double *UT, BAState *st, int lane)
{
    const int j = lane;
    double col[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) col[i] = (i <= j) ? tile[i * TLD + j] : 0.0;
    bool good = true;
    phase_a_sub<0>(col, good, dinv, UT, j);
    phase_a_sub<8>(col, good, dinv, UT, j);
    phase_a_sub<16>(col, good, dinv, UT, j);
    phase_a_sub<24>(col, good, dinv, UT, j);
#pragma unroll
    for (int i = 0; i < NB; i++) dU[i * ULD + j] = col[i];   // entries below the diagonal are never read
    if (!good && lane == 0) st->chol_ok = 0;
}

// z_k = U_kk^-T b_k by one (loader) warp, lane = unknown; publishes z_k for the panel owners and writes it to HBM
__device__ __forceinline__ void z_solve(const double *dU, const double *dinv, const double *bk, double *dz, double *b,
                                        int n, int kb, int lane)
{
    double x = bk[lane];
    double urow[NB];
#pragma unroll
    for (int r = 0; r < NB; r++) urow[r] = dU[r * ULD + lane];   // U[r][lane]
    const double myinv = dinv[lane];
#pragma unroll
    for (int r = 0; r < NB; r++) {
        const double xr = __shfl_sync(0xffffffffu, x * myinv, r);   // z_r (lane r's x is final once rows < r are applied)
        const double upd = fma(-urow[r], xr, x);
        x = (lane > r) ? upd : ((lane == r) ? xr : x);
    }
    dz[lane] = x;
    if (kb * NB + lane < n) b[kb * NB + lane] = x;
}

// b_J -= U_kJ^T z_k for one solved panel tile (column-permuted in its slot), by a loader warp during phase C
__device__ __forceinline__ void panel_rhs(const double *tile, const double *rz /* remote */, double *rb /* remote */, int lane)
{
    const double zl = rz[lane];
    const int pc = pcol(lane);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int r = 0; r < NB; r += 2) {
        s0 = fma(tile[r * TLD + pc], __shfl_sync(0xffffffffu, zl, r), s0);
        s1 = fma(tile[(r + 1) * TLD + pc], __shfl_sync(0xffffffffu, zl, r + 1), s1);
    }
    atomicAdd(rb + lane, -(s0 + s1));
}

// The factored block row leaves for HBM (the back substitution reads it) from a loader warp, off the
// critical path: U_kk from the published copy, z_k likewise.
__device__ __forceinline__ void diag_writeback(const double *dU, double *S, int n, size_t ld, int kb, int lane)
{
    const int k0 = kb * NB, nb = min(NB, n - k0), j = lane;
#pragma unroll 8
    for (int i = 0; i < NB; i++)
        if (i < nb && j < nb && i <= j) S[(size_t)(k0 + i) * ld + k0 + j] = dU[i * ULD + j];
}

// Phase B: U_kJ = U_kk^-T S_kJ for one panel tile (one warp, lane = column), right-looking so the dependent
// chain is 32 multiply + FMA pairs instead of 496 serial FMAs.  U_kk, 1/diag and z_k are first copied from the
// shared memory of the CTA that factored the diagonal tile into a buffer private to this warp (17 coalesced
// 16 B-per-lane requests), then read as broadcasts.
__device__ __forceinline__ void panel_solve(const double *__restrict__ rU /* remote */, const double *__restrict__ rinv /* remote */,
                                            double *ub, double *tile, int lane)
{
    {
        double2 v[17];
        const double2 *src = reinterpret_cast<const double2 *>(rU);
#pragma unroll
        for (int q = 0; q < 17; q++) v[q] = src[q * 32 + lane];          // NB * ULD = 1088 doubles = 17 x 32 double2
        const double iv = rinv[lane];
        double2 *dst = reinterpret_cast<double2 *>(ub);
#pragma unroll
        for (int q = 0; q < 17; q++) dst[q * 32 + lane] = v[q];
        ub[NB * ULD + lane] = iv;
    }
    double x[NB];
#pragma unroll
    for (int r = 0; r < NB; r++) x[r] = tile[r * TLD + lane];
    __syncwarp();
    const double *invd = ub + NB * ULD;
#pragma unroll
    for (int r = 0; r < NB; r++) {
        x[r] *= invd[r];
        const double xr = x[r];
        const double *Ur = ub + r * ULD;
        if ((r & 1) == 0 && r + 1 < NB) x[r + 1] -= Ur[r + 1] * xr;
#pragma unroll
        for (int t = (r + 2) & ~1; t < NB; t += 2) {
            const double2 u = *reinterpret_cast<const double2 *>(Ur + t);
            x[t] -= u.x * xr;
            x[t + 1] -= u.y * xr;
        }
    }
    const int pc = pcol(lane);
#pragma unroll
    for (int r = 0; r < NB; r++) tile[r * TLD + pc] = x[r];   // every lane has read its column (syncwarp above)
}

// A solved panel tile (column-permuted in its slot) leaves for HBM from a loader warp one step later
__device__ __forceinline__ void panel_writeback(const double *tile, double *S, int n, size_t ld, int k0, int J, int lane)
{
    const int col = J * NB + lane, pc = pcol(lane);
    if (col >= n) return;
#pragma unroll 8
    for (int r = 0; r < NB; r++) S[(size_t)(k0 + r) * ld + col] = tile[r * TLD + pc];   // pivot blocks with a panel are full
}

// Phase C: C -= PI^T PJ for one trailing tile (one warp), PI / PJ possibly in another CTA's shared memory.
// DIAG: tile (I, I) -- one operand, and only the 8x8 blocks on or above the diagonal are needed.
template <bool DIAG>
__device__ __forceinline__ void update_tile(double *C, const double *PI, const double *PJ, int lane)
{
    const int g = lane >> 2, q = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c][0] = acc[a][c][1] = 0.0;
#pragma unroll
    for (int kk = 0; kk < NB; kk += 4) {
        // A[row g][k q] = P_I[k][8 a + g], B[k q][col g] = P_J[k][8 c + g]: columns 4 g .. 4 g + 3 of the permuted tile
        double af[4], bf[4];
        {
            const double2 *pa = reinterpret_cast<const double2 *>(PI + (kk + q) * TLD + g * 4);
            const double2 a01 = pa[0], a23 = pa[1];
            af[0] = a01.x; af[1] = a01.y; af[2] = a23.x; af[3] = a23.y;
        }
        if (DIAG) {
#pragma unroll
            for (int c = 0; c < 4; c++) bf[c] = af[c];
        } else {
            const double2 *pb = reinterpret_cast<const double2 *>(PJ + (kk + q) * TLD + g * 4);
            const double2 b01 = pb[0], b23 = pb[1];
            bf[0] = b01.x; bf[1] = b01.y; bf[2] = b23.x; bf[3] = b23.y;
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (!DIAG || c >= a) dmma_8x8x4(acc[a][c][0], acc[a][c][1], af[a], bf[c]);
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (DIAG && c < a) continue;
            double2 *p = reinterpret_cast<double2 *>(C + (a * 8 + g) * TLD + c * 8 + 2 * q);
            double2 v = *p;
            v.x -= acc[a][c][0]; v.y -= acc[a][c][1];
            *p = v;
        }
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
chol_band_cluster_kernel(double *S, double *b, int n, size_t ld, const int *__restrict__ lim, BAState *st, int T, int nslots,
                         long long *trace /* PMV_CHOL_TRACE=1: per-warp clock64 stamps of 8 steps, else nullptr */)
{
    extern __shared__ __align__(16) double band_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    if (st->done) return;   // uniform over the cluster: nobody reaches a barrier
    const int me = (int)cluster.block_rank();
    const int Tp = T + 1;
    const int nblk = (n + NB - 1) / NB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const BandSmem m = carve(band_smem, nslots, Tp, nblk);
    constexpr int TR0 = 64, TRN = 8;   // traced steps
    auto stamp = [&](int kb, int ev) {
        if (trace && lane == 0 && kb >= TR0 && kb < TR0 + TRN)
            trace[(((size_t)(kb - TR0) * CL + me) * NW + warp) * 5 + ev] = clock64();
    };
    auto tile_ptr = [&](int a, int c) { return m.tiles + (size_t)m.slot_of[a * Tp + c] * TSZ; };

    for (int t = tid; t < Tp * Tp; t += THREADS) {
        const int a = t / Tp, c = t % Tp, lo = min(a, c), hi = max(a, c), o = owner_of(lo, hi);
        int cnt = 0;
        bool before = true;
        for (int x = 0; x < Tp; x++)
            for (int y = x; y < Tp; y++) {
                if (x == lo && y == hi) before = false;
                if (before && owner_of(x, y) == o) cnt++;
            }
        m.slot_of[t] = (unsigned char)cnt;
    }
    // tiles [kb, jend[kb]) of block row kb are inside the envelope
    for (int kb = tid; kb < nblk; kb += THREADS) m.jend[kb] = (unsigned short)min(nblk, (min(n, lim[kb]) + NB - 1) / NB);
    // work lists of this CTA for every pivot residue r0 (offsets di / dj from the pivot block, ring residues)
    if (tid < Tp) {
        const int r0 = tid;
        int nu = 0, np = 0, nl = 0;
        for (int di = 1; di < T; di++) {
            const int ri = (r0 + di) % Tp;
            if (owner_of(r0, ri) == me && np < PMAX) m.plist[r0 * PMAX + np++] = make_uchar2(di, ri);
            for (int dj = di; dj < T; dj++) {
                const int rj = (r0 + dj) % Tp;
                if (owner_of(ri, rj) == me && nu < UMAX) m.ulist[r0 * UMAX + nu++] = make_uchar4(di, dj, ri, rj);
            }
        }
        const int rn = (r0 + T) % Tp;   // == (r0 - 1) mod T': residue retired by the previous step, refilled now
        for (int d = 0; d <= T; d++) {
            const int rd = (r0 + d) % Tp;
            if (owner_of(rd, rn) == me && nl < LMAX) m.llist[r0 * LMAX + nl++] = make_uchar2(d, rd);
        }
        m.ucount[r0] = (unsigned char)nu; m.pcount[r0] = (unsigned char)np; m.lcount[r0] = (unsigned char)nl;
    }
    if (me == 0 && tid == 0) st->chol_ok = 1;
    __syncthreads();

    // prologue: the window of pivot block 0 (tile indices < T)
    {
        const int wend = min(T, nblk);
        int idx = 0;
        for (int I = 0; I < wend; I++) {
            const int je = min((int)m.jend[I], wend);
            for (int J = I; J < je; J++) {
                if (owner_of(I, J) != me) continue;
                if (idx++ % NW == warp) load_tile(S, n, ld, I, J, tile_ptr(I, J), lane);
            }
            if (owner_of(I, I) == me && warp == I % NW)
                m.bwin[I * NB + lane] = (I * NB + lane < n) ? b[I * NB + lane] : 0.0;
        }
    }
    cluster.sync();
    if (me == owner_of(0, 0) && warp == FW) phase_a(tile_ptr(0, 0), m.dU, m.dinv, m.UT, st, lane);

    if (warp == FW) {
        // ---- factor warp: after the panels of step kb are ready it updates the next diagonal tile and factors
        // it (phase A of step kb + 1) while the compute warps update the rest of the window
        for (int kb = 0, r0 = 0; kb < nblk; kb++, r0 = (r0 + 1 == Tp ? 0 : r0 + 1)) {
            cluster.sync();
            stamp(kb, 0); stamp(kb, 1);
            cluster.sync();
            stamp(kb, 2);
            const int r1 = (r0 + 1 == Tp ? 0 : r0 + 1);
            if (kb + 1 < nblk && owner_of(r1, r1) == me) {
                double *dt = tile_ptr(r1, r1);
                if (kb + 1 < (int)m.jend[kb]) {
                    const double *PI = cluster.map_shared_rank(tile_ptr(r0, r1), owner_of(r0, r1));
                    update_tile<true>(dt, PI, PI, lane);
                    __syncwarp();
                }
                stamp(kb, 3);
                const int p1 = (kb + 1) & 1;
                phase_a(dt, m.dU + p1 * NB * ULD, m.dinv + p1 * NB, m.UT, st, lane);
            } else {
                stamp(kb, 3);
            }
            stamp(kb, 4);
        }
    } else if (warp >= CW) {
        // ---- loader warps: two barriers per step like everybody else.
        //   phase B: write the block row factored / solved in the previous phases to HBM, start the loads of the
        //            tiles of index kb + T (HBM -> registers), solve z_k = U_kk^-T b_k (one warp of the CTA
        //            that owns the diagonal tile)
        //   phase C: registers -> the slots retired by the previous step; b_J -= U_kJ^T z_k for the panel
        //            tiles this CTA solved
        // At most one incoming tile per loader warp: T' <= 17 gives every CTA at most 3 of the <= 17 slots.
        const int lw = warp - CW;
        for (int kb = 0, r0 = 0; kb < nblk; kb++, r0 = (r0 + 1 == Tp ? 0 : r0 + 1)) {
            cluster.sync();
            stamp(kb, 0);
            const int par = kb & 1;
            double2 lv[16];
            double *ldst = nullptr;
            double lb = 0.0;
            bool lb_set = false;
            const int Jn = kb + T;
            int rn = r0 + T; if (rn >= Tp) rn -= Tp;
            if (owner_of(r0, r0) == me) {
                if (lw == LW - 1) z_solve(m.dU + par * NB * ULD, m.dinv + par * NB, m.bwin + r0 * NB, m.dz + par * NB, b, n, kb, lane);
                if (lw == LW - 2) diag_writeback(m.dU + par * NB * ULD, S, n, ld, kb, lane);
            }
            {
                // slot {rd, rn}: held panel tile (kb - 1, kb + d) of the previous step (d <= T - 2) -> HBM;
                // receives tile (kb + d, kb + T) (d >= 1) if that tile exists and lies inside the envelope
                const int nl = m.lcount[r0];
                const int jprev = kb > 0 ? (int)m.jend[kb - 1] : 0;
                for (int e = lw; e < nl; e += LW) {
                    const uchar2 t = m.llist[r0 * LMAX + e];
                    const int d = t.x, J = kb + d;
                    double *slot = tile_ptr(t.y, rn);
                    if (kb > 0 && d <= T - 2 && J < jprev) panel_writeback(slot, S, n, ld, (kb - 1) * NB, J, lane);
                    if (d >= 1 && Jn < nblk && Jn < (int)m.jend[J]) { tile_fetch(S, n, ld, J, Jn, lv, lane); ldst = slot; }
                }
                if (Jn < nblk && owner_of(rn, rn) == me && lw == 0) { lb = (Jn * NB + lane < n) ? b[Jn * NB + lane] : 0.0; lb_set = true; }
            }
            stamp(kb, 1);
            cluster.sync();
            stamp(kb, 2);
            if (ldst) tile_store(ldst, lv, lane);
            if (lb_set) m.bwin[rn * NB + lane] = lb;
            {
                const int je = m.jend[kb], np = m.pcount[r0];
                const double *rz = cluster.map_shared_rank(m.dz + par * NB, owner_of(r0, r0));
                for (int e = lw; e < np; e += LW) {
                    const uchar2 t = m.plist[r0 * PMAX + e];
                    const int J = kb + t.x, rj = t.y;
                    if (J >= je) continue;
                    double *rb = cluster.map_shared_rank(m.bwin + rj * NB, owner_of(rj, rj));
                    panel_rhs(tile_ptr(r0, rj), rz, rb, lane);
                }
            }
            stamp(kb, 3); stamp(kb, 4);
        }
    } else {
        // ---- compute warps
        for (int kb = 0, r0 = 0; kb < nblk; kb++, r0 = (r0 + 1 == Tp ? 0 : r0 + 1)) {
            const int par = kb & 1;
            const int je = m.jend[kb];
            cluster.sync();   // U_kk published; every trailing update of the previous step is complete
            stamp(kb, 0);
            if (je > kb + 1) {
                // phase B: one warp per owned panel tile (np <= PMAX); each copies U_kk and 1/diag from the shared
                // memory of the CTA that factored the diagonal tile into its private buffer Ubuf[warp]
                const int od = owner_of(r0, r0);
                const double *rU = cluster.map_shared_rank(m.dU + par * NB * ULD, od);
                const double *rinv = cluster.map_shared_rank(m.dinv + par * NB, od);
                const int np = m.pcount[r0];
                if (warp < np) {
                    const uchar2 t = m.plist[r0 * PMAX + warp];
                    if (kb + t.x < je) panel_solve(rU, rinv, m.Ubuf + warp * UBSZ, tile_ptr(r0, t.y), lane);
                }
            }
            stamp(kb, 1);
            cluster.sync();   // panel tiles of block row kb ready in their owners' shared memory
            stamp(kb, 2);
            // phase C: trailing update of the owned tiles, except the next diagonal tile (always the first
            // entry of the list when this CTA owns it), which belongs to the factor warp
            const int r1 = (r0 + 1 == Tp ? 0 : r0 + 1);
            const bool own_next = (kb + 1 < nblk) && owner_of(r1, r1) == me && T > 1;
            const int nu = m.ucount[r0];
            for (int e = (own_next ? 1 : 0) + warp; e < nu; e += CW) {
                const uchar4 t = m.ulist[r0 * UMAX + e];
                const int J = kb + t.y, ri = t.z, rj = t.w;
                if (J >= je) continue;
                const double *PI = cluster.map_shared_rank(tile_ptr(r0, ri), owner_of(r0, ri));
                if (t.x == t.y) {
                    update_tile<true>(tile_ptr(ri, ri), PI, PI, lane);
                } else {
                    const double *PJ = cluster.map_shared_rank(tile_ptr(r0, rj), owner_of(r0, rj));
                    update_tile<false>(tile_ptr(ri, rj), PI, PJ, lane);
                }
            }
            stamp(kb, 3); stamp(kb, 4);
        }
    }
    cluster.sync();   // nobody leaves while its shared memory may still be read
}

}  // namespace

// Window width T (in tiles) the cluster kernel would use for an n x n system with this envelope; 0 = not eligible.
int pmv_internal_ba_cholesky_band_T(int n, const int *lim_host)
{
    if (!lim_host) return 0;
    const int nblk = (n + NB - 1) / NB;
    int T = 1;
    for (int kb = 0; kb < nblk; kb++) {
        const int je = std::min(nblk, (std::min(n, lim_host[kb]) + NB - 1) / NB);
        T = std::max(T, je - kb);
    }
    if (T < 2) T = 2;
    const int Tp = T + 1;
    if (Tp > MAXTP) return 0;
    int cnt[CL] = {0};
    for (int x = 0; x < Tp; x++)
        for (int y = x; y < Tp; y++) cnt[owner_of(x, y)]++;
    const int nslots = *std::max_element(cnt, cnt + CL);
    const size_t smem = sizeof(double) * ((size_t)nslots * TSZ + PMAX * UBSZ + 2 * NB * ULD + 4 * NB + (size_t)Tp * NB + NB * UTLD + 24 * CSLD) +
                        (size_t)Tp * (UMAX * 4 + PMAX * 2 + LMAX * 2 + 3) + (size_t)nblk * 2 + (size_t)Tp * Tp + 32;
    if (nslots > UMAX || nblk > 65535 || smem > 220 * 1024) return 0;
    return T;
}

// Host side: eligibility (window width from the envelope) and launch.  Returns 1 when the cluster kernel was
// launched, 0 when the system is too wide for it (caller falls back to the multi-launch path), < 0 on error.
int pmv_internal_ba_cholesky_band(pmv_ctx *ctx, double *S, double *b, int n, size_t ld, const int *lim_host, const int *lim_dev,
                                  BAState *st, cudaStream_t s)
{
    if (!lim_host || !lim_dev) return 0;
    const char *off = getenv("PMV_CHOL_NO_CLUSTER");
    if (off && off[0] == '1') return 0;
    const int nblk = (n + NB - 1) / NB;
    int T = 1;
    for (int kb = 0; kb < nblk; kb++) {
        const int je = std::min(nblk, (std::min(n, lim_host[kb]) + NB - 1) / NB);
        T = std::max(T, je - kb);
    }
    if (T < 2) T = 2;   // a tile entering the window must not be needed in the step that loads it
    const int Tp = T + 1;
    if (Tp > MAXTP) return 0;
    int cnt[CL] = {0};
    for (int x = 0; x < Tp; x++)
        for (int y = x; y < Tp; y++) cnt[owner_of(x, y)]++;
    const int nslots = *std::max_element(cnt, cnt + CL);
    const size_t smem = sizeof(double) * ((size_t)nslots * TSZ + PMAX * UBSZ + 2 * NB * ULD + 4 * NB + (size_t)Tp * NB + NB * UTLD + 24 * CSLD) +
                        (size_t)Tp * (UMAX * 4 + PMAX * 2 + LMAX * 2 + 3) + (size_t)nblk * 2 + (size_t)Tp * Tp + 32;
    if (nslots > UMAX || nblk > 65535) return 0;
    if (smem > 220 * 1024) return 0;
    if (ctx->attr_first(PMV_ATTR_CHOL_BAND)) {
        cudaError_t e = cudaFuncSetAttribute(chol_band_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return ctx->fail(PMV_ERR_CUDA, "chol_band_cluster_kernel attribute", e);
    }
    const char *tr = getenv("PMV_CHOL_TRACE");
    if (tr && tr[0] == '1') {
        // diagnostic: where the time of a step goes, per CTA and warp (clock64 deltas, averaged over 8 steps)
        const size_t cnt = (size_t)8 * CL * NW * 5;
        long long *d_tr = nullptr;
        std::vector<long long> h(cnt);
        cudaMalloc(&d_tr, cnt * 8);
        cudaMemsetAsync(d_tr, 0, cnt * 8, s);
        chol_band_cluster_kernel<<<CL, THREADS, smem, s>>>(S, b, n, ld, lim_dev, st, T, nslots, d_tr);
        cudaStreamSynchronize(s);
        cudaMemcpy(h.data(), d_tr, cnt * 8, cudaMemcpyDeviceToHost);
        cudaFree(d_tr);
        static int printed = 0;
        if (!printed++) {
            fprintf(stderr, "chol_band trace (T=%d, cycles; B work | wait bar2 | C work | A | wait bar1):\n", T);
            for (int c = 0; c < CL; c++)
                for (int w = 0; w < NW; w++) {
                    double d[5] = {0, 0, 0, 0, 0};
                    for (int k = 0; k < 7; k++) {
                        const long long *e = &h[(((size_t)k * CL + c) * NW + w) * 5];
                        const long long *nx = &h[(((size_t)(k + 1) * CL + c) * NW + w) * 5];
                        d[0] += e[1] - e[0]; d[1] += e[2] - e[1]; d[2] += e[3] - e[2]; d[3] += e[4] - e[3]; d[4] += nx[0] - e[4];
                    }
                    fprintf(stderr, "  cta %d warp %2d: %7.0f %7.0f %7.0f %7.0f %7.0f\n", c, w, d[0] / 7, d[1] / 7, d[2] / 7, d[3] / 7, d[4] / 7);
                }
        }
        ctx->launches++;
        return 1;
    }
    chol_band_cluster_kernel<<<CL, THREADS, smem, s>>>(S, b, n, ld, lim_dev, st, T, nslots, nullptr);
    PMV_LAUNCH_CHECK(ctx, "chol_band_cluster_kernel");
    return 1;
}

// ba_window.cu -- window-batched bundle adjustment (BASELINE config 4: thousands of independent
// windows of <= 22 poses): K8 + K9 + K10 fused per window, no Jacobian ever written to HBM.
//
// One CTA owns one window for a whole LM iteration:
//   win_prepare_kernel   per camera: rotation R(a) and dR/da_k (so an observation costs ~100 flops, no
//                        trig), Jacobi column norms on the first linearisation
//   win_schur_kernel     per chunk of P points: thread (point slot, camera) evaluates residual + Jacobians
//                        (Huber-corrected, Jacobi-scaled), per-point V^-1 in shared memory, Y = W V^-1 and
//                        W blocks staged in shared memory; then thread (ci <= cj) accumulates its 6x6 block
//                        of S -= Y_ci W_cj^T in REGISTERS over all points (the K = 3 Np contraction that
//                        dominates: Nc(Nc+1)/2 * 108 FMA per point).  Afterwards S (+ U_c + D_c^2) is laid
//                        out in shared memory, factorised (Cholesky) and solved in place -> y_c.
//   win_backsub_kernel   per point: y_p, candidate point, model cost change, candidate cost.
// LM bookkeeping reuses ba_lm_update_kernel / ba_accept_kernel.  Same arithmetic as the general path
// (ba_kernels.cuh) up to fp64 summation order; replaces ceres::Solve at CeresBundleAdjustment.cpp:61.
#include <algorithm>

#include "ba.cuh"
#include "ba_kernels.cuh"

namespace {

constexpr int WIN_MAXC = 22;          // Nc(Nc+1)/2 <= 253 tile threads
constexpr int WIN_THREADS = 256;

struct WinDev {
    const unsigned *vis;   // per (window, point): bit c set <=> camera c observes it (observations sorted by camera)
    double *camR;          // per (window, camera): R (9) | dR/da_0 (9) | dR/da_1 (9) | dR/da_2 (9)
    double *candR;         // per (window, camera): R of the candidate pose (9)
    int P;                 // points per chunk
    int LD;                // stride (doubles) between k-columns of the chunk matrices, == 4 (mod 16)
    int passes;            // producer passes per chunk (1 or 2)
};

// rotation matrix and its derivatives w.r.t. the angle-axis vector, via the same code as the residual
__device__ void cam_matrices(const double *pose, double *out36)
{
    const double a[6] = {pose[0], pose[1], pose[2], 0.0, 0.0, 0.0};
    for (int j = 0; j < 3; j++) {
        double e[3] = {0, 0, 0};
        e[j] = 1.0;
        double p[3], R[9], dpa[9];
        ba_project(a, e, p, R, dpa, true);
        for (int i = 0; i < 3; i++) {
            out36[3 * i + j] = p[i];                                  // column j of R
            for (int k = 0; k < 3; k++) out36[9 + 9 * k + 3 * i + j] = dpa[3 * k + i];   // column j of dR/da_k
        }
    }
}

__global__ void __launch_bounds__(128) win_prepare_kernel(const BADev D, const WinDev Wd, int candidate)
{
    const int wc = blockIdx.x * 128 + threadIdx.x;
    if (wc >= D.W * D.Nc) return;
    const BAState *st = &D.st[wc / D.Nc];
    if (st->done) return;
    if (candidate) {
        if (!st->chol_ok) return;
        double m[36];
        cam_matrices(D.cand_poses + 6 * (size_t)wc, m);
        for (int k = 0; k < 9; k++) Wd.candR[9 * (size_t)wc + k] = m[k];
    } else {
        if (!st->need_linearize) return;
        double m[36];
        cam_matrices(D.poses + 6 * (size_t)wc, m);
        for (int k = 0; k < 36; k++) Wd.camR[36 * (size_t)wc + k] = m[k];
    }
}

// residual + Jacobians of one observation from the per-camera matrices (q = X + c)
__device__ __forceinline__ void win_residual_jac(const double *M /* 36, shared */, const double *c3, const double *X,
                                                 double ox, double oy, double fx, double cx, double fy, double cy,
                                                 double r[2], double Jc[12], double Jp[6])
{
    const double q0 = X[0] + c3[0], q1 = X[1] + c3[1], q2 = X[2] + c3[2];
    double p[3], dp[9];
#pragma unroll
    for (int i = 0; i < 3; i++) p[i] = M[3 * i] * q0 + M[3 * i + 1] * q1 + M[3 * i + 2] * q2;
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int i = 0; i < 3; i++)
            dp[3 * k + i] = M[9 + 9 * k + 3 * i] * q0 + M[9 + 9 * k + 3 * i + 1] * q1 + M[9 + 9 * k + 3 * i + 2] * q2;
    const double iz = ba_rcp_fast(p[2]);   // one refined reciprocal instead of three divisions (ba_kernels.cuh)
    r[0] = ox - (-(p[0] * iz) * fx + cx);
    r[1] = oy - (-(p[1] * iz) * fy + cy);
    const double a0 = fx * iz, a2 = -fx * p[0] * iz * iz, b1 = fy * iz, b2 = -fy * p[1] * iz * iz;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        Jc[k] = a0 * dp[3 * k] + a2 * dp[3 * k + 2];
        Jc[6 + k] = b1 * dp[3 * k + 1] + b2 * dp[3 * k + 2];
        const double j0 = a0 * M[k] + a2 * M[6 + k], j1 = b1 * M[3 + k] + b2 * M[6 + k];
        Jc[3 + k] = j0; Jc[9 + k] = j1;
        Jp[k] = j0; Jp[3 + k] = j1;
    }
}

// Jacobi column norms of the first linearisation (unscaled, Huber-corrected): scale = 1 / (1 + |J_:j|)
__global__ void __launch_bounds__(WIN_THREADS) win_colnorm_kernel(const BADev D, const WinDev Wd)
{
    extern __shared__ double sm[];
    const int w = blockIdx.x, tid = threadIdx.x, Nc = D.Nc;
    const BAState *st = &D.st[w];
    if (st->done || st->scale_ready || !st->need_linearize) return;
    double *sM = sm;                    // Nc*36
    double *sC = sM + Nc * 36;          // Nc*3
    double *sAcc = sC + Nc * 3;         // Nc*6 camera column sums
    for (int i = tid; i < Nc * 36; i += WIN_THREADS) sM[i] = Wd.camR[36 * (size_t)w * Nc + i];
    for (int i = tid; i < Nc * 3; i += WIN_THREADS) sC[i] = D.poses[6 * ((size_t)w * Nc + i / 3) + 3 + i % 3];
    for (int i = tid; i < Nc * 6; i += WIN_THREADS) sAcc[i] = 0.0;
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    double cacc[6] = {0, 0, 0, 0, 0, 0};   // lane = camera
    for (int p = warp; p < D.Np; p += WIN_THREADS / 32) {
        const size_t wp = (size_t)w * D.Np + p;
        const unsigned mask = Wd.vis[wp];
        double ps[3] = {0, 0, 0};
        if (lane < Nc && (mask >> lane & 1)) {
            const int i = D.pt_off[wp] + __popc(mask & ((1u << lane) - 1));
            double r[2], jc[12], jp[6];
            win_residual_jac(sM + 36 * lane, sC + 3 * lane, D.points + 3 * wp, D.obs_xy[2 * (size_t)i], D.obs_xy[2 * (size_t)i + 1],
                             D.fx, D.cx, D.fy, D.cy, r, jc, jp);
            double rho0, rho1;
            ba_huber(D.delta, r[0] * r[0] + r[1] * r[1], rho0, rho1);
#pragma unroll
            for (int a = 0; a < 6; a++) cacc[a] += rho1 * (jc[a] * jc[a] + jc[6 + a] * jc[6 + a]);
#pragma unroll
            for (int a = 0; a < 3; a++) ps[a] = rho1 * (jp[a] * jp[a] + jp[3 + a] * jp[3 + a]);
        }
#pragma unroll
        for (int a = 0; a < 3; a++) ps[a] = warp_sum_d(ps[a]);
        if (lane < 3) D.scale_p[3 * wp + lane] = 1.0 / (1.0 + sqrt(ps[lane]));
    }
    if (lane < Nc)
        for (int a = 0; a < 6; a++) atomicAdd(&sAcc[6 * lane + a], cacc[a]);
    __syncthreads();
    for (int i = tid; i < Nc * 6; i += WIN_THREADS) D.scale_c[6 * (size_t)w * Nc + i] = 1.0 / (1.0 + sqrt(sAcc[i]));
}

constexpr int WS_THREADS = 512;       // 8 consumer warps (tiles) + 8 producer warps (observations)
constexpr int WS_PRODUCERS = WS_THREADS - 256;

__device__ __forceinline__ void win_dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void producer_barrier()
{
    asm volatile("bar.sync 1, %0;" ::"n"(WS_PRODUCERS) : "memory");
}

// Warp-specialised: producers (threads 256..383) evaluate chunk k into buffer k&1 while consumers
// (threads 0..255, one 6x6 tile of S each) contract chunk k-1; one CTA-wide barrier per chunk.
// WS_MAXST: 16x16 super-tiles of S per consumer warp (NS(NS+1)/2 super-tiles over 8 warps: 5 up to n = 128, else 6)
template <int WS_MAXST>
__global__ void __launch_bounds__(WS_THREADS, 1) win_schur_kernel(const BADev D, const WinDev Wd)
{
    extern __shared__ double sm[];
    const int w = blockIdx.x, tid = threadIdx.x, Nc = D.Nc, n = D.n, P = Wd.P;
    BAState *st = &D.st[w];
    if (st->done) return;
    const bool lin = st->need_linearize != 0;
    // ---- shared memory carve-up -----------------------------------------------------------------------
    double *sM = sm;                         // Nc*36 camera matrices
    double *sC = sM + Nc * 36;               // Nc*3  camera centres c
    double *sSc = sC + Nc * 3;               // Nc*6  Jacobi scale of the camera columns
    double *sRed = sSc + Nc * 6;             // Nc*33 reduced camera sums (U 21 | g 6 | rhs 6)
    double *sScal = sRed + Nc * 33;          // 8 scalars
    double *sUnion = sScal + 8;
    // phase A (points loop)
    double *sItem = sUnion;                  // P*Nc*20 : Jp (6) | r (2) | Jc (12) per item
    double *sPt = sItem + P * Nc * 20;       // P*12   : Vinv (6) | g (3) | pad
    double *sVg = sPt + P * 12;              // P*9    : raw V (6) | g (3) sums
    // chunk buffers: Y and W stored k-major, M[k = 3*slot + b][row = 6*cam + a] with row stride
    // LD = 16*NS + 4 (== 4 mod 16): a producer writes 6 contiguous doubles per column and the fp64 mma
    // fragment loads (lanes = 8 rows x 4 k) are conflict-free; two buffers
    const int NS = (n + 15) >> 4, LD = Wd.LD, KS = (3 * P + 3) >> 2;
    double *sYW = sVg + P * 9;
    const size_t bufsz = (size_t)4 * KS * LD;
    unsigned *sMask = reinterpret_cast<unsigned *>(sYW + 4 * bufsz);   // 2 x P
    // phase B (factorisation) aliases the union region
    double *A = sUnion;                      // n x (n+1)
    double *bvec = A + (size_t)n * (n + 1);  // n

    for (int i = tid; i < Nc * 36; i += WS_THREADS) sM[i] = Wd.camR[36 * (size_t)w * Nc + i];
    for (int i = tid; i < Nc * 3; i += WS_THREADS) sC[i] = D.poses[6 * ((size_t)w * Nc + i / 3) + 3 + i % 3];
    for (int i = tid; i < Nc * 6; i += WS_THREADS) sSc[i] = D.scale_c[6 * (size_t)w * Nc + i];
    for (int i = tid; i < Nc * 33; i += WS_THREADS) sRed[i] = 0.0;
    if (tid < 8) sScal[tid] = 0.0;
    for (size_t i = tid; i < 4 * bufsz; i += WS_THREADS) sYW[i] = 0.0;   // padding rows / columns stay zero
    __syncthreads();
    const double radius = st->radius, inv_radius = 1.0 / radius;
    const int nchunks = (D.Np + P - 1) / P;
    const bool is_producer = tid >= 256;

    const int ld = n + 1;
    // The two roles keep their accumulators in DISJOINT code paths so the register allocation is the
    // maximum of the two, not the sum.  cta_barrier() = one hand-off per chunk.
    auto cta_barrier = []() { asm volatile("bar.sync 2, %0;" ::"n"(WS_THREADS) : "memory"); };

    if (is_producer) {
        // producer thread (slot s, camera c) evaluates points s and s + P/2 of every chunk for camera c,
        // so its camera sums (U 21 | g 6 | rhs 6) stay in registers for the whole window
        const int ptid = tid - 256;
        const int passes = Wd.passes;            // 1 when every (point slot, camera) item has its own thread, else 2
        const int H = P / passes;                // point slots per pass
        const int pslot = ptid / Nc, pcam = ptid - pslot * Nc;
        const bool producer = pslot < H;
        double accU[21], accG[6], accR[6];
#pragma unroll
        for (int k = 0; k < 21; k++) accU[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) accG[k] = accR[k] = 0.0;
        double cost = 0.0, gmax = 0.0;
        for (int step = 0; step <= nchunks; step++) {
            if (step < nchunks) {
                const int p0 = step * P, buf = step & 1;
                double *sY = sYW + (size_t)buf * 2 * bufsz, *sW = sY + bufsz;
                unsigned *mk = sMask + buf * P;
                // ---- A1: residual + Jacobians of items (point p0+slot, camera pcam) -> shared memory ---------------
                for (int pass = 0; pass < passes; pass++) {
                    const int slot = pslot + pass * H;
                    if (!producer) break;
                    double *it = sItem + (size_t)(slot * Nc + pcam) * 20;
                    if (p0 + slot < D.Np) {
                        const size_t wp = (size_t)w * D.Np + p0 + slot;
                        const unsigned mask = Wd.vis[wp];
                        if (pcam == 0) mk[slot] = mask;
                        if ((mask >> pcam) & 1) {
                            double jc[12], jp[6], r[2];
                            const int i = D.pt_off[wp] + __popc(mask & ((1u << pcam) - 1));
                            win_residual_jac(sM + 36 * pcam, sC + 3 * pcam, D.points + 3 * wp, D.obs_xy[2 * (size_t)i],
                                             D.obs_xy[2 * (size_t)i + 1], D.fx, D.cx, D.fy, D.cy, r, jc, jp);
                            double rho0, rho1;
                            ba_huber(D.delta, r[0] * r[0] + r[1] * r[1], rho0, rho1);
                            cost += 0.5 * rho0;
                            const double sr = ba_sqrt_rho1(rho1);
                            const double *sp = D.scale_p + 3 * wp;
#pragma unroll
                            for (int k = 0; k < 3; k++) { const double sc = sr * sp[k]; it[k] = jp[k] * sc; it[3 + k] = jp[3 + k] * sc; }
                            it[6] = r[0] * sr; it[7] = r[1] * sr;
#pragma unroll
                            for (int k = 0; k < 6; k++) { const double sc = sr * sSc[6 * pcam + k]; it[8 + k] = jc[k] * sc; it[14 + k] = jc[6 + k] * sc; }
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; k++) it[k] = 0.0;
                        }
                    } else {
                        if (pcam == 0) mk[slot] = 0u;
#pragma unroll
                        for (int k = 0; k < 8; k++) it[k] = 0.0;
                    }
                }
                producer_barrier();
                // ---- A2a: raw V (6) and g (3) sums, thread (point slot, component) --------------------------------
                for (int idx = ptid; idx < P * 9; idx += WS_PRODUCERS) {
                    const int q = idx / 9, comp = idx - q * 9;
                    const int ia[9] = {0, 0, 0, 1, 1, 2, 0, 1, 2}, ib[9] = {0, 1, 2, 1, 2, 2, 6, 6, 6};
                    const int a = ia[comp], b2 = ib[comp];
                    double sacc = 0;
                    for (int c = 0; c < Nc; c++) {
                        const double *it = sItem + (size_t)(q * Nc + c) * 20;
                        sacc += (b2 == 6) ? (it[a] * it[6] + it[3 + a] * it[7]) : (it[a] * it[b2] + it[3 + a] * it[3 + b2]);
                    }
                    sVg[q * 9 + comp] = sacc;
                }
                producer_barrier();
                // ---- A2b: (V + D^2)^-1 by adjugate / determinant (one thread per point slot) ---------------------
                if (ptid < P && p0 + ptid < D.Np) {
                    const size_t wpp = (size_t)w * D.Np + p0 + ptid;
                    double V[6], g[3];
                    for (int k = 0; k < 6; k++) V[k] = sVg[ptid * 9 + k];
                    for (int k = 0; k < 3; k++) g[k] = sVg[ptid * 9 + 6 + k];
                    double *pt = sPt + ptid * 12;
                    if (mk[ptid] == 0u) {
                        for (int k = 0; k < 9; k++) pt[k] = 0.0;   // unobserved point: contributes nothing
                    } else {
                        if (lin) {
                            D.diag_p[3 * wpp] = fmin(fmax(V[0], 1e-6), 1e32);
                            D.diag_p[3 * wpp + 1] = fmin(fmax(V[3], 1e-6), 1e32);
                            D.diag_p[3 * wpp + 2] = fmin(fmax(V[5], 1e-6), 1e32);
                            const double *sp = D.scale_p + 3 * wpp;
                            gmax = fmax(gmax, fmax(fabs(g[0] / sp[0]), fmax(fabs(g[1] / sp[1]), fabs(g[2] / sp[2]))));
                        }
                        // D^2 = diag / radius (the LM strategy forms sqrt(diag / radius) and squares it again: the same to an ulp)
                        const double a = fma(D.diag_p[3 * wpp], inv_radius, V[0]), b2 = V[1], c = V[2],
                                     d = fma(D.diag_p[3 * wpp + 1], inv_radius, V[3]), e = V[4], f = fma(D.diag_p[3 * wpp + 2], inv_radius, V[5]);
                        const double A0 = d * f - e * e, A1 = c * e - b2 * f, A2 = b2 * e - c * d;
                        const double idet = ba_rcp_fast(a * A0 + b2 * A1 + c * A2);
                        pt[0] = A0 * idet; pt[1] = A1 * idet; pt[2] = A2 * idet;
                        pt[3] = (a * f - c * c) * idet; pt[4] = (b2 * c - a * e) * idet; pt[5] = (a * d - b2 * b2) * idet;
                        pt[6] = g[0]; pt[7] = g[1]; pt[8] = g[2];
                        for (int k = 0; k < 6; k++) D.Vinv[6 * wpp + k] = pt[k];
                        for (int k = 0; k < 3; k++) D.gp[3 * wpp + k] = g[k];
                    }
                }
                producer_barrier();
                // ---- A3: W = Jc^T Jp, Y = W V^-1 into the chunk matrices; camera sums in registers -----------------
                for (int pass = 0; pass < passes; pass++) {
                    const int slot = pslot + pass * H;
                    if (!producer) break;
                    double *Yo = sY + (size_t)(3 * slot) * LD + 6 * pcam, *Wo = sW + (size_t)(3 * slot) * LD + 6 * pcam;
                    if (p0 + slot >= D.Np || !((mk[slot] >> pcam) & 1)) {
#pragma unroll
                        for (int a = 0; a < 6; a++) {
                            Yo[a] = 0.0; Yo[LD + a] = 0.0; Yo[2 * LD + a] = 0.0;
                            Wo[a] = 0.0; Wo[LD + a] = 0.0; Wo[2 * LD + a] = 0.0;
                        }
                        continue;
                    }
                    const double *it = sItem + (size_t)(slot * Nc + pcam) * 20;
                    const double *pt = sPt + slot * 12;
                    const double vg0 = pt[0] * pt[6] + pt[1] * pt[7] + pt[2] * pt[8], vg1 = pt[1] * pt[6] + pt[3] * pt[7] + pt[4] * pt[8],
                                 vg2 = pt[2] * pt[6] + pt[4] * pt[7] + pt[5] * pt[8];
                    double jp[6], jc[12];
#pragma unroll
                    for (int k = 0; k < 6; k++) jp[k] = it[k];
#pragma unroll
                    for (int k = 0; k < 12; k++) jc[k] = it[8 + k];
                    const double r0 = it[6], r1 = it[7];
                    int t = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) {
                        const double w0 = jc[a] * jp[0] + jc[6 + a] * jp[3], w1 = jc[a] * jp[1] + jc[6 + a] * jp[4],
                                     w2 = jc[a] * jp[2] + jc[6 + a] * jp[5];
                        Wo[a] = w0; Wo[LD + a] = w1; Wo[2 * LD + a] = w2;
                        Yo[a] = w0 * pt[0] + w1 * pt[1] + w2 * pt[2];
                        Yo[LD + a] = w0 * pt[1] + w1 * pt[3] + w2 * pt[4];
                        Yo[2 * LD + a] = w0 * pt[2] + w1 * pt[4] + w2 * pt[5];
                        accR[a] -= w0 * vg0 + w1 * vg1 + w2 * vg2;
                        accG[a] += jc[a] * r0 + jc[6 + a] * r1;
#pragma unroll
                        for (int b2 = a; b2 < 6; b2++) accU[t++] += jc[a] * jc[b2] + jc[6 + a] * jc[6 + b2];
                    }
                }
            }
            cta_barrier();
        }
        // reduce camera sums over the point slots
        if (producer) {
            for (int k = 0; k < 21; k++) atomicAdd(&sRed[33 * pcam + k], accU[k]);
            for (int k = 0; k < 6; k++) { atomicAdd(&sRed[33 * pcam + 21 + k], accG[k]); atomicAdd(&sRed[33 * pcam + 27 + k], accR[k]); }
        }
        cost = warp_sum_d(cost);
        if ((tid & 31) == 0 && cost != 0) atomicAdd(&sScal[0], cost);
        if (lin) {
            for (int o = 16; o; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
            if ((tid & 31) == 0 && gmax > 0) atomicMax(reinterpret_cast<unsigned long long *>(&sScal[1]), (unsigned long long)__double_as_longlong(gmax));
        }
        cta_barrier();   // B1: sRed complete, chunk buffers dead
        cta_barrier();   // B2: camera diagonal written
    } else {
        // ---- consumers: 8 warps, each owns up to WS_MAXST 16x16 super-tiles (2x2 fp64 mma tiles) of the upper
        // triangle of S = sum_p Y W^T; accumulators in registers, operands straight from the chunk matrices
        const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
        const int NST = NS * (NS + 1) / 2;
        int stI[WS_MAXST], stJ[WS_MAXST];
#pragma unroll
        for (int j = 0; j < WS_MAXST; j++) {
            int idx = warp + 8 * j, I = 0;
            if (idx < NST) { while (idx >= NS - I) { idx -= NS - I; I++; } stI[j] = I; stJ[j] = I + idx; }
            else { stI[j] = -1; stJ[j] = -1; }
        }
        double C[WS_MAXST][2][2][2];
#pragma unroll
        for (int j = 0; j < WS_MAXST; j++)
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b2 = 0; b2 < 2; b2++) C[j][a][b2][0] = C[j][a][b2][1] = 0.0;
        for (int step = 0; step <= nchunks; step++) {
            if (step > 0) {
                const int buf = (step - 1) & 1;
                const double *sY = sYW + (size_t)buf * 2 * bufsz, *sW = sY + bufsz;
                for (int ks = 0; ks < KS; ks++) {
                    const int kc = 4 * ks + q;
#pragma unroll
                    for (int j = 0; j < WS_MAXST; j++) {
                        if (stI[j] < 0) continue;
                        const double a0 = sY[(size_t)kc * LD + 16 * stI[j] + g], a1 = sY[(size_t)kc * LD + 16 * stI[j] + 8 + g];
                        const double b0 = sW[(size_t)kc * LD + 16 * stJ[j] + g], b1 = sW[(size_t)kc * LD + 16 * stJ[j] + 8 + g];
                        win_dmma(C[j][0][0][0], C[j][0][0][1], a0, b0);
                        win_dmma(C[j][0][1][0], C[j][0][1][1], a0, b1);
                        win_dmma(C[j][1][0][0], C[j][1][0][1], a1, b0);
                        win_dmma(C[j][1][1][0], C[j][1][1][1], a1, b1);
                    }
                }
            }
            cta_barrier();
        }
        cta_barrier();   // B1
        // camera diagonal: LM diagonal, gradient max
        if (tid < Nc && lin) {
            const int dg[6] = {0, 6, 11, 15, 18, 20};
            double gm = 0;
            for (int a = 0; a < 6; a++) {
                D.diag_c[6 * ((size_t)w * Nc + tid) + a] = fmin(fmax(sRed[33 * tid + dg[a]], 1e-6), 1e32);
                gm = fmax(gm, fabs(sRed[33 * tid + 21 + a] / sSc[6 * tid + a]));
            }
            atomicMax(reinterpret_cast<unsigned long long *>(&sScal[1]), (unsigned long long)__double_as_longlong(gm));
        }
        cta_barrier();   // B2
        // S = -sum Y W^T into the lower triangle of A (A[c][r] = S[r][c], r <= c)
#pragma unroll
        for (int j = 0; j < WS_MAXST; j++) {
            if (stI[j] < 0) continue;
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b2 = 0; b2 < 2; b2++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int r = 16 * stI[j] + 8 * a + g, c = 16 * stJ[j] + 8 * b2 + 2 * q + e;
                        if (r <= c && c < n) A[(size_t)c * ld + r] = -C[j][a][b2][e];
                    }
        }
    }
    __syncthreads();
    // camera diagonal blocks: += U_c + D_c^2
    for (int i = tid; i < Nc * 21; i += WS_THREADS) {
        const int c = i / 21, t = i - c * 21;
        int lo = 0, rem = t;
        while (rem >= 6 - lo) { rem -= 6 - lo; lo++; }
        const int hi = lo + rem;
        double v = sRed[33 * c + t];
        if (lo == hi) { const double d = sqrt(D.diag_c[6 * ((size_t)w * Nc + c) + lo] / radius); v += d * d; }
        A[(size_t)(6 * c + hi) * ld + 6 * c + lo] += v;
    }
    __syncthreads();
    if (tid == 0 && lin) {
        st->cost = sScal[0];
        if (st->iter == 0) st->initial_cost = sScal[0];
        st->gmax = sScal[1];
    }
    __shared__ double s_rhs[6 * WIN_MAXC];
    if (tid < n) s_rhs[tid] = sRed[33 * (tid / 6) + 21 + tid % 6] + sRed[33 * (tid / 6) + 27 + tid % 6];
    __shared__ int s_ok;
    if (tid == 0) {
        s_ok = 1;
        if (st->need_linearize && !(sScal[1] > 1e-10)) { st->done = 1; st->termination = 3; s_ok = -1; }
    }
    __syncthreads();
    if (s_ok < 0) return;
    if (tid < n) bvec[tid] = s_rhs[tid];
    __syncthreads();
    // ---- K10: Cholesky + two triangular solves in shared memory ----------------------------------------------
    for (int j = 0; j < n; j++) {
        if (tid == 0) {
            const double d = A[(size_t)j * ld + j];
            if (!(d > 0) || !isfinite(d)) s_ok = 0; else A[(size_t)j * ld + j] = sqrt(d);
        }
        __syncthreads();
        if (!s_ok) break;
        const double dj = A[(size_t)j * ld + j];
        for (int i = j + 1 + tid; i < n; i += WS_THREADS) A[(size_t)i * ld + j] /= dj;
        __syncthreads();
        const int m = n - j - 1;
        for (int t = tid; t < m * m; t += WS_THREADS) {
            const int ii = t / m, kk = t - ii * m;
            if (kk <= ii) { const int i = j + 1 + ii, k = j + 1 + kk; A[(size_t)i * ld + k] -= A[(size_t)i * ld + j] * A[(size_t)k * ld + j]; }
        }
        __syncthreads();
    }
    if (!s_ok) { if (tid == 0) st->chol_ok = 0; return; }
    for (int j = 0; j < n; j++) {
        if (tid == 0) bvec[j] /= A[(size_t)j * ld + j];
        __syncthreads();
        const double bj = bvec[j];
        for (int i = j + 1 + tid; i < n; i += WS_THREADS) bvec[i] -= A[(size_t)i * ld + j] * bj;
        __syncthreads();
    }
    for (int j = n - 1; j >= 0; j--) {
        if (tid == 0) bvec[j] /= A[(size_t)j * ld + j];
        __syncthreads();
        const double bj = bvec[j];
        for (int i = tid; i < j; i += WS_THREADS) bvec[i] -= A[(size_t)j * ld + i] * bj;
        __syncthreads();
    }
    bool fin = true;
    for (int i = tid; i < n; i += WS_THREADS) { D.yc[(size_t)w * n + i] = bvec[i]; fin = fin && isfinite(bvec[i]); }
    const int allfin = __syncthreads_and(fin);
    if (tid == 0) st->chol_ok = allfin ? 1 : 0;
}

// K11 for windows: blockIdx.y = window, each CTA walks a slice of the points; warp per point, lane = camera
__global__ void __launch_bounds__(WIN_THREADS) win_backsub_kernel(const BADev D, const WinDev Wd, int slices)
{
    extern __shared__ double sm[];
    const int w = blockIdx.y, tid = threadIdx.x, Nc = D.Nc;
    BAState *st = &D.st[w];
    if (st->done || !st->chol_ok) return;
    double *sM = sm;                     // Nc*36
    double *sC = sM + Nc * 36;           // Nc*3
    double *sSc = sC + Nc * 3;           // Nc*6
    double *sY = sSc + Nc * 6;           // Nc*6  y_c .* scale_c
    double *sRc = sY + Nc * 6;           // Nc*9  candidate rotation
    double *sCc = sRc + Nc * 9;          // Nc*3  candidate centre
    for (int i = tid; i < Nc * 36; i += WIN_THREADS) sM[i] = Wd.camR[36 * (size_t)w * Nc + i];
    for (int i = tid; i < Nc * 3; i += WIN_THREADS) {
        sC[i] = D.poses[6 * ((size_t)w * Nc + i / 3) + 3 + i % 3];
        sCc[i] = D.cand_poses[6 * ((size_t)w * Nc + i / 3) + 3 + i % 3];
    }
    for (int i = tid; i < Nc * 6; i += WIN_THREADS) {
        sSc[i] = D.scale_c[6 * (size_t)w * Nc + i];
        sY[i] = D.yc[(size_t)w * D.n + i] * sSc[i];
    }
    for (int i = tid; i < Nc * 9; i += WIN_THREADS) sRc[i] = Wd.candR[9 * (size_t)w * Nc + i];
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    const int per = (D.Np + slices - 1) / slices;
    const int pbeg = blockIdx.x * per, pend = min(D.Np, pbeg + per);
    double mc_acc = 0, cc_acc = 0, sn_acc = 0, xn_acc = 0;
    for (int p = pbeg + warp; p < pend; p += WIN_THREADS / 32) {
        const size_t wp = (size_t)w * D.Np + p;
        const unsigned mask = Wd.vis[wp];
        if (mask == 0u) {
            if (lane < 3) D.cand_points[3 * wp + lane] = D.points[3 * wp + lane];
            continue;
        }
        const double sp[3] = {D.scale_p[3 * wp], D.scale_p[3 * wp + 1], D.scale_p[3 * wp + 2]};
        const double X[3] = {D.points[3 * wp], D.points[3 * wp + 1], D.points[3 * wp + 2]};
        const bool vis = lane < Nc && ((mask >> lane) & 1);
        double r[2] = {0, 0}, jc[12], jp[6], ox = 0, oy = 0;
        double t0 = 0, t1 = 0, t2 = 0, jy0 = 0, jy1 = 0;
        if (vis) {
            const int i = D.pt_off[wp] + __popc(mask & ((1u << lane) - 1));
            ox = D.obs_xy[2 * (size_t)i]; oy = D.obs_xy[2 * (size_t)i + 1];
            win_residual_jac(sM + 36 * lane, sC + 3 * lane, X, ox, oy, D.fx, D.cx, D.fy, D.cy, r, jc, jp);
            double rho0, rho1;
            ba_huber(D.delta, r[0] * r[0] + r[1] * r[1], rho0, rho1);
            const double sr = sqrt(rho1);
            r[0] *= sr; r[1] *= sr;
#pragma unroll
            for (int k = 0; k < 6; k++) { jc[k] *= sr; jc[6 + k] *= sr; jy0 += jc[k] * sY[6 * lane + k]; jy1 += jc[6 + k] * sY[6 * lane + k]; }
#pragma unroll
            for (int k = 0; k < 3; k++) { jp[k] *= sr * sp[k]; jp[3 + k] *= sr * sp[k]; }
            t0 = -(jp[0] * jy0 + jp[3] * jy1); t1 = -(jp[1] * jy0 + jp[4] * jy1); t2 = -(jp[2] * jy0 + jp[5] * jy1);
        }
        t0 = warp_sum_d(t0) + D.gp[3 * wp]; t1 = warp_sum_d(t1) + D.gp[3 * wp + 1]; t2 = warp_sum_d(t2) + D.gp[3 * wp + 2];
        const double *Vi = D.Vinv + 6 * wp;
        const double yp[3] = {Vi[0] * t0 + Vi[1] * t1 + Vi[2] * t2, Vi[1] * t0 + Vi[3] * t1 + Vi[4] * t2,
                              Vi[2] * t0 + Vi[4] * t1 + Vi[5] * t2};
        double cand[3];
#pragma unroll
        for (int k = 0; k < 3; k++) cand[k] = X[k] + (-yp[k] * sp[k]);
        if (lane < 3) D.cand_points[3 * wp + lane] = cand[lane];
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 3; k++) { const double d = X[k] - cand[k]; sn_acc += d * d; xn_acc += X[k] * X[k]; }
        }
        if (vis) {
            double m0 = -jy0, m1 = -jy1;
#pragma unroll
            for (int k = 0; k < 3; k++) { m0 -= jp[k] * yp[k]; m1 -= jp[3 + k] * yp[k]; }
            mc_acc -= m0 * (r[0] + m0 / 2.0) + m1 * (r[1] + m1 / 2.0);
            // candidate residual with the candidate rotation matrix
            const double *Rc = sRc + 9 * lane;
            const double q0 = cand[0] + sCc[3 * lane], q1 = cand[1] + sCc[3 * lane + 1], q2 = cand[2] + sCc[3 * lane + 2];
            const double px = Rc[0] * q0 + Rc[1] * q1 + Rc[2] * q2, py = Rc[3] * q0 + Rc[4] * q1 + Rc[5] * q2,
                         pz = (Rc[6] * q0 + Rc[7] * q1 + Rc[8] * q2) * -1.0;
            const double e0 = ox - (px / pz * D.fx + D.cx), e1 = oy - (py / pz * D.fy + D.cy);
            double rho0, rho1;
            ba_huber(D.delta, e0 * e0 + e1 * e1, rho0, rho1);
            cc_acc += 0.5 * rho0;
        }
    }
    mc_acc = warp_sum_d(mc_acc); cc_acc = warp_sum_d(cc_acc);
    if (lane == 0) {
        atomicAdd(&st->model_change, mc_acc);
        atomicAdd(&st->cand_cost, cc_acc);
        atomicAdd(&st->step_norm2, sn_acc);
        atomicAdd(&st->x_norm2, xn_acc);
    }
}

}  // namespace

// ------------------------------------------------------------------ host side -------------------------
struct pmv_ba_window_ws {
    unsigned *vis = nullptr;
    double *camR = nullptr, *candR = nullptr;
};

bool pmv_internal_ba_window_eligible(int Nc, int Np) { return Nc <= WIN_MAXC && Nc >= 1 && Np >= 1; }

size_t pmv_internal_ba_window_bytes(int W, int Nc, int Np)
{
    return sizeof(unsigned) * (size_t)W * Np + sizeof(double) * (size_t)W * Nc * 45;
}

int pmv_internal_ba_window_iteration(pmv_ctx *ctx, const BADev &D, const unsigned *d_vis, double *d_camR, double *d_candR,
                                     cudaStream_t s)
{
    WinDev Wd;
    Wd.vis = d_vis; Wd.camR = d_camR; Wd.candR = d_candR;
    const int Nc = D.Nc, n = D.n;
    Wd.P = 12;                                  // 3 P = 36 columns = 9 mma k-steps, LD = 36
    Wd.passes = (Wd.P * Nc > WS_PRODUCERS) ? 2 : 1;
    const int wc = D.W * Nc;
    const size_t smem_col = sizeof(double) * (size_t)Nc * 45;
    const int NS = (n + 15) / 16;
    Wd.LD = 16 * NS + 4;
    const int KS = (3 * Wd.P + 3) / 4;
    const size_t unionA = (size_t)Wd.P * Nc * 20 + (size_t)Wd.P * 21 + 4 * (size_t)4 * KS * Wd.LD + Wd.P + 2;
    const size_t unionB = (size_t)n * (n + 1) + n;
    const size_t smem_schur = sizeof(double) * ((size_t)Nc * 78 + 8 + std::max(unionA, unionB) + 2);
    const size_t smem_back = sizeof(double) * (size_t)Nc * 63;
    if (ctx->attr_first(PMV_ATTR_WIN_SCHUR)) {
        cudaFuncSetAttribute(win_schur_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(win_schur_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    }
    win_prepare_kernel<<<(wc + 127) / 128, 128, 0, s>>>(D, Wd, 0);
    PMV_LAUNCH_CHECK(ctx, "win_prepare_kernel");
    win_colnorm_kernel<<<D.W, WIN_THREADS, smem_col, s>>>(D, Wd);
    PMV_LAUNCH_CHECK(ctx, "win_colnorm_kernel");
    if (NS * (NS + 1) / 2 <= 40) win_schur_kernel<5><<<D.W, WS_THREADS, smem_schur, s>>>(D, Wd);
    else win_schur_kernel<6><<<D.W, WS_THREADS, smem_schur, s>>>(D, Wd);
    PMV_LAUNCH_CHECK(ctx, "win_schur_kernel");
    ba_cam_candidate_kernel<<<(wc + 127) / 128, 128, 0, s>>>(D, 1);
    PMV_LAUNCH_CHECK(ctx, "ba_cam_candidate_kernel");
    win_prepare_kernel<<<(wc + 127) / 128, 128, 0, s>>>(D, Wd, 1);
    PMV_LAUNCH_CHECK(ctx, "win_prepare_kernel");
    {
        // enough CTAs to fill the machine: slices per window so that W * slices >= ~4 waves of 148 SMs
        int slices = std::max(1, std::min((148 * 8 + D.W - 1) / D.W, (D.Np + 63) / 64));
        dim3 grid(slices, D.W);
        win_backsub_kernel<<<grid, WIN_THREADS, smem_back, s>>>(D, Wd, slices);
        PMV_LAUNCH_CHECK(ctx, "win_backsub_kernel");
    }
    ba_lm_update_kernel<<<(D.W + 127) / 128, 128, 0, s>>>(D);
    PMV_LAUNCH_CHECK(ctx, "ba_lm_update_kernel");
    {
        size_t tot = (size_t)wc * 6 + (size_t)D.W * D.Np * 3;
        ba_accept_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(D);
        PMV_LAUNCH_CHECK(ctx, "ba_accept_kernel");
    }
    return PMV_OK;
}

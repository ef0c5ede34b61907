// ba.cu -- host side of the bundle adjuster: problem upload, the launch sequence of one
// Levenberg-Marquardt iteration (no host synchronisation inside a solve) and the C ABI.
//
// Drop-in for CeresBundleAdjustment::apply (reference CeresBundleAdjustment.cpp:5-89): the adapter
// marshals tracker->R/t -> pose blocks [rodrigues(R^T), -t] (:26-34), Feature3D -> point blocks (:47-48),
// (column,row) -> observations (:45) and calls pmv_ba_solve with huber_delta 1.0 and
// max_iters = tracker->ba_iterations (:54-61).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <numeric>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "ba.cuh"
#include "ba_kernels.cuh"
#include "ba_runs.cuh"

int pmv_internal_ba_cholesky_large(pmv_ctx *ctx, const BADev &D, const int *lim_host, const BASplit *split, const BAPart *part,
                                   cudaStream_t s);  // ba_chol.cu
int pmv_internal_ba_cholesky_band_T(int n, const int *lim_host);                                    // ba_chol_band.cu
int pmv_internal_ba_allreduce(pmv_ctx *ctx, const double *send, double *recv, size_t count, int op_max,
                              cudaStream_t s);                                      // ba_nccl.cu
bool pmv_internal_ba_window_eligible(int Nc, int Np);                               // ba_window.cu
int pmv_internal_ba_window_iteration(pmv_ctx *ctx, const BADev &D, const unsigned *d_vis, double *d_camR,
                                     double *d_candR, cudaStream_t s);

struct pmv_ba_problem {
    pmv_ctx *ctx = nullptr;
    BADev D{};
    int sharded = 0;     // points sharded over ranks: reduce camera blocks / S / scalars with NCCL
    int rank = 0;
    std::vector<void *> allocs;
    // one-shot solves (pmv_ba_solve / pmv_ba_solve_batched) carve their buffers out of a grow-only arena of
    // the context instead of ~35 cudaMalloc/cudaFree pairs per call (the pipeline calls BA once per keyframe)
    int transient = 0;           // created by a one-shot entry point
    int arena_mode = 0;          // 0 cudaMalloc per buffer, 1 measuring pass, 2 carving pass
    char *arena_base = nullptr;
    size_t arena_off = 0;
    double *d_init_poses = nullptr, *d_init_points = nullptr;
    double *d_Uraw = nullptr, *d_Uraw_red = nullptr;   // 27 doubles per camera (+ W cost slots at the end)
    double *d_scal = nullptr, *d_scal_red = nullptr;   // per window: model_change, cand_cost, step_norm2, x_norm2
    std::vector<int> chol_lim;                           // envelope of S per PMV_CHOL_NB-row block (host copy)
    // one LM iteration is a fixed launch sequence (all decisions live in BAState on the device), so it is
    // captured once into a CUDA graph and replayed: removes the launch gaps of the ~600 dependent launches
    // of the blocked Cholesky
    double *d_band = nullptr;                            // packed envelope of S + rhs (sharded all-reduce buffer)
    long long *d_band_off = nullptr;
    long long band_count = 0;
    cudaGraphExec_t graph_exec = nullptr;
    int graph_max_iters = -1;
    uint64_t graph_launches = 0;                         // kernel launches inside one replay
    size_t bytes = 0;
    // block-organised point elimination (ba_pair_schur_kernel): segments of the per-camera-pair entry lists
    BAPairSeg *d_segs = nullptr;
    int2 *d_entries = nullptr;
    int nsegs = 0;
    // run-organised path of one large problem (ba_runs.cuh): points sorted by camera tuple, runs of equal tuples
    int *d_run_off = nullptr;
    int2 *d_run_pt = nullptr;                           // (point, first observation of the point) per run position
    int nruns = 0;
    RunSide run_side[2];                                // side stream + fork / join events of the elimination / the back-substitution
    int run_kbegin[RUN_MAXK + 2] = {};                  // runs are ordered by tuple size: [kbegin[k], kbegin[k+1]) have k observations
    // two-sided solve of the banded reduced camera system (BASplit, ba.cuh)
    BASplit split;
    std::vector<int> split_lim[3];
    // partitioned solve (BAPart, ba.cuh): P segments, P - 1 separators
    BAPart part;
    std::vector<int> part_lim[PMV_PART_MAX + 1];
    // window-batched path (ba_window.cu): Nc <= 22, every (point, camera) pair observed at most once
    int use_window = 0;
    unsigned *d_vis = nullptr;
    double *d_camR = nullptr, *d_candR = nullptr;
};

namespace {

// host array WITHOUT value initialisation (std::vector would memset gigabytes on one core before the parallel loops
// ever touch them)
template <typename T>
struct RawBuf {
    T *p = nullptr;
    RawBuf() = default;
    explicit RawBuf(size_t n) { alloc(n); }
    void alloc(size_t n) { free(p); p = static_cast<T *>(malloc(std::max<size_t>(n, 1) * sizeof(T))); }
    ~RawBuf() { free(p); }
    RawBuf(const RawBuf &) = delete;
    RawBuf &operator=(const RawBuf &) = delete;
    T *data() const { return p; }
    T &operator[](size_t i) const { return p[i]; }
};

template <typename T>
int dev_alloc(pmv_ba_problem *p, T **out, size_t count)
{
    void *q = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    if (p->arena_mode) {
        const size_t off = (p->arena_off + 255) & ~(size_t)255;
        p->arena_off = off + bytes;
        if (p->arena_mode == 2) { *out = reinterpret_cast<T *>(p->arena_base + off); p->bytes += bytes; }
        return PMV_OK;
    }
    // stream-ordered allocation out of the device's default pool (pmv_create keeps freed blocks cached in it): a
    // problem created after another one was destroyed pays microseconds instead of the page-mapping cost of cudaMalloc
    cudaError_t e = cudaMallocAsync(&q, bytes, p->ctx->stream);
    if (e != cudaSuccess) return p->ctx->fail(PMV_ERR_NOMEM, "ba problem allocation", e);
    p->allocs.push_back(q);
    p->bytes += bytes;
    *out = reinterpret_cast<T *>(q);
    return PMV_OK;
}

__global__ void ba_state_init_kernel(BAState *st, int W)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    BAState s;
    memset(&s, 0, sizeof s);
    s.radius = 1e4; s.decrease_factor = 2.0; s.need_linearize = 1; s.chol_ok = 0;
    st[w] = s;
}

// pack / unpack the per-window scalars that a sharded solve must sum across ranks
__global__ void ba_pack_cost_kernel(const BADev D, double *Uraw_tail)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < D.W) Uraw_tail[w] = D.st[w].new_cost;
}
__global__ void ba_unpack_cost_kernel(const BADev D, const double *Uraw_tail)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < D.W) D.st[w].new_cost = Uraw_tail[w];
}
__global__ void ba_pack_scalars_kernel(const BADev D, double *buf)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= D.W) return;
    const BAState *st = &D.st[w];
    buf[4 * w] = st->model_change; buf[4 * w + 1] = st->cand_cost; buf[4 * w + 2] = st->step_norm2; buf[4 * w + 3] = st->x_norm2;
}
__global__ void ba_unpack_scalars_kernel(const BADev D, const double *buf)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= D.W) return;
    BAState *st = &D.st[w];
    st->model_change = buf[4 * w]; st->cand_cost = buf[4 * w + 1]; st->step_norm2 = buf[4 * w + 2]; st->x_norm2 = buf[4 * w + 3];
}
__global__ void ba_pack_gmax_kernel(const BADev D, double *buf)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < D.W) buf[w] = D.st[w].gmax;
}
__global__ void ba_unpack_gmax_kernel(const BADev D, const double *buf)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < D.W) D.st[w].gmax = buf[w];
}

// Sharded solves exchange only the envelope of the upper triangle of S (plus rhs): row r contributes the
// columns [r, lim(r)), packed back to back -- for the banded BAL-scale system 13 MB instead of 288 MB.
__global__ void __launch_bounds__(256) ba_band_pack_kernel(const double *__restrict__ S, const double *__restrict__ rhs, int n,
                                                           const int *__restrict__ lim, const long long *__restrict__ off,
                                                           double *__restrict__ buf, int unpack, double *S_out, double *rhs_out)
{
    const int r = blockIdx.x;
    if (r == n) {   // last block: the right-hand side
        for (int j = threadIdx.x; j < n; j += 256) {
            if (unpack) rhs_out[j] = buf[off[n] + j]; else buf[off[n] + j] = rhs[j];
        }
        return;
    }
    const int len = lim[r / PMV_CHOL_NB] - r;
    const long long o = off[r];
    for (int j = threadIdx.x; j < len; j += 256) {
        if (unpack) S_out[(size_t)r * n + r + j] = buf[o + j]; else buf[o + j] = S[(size_t)r * n + r + j];
    }
}

int ba_iteration(pmv_ba_problem *p, cudaStream_t s)
{
    pmv_ctx *ctx = p->ctx;
    const BADev &D = p->D;
    const int W = D.W, n = D.n;
    const int wc = W * D.Nc, wp = W * D.Np;
    const int wblocks = (W + 127) / 128;
    if (p->nruns > 0) {
        // runs of equal camera tuples: cost + raw camera blocks straight from the observations (nothing materialised)
        PMV_CUDA_TRY(ctx, cudaMemsetAsync(p->d_Uraw, 0, sizeof(double) * 27 * (size_t)wc, s));
        ba_cam_trig_kernel<<<(wc + 127) / 128, 128, 0, s>>>(D.poses, D.trig, wc, D.st, 0);
        PMV_LAUNCH_CHECK(ctx, "ba_cam_trig_kernel");
        if (run_minb(1) >= 4) ba_run_cam_kernel<4><<<(p->nruns + 3) / 4, 128, 0, s>>>(D, p->d_run_off, p->d_run_pt, p->nruns, p->d_Uraw);
        else if (run_minb(1) == 3) ba_run_cam_kernel<3><<<(p->nruns + 3) / 4, 128, 0, s>>>(D, p->d_run_off, p->d_run_pt, p->nruns, p->d_Uraw);
        else ba_run_cam_kernel<2><<<(p->nruns + 3) / 4, 128, 0, s>>>(D, p->d_run_off, p->d_run_pt, p->nruns, p->d_Uraw);
        PMV_LAUNCH_CHECK(ctx, "ba_run_cam_kernel");
    } else if (D.No > 0) {
        ba_linearize_kernel<<<(D.No + 127) / 128, 128, 0, s>>>(D);
        PMV_LAUNCH_CHECK(ctx, "ba_linearize_kernel");
    }
    if (p->nruns > 0) {
    } else if (D.No >= 256LL * wc) {   // hundreds to thousands of observations per camera: a CTA per camera
        ba_cam_accumulate_wide_kernel<<<wc, 256, 0, s>>>(D, p->d_Uraw);
        PMV_LAUNCH_CHECK(ctx, "ba_cam_accumulate_wide_kernel");
    } else {
        ba_cam_accumulate_kernel<<<(wc + 3) / 4, 128, 0, s>>>(D, p->d_Uraw);
        PMV_LAUNCH_CHECK(ctx, "ba_cam_accumulate_kernel");
    }
    const double *Uraw = p->d_Uraw;
    if (p->sharded) {
        ba_pack_cost_kernel<<<wblocks, 128, 0, s>>>(D, p->d_Uraw + 27 * (size_t)wc);
        PMV_LAUNCH_CHECK(ctx, "ba_pack_cost_kernel");
        int rc = pmv_internal_ba_allreduce(ctx, p->d_Uraw, p->d_Uraw_red, 27 * (size_t)wc + W, 0, s);
        if (rc) return rc;
        ba_unpack_cost_kernel<<<wblocks, 128, 0, s>>>(D, p->d_Uraw_red + 27 * (size_t)wc);
        PMV_LAUNCH_CHECK(ctx, "ba_unpack_cost_kernel");
        Uraw = p->d_Uraw_red;
    }
    ba_latch_cost_kernel<<<wblocks, 128, 0, s>>>(D);
    PMV_LAUNCH_CHECK(ctx, "ba_latch_cost_kernel");
    ba_cam_finalize_kernel<<<(wc + 127) / 128, 128, 0, s>>>(D, Uraw);
    PMV_LAUNCH_CHECK(ctx, "ba_cam_finalize_kernel");
    {
        size_t nn = (size_t)n * n;
        dim3 grid((unsigned)std::min<size_t>((nn + 255) / 256, 1024), W);
        ba_clear_system_kernel<<<grid, 256, 0, s>>>(D);
        PMV_LAUNCH_CHECK(ctx, "ba_clear_system_kernel");
    }
    if (wp > 0 && p->nruns > 0) {
        int rc = launch_run_schur_all(ctx, D, p->d_run_off, p->d_run_pt, p->run_kbegin, s, p->run_side[0]);
        if (rc) return rc;
    } else if (wp > 0) {
        const int by_pairs = p->nsegs > 0;
        if (by_pairs && D.No <= 8 * (long long)D.Np) {
            ba_point_vinv_w1_kernel<8><<<std::min((wp + 15) / 16, 148 * 32), 128, 0, s>>>(D);
            PMV_LAUNCH_CHECK(ctx, "ba_point_vinv_w1_kernel");
        } else {
            ba_point_schur_kernel<<<std::min((wp + 3) / 4, 148 * 32), 128, 0, s>>>(D, by_pairs ? 0 : 1);
            PMV_LAUNCH_CHECK(ctx, "ba_point_schur_kernel");
        }
        if (by_pairs) {
            ba_pair_schur_kernel<<<(p->nsegs + 3) / 4, 128, 0, s>>>(D, p->d_segs, p->nsegs, p->d_entries);
            PMV_LAUNCH_CHECK(ctx, "ba_pair_schur_kernel");
        }
    }
    if (p->sharded) {
        // sum the partial reduced camera systems over the ranks: envelope of the upper triangle + rhs
        ba_band_pack_kernel<<<n + 1, 256, 0, s>>>(D.S, D.rhs, n, D.chol_lim, p->d_band_off, p->d_band, 0, nullptr, nullptr);
        PMV_LAUNCH_CHECK(ctx, "ba_band_pack_kernel");
        int rc = pmv_internal_ba_allreduce(ctx, p->d_band, p->d_band, (size_t)p->band_count, 0, s);
        if (rc) return rc;
        ba_band_pack_kernel<<<n + 1, 256, 0, s>>>(nullptr, nullptr, n, D.chol_lim, p->d_band_off, p->d_band, 1, D.S, D.rhs);
        PMV_LAUNCH_CHECK(ctx, "ba_band_pack_kernel");
        ba_pack_gmax_kernel<<<wblocks, 128, 0, s>>>(D, p->d_scal);
        PMV_LAUNCH_CHECK(ctx, "ba_pack_gmax_kernel");
        rc = pmv_internal_ba_allreduce(ctx, p->d_scal, p->d_scal_red, W, 1, s);
        if (rc) return rc;
        ba_unpack_gmax_kernel<<<wblocks, 128, 0, s>>>(D, p->d_scal_red);
        PMV_LAUNCH_CHECK(ctx, "ba_unpack_gmax_kernel");
    }
    ba_add_cam_blocks_kernel<<<(wc + 127) / 128, 128, 0, s>>>(D);
    PMV_LAUNCH_CHECK(ctx, "ba_add_cam_blocks_kernel");
    if (n <= 160) {
        size_t smem = ((size_t)n * (n + 1) + n) * sizeof(double);
        ba_cholesky_small_kernel<<<W, 256, smem, s>>>(D);
        PMV_LAUNCH_CHECK(ctx, "ba_cholesky_small_kernel");
    } else {
        int rc = pmv_internal_ba_cholesky_large(ctx, D, p->chol_lim.empty() ? nullptr : p->chol_lim.data(), &p->split, &p->part, s);
        if (rc) return rc;
    }
    ba_cam_candidate_kernel<<<(wc + 127) / 128, 128, 0, s>>>(D, p->sharded ? (p->rank == 0) : 1);
    PMV_LAUNCH_CHECK(ctx, "ba_cam_candidate_kernel");
    if (D.cand_trig) {
        ba_cam_trig_kernel<<<(wc + 127) / 128, 128, 0, s>>>(D.cand_poses, D.cand_trig, wc, D.st, 1);
        PMV_LAUNCH_CHECK(ctx, "ba_cam_trig_kernel");
    }
    if (wp > 0 && p->nruns > 0) {
        int rc = launch_run_backsub_all(ctx, D, p->d_run_off, p->d_run_pt, p->run_kbegin, s, p->run_side[1]);
        if (rc) return rc;
    } else if (wp > 0) {
        if ((W == 1 && (D.No >= 100000 || p->nsegs > 0) && D.No <= 8 * (long long)D.Np)) {
            ba_backsub_w1_kernel<8, 3><<<std::min((wp + 15) / 16, 148 * 32), 128, 0, s>>>(D);
            PMV_LAUNCH_CHECK(ctx, "ba_backsub_w1_kernel");
        } else {
            ba_backsub_kernel<<<std::min((wp + 3) / 4, 148 * 32), 128, 0, s>>>(D);
            PMV_LAUNCH_CHECK(ctx, "ba_backsub_kernel");
        }
    }
    if (p->sharded) {
        ba_pack_scalars_kernel<<<wblocks, 128, 0, s>>>(D, p->d_scal);
        PMV_LAUNCH_CHECK(ctx, "ba_pack_scalars_kernel");
        int rc = pmv_internal_ba_allreduce(ctx, p->d_scal, p->d_scal_red, 4 * (size_t)W, 0, s);
        if (rc) return rc;
        ba_unpack_scalars_kernel<<<wblocks, 128, 0, s>>>(D, p->d_scal_red);
        PMV_LAUNCH_CHECK(ctx, "ba_unpack_scalars_kernel");
    }
    ba_lm_update_kernel<<<wblocks, 128, 0, s>>>(D);
    PMV_LAUNCH_CHECK(ctx, "ba_lm_update_kernel");
    {
        size_t tot = (size_t)wc * 6 + (size_t)wp * 3;
        ba_accept_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(D);
        PMV_LAUNCH_CHECK(ctx, "ba_accept_kernel");
    }
    return PMV_OK;
}

// ---- host side of problem creation: the caller's observation list in DEVICE ORDER ---------------------------------
// Device order = sorted by (window, point, camera, original index).  Three routes: the list is already in that order
// (taken in place), an unordered list of many windows (every window sorted on its own), anything else (atomic
// scatter + per-point sorts).  No CUDA in here: pmv_ba_index_observations exposes it to the CPU tests.
// (OpenMP over windows / observations / points: 136 M observations at BASELINE config 4 took 5.5 s on one core)
struct BAIndex {
    // inputs (caller's arrays)
    const double *obs = nullptr;
    const int32_t *cam_idx = nullptr, *pt_idx = nullptr, *obs_off = nullptr;
    int W = 0, Nc = 0, Np = 0, No = 0;
    // results
    std::vector<int> pt_off, cam_off;       // W*Np + 1, W*Nc + 1
    int in_order = 0;                       // the caller's list was already in device order
    bool sorted_per_window = false;         // unordered list of many windows: every window was sorted on its own
    // device-order arrays: the caller's own when the list is in order (no copies: 4 GB at BASELINE config 4), else sorted copies
    const int32_t *Hcam = nullptr, *Hpt = nullptr, *Hwin = nullptr;
    const double *Hobs = nullptr;
    RawBuf<int> win, h_cam, h_pt, h_win, cam_obs;
    RawBuf<double> h_obs;
    bool win_filled = false;

    // window of every observation: only the unordered path and the general path's kernels need the array, so it is
    // filled on demand (touching 572 MB of fresh pages cost more than the whole indexing pass at BASELINE config 4)
    void fill_win()
    {
        if (win_filled) return;
        win_filled = true;
#pragma omp parallel for schedule(static)
        for (int w = 0; w < W; w++) {
            const int o0 = obs_off ? obs_off[w] : 0, o1 = obs_off ? obs_off[w + 1] : No;
            for (int i = o0; i < o1; i++) win[i] = w;
        }
    }

    // point and window of every observation of a list sorted window by window: only the general path reads them
    void fill_pt_win_sorted()
    {
        if (!sorted_per_window || h_pt.data()) return;
        h_pt.alloc(No); h_win.alloc(No);
        const long long nq = (long long)W * Np;
#pragma omp parallel for schedule(static)
        for (long long u = 0; u < nq; u++) {
            const int wq = (int)(u / Np), pq = (int)(u - (long long)wq * Np);
            for (int d = pt_off[u]; d < pt_off[u + 1]; d++) { h_pt[d] = pq; h_win[d] = wq; }
        }
        Hpt = h_pt.data(); Hwin = h_win.data();
    }

    // observations of every (window, camera) in increasing device order: only the general path's camera kernels read
    // this list (ba_cam_accumulate*), so it is built once the path is known
    void build_cam_obs()
    {
        fill_win(); fill_pt_win_sorted();
        cam_obs.alloc(No);
        std::vector<int> cpos(cam_off.begin(), cam_off.end() - 1);
#pragma omp parallel for schedule(static)
        for (int d = 0; d < No; d++) {
            int *cp = &cpos[(size_t)Hwin[d] * Nc + Hcam[d]];
            int k;
#pragma omp atomic capture
            { k = *cp; (*cp)++; }
            cam_obs[k] = d;
        }
        const long long nc = (long long)cam_off.size() - 1;
#pragma omp parallel for schedule(dynamic, 64)
        for (long long c = 0; c < nc; c++)
            if (cam_off[c + 1] - cam_off[c] > 1) std::sort(cam_obs.data() + cam_off[c], cam_obs.data() + cam_off[c + 1]);
    }

    // returns PMV_OK or PMV_ERR_INVALID with *err set
    int build(const char **err)
    {
        win.alloc(No);
        for (int w = 0; w < W; w++) {
            int o0 = obs_off ? obs_off[w] : 0, o1 = obs_off ? obs_off[w + 1] : No;
            if (o0 < 0 || o1 < o0 || o1 > No) { *err = "pmv_ba_problem_create: bad obs_off"; return PMV_ERR_INVALID; }
        }
        if (obs_off && (obs_off[0] != 0 || obs_off[W] != No)) {
            *err = "pmv_ba_problem_create: obs_off must cover the observation list (obs_off[0] == 0, obs_off[W] == No)";
            return PMV_ERR_INVALID;
        }
        // ONE streaming pass over the caller's list (143.7 M observations at BASELINE config 4: five separate passes took 0.38 s):
        // window of every observation, range check, "already in device order?" -- sorted by (window, point, camera), which is
        // how structure-from-motion exports (and BAL files) come --, and, optimistically for that case, the point offsets
        // (boundaries of the list) and thread-private per-camera histograms.  In order: the permutation is the identity and
        // nothing has to be sorted or copied.  Otherwise the offsets are recounted below and the list is sorted.
        pt_off.assign((size_t)W * Np + 1, 0); cam_off.assign((size_t)W * Nc + 1, 0);
        int in_order_l = 1;
        sorted_per_window = false;
        {
            const long long nq = (long long)W * Np;
            const size_t bins = (size_t)W * Nc;
            int T = 1;
    #ifdef _OPENMP
            T = std::max(1, omp_get_max_threads());
    #endif
            std::vector<int> hist((size_t)T * bins, 0);
            int bad = 0, nthreads_used = 1;
    #pragma omp parallel num_threads(T) reduction(| : bad) reduction(& : in_order_l)
            {
                int t = 0, nt = 1;
    #ifdef _OPENMP
                t = omp_get_thread_num(); nt = omp_get_num_threads();
    #endif
                if (t == 0) nthreads_used = nt;
                int *hcnt = hist.data() + (size_t)t * bins;
                const int i0 = (int)((long long)No * t / nt), i1 = (int)((long long)No * (t + 1) / nt);
                int w = 0;
                if (obs_off && i0 < i1) w = (int)(std::upper_bound(obs_off, obs_off + W + 1, i0) - obs_off) - 1;
                long long kp = -1;          // key of the previous observation (the chunk's predecessor for its first one)
                int cprev = -1;
                if (i0 > 0 && i0 < i1) {
                    int wq = 0;
                    if (obs_off) wq = (int)(std::upper_bound(obs_off, obs_off + W + 1, i0 - 1) - obs_off) - 1;
                    kp = (long long)wq * Np + pt_idx[i0 - 1]; cprev = cam_idx[i0 - 1];
                }
                for (int i = i0; i < i1; i++) {
                    if (obs_off) while (i >= obs_off[w + 1]) w++;
                    const int c = cam_idx[i], q = pt_idx[i];
                    if (c < 0 || c >= Nc || q < 0 || q >= Np) { bad = 1; kp = nq; continue; }
                    hcnt[(size_t)w * Nc + c]++;
                    const long long k = (long long)w * Np + q;
                    if (!(kp < k || (kp == k && cprev <= c))) in_order_l = 0;
                    for (long long u = std::max(kp + 1, 0ll); u <= k; u++) pt_off[u] = i;    // pt_off[u] = first observation whose key is >= u
                    kp = k; cprev = c;
                }
            }
            if (bad) {
                *err = "pmv_ba_problem_create: observation index out of range";
                return PMV_ERR_INVALID;
            }
            in_order = in_order_l;
            if (in_order) {
                long long last = -1;
                if (No) {
                    const int wl = obs_off ? (int)(std::upper_bound(obs_off, obs_off + W + 1, No - 1) - obs_off) - 1 : 0;
                    last = (long long)wl * Np + pt_idx[No - 1];
                }
    #pragma omp parallel for schedule(static)
                for (long long u = last + 1; u <= nq; u++) pt_off[u] = No;
    #pragma omp parallel for schedule(static)
                for (long long b = 0; b < (long long)bins; b++) {
                    int c = 0;
                    for (int t = 0; t < nthreads_used; t++) c += hist[(size_t)t * bins + b];
                    cam_off[b + 1] = c;
                }
                std::partial_sum(cam_off.begin(), cam_off.end(), cam_off.begin());
            } else if (W >= 64) {
                // Unordered list of many windows -- e.g. camera-major, the order CeresBundleAdjustment::apply adds its residual
                // blocks in (frame by frame, CeresBundleAdjustment.cpp:27-52): a window's observations are contiguous in both
                // orders, so every window is sorted on its own by one thread with a stable counting sort by point (original
                // order inside a point = ascending camera for camera-major input; a point whose cameras do not come out
                // ascending is insertion-sorted).  No atomics, no per-point std::sort: 0.4 -> 0.15 s at 143.7 M observations.
                sorted_per_window = true;
                h_cam.alloc(No); h_obs.alloc(2 * (size_t)No);     // point / window of an observation: filled on demand (fill_pt_win_sorted)
    #pragma omp parallel
                {
                    std::vector<int> pos(Np + 1);
    #pragma omp for schedule(dynamic, 4)
                    for (int w = 0; w < W; w++) {
                        const int o0 = obs_off[w], o1 = obs_off[w + 1];
                        std::fill(pos.begin(), pos.end(), 0);
                        int *ccnt = &cam_off[(size_t)w * Nc + 1];
                        for (int i = o0; i < o1; i++) { pos[pt_idx[i] + 1]++; ccnt[cam_idx[i]]++; }
                        for (int q = 0; q < Np; q++) pos[q + 1] += pos[q];
                        int *po = &pt_off[(size_t)w * Np];
                        for (int q = 0; q < Np; q++) po[q] = o0 + pos[q];
                        for (int i = o0; i < o1; i++) {
                            const int q = pt_idx[i], d = o0 + pos[q]++;
                            h_cam[d] = cam_idx[i];
                            h_obs[2 * (size_t)d] = obs[2 * (size_t)i]; h_obs[2 * (size_t)d + 1] = obs[2 * (size_t)i + 1];
                        }
                        for (int q = 0; q < Np; q++) {          // cameras of a point ascending (stable): already so for camera-major input
                            const int a = po[q], b = o0 + pos[q];
                            for (int d = a + 1; d < b; d++) {
                                if (h_cam[d - 1] <= h_cam[d]) continue;
                                const int c = h_cam[d];
                                const double ox = h_obs[2 * (size_t)d], oy = h_obs[2 * (size_t)d + 1];
                                int e = d;
                                while (e > a && h_cam[e - 1] > c) {
                                    h_cam[e] = h_cam[e - 1]; h_obs[2 * (size_t)e] = h_obs[2 * (size_t)e - 2]; h_obs[2 * (size_t)e + 1] = h_obs[2 * (size_t)e - 1];
                                    e--;
                                }
                                h_cam[e] = c; h_obs[2 * (size_t)e] = ox; h_obs[2 * (size_t)e + 1] = oy;
                            }
                        }
                    }
                }
                pt_off[(size_t)W * Np] = No;
                std::partial_sum(cam_off.begin(), cam_off.end(), cam_off.begin());
            } else {
                std::fill(pt_off.begin(), pt_off.end(), 0);
                fill_win();
    #pragma omp parallel for schedule(static)
                for (int i = 0; i < No; i++) {
                    int *pc = &pt_off[(size_t)win[i] * Np + pt_idx[i] + 1], *cc = &cam_off[(size_t)win[i] * Nc + cam_idx[i] + 1];
    #pragma omp atomic
                    (*pc)++;
    #pragma omp atomic
                    (*cc)++;
                }
                std::partial_sum(pt_off.begin(), pt_off.end(), pt_off.begin());
                std::partial_sum(cam_off.begin(), cam_off.end(), cam_off.begin());
            }
        }
        Hcam = cam_idx; Hpt = pt_idx; Hwin = win.data();
        Hobs = obs;
        if (sorted_per_window) {
            Hcam = h_cam.data(); Hobs = h_obs.data(); Hpt = nullptr; Hwin = nullptr;
        } else if (!in_order) {
            h_cam.alloc(No); h_pt.alloc(No); h_win.alloc(No); h_obs.alloc(2 * (size_t)No);
            Hcam = h_cam.data(); Hpt = h_pt.data(); Hwin = h_win.data(); Hobs = h_obs.data();
            // scatter by (window, point): slots of a point are claimed atomically, then every point orders its
            // observations by (camera, original index) -- the result is the stable order a serial pass produces
            std::vector<int> pos(pt_off.begin(), pt_off.end() - 1);
            RawBuf<int> orig(No);
    #pragma omp parallel for schedule(static)
            for (int i = 0; i < No; i++) {
                int *pp = &pos[(size_t)win[i] * Np + pt_idx[i]];
                int d;
    #pragma omp atomic capture
                { d = *pp; (*pp)++; }
                orig[d] = i;
            }
            const long long nq = (long long)pt_off.size() - 1;
    #pragma omp parallel
            {
                std::vector<std::pair<int, int>> tmp;
    #pragma omp for schedule(dynamic, 4096)
                for (long long q = 0; q < nq; q++) {
                    const int a = pt_off[q], b = pt_off[q + 1];
                    if (b == a) continue;
                    tmp.clear();
                    for (int d = a; d < b; d++) tmp.push_back({cam_idx[orig[d]], orig[d]});
                    if (b - a > 1) std::sort(tmp.begin(), tmp.end());
                    const int wq = (int)(q / Np), pq = (int)(q - (long long)wq * Np);
                    for (int d = a; d < b; d++) {
                        const int src = tmp[d - a].second;
                        h_cam[d] = tmp[d - a].first; h_pt[d] = pq; h_win[d] = wq;
                        h_obs[2 * (size_t)d] = obs[2 * (size_t)src]; h_obs[2 * (size_t)d + 1] = obs[2 * (size_t)src + 1];
                    }
                }
            }
        }
        return PMV_OK;
    }
};

pmv_ba_problem *ba_problem_create(pmv_ctx *ctx, const double *poses, const double *points, const double *obs,
                                  const int32_t *cam_idx, const int32_t *pt_idx, const int32_t *obs_off, int W, int Nc,
                                  int Np, int No, const double K[9], double huber_delta, int sharded_rank,
                                  int sharded_nranks, bool transient)
{
    if (!ctx) return nullptr;
    if (!poses || !points || (No > 0 && (!obs || !cam_idx || !pt_idx)) || !K || W <= 0 || Nc <= 0 || Np <= 0 || No < 0 ||
        (W > 1 && !obs_off)) {
        ctx->fail(PMV_ERR_INVALID, "pmv_ba_problem_create: bad argument");
        return nullptr;
    }
    if (sharded_nranks > 1 && W != 1) {
        ctx->fail(PMV_ERR_UNSUPPORTED, "sharded bundle adjustment takes one problem (W == 1)");
        return nullptr;
    }
    cudaSetDevice(ctx->device);
    // PMV_BA_TRACE=1: wall time of the phases of this function on stderr (tools / tuning)
    const char *create_trace = getenv("PMV_BA_TRACE");
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!(create_trace && create_trace[0] == '1')) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "ba_problem_create: %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    // ---- host: order observations by (window, point), group them by (window, camera) (BAIndex above) ----
    BAIndex X;
    X.obs = obs; X.cam_idx = cam_idx; X.pt_idx = pt_idx; X.obs_off = obs_off; X.W = W; X.Nc = Nc; X.Np = Np; X.No = No;
    {
        const char *err = nullptr;
        if (X.build(&err) != PMV_OK) { ctx->fail(PMV_ERR_INVALID, err ? err : "pmv_ba_problem_create: bad observation list"); return nullptr; }
    }
    std::vector<int> &pt_off = X.pt_off, &cam_off = X.cam_off;
    const int32_t *&Hcam = X.Hcam, *&Hpt = X.Hpt, *&Hwin = X.Hwin;
    const double *&Hobs = X.Hobs;
    RawBuf<int> &cam_obs = X.cam_obs;
    auto fill_win = [&]() { X.fill_win(); };
    auto fill_pt_win_sorted = [&]() { X.fill_pt_win_sorted(); };
    auto build_cam_obs = [&]() { X.build_cam_obs(); };
    lap("index observations");
    // PMV_BA_FORCE_GENERAL=1 (tests) keeps small problems on the general path so both are exercised
    const char *force_general = getenv("PMV_BA_FORCE_GENERAL");
    // The window kernels give one CTA per window: they win once a batch fills a good part of the GPU, while
    // a lone window (the per-keyframe call of the pipeline) is twice as fast spread over the SMs by the
    // general path (tools/ba_latency.py: 0.61 vs 1.34 ms for 5 poses x 400 points x 5 iterations).
    // PMV_BA_FORCE_WINDOW=1 (tests) keeps any eligible problem on the window path.
    const char *force_window = getenv("PMV_BA_FORCE_WINDOW");
    bool window_ok = pmv_internal_ba_window_eligible(Nc, Np) && sharded_nranks <= 1 &&
                     !(force_general && force_general[0] == '1') &&
                     (W >= 4 || (force_window && force_window[0] == '1'));
    std::vector<unsigned> h_vis;
    if (window_ok) {
        h_vis.assign((size_t)W * Np, 0u);
        int twice = 0;
        const long long nq = (long long)W * Np;
#pragma omp parallel for schedule(static) reduction(| : twice)
        for (long long q = 0; q < nq; q++) {       // the observations of a point are contiguous (and ordered by camera)
            unsigned m = 0;
            for (int d = pt_off[q]; d < pt_off[q + 1]; d++) {
                if (m & (1u << Hcam[d])) twice = 1;   // the same camera sees the point twice -> general path
                m |= 1u << Hcam[d];
            }
            h_vis[q] = m;
        }
        if (twice) window_ok = false;
    }
    // structural envelope of the reduced camera system: camera c couples with cameras up to emax[c]
    std::vector<double> emax(Nc);
    for (int c = 0; c < Nc; c++) emax[c] = c;
    if (W == 1) {
        for (size_t q = 0; q + 1 < pt_off.size(); q++) {
            int cm = -1;
            for (int d = pt_off[q]; d < pt_off[q + 1]; d++) cm = std::max(cm, Hcam[d]);
            for (int d = pt_off[q]; d < pt_off[q + 1]; d++) emax[Hcam[d]] = std::max(emax[Hcam[d]], (double)cm);
        }
    } else {
        for (int c = 0; c < Nc; c++) emax[c] = Nc - 1;   // batched windows use the small-n kernels anyway
    }
    lap("visibility + envelope");
    // ---- runs of points with the same camera tuple (ba_runs.cuh): one large problem whose points see <= 8 cameras.
    // PMV_BA_NO_RUNS=1 keeps the pair-list path, PMV_BA_FORCE_RUNS=1 (tests) sends small problems through the runs.
    std::vector<int> run_off;
    std::vector<int2> run_pt;
    int run_kbegin[RUN_MAXK + 2] = {};
    {
        const char *no_runs = getenv("PMV_BA_NO_RUNS"), *force_runs = getenv("PMV_BA_FORCE_RUNS");
        const bool forced = force_runs && force_runs[0] == '1';
        bool want = W == 1 && !window_ok && !transient && !(no_runs && no_runs[0] == '1') && (No >= 100000 || forced);
        // key buffers live in the context (grow-only): fresh 16 MB blocks fault their pages in again on every call
        typedef std::pair<unsigned long long, int> RunKey;
        RunKey *keys = nullptr;
        size_t nkeys = 0;
        if (want) {
            keys = static_cast<RunKey *>(ctx->host[0].reserve(sizeof(RunKey) * (size_t)Np));
            if (!keys) { ctx->fail(PMV_ERR_NOMEM, "ba run keys"); return nullptr; }
            nkeys = (size_t)Np;
            int too_wide = 0;
#pragma omp parallel for schedule(static) reduction(| : too_wide)
            for (int q = 0; q < Np; q++) {
                const int a = pt_off[q], b = pt_off[q + 1];
                if (b - a > RUN_MAXK) too_wide = 1;
                unsigned long long h = 1469598103934665603ull;
                for (int d = a; d < b; d++) h = (h ^ (unsigned long long)(unsigned)Hcam[d]) * 1099511628211ull;
                // tuple size first, then the tuple; unobserved points sort to the end and are cut off below
                keys[q] = {b == a ? ~0ull : ((unsigned long long)(b - a) << 60) | (h >> 4), q};
            }
            if (too_wide) want = false;
            lap("runs: tuple keys");
        }
        if (want && nkeys > 0) {
            {   // parallel sort by buckets of the key's top 18 bits (tuple size + 14 hash bits): count, scan, scatter, then
                // every bucket -- a handful of tuples -- is sorted on its own; 1 M keys: 25 -> ~6 ms on 16 cores
                constexpr int KB_BITS = 18;
                const size_t nk = nkeys;
                std::vector<int> bstart(((size_t)1 << KB_BITS) + 1, 0);
                auto bucket = [](unsigned long long k) { return (size_t)(k >> (64 - KB_BITS)); };
#pragma omp parallel for schedule(static)
                for (long long q = 0; q < (long long)nk; q++) {
                    int *c = &bstart[bucket(keys[q].first) + 1];
#pragma omp atomic
                    (*c)++;
                }
                std::partial_sum(bstart.begin(), bstart.end(), bstart.begin());
                std::vector<int> cursor(bstart.begin(), bstart.end() - 1);
                RunKey *sorted = static_cast<RunKey *>(ctx->host[1].reserve(sizeof(RunKey) * nk));
                if (!sorted) { ctx->fail(PMV_ERR_NOMEM, "ba run keys"); return nullptr; }
#pragma omp parallel for schedule(static)
                for (long long q = 0; q < (long long)nk; q++) {
                    int *c = &cursor[bucket(keys[q].first)];
                    int d;
#pragma omp atomic capture
                    { d = *c; (*c)++; }
                    sorted[d] = keys[q];
                }
#pragma omp parallel for schedule(dynamic, 256)
                for (long long b = 0; b < ((long long)1 << KB_BITS); b++)
                    if (bstart[b + 1] - bstart[b] > 1) std::sort(sorted + bstart[b], sorted + bstart[b + 1]);
                keys = sorted;
            }
            while (nkeys > 0 && keys[nkeys - 1].first == ~0ull) nkeys--;
            lap("runs: sort");
        }
        if (want && nkeys > 0) {
            auto same_tuple = [&](int qa, int qb) {
                const int a = pt_off[qa], b = pt_off[qb], ka = pt_off[qa + 1] - a;
                if (ka != pt_off[qb + 1] - b) return false;
                for (int d = 0; d < ka; d++) if (Hcam[a + d] != Hcam[b + d]) return false;
                return true;
            };
            run_pt.resize(nkeys);
            std::vector<char> new_tuple(nkeys);
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < (long long)nkeys; i++) {
                run_pt[i] = make_int2(keys[i].second, pt_off[keys[i].second]);
                new_tuple[i] = i == 0 || keys[i].first != keys[i - 1].first || !same_tuple(keys[i].second, keys[i - 1].second);
            }
            for (size_t i = 0; i < nkeys; i++)     // a tuple's points in runs of at most RUN_MAXLEN
                if (new_tuple[i] || (int)i - run_off.back() >= RUN_MAXLEN) run_off.push_back((int)i);
            run_off.push_back((int)nkeys);
            {   // first run of every tuple size
                int r = 0;
                const int nr = (int)run_off.size() - 1;
                for (int k = 1; k <= RUN_MAXK + 1; k++) {
                    while (r < nr && (int)(keys[run_off[r]].first >> 60) < k) r++;
                    run_kbegin[k] = r;
                }
            }
            // short runs would put ~540 atomics per point on S again: the pair lists handle that case better
            if (!forced && nkeys < 3 * (run_off.size() - 1)) { run_off.clear(); run_pt.clear(); }
        }
    }
    const bool use_runs = !run_off.empty();
    const bool need_cam_obs = !use_runs && !window_ok;
    const bool need_pt_win = !use_runs && !window_ok;   // obs_pt / obs_win: ba_linearize_kernel and the pair lists
    if (need_cam_obs) build_cam_obs();
    lap("runs");
    pmv_ba_problem *p = new pmv_ba_problem();
    p->ctx = ctx;
    p->sharded = sharded_nranks > 1;
    p->rank = sharded_rank;
    BADev &D = p->D;
    D.W = W; D.Nc = Nc; D.Np = Np; D.n = 6 * Nc; D.No = No;
    D.fx = K[0]; D.cx = K[2]; D.fy = K[4]; D.cy = K[5]; D.delta = huber_delta;
    const size_t wc = (size_t)W * Nc, wp = (size_t)W * Np, n = D.n;
    int *d_cam, *d_pt, *d_win, *d_ptoff, *d_camoff, *d_camobs, *d_camact;
    double *d_obs;
    int rc = 0;
    const int nblk = ((int)n + PMV_CHOL_NB - 1) / PMV_CHOL_NB;
    int *d_lim = nullptr;
    double *d_sys = nullptr;
    p->use_window = window_ok ? 1 : 0;
    auto alloc_all = [&]() {
    rc |= dev_alloc(p, &d_cam, No); rc |= dev_alloc(p, &d_pt, need_pt_win ? No : 1); rc |= dev_alloc(p, &d_win, need_pt_win ? No : 1);
    rc |= dev_alloc(p, &d_obs, 2 * (size_t)No); rc |= dev_alloc(p, &d_ptoff, wp + 1); rc |= dev_alloc(p, &d_camoff, wc + 1);
    rc |= dev_alloc(p, &d_camobs, need_cam_obs ? No : 1); rc |= dev_alloc(p, &d_camact, wc);
    rc |= dev_alloc(p, &D.poses, wc * 6); rc |= dev_alloc(p, &D.points, wp * 3);
    rc |= dev_alloc(p, &D.cand_poses, wc * 6); rc |= dev_alloc(p, &D.cand_points, wp * 3);
    rc |= dev_alloc(p, &p->d_init_poses, wc * 6); rc |= dev_alloc(p, &p->d_init_points, wp * 3);
    if (use_runs) {     // residuals and Jacobians are recomputed where they are used: no linearisation buffers
        rc |= dev_alloc(p, &p->d_run_off, run_off.size()); rc |= dev_alloc(p, &p->d_run_pt, run_pt.size());
        rc |= dev_alloc(p, &D.trig, wc * 8); rc |= dev_alloc(p, &D.cand_trig, wc * 8);
    } else if (!window_ok) {   // the window path never materialises the linearisation
        rc |= dev_alloc(p, &D.Lr, 2 * (size_t)No); rc |= dev_alloc(p, &D.Ljc, 12 * (size_t)No); rc |= dev_alloc(p, &D.Ljp, 6 * (size_t)No);
    } else {
        rc |= dev_alloc(p, &p->d_vis, wp); rc |= dev_alloc(p, &p->d_camR, wc * 36); rc |= dev_alloc(p, &p->d_candR, wc * 9);
    }
    rc |= dev_alloc(p, &D.scale_c, wc * 6); rc |= dev_alloc(p, &D.scale_p, wp * 3);
    rc |= dev_alloc(p, &D.diag_c, wc * 6); rc |= dev_alloc(p, &D.diag_p, wp * 3);
    rc |= dev_alloc(p, &D.U, wc * 36); rc |= dev_alloc(p, &D.gc, wc * 6);
    // S and rhs contiguous ([S | rhs]) so a sharded solve reduces them with one collective
    rc |= dev_alloc(p, &d_sys, window_ok ? 8 : (size_t)W * (n * n + n) + (n + 1) * 8);   // the window path keeps S on chip
    rc |= dev_alloc(p, &D.yc, (size_t)W * n);
    rc |= dev_alloc(p, &D.Vinv, wp * 6); rc |= dev_alloc(p, &D.gp, wp * 3);
    rc |= dev_alloc(p, &D.st, W);
    rc |= dev_alloc(p, &p->d_Uraw, 27 * wc + W); rc |= dev_alloc(p, &p->d_Uraw_red, 27 * wc + W);
    rc |= dev_alloc(p, &p->d_scal, 4 * (size_t)W); rc |= dev_alloc(p, &p->d_scal_red, 4 * (size_t)W);
    rc |= dev_alloc(p, &d_lim, nblk);
    };
    p->transient = transient ? 1 : 0;
    if (transient && !p->sharded) {
        p->arena_mode = 1;
        alloc_all();
        const size_t total = p->arena_off + 256;
        if (total <= ((size_t)64 << 20)) {
            cudaError_t e = ctx->scratch[7].reserve(total);
            if (e != cudaSuccess) { ctx->fail(PMV_ERR_NOMEM, "ba arena", e); delete p; return nullptr; }
            p->arena_mode = 2; p->arena_base = reinterpret_cast<char *>(ctx->scratch[7].p); p->arena_off = 0;
        } else {
            p->arena_mode = 0;
        }
    }
    alloc_all();
    if (rc) { pmv_ba_problem_destroy(p); return nullptr; }
    D.S = d_sys; D.rhs = d_sys + (size_t)W * n * n;
    D.obs_cam = d_cam; D.obs_pt = d_pt; D.obs_win = d_win; D.obs_xy = d_obs;
    D.pt_off = d_ptoff; D.cam_off = d_camoff; D.cam_obs = need_cam_obs ? d_camobs : nullptr; D.cam_active = d_camact;
    lap("allocate");
    std::vector<int> h_camact(wc);
    for (size_t q = 0; q < wc; q++) h_camact[q] = cam_off[q + 1] > cam_off[q] ? 1 : 0;
    cudaStream_t s = ctx->stream;
    bool ok = true;
    // Large host arrays travel through two pinned staging buffers of the context: OpenMP threads fill one while the
    // other is on the link (a cudaMemcpyAsync from pageable memory moves ~10 GB/s through the driver's own staging).
    constexpr size_t STAGE = (size_t)32 << 20;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    bool stage_busy[2] = {false, false};
    int stage_k = 0;
    const bool stage_ok = !getenv("PMV_BA_NO_STAGING") && ctx->pin[2].reserve(STAGE) == cudaSuccess && ctx->pin[3].reserve(STAGE) == cudaSuccess &&
                          cudaEventCreateWithFlags(&stage_ev[0], cudaEventDisableTiming) == cudaSuccess &&
                          cudaEventCreateWithFlags(&stage_ev[1], cudaEventDisableTiming) == cudaSuccess;
    auto up = [&](void *dst, const void *src, size_t bytes) {
        if (!bytes) return;
        if (!stage_ok || bytes < ((size_t)4 << 20)) {
            if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) ok = false;
            return;
        }
        for (size_t off = 0; off < bytes; off += STAGE, stage_k++) {
            const int b = stage_k & 1;
            const size_t nb = std::min(STAGE, bytes - off);
            char *pin = ctx->pin[2 + b].as<char>();
            if (stage_busy[b] && cudaEventSynchronize(stage_ev[b]) != cudaSuccess) ok = false;
            const char *from = static_cast<const char *>(src) + off;
            const long long pieces = (long long)((nb + ((size_t)1 << 20) - 1) >> 20);
#pragma omp parallel for schedule(static)
            for (long long q = 0; q < pieces; q++) {
                const size_t a = (size_t)q << 20, len = std::min((size_t)1 << 20, nb - a);
                memcpy(pin + a, from + a, len);
            }
            if (cudaMemcpyAsync(static_cast<char *>(dst) + off, pin, nb, cudaMemcpyHostToDevice, s) != cudaSuccess) ok = false;
            if (cudaEventRecord(stage_ev[b], s) != cudaSuccess) ok = false;
            stage_busy[b] = true;
        }
    };
    up(d_cam, Hcam, sizeof(int) * No);
    if (need_pt_win) { fill_win(); fill_pt_win_sorted(); up(d_pt, Hpt, sizeof(int) * No); up(d_win, Hwin, sizeof(int) * No); }   // read by the general path's kernels only
    up(d_obs, Hobs, sizeof(double) * 2 * No); up(d_ptoff, pt_off.data(), sizeof(int) * (wp + 1));
    up(d_camoff, cam_off.data(), sizeof(int) * (wc + 1));
    if (need_cam_obs) up(d_camobs, cam_obs.data(), sizeof(int) * No);
    up(d_camact, h_camact.data(), sizeof(int) * wc);
    up(p->d_init_poses, poses, sizeof(double) * wc * 6); up(p->d_init_points, points, sizeof(double) * wp * 3);
    if (window_ok) up(p->d_vis, h_vis.data(), sizeof(unsigned) * wp);
    if (use_runs) {
        up(p->d_run_off, run_off.data(), sizeof(int) * run_off.size()); up(p->d_run_pt, run_pt.data(), sizeof(int2) * run_pt.size());
        p->nruns = (int)run_off.size() - 1;
        for (int k = 0; k < RUN_MAXK + 2; k++) p->run_kbegin[k] = run_kbegin[k];
        if (!getenv("PMV_RUN_NO_SIDE")) {
            cudaStream_t s2 = nullptr;
            bool okc = cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking) == cudaSuccess;
            for (auto &rs : p->run_side) {
                rs.s2 = s2;
                okc = okc && cudaEventCreateWithFlags(&rs.fork, cudaEventDisableTiming) == cudaSuccess &&
                      cudaEventCreateWithFlags(&rs.join, cudaEventDisableTiming) == cudaSuccess;
            }
            if (!okc) ok = false;
        }
    }
    if (ok && cudaStreamSynchronize(s) != cudaSuccess) ok = false;  // host vectors die at return
    for (auto &e : stage_ev) if (e) cudaEventDestroy(e);
    lap("upload");
    if (!ok) { ctx->fail(PMV_ERR_CUDA, "pmv_ba_problem_create: upload failed", cudaGetLastError()); pmv_ba_problem_destroy(p); return nullptr; }
    {
        // envelope per 64-row block (global over ranks when the points are sharded)
        if (p->sharded) {
            // one max-reduction over the ranks: the envelope (emax) and which cameras are observed at all -- a rank
            // whose point shard never sees camera c must still move it with everybody else (replicated poses)
            double *d_e = nullptr;
            if (dev_alloc(p, &d_e, 4 * (size_t)Nc) != PMV_OK) { pmv_ba_problem_destroy(p); return nullptr; }
            std::vector<double> ex(2 * (size_t)Nc);
            for (int c = 0; c < Nc; c++) { ex[c] = emax[c]; ex[Nc + c] = h_camact[c]; }
            cudaMemcpyAsync(d_e, ex.data(), sizeof(double) * 2 * Nc, cudaMemcpyHostToDevice, s);
            if (pmv_internal_ba_allreduce(ctx, d_e, d_e + 2 * Nc, 2 * (size_t)Nc, 1, s) != PMV_OK) { pmv_ba_problem_destroy(p); return nullptr; }
            cudaMemcpyAsync(ex.data(), d_e + 2 * Nc, sizeof(double) * 2 * Nc, cudaMemcpyDeviceToHost, s);
            if (cudaStreamSynchronize(s) != cudaSuccess) { ctx->fail(PMV_ERR_CUDA, "envelope exchange failed"); pmv_ba_problem_destroy(p); return nullptr; }
            for (int c = 0; c < Nc; c++) { emax[c] = ex[c]; h_camact[c] = ex[Nc + c] > 0.5 ? 1 : 0; }
            cudaMemcpyAsync(d_camact, h_camact.data(), sizeof(int) * Nc, cudaMemcpyHostToDevice, s);
            if (cudaStreamSynchronize(s) != cudaSuccess) { ctx->fail(PMV_ERR_CUDA, "camera activity upload failed"); pmv_ba_problem_destroy(p); return nullptr; }
        }
        p->chol_lim.assign(nblk, (int)n);
        int run = 0;
        for (int kb = 0; kb < nblk; kb++) {
            const int c0 = (kb * PMV_CHOL_NB) / 6, c1 = std::min(Nc - 1, (kb * PMV_CHOL_NB + PMV_CHOL_NB - 1) / 6);
            for (int c = c0; c <= c1; c++) run = std::max(run, (int)emax[c]);
            p->chol_lim[kb] = std::min((int)n, 6 * (run + 1));
        }
        cudaMemcpyAsync(d_lim, p->chol_lim.data(), sizeof(int) * nblk, cudaMemcpyHostToDevice, s);
        D.chol_lim = d_lim;
        // ba_clear_system_kernel clears only the envelope of one large system per iteration: everything else is zeroed here, once
        if (W == 1 && n > 160 && !window_ok) cudaMemsetAsync(D.S, 0, sizeof(double) * n * n, s);
        // ---- two-sided solve: separator of w columns in the middle of a long banded system
        {
            const char *no_split = getenv("PMV_CHOL_NO_SPLIT");
            const int NBk = PMV_CHOL_NB, nn = (int)n;
            int maxw = 0;
            for (int kb = 0; kb < nblk; kb++) maxw = std::max(maxw, p->chol_lim[kb] - kb * NBk);
            const int w = (maxw + NBk - 1) / NBk * NBk;
            const int a = ((nn - w) / 2) / NBk * NBk;
            const bool want = W == 1 && !window_ok && p->arena_mode == 0 && !(no_split && no_split[0] == '1') && nn >= 1024 && w > 0 && 6 * w <= nn && a >= 2 * w &&
                              pmv_internal_ba_cholesky_band_T(nn, p->chol_lim.data()) > 0;
            if (want) {
                BASplit &P = p->split;
                P.a = a; P.w = w; P.h1 = a + w; P.h2 = nn - a;
                std::vector<int> &l1 = p->split_lim[0], &l2 = p->split_lim[1], &l3 = p->split_lim[2];
                const int nb1 = (P.h1 + NBk - 1) / NBk, nb2 = (P.h2 + NBk - 1) / NBk, nb3 = (w + NBk - 1) / NBk;
                l1.resize(nb1); l2.assign(nb2, 0); l3.assign(nb3, w);
                for (int kb = 0; kb < nb1; kb++) l1[kb] = std::min(p->chol_lim[kb], P.h1);
                // reversed system: row i' <-> original index r = n-1-i' ; it reaches the reversed image of the first
                // original row whose envelope covers r
                {
                    std::vector<int> first(nn);
                    int kb = 0;
                    for (int r = 0; r < nn; r++) {
                        while (kb < nblk && p->chol_lim[kb] <= r) kb++;
                        first[r] = std::min(r, kb * NBk);
                    }
                    int run2 = 0;
                    for (int ip = 0; ip < P.h2; ip++) {
                        const int r = nn - 1 - ip;
                        const int reach = nn - first[r];               // one past the last reversed column
                        run2 = std::max(run2, std::min(reach, P.h2));
                        l2[ip / NBk] = std::max(l2[ip / NBk], run2);
                    }
                    for (int q = 1; q < nb2; q++) l2[q] = std::max(l2[q], l2[q - 1]);
                }
                const bool elig = pmv_internal_ba_cholesky_band_T(P.h1, l1.data()) > 0 && pmv_internal_ba_cholesky_band_T(P.h2, l2.data()) > 0 &&
                                  pmv_internal_ba_cholesky_band_T(w, l3.data()) > 0;
                if (elig) {
                    int rc2 = 0;
                    rc2 |= dev_alloc(p, &P.S2, (size_t)P.h2 * P.h2); rc2 |= dev_alloc(p, &P.b2, (size_t)P.h2); rc2 |= dev_alloc(p, &P.y2, (size_t)P.h2);
                    rc2 |= dev_alloc(p, &P.AM, (size_t)w * w); rc2 |= dev_alloc(p, &P.bM, (size_t)w);
                    rc2 |= dev_alloc(p, &P.S3, (size_t)w * w); rc2 |= dev_alloc(p, &P.b3, (size_t)w); rc2 |= dev_alloc(p, &P.y3, (size_t)w);
                    rc2 |= dev_alloc(p, &P.lim1, nb1); rc2 |= dev_alloc(p, &P.lim2, nb2); rc2 |= dev_alloc(p, &P.lim3, nb3);
                    rc2 |= dev_alloc(p, &P.st3, 3);
                    if (rc2) { pmv_ba_problem_destroy(p); return nullptr; }
                    cudaMemsetAsync(P.S2, 0, sizeof(double) * (size_t)P.h2 * P.h2, s);   // entries outside the envelope are never written
                    cudaMemcpyAsync(P.lim1, l1.data(), sizeof(int) * nb1, cudaMemcpyHostToDevice, s);
                    cudaMemcpyAsync(P.lim2, l2.data(), sizeof(int) * nb2, cudaMemcpyHostToDevice, s);
                    cudaMemcpyAsync(P.lim3, l3.data(), sizeof(int) * nb3, cudaMemcpyHostToDevice, s);
                    P.lim1_h = l1.data(); P.lim2_h = l2.data(); P.lim3_h = l3.data();
                    P.lim_orig = d_lim;
                    bool okc = cudaStreamCreateWithFlags(&P.s2, cudaStreamNonBlocking) == cudaSuccess;
                    for (auto &e : P.ev) okc = okc && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
                    if (!okc) { ctx->fail(PMV_ERR_CUDA, "split solve: stream / events"); pmv_ba_problem_destroy(p); return nullptr; }
                    P.enabled = 1;
                }
            }
        }
        // ---- partitioned solve: P segments separated by P - 1 separators of w columns (BAPart).  OPT-IN (PMV_CHOL_PARTS >= 3):
        //      measured on config 5 it does not beat the two-sided solve (2.39 vs 2.31 ms per iteration at P = 4; the spikes and the
        //      sequential separator chain cost what the shorter segment chains save), so the two-sided solve stays the default
        {
            const char *parts_env = getenv("PMV_CHOL_PARTS");
            const int NBk = PMV_CHOL_NB, nn = (int)n;
            int maxw = 0;
            for (int kb = 0; kb < nblk; kb++) maxw = std::max(maxw, p->chol_lim[kb] - kb * NBk);
            const int w = (maxw + NBk - 1) / NBk * NBk;
            int P = parts_env ? atoi(parts_env) : 0;
            while (P >= 3 && ((nn - (P - 1) * w) / P) / NBk * NBk < 2 * w) P--;
            const bool want = p->split.enabled && P >= 3 && P <= PMV_PART_MAX && w > 0;
            if (want) {
                BAPart &Q = p->part;
                Q.P = P; Q.w = w; Q.nR = (P - 1) * w;
                const int ns = ((nn - (P - 1) * w) / P) / NBk * NBk;
                bool elig = true;
                for (int i = 0; i < P; i++) {
                    Q.a[i] = i * (ns + w);
                    Q.ns[i] = i < P - 1 ? ns : nn - Q.a[i];
                    Q.nx[i] = i < P - 1 ? ns + w : Q.ns[i];
                    std::vector<int> &l = p->part_lim[i];
                    const int nbx = (Q.nx[i] + NBk - 1) / NBk, kb0 = Q.a[i] / NBk;
                    l.resize(nbx);
                    for (int kb = 0; kb < nbx; kb++) {
                        l[kb] = std::min(p->chol_lim[kb0 + kb], Q.a[i] + Q.nx[i]) - Q.a[i];
                        elig = elig && (l[kb] - kb * NBk + NBk - 1) / NBk <= 11;           // part_spike_kernel: diagonal + 10 tiles per block row
                    }
                    elig = elig && pmv_internal_ba_cholesky_band_T(Q.nx[i], l.data()) > 0;
                    // the separator right of segment i must not reach beyond segment i + 1
                    if (i < P - 1) elig = elig && p->chol_lim[(Q.a[i] + Q.nx[i] - 1) / NBk] <= Q.a[i] + Q.nx[i] + (i + 1 < P - 1 ? ns : nn);
                }
                std::vector<int> &lr = p->part_lim[PMV_PART_MAX];            // dense envelope of one separator block
                const int nbr = w / NBk;
                lr.assign(nbr, w);
                elig = elig && pmv_internal_ba_cholesky_band_T(w, lr.data()) > 0 && w <= 320;   // 320: staged back-substitution
                if (getenv("PMV_BA_TRACE")) fprintf(stderr, "[pmv] part solve: n %d w %d P %d ns %d nR %d eligible %d\n", nn, w, P, ns, Q.nR, (int)elig);
                if (elig) {
                    int rc2 = 0;
                    for (int i = 0; i < P; i++) {
                        rc2 |= dev_alloc(p, &Q.limX[i], p->part_lim[i].size());
                        if (i > 0) rc2 |= dev_alloc(p, &Q.G[i], (size_t)Q.nx[i] * w);
                    }
                    const size_t ww = (size_t)w * w;
                    rc2 |= dev_alloc(p, &Q.Dsep, (P - 1) * ww); rc2 |= dev_alloc(p, &Q.E, (P - 1) * ww); rc2 |= dev_alloc(p, &Q.F, (P - 1) * ww);
                    rc2 |= dev_alloc(p, &Q.bR, (size_t)Q.nR); rc2 |= dev_alloc(p, &Q.yR, (size_t)Q.nR);
                    rc2 |= dev_alloc(p, &Q.limD, nbr); rc2 |= dev_alloc(p, &Q.stp, 2 * P - 1);
                    if (rc2) { pmv_ba_problem_destroy(p); return nullptr; }
                    for (int i = 0; i < P; i++) {
                        cudaMemcpyAsync(Q.limX[i], p->part_lim[i].data(), sizeof(int) * p->part_lim[i].size(), cudaMemcpyHostToDevice, s);
                        Q.limX_h[i] = p->part_lim[i].data();
                    }
                    cudaMemcpyAsync(Q.limD, lr.data(), sizeof(int) * nbr, cudaMemcpyHostToDevice, s);
                    cudaMemsetAsync(Q.Dsep, 0, sizeof(double) * (P - 1) * ww, s);   // tiles below the diagonal are never written
                    Q.limD_h = lr.data();
                    bool okc = true;
                    for (int i = 1; i < P; i++) okc = okc && cudaStreamCreateWithFlags(&Q.str[i], cudaStreamNonBlocking) == cudaSuccess;
                    for (auto &e : Q.ev_fork) okc = okc && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
                    for (int k = 0; k < 2; k++)
                        for (int i = 1; i < P; i++) okc = okc && cudaEventCreateWithFlags(&Q.ev_join[k][i], cudaEventDisableTiming) == cudaSuccess;
                    if (!okc) { ctx->fail(PMV_ERR_CUDA, "part solve: streams / events"); pmv_ba_problem_destroy(p); return nullptr; }
                    Q.enabled = 1;
                }
            }
        }
        if (p->sharded) {
            std::vector<long long> off(n + 1);
            long long acc = 0;
            for (size_t r = 0; r < n; r++) { off[r] = acc; acc += p->chol_lim[r / PMV_CHOL_NB] - (int)r; }
            off[n] = acc;
            p->band_count = acc + (long long)n;
            if (dev_alloc(p, &p->d_band_off, n + 1) != PMV_OK || dev_alloc(p, &p->d_band, (size_t)p->band_count) != PMV_OK) {
                pmv_ba_problem_destroy(p);
                return nullptr;
            }
            cudaMemcpyAsync(p->d_band_off, off.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, s);
        }
        cudaStreamSynchronize(s);
    }
    lap("envelope exchange + split setup");
    // ---- pair list: for every co-observed camera pair (ci <= ck) the observation pairs (i, k) of the points
    // that see both, cut into segments of <= 512 entries (one warp each).  Large single problems only: below
    // ~100 k observations the per-point kernel has more parallelism than there are segments.
    {
        const char *no_pairs = getenv("PMV_BA_NO_PAIRS");
        const char *force_pairs = getenv("PMV_BA_FORCE_PAIRS");   // tests: small problems through the pair path
        const bool want = !use_runs && W == 1 && !window_ok && p->arena_mode == 0 && (size_t)Nc * Nc <= ((size_t)16 << 20) &&
                          (No >= 100000 || (force_pairs && force_pairs[0] == '1')) && !(no_pairs && no_pairs[0] == '1');
        if (want) {
            std::vector<unsigned> pcnt((size_t)Nc * Nc + 1, 0u);
            for (size_t q = 0; q + 1 < pt_off.size(); q++)
                for (int i = pt_off[q]; i < pt_off[q + 1]; i++)
                    for (int k = pt_off[q]; k < pt_off[q + 1]; k++)
                        if (Hcam[i] <= Hcam[k]) pcnt[(size_t)Hcam[i] * Nc + Hcam[k] + 1]++;
            unsigned long long total = 0;
            for (size_t a = 1; a < pcnt.size(); a++) total += pcnt[a];
            if (total > 0 && total < (1ull << 31)) {
                std::vector<BAPairSeg> segs;
                std::vector<unsigned> cur((size_t)Nc * Nc);
                unsigned run = 0;
                for (size_t a = 0; a < (size_t)Nc * Nc; a++) {
                    const unsigned c = pcnt[a + 1];
                    cur[a] = run;
                    for (unsigned o = 0; o < c; o += 512)
                        segs.push_back({(int)(a / Nc), (int)(a % Nc), (int)(run + o), (int)(run + std::min(c, o + 512))});
                    run += c;
                }
                std::vector<int2> ent((size_t)total);
                for (size_t q = 0; q + 1 < pt_off.size(); q++)
                    for (int i = pt_off[q]; i < pt_off[q + 1]; i++)
                        for (int k = pt_off[q]; k < pt_off[q + 1]; k++)
                            if (Hcam[i] <= Hcam[k]) ent[cur[(size_t)Hcam[i] * Nc + Hcam[k]]++] = make_int2(i, k);
                if (dev_alloc(p, &p->d_segs, segs.size()) != PMV_OK || dev_alloc(p, &p->d_entries, ent.size()) != PMV_OK) {
                    pmv_ba_problem_destroy(p);
                    return nullptr;
                }
                cudaMemcpyAsync(p->d_segs, segs.data(), sizeof(BAPairSeg) * segs.size(), cudaMemcpyHostToDevice, s);
                cudaMemcpyAsync(p->d_entries, ent.data(), sizeof(int2) * ent.size(), cudaMemcpyHostToDevice, s);
                if (cudaStreamSynchronize(s) != cudaSuccess) {
                    ctx->fail(PMV_ERR_CUDA, "pmv_ba_problem_create: pair list upload failed", cudaGetLastError());
                    pmv_ba_problem_destroy(p);
                    return nullptr;
                }
                p->nsegs = (int)segs.size();
            }
        }
    }
    lap("pair lists");
    if (pmv_ba_problem_reset(p, nullptr, nullptr) != PMV_OK) { pmv_ba_problem_destroy(p); return nullptr; }
    lap("reset");
    return p;
}

}  // namespace

extern "C" {

PMV_API pmv_ba_problem *pmv_ba_problem_create(pmv_ctx *ctx, const double *poses, const double *points,
                                              const double *obs, const int32_t *cam_idx, const int32_t *pt_idx,
                                              const int32_t *obs_off, int W, int Nc, int Np, int No,
                                              const double K[9], double huber_delta, int sharded_rank,
                                              int sharded_nranks)
{
    return ba_problem_create(ctx, poses, points, obs, cam_idx, pt_idx, obs_off, W, Nc, Np, No, K, huber_delta,
                             sharded_rank, sharded_nranks, false);
}

PMV_API void pmv_ba_problem_destroy(pmv_ba_problem *p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    if (p->split.s2) { cudaStreamSynchronize(p->split.s2); cudaStreamDestroy(p->split.s2); }
    for (auto &e : p->split.ev) if (e) cudaEventDestroy(e);
    if (p->run_side[0].s2) { cudaStreamSynchronize(p->run_side[0].s2); cudaStreamDestroy(p->run_side[0].s2); }
    for (auto &rs : p->run_side) { if (rs.fork) cudaEventDestroy(rs.fork); if (rs.join) cudaEventDestroy(rs.join); }
    for (auto &st : p->part.str) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (auto &e : p->part.ev_fork) if (e) cudaEventDestroy(e);
    for (auto &row : p->part.ev_join) for (auto &e : row) if (e) cudaEventDestroy(e);
    for (void *q : p->allocs) cudaFreeAsync(q, p->ctx->stream);
    delete p;
}

PMV_API int pmv_ba_problem_reset(pmv_ba_problem *p, const double *poses, const double *points)
{
    if (!p) return PMV_ERR_INVALID;
    pmv_ctx *ctx = p->ctx;
    cudaSetDevice(ctx->device);
    const BADev &D = p->D;
    cudaStream_t s = ctx->stream;
    const size_t pc = (size_t)D.W * D.Nc * 6 * sizeof(double), pp = (size_t)D.W * D.Np * 3 * sizeof(double);
    if (poses) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(p->d_init_poses, poses, pc, cudaMemcpyHostToDevice, s));
    if (points) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(p->d_init_points, points, pp, cudaMemcpyHostToDevice, s));
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(D.poses, p->d_init_poses, pc, cudaMemcpyDeviceToDevice, s));
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(D.points, p->d_init_points, pp, cudaMemcpyDeviceToDevice, s));
    // the run-organised back-substitution writes the candidates of observed points only: the others keep x
    if (p->nruns > 0) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(D.cand_points, p->d_init_points, pp, cudaMemcpyDeviceToDevice, s));
    ba_state_init_kernel<<<(D.W + 127) / 128, 128, 0, s>>>(D.st, D.W);
    PMV_LAUNCH_CHECK(ctx, "ba_state_init_kernel");
    if (poses || points) PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return PMV_OK;
}

PMV_API int pmv_ba_problem_solve(pmv_ba_problem *p, int max_iters)
{
    if (!p) return PMV_ERR_INVALID;
    pmv_ctx *ctx = p->ctx;
    if (max_iters < 0) return ctx->fail(PMV_ERR_INVALID, "pmv_ba_problem_solve: max_iters < 0");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    p->D.max_iters = max_iters;
    if (p->D.n <= 160) {
        if (ctx->attr_first(PMV_ATTR_CHOL_SMALL))
            cudaFuncSetAttribute(ba_cholesky_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    }
    ProfScope ps(ctx, PMV_PHASE_BA, s);
    // iteration 0 evaluates the cost even when max_iters == 0 (Ceres: IterationZero)
    auto one_iteration = [&]() {
        return p->use_window ? pmv_internal_ba_window_iteration(ctx, p->D, p->d_vis, p->d_camR, p->d_candR, s)
                             : ba_iteration(p, s);
    };
    const char *no_graph = getenv("PMV_BA_NO_GRAPH");
    // a one-shot solve of a small problem issues a handful of launches per iteration: capturing and
    // instantiating a graph would cost more than it saves
    const bool want_graph = !(no_graph && no_graph[0] == '1') && !(p->transient && p->D.n <= 160);
    if (want_graph && (!p->graph_exec || p->graph_max_iters != max_iters)) {
        if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
        const char *trace = getenv("PMV_BA_TRACE");
        auto t_last = std::chrono::steady_clock::now();
        auto lap = [&](const char *what) {
            if (!(trace && trace[0] == '1')) return;
            const auto now = std::chrono::steady_clock::now();
            fprintf(stderr, "ba_problem_solve: %-31s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
            t_last = now;
        };
        // first iteration runs eagerly (sets function attributes, validates the launches) ...
        int rc = one_iteration();
        if (rc) return rc;
        lap("first iteration enqueued");
        // ... then the same sequence is captured for replay
        cudaGraph_t graph = nullptr;
        const uint64_t l0 = ctx->launches;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
            rc = one_iteration();
            cudaError_t e = cudaStreamEndCapture(s, &graph);
            lap("iteration captured");
            if (rc == PMV_OK && e == cudaSuccess && graph &&
                cudaGraphInstantiate(&p->graph_exec, graph, 0) == cudaSuccess) {
                lap("graph instantiated");
                p->graph_max_iters = max_iters;
                p->graph_launches = ctx->launches - l0;
            } else {
                p->graph_exec = nullptr;
            }
            if (graph) cudaGraphDestroy(graph);
            ctx->launches = l0;   // capture enqueued nothing
            cudaGetLastError();
        }
        for (int it = 1; it < std::max(max_iters, 1); it++) {
            if (p->graph_exec) {
                PMV_CUDA_TRY(ctx, cudaGraphLaunch(p->graph_exec, s));
                ctx->launches += p->graph_launches;
            } else {
                rc = one_iteration();
                if (rc) return rc;
            }
        }
        return PMV_OK;
    }
    for (int it = 0; it < std::max(max_iters, 1); it++) {
        if (want_graph && p->graph_exec) {
            PMV_CUDA_TRY(ctx, cudaGraphLaunch(p->graph_exec, s));
            ctx->launches += p->graph_launches;
        } else {
            int rc = one_iteration();
            if (rc) return rc;
        }
    }
    return PMV_OK;
}

// Device -> pageable host copy through the context's two pinned staging buffers: the next chunk travels while OpenMP
// threads copy the previous one out (a cudaMemcpyAsync into pageable memory moves ~10 GB/s).  Synchronous on return.
static int staged_download(pmv_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t s)
{
    constexpr size_t STAGE = (size_t)32 << 20;
    if (bytes < ((size_t)4 << 20) || getenv("PMV_BA_NO_STAGING") || ctx->pin[2].reserve(STAGE) != cudaSuccess ||
        ctx->pin[3].reserve(STAGE) != cudaSuccess) {
        PMV_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
        PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
        return PMV_OK;
    }
    cudaEvent_t ev[2] = {nullptr, nullptr};
    if (cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) != cudaSuccess) {
        if (ev[0]) cudaEventDestroy(ev[0]);
        return ctx->fail(PMV_ERR_CUDA, "staged download: events");
    }
    const size_t nchunks = (bytes + STAGE - 1) / STAGE;
    auto issue = [&](size_t k) {
        const size_t off = k * STAGE, nb = std::min(STAGE, bytes - off);
        cudaMemcpyAsync(ctx->pin[2 + (k & 1)].p, static_cast<const char *>(src) + off, nb, cudaMemcpyDeviceToHost, s);
        cudaEventRecord(ev[k & 1], s);
    };
    issue(0);
    bool ok = true;
    for (size_t k = 0; k < nchunks; k++) {
        if (k + 1 < nchunks) issue(k + 1);
        if (cudaEventSynchronize(ev[k & 1]) != cudaSuccess) ok = false;
        const size_t off = k * STAGE, nb = std::min(STAGE, bytes - off);
        const char *pin = ctx->pin[2 + (k & 1)].as<char>();
        char *to = static_cast<char *>(dst) + off;
        const long long pieces = (long long)((nb + ((size_t)1 << 20) - 1) >> 20);
#pragma omp parallel for schedule(static)
        for (long long q = 0; q < pieces; q++) {
            const size_t a = (size_t)q << 20, len = std::min((size_t)1 << 20, nb - a);
            memcpy(to + a, pin + a, len);
        }
    }
    cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    if (!ok || cudaStreamSynchronize(s) != cudaSuccess) return ctx->fail(PMV_ERR_CUDA, "staged download", cudaGetLastError());
    return PMV_OK;
}

PMV_API int pmv_ba_problem_download(pmv_ba_problem *p, double *poses, double *points, pmv_ba_summary *sums)
{
    if (!p) return PMV_ERR_INVALID;
    pmv_ctx *ctx = p->ctx;
    cudaSetDevice(ctx->device);
    const BADev &D = p->D;
    cudaStream_t s = ctx->stream;
    std::vector<BAState> st(D.W);
    if (poses) { int rc = staged_download(ctx, poses, D.poses, (size_t)D.W * D.Nc * 6 * sizeof(double), s); if (rc) return rc; }
    if (points) { int rc = staged_download(ctx, points, D.points, (size_t)D.W * D.Np * 3 * sizeof(double), s); if (rc) return rc; }
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(st.data(), D.st, sizeof(BAState) * D.W, cudaMemcpyDeviceToHost, s));
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    if (sums)
        for (int w = 0; w < D.W; w++) {
            sums[w].initial_cost = st[w].initial_cost; sums[w].final_cost = st[w].cost;
            sums[w].iterations = st[w].iter; sums[w].successful_steps = st[w].successful_steps;
            sums[w].termination = st[w].termination; sums[w].final_radius = st[w].radius;
        }
    return PMV_OK;
}

PMV_API size_t pmv_ba_problem_device_bytes(pmv_ba_problem *p) { return p ? p->bytes : 0; }

PMV_API int pmv_ba_index_observations(const double *obs, const int32_t *cam_idx, const int32_t *pt_idx, const int32_t *obs_off, int W,
                                      int Nc, int Np, int No, int32_t *pt_off, int32_t *cam_off, int32_t *cam, int32_t *pt, int32_t *win,
                                      double *obs_sorted, int32_t *cam_obs, int *route)
{
    if (!obs || !cam_idx || !pt_idx || W <= 0 || Nc <= 0 || Np <= 0 || No < 0 || (W > 1 && !obs_off)) return PMV_ERR_INVALID;
    BAIndex X;
    X.obs = obs; X.cam_idx = cam_idx; X.pt_idx = pt_idx; X.obs_off = obs_off; X.W = W; X.Nc = Nc; X.Np = Np; X.No = No;
    const char *err = nullptr;
    const int rc = X.build(&err);
    if (rc != PMV_OK) return rc;
    X.fill_win(); X.fill_pt_win_sorted();
    if (route) *route = X.in_order ? 0 : X.sorted_per_window ? 1 : 2;
    if (pt_off) std::copy(X.pt_off.begin(), X.pt_off.end(), pt_off);
    if (cam_off) std::copy(X.cam_off.begin(), X.cam_off.end(), cam_off);
    for (int d = 0; d < No; d++) {
        if (cam) cam[d] = X.Hcam[d];
        if (pt) pt[d] = X.Hpt[d];
        if (win) win[d] = X.Hwin[d];
        if (obs_sorted) { obs_sorted[2 * (size_t)d] = X.Hobs[2 * (size_t)d]; obs_sorted[2 * (size_t)d + 1] = X.Hobs[2 * (size_t)d + 1]; }
    }
    if (cam_obs) {
        X.build_cam_obs();
        std::copy(X.cam_obs.data(), X.cam_obs.data() + No, cam_obs);
    }
    return PMV_OK;
}

PMV_API int pmv_ba_solve_batched(pmv_ctx *ctx, double *poses, double *points, const double *obs,
                                 const int32_t *cam_idx, const int32_t *pt_idx, const int32_t *obs_off, int W,
                                 int Nc, int Np, int No, const double K[9], double huber_delta, int max_iters,
                                 pmv_ba_summary *sums)
{
    if (!ctx) return PMV_ERR_INVALID;
    pmv_ba_problem *p = ba_problem_create(ctx, poses, points, obs, cam_idx, pt_idx, obs_off, W, Nc, Np, No, K,
                                          huber_delta, 0, 1, true);
    if (!p) return ctx->last_code ? ctx->last_code : PMV_ERR_INVALID;   // the status ba_problem_create failed with
    int rc = pmv_ba_problem_solve(p, max_iters);
    if (rc == PMV_OK) rc = pmv_ba_problem_download(p, poses, points, sums);
    pmv_ba_problem_destroy(p);
    return rc;
}

PMV_API int pmv_ba_solve(pmv_ctx *ctx, double *poses, double *points, const double *obs, const int32_t *cam_idx,
                         const int32_t *pt_idx, int Nc, int Np, int No, const double K[9], double huber_delta,
                         int max_iters, pmv_ba_summary *summary)
{
    return pmv_ba_solve_batched(ctx, poses, points, obs, cam_idx, pt_idx, nullptr, 1, Nc, Np, No, K, huber_delta,
                                max_iters, summary);
}

PMV_API int pmv_ba_eval(pmv_ctx *ctx, const double *poses, const double *points, const double *obs,
                        const int32_t *cam_idx, const int32_t *pt_idx, int Nc, int Np, int No, const double K[9],
                        double huber_delta, double *r, double *J_pose, double *J_pt, double *cost)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!poses || !points || !obs || !cam_idx || !pt_idx || !K || Nc <= 0 || Np <= 0 || No <= 0)
        return ctx->fail(PMV_ERR_INVALID, "pmv_ba_eval: bad argument");
    for (int i = 0; i < No; i++)
        if (cam_idx[i] < 0 || cam_idx[i] >= Nc || pt_idx[i] < 0 || pt_idx[i] >= Np)
            return ctx->fail(PMV_ERR_INVALID, "pmv_ba_eval: observation index out of range");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    size_t need = sizeof(double) * ((size_t)Nc * 6 + (size_t)Np * 3 + (size_t)No * 22 + 8) + sizeof(int) * 2 * (size_t)No;
    cudaError_t e = ctx->scratch[6].reserve(need);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "ba_eval workspace", e);
    double *d = ctx->scratch[6].as<double>();
    double *d_poses = d; d += (size_t)Nc * 6;
    double *d_pts = d; d += (size_t)Np * 3;
    double *d_obs = d; d += (size_t)No * 2;
    double *d_r = d; d += (size_t)No * 2;
    double *d_jc = d; d += (size_t)No * 12;
    double *d_jp = d; d += (size_t)No * 6;
    double *d_cost = d; d += 8;
    int *d_cam = reinterpret_cast<int *>(d), *d_pt = d_cam + No;
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(d_poses, poses, sizeof(double) * Nc * 6, cudaMemcpyHostToDevice, s));
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(d_pts, points, sizeof(double) * Np * 3, cudaMemcpyHostToDevice, s));
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(d_obs, obs, sizeof(double) * No * 2, cudaMemcpyHostToDevice, s));
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(d_cam, cam_idx, sizeof(int) * No, cudaMemcpyHostToDevice, s));
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(d_pt, pt_idx, sizeof(int) * No, cudaMemcpyHostToDevice, s));
    PMV_CUDA_TRY(ctx, cudaMemsetAsync(d_cost, 0, 8, s));
    {
        ProfScope ps(ctx, PMV_PHASE_BA, s);
        ba_eval_kernel<<<(No + 127) / 128, 128, 0, s>>>(d_poses, d_pts, d_obs, d_cam, d_pt, No, K[0], K[2], K[4], K[5],
                                                       huber_delta, d_r, d_jc, d_jp, d_cost);
        PMV_LAUNCH_CHECK(ctx, "ba_eval_kernel");
    }
    if (r) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(r, d_r, sizeof(double) * No * 2, cudaMemcpyDeviceToHost, s));
    if (J_pose) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(J_pose, d_jc, sizeof(double) * No * 12, cudaMemcpyDeviceToHost, s));
    if (J_pt) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(J_pt, d_jp, sizeof(double) * No * 6, cudaMemcpyDeviceToHost, s));
    if (cost) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(cost, d_cost, 8, cudaMemcpyDeviceToHost, s));
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return PMV_OK;
}

}  // extern "C"

// sort.cuh -- hand-written bitonic sort of 128-bit records (descending, lexicographic on (hi, lo)).
// Used by the corner selectors (K6/K7): GFTT sorts (response bits, pixel index) so that equal
// responses come out "higher address first" exactly like OpenCV's greaterThanPtr; the reference
// ShiTomasi extractor sorts fp64 scores; FAST sorts by ~index to restore raster order after an
// unordered atomic compaction.  Candidate counts are 1e3..3e5, so an O(n log^2 n) network with the
// small strides done in shared memory costs a handful of launches.
#pragma once
#include "common.cuh"

namespace {  // internal linkage: the header is included by several translation units

struct Rec128 {
    unsigned long long hi, lo;
};

__device__ __forceinline__ bool rec_greater(const Rec128 &a, const Rec128 &b)
{
    return a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo);
}

constexpr int SORT_CHUNK = 2048;  // records per CTA in the shared-memory phases (32 KB)

// Full bitonic sort of each SORT_CHUNK block in shared memory; block b sorts descending if
// (b & 1) == 0 else ascending so that the next merge stage sees bitonic sequences.
__global__ void __launch_bounds__(1024) bitonic_local_sort(Rec128 *d, int n_pow2)
{
    __shared__ Rec128 s[SORT_CHUNK];
    const int base = blockIdx.x * SORT_CHUNK;
    for (int i = threadIdx.x; i < SORT_CHUNK; i += 1024) s[i] = d[base + i];
    __syncthreads();
    for (int k = 2; k <= SORT_CHUNK; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < SORT_CHUNK / 2; t += 1024) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int p = i | j;
                bool desc = (((base + i) & k) == 0);
                Rec128 a = s[i], b = s[p];
                if (rec_greater(b, a) == desc) { s[i] = b; s[p] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < SORT_CHUNK; i += 1024) d[base + i] = s[i];
}

// One global compare-exchange step (stride j >= SORT_CHUNK) of merge stage k.
__global__ void __launch_bounds__(256) bitonic_global_step(Rec128 *d, int n_pow2, int k, int j)
{
    int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= n_pow2 / 2) return;
    int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    int p = i | j;
    bool desc = ((i & k) == 0);
    Rec128 a = d[i], b = d[p];
    if (rec_greater(b, a) == desc) { d[i] = b; d[p] = a; }
}

// Strides j < SORT_CHUNK of merge stage k, in shared memory.
__global__ void __launch_bounds__(1024) bitonic_local_merge(Rec128 *d, int n_pow2, int k)
{
    __shared__ Rec128 s[SORT_CHUNK];
    const int base = blockIdx.x * SORT_CHUNK;
    for (int i = threadIdx.x; i < SORT_CHUNK; i += 1024) s[i] = d[base + i];
    __syncthreads();
    for (int j = SORT_CHUNK >> 1; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < SORT_CHUNK / 2; t += 1024) {
            int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            int p = i | j;
            bool desc = (((base + i) & k) == 0);
            Rec128 a = s[i], b = s[p];
            if (rec_greater(b, a) == desc) { s[i] = b; s[p] = a; }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < SORT_CHUNK; i += 1024) d[base + i] = s[i];
}

__global__ void __launch_bounds__(256) sort_pad_kernel(Rec128 *d, int n, int n_pow2)
{
    int i = n + blockIdx.x * 256 + threadIdx.x;
    if (i < n_pow2) d[i] = Rec128{0ull, 0ull};  // smallest record: ends up behind every real one
}

// Sort d[0..n) descending; d must have capacity for the next power of two >= max(n, SORT_CHUNK).
static inline int sort_capacity(int n)
{
    int p = SORT_CHUNK;
    while (p < n) p <<= 1;
    return p;
}

static inline int sort_desc_128(pmv_ctx *ctx, Rec128 *d, int n, cudaStream_t s)
{
    if (n <= 1) return PMV_OK;
    const int np2 = sort_capacity(n);
    if (np2 > n) {
        sort_pad_kernel<<<(np2 - n + 255) / 256, 256, 0, s>>>(d, n, np2);
        PMV_LAUNCH_CHECK(ctx, "sort_pad_kernel");
    }
    bitonic_local_sort<<<np2 / SORT_CHUNK, 1024, 0, s>>>(d, np2);
    PMV_LAUNCH_CHECK(ctx, "bitonic_local_sort");
    for (int k = SORT_CHUNK * 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j >= SORT_CHUNK; j >>= 1) {
            bitonic_global_step<<<(np2 / 2 + 255) / 256, 256, 0, s>>>(d, np2, k, j);
            PMV_LAUNCH_CHECK(ctx, "bitonic_global_step");
        }
        bitonic_local_merge<<<np2 / SORT_CHUNK, 1024, 0, s>>>(d, np2, k);
        PMV_LAUNCH_CHECK(ctx, "bitonic_local_merge");
    }
    return PMV_OK;
}

// ---- bucket sort for large candidate lists whose primary key is a non-negative float (GFTT at 4K: ~220 k records) -----
// The bitonic network above costs 36 launches (190 us) at 2^18 records, most of them latency.  When `hi` holds the bit
// pattern of a float in [q * max, max], its top bits split the list into a few thousand buckets that are already in
// order among themselves: histogram (by the producer kernel), exclusive scan, scatter into buckets, and a final pass in
// which every record counts the records of ITS bucket that precede it -- the exact descending (hi, lo) order in four
// small launches.  Buckets stay small because the response histogram of an image is smooth; the host falls back to the
// bitonic network when the largest bucket says otherwise (plateaus of equal responses).
constexpr int BS_SHIFT = 14;          // bucket = (max_bits >> 14) - (bits >> 14): 512 buckets per octave of response
constexpr int BS_BINS = 1 << 14;      // 32 octaves below the maximum; a wider range falls back to the bitonic network
constexpr int BS_MAXBUCKET = 2048;    // above this the quadratic last pass is no longer cheap

// bucket of a 32-bit key (descending order): shift = BS_SHIFT for float responses, larger for index keys (FAST)
__device__ __forceinline__ int bs_bucket(unsigned bits, unsigned max_bits, int shift = BS_SHIFT)
{
    return (int)(max_bits >> shift) - (int)(bits >> shift);
}

// histogram of a record list whose producer did not build one (FAST: key = ~pixel index, max key = ~0)
__global__ void __launch_bounds__(256) bs_hist_kernel(const Rec128 *__restrict__ in, int n, unsigned max_key, int shift, int *__restrict__ hist)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) atomicAdd(&hist[bs_bucket((unsigned)in[i].hi, max_key, shift)], 1);
}

// exclusive scan of the histogram (one CTA, the histogram staged in 64 KB of dynamic shared memory with coalesced
// loads; slot i of the staging array is skewed by i / 16 so that the 16 bins a thread scans do not collide in one bank),
// largest bucket -> *max_count
constexpr size_t BS_SCAN_SMEM = sizeof(int) * (BS_BINS + BS_BINS / 16);
__global__ void __launch_bounds__(1024) bs_scan_kernel(const int *__restrict__ hist, int *__restrict__ start, int *__restrict__ max_count)
{
    extern __shared__ int s_h[];
    __shared__ int s_sum[1024];
    __shared__ int s_max[32];
    constexpr int PER = BS_BINS / 1024;   // 16
    const int t = threadIdx.x;
    for (int i = t; i < BS_BINS; i += 1024) s_h[i + i / 16] = hist[i];
    __syncthreads();
    int local = 0, mx = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) { const int c = s_h[t * PER + k + t]; local += c; mx = c > mx ? c : mx; }
    s_sum[t] = local;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {        // Hillis-Steele inclusive scan of the per-thread sums
        const int v = t >= o ? s_sum[t - o] : 0;
        __syncthreads();
        s_sum[t] += v;
        __syncthreads();
    }
    int run = s_sum[t] - local;
#pragma unroll
    for (int k = 0; k < PER; k++) { const int c = s_h[t * PER + k + t]; s_h[t * PER + k + t] = run; run += c; }
    __syncthreads();
    for (int i = t; i < BS_BINS; i += 1024) start[i] = s_h[i + i / 16];
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((t & 31) == 0) s_max[t >> 5] = mx;
    __syncthreads();
    if (t < 32) {
        mx = s_max[t];
        for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (t == 0) atomicMax(max_count, mx);
    }
}

// max_bits: device pointer to the largest key (GFTT: the response maximum found by the response kernel), or nullptr with
// max_key given by value
__global__ void __launch_bounds__(256) bs_scatter_kernel(const Rec128 *__restrict__ in, int n, const int *__restrict__ max_bits,
                                                         const int *__restrict__ start, int *__restrict__ cursor, Rec128 *__restrict__ tmp,
                                                         unsigned max_key = 0u, int shift = BS_SHIFT)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const Rec128 r = in[i];
    const int b = bs_bucket((unsigned)r.hi, max_bits ? (unsigned)*max_bits : max_key, shift);
    tmp[start[b] + atomicAdd(&cursor[b], 1)] = r;
}

__global__ void __launch_bounds__(256) bs_rank_kernel(const Rec128 *__restrict__ tmp, int n, const int *__restrict__ max_bits,
                                                      const int *__restrict__ start, const int *__restrict__ hist, Rec128 *__restrict__ out,
                                                      unsigned max_key = 0u, int shift = BS_SHIFT)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const Rec128 r = tmp[i];
    const int b = bs_bucket((unsigned)r.hi, max_bits ? (unsigned)*max_bits : max_key, shift);
    const int s0 = start[b], c = hist[b];
    int before = 0;
    for (int k = 0; k < c; k++) before += rec_greater(tmp[s0 + k], r) ? 1 : 0;
    out[s0 + before] = r;
}

}  // namespace

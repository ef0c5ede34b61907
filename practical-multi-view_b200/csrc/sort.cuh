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

}  // namespace

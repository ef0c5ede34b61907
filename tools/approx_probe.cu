// approx_probe.cu -- accuracy of the fp64 reciprocal / reciprocal-square-root seeds (rcp.approx.ftz.f64,
// rsqrt.approx.ftz.f64) and of the refined versions the large-problem BA kernels use (ba_rcp_fast, ba_rsqrt_fast in
// csrc/ba_kernels.cuh: one cubic step).  Prints the maximum relative error in units of 2^-53 over
// 2^24 arguments spread over [2^-40, 2^40].   nvcc -arch=sm_100a -o approx_probe tools/approx_probe.cu
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double rcp_seed(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ double rsqrt_seed(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ double rcp_fast(double x)
{
    double y = rcp_seed(x);
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}
__device__ __forceinline__ double rsqrt_fast(double x)
{
    double y = rsqrt_seed(x);
    const double e = fma(-x * y, y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
}

__global__ void probe(double *out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    // argument: 2^(e) * (1 + m), e in [-40, 40), m from a hash of i
    unsigned h = i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const double m = (double)h / 4294967296.0, x = ldexp(1.0 + m, (int)(i % 80u) - 40);
    const double r_true = 1.0 / x, s_true = 1.0 / sqrt(x);   // correctly rounded division; rsqrt reference within 1 ulp
    double e[4] = {fabs(rcp_seed(x) - r_true) / r_true, fabs(rcp_fast(x) - r_true) / r_true,
                   fabs(rsqrt_seed(x) - s_true) / s_true, fabs(rsqrt_fast(x) - s_true) / s_true};
    for (int k = 0; k < 4; k++) {
        double v = e[k];
        for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long *>(out + k), (unsigned long long)__double_as_longlong(v));
    }
}

int main()
{
    double *d, h[4];
    cudaMalloc(&d, sizeof h);
    cudaMemset(d, 0, sizeof h);
    probe<<<(1 << 24) / 256, 256>>>(d);
    if (cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
    const double u = ldexp(1.0, -53);
    printf("{\"rcp_seed_rel\": %.3e, \"rcp_fast_ulp53\": %.2f, \"rsqrt_seed_rel\": %.3e, \"rsqrt_fast_ulp53\": %.2f}\n", h[0], h[1] / u, h[2], h[3] / u);
    return 0;
}

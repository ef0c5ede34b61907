// lat_probe.cu -- dependent-chain latencies (cycles per op) of the instructions on the critical path of the
// banded Cholesky (ba_chol_band.cu), measured with clock64 on one warp.  Build: nvcc -arch=sm_100a -O3 -o lat_probe lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 4096;

template <int OP>
__global__ void probe(double *out, long long *cyc, double seed)
{
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
    __shared__ double sm[64];
    sm[threadIdx.x] = x; sm[threadIdx.x + 32] = y;
    __syncwarp();
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = fma(x, y, 1e-9);
        if (OP == 1) x = x * y;
        if (OP == 2) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);
        if (OP == 3) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x));
        if (OP == 4) asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x));
        if (OP == 5) x = rsqrt(x) + 1.0;
        if (OP == 6) x = 1.0 / x + 1.0;
        if (OP == 7) { int idx = (int)(__double2loint(x) & 31); x = sm[idx] ; }
        if (OP == 8) x = __fmaf_rn((float)x, 1.0000001f, 1e-9f);
        if (OP == 9) { float f = __shfl_sync(0xffffffffu, (float)x, (i + 1) & 31); x = f; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x;
}

template <int OP>
__global__ void tput(double *out, long long *cyc, double seed)
{
    double x[8];
    for (int i = 0; i < 8; i++) x[i] = seed + i + threadIdx.x * 1e-9;
    const double y = 1.0000001;
    __shared__ __align__(16) double sm[256];
    for (int i = threadIdx.x; i < 256; i += 32) sm[i] = seed + i;
    __syncwarp();
    double d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0, d5 = 0, d6 = 0, d7 = 0;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 8; i++) {
        if (OP == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = fma(x[k], y, 1e-9);
        }
        if (OP == 1) {
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = __shfl_sync(0xffffffffu, x[k], (k + i) & 31);
        }
        if (OP == 2) {
#pragma unroll
            for (int k = 0; k < 8; k++) { const double2 v = *reinterpret_cast<const double2 *>(sm + ((i + k) & 63) * 2); x[k] += v.x; }
        }
        if (OP == 3) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(x[0]), "d"(x[1]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d2), "+d"(d3) : "d"(x[2]), "d"(x[3]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d4), "+d"(d5) : "d"(x[4]), "d"(x[5]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d6), "+d"(d7) : "d"(x[6]), "d"(x[7]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(x[1]), "d"(x[0]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d2), "+d"(d3) : "d"(x[3]), "d"(x[2]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d4), "+d"(d5) : "d"(x[5]), "d"(x[4]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d6), "+d"(d7) : "d"(x[7]), "d"(x[6]));
        }
        if (OP == 4) {   // dependent DMMA chain
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(x[0]), "d"(x[1]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(x[2]), "d"(x[3]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(x[4]), "d"(x[5]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(x[6]), "d"(x[7]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(x[1]), "d"(x[0]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(x[3]), "d"(x[2]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(x[5]), "d"(x[4]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(x[7]), "d"(x[6]));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double sacc = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
    for (int i = 0; i < 8; i++) sacc += x[i];
    out[threadIdx.x] = sacc;
}

int main()
{
    double *d; long long *c, h;
    cudaMalloc(&d, 256); cudaMalloc(&c, 8);
    const char *names[] = {"DFMA", "DMUL", "SHFL f64 (2x SHFL)", "rcp.approx.f64 (MUFU.RCP64H)", "rsqrt.approx.f64 (MUFU.RSQ64H)",
                           "rsqrt() + DADD", "1.0/x + DADD", "LDS.64 dependent (cvt+addr+lds)", "cvt+FFMA+cvt", "cvt+SHFL f32+cvt"};
#define RUN(OP) probe<OP><<<1, 32>>>(d, c, 1.5); probe<OP><<<1, 32>>>(d, c, 1.5); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-36s %7.1f cycles/op\n", names[OP], (double)h / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    const char *tn[] = {"DFMA x8 independent", "SHFL f64 x8 independent", "LDS.128 broadcast + DADD x8", "DMMA.8x8x4 x4 accumulators", "DMMA.8x8x4 dependent chain"};
#define RUNT(OP) tput<OP><<<1, 32>>>(d, c, 1.5); tput<OP><<<1, 32>>>(d, c, 1.5); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-36s %7.1f cycles/op (single warp throughput)\n", tn[OP], (double)h / N);
    RUNT(0) RUNT(1) RUNT(2) RUNT(3) RUNT(4)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

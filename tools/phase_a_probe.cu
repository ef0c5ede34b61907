// phase_a_probe.cu -- times phase A of the banded cluster Cholesky (one warp factoring a 32x32 diagonal tile)
// in isolation.  Build on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include \
//        -I practical-multi-view_b200/csrc -o /tmp/phase_a_probe tools/phase_a_probe.cu
#include "../practical-multi-view_b200/csrc/ba_chol_band.cu"

__global__ void probe_kernel(long long *cyc, double *out, BAState *st)
{
    __shared__ __align__(16) double tile[TSZ], bw[NB], dU[NB * ULD], dinv[NB], dz[NB], UT[NB * UTLD + 24 * CSLD];
    const int lane = threadIdx.x;
    for (int i = 0; i < NB; i++) tile[i * TLD + lane] = (i == lane ? 40.0 : 1.0 / (1 + abs(i - lane)));
    bw[lane] = 1.0 + lane;
    __syncwarp();
    for (int rep = 0; rep < 3; rep++) {
        long long t0 = clock64();
        phase_a(tile, dU, dinv, UT, st, lane);
        long long t1 = clock64();
        if (lane == 0) cyc[rep] = t1 - t0;
    }
    out[lane] = dU[lane * ULD + lane] + dinv[lane] + bw[lane] + dz[lane];
}

int main()
{
    long long *c, h[3]; double *o; BAState *st;
    cudaMalloc(&c, 24); cudaMalloc(&o, 256); cudaMalloc(&st, sizeof(BAState));
    probe_kernel<<<1, 32>>>(c, o, st);
    cudaDeviceSynchronize();
    cudaMemcpy(h, c, 24, cudaMemcpyDeviceToHost);
    printf("phase_a alone: %lld %lld %lld cycles (%s)\n", h[0], h[1], h[2], cudaGetErrorString(cudaGetLastError()));
    return 0;
}

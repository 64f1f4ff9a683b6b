// smsp_map.cu -- which warps of a CTA share an SM sub-partition (FP64 pipe)?  Warp 0 runs the TRSM chain, warp X
// streams DMMAs; the chain slows down only when X sits on warp 0's sub-partition.
#include <cstdio>
#include <vector>
#include <cmath>
#include "../gpyrn_b200/csrc/common.cuh"
using namespace gprn;

__global__ void __launch_bounds__(512, 1) map_kernel(const double* A, double* out, long long* cyc, int nrep, int mma_warp) {
    extern __shared__ double smem[];
    double* Bs = smem;
    double* As = smem + NB * LDT;
    double* col = smem + 2 * NB * LDT;
    double* pivs = col + 2 * NB;
    double* rd = pivs + NB;
    __shared__ int bad;
    __shared__ volatile int done;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, w4 = warp & 3;
    const int r = lane >> 2, c = lane & 3;
    if (tid == 0) { bad = 0; done = 0; }
    for (int e = tid; e < NB * NB; e += blockDim.x) { Bs[(e >> 6) * LDT + (e & 63)] = A[e]; As[(e >> 6) * LDT + (e & 63)] = A[e] * 1e-3; }
    __syncthreads();
    if (tid < 256) potrf64_t<8, 4>(Bs, LDT, Bs, rd, col, pivs, &bad, tid);
    __syncthreads();
    double acc[2][8][2];
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++) { acc[x][y][0] = A[(16 * w4 + 8 * x + r) * 64 + 8 * y + 2 * c]; acc[x][y][1] = A[(16 * w4 + 8 * x + r) * 64 + 8 * y + 2 * c + 1]; }
    __syncthreads();
    if (warp == 0) {
        long long t0 = clock64();
        for (int it = 0; it < nrep; it++) trsm_rows_inreg(acc, Bs, rd, lane);
        long long t1 = clock64();
        if (tid == 0) cyc[0] = (t1 - t0) / nrep;
        __threadfence_block();
        if (lane == 0) done = 1;
    } else if (warp == mma_warp) {
        while (!done) mma_slab<true>(acc, As, Bs, w4, lane);
    }
    double s = 0;
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++) s += acc[x][y][0] + acc[x][y][1];
    out[tid] = s;
}

int main() {
    double *d_out; long long* d_cyc; double* d_A;
    cudaMalloc(&d_out, 8 * 1024); cudaMalloc(&d_cyc, 64); cudaMalloc(&d_A, 8 * 64 * 64);
    std::vector<double> A(64 * 64);
    for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) A[i * 64 + j] = std::exp(-0.5 * (i - j) * (i - j) / 400.0) + (i == j ? 0.1 : 0.0);
    cudaMemcpy(d_A, A.data(), 8 * 64 * 64, cudaMemcpyHostToDevice);
    const size_t smem = (2 * NB * LDT + 4 * NB) * sizeof(double);
    cudaFuncSetAttribute(map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int x = 0; x < 16; x++) {
        map_kernel<<<1, 512, smem>>>(d_A, d_out, d_cyc, 20, x == 0 ? 99 : x);
        cudaError_t e = cudaDeviceSynchronize();
        long long cy;
        cudaMemcpy(&cy, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("DMMA stream on warp %2d: TRSM on warp 0 takes %6lld cycles (%s)\n", x == 0 ? -1 : x, cy, cudaGetErrorString(e));
    }
    return 0;
}

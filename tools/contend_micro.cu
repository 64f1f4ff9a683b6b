// contend_micro.cu -- does a DMMA-streaming warp starve a dependent FP64 chain on the same SM sub-partition?
// One CTA, 8 warps: warps 0-3 run trsm_rows_inreg (latency chain), warps 4-7 (same sub-partitions) run mma_slab in a
// loop.  Prints the cycles per TRSM alone, next to the MMA warps, and with the MMA warps yielding between k-steps.
#include <cstdio>
#include <vector>
#include <cmath>
#include "../gpyrn_b200/csrc/common.cuh"
using namespace gprn;

template <int YIELD>
__device__ __forceinline__ void mma_slab_y(double (&acc)[2][8][2], const double* __restrict__ As,
                                           const double* __restrict__ Bs, int w4, int lane) {
    const int r = lane >> 2, c = lane & 3;
    const double* ap = As + (w4 * 16 + r) * LDT + c;
    const double* bp = Bs + r * LDT + c;
#pragma unroll 2
    for (int k0 = 0; k0 < NB; k0 += 4) {
        double a[2], b[8];
#pragma unroll
        for (int x = 0; x < 2; x++) a[x] = -ap[x * 8 * LDT + k0];
#pragma unroll
        for (int y = 0; y < 8; y++) b[y] = bp[y * 8 * LDT + k0];
#pragma unroll
        for (int x = 0; x < 2; x++)
#pragma unroll
            for (int y = 0; y < 8; y++) {
                dmma884(acc[x][y], a[x], b[y]);
                if (YIELD == 1 && (y & 3) == 3) __nanosleep(0);
                if (YIELD == 2 && (y & 1) == 1) asm volatile("nanosleep.u32 20;");
            }
    }
}


// KU consecutive k-steps per accumulator visit: the KU DMMAs on one accumulator are dependent (26 cycles apart), so
// the warp cannot keep the FP64 pipe's queue full
template <int KU, int YH>
__device__ __forceinline__ void mma_slab_dep(double (&acc)[2][8][2], const double* __restrict__ As,
                                             const double* __restrict__ Bs, int w4, int lane) {
    const int r = lane >> 2, c = lane & 3;
    const double* ap = As + (w4 * 16 + r) * LDT + c;
    const double* bp = Bs + r * LDT + c;
#pragma unroll 1
    for (int k0 = 0; k0 < NB; k0 += 4 * KU) {
#pragma unroll
        for (int yh = 0; yh < 8; yh += YH) {
            double a[2][KU], b[YH][KU];
#pragma unroll
            for (int x = 0; x < 2; x++)
#pragma unroll
                for (int u = 0; u < KU; u++) a[x][u] = -ap[x * 8 * LDT + k0 + 4 * u];
#pragma unroll
            for (int y = 0; y < YH; y++)
#pragma unroll
                for (int u = 0; u < KU; u++) b[y][u] = bp[(yh + y) * 8 * LDT + k0 + 4 * u];
#pragma unroll
            for (int x = 0; x < 2; x++)
#pragma unroll
                for (int y = 0; y < YH; y++)
#pragma unroll
                    for (int u = 0; u < KU; u++) dmma884(acc[x][yh + y], a[x][u], b[y][u]);
        }
    }
}

// mode 0: TRSM alone (warps 4-7 idle); 1: + MMA warps; 2: + MMA warps with nanosleep(0) every 4 DMMAs; 3: nanosleep 20 every 2
template <int MODE>
__global__ void __launch_bounds__(256, 1) contend_kernel(const double* A, double* out, long long* cyc, int nrep, volatile int* stop) {
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
    potrf64(Bs, LDT, Bs, rd, col, pivs, &bad);
    __syncthreads();
    double acc[2][8][2];
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++) { acc[x][y][0] = A[(16 * w4 + 8 * x + r) * 64 + 8 * y + 2 * c]; acc[x][y][1] = A[(16 * w4 + 8 * x + r) * 64 + 8 * y + 2 * c + 1]; }
    __syncthreads();
    long long mma_iters = 0;
    if (warp < 4) {
        long long t0 = clock64();
        for (int it = 0; it < nrep; it++) trsm_rows_inreg(acc, Bs, rd, lane);
        long long t1 = clock64();
        if (tid == 0) cyc[0] = (t1 - t0) / nrep;
        __threadfence_block();
        if (lane == 0) atomicAdd((int*)&done, 1);
    } else if (MODE >= 1) {
        long long t0 = clock64();
        while (done < 4) {
            if (MODE == 1) mma_slab_y<0>(acc, As, Bs, w4, lane);
            if (MODE == 2) mma_slab_y<1>(acc, As, Bs, w4, lane);
            if (MODE == 3) mma_slab_y<2>(acc, As, Bs, w4, lane);
            if (MODE == 4) mma_slab_dep<2, 8>(acc, As, Bs, w4, lane);
            if (MODE == 5) mma_slab_dep<4, 4>(acc, As, Bs, w4, lane);
            if (MODE == 6) mma_slab_dep<8, 2>(acc, As, Bs, w4, lane);
            if (MODE == 7) mma_slab_dep<16, 1>(acc, As, Bs, w4, lane);
            mma_iters++;
        }
        long long t1 = clock64();
        if (tid == 128) { cyc[1] = (t1 - t0) / (mma_iters ? mma_iters : 1); cyc[2] = mma_iters; }
    }
    double s = 0;
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++) s += acc[x][y][0] + acc[x][y][1];
    out[tid] = s;
}

template <int V>
__global__ void __launch_bounds__(256, 1) mma_rate_kernel(const double* A, double* out, long long* cyc, int nrep, int nwarps) {
    extern __shared__ double smem[];
    double* Bs = smem;
    double* As = smem + NB * LDT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, w4 = warp & 3;
    for (int e = tid; e < NB * NB; e += blockDim.x) { Bs[(e >> 6) * LDT + (e & 63)] = A[e]; As[(e >> 6) * LDT + (e & 63)] = A[e] * 1e-3; }
    __syncthreads();
    double acc[2][8][2];
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++) acc[x][y][0] = acc[x][y][1] = 0.0;
    long long t0 = clock64();
    if (warp < nwarps)
        for (int it = 0; it < nrep; it++) {
            if (V == 0) mma_slab_y<0>(acc, As, Bs, w4, lane);
            if (V == 1) mma_slab_dep<2, 8>(acc, As, Bs, w4, lane);
            if (V == 2) mma_slab_dep<4, 4>(acc, As, Bs, w4, lane);
            if (V == 3) mma_slab_dep<8, 2>(acc, As, Bs, w4, lane);
            if (V == 4) mma_slab_dep<16, 1>(acc, As, Bs, w4, lane);
        }
    long long t1 = clock64();
    if (tid == 0) cyc[0] = (t1 - t0) / nrep;
    double s = 0;
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++) s += acc[x][y][0] + acc[x][y][1];
    out[tid] = s;
}

int main() {
    double *d_out; long long* d_cyc; double* d_A; int* d_stop;
    cudaMalloc(&d_out, 8 * 1024); cudaMalloc(&d_cyc, 64); cudaMalloc(&d_A, 8 * 64 * 64); cudaMalloc(&d_stop, 4);
    std::vector<double> A(64 * 64);
    for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) A[i * 64 + j] = std::exp(-0.5 * (i - j) * (i - j) / 400.0) + (i == j ? 0.1 : 0.0);
    cudaMemcpy(d_A, A.data(), 8 * 64 * 64, cudaMemcpyHostToDevice);
    const size_t smem = (2 * NB * LDT + 4 * NB) * sizeof(double);
    typedef void (*kfn)(const double*, double*, long long*, int, volatile int*);
    kfn fns[8] = {contend_kernel<0>, contend_kernel<1>, contend_kernel<2>, contend_kernel<3>, contend_kernel<4>, contend_kernel<5>, contend_kernel<6>, contend_kernel<7>};
    const char* nm[8] = {"TRSM alone", "TRSM next to MMA warps", "... MMA warps nanosleep(0) every 4 DMMAs", "... MMA warps nanosleep(20) every 2 DMMAs",
                         "... MMA: 2 dependent k-steps per accumulator", "... MMA: 4 dependent k-steps per accumulator", "... MMA: 8 dependent k-steps", "... MMA: 16 dependent k-steps"};
    for (int m = 0; m < 8; m++) {
        cudaFuncSetAttribute(fns[m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaMemset(d_cyc, 0, 64);
        fns[m]<<<1, 256, smem>>>(d_A, d_out, d_cyc, 20, d_stop);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc[3];
        cudaMemcpy(cyc, d_cyc, 24, cudaMemcpyDeviceToHost);
        printf("%-44s (%s): %7lld cycles per TRSM; MMA warps: %7lld cycles per 64x64x64 slab product, %lld products\n", nm[m], cudaGetErrorString(e), cyc[0], cyc[1], cyc[2]);
    }
    typedef void (*rfn)(const double*, double*, long long*, int, int);
    rfn rf[5] = {mma_rate_kernel<0>, mma_rate_kernel<1>, mma_rate_kernel<2>, mma_rate_kernel<3>, mma_rate_kernel<4>};
    const char* rn[5] = {"independent order", "2 dependent k-steps", "4 dependent k-steps", "8 dependent k-steps", "16 dependent k-steps"};
    for (int v = 0; v < 5; v++)
        for (int nw : {4, 8}) {
            cudaFuncSetAttribute(rf[v], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            rf[v]<<<1, 256, smem>>>(d_A, d_out, d_cyc, 50, nw);
            cudaDeviceSynchronize();
            long long cy;
            cudaMemcpy(&cy, d_cyc, 8, cudaMemcpyDeviceToHost);
            printf("MMA alone, %-22s %d warps: %6lld cycles per slab product (ideal %d)\n", rn[v], nw, cy, nw == 4 ? 4288 : 8576);
        }
    return 0;
}

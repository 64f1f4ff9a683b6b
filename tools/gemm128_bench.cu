// Micro-benchmark of the 128x128 DMMA GEMM core (gpyrn_b200/csrc/gemm128.cuh) in isolation: one CTA per SM (x waves),
// each forming a 128x128 tile over K, operands L2 resident.  Prints TFLOP/s against K, with and without an epilogue
// that reads / writes C, so that main-loop efficiency can be separated from prologue / epilogue cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -I gpyrn_b200/csrc -o tools/gemm128_bench tools/gemm128_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gemm128.cuh"
using namespace gprn;

template <bool KMAJOR, bool EPI>
__global__ void __launch_bounds__(G_THREADS, G_MINB) bench_kernel(const double* A, const double* B, double* C, int K, int ld) {
    extern __shared__ double smem[];
    double acc[G_MI][G_NI][2];
#pragma unroll
    for (int i = 0; i < G_MI; i++)
#pragma unroll
        for (int j = 0; j < G_NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int t = blockIdx.x % 8;      // 8 different operand panels
    gemm128_mainloop<KMAJOR>(acc, smem, A + (size_t)t * G_BM * ld, ld, KMAJOR ? B + t * 128 : B + (size_t)t * 128 * ld, ld, K);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp / G_WARPS_N, wn = warp % G_WARPS_N, r = lane >> 2, c = lane & 3;
    double* Ct = C + (size_t)blockIdx.x * G_BM * 128;
#pragma unroll
    for (int i = 0; i < G_MI; i++)
#pragma unroll
        for (int j = 0; j < G_NI; j++) {
            double2* p = reinterpret_cast<double2*>(Ct + (size_t)(wm * G_WM + i * 8 + r) * 128 + wn * G_WN + j * 8 + 2 * c);
            double2 v = EPI ? *p : make_double2(0.0, 0.0);
            v.x -= acc[i][j][0];
            v.y -= acc[i][j][1];
            *p = v;
        }
}

template <bool KMAJOR, bool EPI>
static void run(const double* A, const double* B, double* C, int K, int ld, int ctas, const char* name) {
    cudaFuncSetAttribute(bench_kernel<KMAJOR, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM128_SMEM);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) bench_kernel<KMAJOR, EPI><<<ctas, G_THREADS, GEMM128_SMEM>>>(A, B, C, K, ld);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int w = 0; w < reps; w++) bench_kernel<KMAJOR, EPI><<<ctas, G_THREADS, GEMM128_SMEM>>>(A, B, C, K, ld);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * G_BM * 128 * (double)K * ctas * reps;
    printf("%-22s K=%5d ctas=%5d  %8.1f us/launch  %6.2f TFLOP/s  (%s)\n", name, K, ctas, ms * 1e3 / reps, fl / (ms * 1e-3) * 1e-12,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int ld = 4096;
    double *A, *B, *C;
    cudaMalloc(&A, sizeof(double) * 8 * 128 * ld);
    cudaMalloc(&B, sizeof(double) * (size_t)ld * ld);
    cudaMalloc(&C, sizeof(double) * 148 * 32 * 128 * 128);
    std::vector<double> h((size_t)ld * ld);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)((i * 2654435761u) % 1000) * 1e-3;
    cudaMemcpy(A, h.data(), sizeof(double) * 8 * 128 * ld, cudaMemcpyHostToDevice);
    cudaMemcpy(B, h.data(), sizeof(double) * (size_t)ld * ld, cudaMemcpyHostToDevice);
    cudaMemset(C, 0, sizeof(double) * 148 * 32 * 128 * 128);
    for (int K : {256, 512, 1024, 4096}) {
        const int per = 148 * G_MINB * (128 / G_BM);
        run<false, true>(A, B, C, K, ld, per, "NT  C rmw,    1 wave");
        run<false, true>(A, B, C, K, ld, per * 4, "NT  C rmw,    4 waves");
        run<true, false>(A, B, C, K, ld, per * 4, "NN  store,    4 waves");
    }
    return 0;
}

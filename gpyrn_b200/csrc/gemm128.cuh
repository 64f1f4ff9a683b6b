// gemm128.cuh -- FP64 tensor-core GEMM core for the large-N path: one CTA computes a G_BM x 128 tile of
//     acc = sum_k A[m][k] * B(k, n)
// on DMMA m8n8k4 with a multi-stage cp.async (LDGSTS, 16-byte) shared-memory pipeline.
//   A : row-major, k contiguous ("[m][k]").
//   B : !B_KMAJOR  row-major [n][k], k contiguous   (NT product, e.g. L_ik L_jk^T)
//        B_KMAJOR  row-major [k][n], n contiguous   (NN product, e.g. L_ik X_kj)
// Default geometry: 64 x 128 CTA tile, 128 threads = 2 (M) x 2 (N) warps of 32 x 64 (64 accumulator doubles per
// thread, 12 fragment loads for 32 DMMAs per k-step of 4), 3 stages of 16, TWO CTAs per SM (92 KB of shared memory
// and < 256 registers each): while one CTA sits in its prologue, epilogue or at the k-step barrier the other one
// keeps the DMMA pipe fed.  Measured in isolation (tools/gemm128_bench.cu, L2-resident operands, B200): 35.0 TFLOP/s
// at K = 4096 (94 % of the 37.09 DMMA peak; cuBLAS DGEMM 8192^3: 35.5), 32.7 at K = 512, 30.4 at K = 256.  The
// earlier 128 x 128 tile with 512 threads, 4 stages and ONE CTA per SM (-DG_BM=128 -DG_WARPS_M=4 -DG_WARPS_N=4
// -DG_STAGES=4 -DG_MINB=1) reached 30-32 / 28.7 / 27.1: its per-tile prologue + epilogue (6 us) and barrier bubbles
// were exposed.  Shared-memory strides (20 / 132 doubles = 32 B mod 128 B) make the fragment loads of a half-warp hit
// 16 distinct 8-byte slots.
#pragma once
#include "common.cuh"

namespace gprn {

#ifndef G_BM
#define G_BM 64                    // CTA tile rows
#endif
#define G_BN 128                   // CTA tile columns
#ifndef G_BK
#define G_BK 16
#endif
#define G_KCH (G_BK / 2)            // 16-byte chunks per operand row and k-slab
#ifndef G_WARPS_M
#define G_WARPS_M 2
#endif
#ifndef G_WARPS_N
#define G_WARPS_N 2
#endif
#define G_THREADS (32 * G_WARPS_M * G_WARPS_N)
#define G_WM (G_BM / G_WARPS_M)     // warp tile rows
#define G_WN (G_BN / G_WARPS_N)     // warp tile columns
#define G_MI (G_WM / 8)             // m8 tiles per warp
#define G_NI (G_WN / 8)             // n8 tiles per warp
#ifndef G_STAGES
#define G_STAGES 3
#endif
#ifndef G_MINB
#define G_MINB 2                   // CTAs per SM the register budget is sized for
#endif
#define G_LDA (G_BK + 4)            // As[m][k]; (BK+4)*8 B = 32 (mod 128) for BK = 16, 32
#define G_LDB_NT (G_BK + 4)         // Bs[n][k]
#define G_LDB_NN 132                // Bs[k][n]
#define G_A_STAGE (G_BM * G_LDA)    // doubles
#define G_B_STAGE_NT (G_BN * G_LDB_NT)
#define G_B_STAGE_NN (G_BK * G_LDB_NN)
#define GEMM128_SMEM (G_STAGES * (G_A_STAGE + G_B_STAGE_NT) * sizeof(double))   // NT is the larger one

// Issue the loads of one k-slab (16 deep) into one pipeline stage.
template <bool B_KMAJOR>
__device__ __forceinline__ void gemm128_load_stage(double* As, double* Bs, const double* __restrict__ A, size_t lda,
                                                   const double* __restrict__ B, size_t ldb, int k0, int tid) {
    // A: 128 rows x 8 chunks of 2 doubles
#pragma unroll
    for (int u = 0; u < G_BM * G_KCH / G_THREADS; u++) {
        int ch = tid + G_THREADS * u, row = ch / G_KCH, kc = ch % G_KCH;
        cp_async16(As + row * G_LDA + 2 * kc, A + (size_t)row * lda + k0 + 2 * kc);
    }
    if (!B_KMAJOR) {
#pragma unroll
        for (int u = 0; u < G_BN * G_KCH / G_THREADS; u++) {
            int ch = tid + G_THREADS * u, row = ch / G_KCH, kc = ch % G_KCH;
            cp_async16(Bs + row * G_LDB_NT + 2 * kc, B + (size_t)row * ldb + k0 + 2 * kc);
        }
    } else {
#pragma unroll
        for (int u = 0; u < G_BK * 64 / G_THREADS; u++) {
            int ch = tid + G_THREADS * u, k = ch >> 6, nc = ch & 63;
            cp_async16(Bs + k * G_LDB_NN + 2 * nc, B + (size_t)(k0 + k) * ldb + 2 * nc);
        }
    }
}

// acc += A(128 x K) * B over k in [0, K), K a multiple of 16.  A / B point at the k = 0 corner of
// the CTA's row / column panel.  smem: GEMM128_SMEM bytes.  All G_THREADS threads call.
// Accumulator (i, j, e) of a thread is element (wm*G_WM + i*8 + lane/4, wn*G_WN + j*8 + 2*(lane%4) + e) of the tile.
template <bool B_KMAJOR>
__device__ __forceinline__ void gemm128_mainloop(double (&acc)[G_MI][G_NI][2], double* smem, const double* __restrict__ A,
                                                 size_t lda, const double* __restrict__ B, size_t ldb, int K) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp / G_WARPS_N, wn = warp % G_WARPS_N;
    const int r = lane >> 2, c = lane & 3;
    constexpr int BST = B_KMAJOR ? G_B_STAGE_NN : G_B_STAGE_NT;
    double* As0 = smem;
    double* Bs0 = smem + G_STAGES * G_A_STAGE;
    const int nk = K / G_BK;
#pragma unroll
    for (int s = 0; s < G_STAGES - 1; s++) {
        if (s < nk) gemm128_load_stage<B_KMAJOR>(As0 + s * G_A_STAGE, Bs0 + s * BST, A, lda, B, ldb, s * G_BK, tid);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<G_STAGES - 2>();
        __syncthreads();
        {
            const int kn = kt + G_STAGES - 1;
            if (kn < nk) {
                const int sn = kn % G_STAGES;
                gemm128_load_stage<B_KMAJOR>(As0 + sn * G_A_STAGE, Bs0 + sn * BST, A, lda, B, ldb, kn * G_BK, tid);
            }
            cp_async_commit();
        }
        const int st = kt % G_STAGES;
        const double* as = As0 + st * G_A_STAGE + (wm * G_WM + r) * G_LDA + c;
        const double* bs = B_KMAJOR ? Bs0 + st * BST + c * G_LDB_NN + wn * G_WN + r
                                    : Bs0 + st * BST + (wn * G_WN + r) * G_LDB_NT + c;
#pragma unroll
        for (int kk = 0; kk < G_BK; kk += 4) {
            double a[G_MI], b[G_NI];
#pragma unroll
            for (int i = 0; i < G_MI; i++) a[i] = as[i * 8 * G_LDA + kk];
#pragma unroll
            for (int j = 0; j < G_NI; j++) b[j] = B_KMAJOR ? bs[kk * G_LDB_NN + j * 8] : bs[j * 8 * G_LDB_NT + kk];
#pragma unroll
            for (int i = 0; i < G_MI; i++)
#pragma unroll
                for (int j = 0; j < G_NI; j++) dmma884(acc[i][j], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

}  // namespace gprn

// gemm128.cuh -- FP64 tensor-core GEMM core for the large-N path: one CTA computes a 128x128 tile of
//     acc = sum_k A[m][k] * B(k, n)
// on DMMA m8n8k4 with a multi-stage cp.async (LDGSTS, 16-byte) shared-memory pipeline.
//   A : row-major, k contiguous ("[m][k]").
//   B : !B_KMAJOR  row-major [n][k], k contiguous   (NT product, e.g. L_ik L_jk^T)
//        B_KMAJOR  row-major [k][n], n contiguous   (NN product, e.g. L_ik X_kj)
// G_THREADS = 512: 16 warps laid out 4 (M) x 4 (N), each owning 32x32 = 4x4 DMMA accumulator tiles (32 doubles
// per thread; 4 warps per scheduler hide the DMMA / shared-memory latencies better than the 8-warp, 64x32
// variant, which measured 71 % DMMA-pipe utilisation).  Per k-step of 4 a warp issues 8 fragment loads for 16 DMMAs.  Shared-memory strides (20 / 132 doubles = 32 B mod 128 B) make the fragment loads of a
// half-warp hit 16 distinct 8-byte slots.
#pragma once
#include "common.cuh"

namespace gprn {

#define G_BM 128
#define G_BN 128
#ifndef G_BK
#define G_BK 16
#endif
#define G_KCH (G_BK / 2)            // 16-byte chunks per operand row and k-slab
#ifndef G_THREADS
#define G_THREADS 512             // 16 warps, 4 (M) x 4 (N), 32x32 accumulator tile per warp
#endif
#define G_WARPS_N 4
#define G_WARPS_M (G_THREADS / 32 / G_WARPS_N)
#define G_WM (G_BM / G_WARPS_M)     // rows per warp: 32 (512 threads) or 64 (256 threads)
#define G_MI (G_WM / 8)             // m8 tiles per warp
#ifndef G_STAGES
#define G_STAGES 4
#endif
#define G_LDA (G_BK + 4)            // As[m][k]; (BK+4)*8 B = 32 (mod 128) for BK = 16, 32
#define G_LDB_NT (G_BK + 4)         // Bs[n][k]
#define G_LDB_NN 132                // Bs[k][n]
#define G_A_STAGE (G_BM * G_LDA)    // doubles
#define G_B_STAGE_NT (G_BN * G_LDB_NT)
#define G_B_STAGE_NN (G_BK * G_LDB_NN)
#define GEMM128_SMEM (G_STAGES * (G_A_STAGE + G_B_STAGE_NT) * sizeof(double))   // NT is the larger one

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

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
// Accumulator (i, j, e) of a thread is element (wm*G_WM + i*8 + lane/4, wn*32 + j*8 + 2*(lane%4) + e) of the tile.
template <bool B_KMAJOR>
__device__ __forceinline__ void gemm128_mainloop(double (&acc)[G_MI][4][2], double* smem, const double* __restrict__ A,
                                                 size_t lda, const double* __restrict__ B, size_t ldb, int K) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;
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
        const double* bs = B_KMAJOR ? Bs0 + st * BST + c * G_LDB_NN + wn * 32 + r
                                    : Bs0 + st * BST + (wn * 32 + r) * G_LDB_NT + c;
#pragma unroll
        for (int kk = 0; kk < G_BK; kk += 4) {
            double a[G_MI], b[4];
#pragma unroll
            for (int i = 0; i < G_MI; i++) a[i] = as[i * 8 * G_LDA + kk];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = B_KMAJOR ? bs[kk * G_LDB_NN + j * 8] : bs[j * 8 * G_LDB_NT + kk];
#pragma unroll
            for (int i = 0; i < G_MI; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

}  // namespace gprn

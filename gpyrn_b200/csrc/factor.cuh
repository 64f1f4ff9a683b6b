// factor.cuh -- batched blocked Cholesky and triangular inverse (FP64 compute bound).
//
// Right-looking blocked Cholesky with 64x64 tiles, one launch per role and block step, batched over
// an arbitrary list of matrices (blockIdx.y):
//     potrf_diag   : L_kk = chol(A_kk)                 (shared memory, one CTA per matrix)
//     trsm_panel   : L_ik = A_ik L_kk^-T               (forward substitution, one row per thread)
//     syrk_update  : A_ij -= L_ik L_jk^T               (DMMA m8n8k4 tiles, the only dense contraction)
// and the inverse factor X = L^-1 by block rows:
//     trtri_diag   : X_ii = L_ii^-1
//     trtri_row    : X_ij = -L_ii^-1 * sum_{k=j}^{i-1} L_ik X_kj     (DMMA accumulate + substitution)
// Replaces jnp.linalg.cholesky (gpyrn/meanfield.py:88), np.linalg.solve (:771,:850),
// cho_solve (:1032-1051) and scipy cho_factor/cho_solve (gpyrn/_gp.py:126-135).
#pragma once
#include "common.cuh"
#include "gemm128.cuh"

namespace gprn {

// W_tiles(I>=J) = K_tiles, plus dvec on the diagonal (dvec may be null).  grid = (lower tiles, nmat).
__global__ void __launch_bounds__(256) form_a_kernel(double* __restrict__ W, const double* __restrict__ K,
                                                     const double* __restrict__ dvec, const int* __restrict__ ids,
                                                     int Np) {
    int I, J;
    tri_decode(blockIdx.x, I, J);
    const int id = ids[blockIdx.y];
    const size_t base = (size_t)id * Np * Np + (size_t)(I * NB) * Np + J * NB;
    const double* src = K + base;
    double* dst = W + base;
    const double* dv = dvec ? dvec + (size_t)id * Np + I * NB : nullptr;
    for (int e = threadIdx.x; e < NB * (NB / 2); e += 256) {
        int r = e >> 5, c2 = e & 31;
        double2 v = *reinterpret_cast<const double2*>(src + (size_t)r * Np + 2 * c2);
        if (dv && I == J) {
            if (2 * c2 == r) v.x += dv[r];
            if (2 * c2 + 1 == r) v.y += dv[r];
        }
        *reinterpret_cast<double2*>(dst + (size_t)r * Np + 2 * c2) = v;
    }
}

// Cholesky of diagonal tile k of every listed matrix; accumulates 2*sum(log diag) into logdet[id]
// and raises status[id] when a pivot is not positive.  grid = (nmat), block = 256.
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ W, const int* __restrict__ ids, int Np,
                                                         int k, double* __restrict__ logdet,
                                                         int* __restrict__ status) {
    __shared__ double T[NB * LDV];
    __shared__ int bad;
    const int id = ids[blockIdx.x];
    double* A = W + (size_t)id * Np * Np + (size_t)(k * NB) * Np + k * NB;
    const int tid = threadIdx.x;
    if (tid == 0) bad = 0;
    for (int e = tid; e < NB * NB; e += 256) {
        int r = e >> 6, c = e & 63;
        T[r * LDV + c] = A[(size_t)r * Np + c];
    }
    __syncthreads();
    for (int c = 0; c < NB; c++) {
        if (tid == 0) {
            double piv = T[c * LDV + c];
            if (!(piv > 0.0)) bad = 1;
            T[c * LDV + c] = sqrt(piv);
        }
        __syncthreads();
        const double d = T[c * LDV + c];
        if (tid > c && tid < NB) T[tid * LDV + c] = T[tid * LDV + c] / d;
        __syncthreads();
        for (int e = (c + 1) * NB + tid; e < NB * NB; e += 256) {
            int r = e >> 6, cc = e & 63;
            if (cc > c && cc <= r) T[r * LDV + cc] = fma(-T[r * LDV + c], T[cc * LDV + c], T[r * LDV + cc]);
        }
        __syncthreads();
    }
    for (int e = tid; e < NB * NB; e += 256) {
        int r = e >> 6, c = e & 63;
        A[(size_t)r * Np + c] = (c <= r) ? T[r * LDV + c] : 0.0;
    }
    if (tid < 32) {
        double s = log(T[tid * LDV + tid]) + log(T[(tid + 32) * LDV + tid + 32]);
        s = warp_sum(s);
        if (tid == 0) {
            atomicAdd(&logdet[id], 2.0 * s);
            if (bad) status[id] = 1;
        }
    }
}

// Panel solve below diagonal tile k: rows of tiles i = k+1+2*blockIdx.x (+1).  grid = (ceil((nt-k-1)/2), nmat),
// block = 128 (one matrix row per thread).  Shared memory: Ls[64*LDT] + V[64*129].
#define TRSM_LDV 129
#define TRSM_SMEM ((NB * LDT + NB * TRSM_LDV) * sizeof(double))
__global__ void __launch_bounds__(128) trsm_panel_kernel(double* __restrict__ W, const int* __restrict__ ids, int Np,
                                                         int k) {
    extern __shared__ double smem[];
    double* Ls = smem;
    double* V = smem + NB * LDT;
    const int id = ids[blockIdx.y];
    const int nt = Np / NB;
    double* Wm = W + (size_t)id * Np * Np;
    const int row0 = (k + 1 + 2 * blockIdx.x) * NB;
    const int nrows = min(2 * NB, Np - row0);
    const int tid = threadIdx.x;
    (void)nt;
    load_tile<false>(Ls, Wm + (size_t)(k * NB) * Np + k * NB, Np, tid, 128);
    // A rows -> V[c][r] (vector index = row)
    for (int e = tid; e < nrows * NB; e += 128) {
        int r = e >> 6, c = e & 63;
        V[c * TRSM_LDV + r] = Wm[(size_t)(row0 + r) * Np + k * NB + c];
    }
    __syncthreads();
    if (tid < nrows) subst_lower(Ls, LDT, V, TRSM_LDV, tid);
    __syncthreads();
    for (int e = tid; e < nrows * NB; e += 128) {
        int r = e >> 6, c = e & 63;
        Wm[(size_t)(row0 + r) * Np + k * NB + c] = V[c * TRSM_LDV + r];
    }
}

// Trailing update after panel k: A_ij -= L_ik L_jk^T for k < j <= i.  grid = (n(n+1)/2 with n = nt-k-1, nmat),
// block = 128 (2x2 warps, 32x32 each).  Dynamic shared memory 2*TILE_SMEM.
__global__ void __launch_bounds__(128) syrk_update_kernel(double* __restrict__ W, const int* __restrict__ ids, int Np,
                                                          int k) {
    extern __shared__ double smem[];
    double* As = smem;
    double* Bs = smem + NB * LDT;
    int ti, tj;
    tri_decode(blockIdx.x, ti, tj);
    const int I = k + 1 + ti, J = k + 1 + tj;
    const int id = ids[blockIdx.y];
    double* Wm = W + (size_t)id * Np * Np;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    load_tile<false>(As, Wm + (size_t)(I * NB) * Np + k * NB, Np, tid, 128);
    if (I != J) load_tile<false>(Bs, Wm + (size_t)(J * NB) * Np + k * NB, Np, tid, 128);
    const double* Bp = (I != J) ? Bs : As;
    double* C = Wm + (size_t)(I * NB) * Np + J * NB;
    double acc[4][4][2];
    const int r = lane >> 2, c = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2 v = *reinterpret_cast<const double2*>(C + (size_t)(wm * 32 + i * 8 + r) * Np + wn * 32 + j * 8 + 2 * c);
            acc[i][j][0] = v.x;
            acc[i][j][1] = v.y;
        }
    __syncthreads();
    mma_tile<true>(acc, As, Bp, wm, wn, lane);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<double2*>(C + (size_t)(wm * 32 + i * 8 + r) * Np + wn * 32 + j * 8 + 2 * c) = v;
        }
}

// X_ii = L_ii^-1 for every diagonal tile.  grid = (nt, nmat), block = 64 (one column per thread).
#define TRTRI_DIAG_SMEM ((NB * LDT + NB * LDV) * sizeof(double))
__global__ void __launch_bounds__(64) trtri_diag_kernel(double* __restrict__ X, const double* __restrict__ W,
                                                        const int* __restrict__ ids, int Np) {
    extern __shared__ double smem[];
    double* Ls = smem;
    double* V = smem + NB * LDT;
    const int id = ids[blockIdx.y], i = blockIdx.x, tid = threadIdx.x;
    const size_t off = (size_t)id * Np * Np + (size_t)(i * NB) * Np + i * NB;
    load_tile<false>(Ls, W + off, Np, tid, 64);
    for (int m = 0; m < NB; m++) V[m * LDV + tid] = (m == tid) ? 1.0 : 0.0;
    __syncthreads();
    subst_lower(Ls, LDT, V, LDV, tid, tid >> 3);
    __syncthreads();
    double* Xt = X + off;
    for (int m = 0; m < NB; m++) Xt[(size_t)m * Np + tid] = V[m * LDV + tid];
}

// Block row i of the inverse: X_ij for j = blockIdx.x < i.  grid = (i, nmat), block = 128.
// Dynamic shared memory: As, Bs (MMA operands) + V (64*LDV).
#define TRTRI_SMEM (2 * TILE_SMEM + NB * LDV * sizeof(double))
__global__ void __launch_bounds__(128) trtri_row_kernel(double* __restrict__ X, const double* __restrict__ W,
                                                        const int* __restrict__ ids, int Np, int i) {
    extern __shared__ double smem[];
    double* As = smem;
    double* Bs = smem + NB * LDT;
    double* V = smem + 2 * NB * LDT;
    const int id = ids[blockIdx.y], j = blockIdx.x;
    const double* Wm = W + (size_t)id * Np * Np;
    double* Xm = X + (size_t)id * Np * Np;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int k = j; k < i; k++) {
        load_tile<false>(As, Wm + (size_t)(i * NB) * Np + k * NB, Np, tid, 128);   // L_ik[m][kk]
        load_tile<true>(Bs, Xm + (size_t)(k * NB) * Np + j * NB, Np, tid, 128);    // Bs[n][kk] = X_kj[kk][n]
        __syncthreads();
        mma_tile<true>(acc, As, Bs, wm, wn, lane);                                  // acc = -sum L_ik X_kj
        __syncthreads();
    }
    const int r = lane >> 2, c = lane & 3;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int m = wm * 32 + a * 8 + r, n = wn * 32 + b * 8 + 2 * c;
            V[m * LDV + n] = acc[a][b][0];
            V[m * LDV + n + 1] = acc[a][b][1];
        }
    load_tile<false>(As, Wm + (size_t)(i * NB) * Np + i * NB, Np, tid, 128);        // L_ii
    __syncthreads();
    if (tid < NB) subst_lower(As, LDT, V, LDV, tid);
    __syncthreads();
    double* Xt = Xm + (size_t)(i * NB) * Np + j * NB;
    for (int e = tid; e < NB * NB; e += 128) {
        int m = e >> 6, n = e & 63;
        Xt[(size_t)m * Np + n] = V[m * LDV + n];
    }
}

// ------------------------------------------------------------------------------------------------
// Large-N path (Np a multiple of 256): two-level blocking.  Panels of 256 columns are factored with the
// 64-tile kernels above (syrk restricted to the panel's own columns); the trailing matrix is then updated
// once per panel with 128x128 tiles and a K = 256 deep DMMA product (gemm128.cuh).
// ------------------------------------------------------------------------------------------------
#define OUTER_KB 256

// In-panel trailing update after tile column k: A_ij -= L_ik L_jk^T for J in (k, jend), I in [J, nt).
// grid = ((nt-k-1) * (jend-k-1), nmat), block = 128.
__global__ void __launch_bounds__(128) syrk_inpanel_kernel(double* __restrict__ W, const int* __restrict__ ids, int Np,
                                                           int k, int jend) {
    extern __shared__ double smem[];
    double* As = smem;
    double* Bs = smem + NB * LDT;
    const int ncols = jend - k - 1;
    const int I = k + 1 + blockIdx.x / ncols, J = k + 1 + blockIdx.x % ncols;
    if (J > I) return;
    const int id = ids[blockIdx.y];
    double* Wm = W + (size_t)id * Np * Np;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    load_tile<false>(As, Wm + (size_t)(I * NB) * Np + k * NB, Np, tid, 128);
    if (I != J) load_tile<false>(Bs, Wm + (size_t)(J * NB) * Np + k * NB, Np, tid, 128);
    const double* Bp = (I != J) ? Bs : As;
    double* C = Wm + (size_t)(I * NB) * Np + J * NB;
    double acc[4][4][2];
    const int r = lane >> 2, c = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2 v = *reinterpret_cast<const double2*>(C + (size_t)(wm * 32 + i * 8 + r) * Np + wn * 32 + j * 8 + 2 * c);
            acc[i][j][0] = v.x;
            acc[i][j][1] = v.y;
        }
    __syncthreads();
    mma_tile<true>(acc, As, Bp, wm, wn, lane);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<double2*>(C + (size_t)(wm * 32 + i * 8 + r) * Np + wn * 32 + j * 8 + 2 * c) = v;
        }
}

// Trailing update after the 256-column panel starting at column c0: for 128x128 tiles (TI >= TJ) of the
// trailing square starting at t0 = c0 + 256:  C -= L[rows, c0:c0+256] L[cols, c0:c0+256]^T.
// grid = (n(n+1)/2 with n = (Np - t0)/128, nmat), block = 256, dynamic smem GEMM128_SMEM.
__global__ void __launch_bounds__(256) syrk_outer_kernel(double* __restrict__ W, const int* __restrict__ ids, int Np,
                                                         int c0) {
    extern __shared__ double smem[];
    int TI, TJ;
    tri_decode(blockIdx.x, TI, TJ);
    const int t0 = c0 + OUTER_KB;
    const int r0 = t0 + TI * G_BM, n0 = t0 + TJ * G_BN;
    const int id = ids[blockIdx.y];
    double* Wm = W + (size_t)id * Np * Np;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    gemm128_mainloop<false>(acc, smem, Wm + (size_t)r0 * Np + c0, Np, Wm + (size_t)n0 * Np + c0, Np, OUTER_KB);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 2, wn = warp & 3;
    const int r = lane >> 2, c = lane & 3;
    double* C = Wm + (size_t)r0 * Np + n0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2* p = reinterpret_cast<double2*>(C + (size_t)(wm * 64 + i * 8 + r) * Np + wn * 32 + j * 8 + 2 * c);
            double2 v = *p;
            v.x -= acc[i][j][0];
            v.y -= acc[i][j][1];
            *p = v;
        }
}

// Inverse, block row of 256 rows starting at R0: X[rows, 0:R0] = - L[rows, n0:R0] X[n0:R0, cols] (the part of
// the row-sweep sum that lies above the block), 128x128 tiles.  X must hold zeros in its upper tiles.
// grid = (2 * R0/128, nmat), block = 256, dynamic smem GEMM128_SMEM.
__global__ void __launch_bounds__(256) trtri_outer_kernel(double* __restrict__ X, const double* __restrict__ W,
                                                          const int* __restrict__ ids, int Np, int R0) {
    extern __shared__ double smem[];
    const int ncol = R0 / G_BN;
    const int TJ = blockIdx.x % ncol, half = blockIdx.x / ncol;
    const int r0 = R0 + half * G_BM, n0 = TJ * G_BN;
    const int id = ids[blockIdx.y];
    const double* Wm = W + (size_t)id * Np * Np;
    double* Xm = X + (size_t)id * Np * Np;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    gemm128_mainloop<true>(acc, smem, Wm + (size_t)r0 * Np + n0, Np, Xm + (size_t)n0 * Np + n0, Np, R0 - n0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 2, wn = warp & 3;
    const int r = lane >> 2, c = lane & 3;
    double* C = Xm + (size_t)r0 * Np + n0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2 v = make_double2(-acc[i][j][0], -acc[i][j][1]);
            *reinterpret_cast<double2*>(C + (size_t)(wm * 64 + i * 8 + r) * Np + wn * 32 + j * 8 + 2 * c) = v;
        }
}

// Inverse, in-block sweep for the 4 tile rows i = i0 .. i0+3 of a 256-row block: one CTA per 64-column tile j
// walks the rows in order,  X_ij = L_ii^-1 ( G'_ij - sum_{k=max(j,i0)}^{i-1} L_ik X_kj ),  with G' the partial
// sum left in place by trtri_outer_kernel (zero for columns inside the block).  grid = (i0 + 3, nmat),
// block = 128, dynamic smem TRTRI_SMEM.
__global__ void __launch_bounds__(128) trtri_inblock_kernel(double* __restrict__ X, const double* __restrict__ W,
                                                            const int* __restrict__ ids, int Np, int i0) {
    extern __shared__ double smem[];
    double* As = smem;
    double* Bs = smem + NB * LDT;
    double* V = smem + 2 * NB * LDT;
    const int id = ids[blockIdx.y], j = blockIdx.x;
    const double* Wm = W + (size_t)id * Np * Np;
    double* Xm = X + (size_t)id * Np * Np;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int r = lane >> 2, c = lane & 3;
    const int kstart = max(j, i0);
    for (int i = max(i0, j + 1); i < i0 + 4; i++) {
        double acc[4][4][2];
        double* Xt = Xm + (size_t)(i * NB) * Np + j * NB;
        if (j < i0) {
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    double2 v = *reinterpret_cast<const double2*>(Xt + (size_t)(wm * 32 + a * 8 + r) * Np + wn * 32 + b * 8 + 2 * c);
                    acc[a][b][0] = v.x;
                    acc[a][b][1] = v.y;
                }
        } else {
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
        }
        for (int k = kstart; k < i; k++) {
            load_tile<false>(As, Wm + (size_t)(i * NB) * Np + k * NB, Np, tid, 128);
            load_tile<true>(Bs, Xm + (size_t)(k * NB) * Np + j * NB, Np, tid, 128);
            __syncthreads();
            mma_tile<true>(acc, As, Bs, wm, wn, lane);
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                int m = wm * 32 + a * 8 + r, n = wn * 32 + b * 8 + 2 * c;
                V[m * LDV + n] = acc[a][b][0];
                V[m * LDV + n + 1] = acc[a][b][1];
            }
        load_tile<false>(As, Wm + (size_t)(i * NB) * Np + i * NB, Np, tid, 128);
        __syncthreads();
        if (tid < NB) subst_lower(As, LDT, V, LDV, tid);
        __syncthreads();
        for (int e = tid; e < NB * NB; e += 128) {
            int m = e >> 6, n = e & 63;
            Xt[(size_t)m * Np + n] = V[m * LDV + n];
        }
        __threadfence_block();
        __syncthreads();     // X_ij is read back (as a B operand) by this CTA for the next rows
    }
}

// z = X v  (X lower triangular, row-major).  One warp per row; grid = (Np/8, nmat), block = 256.
__global__ void __launch_bounds__(256) trmv_lower_kernel(double* __restrict__ z, const double* __restrict__ X,
                                                         const double* __restrict__ v, const int* __restrict__ ids,
                                                         const int* __restrict__ xids, int Np) {
    const int id = ids[blockIdx.y];
    const int xid = xids ? xids[blockIdx.y] : id;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a = blockIdx.x * 8 + warp;
    const double* row = X + (size_t)xid * Np * Np + (size_t)a * Np;
    const double* vv = v + (size_t)id * Np;
    double s = 0.0;
    for (int n = lane; n <= a; n += 32) s = fma(row[n], vv[n], s);
    s = warp_sum(s);
    if (lane == 0) z[(size_t)id * Np + a] = s;
}

// u[n] = sum_{a>=n} X[a][n] z[a],  g[n] = sum_{a>=n} X[a][n]^2   (z / u may be null).
// One CTA owns a 64-column tile and walks all rows below it (fixed summation order: deterministic).
// grid = (nt column tiles, nmat), block = 256 (64 columns x 4 row groups).
__global__ void __launch_bounds__(256) trmv_upper_norm_kernel(double* __restrict__ u, double* __restrict__ g,
                                                              const double* __restrict__ X,
                                                              const double* __restrict__ z,
                                                              const int* __restrict__ ids, int Np) {
    __shared__ double su[4][NB], sg[4][NB];
    const int id = ids[blockIdx.y];
    const int ct = blockIdx.x, n = ct * NB + (threadIdx.x & 63), rg = threadIdx.x >> 6;
    const double* Xm = X + (size_t)id * Np * Np;
    const double* zz = z ? z + (size_t)id * Np : nullptr;
    double pu = 0.0, pg = 0.0;
    for (int a = ct * NB + rg; a < Np; a += 4) {
        double x = Xm[(size_t)a * Np + n];
        pg = fma(x, x, pg);
        if (zz) pu = fma(x, zz[a], pu);
    }
    su[rg][threadIdx.x & 63] = pu;
    sg[rg][threadIdx.x & 63] = pg;
    __syncthreads();
    if (rg == 0) {
        int cidx = threadIdx.x;
        pu = (su[0][cidx] + su[1][cidx]) + (su[2][cidx] + su[3][cidx]);
        pg = (sg[0][cidx] + sg[1][cidx]) + (sg[2][cidx] + sg[3][cidx]);
        if (zz) u[(size_t)id * Np + n] = pu;
        g[(size_t)id * Np + n] = pg;
    }
}

}  // namespace gprn

// factor.cuh -- batched blocked Cholesky and triangular inverse (FP64 compute bound), batched over an arbitrary
// list of matrices (blockIdx.y).
//
// Cholesky, 64x64 tiles, panels of 4 tile columns (left-looking inside a panel, right-looking between panels):
//     panel_col              : one tile column in ONE launch -- in-panel update, potrf64 of the diagonal tile
//                              (redundantly per CTA), forward substitution of two row tiles       [latency-bound regime]
//     potrf_col + trsm_col   : the same step as two launches, diagonal tile factored once          [SM-bound regime]
//     syrk_update            : trailing update with 64x64 DMMA tiles (sizes that are not 256-aligned)
//     syrk_outer             : trailing update on the GEMM core of gemm128.cuh (two-level path, Np % 256 == 0)
// Inverse factor X = L^-1 by block rows:
//     trtri_diag             : X_ii = L_ii^-1
//     trtri_row              : X_ij = -L_ii^-1 * sum_{k=j}^{i-1} L_ik X_kj   (64-tile path)
//     trtri_outer + trtri_inblock : the same sum split into a GEMM-core part (rows above the 256-row block) and an
//                              in-block sweep with substitution (two-level path)
// plus the triangular matrix-vector passes (trmv_lower, trmv_upper_norm).
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
    GPRN_TRACE_SCOPE(TK_FORM_A);
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

// Left-looking panel step for tile column k of the panel that starts at tile column k0 (panels are 4 tile
// columns wide): every CTA (a) forms A_kk - sum_{k'} L_kk' L_kk'^T over the panel's finished columns and
// factors it (redundantly -- it is 64^3/3 flops and saves a launch plus a grid-wide dependency), (b) does the
// same update for its own two row tiles i0, i0+1 and solves them against L_kk.  L_kk overwrites A_kk, which
// every CTA of the matrix reads, so it is stored by whichever CTA is the LAST to have consumed A_kk (ticket in
// ctr[id], self-resetting; no spinning); that CTA also adds sum(log pivots) = 2 sum(log diag L) to logdet[id]
// and raises status[id] on a non-positive pivot.
// grid = (max(1, ceil((nt-k-1)/2)), nmat), block = 256 (warps 0-3: diagonal + row tile i0, warps 4-7: i0+1).
#define PANEL_SMEM ((3 * NB * LDT + 4 * NB) * sizeof(double))
#define TRSM_LDV LDV2
__global__ void __launch_bounds__(256) panel_col_kernel(double* __restrict__ W, const int* __restrict__ ids, int Np,
                                                        int k, int k0, double* __restrict__ logdet,
                                                        int* __restrict__ status, int* __restrict__ ctr) {
    GPRN_TRACE_SCOPE(TK_PANEL);
    extern __shared__ double smem[];
    double* Bs = smem;                 // L_kk' operand; then potrf input (stride LDV) and output L_kk (stride LDT)
    double* As0 = smem + NB * LDT;     // L_i0k' / L_i1k' operands; then V (substitution vectors, stride 129)
    double* As1 = smem + 2 * NB * LDT;
    double* V = As0;
    double* col = smem + 3 * NB * LDT;
    double* pivs = col + 2 * NB;
    double* rd = pivs + NB;
    __shared__ int bad, last;
    const int nt = Np / NB;
    const int id = ids[blockIdx.y];
    double* Wm = W + (size_t)id * Np * Np;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp >> 2, w4 = warp & 3, wm = w4 >> 1, wn = w4 & 1, tid4 = tid & 127;
    const int r = lane >> 2, c = lane & 3;
    const int i_own = k + 1 + 2 * blockIdx.x + grp;
    const bool has = i_own < nt;
    if (tid == 0) bad = 0;
    double accd[4][4][2], accr[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int m = wm * 32 + a * 8 + r, n = wn * 32 + b * 8 + 2 * c;
            accd[a][b][0] = accd[a][b][1] = accr[a][b][0] = accr[a][b][1] = 0.0;
            if (grp == 0) {
                double2 v = *reinterpret_cast<const double2*>(Wm + (size_t)(k * NB + m) * Np + k * NB + n);
                accd[a][b][0] = v.x; accd[a][b][1] = v.y;
            }
            if (has) {
                double2 v = *reinterpret_cast<const double2*>(Wm + (size_t)(i_own * NB + m) * Np + k * NB + n);
                accr[a][b][0] = v.x; accr[a][b][1] = v.y;
            }
        }
    for (int kp = k0; kp < k; kp++) {
        load_tile<false, false>(Bs, Wm + (size_t)(k * NB) * Np + kp * NB, Np, tid, 256);
        if (has) load_tile<false, false>(grp ? As1 : As0, Wm + (size_t)(i_own * NB) * Np + kp * NB, Np, tid4, 128);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        if (grp == 0) mma_tile<true>(accd, Bs, Bs, wm, wn, lane);
        if (has) mma_tile<true>(accr, grp ? As1 : As0, Bs, wm, wn, lane);
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int m = wm * 32 + a * 8 + r, n = wn * 32 + b * 8 + 2 * c;
            if (grp == 0) {
                Bs[m * LDV + n] = accd[a][b][0];
                Bs[m * LDV + n + 1] = accd[a][b][1];
            }
            if (has) {
                V[n * TRSM_LDV + grp * NB + m] = accr[a][b][0];
                V[(n + 1) * TRSM_LDV + grp * NB + m] = accr[a][b][1];
            }
        }
    __syncthreads();
    if (tid == 0) {      // every value of A_kk this CTA needs now sits in shared memory
        const int prev = atomicAdd(&ctr[id], 1);
        last = (prev == (int)gridDim.x - 1);
        if (last) ctr[id] = 0;
    }
    potrf64(Bs, LDV, Bs, rd, col, pivs, &bad);
    if (k + 1 + 2 * (int)blockIdx.x + grp < nt) subst_lower_mma<2>(Bs, LDT, rd, V, TRSM_LDV, grp * NB + w4 * 16);
    __syncthreads();
    if (has) {
        double* dst = Wm + (size_t)(i_own * NB) * Np + k * NB;
        for (int e = tid4; e < NB * NB; e += 128) {
            int m = e >> 6, n = e & 63;
            dst[(size_t)m * Np + n] = V[n * TRSM_LDV + grp * NB + m];
        }
    }
    if (last) {
        double* dst = Wm + (size_t)(k * NB) * Np + k * NB;
        for (int e = tid; e < NB * NB; e += 256) {
            int m = e >> 6, n = e & 63;
            dst[(size_t)m * Np + n] = Bs[m * LDT + n];
        }
        if (tid < 32) {
            double sl = log(pivs[tid]) + log(pivs[tid + 32]);
            sl = warp_sum(sl);
            if (tid == 0) {
                atomicAdd(&logdet[id], sl);
                if (bad) status[id] = 1;
            }
        }
    }
}

// The same panel step as two launches (default): potrf_col_kernel factors the diagonal tile once per matrix,
// trsm_col_kernel then solves FOUR row tiles per CTA (all 256 threads carry a substitution vector).  Compared with
// panel_col_kernel this drops the redundant per-CTA potrf64 and halves the CTA count, so a panel step holds
// roughly a third of the SM time -- SMs that the GEMM launches of the other stream groups can use.  Same arithmetic, same
// summation order as panel_col_kernel: the factors are bit-identical.
// potrf_col_kernel: grid = (nmat), block = 256, dynamic smem POTRF_COL_SMEM.
#define POTRF_COL_SMEM ((NB * LDT + 4 * NB) * sizeof(double))
__global__ void __launch_bounds__(256) potrf_col_kernel(double* __restrict__ W, const int* __restrict__ ids, int Np,
                                                        int k, int k0, double* __restrict__ logdet,
                                                        int* __restrict__ status) {
    GPRN_TRACE_SCOPE(TK_POTRF);
    extern __shared__ double smem[];
    double* Bs = smem;                 // L_kk' operand; then potrf input (stride LDV) and output L_kk (stride LDT)
    double* col = smem + NB * LDT;
    double* pivs = col + 2 * NB;
    double* rd = pivs + NB;
    __shared__ int bad;
    const int id = ids[blockIdx.x];
    double* Wm = W + (size_t)id * Np * Np;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp >> 2, w4 = warp & 3, wm = w4 >> 1, wn = w4 & 1;
    const int r = lane >> 2, c = lane & 3;
    if (tid == 0) bad = 0;
    double accd[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int m = wm * 32 + a * 8 + r, n = wn * 32 + b * 8 + 2 * c;
            accd[a][b][0] = accd[a][b][1] = 0.0;
            if (grp == 0) {
                double2 v = *reinterpret_cast<const double2*>(Wm + (size_t)(k * NB + m) * Np + k * NB + n);
                accd[a][b][0] = v.x; accd[a][b][1] = v.y;
            }
        }
    for (int kp = k0; kp < k; kp++) {
        load_tile<false>(Bs, Wm + (size_t)(k * NB) * Np + kp * NB, Np, tid, 256);
        __syncthreads();
        if (grp == 0) mma_tile<true>(accd, Bs, Bs, wm, wn, lane);
        __syncthreads();
    }
    if (grp == 0) {
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int m = wm * 32 + a * 8 + r, n = wn * 32 + b * 8 + 2 * c;
                Bs[m * LDV + n] = accd[a][b][0];
                Bs[m * LDV + n + 1] = accd[a][b][1];
            }
    }
    __syncthreads();
    potrf64(Bs, LDV, Bs, rd, col, pivs, &bad);
    double* dst = Wm + (size_t)(k * NB) * Np + k * NB;
    for (int e = tid; e < NB * NB; e += 256) {
        int m = e >> 6, n = e & 63;
        dst[(size_t)m * Np + n] = Bs[m * LDT + n];
    }
    if (tid < 32) {
        double sl = log(pivs[tid]) + log(pivs[tid + 32]);
        sl = warp_sum(sl);
        if (tid == 0) {
            logdet[id] += sl;          // one CTA per matrix and launch, launches of a matrix are stream ordered
            if (bad) status[id] = 1;
        }
    }
}

// trsm_col_kernel<GROUPS>: L_ik = (A_ik - sum_{k'} L_ik' L_kk'^T) L_kk^-T for the 2*GROUPS row tiles
// i = k+1 + 2*GROUPS*blockIdx.x ...  Each group of 4 warps owns two tiles (one accumulator set each).
// GROUPS = 1 (128 threads, 105 KB, ~190 registers) is the default: such a CTA fits on an SM NEXT TO one CTA of the
// GEMM core (gemm128.cuh: 128 threads, 92 KB), so the substitution latency of one stream group hides behind the
// DMMA work of another instead of idling an SM; GROUPS = 2 (256 threads) does not.
// grid = (ceil((nt-k-1)/(2*GROUPS)), nmat), block = 128*GROUPS, dynamic smem TRSM_COL_SMEM(GROUPS).
#define TRSM_COL_SMEM(G) (((1 + 2 * (G)) * NB * LDT + NB) * sizeof(double))
template <int GROUPS>
__global__ void __launch_bounds__(128 * GROUPS) trsm_col_kernel(double* __restrict__ W, const int* __restrict__ ids,
                                                                int Np, int k, int k0) {
    GPRN_TRACE_SCOPE(TK_TRSM);
    constexpr int NTHR = 128 * GROUPS, VLD = (GROUPS == 2) ? LDV4 : LDV2;
    extern __shared__ double smem[];
    double* Bs = smem;                 // L_kk' operand, then L_kk
    double* As = smem + NB * LDT;      // 2*GROUPS row-tile operands; then V (substitution vectors, stride VLD)
    double* V = As;
    double* rd = smem + (1 + 2 * GROUPS) * NB * LDT;
    const int nt = Np / NB;
    const int id = ids[blockIdx.y];
    double* Wm = W + (size_t)id * Np * Np;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp >> 2, w4 = warp & 3, wm = w4 >> 1, wn = w4 & 1, tid4 = tid & 127;
    const int r = lane >> 2, c = lane & 3;
    const int ibase = k + 1 + 2 * GROUPS * blockIdx.x;
    const int ia = ibase + 2 * grp, ib = ia + 1;
    const bool hasa = ia < nt, hasb = ib < nt;
    double acca[4][4][2], accb[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int m = wm * 32 + a * 8 + r, n = wn * 32 + b * 8 + 2 * c;
            acca[a][b][0] = acca[a][b][1] = accb[a][b][0] = accb[a][b][1] = 0.0;
            if (hasa) {
                double2 v = *reinterpret_cast<const double2*>(Wm + (size_t)(ia * NB + m) * Np + k * NB + n);
                acca[a][b][0] = v.x; acca[a][b][1] = v.y;
            }
            if (hasb) {
                double2 v = *reinterpret_cast<const double2*>(Wm + (size_t)(ib * NB + m) * Np + k * NB + n);
                accb[a][b][0] = v.x; accb[a][b][1] = v.y;
            }
        }
    double* Aa = As + (2 * grp) * NB * LDT;
    double* Ab = Aa + NB * LDT;
    for (int kp = k0; kp < k; kp++) {
        load_tile<false, false>(Bs, Wm + (size_t)(k * NB) * Np + kp * NB, Np, tid, NTHR);
        if (hasa) load_tile<false, false>(Aa, Wm + (size_t)(ia * NB) * Np + kp * NB, Np, tid4, 128);
        if (hasb) load_tile<false, false>(Ab, Wm + (size_t)(ib * NB) * Np + kp * NB, Np, tid4, 128);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        if (hasa) mma_tile<true>(acca, Aa, Bs, wm, wn, lane);
        if (hasb) mma_tile<true>(accb, Ab, Bs, wm, wn, lane);
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int m = wm * 32 + a * 8 + r, n = wn * 32 + b * 8 + 2 * c;
            if (hasa) {
                V[n * VLD + (2 * grp) * NB + m] = acca[a][b][0];
                V[(n + 1) * VLD + (2 * grp) * NB + m] = acca[a][b][1];
            }
            if (hasb) {
                V[n * VLD + (2 * grp + 1) * NB + m] = accb[a][b][0];
                V[(n + 1) * VLD + (2 * grp + 1) * NB + m] = accb[a][b][1];
            }
        }
    load_tile<false>(Bs, Wm + (size_t)(k * NB) * Np + k * NB, Np, tid, NTHR);
    __syncthreads();
    if (tid < NB) rd[tid] = 1.0 / Bs[tid * LDT + tid];
    __syncthreads();
    if (ibase + (warp >> 1) < nt) subst_lower_mma<4>(Bs, LDT, rd, V, VLD, (warp >> 1) * NB + (warp & 1) * 32);
    __syncthreads();
    if (hasa) {
        double* dst = Wm + (size_t)(ia * NB) * Np + k * NB;
        for (int e = tid4; e < NB * NB; e += 128) {
            int m = e >> 6, n = e & 63;
            dst[(size_t)m * Np + n] = V[n * VLD + (2 * grp) * NB + m];
        }
    }
    if (hasb) {
        double* dst = Wm + (size_t)(ib * NB) * Np + k * NB;
        for (int e = tid4; e < NB * NB; e += 128) {
            int m = e >> 6, n = e & 63;
            dst[(size_t)m * Np + n] = V[n * VLD + (2 * grp + 1) * NB + m];
        }
    }
}

// Trailing update after tile columns [kb, ke): A_ij -= sum_{k=kb}^{ke-1} L_ik L_jk^T for ke <= j <= i (64x64 tiles).
// grid = (n(n+1)/2 with n = nt-ke, nmat), block = 128 (2x2 warps, 32x32 each).  Dynamic shared memory 2*TILE_SMEM.
__global__ void __launch_bounds__(128) syrk_update_kernel(double* __restrict__ W, const int* __restrict__ ids, int Np,
                                                          int kb, int ke) {
    GPRN_TRACE_SCOPE(TK_SYRK64);
    extern __shared__ double smem[];
    double* As = smem;
    double* Bs = smem + NB * LDT;
    int ti, tj;
    tri_decode(blockIdx.x, ti, tj);
    const int I = ke + ti, J = ke + tj;
    const int id = ids[blockIdx.y];
    double* Wm = W + (size_t)id * Np * Np;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
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
    for (int k = kb; k < ke; k++) {
        load_tile<false>(As, Wm + (size_t)(I * NB) * Np + k * NB, Np, tid, 128);
        if (I != J) load_tile<false>(Bs, Wm + (size_t)(J * NB) * Np + k * NB, Np, tid, 128);
        __syncthreads();
        mma_tile<true>(acc, As, (I != J) ? Bs : As, wm, wn, lane);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<double2*>(C + (size_t)(wm * 32 + i * 8 + r) * Np + wn * 32 + j * 8 + 2 * c) = v;
        }
}

// X_ii = L_ii^-1 for every diagonal tile.  grid = (nt, nmat), block = 64 (one column per thread).
#define TRTRI_DIAG_SMEM ((NB * LDT + NB * LDV + NB) * sizeof(double))
__global__ void __launch_bounds__(64) trtri_diag_kernel(double* __restrict__ X, const double* __restrict__ W,
                                                        const int* __restrict__ ids, int Np) {
    GPRN_TRACE_SCOPE(TK_TRTRI_DIAG);
    extern __shared__ double smem[];
    double* Ls = smem;
    double* V = smem + NB * LDT;
    const int id = ids[blockIdx.y], i = blockIdx.x, tid = threadIdx.x;
    const size_t off = (size_t)id * Np * Np + (size_t)(i * NB) * Np + i * NB;
    double* rd = V + NB * LDV;
    load_tile<false>(Ls, W + off, Np, tid, 64);
    for (int m = 0; m < NB; m++) V[m * LDV + tid] = (m == tid) ? 1.0 : 0.0;
    __syncthreads();
    rd[tid] = 1.0 / Ls[tid * LDT + tid];
    __syncthreads();
    subst_lower_mma<4>(Ls, LDT, rd, V, LDV, (tid >> 5) * 32, (tid >> 5) * 4);
    __syncthreads();
    double* Xt = X + off;
    for (int m = 0; m < NB; m++) Xt[(size_t)m * Np + tid] = V[m * LDV + tid];
}

// Block row i of the inverse: X_ij for j = blockIdx.x < i.  grid = (i, nmat), block = 128.
// Dynamic shared memory: As, Bs (MMA operands) + V (64*LDV).
#define TRTRI_SMEM (2 * TILE_SMEM + (NB * LDV + NB) * sizeof(double))
__global__ void __launch_bounds__(128) trtri_row_kernel(double* __restrict__ X, const double* __restrict__ W,
                                                        const int* __restrict__ ids, int Np, int i) {
    GPRN_TRACE_SCOPE(TK_TRTRI_ROW);
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
    double* rd = V + NB * LDV;
    if (tid < NB) rd[tid] = 1.0 / As[tid * LDT + tid];
    __syncthreads();
    subst_lower_mma<2>(As, LDT, rd, V, LDV, warp * 16);
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
// with 128x128 tiles (two 64x128 CTAs each) and K = 256 / 512 deep DMMA products (gemm128.cuh).
// ------------------------------------------------------------------------------------------------
#define OUTER_KB 256

// Trailing update with the K-slab of columns [c0, c0 + KB): for 128x128 tiles (TI >= TJ) of the trailing square that
// starts at column t0,  C -= L[rows, c0:c0+KB] L[cols, c0:c0+KB]^T.  jcols = 0: all tiles; jcols = 2: only the first
// two tile columns (the next 256-column panel).  The host applies slabs lazily -- after an even panel only the next
// panel's columns are updated (KB = 256), after an odd panel everything to the right gets both slabs at once
// (KB = 512) -- which halves the read-modify-write passes over the trailing matrix.
// A 128x128 tile is formed by G_SPLIT = 128 / G_BM CTAs (row slabs of G_BM rows; blockIdx.x = tile * G_SPLIT + slab).
// grid = (G_SPLIT * tiles, nmat), block = G_THREADS, dynamic smem GEMM128_SMEM.
#define G_SPLIT (128 / G_BM)
__global__ void __launch_bounds__(G_THREADS, G_MINB) syrk_outer_kernel(double* __restrict__ W, const int* __restrict__ ids,
                                                               int Np, int c0, int KB, int t0, int jcols) {
    GPRN_TRACE_SCOPE(TK_SYRK_OUTER);
    extern __shared__ double smem[];
    int TI, TJ;
    const int tile = blockIdx.x / G_SPLIT, slab = blockIdx.x % G_SPLIT;
    if (jcols == 0) {
        tri_decode(tile, TI, TJ);
    } else {
        const int n128 = (Np - t0) / 128;
        if (tile < n128) { TI = tile; TJ = 0; }
        else { TI = tile - n128 + 1; TJ = 1; }
    }
    const int r0 = t0 + TI * 128 + slab * G_BM, n0 = t0 + TJ * G_BN;
    const int id = ids[blockIdx.y];
    double* Wm = W + (size_t)id * Np * Np;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp / G_WARPS_N, wn = warp % G_WARPS_N;
    const int r = lane >> 2, c = lane & 3;
    double* C = Wm + (size_t)r0 * Np + n0;
    // The products are summed on their own and subtracted from C once at the end: starting the accumulators at
    // -C (which would make the epilogue a pure store) rounds every partial sum at the magnitude of C and costs
    // measurable ELBO parity at N >= 2048 (c5 anchor: 1e-10 bar missed), for no gain in time.
    double acc[G_MI][G_NI][2];
#pragma unroll
    for (int i = 0; i < G_MI; i++)
#pragma unroll
        for (int j = 0; j < G_NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    gemm128_mainloop<false>(acc, smem, Wm + (size_t)r0 * Np + c0, Np, Wm + (size_t)n0 * Np + c0, Np, KB);
#pragma unroll
    for (int i = 0; i < G_MI; i++)
#pragma unroll
        for (int j = 0; j < G_NI; j++) {
            double2* p = reinterpret_cast<double2*>(C + (size_t)(wm * G_WM + i * 8 + r) * Np + wn * G_WN + j * 8 + 2 * c);
            double2 v = *p;
            v.x -= acc[i][j][0];
            v.y -= acc[i][j][1];
            *p = v;
        }
}

// Inverse, block row of 256 rows starting at R0: X[rows, 0:R0] = - L[rows, n0:R0] X[n0:R0, cols] (the part of
// the row-sweep sum that lies above the block), G_BM x 128 tiles.  X must hold zeros in its upper tiles.
// The K range of a tile, R0 - n0, shrinks with its column, which leaves most SMs idle behind the few long tiles
// when few matrices are in flight.  The host can therefore cut the range into chunks of kc (>= TRTRI_KC): chunk 0
// writes to X as before, chunk s >= 1 to the partial buffer Gp[id][s-1] (256 x Np strip, same row/column indexing);
// the in-block kernel adds the partials in a fixed order (deterministic, no atomics).  With enough matrices in
// flight the launch is already full and kc = R0 (no split) is faster: shorter K means more prologue / epilogue.
// grid = ((256 / G_BM) * trtri_outer_units(R0), nmat), block = G_THREADS, dynamic smem GEMM128_SMEM.
#define TRTRI_KC 1024
#define TRTRI_MAXCH(Np) (((Np) + TRTRI_KC - 1) / TRTRI_KC)      // chunks of the longest tile
__host__ __device__ __forceinline__ int trtri_nchunk(int R0, int n0, int kc) { return (R0 - n0 + kc - 1) / kc; }
__host__ __forceinline__ int trtri_outer_units(int R0, int kc) {
    int u = 0;
    for (int n0 = 0; n0 < R0; n0 += G_BN) u += trtri_nchunk(R0, n0, kc);
    return u;
}
__global__ void __launch_bounds__(G_THREADS, G_MINB) trtri_outer_kernel(double* __restrict__ X, const double* __restrict__ W,
                                                                double* __restrict__ Gp, const int* __restrict__ ids,
                                                                int Np, int R0, int units, int kc) {
    GPRN_TRACE_SCOPE(TK_TRTRI_OUTER);
    extern __shared__ double smem[];
    const int half = blockIdx.x / units;          // row slab (G_BM rows) of the 256-row block
    int u = blockIdx.x % units, n0 = 0, s = 0;
    for (;; n0 += G_BN) {                      // decode (column tile, chunk)
        const int nc = trtri_nchunk(R0, n0, kc);
        if (u < nc) { s = u; break; }
        u -= nc;
    }
    const int r0 = R0 + half * G_BM;
    const int k0 = n0 + s * kc, k1 = min(k0 + kc, R0);
    const int id = ids[blockIdx.y];
    const double* Wm = W + (size_t)id * Np * Np;
    double* Xm = X + (size_t)id * Np * Np;
    double acc[G_MI][G_NI][2];
#pragma unroll
    for (int i = 0; i < G_MI; i++)
#pragma unroll
        for (int j = 0; j < G_NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    gemm128_mainloop<true>(acc, smem, Wm + (size_t)r0 * Np + k0, Np, Xm + (size_t)k0 * Np + n0, Np, k1 - k0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp / G_WARPS_N, wn = warp % G_WARPS_N;
    const int r = lane >> 2, c = lane & 3;
    double* C = (s == 0) ? Xm + (size_t)r0 * Np + n0
                         : Gp + ((size_t)id * (TRTRI_MAXCH(Np) - 1) + (s - 1)) * OUTER_KB * Np + (size_t)(half * G_BM) * Np + n0;
#pragma unroll
    for (int i = 0; i < G_MI; i++)
#pragma unroll
        for (int j = 0; j < G_NI; j++) {
            double2 v = make_double2(-acc[i][j][0], -acc[i][j][1]);
            *reinterpret_cast<double2*>(C + (size_t)(wm * G_WM + i * 8 + r) * Np + wn * G_WN + j * 8 + 2 * c) = v;
        }
}

// Inverse, in-block sweep for the 4 tile rows i = i0 .. i0+3 of a 256-row block: one CTA per 64-column tile j
// walks the rows in order,  X_ij = L_ii^-1 ( G'_ij - sum_{k=max(j,i0)}^{i-1} L_ik X_kj ),  with G' the partial
// sum left in place by trtri_outer_kernel (zero for columns inside the block).  grid = (min(i0+4, nt) - 1, nmat),
// block = 128, dynamic smem TRTRI_SMEM.
__global__ void __launch_bounds__(128) trtri_inblock_kernel(double* __restrict__ X, const double* __restrict__ W,
                                                            const double* __restrict__ Gp, const int* __restrict__ ids,
                                                            int Np, int i0, int kc) {
    GPRN_TRACE_SCOPE(TK_TRTRI_INBLOCK);
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
    const int iend = min(i0 + 4, Np / NB);
    for (int i = max(i0, j + 1); i < iend; i++) {
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
            // split-K partials of trtri_outer_kernel, added in chunk order
            const int nch = Gp ? trtri_nchunk(i0 * NB, (j >> 1) * G_BN, kc) : 1;
            for (int sp = 1; sp < nch; sp++) {
                const double* gp = Gp + ((size_t)id * (TRTRI_MAXCH(Np) - 1) + (sp - 1)) * OUTER_KB * Np +
                                   (size_t)((i - i0) * NB) * Np + j * NB;
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        double2 v = *reinterpret_cast<const double2*>(gp + (size_t)(wm * 32 + a * 8 + r) * Np + wn * 32 + b * 8 + 2 * c);
                        acc[a][b][0] += v.x;
                        acc[a][b][1] += v.y;
                    }
            }
        } else {
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
        }
        for (int k = kstart; k < i; k++) {
            load_tile<false, false>(As, Wm + (size_t)(i * NB) * Np + k * NB, Np, tid, 128);
            load_tile<true, false>(Bs, Xm + (size_t)(k * NB) * Np + j * NB, Np, tid, 128);
            cp_async_commit();
            cp_async_wait<0>();
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
        double* rd = V + NB * LDV;
        if (tid < NB) rd[tid] = 1.0 / As[tid * LDT + tid];
        __syncthreads();
        subst_lower_mma<2>(As, LDT, rd, V, LDV, warp * 16);
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
    GPRN_TRACE_SCOPE(TK_TRMV_LOWER);
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
    GPRN_TRACE_SCOPE(TK_TRMV_UPPER);
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

// Bitwise comparison of the lower triangle of matrix blockIdx.x of W (and of its log-det) with a reference factor:
// adds 1 to *mismatches when anything differs (gprn_debug_panel_stress).  grid = (nmat), block = 256.
__global__ void lower_mismatch_kernel(const double* __restrict__ W, const double* __restrict__ ref,
                                      const double* __restrict__ logdet, const double* __restrict__ logdet_ref,
                                      int Np, unsigned long long* __restrict__ mismatches) {
    __shared__ int diff;
    if (threadIdx.x == 0) diff = 0;
    __syncthreads();
    const double* Wm = W + (size_t)blockIdx.x * Np * Np;
    int d = 0;
    for (size_t e = threadIdx.x; e < (size_t)Np * Np; e += blockDim.x) {
        const int r = (int)(e / Np), c = (int)(e % Np);
        if (c <= r && __double_as_longlong(Wm[e]) != __double_as_longlong(ref[e])) d = 1;
    }
    if (threadIdx.x == 0 && __double_as_longlong(logdet[blockIdx.x]) != __double_as_longlong(logdet_ref[0])) d = 1;
    if (d) diff = 1;
    __syncthreads();
    if (threadIdx.x == 0 && diff) atomicAdd(mismatches, 1ULL);
}

}  // namespace gprn

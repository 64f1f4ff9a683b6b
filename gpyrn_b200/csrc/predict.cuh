// predict.cuh -- GP conditional mean / variance at T test epochs (gpyrn/_gp.py:107-138) and the GPRN
// combination step (gpyrn/meanfield.py:1364-1372).
//
// The reference loops over the T test points, doing a 1-RHS cho_solve and a T x N dgemv per point and
// keeping only the diagonal of a T x T matrix.  Here, with X = chol(K + diag(v))^-1:
//     mean[t] = Kstar[t,:] . alpha,  alpha = X^T X m
//     var[t]  = k(0) + 1.25e-12 - || X Kstar[t,:]^T ||^2
// Kstar is assembled in chunks of test points; the squared norms come from a DMMA product
// Kstar_chunk * X^T whose 64x64 output tiles are squared and row-reduced in registers, never stored.
#pragma once
#include "common.cuh"
#include "assemble.cuh"
#include "gemm128.cuh"

namespace gprn {

// All kernels below are batched over the GPs of one hyper-parameter set through the last grid dimension b:
// Kstar chunk b at Ks + b*kstride, inverse factor b at X + (xid0 + b)*Np*Np, vectors / outputs with their own strides.

// out[t] = sum_n Ks[t][n] * alpha[n].  One warp per row.  grid = (ceil(T/8), nb), block = 256.
__global__ void __launch_bounds__(256) rect_gemv_kernel(double* __restrict__ out, size_t ostride,
                                                        const double* __restrict__ Ks, size_t kstride, size_t ld,
                                                        const double* __restrict__ alpha, size_t astride, int T, int N) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + warp;
    if (t >= T) return;
    out += (size_t)blockIdx.y * ostride;
    Ks += (size_t)blockIdx.y * kstride;
    alpha += (size_t)blockIdx.y * astride;
    const double* row = Ks + (size_t)t * ld;
    double s = 0.0;
    for (int n = lane; n < N; n += 32) s = fma(row[n], alpha[n], s);
    s = warp_sum(s);
    if (lane == 0) out[t] = s;
}

// rownorm2[t] = sum_a ( sum_{n<=a} Ks[t][n] X[a][n] )^2 for the 64 test rows of this CTA.
// Ks: [Tpad][Np] (columns >= N zero, rows >= T zero), X: [Np][Np] lower.  grid = (Tpad/64, nb), block = 128,
// dynamic shared memory 2*TILE_SMEM.
__global__ void __launch_bounds__(128) predict_norm_kernel(double* __restrict__ rownorm2, size_t rstride,
                                                           const double* __restrict__ Ks, size_t kstride,
                                                           const double* __restrict__ X, int Np, int N) {
    rownorm2 += (size_t)blockIdx.y * rstride;
    Ks += (size_t)blockIdx.y * kstride;
    X += (size_t)blockIdx.y * Np * Np;
    extern __shared__ double smem[];
    double* As = smem;
    double* Bs = smem + NB * LDT;
    __shared__ double rs[2][NB];
    const int nt = Np / NB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    const int r = lane >> 2, cc = lane & 3;
    const double* Kt = Ks + (size_t)(blockIdx.x * NB) * Np;
    double rowacc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int a = 0; a < nt; a++) {
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kt = 0; kt <= a; kt++) {
            load_tile<false>(As, Kt + kt * NB, Np, tid, 128);
            load_tile<false>(Bs, X + (size_t)(a * NB) * Np + kt * NB, Np, tid, 128);
            __syncthreads();
            mma_tile<false>(acc, As, Bs, wm, wn, lane);
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int ga = a * NB + wn * 32 + j * 8 + 2 * cc;      // identity padding of X must not count
                if (ga < N) rowacc[i] = fma(acc[i][j][0], acc[i][j][0], rowacc[i]);
                if (ga + 1 < N) rowacc[i] = fma(acc[i][j][1], acc[i][j][1], rowacc[i]);
            }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double s = rowacc[i];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (cc == 0) rs[wn][wm * 32 + i * 8 + r] = s;
    }
    __syncthreads();
    if (tid < NB) rownorm2[blockIdx.x * NB + tid] = rs[0][tid] + rs[1][tid];
}

// Large-N variant of predict_norm_kernel on the DMMA GEMM core (gemm128.cuh): CTA (tt, at) forms the tile
// C[t][a] = sum_{n <= a} Ks[t][n] X[a][n] for G_BM test epochs x 128 rows of X (K = a0 + 128, X lower triangular),
// squares it and writes the 128 row sums to partial[at][t]; predict_var_kernel adds the partials in tile order.
// Ks: [Tpad][Np] with zero padding (columns >= N, rows >= T).  grid = (Tpad/G_BM, Np/128, nb), block = G_THREADS,
// dynamic smem GEMM128_SMEM.  blockIdx.y is mapped to descending a so that the long-K tiles start first.
__global__ void __launch_bounds__(G_THREADS, G_MINB) predict_norm128_kernel(double* __restrict__ partial, size_t pstride,
                                                                    int Tpad, const double* __restrict__ Ks,
                                                                    size_t kstride, const double* __restrict__ X,
                                                                    int Np) {
    extern __shared__ double smem[];
    partial += (size_t)blockIdx.z * pstride;
    Ks += (size_t)blockIdx.z * kstride;
    X += (size_t)blockIdx.z * Np * Np;
    __shared__ double rs[G_WARPS_N][G_BM];
    const int na = Np / G_BN;
    const int at = na - 1 - blockIdx.y, a0 = at * G_BN, t0 = blockIdx.x * G_BM;
    double acc[G_MI][G_NI][2];
#pragma unroll
    for (int i = 0; i < G_MI; i++)
#pragma unroll
        for (int j = 0; j < G_NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    gemm128_mainloop<false>(acc, smem, Ks + (size_t)t0 * Np, Np, X + (size_t)a0 * Np, Np, a0 + G_BN);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp / G_WARPS_N, wn = warp % G_WARPS_N;
    const int r = lane >> 2, c = lane & 3;
#pragma unroll
    for (int i = 0; i < G_MI; i++) {
        double sr = 0.0;
#pragma unroll
        for (int j = 0; j < G_NI; j++) {
            sr = fma(acc[i][j][0], acc[i][j][0], sr);
            sr = fma(acc[i][j][1], acc[i][j][1], sr);
        }
        sr += __shfl_xor_sync(0xffffffffu, sr, 1);
        sr += __shfl_xor_sync(0xffffffffu, sr, 2);
        if (c == 0) rs[wn][wm * G_WM + i * 8 + r] = sr;
    }
    __syncthreads();
    if (threadIdx.x < G_BM) {
        const int t = threadIdx.x;
        double sp = rs[0][t];
#pragma unroll
        for (int w = 1; w < G_WARPS_N; w++) sp += rs[w][t];
        partial[(size_t)at * Tpad + t0 + t] = sp;
    }
}

// var[t] = k(0) + nugget - rownorm2[t]   (diagonal of Kstarstar, _gp.py:131,136-137)
// rownorm2: [nparts][stride] partial sums (nparts = 1 for predict_norm_kernel), added in order.
// grid = (ceil(T/256), nb); GP b uses program tok[b*GPRN_MAX_PROG..], len[b], parameters hyper + paroff[b].
__global__ void predict_var_kernel(double* __restrict__ var, size_t vstride, const double* __restrict__ rownorm2,
                                   size_t rstride, int nparts, int stride, int T, const int32_t* __restrict__ tok,
                                   const int32_t* __restrict__ len, const double* __restrict__ hyper,
                                   const int32_t* __restrict__ paroff, double nugget) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int b = blockIdx.y;
    var += (size_t)b * vstride;
    rownorm2 += (size_t)b * rstride;
    double k0 = eval_prog(tok + (size_t)b * GPRN_MAX_PROG, len[b], hyper + paroff[b], 0.0, true, false) + nugget;
    double s = 0.0;
    for (int a = 0; a < nparts; a++) s += rownorm2[(size_t)a * stride + t];
    var[t] = k0 - s;
}

// GPRN combination (meanfield.py:1364-1372; jitter^2 added once per node, quirk Q6).
// gp_mean/gp_var: [M][T] (m = j nodes, q + j*p + i weights).  out mean/var: [T][p].
__global__ void predict_combine_kernel(double* __restrict__ pmean, double* __restrict__ pvar,
                                       const double* __restrict__ gp_mean, const double* __restrict__ gp_var,
                                       const double* __restrict__ mean_t, const double* __restrict__ jit, int T,
                                       int p, int q) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T * p) return;
    int t = e / p, i = e % p;
    double m = 0.0, v = 0.0;
    m += mean_t[(size_t)i * T + t];
    double j2 = jit[i] * jit[i];
    for (int j = 0; j < q; j++) {
        double nP = gp_mean[(size_t)j * T + t], nV = gp_var[(size_t)j * T + t];
        size_t w = (size_t)(q + j * p + i) * T + t;
        double wP = gp_mean[w], wV = gp_var[w];
        m += nP * wP;
        v += ((wP * wP) * nV + wV * (nV + nP * nP)) + j2;
    }
    pmean[e] = m;
    pvar[e] = v;
}

}  // namespace gprn

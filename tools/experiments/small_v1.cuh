// small_v1.cuh -- FIRST version (kept for A/B runs: GPRN_SMALL_V1=1) of the fused per-matrix pipeline for small and mid-size matrices (Np <= 512, i.e. at most 8x8 tiles of 64).
//
// One persistent CTA takes a matrix through the WHOLE per-iteration chain
//     A = K + diag(D)  ->  L = chol(A)  ->  X = L^-1  ->  g = colnorm2(X),  u = X^T (X v),  logdet(A)
// and only the vectors g, u and the scalar log-det leave the chip.  K is read from HBM exactly once
// (lower tiles); L / X live in a per-CTA scratch of nt(nt+1)/2 tiles -- 10 tiles (320 KB) for Np = 256, which stays
// L2 resident (296 CTAs x 320 KB = 95 MB < 126 MB L2); 36 tiles (1.15 MB) for Np = 512, which streams through HBM --
// and X tiles overwrite the L tiles they no longer need.
// The building blocks are the same as in factor.cuh: DMMA m8n8k4 tile products (mma_tile), the
// register-resident 64x64 Cholesky (potrf64) and thread-per-vector substitution (subst_lower).
// Replaces, for q == 1 and N <= 256 (and for N <= 512 when enough matrices are in flight to give every SM its own:
// gprn_api.cu: decide_small_path), form_a + panel_col + trtri_* + trmv_* (one launch per phase instead of ~10-70 and
// none of their HBM round trips).  256 threads: warps 0-3 and 4-7 work on two tiles at a time.
#pragma once
#include "common.cuh"
#include "small.cuh"     // SmallArgs, small_tile, SMALL_MAX_NT, SMALL_SCRATCH_DOUBLES, the phase counters

namespace gprn {

#define SMALLV1_LDV 129              // odd strides: the fused kernel keeps the thread-per-vector substitution
#ifdef GPRN_POTRF_V1
#define SMALLV1_LDP 65               // (measured: the DMMA variant is 11 % slower here, 2 CTAs/SM already hide its latency)
#else
#define SMALLV1_LDP LDT              // potrf64 v2 works in place on the LDT layout
#endif
// 3 operand tiles + col(128) + pivs(64) + rd(64) + gacc(Np) + zacc(Np): 112 KB, two CTAs per SM
#define SMALLV1_SMEM ((3 * NB * LDT + 4 * NB + 2 * SMALL_MAX_NT * NB) * sizeof(double))

#ifdef GPRN_TRACE
#define SMALLV1_PH(i)                                                        \
    do {                                                                   \
        if (threadIdx.x == 0) {                                            \
            long long t_ = clock64();                                      \
            ph_acc[ph_cur] += (unsigned long long)(t_ - ph_last);          \
            ph_last = t_;                                                  \
            ph_cur = (i);                                                  \
        }                                                                  \
    } while (0)
#else
#define SMALLV1_PH(i)
#endif

__global__ void __launch_bounds__(256, 2) small_pipeline_v1_kernel(SmallArgs a) {
    GPRN_TRACE_SCOPE(TK_SMALL);
    extern __shared__ double smem[];
    double* Bs = smem;                 // B operand / potrf input+output (L_kk) / A operand in the inverse
    double* As0 = smem + NB * LDT;
    double* As1 = smem + 2 * NB * LDT;
    double* V = As0;                   // substitution vectors (stride 129), aliases As0/As1
    double* col = smem + 3 * NB * LDT;
    double* pivs = col + 2 * NB;
    double* rd = pivs + NB;
    double* gacc = rd + NB;            // [Np]
    double* zacc = gacc + SMALL_MAX_NT * NB;
    __shared__ int bad;
    const int Np = a.Np, nt = Np / NB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp >> 2, w4 = warp & 3, wm = w4 >> 1, wn = w4 & 1, tid4 = tid & 127;
    const int r = lane >> 2, c = lane & 3;
    double* sc = a.scratch + (size_t)blockIdx.x * SMALL_SCRATCH_DOUBLES;
#ifdef GPRN_TRACE
    unsigned long long ph_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_last = clock64();
    int ph_cur = 0;
#endif

    for (int mi = blockIdx.x; mi < a.nmat; mi += gridDim.x) {
        const int id = a.ids[mi];
        const double* Km = a.K + (size_t)id * Np * Np;
        const double* dv = a.dvec ? a.dvec + (size_t)id * Np : nullptr;
        double logsum = 0.0;
        if (tid == 0) bad = 0;
        __syncthreads();

        // ================= Cholesky, left-looking over tile columns =================
        // Column k is done in sub-rounds of two tiles: s = 0 -> (k, k+1), s = 1 -> (k+2, k+3), ...; warps 0-3 take
        // the first tile of a sub-round, warps 4-7 the second.  One accumulator set is live at a time.
        for (int k = 0; k < nt; k++) {
            for (int s = 0; k + 2 * s < nt; s++) {
                const int it = k + 2 * s + grp;
                const bool have = it < nt;
                SMALLV1_PH(1);
                double acc[4][4][2];
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) {
                        const int m = wm * 32 + x * 8 + r, n = wn * 32 + y * 8 + 2 * c;
                        double2 v = make_double2(0.0, 0.0);
                        if (have) {
                            v = *reinterpret_cast<const double2*>(Km + (size_t)(it * NB + m) * Np + k * NB + n);
                            if (dv && it == k) {
                                if (m == n) v.x += dv[k * NB + m];
                                if (m == n + 1) v.y += dv[k * NB + m];
                            }
                        }
                        acc[x][y][0] = v.x;
                        acc[x][y][1] = v.y;
                    }
                const bool diag = (s == 0 && grp == 0);
                SMALLV1_PH(2);
                for (int kp = 0; kp < k; kp++) {
                    load_tile<false, false>(Bs, small_tile(sc, k, kp), NB, tid, 256);
                    if (have && !diag) load_tile<false, false>(grp ? As1 : As0, small_tile(sc, it, kp), NB, tid4, 128);
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncthreads();
                    if (have) mma_tile<true>(acc, diag ? Bs : (grp ? As1 : As0), Bs, wm, wn, lane);
                    __syncthreads();
                }
                SMALLV1_PH(3);
                if (s == 0) {
#pragma unroll
                    for (int x = 0; x < 4; x++)
#pragma unroll
                        for (int y = 0; y < 4; y++) {
                            const int m = wm * 32 + x * 8 + r, n = wn * 32 + y * 8 + 2 * c;
                            if (grp == 0) {
                                Bs[m * SMALLV1_LDP + n] = acc[x][y][0];
                                Bs[m * SMALLV1_LDP + n + 1] = acc[x][y][1];
                            } else if (have) {
                                V[n * SMALLV1_LDV + NB + m] = acc[x][y][0];
                                V[(n + 1) * SMALLV1_LDV + NB + m] = acc[x][y][1];
                            }
                        }
                    __syncthreads();
                    SMALLV1_PH(4);
                    potrf64(Bs, SMALLV1_LDP, Bs, rd, col, pivs, &bad);
                    SMALLV1_PH(5);
                    if (tid >= NB && tid < 2 * NB && k + 1 < nt) subst_lower(Bs, LDT, rd, V, SMALLV1_LDV, tid);
                    if (tid < 32) logsum += log(pivs[tid]) + log(pivs[tid + 32]);
                    __syncthreads();
                    SMALLV1_PH(6);
                    double* dkk = small_tile(sc, k, k);
                    for (int e = tid; e < NB * NB; e += 256) dkk[e] = Bs[(e >> 6) * LDT + (e & 63)];
                    if (k + 1 < nt) {
                        double* d1 = small_tile(sc, k + 1, k);
                        for (int e = tid; e < NB * NB; e += 256) d1[e] = V[(e & 63) * SMALLV1_LDV + NB + (e >> 6)];
                    }
                } else {
#pragma unroll
                    for (int x = 0; x < 4; x++)
#pragma unroll
                        for (int y = 0; y < 4; y++) {
                            const int m = wm * 32 + x * 8 + r, n = wn * 32 + y * 8 + 2 * c;
                            if (have) {
                                V[n * SMALLV1_LDV + grp * NB + m] = acc[x][y][0];
                                V[(n + 1) * SMALLV1_LDV + grp * NB + m] = acc[x][y][1];
                            }
                        }
                    load_tile<false>(Bs, small_tile(sc, k, k), NB, tid, 256);     // L_kk back from the scratch
                    __syncthreads();
                    if (tid < NB) rd[tid] = 1.0 / Bs[tid * LDT + tid];
                    __syncthreads();
                    SMALLV1_PH(5);
                    if (tid < 2 * NB && k + 2 + (tid >> 6) < nt) subst_lower(Bs, LDT, rd, V, SMALLV1_LDV, tid);
                    __syncthreads();
                    SMALLV1_PH(6);
                    if (have) {
                        double* d2 = small_tile(sc, it, k);
                        for (int e = tid4; e < NB * NB; e += 128) d2[e] = V[(e & 63) * SMALLV1_LDV + grp * NB + (e >> 6)];
                    }
                }
                __syncthreads();
            }
        }
        if (tid < 32) {
            logsum = warp_sum(logsum);
            if (tid == 0) {
                a.logdet[id] = logsum;
                if (bad) a.mstatus[id] = 1;
            }
        }
        SMALLV1_PH(0);
        if (!a.do_inverse) continue;

        // ================= inverse by block rows; X tiles overwrite the L tiles =================
        const double* vglob = a.vv + (size_t)id * Np;      // right-hand side v: broadcast reads, L1 / L2 resident
        for (int e = tid; e < Np; e += 256) {
            gacc[e] = 0.0;
            zacc[e] = 0.0;
        }
        __syncthreads();
        for (int i = 0; i < nt; i++) {
            for (int j0 = 0; j0 <= i; j0 += 2) {
                const int j = j0 + grp;                 // this group's right-hand-side tile (j == i: identity)
                SMALLV1_PH(7);
                double acc[4][4][2];
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) acc[x][y][0] = acc[x][y][1] = 0.0;
                for (int k = j0; k < i; k++) {
                    load_tile<false, false>(Bs, small_tile(sc, i, k), NB, tid, 256);                // L_ik (shared A operand)
                    const bool part = (j < i) && (k >= j);
                    if (part) load_tile<true, false>(grp ? As1 : As0, small_tile(sc, k, j), NB, tid4, 128);   // X_kj^T
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncthreads();
                    if (part) mma_tile<true>(acc, Bs, grp ? As1 : As0, wm, wn, lane);
                    __syncthreads();
                }
                SMALLV1_PH(8);
                // stage right-hand sides: vector = column n of the tile, element m at V[m*ldv + grp*64 + n]
                if (j < i) {
#pragma unroll
                    for (int x = 0; x < 4; x++)
#pragma unroll
                        for (int y = 0; y < 4; y++) {
                            const int m = wm * 32 + x * 8 + r, n = wn * 32 + y * 8 + 2 * c;
                            V[m * SMALLV1_LDV + grp * NB + n] = acc[x][y][0];
                            V[m * SMALLV1_LDV + grp * NB + n + 1] = acc[x][y][1];
                        }
                } else if (j == i) {
                    for (int e = tid4; e < NB * NB; e += 128) V[(e >> 6) * SMALLV1_LDV + grp * NB + (e & 63)] = ((e >> 6) == (e & 63)) ? 1.0 : 0.0;
                }
                load_tile<false>(Bs, small_tile(sc, i, i), NB, tid, 256);                           // L_ii
                __syncthreads();
                if (tid < NB) rd[tid] = 1.0 / Bs[tid * LDT + tid];
                __syncthreads();
                const int jt = j0 + (tid >> 6);          // tile handled by thread tid < 128 in the substitution
                SMALLV1_PH(9);
                if (tid < 2 * NB && jt <= i) subst_lower(Bs, LDT, rd, V, SMALLV1_LDV, tid, jt == i ? ((tid & 63) >> 3) : 0);
                __syncthreads();
                SMALLV1_PH(10);
                if (tid < 2 * NB && jt <= i) {
                    // column sums of squares of X_ij (thread = column) -> g_j ; single writer per (j, n) and round
                    double sg = 0.0;
                    for (int m = 0; m < NB; m++) { double x = V[m * SMALLV1_LDV + tid]; sg = fma(x, x, sg); }
                    gacc[jt * NB + (tid & 63)] += sg;
                }
                if (tid >= 2 * NB && tid < 3 * NB) {
                    // row sums X_ij v_j (thread = row) -> z_i ; one thread adds both tiles of the round (fixed order)
                    const int m = tid - 2 * NB;
                    double sz = 0.0;
                    for (int g2 = 0; g2 < 2; g2++) {
                        const int jj = j0 + g2;
                        if (jj <= i)
                            for (int n = 0; n < NB; n++) sz = fma(V[m * SMALLV1_LDV + g2 * NB + n], vglob[jj * NB + n], sz);
                    }
                    zacc[i * NB + m] += sz;
                }
                if (j <= i) {
                    double* dx = small_tile(sc, i, j);
                    for (int e = tid4; e < NB * NB; e += 128) dx[e] = V[(e >> 6) * SMALLV1_LDV + grp * NB + (e & 63)];
                }
                __syncthreads();
            }
        }
        SMALLV1_PH(11);
        // u_j[n] = sum_{i >= j} sum_m X_ij[m][n] z_i[m]   (X tiles from the scratch)
        for (int e = tid; e < Np; e += 256) {
            const int j = e >> 6, n = e & 63;
            double su = 0.0;
            for (int i = j; i < nt; i++) {
                const double* xt = small_tile(sc, i, j);
                for (int m = 0; m < NB; m++) su = fma(xt[m * NB + n], zacc[i * NB + m], su);
            }
            a.uv[(size_t)id * Np + e] = su;
            a.gv[(size_t)id * Np + e] = gacc[e];
        }
        __syncthreads();
        SMALLV1_PH(0);
    }
#ifdef GPRN_TRACE
    if (threadIdx.x == 0)
        for (int i = 0; i < 12; i++) atomicAdd(&g_small_phase[i], ph_acc[i]);
#endif
}

}  // namespace gprn

// small.cuh -- fused per-matrix pipeline for small and mid-size matrices (Np <= 512, i.e. at most 8x8 tiles of 64).
//
// One persistent CTA takes a matrix through the WHOLE per-iteration chain
//     A = K + diag(D)  ->  L = chol(A)  ->  X = L^-1  ->  g = colnorm2(X),  u = X^T (X v),  logdet(A)
// and only the vectors g, u and the scalar log-det leave the chip.  K is read from HBM exactly once
// (lower tiles); L / X live in a per-CTA scratch of nt(nt+1)/2 tiles -- 10 tiles (320 KB) for Np = 256, which stays
// L2 resident (296 CTAs x 320 KB = 95 MB < 126 MB L2); 36 tiles (1.15 MB) for Np = 512, which streams through HBM --
// and the tiles of the inverse overwrite the L tiles they no longer need.
//
// The triangular solves run in the REGISTERS of
// the warps that hold the right-hand sides as DMMA accumulators (common.cuh: trsm_rows_inreg, shuffles inside a
// quad + DMMA for the off-diagonal blocks) instead of being staged through shared memory for a thread-per-vector
// substitution.  For that a 64x64 tile belongs to four warps as 16 x 64 row slabs (mma_slab), and the inverse factor
// is kept TRANSPOSED in the scratch, Y_ij = X_ij^T: its row sweep  X_ij = -L_ii^-1 sum_k L_ik X_kj  becomes
//     Y_ij L_ii^T = - sum_{k=j}^{i-1} Y_kj L_ik^T ,      Y_ii L_ii^T = I,
// the same right-hand solve as the Cholesky panel step  L_ik L_kk^T = A_ik - sum L_ik' L_kk'^T , with both operands
// of every product read row-major (no transposing loads).  g and z = X v are reduced from the registers
// (row sums of squares: quad shuffles; column sums: butterfly over the eight row lanes), so the solved tile goes to
// the scratch once and is never re-read for them.
// Replaces, for q == 1 and N <= 256 (and for N <= 512 when enough matrices are in flight to give every SM its own:
// gprn_api.cu: decide_small_path), form_a + panel_col + trtri_* + trmv_* (one launch per phase instead of ~10-70 and
// none of their HBM round trips).  256 threads: warps 0-3 and 4-7 work on two tiles at a time.
#pragma once
#include "common.cuh"

namespace gprn {

#define SMALL_MAX_NT 8
#define SMALL_TILES (SMALL_MAX_NT * (SMALL_MAX_NT + 1) / 2)
// scratch tiles keep the padded row stride of the shared-memory operand tiles (LDT), so that ONE TMA bulk copy moves a
// tile verbatim into an operand buffer
#define SMALL_LDS LDT
#define SMALL_TILE_DOUBLES (NB * SMALL_LDS)
#define SMALL_SCRATCH_DOUBLES (SMALL_TILES * SMALL_TILE_DOUBLES)
#define SMALL_CTAS_PER_SM 2
// 3 operand tiles + col(128) + pivs(64) + rd(64) + gacc(Np) + zacc(Np): 112 KB, two CTAs per SM
#define SMALL_SMEM ((3 * NB * LDT + 4 * NB + 2 * SMALL_MAX_NT * NB) * sizeof(double))

#ifdef GPRN_TRACE
__device__ unsigned long long g_small_phase[16];
#define SMALL_PH(i)                                                        \
    do {                                                                   \
        if (threadIdx.x == 0) {                                            \
            long long t_ = clock64();                                      \
            ph_acc[ph_cur] += (unsigned long long)(t_ - ph_last);          \
            ph_last = t_;                                                  \
            ph_cur = (i);                                                  \
        }                                                                  \
    } while (0)
#else
#define SMALL_PH(i)
#endif

struct SmallArgs {
    const double* K;       // [.][Np][Np] assembled covariance matrices (lower tiles)
    const int* ids;        // matrix ids
    int nmat, Np;
    const double* dvec;    // [id][Np] diagonal to add, or null
    const double* vv;      // [id][Np] right-hand side v (needed when do_inverse)
    double* scratch;       // gridDim.x * SMALL_SCRATCH_DOUBLES
    double* uv;            // out [id][Np]  u = X^T X v
    double* gv;            // out [id][Np]  g = colnorm2(X) = diag(A^-1)
    double* logdet;        // out [id]
    int* mstatus;          // out [id], set to 1 on a non-positive pivot
    int do_inverse;
};

__device__ __forceinline__ double* small_tile(double* scratch, int I, int J) {
    return scratch + (size_t)(I * (I + 1) / 2 + J) * SMALL_TILE_DOUBLES;
}

// slab (accumulator layout) <-> a row-major 64 x 64 tile with leading dimension ld (global scratch: NB, shared: LDT)
__device__ __forceinline__ void slab_store(const double (&acc)[2][8][2], double* __restrict__ dst, int ld, int w4, int lane) {
    const int r = lane >> 2, c = lane & 3;
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++)
            *reinterpret_cast<double2*>(dst + (size_t)(16 * w4 + 8 * x + r) * ld + 8 * y + 2 * c) = make_double2(acc[x][y][0], acc[x][y][1]);
}
__device__ __forceinline__ void slab_load(double (&acc)[2][8][2], const double* __restrict__ src, int ld, int w4, int lane) {
    const int r = lane >> 2, c = lane & 3;
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++) {
            const double2 v = *reinterpret_cast<const double2*>(src + (size_t)(16 * w4 + 8 * x + r) * ld + 8 * y + 2 * c);
            acc[x][y][0] = v.x;
            acc[x][y][1] = v.y;
        }
}

// slab_load from the L2 scratch (ld.global.cg: the tile was written by OTHER warps of this CTA)
__device__ __forceinline__ void slab_load_cg(double (&acc)[2][8][2], const double* __restrict__ src, int ld, int w4, int lane) {
    const int r = lane >> 2, c = lane & 3;
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++) {
            const double2 v = __ldcg(reinterpret_cast<const double2*>(src + (size_t)(16 * w4 + 8 * x + r) * ld + 8 * y + 2 * c));
            acc[x][y][0] = v.x;
            acc[x][y][1] = v.y;
        }
}

// One load round of the kernel: the B tile (all warps) and up to two A tiles (one per warp group) from the scratch.
// TMA: thread 0 issues one bulk copy per tile (34 816 bytes) on the mbarrier and every thread waits for the phase; otherwise each
// thread issues its share as 16-byte cp.async copies and the CTA meets at a barrier.  The caller guarantees (by the
// barrier that ended the previous use) that nobody still reads the destination buffers.
template <bool TMA>
__device__ __forceinline__ void small_load_round(double* Bs, const double* srcB, double* A0, const double* srcA0,
                                                 double* A1, const double* srcA1, int tid, unsigned long long* bar,
                                                 unsigned& phase, int* bad) {
    if (TMA) {
        if (tid == 0) {
            constexpr unsigned TB = (unsigned)(SMALL_TILE_DOUBLES * sizeof(double));
            fence_proxy_async();
            mbar_expect_tx(bar, TB * (1u + (srcA0 != nullptr) + (srcA1 != nullptr)));
            bulk_g2s(Bs, srcB, TB, bar);
            if (srcA0) bulk_g2s(A0, srcA0, TB, bar);
            if (srcA1) bulk_g2s(A1, srcA1, TB, bar);
        }
        if (!mbar_wait(bar, phase)) *bad = 2;
        phase ^= 1u;
    } else {
        load_tile<false, false>(Bs, srcB, SMALL_LDS, tid, 256);
        if (tid < 128) { if (srcA0) load_tile<false, false>(A0, srcA0, SMALL_LDS, tid, 128); }
        else if (srcA1) load_tile<false, false>(A1, srcA1, SMALL_LDS, tid - 128, 128);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    }
}

template <bool TMA>
__global__ void __launch_bounds__(256, 2) small_pipeline_kernel(SmallArgs a) {
    GPRN_TRACE_SCOPE(TK_SMALL);
    extern __shared__ double smem[];
    double* Bs = smem;                 // B operand / potrf input + output (L_kk) / L_ii of the solves
    double* As0 = smem + NB * LDT;     // A operand of warps 0-3; after the products: per-warp column sums (zpart)
    double* As1 = smem + 2 * NB * LDT; // A operand of warps 4-7; their tile waits here while potrf64 runs
    double* col = smem + 3 * NB * LDT;
    double* pivs = col + 2 * NB;
    double* rd = pivs + NB;
    double* gacc = rd + NB;            // [Np]
    double* zacc = gacc + SMALL_MAX_NT * NB;
    double* zpart = As0;               // [8 warps][64]
    __shared__ int bad;
    __shared__ unsigned long long tma_bar;          // mbarrier of the bulk tile loads (one phase per load round)
    unsigned phase = 0;
    const int Np = a.Np, nt = Np / NB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp >> 2, w4 = warp & 3, tid4 = tid & 127;
    const int r = lane >> 2, c = lane & 3;
    double* Ag = grp ? As1 : As0;
    double* sc = a.scratch + (size_t)blockIdx.x * SMALL_SCRATCH_DOUBLES;
    if (TMA) {
        if (tid == 0) mbar_init(&tma_bar, 1);
        __syncthreads();
    }
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
        // the first tile of a sub-round, warps 4-7 the second.
        for (int k = 0; k < nt; k++) {
            for (int s = 0; k + 2 * s < nt; s++) {
                const int it = k + 2 * s + grp;
                const bool have = it < nt;
                const bool diag = (s == 0 && grp == 0);
                SMALL_PH(1);
                double acc[2][8][2];
#pragma unroll
                for (int x = 0; x < 2; x++)
#pragma unroll
                    for (int y = 0; y < 8; y++) {
                        const int m = 16 * w4 + 8 * x + r, n = 8 * y + 2 * c;
                        double2 v = make_double2(0.0, 0.0);
                        if (have) {
                            v = *reinterpret_cast<const double2*>(Km + (size_t)(it * NB + m) * Np + k * NB + n);
                            if (dv && diag) {
                                if (m == n) v.x += dv[k * NB + m];
                                if (m == n + 1) v.y += dv[k * NB + m];
                            }
                        }
                        acc[x][y][0] = v.x;
                        acc[x][y][1] = v.y;
                    }
                {   // the K tile of this group's NEXT sub-round goes to L2 now: its HBM latency hides behind this round
                    int kn = k, sn = s + 1;
                    if (k + 2 * sn >= nt) { kn = k + 1; sn = 0; }
                    const int itn = kn + 2 * sn + grp;
                    if (kn < nt && itn < nt) {
                        const double* nx = Km + (size_t)(itn * NB + (tid4 >> 1)) * Np + kn * NB + (tid4 & 1) * 32;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 16));
                    }
                }
                SMALL_PH(2);
                for (int kp = 0; kp < k; kp++) {
                    const int it0 = k + 2 * s;           // tile of warps 0-3 (the diagonal one when s == 0), it0 + 1: warps 4-7
                    small_load_round<TMA>(Bs, small_tile(sc, k, kp), As0, s > 0 ? small_tile(sc, it0, kp) : nullptr,
                                          As1, it0 + 1 < nt ? small_tile(sc, it0 + 1, kp) : nullptr, tid, &tma_bar, phase, &bad);
                    if (have) mma_slab<true>(acc, diag ? Bs : Ag, Bs, w4, lane);
                    __syncthreads();
                }
                SMALL_PH(3);
                if (s == 0) {
                    // the diagonal tile goes to Bs for potrf64; the tile below it waits in As1 (its accumulators
                    // would not survive the register demand of the factorisation)
                    if (grp == 0) slab_store(acc, Bs, LDT, w4, lane);
                    else if (have) slab_store(acc, As1, LDT, w4, lane);
                    __syncthreads();
                    SMALL_PH(4);
                    potrf64(Bs, LDT, Bs, rd, col, pivs, &bad);
                    SMALL_PH(5);
                    if (tid < 32) logsum += log(pivs[tid]) + log(pivs[tid + 32]);
                    // every solve of the column uses 1 / L_kk by IEEE division (the sub-rounds s > 0 and mid.cuh reload
                    // L_kk and have nothing else), not potrf64's Newton reciprocal
                    if (tid < NB) rd[tid] = 1.0 / Bs[tid * LDT + tid];
                    __syncthreads();
                    if (grp == 0) {
                        double* dkk = small_tile(sc, k, k);
                        for (int e = tid4; e < NB * (NB / 2); e += 128) {
                            const int m = e >> 5, c2 = e & 31;
                            *reinterpret_cast<double2*>(dkk + m * SMALL_LDS + 2 * c2) = *reinterpret_cast<const double2*>(Bs + m * LDT + 2 * c2);
                        }
                        if (TMA) fence_proxy_async();
                    } else if (have) {
                        slab_load(acc, As1, LDT, w4, lane);
                        trsm_rows_inreg(acc, Bs, rd, lane);
                        SMALL_PH(6);
                        slab_store(acc, small_tile(sc, it, k), SMALL_LDS, w4, lane);
                        if (TMA) fence_proxy_async();
                    }
                } else {
                    small_load_round<TMA>(Bs, small_tile(sc, k, k), nullptr, nullptr, nullptr, nullptr, tid, &tma_bar, phase, &bad);   // L_kk back
                    if (tid < NB) rd[tid] = 1.0 / Bs[tid * LDT + tid];
                    __syncthreads();
                    SMALL_PH(5);
                    if (have) {
                        trsm_rows_inreg(acc, Bs, rd, lane);
                        SMALL_PH(6);
                        slab_store(acc, small_tile(sc, it, k), SMALL_LDS, w4, lane);
                        if (TMA) fence_proxy_async();
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
        SMALL_PH(0);
        if (!a.do_inverse) continue;

        // ================= inverse by block rows, transposed: Y_ij = X_ij^T overwrites L_ij =================
        const double* vglob = a.vv + (size_t)id * Np;
        for (int e = tid; e < Np; e += 256) {
            gacc[e] = 0.0;
            zacc[e] = 0.0;
        }
        __syncthreads();
        for (int i = 0; i < nt; i++) {
            for (int j0 = 0; j0 <= i; j0 += 2) {
                const int j = j0 + grp;                 // this group's column tile (j == i: identity right-hand side)
                const bool act = j <= i;
                SMALL_PH(7);
                double acc[2][8][2];
#pragma unroll
                for (int x = 0; x < 2; x++)
#pragma unroll
                    for (int y = 0; y < 8; y++) acc[x][y][0] = acc[x][y][1] = 0.0;
                for (int k = j0; k < i; k++) {
                    const bool part = (j < i) && (k >= j);
                    // L_ik (B operand of both groups) and the groups' A operands Y_k,j0 / Y_k,j0+1
                    small_load_round<TMA>(Bs, small_tile(sc, i, k), As0, small_tile(sc, k, j0),
                                          As1, (j0 + 1 < i && k >= j0 + 1) ? small_tile(sc, k, j0 + 1) : nullptr, tid, &tma_bar, phase, &bad);
                    if (part) mma_slab<true>(acc, Ag, Bs, w4, lane);                              // - sum Y_kj L_ik^T
                    __syncthreads();
                }
                SMALL_PH(8);
                if (j == i) {
#pragma unroll
                    for (int x = 0; x < 2; x++)
#pragma unroll
                        for (int y = 0; y < 8; y++) {
                            const int n = 16 * w4 + 8 * x + r, m = 8 * y + 2 * c;
                            acc[x][y][0] = (n == m) ? 1.0 : 0.0;
                            acc[x][y][1] = (n == m + 1) ? 1.0 : 0.0;
                        }
                }
                small_load_round<TMA>(Bs, small_tile(sc, i, i), nullptr, nullptr, nullptr, nullptr, tid, &tma_bar, phase, &bad);   // L_ii
                if (tid < NB) rd[tid] = 1.0 / Bs[tid * LDT + tid];
                __syncthreads();
                SMALL_PH(9);
                if (act) {
                    trsm_rows_inreg(acc, Bs, rd, lane, j == i ? 2 * w4 : 0);
                    SMALL_PH(10);
                    // g_j[n] += sum_m Y_ij[n][m]^2 : the four lanes of a quad hold one row
                    double vn[2];
#pragma unroll
                    for (int x = 0; x < 2; x++) {
                        double sg = 0.0;
#pragma unroll
                        for (int y = 0; y < 8; y++) {
                            sg = fma(acc[x][y][0], acc[x][y][0], sg);
                            sg = fma(acc[x][y][1], acc[x][y][1], sg);
                        }
                        sg += __shfl_xor_sync(0xffffffffu, sg, 1);
                        sg += __shfl_xor_sync(0xffffffffu, sg, 2);
                        const int n = 16 * w4 + 8 * x + r;
                        if (c == 0) gacc[j * NB + n] += sg;          // single writer per (j, n) and round
                        vn[x] = vglob[j * NB + n];
                    }
                    // z_i[m] += sum_n Y_ij[n][m] v_j[n] : per-warp column sums (butterfly over the row lanes)
#pragma unroll
                    for (int y = 0; y < 8; y++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            double t = acc[0][y][e] * vn[0];
                            t = fma(acc[1][y][e], vn[1], t);
                            t += __shfl_xor_sync(0xffffffffu, t, 4);
                            t += __shfl_xor_sync(0xffffffffu, t, 8);
                            t += __shfl_xor_sync(0xffffffffu, t, 16);
                            if (r == 0) zpart[warp * NB + 8 * y + 2 * c + e] = t;
                        }
                    slab_store(acc, small_tile(sc, i, j), SMALL_LDS, w4, lane);
                    if (TMA) fence_proxy_async();
                }
                __syncthreads();
                if (tid < NB) {        // fixed order: the four slabs of the first tile, then those of the second
                    double t = (zpart[tid] + zpart[NB + tid]) + (zpart[2 * NB + tid] + zpart[3 * NB + tid]);
                    if (j0 + 1 <= i) t += (zpart[4 * NB + tid] + zpart[5 * NB + tid]) + (zpart[6 * NB + tid] + zpart[7 * NB + tid]);
                    zacc[i * NB + tid] += t;
                }
                __syncthreads();
            }
        }
        SMALL_PH(11);
        // u_j[n] = sum_{i >= j} sum_m Y_ij[n][m] z_i[m] : thread = (row n, quarter of the row), eight independent
        // 16-byte loads in flight per thread and tile (the tiles come from the L2 scratch); the four quarters of a
        // row are combined by quad shuffles
        {
            const int n = tid >> 2, qd = tid & 3;
            for (int j = 0; j < nt; j++) {
                double su = 0.0;
                for (int i = j; i < nt; i++) {
                    const double* yrow = small_tile(sc, i, j) + n * SMALL_LDS + qd * 16;
                    const double* zi = zacc + i * NB + qd * 16;
                    double2 v[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) v[e] = *reinterpret_cast<const double2*>(yrow + 2 * e);
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        su = fma(v[e].x, zi[2 * e], su);
                        su = fma(v[e].y, zi[2 * e + 1], su);
                    }
                }
                su += __shfl_xor_sync(0xffffffffu, su, 1);
                su += __shfl_xor_sync(0xffffffffu, su, 2);
                if (qd == 0) {
                    a.uv[(size_t)id * Np + j * NB + n] = su;
                    a.gv[(size_t)id * Np + j * NB + n] = gacc[j * NB + n];
                }
            }
        }
        __syncthreads();
        SMALL_PH(0);
    }
#ifdef GPRN_TRACE
    if (threadIdx.x == 0)
        for (int i = 0; i < 12; i++) atomicAdd(&g_small_phase[i], ph_acc[i]);
#endif
}

}  // namespace gprn

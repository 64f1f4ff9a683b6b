// mid.cuh -- dataflow kernels for small / mid-size matrices (Np <= 1024, any q): one matrix is factored and inverted
// by nt CTAs at once, coupled by per-tile flags in global memory, instead of by one CTA walking all tiles (small.cuh:
// the throughput path for thousands of q = 1 matrices) or by ~20 dependent launches per matrix (factor.cuh).
// Latency mode: the single-evaluation case -- one ELBOcalc inside an optimiser or sampler step (BASELINE configs C1 /
// C2) -- where the GPU holds M = q(p+1) matrices, every CTA is resident and the evaluation is ONE dependency chain.
// Throughput mode: more CTAs than the GPU holds (tens of walkers, q > 1 batches); CTAs numbered by a start-order ticket.
//
//   mid_pipeline_kernel<NW>, grid = (nt, matrices); four warps own the four 16 x 64 row slabs of a tile (NW = 8: four
//   more warps that only stage tiles and join potrf64):
//     Cholesky, left-looking: CTA r owns tile ROW r.  For k = 0..r-1 it forms
//         T_rk = A_rk - sum_{k' < k} L_rk' L_kk'^T
//       (A fragments straight from its own finished tiles in L2; B tile L_kk' of CTA k through two cp.async buffers
//       in shared memory, after that tile's flag is up), solves against L_kk (flag of CTA k) in registers
//       (trsm_rows_inreg), stores L_rk and raises its flag -- and subtracts L_rk L_rk^T from the DIAGONAL tile's own
//       accumulator at once, so that after the row's last solve one product, not r, precedes potrf64.
//     Inverse, transposed (Y_ij = X_ij^T, see small.cuh): CTA c owns tile COLUMN c,
//         Y_ic L_ii^T = [i == c] I - sum_{k=c}^{i-1} Y_kc L_ik^T ,   i = c .. nt-1,
//       whose only inputs from other CTAs are Cholesky tiles (flags) -- the columns of the inverse are independent
//       chains.  While L_ii is not there yet the CTA works ahead on row i + 1 in a second accumulator.  Y goes to the X
//       buffer (L tiles are still being read by the other columns).  g_c (row sums of squares) is complete inside the
//       CTA; the column sums that make up z = X v are written per (column, row tile).
//   mid_finish_kernel, grid = (nt, matrices): adds the z partials and the per-row log-det partials in a fixed order,
//     forms u_c = sum_{i >= c} Y_ic z_i and resets the flags (and the ticket) for the next launch.
//   mid_trmv_lower_kernel: z = X v on transposed tiles (prior quadratic forms, q > 1).
// Dependencies only point to CTAs with a smaller tile row of the same matrix (Cholesky) or to Cholesky tiles
// (inverse).  While all CTAs of a launch are co-resident (the latency case) that is enough; a launch with more CTAs
// than the GPU holds (throughput mode: tens of walkers at N ~ 500, q > 1 batches) numbers its CTAs by a start-order
// ticket instead of the block index, so a CTA only ever waits for CTAs that started before it.  The waits are bounded
// anyway (a protocol error becomes mstatus = 2, not a hung GPU).
// Same building blocks and the same summation orders as small.cuh: an evaluation gives bit-identical results whether
// it runs alone (this path) or inside a large batch (small.cuh) -- tests/test_gpu_parity.py::test_c3_full_size_properties.
#pragma once
#include "common.cuh"
#include "small.cuh"

namespace gprn {

#define MID_MAX_NT 16
#define MID_TILES (MID_MAX_NT * (MID_MAX_NT + 1) / 2)
// P (potrf64 / L_kk / L_ii) + Bm (B operand, two buffers) + col(128) + pivs(64) + rd(64) + gacc(64) + zpart(4 x 64)
#define MID_SMEM ((3 * NB * LDT + 4 * NB + NB + 4 * NB) * sizeof(double))

struct MidArgs {
    const double* K;       // [.][Np][Np] assembled covariance matrices (lower tiles)
    double* W;             // [.][Np][Np] out: Cholesky factor of K + diag(dvec), lower tiles, row-major
    double* X;             // [.][Np][Np] out: transposed tiles of the inverse factor (tile (i, j) holds X_ij^T)
    const int* ids;        // matrix ids
    int Np;
    const double* dvec;    // [id][Np] diagonal to add, or null
    const double* vv;      // [id][Np] right-hand side v (with do_inverse; null together with uv: no z / u)
    double* uv;            // out [id][Np]  u = X^T X v           (mid_finish_kernel)
    double* gv;            // out [id][Np]  g = colnorm2(X) = diag(A^-1)
    double* logdet;        // out [id]                            (mid_finish_kernel)
    int* mstatus;          // out [id], 1: non-positive pivot, 2: flag wait timed out
    int* tstate;           // [id][MID_TILES] tile flags, zero at launch (reset by mid_finish_kernel): 1 = L tile final
    double* zp;            // [id][MID_MAX_NT (column)][MID_MAX_NT (row tile)][64] column sums of Y_ic weighted by v_c
    double* ldpart;        // [id][MID_MAX_NT][32] per-row, per-lane log-det partials (summed as small.cuh sums them)
    int do_inverse;
    int* ticket;           // null, or a counter (zero at launch, reset by mid_finish_kernel) that orders the CTAs
};

__device__ __forceinline__ int mid_tile_index(int I, int J) { return I * (I + 1) / 2 + J; }

#ifdef GPRN_TRACE
// thread-0 clock64 by phase and CTA row (development aid): 0 other, 1 flag waits, 2 K loads, 3 products, 4 potrf64,
// 5 solves, 6 stores + publish, 7 inverse reductions
__device__ unsigned long long g_mid_phase[2][MID_MAX_NT][8];   // [0: set-up launches, 1: iteration launches]
#define MID_PH(i)                                                          \
    do {                                                                   \
        if (threadIdx.x == 0) {                                            \
            long long t_ = clock64();                                      \
            ph_acc[ph_cur] += (unsigned long long)(t_ - ph_last);          \
            ph_last = t_;                                                  \
            ph_cur = (i);                                                  \
        }                                                                  \
    } while (0)
#else
#define MID_PH(i)
#endif

// All threads of the CTA return once flag[idx] >= want (thread 0 polls).  The tile behind the flag was written by
// another CTA: readers use L2 loads (ld.global.cg / cp.async.cg).
__device__ __forceinline__ void mid_wait(const int* flags, int idx, int want, int* timed_out) {
    if (threadIdx.x == 0) {
        const volatile int* f = flags + idx;
        int spins = 0;
        while (*f < want) {
            __nanosleep(32);
            if (++spins > (1 << 22)) { *timed_out = 1; break; }
        }
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ void mid_publish(int* flags, int idx, int value) {
    __threadfence();                   // every thread: its tile stores are visible device-wide ...
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(flags + idx) = value;      // ... before the flag is
}

// acc -= sum_{t = t0}^{t1 - 1} A_t B_t^T.  A_t = a_tile(t): a finished tile of this CTA, fragments straight from L2
// (mma_slab_ga); B_t = b_tile(t): staged through the two halves of Bm, the copy of step t + 1 in flight during the
// products of step t.  ready(t) returns, for all threads and behind a CTA barrier, once B_t may be read (the flag of
// another CTA's tile, or nothing for an own tile); that barrier also retires the last reader of the buffer about to
// be overwritten and orders this CTA's own earlier tile stores before the loads.
// mma_slab_ga with the A fragments of the first four k-steps handed in (`an`, loaded by the previous call or by
// mid_a_prefetch) and, while the last chunk is being multiplied, the first four k-steps of the NEXT product's A tile
// loaded into `an` again (Anext; null: none) -- the L2 latency of a product's first fragments is then hidden behind
// the previous product instead of following the CTA barrier that precedes every product.  Same DMMA sequence.
__device__ __forceinline__ void mid_a_prefetch(double (&an)[4][2], const double* __restrict__ Ag, int lda, int w4, int lane) {
    const double* ap = Ag + (size_t)(w4 * 16 + (lane >> 2)) * lda + (lane & 3);
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int x = 0; x < 2; x++) an[u][x] = __ldcg(ap + (size_t)x * 8 * lda + 4 * u);
}
template <bool NEG>
__device__ __forceinline__ void mma_slab_ga_pf(double (&acc)[2][8][2], double (&an)[4][2], const double* __restrict__ Ag,
                                               const double* __restrict__ Anext, int lda, const double* __restrict__ Bs,
                                               int w4, int lane) {
    const int r = lane >> 2, c = lane & 3;
    const double* ap = Ag + (size_t)(w4 * 16 + r) * lda + c;
    const double* bp = Bs + r * LDT + c;
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
        double ac[4][2];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int x = 0; x < 2; x++) ac[u][x] = NEG ? -an[u][x] : an[u][x];
        if (ch < 3) {
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int x = 0; x < 2; x++) an[u][x] = __ldcg(ap + (size_t)x * 8 * lda + 16 * (ch + 1) + 4 * u);
        } else if (Anext) {
            mid_a_prefetch(an, Anext, lda, w4, lane);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k0 = 16 * ch + 4 * u;
            double b[8];
#pragma unroll
            for (int y = 0; y < 8; y++) b[y] = bp[y * 8 * LDT + k0];
#pragma unroll
            for (int x = 0; x < 2; x++)
#pragma unroll
                for (int y = 0; y < 8; y++) dmma884(acc[x][y], ac[u][x], b[y]);
        }
    }
}

// NTHR threads stage the B tiles; the warps with `compute` set (the four slab owners) multiply.
template <int NTHR, class FA, class FB, class FR>
__device__ __forceinline__ void mid_products(double (&acc)[2][8][2], int t0, int t1, FA a_tile, FB b_tile, FR ready,
                                             double* Bm, int Np, int tid, int w4, int lane, bool compute) {
    if (t0 >= t1) return;
    ready(t0);
    load_tile<false, false>(Bm, b_tile(t0), Np, tid, NTHR);
    cp_async_commit();
    double an[4][2];
    if (compute) mid_a_prefetch(an, a_tile(t0), Np, w4, lane);
    for (int t = t0; t < t1; t++) {
        const int s = t - t0;
        if (t + 1 < t1) {
            ready(t + 1);
            load_tile<false, false>(Bm + ((s + 1) & 1) * NB * LDT, b_tile(t + 1), Np, tid, NTHR);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (compute)
            mma_slab_ga_pf<true>(acc, an, a_tile(t), t + 1 < t1 ? a_tile(t + 1) : nullptr, Np, Bm + (s & 1) * NB * LDT, w4, lane);
    }
    __syncthreads();
}

// NW = 4: 128 threads, two CTAs per SM (up to 2 x SMs CTAs in flight).  NW = 8: 256 threads, one CTA per SM -- warps
// 4-7 only stage tiles and join potrf64 (13 k instead of ~20 k cycles per diagonal tile: the link of the dependency
// chain that every tile row waits for); the tile arithmetic stays with warps 0-3, so the bits are the same.
template <int NW>
__global__ void __launch_bounds__(32 * NW, NW == 4 ? 2 : 1) mid_pipeline_kernel(MidArgs a) {
    constexpr int NTHR = 32 * NW;
    GPRN_TRACE_SCOPE(TK_SMALL);
    extern __shared__ double smem[];
    double* P = smem;
    double* Bm = smem + NB * LDT;      // two buffers (mid_products)
    double* col = smem + 3 * NB * LDT;
    double* pivs = col + 2 * NB;
    double* rd = pivs + NB;
    double* gacc = rd + NB;            // [64] g of this CTA's column
    double* zpart = gacc + NB;         // [4 warps][64]
    __shared__ int bad, timed_out, peek;
    const int Np = a.Np, nt = Np / NB;
    const int tid = threadIdx.x, w4 = (tid >> 5) & 3, lane = tid & 31;
    const bool compute = tid < 128;                    // the four warps that own the 16 x 64 slabs of a tile
    const int r = lane >> 2, c = lane & 3;
    // Which (tile row, matrix) this CTA works on: the order in which the CTAs START, taken from a ticket counter --
    // the rows a CTA waits for then belong to CTAs that have started before it, whatever order the hardware dispatches
    // blocks in and however many of them are resident, so the flag waits cannot deadlock.  (ticket == null: block
    // index; only safe when every CTA of the launch is resident.)
    int me = blockIdx.x, mat = blockIdx.y;
    if (a.ticket) {
        __shared__ int s_ticket;
        if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1);
        __syncthreads();
        me = s_ticket % (int)gridDim.x;                // tile row (Cholesky) / tile column (inverse)
        mat = s_ticket / (int)gridDim.x;
    }
    const int id = a.ids[mat];
    const double* Km = a.K + (size_t)id * Np * Np;
    double* Wm = a.W + (size_t)id * Np * Np;
    double* Xm = a.X + (size_t)id * Np * Np;
    const double* dv = a.dvec ? a.dvec + (size_t)id * Np : nullptr;
    int* flags = a.tstate + (size_t)id * MID_TILES;
    if (tid == 0) { bad = 0; timed_out = 0; }
    if (tid < NB) gacc[tid] = 0.0;
    __syncthreads();
#ifdef GPRN_TRACE
    unsigned long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_last = clock64();
    int ph_cur = 0;
#endif
#define MID_TILE(base, I, J) ((base) + (size_t)((I) * NB) * Np + (J) * NB)
    auto load_k = [&](double (&t)[2][8][2], int J) {   // K_me,J (+ D on the diagonal tile) in accumulator layout
#pragma unroll
        for (int x = 0; x < 2; x++)
#pragma unroll
            for (int y = 0; y < 8; y++) {
                const int m = 16 * w4 + 8 * x + r, n = 8 * y + 2 * c;
                double2 v = *reinterpret_cast<const double2*>(Km + (size_t)(me * NB + m) * Np + J * NB + n);
                if (dv && J == me) {
                    if (m == n) v.x += dv[J * NB + m];
                    if (m == n + 1) v.y += dv[J * NB + m];
                }
                t[x][y][0] = v.x;
                t[x][y][1] = v.y;
            }
    };
    auto prefetch_k = [&](int J) {     // K_me,J to L2: its HBM latency hides behind the current tile's work
        if (tid >= 128) return;
        const double* nx = Km + (size_t)(me * NB + (tid >> 1)) * Np + J * NB + (tid & 1) * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 16));
    };

    // ================= Cholesky: tile row `me`, columns k = 0 .. me =================
    // The diagonal tile's accumulator accd lives beside the current column's: L_me,k is subtracted from it as soon as
    // it exists (same order k = 0, 1, .. as the left-looking sum), so that after the last solve of the row only ONE
    // product stands between L_me,me-1 and potrf64 -- the row's critical path -- instead of `me` of them.
    double accd[2][8][2];
    MID_PH(2);
    if (compute) load_k(accd, me);
    if (me > 0) prefetch_k(0);
    for (int k = 0; k < me; k++) {
        MID_PH(2);
        double acc[2][8][2];
        if (compute) load_k(acc, k);
        if (k + 1 < me) prefetch_k(k + 1);
        MID_PH(3);
        mid_products<NTHR>(acc, 0, k,
                     [&](int kp) { return MID_TILE(Wm, me, kp); },
                     [&](int kp) { return MID_TILE(Wm, k, kp); },                                  // L_k,kp of CTA k
                     [&](int kp) { mid_wait(flags, mid_tile_index(k, kp), 1, &timed_out); },
                     Bm, Np, tid, w4, lane, compute);
        MID_PH(1);
        mid_wait(flags, mid_tile_index(k, k), 1, &timed_out);
        MID_PH(5);
        load_tile<false>(P, MID_TILE(Wm, k, k), Np, tid, NTHR);
        __syncthreads();
        if (tid < NB) rd[tid] = 1.0 / P[tid * LDT + tid];
        __syncthreads();
        if (compute) {
            trsm_rows_inreg(acc, P, rd, lane);
            MID_PH(6);
            slab_store(acc, MID_TILE(Wm, me, k), Np, w4, lane);
            slab_store(acc, Bm, LDT, w4, lane);        // and kept on chip: both operands of the diagonal update
        }
        // the diagonal update comes first -- it is what potrf64 of this row waits for -- while the global stores of
        // L_me,k drain; the flag follows (the other rows have the time of a potrf64 to spare before they need the tile)
        __syncthreads();
        MID_PH(3);
        if (compute) mma_slab<true>(accd, Bm, Bm, w4, lane);      // the DMMA sequence of mma_slab_ga: same bits
        MID_PH(6);
        mid_publish(flags, mid_tile_index(me, k), 1);
    }
    {
        MID_PH(4);
        __syncthreads();
        if (compute) slab_store(accd, P, LDT, w4, lane);
        __syncthreads();
        potrf64_t<NW>(P, LDT, P, rd, col, pivs, &bad, tid);
        MID_PH(6);
        if (tid < 32) a.ldpart[((size_t)id * MID_MAX_NT + me) * 32 + tid] = log(pivs[tid]) + log(pivs[tid + 32]);
        double* dkk = MID_TILE(Wm, me, me);
        for (int e = tid; e < NB * (NB / 2); e += NTHR) {
            const int m = e >> 5, c2 = e & 31;
            *reinterpret_cast<double2*>(dkk + (size_t)m * Np + 2 * c2) = *reinterpret_cast<const double2*>(P + m * LDT + 2 * c2);
        }
        mid_publish(flags, mid_tile_index(me, me), 1);
    }
    MID_PH(0);
    if (tid == 0 && (bad || timed_out)) a.mstatus[id] = timed_out ? 2 : 1;
    if (!a.do_inverse) {
#ifdef GPRN_TRACE
        if (threadIdx.x == 0)
            for (int i = 0; i < 8; i++) atomicAdd(&g_mid_phase[0][me][i], ph_acc[i]);
#endif
        return;
    }

    // ================= inverse: tile column `me`, rows i = me .. nt-1 =================
    // Row i is  Y_i,me L_ii^T = [i == me] I - sum_{k = me}^{i-1} Y_k,me L_ik^T.  While L_ii is not there yet the CTA
    // works ahead on row i + 1 in a second accumulator (its terms k < i: all inputs exist), so that once L_ii arrives
    // one product, one solve and the reductions finish the row.  Every accumulator still receives its terms in
    // ascending k: the bits do not depend on how far ahead the CTA got.
    const double* vglob = a.vv ? a.vv + (size_t)id * Np : nullptr;      // null: inverse and g only (set-up, q > 1)
    double acc[2][8][2], accn[2][8][2];
    int kn = me;                       // accn holds the terms k < kn of row i + 1
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++) {
            const int n = 16 * w4 + 8 * x + r, m = 8 * y + 2 * c;
            acc[x][y][0] = (n == m) ? 1.0 : 0.0;
            acc[x][y][1] = (n == m + 1) ? 1.0 : 0.0;
            accn[x][y][0] = accn[x][y][1] = 0.0;
        }
    for (int i = me; i < nt; i++) {
        MID_PH(3);
        int k0 = kn;                   // acc holds the terms k < k0 of row i
        mid_products<NTHR>(acc, k0, i,
                     [&](int k) { return MID_TILE(Xm, k, me); },                                   // own tiles
                     [&](int k) { return MID_TILE(Wm, i, k); },                                    // L_ik of CTA i
                     [&](int k) { mid_wait(flags, mid_tile_index(i, k), 1, &timed_out); },
                     Bm, Np, tid, w4, lane, compute);
        kn = me;
        if (i + 1 < nt) {
            if (tid == 0) peek = *reinterpret_cast<const volatile int*>(flags + mid_tile_index(i, i));
            __syncthreads();
            if (peek < 1) {            // CTA-uniform
                mid_products<NTHR>(accn, me, i,
                             [&](int k) { return MID_TILE(Xm, k, me); },
                             [&](int k) { return MID_TILE(Wm, i + 1, k); },
                             [&](int k) { mid_wait(flags, mid_tile_index(i + 1, k), 1, &timed_out); },
                             Bm, Np, tid, w4, lane, compute);
                kn = i;
            }
        }
        MID_PH(1);
        mid_wait(flags, mid_tile_index(i, i), 1, &timed_out);
        MID_PH(5);
        load_tile<false>(P, MID_TILE(Wm, i, i), Np, tid, NTHR);              // L_ii
        __syncthreads();
        if (tid < NB) rd[tid] = 1.0 / P[tid * LDT + tid];
        __syncthreads();
        if (compute) {
            trsm_rows_inreg(acc, P, rd, lane, i == me ? 2 * w4 : 0);
            MID_PH(7);
            // g_me[n] += sum_m Y[n][m]^2 : the four lanes of a quad hold one row
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
                if (c == 0) gacc[n] += sg;                   // single writer per n
                vn[x] = vglob ? vglob[me * NB + n] : 0.0;
            }
            // column sums of Y_i,me weighted by v_me: this column's share of z_i
#pragma unroll
            for (int y = 0; y < 8; y++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    double t = acc[0][y][e] * vn[0];
                    t = fma(acc[1][y][e], vn[1], t);
                    t += __shfl_xor_sync(0xffffffffu, t, 4);
                    t += __shfl_xor_sync(0xffffffffu, t, 8);
                    t += __shfl_xor_sync(0xffffffffu, t, 16);
                    if (r == 0) zpart[w4 * NB + 8 * y + 2 * c + e] = t;
                }
        }
        MID_PH(6);
        if (compute) slab_store(acc, MID_TILE(Xm, i, me), Np, w4, lane);
        __syncthreads();
        if (tid < NB)
            a.zp[(((size_t)id * MID_MAX_NT + me) * MID_MAX_NT + i) * NB + tid] =
                (zpart[tid] + zpart[NB + tid]) + (zpart[2 * NB + tid] + zpart[3 * NB + tid]);
#pragma unroll
        for (int x = 0; x < 2; x++)
#pragma unroll
            for (int y = 0; y < 8; y++) {
                acc[x][y][0] = accn[x][y][0];
                acc[x][y][1] = accn[x][y][1];
                accn[x][y][0] = accn[x][y][1] = 0.0;
            }
        __syncthreads();
    }
    MID_PH(0);
#ifdef GPRN_TRACE
    if (threadIdx.x == 0)
        for (int i = 0; i < 8; i++) atomicAdd(&g_mid_phase[1][me][i], ph_acc[i]);
#endif
    if (tid < NB) a.gv[(size_t)id * Np + me * NB + tid] = gacc[tid];
    if (tid == 0 && timed_out) a.mstatus[id] = 2;
#undef MID_TILE
}

// z_i = sum_{j <= i} zp[j][i], u_c = sum_{i >= c} Y_ic z_i, logdet = sum_r ldpart[r] -- each in exactly the order in
// which small.cuh adds the same terms, so that an evaluation gives the same bits on either path.
// grid = (nt, matrices), block = 256: thread = (row n of the column's tiles, quarter of the row).
__global__ void __launch_bounds__(256) mid_finish_kernel(MidArgs a) {
    __shared__ double z[MID_MAX_NT * NB];
    const int Np = a.Np, nt = Np / NB, tid = threadIdx.x, me = blockIdx.x;
    const int id = a.ids[blockIdx.y];
    // the pipeline launch is over: its tile flags go back to zero for the next one (they are zero when allocated)
    if (me == 0 && tid >= 64 && tid < 64 + MID_TILES) a.tstate[(size_t)id * MID_TILES + tid - 64] = 0;
    if (a.ticket && me == 0 && blockIdx.y == 0 && tid == 0) *a.ticket = 0;
    if (me == 0 && tid < 32) {         // lane-wise over the rows, then the warp butterfly: the order of small.cuh
        double s = 0.0;
        for (int r = 0; r < nt; r++) s += a.ldpart[((size_t)id * MID_MAX_NT + r) * 32 + tid];
        s = warp_sum(s);
        if (tid == 0) a.logdet[id] = s;
    }
    if (!a.do_inverse || !a.uv) return;
    for (int e = tid; e < (nt - me) * NB; e += 256) {
        const int i = me + (e >> 6), m = e & 63;
        double zj[MID_MAX_NT];                         // all partials in flight at once, then added in order
#pragma unroll
        for (int j = 0; j < MID_MAX_NT; j++)
            zj[j] = j <= i ? __ldcg(a.zp + (((size_t)id * MID_MAX_NT + j) * MID_MAX_NT + i) * NB + m) : 0.0;
        double s = 0.0;
#pragma unroll
        for (int j0 = 0; j0 < MID_MAX_NT; j0 += 2) {   // pairs of column tiles, as small.cuh adds them
            if (j0 <= i) {
                double t = zj[j0];
                if (j0 + 1 <= i) t += zj[j0 + 1];
                s += t;
            }
        }
        z[i * NB + m] = s;
    }
    __syncthreads();
    const double* Xm = a.X + (size_t)id * Np * Np;
    const int n = tid >> 2, qd = tid & 3;
    double su = 0.0;
    double2 v[8], w[8];                // the next tile's row segment is in flight during this one's FMA chain
    const double* ybase = Xm + (size_t)n * Np + me * NB + qd * 16;
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = __ldcg(reinterpret_cast<const double2*>(ybase + (size_t)(me * NB) * Np + 2 * e));
    for (int i = me; i < nt; i++) {
        if (i + 1 < nt) {
#pragma unroll
            for (int e = 0; e < 8; e++) w[e] = __ldcg(reinterpret_cast<const double2*>(ybase + (size_t)((i + 1) * NB) * Np + 2 * e));
        }
        const double* zi = z + i * NB + qd * 16;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            su = fma(v[e].x, zi[2 * e], su);
            su = fma(v[e].y, zi[2 * e + 1], su);
        }
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = w[e];
    }
    su += __shfl_xor_sync(0xffffffffu, su, 1);
    su += __shfl_xor_sync(0xffffffffu, su, 2);
    if (qd == 0) a.uv[(size_t)id * Np + me * NB + n] = su;
}

// z = X v for an inverse factor stored as TRANSPOSED tiles (tile (I, J) of the buffer holds X_IJ^T, as this path writes
// it):  z_I[a] = sum_{J <= I} sum_n T_IJ[n][a] v_J[n].  Used for the prior's quadratic forms m^T K^-1 m = ||X_K m||^2
// when q > 1 (quirk Q4 pairs K with a vector that is not its own mean, so the q = 1 identity does not apply).
// grid = (nt, matrices), block = 256: thread = (column a of the tile row, quarter of the 64 n of a tile); the four
// quarters are added in a fixed order.
__global__ void __launch_bounds__(256) mid_trmv_lower_kernel(double* __restrict__ z, const double* __restrict__ XT,
                                                             const double* __restrict__ v, const int* __restrict__ ids,
                                                             int Np) {
    __shared__ double part[4][NB];
    const int id = ids[blockIdx.y], I = blockIdx.x, a = threadIdx.x & 63, qd = threadIdx.x >> 6;
    const double* Xm = XT + (size_t)id * Np * Np;
    const double* vv = v + (size_t)id * Np;
    double s = 0.0;
    for (int J = 0; J <= I; J++) {
        const double* T = Xm + (size_t)(I * NB + qd * 16) * Np + J * NB + a;      // T[n][a], n = 16 qd ..
        const double* vj = vv + J * NB + qd * 16;
#pragma unroll
        for (int n = 0; n < 16; n++) s = fma(__ldcg(T + (size_t)n * Np), vj[n], s);
    }
    part[qd][a] = s;
    __syncthreads();
    if (qd == 0) z[(size_t)id * Np + I * NB + a] = (part[0][a] + part[1][a]) + (part[2][a] + part[3][a]);
}

}  // namespace gprn

// small_v3_v4.cuh -- two designs of the fused small-N kernel that were built, verified bit-identical on the B200 and
// measured SLOWER than the shipped one (gpyrn_b200/csrc/small.cuh); kept out of the library, for the record only
// (DESIGN.md "Tried and rejected", profiles/README.md "fused small-N kernel").  They compile when included after
// small.cuh (they use its SmallArgs / small_tile / slab_* helpers and common.cuh's potrf64_t / mma_slab_ga; they predate the padded scratch layout).
//   version 3: 128-thread CTAs, four per SM, one operand tile in shared memory         C3: 14.8 k evaluations/s
//   version 4: product warps and chain warps on different SM sub-partitions             C3: 13.2 k
//   shipped  : 256-thread CTAs, two per SM                                              C3: 16.3 k
#pragma once
namespace gprn {

// ------------------------------------------------------------------------------------------------
// Version 3 (GPRN_SMALL_V3=1): the same algorithm with FOUR matrices resident per SM instead of two.
// The fused kernel is latency bound -- 64-step pivot / substitution chains, L2 round trips between them -- and what
// fills the FP64 pipe is the number of independent matrices per SM (measured: one 256-thread CTA per SM 11.4 k
// evaluations/s on C3, two 16.2 k).  A 128-thread CTA (four warps = one tile at a time) loses little per matrix,
// because most phases are chains that eight warps do not shorten, and needs ONE operand tile in shared memory:
//   * the A fragments of every product come straight from the L2-resident scratch (mma_slab_ga),
//   * the B tile, L_kk for potrf64 and L_ii / L_kk for the solves share one buffer P,
// so a CTA takes 46 KB and 128 registers per thread: four per SM.
// grid = min(nmat, 4 * SMs) persistent CTAs of 128 threads; dynamic shared memory SMALL_SMEM.
// ------------------------------------------------------------------------------------------------
#define SMALLV3_THREADS 128
#define SMALLV3_CTAS_PER_SM 4
// P + col(128) + pivs(64) + rd(64) + gacc(Np) + zacc(Np) + zpart(4 x 64)
#define SMALLV3_SMEM ((NB * LDT + 4 * NB + 2 * SMALL_MAX_NT * NB + 4 * NB) * sizeof(double))

__global__ void __launch_bounds__(SMALLV3_THREADS, SMALLV3_CTAS_PER_SM) small_pipeline_v3_kernel(SmallArgs a) {
    GPRN_TRACE_SCOPE(TK_SMALL);
    extern __shared__ double smem[];
    double* P = smem;                  // B operand of the products / potrf64 in place / L_kk, L_ii of the solves
    double* col = smem + NB * LDT;
    double* pivs = col + 2 * NB;
    double* rd = pivs + NB;
    double* gacc = rd + NB;            // [Np]
    double* zacc = gacc + SMALL_MAX_NT * NB;
    double* zpart = zacc + SMALL_MAX_NT * NB;      // [4 warps][64]
    __shared__ int bad;
    const int Np = a.Np, nt = Np / NB;
    const int tid = threadIdx.x, w4 = tid >> 5, lane = tid & 31;
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

        // ================= Cholesky, left-looking, one tile at a time =================
        for (int k = 0; k < nt; k++) {
            for (int it = k; it < nt; it++) {
                SMALL_PH(1);
                double acc[2][8][2];
#pragma unroll
                for (int x = 0; x < 2; x++)
#pragma unroll
                    for (int y = 0; y < 8; y++) {
                        const int m = 16 * w4 + 8 * x + r, n = 8 * y + 2 * c;
                        double2 v = *reinterpret_cast<const double2*>(Km + (size_t)(it * NB + m) * Np + k * NB + n);
                        if (dv && it == k) {
                            if (m == n) v.x += dv[k * NB + m];
                            if (m == n + 1) v.y += dv[k * NB + m];
                        }
                        acc[x][y][0] = v.x;
                        acc[x][y][1] = v.y;
                    }
                {   // the next tile's K goes to L2 now: its HBM latency hides behind this round
                    int kn = k, itn = it + 1;
                    if (itn >= nt) { kn = k + 1; itn = kn; }
                    if (kn < nt) {
                        const double* nx = Km + (size_t)(itn * NB + (tid >> 1)) * Np + kn * NB + (tid & 1) * 32;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 16));
                    }
                }
                SMALL_PH(2);
                for (int kp = 0; kp < k; kp++) {
                    load_tile<false>(P, small_tile(sc, k, kp), NB, tid, SMALLV3_THREADS);
                    __syncthreads();
                    mma_slab_ga<true>(acc, small_tile(sc, it, kp), NB, P, w4, lane);
                    __syncthreads();
                }
                SMALL_PH(3);
                if (it == k) {
                    slab_store(acc, P, LDT, w4, lane);
                    __syncthreads();
                    SMALL_PH(4);
                    potrf64_t<4>(P, LDT, P, rd, col, pivs, &bad, tid);
                    SMALL_PH(6);
                    if (tid < 32) logsum += log(pivs[tid]) + log(pivs[tid + 32]);
                    if (tid < NB) rd[tid] = 1.0 / P[tid * LDT + tid];      // the reciprocal every solve of this column uses
                    double* dkk = small_tile(sc, k, k);
                    for (int e = tid; e < NB * (NB / 2); e += SMALLV3_THREADS) {
                        const int m = e >> 5, c2 = e & 31;
                        *reinterpret_cast<double2*>(dkk + m * NB + 2 * c2) = *reinterpret_cast<const double2*>(P + m * LDT + 2 * c2);
                    }
                } else {
                    if (k > 0) {                     // the products went through P: L_kk comes back from the scratch
                        load_tile<false>(P, small_tile(sc, k, k), NB, tid, SMALLV3_THREADS);
                        __syncthreads();
                        if (tid < NB) rd[tid] = 1.0 / P[tid * LDT + tid];
                        __syncthreads();
                    }
                    SMALL_PH(5);
                    trsm_rows_inreg(acc, P, rd, lane);
                    SMALL_PH(6);
                    slab_store(acc, small_tile(sc, it, k), NB, w4, lane);
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
        for (int e = tid; e < Np; e += SMALLV3_THREADS) {
            gacc[e] = 0.0;
            zacc[e] = 0.0;
        }
        __syncthreads();
        for (int i = 0; i < nt; i++) {
            for (int j = 0; j <= i; j++) {
                SMALL_PH(7);
                double acc[2][8][2];
#pragma unroll
                for (int x = 0; x < 2; x++)
#pragma unroll
                    for (int y = 0; y < 8; y++) {
                        const int n = 16 * w4 + 8 * x + r, m = 8 * y + 2 * c;
                        acc[x][y][0] = (j == i && n == m) ? 1.0 : 0.0;
                        acc[x][y][1] = (j == i && n == m + 1) ? 1.0 : 0.0;
                    }
                for (int k = j; k < i; k++) {
                    load_tile<false>(P, small_tile(sc, i, k), NB, tid, SMALLV3_THREADS);              // L_ik   (B operand)
                    __syncthreads();
                    mma_slab_ga<true>(acc, small_tile(sc, k, j), NB, P, w4, lane);                  // - sum Y_kj L_ik^T
                    __syncthreads();
                }
                SMALL_PH(8);
                if (j < i || i == 0) {               // (j == i > 0: P still holds L_ii from the tile before)
                    load_tile<false>(P, small_tile(sc, i, i), NB, tid, SMALLV3_THREADS);              // L_ii
                    __syncthreads();
                    if (tid < NB) rd[tid] = 1.0 / P[tid * LDT + tid];
                    __syncthreads();
                }
                SMALL_PH(9);
                trsm_rows_inreg(acc, P, rd, lane, j == i ? 2 * w4 : 0);
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
                        if (r == 0) zpart[w4 * NB + 8 * y + 2 * c + e] = t;
                    }
                slab_store(acc, small_tile(sc, i, j), NB, w4, lane);
                __syncthreads();
                if (tid < NB) zacc[i * NB + tid] += (zpart[tid] + zpart[NB + tid]) + (zpart[2 * NB + tid] + zpart[3 * NB + tid]);
                __syncthreads();
            }
        }
        SMALL_PH(11);
        // u_j[n] = sum_{i >= j} sum_m Y_ij[n][m] z_i[m] : thread = (row n, half of the row), eight independent 16-byte
        // loads in flight per thread (the tiles come from the L2 scratch); the halves are combined by a shuffle
        {
            const int n = tid >> 1, hf = tid & 1;
            for (int j = 0; j < nt; j++) {
                double su = 0.0;
                for (int i = j; i < nt; i++) {
                    const double* yrow = small_tile(sc, i, j) + n * NB + hf * 32;
                    const double* zi = zacc + i * NB + hf * 32;
#pragma unroll
                    for (int h2 = 0; h2 < 2; h2++) {
                        double2 v[8];
#pragma unroll
                        for (int e = 0; e < 8; e++) v[e] = __ldcg(reinterpret_cast<const double2*>(yrow + 16 * h2 + 2 * e));
#pragma unroll
                        for (int e = 0; e < 8; e++) {
                            su = fma(v[e].x, zi[16 * h2 + 2 * e], su);
                            su = fma(v[e].y, zi[16 * h2 + 2 * e + 1], su);
                        }
                    }
                }
                su += __shfl_xor_sync(0xffffffffu, su, 1);
                if (hf == 0) {
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

// ------------------------------------------------------------------------------------------------
// Version 4: the products and the dependency chains run on DIFFERENT SM sub-partitions.
//
// What limits versions 1-3 (tools/contend_micro.cu, profiles/README.md): a warp that streams DMMAs keeps the FP64 pipe
// of its sub-partition ~90 % busy, and a dependent mul -> shuffle -> fma chain that shares the sub-partition then
// waits behind two or three queued DMMAs at every step -- trsm_rows_inreg takes 6.1 k cycles alone and 24.8 k next to
// a product warp.  With every warp alternating between products and chains, some co-resident warp is always in a
// product phase, so all chains crawl and the pipe idles 60 % of the time.
//
// Here the eight warps of a CTA are split by sub-partition (warp w runs on sub-partition w % 4):
//   * product group  M = warps 2, 3, 6, 7 (sub-partitions 2, 3): forms the right-hand side of every tile,
//         T_ik = A_ik - sum_k' L_ik' L_kk'^T     and     T'_ij = - sum_k Y_kj L_ik^T ,
//     A fragments straight from the L2 scratch (mma_slab_ga), the B tile through its own shared-memory buffer, and
//     hands the tile over through the scratch (it is written where L_ik / Y_ij will live);
//   * chain group    C = warps 0, 1, 4, 5 (sub-partitions 0, 1): potrf64 of the diagonal tiles and the in-register
//     triangular solves of all tiles, the g / z reductions, and publishes the finished tile.
// The groups synchronise on named barriers of 128 threads and hand tiles over through per-tile state flags in shared
// memory (1 = right-hand side ready, 2 = L tile final, 3 = inverse right-hand side ready, 4 = Y tile final); the
// product group runs ahead of the chain group as far as the tile dependencies allow (e.g. the right-hand sides of a
// whole tile column are formed while the column's diagonal tile is being factored).  Two CTAs per SM: the chain
// sub-partitions host four chain warps (never a DMMA stream), the product sub-partitions four product warps.
// Same arithmetic as versions 2 / 3 (bit-identical results).
// ------------------------------------------------------------------------------------------------
#define SMALLV4_THREADS 256
#define SMALLV4_CTAS_PER_SM 2
#define SMALLV4_MAX_CTAS_PER_SM 4            // largest residency of any version: sizes the scratch
// P + Bm + col(128) + pivs(64) + rd(64) + gacc(Np) + zacc(Np) + zpart(4 x 64)
#define SMALLV4_SMEM ((2 * NB * LDT + 4 * NB + 2 * SMALL_MAX_NT * NB + 4 * NB) * sizeof(double))
#define SMALLV4_BAR_C 2
#define SMALLV4_BAR_M 3

// Spin until state[idx] >= want.  Bounded: a scheduling bug shows up as a flagged evaluation, not as a hung GPU.
__device__ __forceinline__ void small_wait(volatile int* state, int idx, int want, volatile int* abort_flag) {
    int spins = 0;
    while (state[idx] < want) {
        __nanosleep(64);
        if (++spins > (1 << 20)) { *abort_flag = 1; break; }
    }
    __threadfence();
}
__device__ __forceinline__ int small_tile_index(int I, int J) { return I * (I + 1) / 2 + J; }

__global__ void __launch_bounds__(SMALLV4_THREADS, SMALLV4_CTAS_PER_SM) small_pipeline_v4_kernel(SmallArgs a) {
    GPRN_TRACE_SCOPE(TK_SMALL);
    extern __shared__ double smem[];
    double* P = smem;                  // chain group: potrf64 in place / L_kk, L_ii of the solves
    double* Bm = smem + NB * LDT;      // product group: B operand of the products
    double* col = smem + 2 * NB * LDT;
    double* pivs = col + 2 * NB;
    double* rd = pivs + NB;
    double* gacc = rd + NB;            // [Np]
    double* zacc = gacc + SMALL_MAX_NT * NB;
    double* zpart = zacc + SMALL_MAX_NT * NB;      // [4 warps][64]
    __shared__ int bad;
    __shared__ volatile int tstate[SMALL_TILES];
    __shared__ volatile int abort_flag;
    const int Np = a.Np, nt = Np / NB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool role_m = (warp & 2) != 0;                       // warps 2, 3, 6, 7
    const int lw = (warp & 1) | ((warp >> 2) << 1);            // warp index inside the group, 0..3
    const int gtid = lw * 32 + lane;
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
        if (tid == 0) { bad = 0; abort_flag = 0; }
        if (tid < SMALL_TILES) tstate[tid] = 0;
        for (int e = tid; e < Np; e += SMALLV4_THREADS) {
            gacc[e] = 0.0;
            zacc[e] = 0.0;
        }
        __syncthreads();

        if (role_m) {
            // ================= product group =================
            for (int k = 0; k < nt; k++) {
                for (int it = k; it < nt; it++) {
                    double acc[2][8][2];
#pragma unroll
                    for (int x = 0; x < 2; x++)
#pragma unroll
                        for (int y = 0; y < 8; y++) {
                            const int m = 16 * lw + 8 * x + r, n = 8 * y + 2 * c;
                            double2 v = *reinterpret_cast<const double2*>(Km + (size_t)(it * NB + m) * Np + k * NB + n);
                            if (dv && it == k) {
                                if (m == n) v.x += dv[k * NB + m];
                                if (m == n + 1) v.y += dv[k * NB + m];
                            }
                            acc[x][y][0] = v.x;
                            acc[x][y][1] = v.y;
                        }
                    {   // the next tile's K goes to L2 now: its HBM latency hides behind this tile
                        int kn = k, itn = it + 1;
                        if (itn >= nt) { kn = k + 1; itn = kn; }
                        if (kn < nt) {
                            const double* nx = Km + (size_t)(itn * NB + (gtid >> 1)) * Np + kn * NB + (gtid & 1) * 32;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 16));
                        }
                    }
                    for (int kp = 0; kp < k; kp++) {
                        small_wait(tstate, small_tile_index(k, kp), 2, &abort_flag);
                        small_wait(tstate, small_tile_index(it, kp), 2, &abort_flag);
                        load_tile<false>(Bm, small_tile(sc, k, kp), NB, gtid, 128);
                        group_sync<SMALLV4_BAR_M, 128>();
                        mma_slab_ga<true>(acc, small_tile(sc, it, kp), NB, Bm, lw, lane);
                        group_sync<SMALLV4_BAR_M, 128>();
                    }
                    slab_store(acc, small_tile(sc, it, k), NB, lw, lane);
                    __threadfence();
                    group_sync<SMALLV4_BAR_M, 128>();
                    if (gtid == 0) tstate[small_tile_index(it, k)] = 1;
                }
            }
            if (a.do_inverse) {
                for (int i = 1; i < nt; i++) {
                    for (int j = 0; j < i; j++) {
                        double acc[2][8][2];
#pragma unroll
                        for (int x = 0; x < 2; x++)
#pragma unroll
                            for (int y = 0; y < 8; y++) acc[x][y][0] = acc[x][y][1] = 0.0;
                        for (int k = j; k < i; k++) {
                            small_wait(tstate, small_tile_index(k, j), 4, &abort_flag);                 // Y_kj final
                            if (k == j) small_wait(tstate, small_tile_index(i, k), 2, &abort_flag);     // (the Cholesky is complete)
                            load_tile<false>(Bm, small_tile(sc, i, k), NB, gtid, 128);                  // L_ik   (B operand)
                            group_sync<SMALLV4_BAR_M, 128>();
                            mma_slab_ga<true>(acc, small_tile(sc, k, j), NB, Bm, lw, lane);             // - sum Y_kj L_ik^T
                            group_sync<SMALLV4_BAR_M, 128>();
                        }
                        slab_store(acc, small_tile(sc, i, j), NB, lw, lane);
                        __threadfence();
                        group_sync<SMALLV4_BAR_M, 128>();
                        if (gtid == 0) tstate[small_tile_index(i, j)] = 3;
                    }
                }
            }
        } else {
            // ================= chain group =================
            double logsum = 0.0;
            for (int k = 0; k < nt; k++) {
                SMALL_PH(1);
                small_wait(tstate, small_tile_index(k, k), 1, &abort_flag);
                load_tile<false>(P, small_tile(sc, k, k), NB, gtid, 128);
                group_sync<SMALLV4_BAR_C, 128>();
                SMALL_PH(4);
                potrf64_t<4, SMALLV4_BAR_C>(P, LDT, P, rd, col, pivs, &bad, gtid);
                SMALL_PH(6);
                if (gtid < 32) logsum += log(pivs[gtid]) + log(pivs[gtid + 32]);
                if (gtid < NB) rd[gtid] = 1.0 / P[gtid * LDT + gtid];      // the reciprocal every solve of this column uses
                double* dkk = small_tile(sc, k, k);
                for (int e = gtid; e < NB * (NB / 2); e += 128) {
                    const int m = e >> 5, c2 = e & 31;
                    *reinterpret_cast<double2*>(dkk + m * NB + 2 * c2) = *reinterpret_cast<const double2*>(P + m * LDT + 2 * c2);
                }
                __threadfence();
                group_sync<SMALLV4_BAR_C, 128>();
                if (gtid == 0) tstate[small_tile_index(k, k)] = 2;
                for (int it = k + 1; it < nt; it++) {
                    SMALL_PH(1);
                    small_wait(tstate, small_tile_index(it, k), 1, &abort_flag);
                    double acc[2][8][2];
                    slab_load_cg(acc, small_tile(sc, it, k), NB, lw, lane);
                    SMALL_PH(5);
                    trsm_rows_inreg(acc, P, rd, lane);
                    SMALL_PH(6);
                    slab_store(acc, small_tile(sc, it, k), NB, lw, lane);
                    __threadfence();
                    group_sync<SMALLV4_BAR_C, 128>();
                    if (gtid == 0) tstate[small_tile_index(it, k)] = 2;
                }
            }
            if (gtid < 32) {
                logsum = warp_sum(logsum);
                if (gtid == 0) {
                    a.logdet[id] = logsum;
                    if (bad) a.mstatus[id] = 1;
                }
            }
            SMALL_PH(0);
            if (a.do_inverse) {
                const double* vglob = a.vv + (size_t)id * Np;
                for (int i = 0; i < nt; i++) {
                    SMALL_PH(8);
                    load_tile<false>(P, small_tile(sc, i, i), NB, gtid, 128);                            // L_ii
                    group_sync<SMALLV4_BAR_C, 128>();
                    if (gtid < NB) rd[gtid] = 1.0 / P[gtid * LDT + gtid];
                    group_sync<SMALLV4_BAR_C, 128>();
                    for (int j = 0; j <= i; j++) {
                        double acc[2][8][2];
                        if (j < i) {
                            SMALL_PH(7);
                            small_wait(tstate, small_tile_index(i, j), 3, &abort_flag);
                            slab_load_cg(acc, small_tile(sc, i, j), NB, lw, lane);
                        } else {
#pragma unroll
                            for (int x = 0; x < 2; x++)
#pragma unroll
                                for (int y = 0; y < 8; y++) {
                                    const int n = 16 * lw + 8 * x + r, m = 8 * y + 2 * c;
                                    acc[x][y][0] = (n == m) ? 1.0 : 0.0;
                                    acc[x][y][1] = (n == m + 1) ? 1.0 : 0.0;
                                }
                        }
                        SMALL_PH(9);
                        trsm_rows_inreg(acc, P, rd, lane, j == i ? 2 * lw : 0);
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
                            const int n = 16 * lw + 8 * x + r;
                            if (c == 0) gacc[j * NB + n] += sg;          // single writer per (j, n) and tile
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
                                if (r == 0) zpart[lw * NB + 8 * y + 2 * c + e] = t;
                            }
                        slab_store(acc, small_tile(sc, i, j), NB, lw, lane);
                        __threadfence();
                        group_sync<SMALLV4_BAR_C, 128>();
                        if (gtid == 0) tstate[small_tile_index(i, j)] = 4;
                        if (gtid < NB) zacc[i * NB + gtid] += (zpart[gtid] + zpart[NB + gtid]) + (zpart[2 * NB + gtid] + zpart[3 * NB + gtid]);
                        group_sync<SMALLV4_BAR_C, 128>();
                    }
                }
            }
            SMALL_PH(11);
        }
        __syncthreads();                   // both groups are through with this matrix: every Y tile is in the scratch
        if (abort_flag && tid == 0) a.mstatus[id] = 2;
        if (a.do_inverse) {
            // u_j[n] = sum_{i >= j} sum_m Y_ij[n][m] z_i[m] : thread = (row n, quarter of the row), eight independent
            // 16-byte loads in flight per thread and tile; the quarters are combined by quad shuffles
            const int n = tid >> 2, qd = tid & 3;
            for (int j = 0; j < nt; j++) {
                double su = 0.0;
                for (int i = j; i < nt; i++) {
                    const double* yrow = small_tile(sc, i, j) + n * NB + qd * 16;
                    const double* zi = zacc + i * NB + qd * 16;
                    double2 v[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) v[e] = __ldcg(reinterpret_cast<const double2*>(yrow + 2 * e));
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

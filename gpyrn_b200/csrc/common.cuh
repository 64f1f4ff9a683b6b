// common.cuh -- shared device helpers for the GPRN hot path (sm_100a, FP64).
//
// Layout conventions (see DESIGN.md "Data layout in HBM"):
//   * every dense matrix is row-major, Np x Np with Np = N rounded up to a multiple of the tile
//     size NB = 64; the padding block is the identity, so Cholesky factors / inverses / log-dets of
//     the padded matrix restrict exactly to those of the N x N matrix;
//   * only the lower triangle (tiles I >= J) is ever read or written, diagonal tiles are stored full;
//   * a batch of matrices is addressed through a device array of matrix ids: matrix `id` lives at
//     base + id * Np * Np, its work vectors at base + id * Np.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NB 64           // tile size
#define LDT 68          // shared-memory row stride (doubles) of an MMA operand tile: 68*8 B = 32 (mod 128)
                        // -> the m8n8k4 fragment loads of a half-warp hit 16 distinct 8-byte slots
#define LDV 68          // substitution vectors V[elem][vec]: stride = 4 (mod 16) doubles, so the DMMA B-fragment loads of
#define LDV2 132        // subst_lower_mma (lane (r,c) -> element 4c.., vector r) hit 16 distinct 8-byte slots per half-warp
#define LDV4 260        // (LDV2 / LDV4: two / four tiles of vectors side by side)
#define TILE_SMEM (NB * LDT * sizeof(double))

namespace gprn {

// D(8x8) += A(8x4) * B(4x8), FP64 tensor-core path (SASS: DMMA.8x8x4).
// a = A[lane/4][lane%4], b = B[lane%4][lane/4], c = {C[lane/4][2*(lane%4)], C[lane/4][2*(lane%4)+1]}.
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// Asynchronous global -> shared copies (LDGSTS): 16 bytes bypassing L1, or 8 bytes (transposing scatters).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ------------------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, SASS: UBLKCP) completing on an mbarrier: one warp moves a 64x64 tile as 64 row
// copies of 512 bytes -- each lane issues two -- instead of every thread of the CTA computing addresses for eight
// 16-byte LDGSTS, and the threads that consume the tile wait on the mbarrier's phase instead of on a CTA barrier.
// The destination keeps the padded row stride LDT (a 1-D bulk copy cannot swizzle; the pad is what keeps the m8n8k4
// fragment loads conflict-free).  Data written with ordinary stores that a bulk copy reads or overwrites (scratch
// tiles in global memory, operand buffers in shared memory) needs fence_proxy_async() between the two proxies.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Wait for the phase with the given parity.  Bounded: a protocol error becomes a flagged result, not a hung GPU.
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned parity) {
    for (int i = 0; i < (1 << 24); i++) {
        unsigned ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
// One warp issues the 64 row copies of a tile (row-major source with leading dimension ld doubles -> dst[64][LDT]).
__device__ __forceinline__ void load_tile_bulk(double* dst, const double* __restrict__ src, size_t ld, int lane,
                                               unsigned long long* bar) {
    bulk_g2s(dst + lane * LDT, src + (size_t)lane * ld, NB * sizeof(double), bar);
    bulk_g2s(dst + (lane + 32) * LDT, src + (size_t)(lane + 32) * ld, NB * sizeof(double), bar);
}

// Load a 64x64 tile from global memory (row-major, leading dimension ld) into shared memory
// dst[64][LDT].  TRANS: dst[m][k] = src[k*ld + m].  All `nthreads` threads of the CTA call.
// The copy is issued as cp.async (every element of the tile is in flight at once: one memory round trip per tile
// instead of one per unrolled batch of register loads) and, with WAIT, completed for the calling thread before
// returning; the caller's __syncthreads() then publishes it.  WAIT = false: the caller batches several tiles and
// ends with cp_async_commit(); cp_async_wait<0>().
template <bool TRANS, bool WAIT = true>
__device__ __forceinline__ void load_tile(double* __restrict__ dst, const double* __restrict__ src, size_t ld,
                                          int tid, int nthreads) {
    if (!TRANS) {
        for (int e = tid; e < NB * (NB / 2); e += nthreads) {
            int r = e >> 5, c2 = e & 31;
            cp_async16(dst + r * LDT + 2 * c2, src + (size_t)r * ld + 2 * c2);
        }
    } else {
        for (int e = tid; e < NB * NB; e += nthreads) {
            int k = e >> 6, m = e & 63;          // coalesced along m in global memory
            cp_async8(dst + m * LDT + k, src + (size_t)k * ld + m);
        }
    }
    if (WAIT) {
        cp_async_commit();
        cp_async_wait<0>();
    }
}

// acc(32x32 warp tile) += As(rows wm*32.., k 0..63) * Bs(rows wn*32.., k 0..63)^T, with As/Bs in
// shared memory [64][LDT].  NEG: subtract instead (a fragment negated).
template <bool NEG>
__device__ __forceinline__ void mma_tile(double (&acc)[4][4][2], const double* __restrict__ As,
                                         const double* __restrict__ Bs, int wm, int wn, int lane) {
    const int r = lane >> 2, c = lane & 3;
    const double* ap = As + (wm * 32 + r) * LDT + c;
    const double* bp = Bs + (wn * 32 + r) * LDT + c;
#pragma unroll 4
    for (int k0 = 0; k0 < NB; k0 += 4) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            a[i] = ap[i * 8 * LDT + k0];
            if (NEG) a[i] = -a[i];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = bp[j * 8 * LDT + k0];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j], a[i], b[j]);
    }
}

// acc += As * diag(sc) * Bs^T: as mma_tile, with the B fragment multiplied by sc[k] (64 doubles in shared memory) as it
// is loaded -- the same rounding as scaling the tile when it is stored, without a register round trip at load time.
__device__ __forceinline__ void mma_tile_scaled(double (&acc)[4][4][2], const double* __restrict__ As,
                                                const double* __restrict__ Bs, const double* __restrict__ sc, int wm,
                                                int wn, int lane) {
    const int r = lane >> 2, c = lane & 3;
    const double* ap = As + (wm * 32 + r) * LDT + c;
    const double* bp = Bs + (wn * 32 + r) * LDT + c;
#pragma unroll 4
    for (int k0 = 0; k0 < NB; k0 += 4) {
        double a[4], b[4];
        const double s = sc[k0 + c];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = ap[i * 8 * LDT + k0];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = bp[j * 8 * LDT + k0] * s;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j], a[i], b[j]);
    }
}

// ------------------------------------------------------------------------------------------------
// Register-resident tile routines: a 64x64 tile is owned by FOUR warps, warp w4 holding the 16 x 64 slab of rows
// 16*w4 .. 16*w4+15 in DMMA accumulator layout,
//     acc[x][y][e] = T[16*w4 + 8x + r][8y + 2c + e],     r = lane / 4, c = lane % 4.
// Every warp then owns COMPLETE rows, so a triangular solve from the right (X L^T = T: the Cholesky panel solve and,
// with the inverse factor stored transposed, the row sweep of the inverse) runs inside the warp's registers: no
// staging of the right-hand sides in shared memory, no transposition, no CTA barrier.  oracle/warp_model.py is a
// lane-level numpy model of these routines (tests/test_warp_model.py checks the index arithmetic on the CPU).
// ------------------------------------------------------------------------------------------------

// acc(16 x 64 slab of warp w4) += As[slab rows][k] * Bs[n][k]^T  (As, Bs: 64 x 64 in shared memory, stride LDT).
template <bool NEG>
__device__ __forceinline__ void mma_slab(double (&acc)[2][8][2], const double* __restrict__ As,
                                         const double* __restrict__ Bs, int w4, int lane) {
    const int r = lane >> 2, c = lane & 3;
    const double* ap = As + (w4 * 16 + r) * LDT + c;
    const double* bp = Bs + r * LDT + c;
#pragma unroll 2
    for (int k0 = 0; k0 < NB; k0 += 4) {
        double a[2], b[8];
#pragma unroll
        for (int x = 0; x < 2; x++) {
            a[x] = ap[x * 8 * LDT + k0];
            if (NEG) a[x] = -a[x];
        }
#pragma unroll
        for (int y = 0; y < 8; y++) b[y] = bp[y * 8 * LDT + k0];
#pragma unroll
        for (int x = 0; x < 2; x++)
#pragma unroll
            for (int y = 0; y < 8; y++) dmma884(acc[x][y], a[x], b[y]);
    }
}

// As mma_slab, with the A fragments read STRAIGHT FROM GLOBAL MEMORY (Ag: the 64 x 64 A tile, row-major, leading
// dimension lda; it sits in the L2-resident scratch): a lane's fragment element A[row][4 ks + c] is an 8-byte load, a
// quad reads one 32-byte sector, and every element is used by exactly one lane -- staging the A tile in shared memory
// would only add a copy.  The loads run a chunk of four k-steps (64 DMMAs, ~1000 cycles) ahead of their use, which
// covers the L2 latency.  Only the B tile (read by all four warps) needs shared memory.
template <bool NEG>
__device__ __forceinline__ void mma_slab_ga(double (&acc)[2][8][2], const double* __restrict__ Ag, int lda,
                                            const double* __restrict__ Bs, int w4, int lane) {
    const int r = lane >> 2, c = lane & 3;
    const double* ap = Ag + (size_t)(w4 * 16 + r) * lda + c;
    const double* bp = Bs + r * LDT + c;
    double an[4][2];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int x = 0; x < 2; x++) an[u][x] = __ldcg(ap + (size_t)x * 8 * lda + 4 * u);
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

// Solve  X L^T = T  in place for the warp's slab (forward substitution along the columns, right-looking over the
// eight 8-column blocks).  Ls: L, 64 x 64 lower triangular in shared memory (stride LDT); rd[j] = 1 / L[j][j].
// Block step y:  (a) the 8 x 8 diagonal block is solved inside each quad (the four lanes of a quad hold one row's
// eight values of the block): column j is finished by its owner (multiplication by the reciprocal diagonal), broadcast
// to the quad by one shuffle and eliminated from the columns j' > j by the lanes that own them -- scalar FMAs in
// ascending column order, exactly the arithmetic of subst_lower_mma; (b) the solved block (negated, turned into A
// fragments by two more quad shuffles per k-step) updates the blocks y' > y,  T[:, y'] -= X[:, y] L[y', y]^T : two
// DMMA k-steps per (row block, y'), B fragments straight from Ls -- block y + 1 at once, the others interleaved with
// the column steps of the next solve.  A genuine substitution: no inverse of a block is
// formed.  ymin: column blocks below ymin are zero in T for all rows of the warp (identity right-hand sides).
// (A variant that gathers the eight values into every lane and solves the block redundantly halves the shuffle
// latency but needs ~170 registers; this one takes 102.)
__device__ __forceinline__ void trsm_rows_inreg(double (&acc)[2][8][2], const double* __restrict__ Ls,
                                                const double* __restrict__ rd, int lane, int ymin = 0) {
    const int r = lane >> 2, c = lane & 3, qbase = lane & ~3;
    double pa0[2] = {0.0, 0.0}, pa1[2] = {0.0, 0.0};     // A fragments of the previous block (its deferred updates)
#pragma unroll
    for (int y = 0; y < 8; y++) {
        if (y < ymin) continue;                       // warp-uniform
        const double* Ld = Ls + (8 * y + 2 * c) * LDT + 8 * y;     // this lane's two rows of the diagonal block of L
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double rdj = rd[8 * y + j];
            const double l0 = Ld[j], l1 = Ld[LDT + j];
#pragma unroll
            for (int x = 0; x < 2; x++) {
                double xj = ((j & 1) ? acc[x][y][1] : acc[x][y][0]) * rdj;
                xj = __shfl_sync(0xffffffffu, xj, qbase | (j >> 1));
                if (c == (j >> 1)) {
                    if (j & 1) acc[x][y][1] = xj;
                    else acc[x][y][0] = xj;
                }
                if (2 * c > j) acc[x][y][0] = fma(-l0, xj, acc[x][y][0]);
                if (2 * c + 1 > j) acc[x][y][1] = fma(-l1, xj, acc[x][y][1]);
            }
            // software pipelining: the updates that block y-1 owes to the blocks beyond y are issued between the
            // column steps of block y, whose mul -> shuffle -> fma chain (46 cycles a step) leaves the FP64 pipe idle;
            // only the update of block y itself was applied before this solve started.  Every target still receives
            // its updates in ascending source order.
            if (y >= 1 && y + 1 + j <= 7 && y > ymin) {
                const int yy = y + 1 + j;
                const double b0 = Ls[(8 * yy + r) * LDT + 8 * (y - 1) + c], b1 = Ls[(8 * yy + r) * LDT + 8 * (y - 1) + 4 + c];
#pragma unroll
                for (int x = 0; x < 2; x++) {
                    dmma884(acc[x][yy], pa0[x], b0);
                    dmma884(acc[x][yy], pa1[x], b1);
                }
            }
        }
        if (y == 7) break;
        // A fragments of the solved block: lane (r, c) needs columns c and 4 + c of row r, held by lanes c/2 and 2 + c/2
#pragma unroll
        for (int x = 0; x < 2; x++) {
            const double v0 = __shfl_sync(0xffffffffu, acc[x][y][0], qbase | (c >> 1));
            const double v1 = __shfl_sync(0xffffffffu, acc[x][y][1], qbase | (c >> 1));
            const double w0 = __shfl_sync(0xffffffffu, acc[x][y][0], qbase | 2 | (c >> 1));
            const double w1 = __shfl_sync(0xffffffffu, acc[x][y][1], qbase | 2 | (c >> 1));
            pa0[x] = -((c & 1) ? v1 : v0);
            pa1[x] = -((c & 1) ? w1 : w0);
        }
        {   // the next block needs this one's update now
            const int yy = y + 1;
            const double b0 = Ls[(8 * yy + r) * LDT + 8 * y + c], b1 = Ls[(8 * yy + r) * LDT + 8 * y + 4 + c];
#pragma unroll
            for (int x = 0; x < 2; x++) {
                dmma884(acc[x][yy], pa0[x], b0);
                dmma884(acc[x][yy], pa1[x], b1);
            }
        }
    }
}

// Forward substitution  L y = g  for ONE vector per thread, everything in shared memory.
//   Ls : 64x64 lower-triangular tile, element (m,k) at Ls[m*lds + k]   (read by broadcast)
//   rd : reciprocals of the diagonal of Ls (64 values)
//   V  : vectors, element m of vector n at V[m*ldv + n]
//   first_block : rows below first_block*8 are known to be zero on input (identity right-hand sides)
// Blocked by 8 so that the 8 running values stay in registers.  This is a genuine substitution (LAPACK
// dtrsm semantics) -- multiplying by an explicit inverse of the tile instead costs 1-2 digits of ELBO
// parity on these ill-conditioned matrices (DESIGN.md "Numerics"); only the 64 scalar divisions by the
// diagonal are replaced by multiplications with its reciprocal (<= 1 ulp each, off the critical path).
__device__ __forceinline__ void subst_lower(const double* __restrict__ Ls, int lds, const double* __restrict__ rd,
                                            double* __restrict__ V, int ldv, int n, int first_block = 0) {
    for (int mb = first_block; mb < 8; mb++) {
        double y[8];
#pragma unroll
        for (int u = 0; u < 8; u++) y[u] = V[(mb * 8 + u) * ldv + n];
        for (int kb = first_block; kb < mb; kb++) {
            double x[8];
#pragma unroll
            for (int w = 0; w < 8; w++) x[w] = V[(kb * 8 + w) * ldv + n];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const double* lrow = Ls + (mb * 8 + u) * lds + kb * 8;
#pragma unroll
                for (int w = 0; w < 8; w++) y[u] = fma(-lrow[w], x[w], y[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const double* lrow = Ls + (mb * 8 + u) * lds + mb * 8;
#pragma unroll
            for (int w = 0; w < u; w++) y[u] = fma(-lrow[w], y[w], y[u]);
            y[u] = y[u] * rd[mb * 8 + u];
        }
#pragma unroll
        for (int u = 0; u < 8; u++) V[(mb * 8 + u) * ldv + n] = y[u];
    }
}

// Warp-cooperative forward substitution  L y = g  for the 8*NT vectors vec0 .. vec0+8*NT-1 (element m of vector n at
// V[m*ldv + n], ldv = 4 mod 16), called by all 32 lanes of a warp; different warps take different vector ranges and
// never synchronise with each other.  Blocked by 8 like subst_lower, but the off-diagonal part
//     Y[8 rows of block mb][vectors] -= L[mb][kb] (8x8) * X[kb][vectors]
// runs on DMMA m8n8k4 (2 k-steps per 8x8 block and 8-vector tile; accumulator = the warp's 8 x 8NT slab of V), and
// only the 8x8 diagonal blocks are solved by scalar FMA chains, one vector per lane.  Still a genuine substitution:
// no inverse of L or of its diagonal blocks is formed.  first_block as in subst_lower (all vectors of the warp).
template <int NT>
__device__ __forceinline__ void subst_lower_mma(const double* __restrict__ Ls, int lds, const double* __restrict__ rd,
                                                double* V, int ldv, int vec0, int first_block = 0) {
    const int lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
    constexpr int NV = (8 * NT + 31) / 32;          // vectors per lane in the diagonal solves
    for (int mb = first_block; mb < 8; mb++) {
        double acc[NT][2];
        double* vrow = V + (mb * 8 + r) * ldv + vec0 + 2 * c;
#pragma unroll
        for (int t = 0; t < NT; t++) {
            acc[t][0] = vrow[8 * t];
            acc[t][1] = vrow[8 * t + 1];
        }
        const double* lrow = Ls + (mb * 8 + r) * lds + c;
        for (int kb = first_block; kb < mb; kb++) {
            const double a0 = -lrow[kb * 8], a1 = -lrow[kb * 8 + 4];
            const double* xb = V + (kb * 8 + c) * ldv + vec0 + r;
#pragma unroll
            for (int t = 0; t < NT; t++) {
                dmma884(acc[t], a0, xb[8 * t]);
                dmma884(acc[t], a1, xb[4 * ldv + 8 * t]);
            }
        }
#pragma unroll
        for (int t = 0; t < NT; t++) {
            vrow[8 * t] = acc[t][0];
            vrow[8 * t + 1] = acc[t][1];
        }
        __syncwarp();
        double y[NV][8];
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int v = lane + 32 * j;
            if (v < 8 * NT) {
#pragma unroll
                for (int u = 0; u < 8; u++) y[j][u] = V[(mb * 8 + u) * ldv + vec0 + v];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const double* ld = Ls + (mb * 8 + u) * lds + mb * 8;
            const double rdu = rd[mb * 8 + u];
#pragma unroll
            for (int j = 0; j < NV; j++) {
#pragma unroll
                for (int w = 0; w < u; w++) y[j][u] = fma(-ld[w], y[j][w], y[j][u]);
                y[j][u] = y[j][u] * rdu;
            }
        }
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int v = lane + 32 * j;
            if (v < 8 * NT) {
#pragma unroll
                for (int u = 0; u < 8; u++) V[(mb * 8 + u) * ldv + vec0 + v] = y[j][u];
            }
        }
        __syncwarp();
    }
}

// Cholesky of a 64x64 tile by 256 threads.  Two implementations with the same contract:
//   Td  : input tile in shared memory, element (m,n) at Td[m*ldd + n] (lower triangle used)
//   Ls  : output, lower-triangular L with zeros above the diagonal, stride LDT (may alias Td)
//   rd  : output, 1 / L[c][c];   col: scratch 128 doubles (v1 only);   pivs: output, pivots L[c][c]^2
//   bad : set to 1 when a pivot is not positive (caller zeroes it)
//
// potrf64 ("v2" arithmetic; the default is its look-ahead form "v3" below): right-looking in 8 block steps of 8
// columns -- 16 CTA barriers instead of 64.
//   Phase A (warps 0-1, thread t = row t): every thread loads the 8x8 diagonal block (broadcast reads) and factors
//     it redundantly in registers on UNSCALED columns (T[i][k] = L[i][k] L[k][k], pivot p_k = L[k][k]^2: one
//     reciprocal per pivot on the dependency chain, no square root), then solves its own row of the block column
//     against it by forward substitution and writes the (unscaled) row.
//   Phase B (all 8 warps): the trailing 8x8 blocks (I >= K > j) get C_IK -= T_Ij diag(1/p) T_Kj^T on DMMA m8n8k4.
//   The columns are scaled by 1/sqrt(p) once at the end.
// A genuine substitution / factorisation: no inverse of a block is formed (DESIGN.md "Numerics").
// Measured against v1 (one barrier per pivot, 64 x 290 ns = 19 us per tile): see profiles/README.md.
#ifndef GPRN_POTRF_V1
// 1/x for a positive normal x: hardware seed (MUFU.RCP64H, relative error <= 2^-23) + two Newton steps
// (2^-23 -> 2^-46 -> below 1 ulp).  Shorter dependency chain than the IEEE division / __drcp_rn sequence, which is
// what a pivot step waits on.  Non-positive, NaN or subnormal inputs give garbage; the caller's pivot test flags them.
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// potrf64 (default, "v3"): the arithmetic of v2, operation for operation (bit-identical factors), with the
// dependency chain shortened:
//   * look-ahead: after block step j only the NEXT block column (I, j+1) is updated by all warps; warps 0-1 then start
//     phase A of step j+1 while warps 2-7 apply step j to the remaining blocks (I >= K >= j+2) -- the trailing update
//     leaves the critical path (every block still receives its updates in ascending j);
//   * phase A eliminates the thread's own row together with the 8x8 diagonal block (column by column: independent
//     FMAs) instead of solving it afterwards with one serial chain per element.
// Measured in isolation (tools/small_micro.cu): see profiles/README.md.
__device__ __forceinline__ void potrf64_phase_a(double* Ls, double* __restrict__ sinv, double* __restrict__ pivs,
                                                int* bad, int j, int tid) {
    const int c0 = 8 * j;
    double T[36], a[8];
#define GPRN_PT(u, w) T[(u) * ((u) + 1) / 2 + (w)]
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
        for (int w = 0; w <= u; w++) GPRN_PT(u, w) = Ls[(c0 + u) * LDT + c0 + w];
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const double2 v = *reinterpret_cast<const double2*>(Ls + tid * LDT + c0 + k);
        a[k] = v.x;
        a[k + 1] = v.y;
    }
    // every read of the diagonal block precedes every write of this phase (rows c0..c0+7 are rewritten)
    asm volatile("bar.sync 1, 64;" ::: "memory");
    // 8x8 diagonal block, redundantly in every thread (registers only; the warp pays for one thread), and the
    // thread's own row of the block column with it:  xs_k = a_k - sum_{c<k} (xs_c / p_c) T[k][c]   (unscaled, like T).
    // For a row of the diagonal block itself this reproduces T[u][k] (k <= u) with the same operations.
    double inv[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        inv[k] = rcp_fast(GPRN_PT(k, k));
#pragma unroll
        for (int i = k + 1; i < 8; i++) {
            const double t = GPRN_PT(i, k) * inv[k];
#pragma unroll
            for (int c = k + 1; c <= i; c++) GPRN_PT(i, c) = fma(-t, GPRN_PT(c, k), GPRN_PT(i, c));
        }
        const double yk = a[k] * inv[k];
#pragma unroll
        for (int c = k + 1; c < 8; c++) a[c] = fma(-yk, GPRN_PT(c, k), a[c]);
    }
    if (tid >= c0) {
        const int u = tid - c0;    // 0..7: row of the diagonal block; >= 8: row below it
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k == u) {
                const double p = GPRN_PT(k, k);
                pivs[tid] = p;
                sinv[tid] = inv[k];
                if (!(p > 0.0)) *bad = 1;
            }
            Ls[tid * LDT + c0 + k] = a[k];
        }
    }
#undef GPRN_PT
}

// C_IK -= T_Ij diag(1/p) T_Kj^T for one 8x8 block (I >= K > j), by one warp, on DMMA m8n8k4
__device__ __forceinline__ void potrf64_update_block(double* Ls, const double* __restrict__ sinv, int I, int K, int j,
                                                     int lane) {
    const int r = lane >> 2, c = lane & 3, c0 = 8 * j;
    const double s0 = -sinv[c0 + c], s1 = -sinv[c0 + c + 4];
    double* cp = Ls + (8 * I + r) * LDT + 8 * K + 2 * c;
    double2 cv = *reinterpret_cast<double2*>(cp);
    double acc[2] = {cv.x, cv.y};
    const double* ap = Ls + (8 * I + r) * LDT + c0 + c;
    const double* bp = Ls + (8 * K + r) * LDT + c0 + c;
    dmma884(acc, ap[0] * s0, bp[0]);
    dmma884(acc, ap[4] * s1, bp[4]);
    *reinterpret_cast<double2*>(cp) = make_double2(acc[0], acc[1]);
}

// NW: warps of the calling CTA (8: the 256-thread kernels; 4: the 128-thread fused small-N kernel).  Warps 0-1 run
// phase A, warps 2..NW-1 the rest of the trailing update next to it; the factors do not depend on NW.
// BAR / tid: the NW warps may be a GROUP of a larger CTA (the fused small-N kernel runs the factorisation on four of
// its eight warps while the others stream DMMA products): tid is the thread's index inside the group and BAR the
// named barrier the group synchronises on (0: the whole CTA, __syncthreads).
template <int BAR, int NTHR>
__device__ __forceinline__ void group_sync() {
    if (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NTHR) : "memory");
}
template <int NW, int BAR = 0>
__device__ __forceinline__ void potrf64_t(const double* Td, int ldd, double* Ls, double* __restrict__ rd,
                                          double* __restrict__ col, double* __restrict__ pivs, int* bad, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    double* sinv = col;                    // 1 / pivot of the finished columns (64 doubles of the scratch)
    if (Td != Ls || ldd != LDT) {          // bring the tile into Ls with stride LDT (alias-safe)
        static_assert(NW == 8 || NW == 4, "potrf64_t: 4 or 8 warps");
        constexpr int PER = 16 * 8 / NW;   // elements per thread
        const int r = tid / (NW / 2), q4 = tid % (NW / 2);
        double a[PER];
#pragma unroll
        for (int u = 0; u < PER; u++) a[u] = Td[r * ldd + q4 + (NW / 2) * u];
        group_sync<BAR, 32 * NW>();
#pragma unroll
        for (int u = 0; u < PER; u++) Ls[r * LDT + q4 + (NW / 2) * u] = a[u];
        group_sync<BAR, 32 * NW>();
    }
    if (tid < NB) potrf64_phase_a(Ls, sinv, pivs, bad, 0, tid);
    group_sync<BAR, 32 * NW>();
    for (int j = 0; j < 7; j++) {
        // look-ahead: block column j+1 first (7 - j blocks) ...
        for (int b = warp; b < 7 - j; b += NW) potrf64_update_block(Ls, sinv, j + 1 + b, j + 1, j, lane);
        group_sync<BAR, 32 * NW>();
        if (tid < NB) {
            potrf64_phase_a(Ls, sinv, pivs, bad, j + 1, tid);                  // ... so that the next step can start
        } else {
            // the other warps: the rest of step j, blocks (I, K) with I >= K >= j + 2
            // (restricting this to the warps that do not share an SM sub-partition with a phase-A warp -- 2, 3, 6, 7 of
            // eight -- was measured slower: 14.07 k instead of 13.38 k cycles; four warps make it the critical path)
            const int nb = 6 - j, cnt = nb * (nb + 1) / 2;
            for (int pi = warp - 2; pi < cnt; pi += NW - 2) {
                int ii = 0, kk = pi;
                while (kk > ii) { kk -= ii + 1; ii++; }
                potrf64_update_block(Ls, sinv, j + 2 + ii, j + 2 + kk, j, lane);
            }
        }
        group_sync<BAR, 32 * NW>();
    }
    // scale the columns: L[r][c] = T[r][c] / sqrt(p_c), L[c][c] = sqrt(p_c), zeros above the diagonal
    if (tid < NB) rd[tid] = rcp_fast(sqrt(pivs[tid]));
    group_sync<BAR, 32 * NW>();
    for (int e = tid; e < NB * 4; e += 32 * NW) {
        const int r = e >> 2, q4 = e & 3;
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int cc = q4 + 4 * u;
            double v = 0.0;
            if (cc < r) v = Ls[r * LDT + cc] * rd[cc];
            else if (cc == r) v = sqrt(pivs[r]);
            Ls[r * LDT + cc] = v;
        }
    }
    group_sync<BAR, 32 * NW>();
}
__device__ __forceinline__ void potrf64(const double* Td, int ldd, double* Ls, double* __restrict__ rd,
                                        double* __restrict__ col, double* __restrict__ pivs, int* bad) {
    potrf64_t<8>(Td, ldd, Ls, rd, col, pivs, bad, threadIdx.x);
}

#else
// potrf64 v1 (-DGPRN_POTRF_V1): tile held in registers (thread (r, q4) owns row r, columns q4 + 4u), outer-product
// form on UNSCALED columns, one barrier per pivot: the owners of column c+1 publish it to shared memory as soon as
// step c has updated it.  Scaling by 1/sqrt(pivot) happens once at the end.
__device__ __forceinline__ void potrf64(const double* Td, int ldd, double* Ls, double* __restrict__ rd,
                                        double* __restrict__ col, double* __restrict__ pivs, int* bad) {
    const int tid = threadIdx.x, r = tid >> 2, q4 = tid & 3;
    double a[16];
#pragma unroll
    for (int u = 0; u < 16; u++) a[u] = Td[r * ldd + q4 + 4 * u];
    __syncthreads();                       // Td fully read: Ls (which may alias it) can be written from here on
    // Finished (unscaled) columns are parked in Ls the moment they are published, so the update below needs no
    // "column already final" / "above the diagonal" predicates: registers of dead elements may hold garbage, it is
    // never read.  This halves the instructions of a pivot step (the step is issue- as much as latency-bound).
    if (q4 == 0) { col[r] = a[0]; Ls[r * LDT] = a[0]; }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NB; c++) {
        const double* cb = col + (c & 1) * NB;
        const double piv = cb[c];
        if (tid == 0) {
            pivs[c] = piv;
            if (!(piv > 0.0)) *bad = 1;
        }
        if (r > c) {
            const double t = cb[r] * (1.0 / piv);
#pragma unroll
            for (int u = c >> 2; u < 16; u++) a[u] = fma(-t, cb[q4 + 4 * u], a[u]);
        }
        if (c + 1 < NB) {
            if (q4 == ((c + 1) & 3)) {
                const double v = a[(c + 1) >> 2];
                col[((c + 1) & 1) * NB + r] = v;
                Ls[r * LDT + c + 1] = v;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (tid < NB) rd[tid] = 1.0 / sqrt(pivs[tid]);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 16; u++) {
        const int cc = q4 + 4 * u;
        double v = 0.0;
        if (cc < r) v = Ls[r * LDT + cc] * rd[cc];
        else if (cc == r) v = sqrt(pivs[r]);
        Ls[r * LDT + cc] = v;
    }
    __syncthreads();
}
#endif

// ------------------------------------------------------------------------------------------------
// Per-CTA execution trace (development aid, compiled in only with -DGPRN_TRACE; see tools/trace_run.py):
// every CTA of the traced kernels appends (kernel id, SM id, start, end in globaltimer ns) to a device buffer.
// ------------------------------------------------------------------------------------------------
#ifdef GPRN_TRACE
struct TraceRec { unsigned long long t0, t1; int kid, smid; };
__device__ TraceRec* g_trace_buf = nullptr;
__device__ unsigned int g_trace_n = 0, g_trace_cap = 0;
struct TraceScope {
    unsigned long long t0;
    int kid;
    __device__ __forceinline__ TraceScope(int k) : t0(0), kid(k) {
        if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    }
    __device__ __forceinline__ ~TraceScope() {
        if (threadIdx.x == 0 && g_trace_buf) {
            unsigned long long t1;
            unsigned smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            unsigned i = atomicAdd(&g_trace_n, 1u);
            if (i < g_trace_cap) g_trace_buf[i] = TraceRec{t0, t1, kid, (int)smid};
        }
    }
};
#define GPRN_TRACE_SCOPE(k) TraceScope trace_scope_(k)
#else
#define GPRN_TRACE_SCOPE(k)
#endif
enum { TK_PANEL = 1, TK_POTRF = 2, TK_TRSM = 3, TK_SYRK64 = 4, TK_SYRK_OUTER = 5, TK_TRTRI_DIAG = 6, TK_TRTRI_ROW = 7,
       TK_TRTRI_OUTER = 8, TK_TRTRI_INBLOCK = 9, TK_TRMV_LOWER = 10, TK_TRMV_UPPER = 11, TK_CROSS_FROB = 12,
       TK_FORM_A = 13, TK_KASSEMBLE = 14, TK_SMALL = 15, TK_OTHER = 16 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum (blockDim.x <= 1024, multiple of 32); result valid in every thread.
__device__ __forceinline__ double block_sum(double v, double* red /* >= 33 doubles of shared memory */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// linear index over the lower triangle (including diagonal) of an n x n tile grid -> (I, J), I >= J
__device__ __forceinline__ void tri_decode(int t, int& I, int& J) {
    int i = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= t) i++;
    while (i * (i + 1) / 2 > t) i--;
    I = i;
    J = t - i * (i + 1) / 2;
}

}  // namespace gprn

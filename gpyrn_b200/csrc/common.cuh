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
// potrf64 (default, "v2"): right-looking in 8 block steps of 8 columns -- 16 CTA barriers instead of 64.
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

__device__ __forceinline__ void potrf64(const double* Td, int ldd, double* Ls, double* __restrict__ rd,
                                        double* __restrict__ col, double* __restrict__ pivs, int* bad) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double* sinv = col;                    // 1 / pivot of the finished columns (64 doubles of the scratch)
    if (Td != Ls || ldd != LDT) {          // bring the tile into Ls with stride LDT (alias-safe)
        const int r = tid >> 2, q4 = tid & 3;
        double a[16];
#pragma unroll
        for (int u = 0; u < 16; u++) a[u] = Td[r * ldd + q4 + 4 * u];
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 16; u++) Ls[r * LDT + q4 + 4 * u] = a[u];
        __syncthreads();
    }
    // The tile is factored on UNSCALED columns, T[i][k] = L[i][k] * L[k][k] with pivots p_k = L[k][k]^2 (as in v1):
    // a pivot step then needs one reciprocal and no square root; the columns are scaled once at the end.
#define GPRN_PT(u, w) T[(u) * ((u) + 1) / 2 + (w)]
    for (int j = 0; j < 8; j++) {
        const int c0 = 8 * j;
        if (tid < NB) {                    // warps 0-1, converged: thread = row of the tile
            double T[36], inv[8], a[8], y[8];
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
            // 8x8 diagonal block, redundantly in every thread (registers only; the warp pays for one thread)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                inv[k] = rcp_fast(GPRN_PT(k, k));
#pragma unroll
                for (int i = k + 1; i < 8; i++) {
                    const double t = GPRN_PT(i, k) * inv[k];
#pragma unroll
                    for (int c = k + 1; c <= i; c++) GPRN_PT(i, c) = fma(-t, GPRN_PT(c, k), GPRN_PT(i, c));
                }
            }
            // own row of the block column:  xs_k = a_k - sum_{c<k} (xs_c / p_c) T[k][c]   (unscaled, like T).
            // For a row of the diagonal block itself this reproduces T[u][k] (k <= u) with the same operations.
#pragma unroll
            for (int k = 0; k < 8; k++) {
#pragma unroll
                for (int c = 0; c < k; c++) a[k] = fma(-y[c], GPRN_PT(k, c), a[k]);
                y[k] = a[k] * inv[k];
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
        }
        __syncthreads();
        if (j < 7) {
            // trailing blocks (I >= K > j):  C_IK -= T_Ij diag(1/p) T_Kj^T  on DMMA m8n8k4
            const int nb = 7 - j, cnt = nb * (nb + 1) / 2;
            const int r = lane >> 2, c = lane & 3;
            const double s0 = -sinv[c0 + c], s1 = -sinv[c0 + c + 4];
            for (int pi = warp; pi < cnt; pi += 8) {
                int ii = 0, kk = pi;
                while (kk > ii) { kk -= ii + 1; ii++; }
                const int I = j + 1 + ii, K = j + 1 + kk;
                double* cp = Ls + (8 * I + r) * LDT + 8 * K + 2 * c;
                double2 cv = *reinterpret_cast<double2*>(cp);
                double acc[2] = {cv.x, cv.y};
                const double* ap = Ls + (8 * I + r) * LDT + c0 + c;
                const double* bp = Ls + (8 * K + r) * LDT + c0 + c;
                dmma884(acc, ap[0] * s0, bp[0]);
                dmma884(acc, ap[4] * s1, bp[4]);
                *reinterpret_cast<double2*>(cp) = make_double2(acc[0], acc[1]);
            }
            __syncthreads();
        }
    }
#undef GPRN_PT
    // scale the columns: L[r][c] = T[r][c] / sqrt(p_c), L[c][c] = sqrt(p_c), zeros above the diagonal
    if (tid < NB) rd[tid] = rcp_fast(sqrt(pivs[tid]));
    __syncthreads();
    {
        const int r = tid >> 2, q4 = tid & 3;
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int cc = q4 + 4 * u;
            double v = 0.0;
            if (cc < r) v = Ls[r * LDT + cc] * rd[cc];
            else if (cc == r) v = sqrt(pivs[r]);
            Ls[r * LDT + cc] = v;
        }
    }
    __syncthreads();
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

// small_micro.cu -- cycle counts of the building blocks of the fused small-N kernel in isolation (one CTA on one SM):
// dependent-issue latencies of DFMA / DMUL / SHFL.64 / DMMA, and potrf64 / trsm_rows_inreg / mma_slab per call.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -o tools/small_micro tools/small_micro.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "../gpyrn_b200/csrc/common.cuh"
using namespace gprn;

namespace gprn {
// potrf64_v2: the previous version (trailing update of a block step completed before the next step starts); kept for
// tools/small_micro.cu, which times it and bit-compares it against the look-ahead version.
__device__ __forceinline__ void potrf64_v2(const double* Td, int ldd, double* Ls, double* __restrict__ rd,
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
}  // namespace gprn


// the un-pipelined form of trsm_rows_inreg (all updates of a solved block issued before the next solve starts)
__device__ __forceinline__ void trsm_rows_inreg_v0(double (&acc)[2][8][2], const double* __restrict__ Ls,
                                                   const double* __restrict__ rd, int lane, int ymin = 0) {
    const int r = lane >> 2, c = lane & 3, qbase = lane & ~3;
#pragma unroll
    for (int y = 0; y < 8; y++) {
        if (y < ymin) continue;
        const double* Ld = Ls + (8 * y + 2 * c) * LDT + 8 * y;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double rdj = rd[8 * y + j];
            const double l0 = Ld[j], l1 = Ld[LDT + j];
#pragma unroll
            for (int x = 0; x < 2; x++) {
                double xj = ((j & 1) ? acc[x][y][1] : acc[x][y][0]) * rdj;
                xj = __shfl_sync(0xffffffffu, xj, qbase | (j >> 1));
                if (c == (j >> 1)) { if (j & 1) acc[x][y][1] = xj; else acc[x][y][0] = xj; }
                if (2 * c > j) acc[x][y][0] = fma(-l0, xj, acc[x][y][0]);
                if (2 * c + 1 > j) acc[x][y][1] = fma(-l1, xj, acc[x][y][1]);
            }
        }
        if (y == 7) break;
        double a0[2], a1[2];
#pragma unroll
        for (int x = 0; x < 2; x++) {
            const double v0 = __shfl_sync(0xffffffffu, acc[x][y][0], qbase | (c >> 1));
            const double v1 = __shfl_sync(0xffffffffu, acc[x][y][1], qbase | (c >> 1));
            const double w0 = __shfl_sync(0xffffffffu, acc[x][y][0], qbase | 2 | (c >> 1));
            const double w1 = __shfl_sync(0xffffffffu, acc[x][y][1], qbase | 2 | (c >> 1));
            a0[x] = -((c & 1) ? v1 : v0);
            a1[x] = -((c & 1) ? w1 : w0);
        }
#pragma unroll
        for (int yy = y + 1; yy < 8; yy++) {
            const double b0 = Ls[(8 * yy + r) * LDT + 8 * y + c], b1 = Ls[(8 * yy + r) * LDT + 8 * y + 4 + c];
#pragma unroll
            for (int x = 0; x < 2; x++) { dmma884(acc[x][yy], a0[x], b0); dmma884(acc[x][yy], a1[x], b1); }
        }
    }
}
template <int UNR>
__device__ __forceinline__ void mma_slab_u(double (&acc)[2][8][2], const double* __restrict__ As,
                                           const double* __restrict__ Bs, int w4, int lane) {
    const int r = lane >> 2, c = lane & 3;
    const double* ap = As + (w4 * 16 + r) * LDT + c;
    const double* bp = Bs + r * LDT + c;
#pragma unroll UNR
    for (int k0 = 0; k0 < NB; k0 += 4) {
        double a[2], b[8];
#pragma unroll
        for (int x = 0; x < 2; x++) a[x] = -ap[x * 8 * LDT + k0];
#pragma unroll
        for (int y = 0; y < 8; y++) b[y] = bp[y * 8 * LDT + k0];
#pragma unroll
        for (int x = 0; x < 2; x++)
#pragma unroll
            for (int y = 0; y < 8; y++) dmma884(acc[x][y], a[x], b[y]);
    }
}

__global__ void lat_kernel(double* out, long long* cyc, double seed) {
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001, c2[2] = {seed, seed};
    const int lane = threadIdx.x & 31;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) x = fma(x, y, 1e-9);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) x = x * y;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // SHFL.64 chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) x = __shfl_sync(0xffffffffu, x, (lane & ~3) | ((lane + 1) & 3));
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // DMMA dependent chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) dmma884(c2, x, y);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // DMMA independent (16 accumulators)
    double acc[16][2];
#pragma unroll
    for (int u = 0; u < 16; u++) acc[u][0] = acc[u][1] = seed;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) dmma884(acc[u], x, y);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // mul -> shfl -> fma step (the TRSM column step)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            double xj = x * y;
            xj = __shfl_sync(0xffffffffu, xj, (lane & ~3) | (u & 3));
            x = fma(-y, xj, x);
        }
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    double s = x + c2[0] + c2[1];
#pragma unroll
    for (int u = 0; u < 16; u++) s += acc[u][0] + acc[u][1];
    out[threadIdx.x] = s;
}

// potrf64 / trsm / mma on an SPD tile; nrep repetitions, cycles of the whole loop from thread 0
template <int mode>
__global__ void __launch_bounds__(256, 2) block_kernel(const double* A, double* out, long long* cyc, int nrep, int nwarps_active) {
    extern __shared__ double smem[];
    double* Bs = smem;
    double* As = smem + NB * LDT;
    double* col = smem + 2 * NB * LDT;
    double* pivs = col + 2 * NB;
    double* rd = pivs + NB;
    __shared__ int bad;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, w4 = warp & 3;
    const int r = lane >> 2, c = lane & 3;
    if (tid == 0) bad = 0;
    for (int e = tid; e < NB * NB; e += blockDim.x) { Bs[(e >> 6) * LDT + (e & 63)] = A[e]; As[(e >> 6) * LDT + (e & 63)] = A[e] * 0.5; }
    __syncthreads();
    double acc[2][8][2];
    if (mode != 0 && mode != 4) {           // factor once so that Bs holds L and rd its reciprocal diagonal
        potrf64(Bs, LDT, Bs, rd, col, pivs, &bad);
    }
    __syncthreads();
    if (mode != 0 && mode != 4) {
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++) { acc[x][y][0] = A[(16 * w4 + 8 * x + r) * 64 + 8 * y + 2 * c]; acc[x][y][1] = A[(16 * w4 + 8 * x + r) * 64 + 8 * y + 2 * c + 1]; }
    }
    long long t0 = clock64();
    for (int it = 0; it < nrep; it++) {
        if (mode == 0) {
            for (int e = tid; e < NB * NB; e += blockDim.x) Bs[(e >> 6) * LDT + (e & 63)] = A[e];
            __syncthreads();
            potrf64(Bs, LDT, Bs, rd, col, pivs, &bad);
        } else if (mode == 1) {
            if (warp < nwarps_active) trsm_rows_inreg(acc, Bs, rd, lane);
            __syncthreads();
        } else if (mode == 2) {
            if (warp < nwarps_active) mma_slab<true>(acc, As, Bs, w4, lane);
            __syncthreads();
        } else if (mode == 4) {
            for (int e = tid; e < NB * NB; e += blockDim.x) Bs[(e >> 6) * LDT + (e & 63)] = A[e];
            __syncthreads();
            potrf64_v2(Bs, LDT, Bs, rd, col, pivs, &bad);
        } else if (mode == 5) {
            if (warp < nwarps_active) trsm_rows_inreg_v0(acc, Bs, rd, lane);
            __syncthreads();
        } else if (mode == 6) {
            if (warp < nwarps_active) mma_slab_u<1>(acc, As, Bs, w4, lane);
            __syncthreads();
        } else if (mode == 7) {
            if (warp < nwarps_active) mma_slab_u<4>(acc, As, Bs, w4, lane);
            __syncthreads();
        } else if (mode == 3) {   // staged thread-per-vector substitution of version 1 (64 vectors per tile)
            double* V = As;
            if (tid < 64 * (nwarps_active / 4 > 0 ? nwarps_active / 4 : 1)) subst_lower(Bs, LDT, rd, V, 129, tid);
            __syncthreads();
        }
    }
    long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0;
    if (mode != 0 && mode != 4) {
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++) s += acc[x][y][0] + acc[x][y][1];
    }
    out[blockIdx.x * 256 + tid] = s + Bs[tid] + rd[tid & 63];
}

// bit comparison: potrf64 v3 vs v2 and trsm pipelined vs v0 on the same data
__global__ void __launch_bounds__(256, 2) compare_kernel(const double* A, unsigned long long* mism) {
    extern __shared__ double smem[];
    double* B3 = smem;
    double* B2 = smem + NB * LDT;
    double* col = smem + 2 * NB * LDT;
    double* pivs = col + 2 * NB;
    double* rd = pivs + NB;
    double* rd2 = rd + NB;
    __shared__ int bad;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, w4 = warp & 3, r = lane >> 2, c = lane & 3;
    if (tid == 0) bad = 0;
    for (int e = tid; e < NB * NB; e += 256) { B3[(e >> 6) * LDT + (e & 63)] = A[e]; B2[(e >> 6) * LDT + (e & 63)] = A[e]; }
    __syncthreads();
    potrf64(B3, LDT, B3, rd, col, pivs, &bad);
    __syncthreads();
    potrf64_v2(B2, LDT, B2, rd2, col, pivs, &bad);
    __syncthreads();
    unsigned long long d = 0;
    for (int e = tid; e < NB * NB; e += 256)
        if (__double_as_longlong(B3[(e >> 6) * LDT + (e & 63)]) != __double_as_longlong(B2[(e >> 6) * LDT + (e & 63)])) d++;
    if (tid < NB && __double_as_longlong(rd[tid]) != __double_as_longlong(rd2[tid])) d++;
    if (d) atomicAdd(&mism[0], d);
    double acc[2][8][2], acc0[2][8][2];
#pragma unroll
    for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 8; y++) {
            acc[x][y][0] = acc0[x][y][0] = A[(16 * w4 + 8 * x + r) * 64 + 8 * y + 2 * c] + 0.01 * (x + y);
            acc[x][y][1] = acc0[x][y][1] = A[(16 * w4 + 8 * x + r) * 64 + 8 * y + 2 * c + 1] - 0.02 * r;
        }
    if (warp < 4) {
        trsm_rows_inreg(acc, B3, rd, lane);
        trsm_rows_inreg_v0(acc0, B3, rd, lane);
        unsigned long long d2 = 0;
#pragma unroll
        for (int x = 0; x < 2; x++)
#pragma unroll
            for (int y = 0; y < 8; y++)
                for (int e = 0; e < 2; e++)
                    if (__double_as_longlong(acc[x][y][e]) != __double_as_longlong(acc0[x][y][e])) d2++;
        if (d2) atomicAdd(&mism[1], d2);
    }
    if (bad && tid == 0) atomicAdd(&mism[2], 1ULL);
}

int main() {
    double *d_out; long long* d_cyc; double* d_A;
    cudaMalloc(&d_out, 8 * 1024 * 512); cudaMalloc(&d_cyc, 8 * 1024); cudaMalloc(&d_A, 8 * 64 * 64);
    std::vector<double> A(64 * 64);
    for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) A[i * 64 + j] = std::exp(-0.5 * (i - j) * (i - j) / 400.0) + (i == j ? 0.1 : 0.0);
    cudaMemcpy(d_A, A.data(), 8 * 64 * 64, cudaMemcpyHostToDevice);
    long long cyc[1024];
    {
        unsigned long long* d_m; unsigned long long hm[3] = {0, 0, 0};
        cudaMalloc(&d_m, 24); cudaMemset(d_m, 0, 24);
        const size_t sm2 = (2 * NB * LDT + 5 * NB) * sizeof(double);
        cudaFuncSetAttribute(compare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
        compare_kernel<<<1, 256, sm2>>>(d_A, d_m);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hm, d_m, 24, cudaMemcpyDeviceToHost);
        printf("bit comparison (%s): potrf64 v3 vs v2 mismatches %llu, trsm pipelined vs v0 mismatches %llu, bad pivots %llu\n",
               cudaGetErrorString(e), hm[0], hm[1], hm[2]);
    }
    lat_kernel<<<1, 32>>>(d_out, d_cyc, 1.0);
    cudaMemcpy(cyc, d_cyc, 8 * 8, cudaMemcpyDeviceToHost);
    const char* nm[] = {"DFMA dependent", "DMUL dependent", "SHFL.64 dependent", "DMMA dependent", "DMMA 16 independent (issue)", "mul->shfl->fma step"};
    for (int i = 0; i < 6; i++) printf("%-32s %7.2f cycles/op\n", nm[i], cyc[i] / 1024.0);
    const size_t smem = (2 * NB * LDT + 4 * NB + 64 * 129) * sizeof(double);
    typedef void (*kfn)(const double*, double*, long long*, int, int);
    kfn fns[8] = {block_kernel<0>, block_kernel<1>, block_kernel<2>, block_kernel<3>, block_kernel<4>, block_kernel<5>, block_kernel<6>, block_kernel<7>};
    for (int i = 0; i < 8; i++) cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* mn[] = {"potrf64 v3 (incl. tile copy)", "trsm_rows_inreg (pipelined)", "mma_slab", "subst_lower v1 (64 thr/tile)",
                        "potrf64 v2 (incl. tile copy)", "trsm_rows_inreg v0", "mma_slab unroll 1", "mma_slab unroll 4"};
    for (int grid : {1, 296}) {
        for (int mode = 0; mode < 8; mode++) {
            for (int nw : {4, 8}) {
                if ((mode == 0 || mode == 4) && nw == 4) continue;
                const int nrep = 20;
                fns[mode]<<<grid, 256, smem>>>(d_A, d_out, d_cyc, nrep, nw);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(cyc, d_cyc, 8 * grid, cudaMemcpyDeviceToHost);
                double m = 0; for (int i = 0; i < grid; i++) m += cyc[i]; m /= grid;
                printf("grid %3d  %-30s warps active %d : %9.0f cycles per call\n", grid, mn[mode], nw, m / nrep);
            }
        }
    }
    return 0;
}

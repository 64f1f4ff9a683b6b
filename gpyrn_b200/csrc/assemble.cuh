// assemble.cuh -- covariance-matrix assembly kernels (HBM-write bound).
//
// One thread evaluates k(t_i - t_j) for 16 elements of a 64x64 tile; the time stamps of the tile's
// rows and columns are staged in shared memory, stores are coalesced along the column index.
// The kernel function is a postfix program (opcodes of include/gprn_b200.h) interpreted per
// element; every thread of a CTA follows the same control path, so there is no divergence.
// Floating-point expressions follow the operation order of gpyrn/covfunc.py so that results agree
// with numpy to a few ulp.
#pragma once
#include "common.cuh"
#include "../../include/gprn_b200.h"

namespace gprn {

struct ProgTable {  // device pointers, one entry per matrix m of the model (nodes first, then weights)
    const int32_t* tok;      // [M][GPRN_MAX_PROG]
    const int32_t* len;      // [M]
    const int32_t* par_off;  // [M] offset of the matrix' first parameter inside a hyper set
};

// k(r).  `on_diag_pos`: element sits on the main diagonal by POSITION (WhiteNoise quirk Q9,
// covfunc.py:144-148); `wn_const`: rectangular evaluation -> WhiteNoise is the constant w^2.
__device__ __forceinline__ double eval_prog(const int32_t* __restrict__ tok, int ntok, const double* __restrict__ par,
                                            double r, bool on_diag_pos, bool wn_const) {
    double st[6];
    int sp = 0, pp = 0;
    const double ar = fabs(r);
    for (int t = 0; t < ntok; t++) {
        const int op = tok[t];
        switch (op) {
            case GPRN_OP_SE: {  // covfunc.py:169-170
                double th = par[pp], l = par[pp + 1];
                st[sp++] = (th * th) * exp(((-0.5) * (r * r)) / (l * l));
                pp += 2;
            } break;
            case GPRN_OP_PER: {  // covfunc.py:211-213
                double th = par[pp], P = par[pp + 1], l = par[pp + 2];
                double s = sin((M_PI * ar) / P);
                st[sp++] = (th * th) * exp(((-2.0) * (s * s)) / (l * l));
                pp += 3;
            } break;
            case GPRN_OP_QP: {  // covfunc.py:251-255
                double th = par[pp], le = par[pp + 1], P = par[pp + 2], lp = par[pp + 3];
                double s = sin((M_PI * ar) / P);
                double t1 = ((-2.0) * (s * s)) / (lp * lp);
                double t2 = (r * r) / (2.0 * (le * le));
                st[sp++] = (th * th) * exp(t1 - t2);
                pp += 4;
            } break;
            case GPRN_OP_RQ: {  // covfunc.py:286-288
                double th = par[pp], al = par[pp + 1], l = par[pp + 2];
                st[sp++] = (th * th) * pow(1.0 + (0.5 * (r * r)) / (al * (l * l)), -al);
                pp += 3;
            } break;
            case GPRN_OP_M32: {  // covfunc.py:370-373
                double th = par[pp], l = par[pp + 1];
                double s = (sqrt(3.0) * ar) / l;
                st[sp++] = ((th * th) * (1.0 + s)) * exp(-s);
                pp += 2;
            } break;
            case GPRN_OP_M52: {  // covfunc.py:391-396
                double th = par[pp], l = par[pp + 1];
                double num = ((3.0 * sqrt(5.0)) * l) * ar + 5.0 * (ar * ar);
                st[sp++] = ((th * th) * (1.0 + num / (3.0 * (l * l)))) * exp((-sqrt(5.0) * ar) / l);
                pp += 2;
            } break;
            case GPRN_OP_WN: {  // covfunc.py:144-148
                double w = par[pp];
                st[sp++] = (on_diag_pos || wn_const) ? w * w : 0.0;
                pp += 1;
            } break;
            case GPRN_OP_CONST: {  // covfunc.py:122-125
                double cst = par[pp];
                st[sp++] = cst * cst;
                pp += 1;
            } break;
            case GPRN_OP_RQP: {  // covfunc.py:307-310
                double th = par[pp], al = par[pp + 1], le = par[pp + 2], P = par[pp + 3], lp = par[pp + 4];
                double s = sin((M_PI * ar) / P);
                st[sp++] = ((th * th) * exp(((-2.0) * (s * s)) / (lp * lp))) *
                           pow(1.0 + (r * r) / ((2.0 * al) * (le * le)), -al);
                pp += 5;
            } break;
            case GPRN_OP_COS: {  // covfunc.py:327-328
                double th = par[pp], P = par[pp + 1];
                st[sp++] = (th * th) * cos(((2.0 * M_PI) * ar) / P);
                pp += 2;
            } break;
            case GPRN_OP_EXP: {  // covfunc.py:351-352
                double th = par[pp], l = par[pp + 1];
                st[sp++] = (th * th) * exp((-ar) / l);
                pp += 2;
            } break;
            case GPRN_OP_DSE: {  // covfunc.py:182-185
                double th = par[pp], l = par[pp + 1];
                double term1 = (th * th) / pow(l, 4.0);
                double term2 = l * l - r * r;
                st[sp++] = (term1 * term2) * exp(((-0.5) * (r * r)) / (l * l));
                pp += 2;
            } break;
            case GPRN_OP_DPER: {  // covfunc.py:215-221
                double th = par[pp], P = par[pp + 1], l = par[pp + 2];
                double rP = (M_PI * r) / P;
                double s = sin(rP), c = cos(rP);
                double term1 = (4.0 * (M_PI * M_PI)) * (th * th);
                double term2 = (l * l) * cos(2.0 * rP) - (4.0 * (s * s)) * (c * c);
                double term3 = exp(((-2.0) * (s * s)) / (l * l));
                st[sp++] = (term1 * term2) * term3;
                pp += 3;
            } break;
            case GPRN_OP_DQP: {  // covfunc.py:257-266
                double th = par[pp], le = par[pp + 1], P = par[pp + 2], lp = par[pp + 3];
                double P2 = P * P, lp2 = lp * lp, le2 = le * le, lp4 = pow(lp, 4.0), le4 = pow(le, 4.0);
                double term1 = (2.0 * (th * th)) / ((P2 * lp4) * le4);
                double sp1 = sin((M_PI * r) / P), cp1 = cos((M_PI * r) / P);
                double term2 = (((P2 * lp4) * le2 - ((2.0 * P2) * lp4) * (r * r)) -
                                ((((4.0 * M_PI) * P) * lp2) * le2) * r * sin(((2.0 * M_PI) * r) / P)) +
                               (((2.0 * (M_PI * M_PI)) * lp2) * le4) * cos(((2.0 * M_PI) * r) / P) -
                               ((((8.0 * (M_PI * M_PI)) * le4) * (sp1 * sp1)) * (cp1 * cp1));
                double term3 = exp((-(lp2 * (r * r) + (2.0 * le2) * (sp1 * sp1))) / (lp2 * le2));
                st[sp++] = (term1 * term2) * term3;
                pp += 4;
            } break;
            case GPRN_OP_GEXP: {  // covfunc.py:431-432
                double th = par[pp], ga = par[pp + 1], l = par[pp + 2];
                st[sp++] = (th * th) * exp(-pow(ar / l, ga));
                pp += 3;
            } break;
            case GPRN_OP_PIECE: {  // covfunc.py:470-474 (compact support: 0 beyond |r| = eta / 2)
                double eta = par[pp];
                double x = fabs(r / (0.5 * eta));
                double om = 1.0 - x;
                st[sp++] = x > 1.0 ? 0.0 : (3.0 * x + 1.0) * ((om * om) * om);
                pp += 1;
            } break;
            case GPRN_OP_PAC: {  // covfunc.py:493-496
                double am = par[pp], l1 = par[pp + 1], l2 = par[pp + 2];
                double den = l1 * l1 + l2 * l2;
                double a = sqrt(((2.0 * l1) * l2) / den);
                double b = exp((((-2.0) * r) * r) / den);
                st[sp++] = ((am * am) * a) * b;
                pp += 3;
            } break;
            case GPRN_OP_NPER: {  // covfunc.py:517-519
                double am = par[pp], al = par[pp + 1], P = par[pp + 2], l = par[pp + 3];
                double s = sin((M_PI * ar) / P);
                st[sp++] = (am * am) * pow(1.0 + (2.0 * (s * s)) / (al * (l * l)), -al);
                pp += 4;
            } break;
            case GPRN_OP_QNPER: {  // covfunc.py:543-546
                double am = par[pp], al = par[pp + 1], le = par[pp + 2], P = par[pp + 3], lp = par[pp + 4];
                double s = sin((M_PI * ar) / P);
                double a = pow(1.0 + (2.0 * (s * s)) / (al * (lp * lp)), -al);
                double b = exp(((-0.5) * (r * r)) / (le * le));
                st[sp++] = ((am * am) * a) * b;
                pp += 5;
            } break;
            case GPRN_OP_COSP: {  // covfunc.py:664-665
                double am = par[pp], P = par[pp + 1], l = par[pp + 2];
                double cs = cos((M_PI * ar) / P);
                st[sp++] = (am * am) * exp(((-2.0) * (cs * cs)) / (l * l));
                pp += 3;
            } break;
            case GPRN_OP_QCOSP: {  // covfunc.py:686-688
                double am = par[pp], le = par[pp + 1], P = par[pp + 2], lp = par[pp + 3];
                double cs = cos((M_PI * ar) / P);
                st[sp++] = (am * am) * exp(((-2.0) * (cs * cs)) / (lp * lp) - (r * r) / (2.0 * (le * le)));
                pp += 4;
            } break;
            case GPRN_OP_ADD: {
                sp--;
                st[sp - 1] = st[sp - 1] + st[sp];
            } break;
            case GPRN_OP_MUL: {
                sp--;
                st[sp - 1] = st[sp - 1] * st[sp];
            } break;
            default: break;
        }
    }
    return st[0];
}

// Symmetric assembly of K_m = k_m(t - t^T) + nugget*I for every matrix of every set in the chunk.
// grid = (nt*(nt+1)/2 lower tiles, M, number of slots), block = 256.  Writes tiles I >= J; diagonal tiles full.
// Padding (index >= N): identity.  slots: workspace slots to assemble (null: slot = blockIdx.z);
// slot_set: row of `hyper` that each slot holds (null: the slot index itself).
__global__ void __launch_bounds__(256) kassemble_sym_kernel(double* __restrict__ K, const double* __restrict__ time,
                                                            const double* __restrict__ hyper, int H, ProgTable pt,
                                                            int M, int N, int Np, double nugget,
                                                            const int* __restrict__ slots,
                                                            const int* __restrict__ slot_set) {
    GPRN_TRACE_SCOPE(TK_KASSEMBLE);
    __shared__ double ti[NB], tj[NB];
    __shared__ int32_t stok[GPRN_MAX_PROG];
    __shared__ double spar[GPRN_MAX_PROG * 4];
    int I, J;
    tri_decode(blockIdx.x, I, J);
    const int m = blockIdx.y, set = slots ? slots[blockIdx.z] : (int)blockIdx.z;
    const int hrow = slot_set ? slot_set[set] : set;
    const int tid = threadIdx.x;
    const int ntok = pt.len[m];
    if (tid < NB) {
        int gi = I * NB + tid, gj = J * NB + tid;
        ti[tid] = gi < N ? time[gi] : 0.0;
        tj[tid] = gj < N ? time[gj] : 0.0;
    }
    if (tid < ntok) stok[tid] = pt.tok[m * GPRN_MAX_PROG + tid];
    if (tid < GPRN_MAX_PROG * 4) {
        int po = pt.par_off[m] + tid;
        spar[tid] = po < H ? hyper[(size_t)hrow * H + po] : 0.0;
    }
    __syncthreads();
    double* Kt = K + ((size_t)set * M + m) * Np * Np;
    // thread = two adjacent columns x 8 rows: a warp stores one full 512-byte tile row per instruction (16-byte stores)
    const int c2 = tid & 31, rg = tid >> 5;
    const int gj = J * NB + 2 * c2;
#pragma unroll 2
    for (int u = 0; u < 8; u++) {
        const int rr = rg * 8 + u, gi = I * NB + rr;
        double2 v;
        if (gi < N && gj < N) {
            v.x = eval_prog(stok, ntok, spar, ti[rr] - tj[2 * c2], gi == gj, false);
            if (gi == gj) v.x += nugget;
        } else {
            v.x = (gi == gj) ? 1.0 : 0.0;
        }
        if (gi < N && gj + 1 < N) {
            v.y = eval_prog(stok, ntok, spar, ti[rr] - tj[2 * c2 + 1], gi == gj + 1, false);
            if (gi == gj + 1) v.y += nugget;
        } else {
            v.y = (gi == gj + 1) ? 1.0 : 0.0;
        }
        *reinterpret_cast<double2*>(Kt + (size_t)gi * Np + gj) = v;
    }
}

// Rectangular assembly K[r][c] = k(trow[r] - tcol[c]) (+ nugget on the diagonal when `square`),
// row-major with leading dimension ld, no padding.  grid = (ceil(ncols/64), ceil(nrows/64), nbatch), block 256.
// Used for Kstar (prediction) and for gprn_kmatrix.  `row0`: index of trow[0] in the caller's full row range (the
// diagonal-by-position rule of WhiteNoise, quirk Q9, applies to global positions when rows are processed in chunks).
// Batched form (blockIdx.z = b): matrix b is written at K + b*kstride and uses program / parameter set
// tokb[b*GPRN_MAX_PROG..], lenb[b], parb + paroff[b]; with tokb == null the single (tok, ntok, par) is used.
__global__ void __launch_bounds__(256) kassemble_rect_kernel(double* __restrict__ K, size_t ld,
                                                             const double* __restrict__ trow, int nrows,
                                                             const double* __restrict__ tcol, int ncols,
                                                             const int32_t* __restrict__ tok, int ntok,
                                                             const double* __restrict__ par, int npar, int square,
                                                             double nugget, int row0, size_t kstride,
                                                             const int32_t* __restrict__ lenb,
                                                             const int32_t* __restrict__ paroff) {
    __shared__ double ti[NB], tj[NB];
    __shared__ int32_t stok[GPRN_MAX_PROG];
    __shared__ double spar[GPRN_MAX_PROG * 4];
    const int tid = threadIdx.x;
    int poff = 0;
    if (lenb) {                    // batched: per-matrix program; parameters at par[paroff[b] ..] of an npar-long vector
        const int b = blockIdx.z;
        tok += (size_t)b * GPRN_MAX_PROG;
        ntok = lenb[b];
        poff = paroff[b];
        K += (size_t)b * kstride;
    }
    if (tid < NB) {
        int gi = blockIdx.y * NB + tid, gj = blockIdx.x * NB + tid;
        ti[tid] = gi < nrows ? trow[gi] : 0.0;
        tj[tid] = gj < ncols ? tcol[gj] : 0.0;
    }
    if (tid < ntok) stok[tid] = tok[tid];
    if (tid < GPRN_MAX_PROG * 4) spar[tid] = poff + tid < npar ? par[poff + tid] : 0.0;
    __syncthreads();
    const int c = tid & 63, rg = tid >> 6;
    const int gj = blockIdx.x * NB + c;
    if (gj >= ncols) return;
    for (int u = 0; u < 16; u++) {
        const int rr = rg * 16 + u, gi = blockIdx.y * NB + rr;
        if (gi >= nrows) break;
        double v = eval_prog(stok, ntok, spar, ti[rr] - tj[c], square && gi + row0 == gj, !square);
        if (square && gi + row0 == gj) v += nugget;
        K[(size_t)gi * ld + gj] = v;
    }
}

// out[e] = k(r[e]) for an arbitrary array of lags (covFunction.__call__).  1-D grid-stride.
__global__ void keval_kernel(double* __restrict__ out, const double* __restrict__ r, long long n, long long ncols,
                             const int32_t* __restrict__ tok, int ntok, const double* __restrict__ par, int square) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        bool dg = square && (e / ncols == e % ncols);
        out[e] = eval_prog(tok, ntok, par, r[e], dg, !square);
    }
}

}  // namespace gprn

"""warp_model.py -- lane-level numpy model of the register-resident tile routines of gpyrn_b200/csrc/common.cuh
(mma_slab, trsm_rows_inreg, the row-sum / column-sum reductions of the fused small-N kernel).

TEST INFRASTRUCTURE ONLY (like everything under oracle/): tests/test_warp_model.py uses it to check, on the CPU and
without a GPU, that the fragment index arithmetic of those device routines solves  X L^T = T  and reduces the right
elements.  A "register" is a length-32 array (one value per lane); shfl / dmma884 follow the PTX semantics that
common.cuh states for mma.sync.m8n8k4.f64:
    a = A[lane/4][lane%4],  b = B[lane%4][lane/4],  c = {C[lane/4][2*(lane%4)], C[lane/4][2*(lane%4)+1]}.
"""
import numpy as np

LANES = np.arange(32)
R = LANES >> 2
C = LANES & 3


def shfl(v, src):
    return v[src]


def dmma884(c0, c1, a, b):
    """(c0, c1) += A(8x4) * B(4x8) in the m8n8k4 fragment layout; returns the new (c0, c1)."""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    Cm = np.zeros((8, 8))
    A[R, C] = a
    B[C, R] = b
    Cm[R, 2 * C] = c0
    Cm[R, 2 * C + 1] = c1
    Cm = Cm + A @ B
    return Cm[R, 2 * C], Cm[R, 2 * C + 1]


def slab_from_matrix(T, w4):
    """acc[x][y][e][lane] = T[16*w4 + 8x + r][8y + 2c + e]  (the 16 x 64 slab of warp w4)."""
    acc = np.zeros((2, 8, 2, 32))
    for x in range(2):
        for y in range(8):
            for e in range(2):
                acc[x, y, e] = T[16 * w4 + 8 * x + R, 8 * y + 2 * C + e]
    return acc


def matrix_from_slabs(accs):
    T = np.zeros((64, 64))
    for w4, acc in enumerate(accs):
        for x in range(2):
            for y in range(8):
                for e in range(2):
                    T[16 * w4 + 8 * x + R, 8 * y + 2 * C + e] = acc[x, y, e]
    return T


def mma_slab(acc, As, Bs, w4, neg=False):
    """acc(16x64 slab of warp w4) += As[rows of the slab][k] * Bs[n][k]^T (common.cuh: mma_slab)."""
    acc = acc.copy()
    for k0 in range(0, 64, 4):
        a = [As[16 * w4 + 8 * x + R, k0 + C] * (-1.0 if neg else 1.0) for x in range(2)]
        b = [Bs[8 * y + R, k0 + C] for y in range(8)]
        for x in range(2):
            for y in range(8):
                acc[x, y, 0], acc[x, y, 1] = dmma884(acc[x, y, 0], acc[x, y, 1], a[x], b[y])
    return acc


def trsm_rows_inreg(acc, Ls, rd, ymin=0):
    """Solve X L^T = T in place for a warp's 16 x 64 slab in accumulator layout (common.cuh: trsm_rows_inreg),
    including its software pipelining: block y + 1 receives the update of block y at once, the blocks beyond it
    between the column steps of the next solve."""
    acc = acc.copy()
    qbase = LANES & ~3
    pa0, pa1 = [None, None], [None, None]

    def update(yy, ysrc):
        b0 = Ls[8 * yy + R, 8 * ysrc + C]
        b1 = Ls[8 * yy + R, 8 * ysrc + 4 + C]
        for x in range(2):
            acc[x, yy, 0], acc[x, yy, 1] = dmma884(acc[x, yy, 0], acc[x, yy, 1], pa0[x], b0)
            acc[x, yy, 0], acc[x, yy, 1] = dmma884(acc[x, yy, 0], acc[x, yy, 1], pa1[x], b1)

    for y in range(8):
        if y < ymin:
            continue
        row0 = 8 * y + 2 * C                      # this lane's two rows of the diagonal block of L
        for j in range(8):
            rdj = rd[8 * y + j]
            l0 = Ls[row0, 8 * y + j]
            l1 = Ls[row0 + 1, 8 * y + j]
            for x in range(2):
                xj = (acc[x, y, 1] if (j & 1) else acc[x, y, 0]) * rdj
                xj = shfl(xj, qbase | (j >> 1))
                own = C == (j >> 1)
                acc[x, y, j & 1] = np.where(own, xj, acc[x, y, j & 1])
                acc[x, y, 0] = np.where(2 * C > j, acc[x, y, 0] - l0 * xj, acc[x, y, 0])
                acc[x, y, 1] = np.where(2 * C + 1 > j, acc[x, y, 1] - l1 * xj, acc[x, y, 1])
            if y >= 1 and y + 1 + j <= 7 and y > ymin:
                update(y + 1 + j, y - 1)
        if y == 7:
            break
        for x in range(2):
            v0 = shfl(acc[x, y, 0], qbase | (C >> 1))
            v1 = shfl(acc[x, y, 1], qbase | (C >> 1))
            w0 = shfl(acc[x, y, 0], qbase | 2 | (C >> 1))
            w1 = shfl(acc[x, y, 1], qbase | 2 | (C >> 1))
            pa0[x] = -np.where(C & 1, v1, v0)
            pa1[x] = -np.where(C & 1, w1, w0)
        update(y + 1, y)
    return acc


def identity_slab(w4):
    acc = np.zeros((2, 8, 2, 32))
    for x in range(2):
        for y in range(8):
            for e in range(2):
                acc[x, y, e] = (16 * w4 + 8 * x + R == 8 * y + 2 * C + e).astype(float)
    return acc


def row_sumsq(acc):
    """Per slab row: sum of squares over the 64 columns; returns out[x][lane] (valid in every lane of the quad)."""
    out = np.zeros((2, 32))
    for x in range(2):
        s = np.zeros(32)
        for y in range(8):
            for e in range(2):
                s = s + acc[x, y, e] * acc[x, y, e]
        s = s + shfl(s, LANES ^ 1)
        s = s + shfl(s, LANES ^ 2)
        out[x] = s
    return out


def col_weighted_sums(acc, vrow):
    """Per column m: sum over the slab's 16 rows n of acc[n][m] * v[n]; vrow[x][lane] = v of row 8x + r.
    Returns pz[y][e][lane], valid in the lanes with r == 0 (lane < 4) -- and, the butterfly being symmetric, in all."""
    pz = np.zeros((8, 2, 32))
    for y in range(8):
        for e in range(2):
            t = acc[0, y, e] * vrow[0]
            t = t + acc[1, y, e] * vrow[1]
            t = t + shfl(t, LANES ^ 4)
            t = t + shfl(t, LANES ^ 8)
            t = t + shfl(t, LANES ^ 16)
            pz[y, e] = t
    return pz

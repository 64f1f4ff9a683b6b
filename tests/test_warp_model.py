"""CPU check of the fragment index arithmetic of the register-resident tile routines (common.cuh: mma_slab,
trsm_rows_inreg, and the reductions of the fused small-N kernel) through the lane-level model oracle/warp_model.py."""
import numpy as np

from oracle import warp_model as wm


def _lower(rng):
    L = np.tril(rng.standard_normal((64, 64))) * 0.3
    L[np.arange(64), np.arange(64)] = 1.0 + rng.random(64)
    return L


def test_mma_slab_is_a_times_b_transposed():
    rng = np.random.default_rng(0)
    A, B, C0 = rng.standard_normal((64, 64)), rng.standard_normal((64, 64)), rng.standard_normal((64, 64))
    for neg in (False, True):
        accs = [wm.mma_slab(wm.slab_from_matrix(C0, w), A, B, w, neg=neg) for w in range(4)]
        got = wm.matrix_from_slabs(accs)
        want = C0 + (-1.0 if neg else 1.0) * A @ B.T
        assert np.max(np.abs(got - want)) < 1e-12


def test_trsm_rows_inreg_solves_x_lt_equals_t():
    rng = np.random.default_rng(1)
    L = _lower(rng)
    T = rng.standard_normal((64, 64))
    rd = 1.0 / np.diag(L)
    got = wm.matrix_from_slabs([wm.trsm_rows_inreg(wm.slab_from_matrix(T, w), L, rd) for w in range(4)])
    want = np.linalg.solve(L, T.T).T
    assert np.max(np.abs(got - want)) < 1e-10 * np.max(np.abs(want))


def test_identity_right_hand_side_with_skipped_blocks_gives_the_transposed_inverse():
    rng = np.random.default_rng(2)
    L = _lower(rng)
    rd = 1.0 / np.diag(L)
    got = wm.matrix_from_slabs([wm.trsm_rows_inreg(wm.identity_slab(w), L, rd, ymin=2 * w) for w in range(4)])
    want = np.linalg.inv(L).T                       # Y = X^T, upper triangular
    assert np.max(np.abs(got - want)) < 1e-10 * np.max(np.abs(want))
    assert np.all(np.tril(got, -1) == 0.0)


def test_row_and_column_reductions():
    rng = np.random.default_rng(3)
    Y = rng.standard_normal((64, 64))
    v = rng.standard_normal(64)
    g = np.zeros(64)
    z = np.zeros(64)
    for w in range(4):
        acc = wm.slab_from_matrix(Y, w)
        rs = wm.row_sumsq(acc)
        for x in range(2):
            g[16 * w + 8 * x + wm.R[wm.C == 0]] = rs[x][wm.C == 0]
        vrow = np.stack([v[16 * w + 8 * x + wm.R] for x in range(2)])
        pz = wm.col_weighted_sums(acc, vrow)
        for y in range(8):
            for e in range(2):
                z[8 * y + 2 * wm.C[:4] + e] += pz[y, e][:4]
    assert np.allclose(g, np.sum(Y * Y, axis=1), rtol=1e-13)
    assert np.allclose(z, Y.T @ v, rtol=1e-12, atol=1e-13)

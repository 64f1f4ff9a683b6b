"""Per-CTA execution trace of one batched ELBO call (development aid; run on the GPU box).

    python tools/trace_run.py N p q NODE B MAX_ITER [out.npz]
    python tools/trace_run.py --analyse out.npz          (analysis only, no GPU)

Builds a second library with -DGPRN_TRACE (gpyrn_b200/csrc/libgprn_b200_trace.so; every CTA of the factorisation
kernels stamps %globaltimer / %smid), runs one warm-up call and one traced call, and prints where the SM time of the
traced call went: per kernel the CTA count, SM-time share and mean CTA duration; overall SM occupancy by
"GEMM-class" kernels (the 128x128 DMMA launches), by latency-class kernels, and idle.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpyrn_b200 import _lib  # noqa: E402

NAMES = {1: "panel_col", 2: "potrf_col", 3: "trsm_col", 4: "syrk64", 5: "syrk_outer", 6: "trtri_diag", 7: "trtri_row",
         8: "trtri_outer", 9: "trtri_inblock", 10: "trmv_lower", 11: "trmv_upper", 12: "cross_frob", 13: "form_a",
         14: "kassemble", 15: "small", 16: "other"}
GEMM = {5, 8, 12}


def build_trace_lib():
    out = os.path.join(_lib.CSRC, "libgprn_b200_trace.so")
    src = os.path.join(_lib.CSRC, "gprn_api.cu")
    if not os.path.exists(out) or os.path.getmtime(out) < max(
            os.path.getmtime(os.path.join(_lib.CSRC, f)) for f in os.listdir(_lib.CSRC) if f.endswith((".cu", ".cuh"))):
        subprocess.run(["nvcc"] + _lib.NVCC_FLAGS + ["-DGPRN_TRACE", "-o", out, src], check=True, cwd=_lib.CSRC)
    return out


def analyse(rec, nsm=148):
    t0, t1 = rec["t0"].astype(np.int64), rec["t1"].astype(np.int64)
    T0, T1 = t0.min(), t1.max()
    span = (T1 - T0) * 1e-3
    print(f"traced span {span:.1f} us, {rec.size} CTAs, SMs seen {np.unique(rec['smid']).size}")
    dur = (t1 - t0) * 1e-3
    tot = span * nsm
    print(f"{'kernel':16s} {'CTAs':>8s} {'SM-time %':>10s} {'mean us':>9s} {'max us':>9s}")
    for k in sorted(set(rec["kid"].tolist())):
        m = rec["kid"] == k
        print(f"{NAMES.get(k, k):16s} {m.sum():8d} {100 * dur[m].sum() / tot:10.2f} {dur[m].mean():9.1f} {dur[m].max():9.1f}")
    # time-resolved occupancy on a 4 us grid: which kernel classes hold each SM
    step = 4000
    ng = int((T1 - T0) // step + 1)
    kinds = sorted(set(rec["kid"].tolist()))
    occ = {k: np.zeros((nsm, ng), dtype=bool) for k in kinds}
    for k in kinds:
        m = rec["kid"] == k
        for a, b, sm in zip((t0[m] - T0) // step, (t1[m] - T0) // step + 1, rec["smid"][m] % nsm):
            occ[k][sm, a:b] = True
    g = np.zeros((nsm, ng), dtype=bool)
    for k in kinds:
        if k in GEMM:
            g |= occ[k]
    npres = sum(occ[k].astype(np.int8) for k in kinds if k not in GEMM)
    lat = (~g) & (npres > 0)
    print(f"SM-time with a GEMM-class CTA (syrk_outer / trtri_outer / cross_frob) resident: {100 * g.mean():.1f} %")
    print("SM-time with only latency-class CTAs resident, attributed by kernel (equal split among those present):")
    for k in kinds:
        if k not in GEMM:
            share = ((occ[k] & lat) / np.maximum(npres, 1)).sum() / (nsm * ng)
            if share > 5e-4:
                print(f"    {NAMES.get(k, k):14s} {100 * share:6.2f} %")
    print(f"SM-time idle: {100 * ((~g) & (npres == 0)).mean():.1f} %")
    nge = g.sum(axis=0)
    print("fraction of wall time by number of SMs running GEMM-class CTAs: "
          + ", ".join(f"{lo}-{hi}: {100 * np.mean((nge >= lo) & (nge <= hi)):.1f} %" for lo, hi in
                      ((0, 0), (1, 36), (37, 73), (74, 110), (111, 140), (141, 148))))


def main():
    if sys.argv[1] == "--analyse":
        analyse(np.load(sys.argv[2])["rec"])
        return
    N, p, q = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    node, B, max_iter = sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
    out = sys.argv[7] if len(sys.argv) > 7 else None
    _lib.LIB_PATH = build_trace_lib()
    import gpyrn_b200 as gp
    from gpyrn_b200 import covfunc, meanfunc
    from oracle import gprn_oracle as orc
    m = orc.synth(N, p, q, seed=1, node=node)
    theta = orc.perturbed_hyper_sets(m, B, 101)
    args = []
    for y, e in zip(m.y, m.yerr):
        args += [y, e]
    g = gp.inference(q, m.time, *args)
    K = {"QP": covfunc.QuasiPeriodic, "M52": covfunc.Matern52, "SE": covfunc.SquaredExponential}
    g.set_components([K[s[0]](*s[1:]) for s in m.nodes], [K[s[0]](*s[1:]) for s in m.weights],
                     [meanfunc.Constant(0.0)] * p, [0.1] * p)
    P = np.concatenate([theta[:, :-p], np.zeros((B, p)), theta[:, -p:]], axis=1)
    L = _lib.lib()
    L.gprn_trace_begin.argtypes = [ctypes.c_uint]
    L.gprn_trace_dump.argtypes = [ctypes.c_char_p]
    g.ELBO_batch(P, max_iter=max_iter)
    _lib.check(L.gprn_trace_begin(8 << 20))
    g.ELBO_batch(P, max_iter=max_iter)
    path = os.path.join(ROOT, "gpurun_out", "trace.bin")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    n = L.gprn_trace_dump(path.encode())
    ph = (ctypes.c_ulonglong * 12)()
    L.gprn_trace_small_phases(ph, 0)
    if sum(ph):
        names = ["other/loop", "chol: load A tiles", "chol: MMA update", "chol: stage", "chol: potrf64", "chol: subst",
                 "chol: store L", "inv: MMA", "inv: stage + L_ii", "inv: subst", "inv: g/z sums + store X", "u pass"]
        tot = float(sum(ph))
        print("fused small kernel, thread-0 clock64 share by phase (warm-up + traced call):")
        for nm, v in zip(names, ph):
            print(f"    {nm:26s} {100 * v / tot:6.2f} %")
    rec = np.fromfile(path, dtype=np.dtype([("t0", "<u8"), ("t1", "<u8"), ("kid", "<i4"), ("smid", "<i4")]))
    print(f"{n} records, device time {L.gprn_last_elbo_ms(g._h()):.1f} ms")
    analyse(rec)
    if out:
        np.savez_compressed(out, rec=rec)
    os.remove(path)


if __name__ == "__main__":
    main()

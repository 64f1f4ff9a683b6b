#!/usr/bin/env python
"""bench.py -- GPRN ELBO evaluations per second on B200 (BASELINE.json metric), with FP64 roofline and
the reference's CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c2] [--impl ours|reference]

A *step* is one batched pass of the hot path: every rank evaluates its `sets_per_gpu` hyper-parameter
sets (each a full ``ELBOcalc`` from mu=var='init' to its own convergence).  Weak scaling: per-GPU work is
fixed, the sets are independent (no data-path collective); the only exchange is the gather of ELBO values.

Workloads (BASELINE.json configs / SURVEY.md 8d):
    c4 (default): synth(N=4096, p=4, q=2, Matern52 nodes), theta_b = theta_0 * exp(0.1 z), seed 102
                  -- the configuration the north-star target is quoted on;
    c3: synth(N=256, p=4, q=1, QuasiPeriodic), 8192 sets per GPU, seed 101;
    c2: synth(N=500, p=4, q=1, QuasiPeriodic), single evaluation per GPU (latency case).

JSON line: see README / DESIGN.md "Measurement".  `value` = evaluations/s with hyper-parameters already
resident in HBM (device-pointer C-ABI entry, CUDA events on the launching stream); `e2e` = the same through
``inference.ELBO_batch`` with host buffers (H2D of the hyper sets and y - mean, D2H of ELBO/iters/status inside
the timed region).  `roofline` = algorithmic FP64 flops (M N^3/3 + n_it [(2/3) M + q(q-1)/2] N^3
per evaluation, with the iteration counts the run actually took) / device time, against the FP64 DMMA peak
measured on this pool (profiles/fp64_peaks_r01.json; MEASURED_PEAKS.json carries no FP64 entry).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...", printed when NCCL_DEBUG is set in the
# environment) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c4": dict(N=4096, p=4, q=2, node="M52", sets_per_gpu=4, seed=102,
               name="C4 synth(N=4096,p=4,q=2,Matern52 nodes, SE weights), batched ELBOcalc to convergence"),
    "c3": dict(N=256, p=4, q=1, node="QP", sets_per_gpu=8192, seed=101,
               name="C3 synth(N=256,p=4,q=1,QuasiPeriodic node, SE weights), 8192 hyper sets per GPU"),
    "c2": dict(N=500, p=4, q=1, node="QP", sets_per_gpu=1, seed=101,
               name="C2 synth(N=500,p=4,q=1,QuasiPeriodic node, SE weights), single ELBOcalc"),
}


def fp64_peak():
    """(TFLOP/s, description).  Measured FP64 DMMA peak of this pool's B200 (tools/fp64_peaks.cu)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")))
        return float(d["fp64_dmma_tflops_sustained"]), "measured (DMMA m8n8k4 micro-kernel, profiles/fp64_peaks_r01.json)"
    except Exception:
        return 37.0, "fallback (B200 FP64 nominal 148 SM x 64 FMA x 1.965 GHz)"


def algorithmic_flops(N, p, q, iters):
    """F_eval = M N^3/3 + n_it [(2/3) M + q(q-1)/2] N^3, summed over evaluations."""
    M = q * (p + 1)
    iters = np.asarray(iters, dtype=np.float64)
    return float(np.sum(M * N ** 3 / 3.0 + iters * ((2.0 / 3.0) * M + q * (q - 1) / 2.0) * N ** 3))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [s for s in sm if mx and s > 0.3 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_problem(w):
    import workloads
    a = workloads.synth_arrays(w["N"], w["p"], w["q"], seed=1, node=w["node"])
    return a, workloads.theta0(a)


def oracle_model(a):
    """cpu_baseline / reference arm only: the CPU oracle's model container for these inputs."""
    from oracle import gprn_oracle as orc
    return orc, orc.Model(a["t"], a["y"], a["yerr"], a["nodes"], a["weights"], None, a["jitters"])


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's numpy/scipy algorithm on host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample(wname, w, budget_s=30.0, iters_hint=None):
    """Time a bounded sample of the workload with the CPU oracle (all BLAS threads).  Returns dict."""
    import workloads
    a, th0 = build_problem(w)
    orc, m = oracle_model(a)
    theta = workloads.perturbed_sets(th0, 64, w["seed"])
    cores = os.cpu_count()
    t_start = time.perf_counter()
    if w["N"] >= 1024:
        # one ELBOaux iteration of one set; an evaluation is setup + (n_it + 1) of these (meanfield.py:627-636)
        mb = orc.model_with_hyper(m, theta[0])
        t0 = time.perf_counter()
        Kf, Kw, Lf, Lw = orc.build_matrices(mb)
        mu, var = orc.init_mu_var(mb)
        t_setup = time.perf_counter() - t0
        t0 = time.perf_counter()
        orc.elbo_aux(mb, Kf, Kw, Lf, Lw, mb.y - mb.mean_vals, mb.jitters ** 2, mu, var)
        t_aux = time.perf_counter() - t0
        src = "iteration count of this set from the GPU run beside it"
        if not iters_hint:
            try:    # counts recorded by tools/record_iterations.py on a B200 (parity tests: counts equal the reference's)
                iters_hint = int(json.load(open(os.path.join(ROOT, "profiles", f"iterations_{wname}.json")))["iterations"][0])
                src = f"iteration count of this set recorded in profiles/iterations_{wname}.json"
            except Exception:
                iters_hint, src = 50, "iteration count assumed"
        n_it = iters_hint
        per_eval = t_setup + (n_it + 1) * t_aux
        return {"value": 1.0 / per_eval, "unit": "elbo_evals/s", "cores": cores, "kind": "port",
                "sample": f"1 set: setup {t_setup:.1f} s + 1 ELBOaux iteration {t_aux:.1f} s measured; evaluation = setup + "
                          f"(n_it+1) iterations with n_it={n_it} ({src}); "
                          f"numpy/scipy oracle port of the reference algorithm, OpenBLAS threads={cores}",
                "seconds": time.perf_counter() - t_start}
    done, its = 0, 0
    while done < len(theta) and (time.perf_counter() - t_start) < budget_s:
        mb = orc.model_with_hyper(m, theta[done])
        try:
            _, _, _, it = orc.elbo_calc(mb)
            its += it
        except np.linalg.LinAlgError:
            pass
        done += 1
    dt = time.perf_counter() - t_start
    return {"value": done / dt, "unit": "elbo_evals/s", "cores": cores, "kind": "port",
            "sample": f"first {done} hyper sets of the workload, full ELBOcalc each (mean {its / max(done, 1):.1f} iterations), "
                      f"numpy/scipy oracle port of the reference algorithm, OpenBLAS threads={cores}",
            "seconds": dt}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    samples = []
    total = args.warmup + args.steps
    budget = 240.0
    for s in range(total):
        if s > 0 and time.perf_counter() - t0 > budget:
            break
        r = cpu_sample(args.workload, w, budget_s=20.0 if w["N"] < 1024 else 0.0)
        if s >= min(args.warmup, total - 1) or s == total - 1:
            samples.append(r)
    vals = [r["value"] for r in samples]
    v = float(np.mean(vals))
    out = {"metric": "elbo_evals_per_sec", "value": v, "unit": "elbo_evals/s", "n_gpus": args.gpus, "steps": len(samples),
           "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([r["seconds"] for r in samples])),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "impl": "reference", "config": {"workload": w["name"], "sets_per_gpu": w["sets_per_gpu"]},
           "cpu_baseline": {"value": v, "unit": "elbo_evals/s", "cores": samples[-1]["cores"], "kind": "port",
                            "sample": samples[-1]["sample"]},
           "e2e": {"value": v, "unit": "elbo_evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import gpyrn_b200 as gp
    from gpyrn_b200 import _lib, covfunc, meanfunc, distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; gpyrn_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import workloads
    a, th0 = build_problem(w)
    N, p, q = w["N"], w["p"], w["q"]
    B_local = args.sets_per_gpu or w["sets_per_gpu"]
    B = B_local * world
    # Equal-work shards: the sets differ a lot in cost (C4: 34...94 iterations), so a fair weak-scaling series needs
    # per-GPU work that does not depend on N.  With recorded iteration counts (profiles/iterations_<w>.json, written by
    # tools/record_iterations.py) the pool of 8 x sets_per_gpu sets is cut into 8 blocks of equal summed cost; rank r
    # always evaluates block r, whatever N is.  Without the record: round-robin shards, re-dealt after the warm-up.
    pool_B = B_local * max(8, world)
    theta = workloads.perturbed_sets(th0, pool_B, w["seed"])
    blocks = None
    try:
        recd = json.load(open(os.path.join(ROOT, "profiles", f"iterations_{args.workload}.json")))
        rec = recd["iterations"]
        if len(rec) >= pool_B and not args.no_balance:
            # lock-step cost model (time of one iteration with 1..k active sets, measured): a block that ends in a
            # long single-set tail costs more than its iteration sum says
            st = recd.get("lockstep_ms") if B_local == len(recd.get("lockstep_ms", [])) else None
            blocks = D.balanced_assignment(rec[:pool_B], pool_B // B_local, step_times=st)
            # blocks have (nearly) equal cost; order them by the distance of their cost from the mean so that block 0
            # -- the N = 1 workload -- is the most typical one and small N stay representative of the pool
            cost = [D.lockstep_cost([rec[i] for i in b], st) if st else float(sum(rec[i] for i in b)) for b in blocks]
            mean = sum(cost) / len(cost)
            blocks = [b for _, _, b in sorted(zip([abs(c - mean) for c in cost], range(len(blocks)), blocks),
                                              key=lambda x: (x[0], x[1]))]
    except Exception:
        blocks = None
    if blocks is not None:
        idx = blocks[rank]
        used = np.sort(np.concatenate(blocks[:world]))
    else:
        idx = D.shard_indices(B, world, rank, "strided")
        used = np.arange(B)
    theta_l = np.ascontiguousarray(theta[idx])
    KC = {"QP": covfunc.QuasiPeriodic, "M52": covfunc.Matern52, "SE": covfunc.SquaredExponential}
    ya = []
    for y, e in zip(a["y"], a["yerr"]):
        ya += [y, e]
    g = gp.inference(q, a["t"], *ya, device=local)
    g.set_components([KC[s[0]](*s[1:]) for s in a["nodes"]], [KC[s[0]](*s[1:]) for s in a["weights"]],
                     [meanfunc.Constant(0.0)] * p, [0.1] * p)
    P_l = np.concatenate([theta_l[:, :-p], np.zeros((B_local, p)), theta_l[:, -p:]], axis=1)
    L = _lib.lib()
    h = g._h()
    g._bind_model(g.nodes, g.weights)
    _lib.check(L.gprn_upload_ysub(h, _lib.dptr(_lib.f64(a["y"]))))
    H = theta_l.shape[1]
    d_hyper = torch.from_numpy(theta_l).cuda()
    d_elbo = torch.empty(B_local, dtype=torch.float64, device="cuda")
    d_iters = torch.empty(B_local, dtype=torch.int32, device="cuda")
    d_status = torch.empty(B_local, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def step_dev():
        _lib.check(L.gprn_elbo_batched_dev(h, B_local, d_hyper.data_ptr(), -1, d_elbo.data_ptr(), d_iters.data_ptr(),
                                           d_status.data_ptr(), stream.cuda_stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_dev()
    barrier()
    if world > 1 and blocks is None and not args.no_balance:
        # re-deal the sets so that every rank carries the same summed iteration count (measured by the warm-up
        # pass); counts per rank stay equal, so this is still weak scaling over the same global batch
        it_all = D.gather_results(idx, {"iters": d_iters.cpu().numpy().astype(np.int64)}, pool_B)["iters"]
        idx = used[D.balanced_assignment(it_all[used], world)[rank]]
        theta_l = np.ascontiguousarray(theta[idx])
        P_l = np.concatenate([theta_l[:, :-p], np.zeros((B_local, p)), theta_l[:, -p:]], axis=1)
        d_hyper.copy_(torch.from_numpy(theta_l))
        step_dev()
        barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    L.gprn_reset_launch_count(h)
    ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)                                  # evict L2 between steps (outside the timed events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_dev()
        e1.record(stream)
        e1.synchronize()
        ms += e0.elapsed_time(e1)
    barrier()
    launches = int(L.gprn_launch_count(h))
    iters_l = d_iters.cpu().numpy().astype(np.int64)
    status_l = d_status.cpu().numpy()
    elbo_l = d_elbo.cpu().numpy()
    # end to end through the public API (host buffers in, host results out)
    g.ELBO_batch(P_l[: max(1, min(B_local, 2))])
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_api, it_api, st_api = g.ELBO_batch(P_l, return_info=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    assert np.array_equal(e_api, elbo_l), "public API and device-pointer entry disagree"

    t_ms = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    fl = torch.tensor([algorithmic_flops(N, p, q, iters_l), float(iters_l.sum()), float(launches),
                       float((status_l != 0).sum())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(fl, op=dist.ReduceOp.SUM)
        # the one collective of the path: gather the ELBO values (NCCL over NVLink)
        res = D.gather_results(idx, {"elbo": elbo_l, "iters": iters_l}, pool_B)
        elbo_all = res["elbo"][used]
    else:
        elbo_all = elbo_l
    if rank == 0:
        ms_tot, e2e_ms = float(t_ms[0]), float(t_ms[1])
        flops_step, iters_step, launches_all, bad = [float(x) for x in fl]
        peak, peak_src = fp64_peak()
        value = B * args.steps / (ms_tot * 1e-3)
        achieved = flops_step * args.steps / (ms_tot * 1e-3) / 1e12 / world      # per GPU
        out = {"metric": "elbo_evals_per_sec", "value": value, "unit": "elbo_evals/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_tot / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": w["name"], "N": N, "p": p, "q": q, "sets_per_gpu": B_local, "global_sets": B,
                          "mean_iterations": iters_step / B, "elbo_iterations_per_sec": iters_step * args.steps / (ms_tot * 1e-3),
                          "not_converged_or_failed": int(bad),
                          "sharding": ("equal-cost blocks from recorded iteration counts and the measured lock-step iteration times (profiles/iterations_%s.json); rank r always evaluates block r" % args.workload)
                          if blocks is not None else ("round-robin" if (world == 1 or args.no_balance) else
                                                      "round-robin, re-dealt by warm-up iteration counts (equal sets per rank)"),
                          "l2": "256 MiB flush between steps; per-step working set (K, L, L^-1 per matrix) >> 126 MB L2",
                          "elbo_checksum": float(np.sum(elbo_all))},
               "e2e": {"value": B * args.steps / (e2e_ms * 1e-3), "unit": "elbo_evals/s",
                       "h2d_bytes_per_step": int(B_local * H * 8 + p * N * 8), "d2h_bytes_per_step": int(B_local * 16)},
               "gpu_launches": int(launches_all),
               "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                            "traffic": None, "peak_source": "FP64 " + peak_src,
                            "what": "whole batched evaluation, algorithmic FP64 flops (SURVEY.md 8d) / device time, per GPU"},
               "clocks": clocks}
        if world == 1 and not args.no_cpu:
            hint = int(iters_l[0]) if (N >= 1024 and blocks is None) else None   # set 0 is what the CPU leg times
            cb = cpu_sample(args.workload, w, budget_s=20.0, iters_hint=hint)
            cb.pop("seconds", None)
            out["cpu_baseline"] = cb
        print(json.dumps(out), flush=True)
    g.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GPRN_BENCH_WORKLOAD", "c4"), choices=sorted(WORKLOADS))
    ap.add_argument("--sets-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-balance", action="store_true", help="keep the round-robin shards (no cost-balanced re-deal)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()

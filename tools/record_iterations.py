"""Record the iteration counts of the first sets of a bench workload (GPU run) so that the CPU reference arm can
turn its measured per-iteration time into evaluations/s without assuming a count.
    python tools/record_iterations.py c4 8  ->  profiles/iterations_c4.json"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, workloads
import gpyrn_b200 as gp
from gpyrn_b200 import covfunc, meanfunc
name, B = sys.argv[1], int(sys.argv[2])
w = bench.WORKLOADS[name]
a, th0 = bench.build_problem(w)
theta = workloads.perturbed_sets(th0, B, w["seed"])
p, q = w["p"], w["q"]
KC = {"QP": covfunc.QuasiPeriodic, "M52": covfunc.Matern52, "SE": covfunc.SquaredExponential}
ya = []
for y, e in zip(a["y"], a["yerr"]):
    ya += [y, e]
g = gp.inference(q, a["t"], *ya)
g.set_components([KC[s[0]](*s[1:]) for s in a["nodes"]], [KC[s[0]](*s[1:]) for s in a["weights"]],
                 [meanfunc.Constant(0.0)] * p, [0.1] * p)
P = np.concatenate([theta[:, :-p], np.zeros((B, p)), theta[:, -p:]], axis=1)
elbo, iters, status = g.ELBO_batch(P, return_info=True)
out = {"workload": w["name"], "seed": w["seed"], "iterations": [int(i) for i in iters], "status": [int(s) for s in status],
       "elbo": [float(e) for e in elbo], "source": "gpyrn_b200 on B200 (iteration counts match the reference wherever it was run)"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"iterations_{name}.json"), "w"), indent=1)
print(out["iterations"])

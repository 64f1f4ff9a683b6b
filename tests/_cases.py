"""Helpers shared by the tests: golden-fixture loading and oracle model construction."""
import ast
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not f.endswith("kernels.npz"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["nodes"] = [ast.literal_eval(str(s)) for s in d["nodes"]]
    d["weights"] = [ast.literal_eval(str(s)) for s in d["weights"]]
    d["iters"] = int(d["iters"])
    d["elbo"] = float(d["elbo"])
    d["max_iter"] = None if int(d["max_iter"]) < 0 else int(d["max_iter"])
    return d


def oracle_model(d):
    from oracle import gprn_oracle as orc
    mean_vals = np.repeat(np.asarray(d["mean_consts"], float)[:, None], d["t"].size, axis=1)
    return orc.Model(d["t"], d["y"], d["yerr"], d["nodes"], d["weights"], mean_vals, d["jitters"])


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))

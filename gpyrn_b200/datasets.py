"""Loader for whitespace-separated observation tables with a one-line header, the format of the reference's bundled
``gpyrn/datasets/Solar_observations.txt`` (columns ``BJD RV RVerr RHK RHKerr S Serr BIS BISerr FWHM FWHMerr ...``;
the reference ships the file but no loader -- SURVEY.md 8(f).4).  Host-side convenience only: nothing here touches
the GPU.

    t, series = load_observations("Solar_observations.txt", ("RV", "FWHM", "BIS", "RHK"))
    gprn = gpyrn_b200.inference(1, t, *series)          # series = [RV, RVerr, FWHM, FWHMerr, ...]
"""
import numpy as np


def load_table(path):
    """Read the table into a dict ``column name -> float64 array`` (header order preserved)."""
    with open(path) as f:
        names = f.readline().split()
    data = np.loadtxt(path, skiprows=1, ndmin=2)
    if data.shape[1] != len(names):
        raise ValueError(f"{path}: header names {len(names)} columns, the rows have {data.shape[1]}")
    return {n: np.ascontiguousarray(data[:, i]) for i, n in enumerate(names)}


def load_observations(path, columns=("RV",), time="BJD", sort=True):
    """Time stamps and (value, error) pairs of the requested columns, in the argument order of ``inference``.

    The error of column ``X`` is the column named ``Xerr``; when the header spells it differently (the bundled file has
    ``Constrast`` / ``Contrasterr``) the column that follows ``X`` is taken if its name ends in ``err``."""
    tab = load_table(path)
    names = list(tab)
    if time not in tab:
        raise KeyError(f"no time column {time!r} in {path} (columns: {names})")
    order = np.argsort(tab[time], kind="stable") if sort else np.arange(tab[time].size)
    series = []
    for c in columns:
        if c not in tab:
            raise KeyError(f"no column {c!r} in {path} (columns: {names})")
        e = c + "err"
        if e not in tab:
            nxt = names[names.index(c) + 1] if names.index(c) + 1 < len(names) else None
            if nxt is None or not nxt.lower().endswith("err"):
                raise KeyError(f"no error column for {c!r} in {path}")
            e = nxt
        series += [tab[c][order].copy(), tab[e][order].copy()]
    return tab[time][order].copy(), series

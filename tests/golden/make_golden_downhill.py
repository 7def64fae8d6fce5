"""Frozen trajectories of the oracle's cv::DownhillSolver restatement (oracle/dp_oracle.c
orc_downhill; reference call site optimization_opencv.cpp:46-63) on analytic objectives.

cv::DownhillSolver's source is not in the reference tree, not in this image, and not in the
Python cv2 binding, so these vectors cannot pin the restatement against upstream.
What they do pin is DRIFT: every point the solver evaluates, in order (which encodes the
ilo / ihi / inhi choice and the reflect / expand / contract / shrink action of every
iteration), the evaluation count and the result, for
  * a smooth bowl, a Rosenbrock valley,
  * a piecewise-constant objective (values quantised to 1/64, like the photometric objective's
    1/32-px and u8 rounding): ties between vertices, the `<=` / `==` tie rules, the shrink step,
  * the evaluation cap.
Run from the repo root:  python tests/golden/make_golden_downhill.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402

CASES = {
    "bowl3": dict(x0=[0, 0, 0], step=[0.02, 0.2, 0.2], max_evals=500, eps=1e-4),
    "rosenbrock3": dict(x0=[-0.5, 0.3, 0.1], step=[0.02, 0.2, 0.2], max_evals=500, eps=1e-4),
    "quantised3": dict(x0=[0, 0, 0], step=[0.02, 0.2, 0.2], max_evals=500, eps=1e-4),
    "capped3": dict(x0=[0, 0, 0], step=[0.5, 0.5, 0.5], max_evals=60, eps=0.0),
}


def objective(name):
    if name == "bowl3":
        c = np.array([0.013, -0.07, 0.11])
        return lambda x: float(((x - c) ** 2 * np.array([40.0, 1.0, 2.0])).sum())
    if name == "rosenbrock3":
        return lambda x: float(100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2 +
                               100 * (x[2] - x[1] ** 2) ** 2 + (1 - x[1]) ** 2)
    if name == "quantised3":
        c = np.array([0.004, 0.05, -0.08])
        return lambda x: float(np.floor(64.0 * (np.abs(x - c) * np.array([30.0, 2.0, 3.0])).sum()) / 64.0)
    if name == "capped3":
        return lambda x: float(np.sin(37 * x[0]) + np.cos(23 * x[1]) + x[2] ** 2 + 3)
    raise KeyError(name)


def run(name):
    cfg = CASES[name]
    f = objective(name)
    pts = []
    x, res, fc = orc.downhill(lambda v: (pts.append([float(t) for t in v]), f(v))[1], cfg["x0"],
                              cfg["step"], max_evals=cfg["max_evals"], eps=cfg["eps"])
    return dict(points=pts, x=[float(t) for t in x], res=float(res), fcount=int(fc))


if __name__ == "__main__":
    out = {n: run(n) for n in CASES}
    for n, r in out.items():
        print(n, "fcount", r["fcount"], "evaluated", len(r["points"]), "res", r["res"])
    json.dump(out, open(os.path.join(HERE, "golden_downhill.json"), "w"))

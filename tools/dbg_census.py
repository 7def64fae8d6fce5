import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densepoints_b200 import capi, scenes
os.environ["DP_REFINE_TRACE"] = "/tmp/tr.bin"
os.environ["DP_LANE_MIN_PATCHES"] = "0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
sc = scenes.make_sphere_scene(seed=2, n_views=16, width=1280, height=960, f=1000.0)
seeds = scenes.make_seeds(sc, n, seed=200)
ctx = capi.Context(0)
ctx.set_views(sc.P, sc.images)
nvis, vis, _, _ = ctx.visibility(seeds["pos"], seeds["nrm"], seeds["ref"])
keep, fnv, fvi = ctx.filter(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis, 7)
pos, nrm, ev, xb = ctx.refine(seeds["pos"], seeds["nrm"], seeds["ref"], fnv, fvi, 7, mask=keep)
print("refined", int(keep.sum()), "evals", int((ev.astype(np.int64) * fnv * keep).sum()))

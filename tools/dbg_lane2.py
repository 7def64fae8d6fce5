"""Debug aid: textures / scores of the score kernel against the oracle, mismatch map."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from densepoints_b200 import capi, scenes
from oracle import oracle as orc

s = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sc = scenes.make_plane_scene(seed=1, n_views=3, width=640, height=480)
seeds = scenes.make_seeds(sc, n, seed=1)
orc.set_homography_mode(1)
V = orc.Views(sc.P, sc.images)
nvis, vis, _, _ = orc.visibility_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"])
a = (seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
ctx = capi.Context(0, capi.default_params(minimum_visible_image=2))
ctx.set_views(sc.P, sc.images)
ncc, tex, valid = ctx.score(*a, s, want_tex=True)
o_ncc, o_tex, o_valid = orc.score_batch(V, *a, s, want_tex=True)
print("valid equal", np.array_equal(valid, o_valid), "tex equal", np.array_equal(tex, o_tex),
      "ncc max diff", np.abs(ncc - o_ncc).max())
d = (tex != o_tex).any(-1)            # n, V, s, s
print("mismatching texels per (y,x):\n", d.sum(axis=(0, 1)))
bad = np.argwhere(d)
for b in bad[:6]:
    i, k, y, x = b
    print("patch", i, "view slot", k, "texel", (y, x), "gpu", tex[i, k, y, x], "oracle", o_tex[i, k, y, x])
print("patches with ncc diff > 1e-6:", np.where(np.abs(ncc - o_ncc).max(1) > 1e-6)[0][:20])
keep, fnv, fvi = ctx.filter(*a, s)
ok, onv, ovi = orc.filter_batch(V, *a, s, 0.6, 2)
print("filter equal", np.array_equal(keep, ok), np.array_equal(fnv, onv), np.array_equal(fvi, ovi))

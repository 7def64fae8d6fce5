"""Debug aid: refine a few C1 patches with nm_max_evals = 4 (the solver evaluates the four
vertices of the initial simplex and returns the best) and compare with the objective computed
through dp_score_at at the same four points."""
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
for mx in (4, 500):
    ctx = capi.Context(0, capi.default_params(minimum_visible_image=2, nm_max_evals=mx))
    ctx.set_views(sc.P, sc.images)
    pos, nrm, ev, xb = ctx.refine(*a, s)
    op, on, oev, oxb = orc.refine_batch(V, *a, s, orc.default_params(minimum_visible_image=2, nm_max_evals=mx))
    bad = np.where((ev != oev) | (np.abs(xb - oxb).max(1) > 1e-12))[0]
    print(f"s={s} max_evals={mx}: {len(bad)} of {n} differ; first: {bad[:10]}")
    for i in bad[:4]:
        print("  patch", i, "nvis", nvis[i], "gpu ev/x", ev[i], xb[i], "oracle", oev[i], oxb[i])
        verts = [(-0.01, -0.1, -0.1), (0.01, 0.0, 0.0), (0.0, 0.1, 0.0), (0.0, 0.0, 0.1)]
        print("   oracle f at vertices", [orc.objective(V, a[2][i], vis[i, :nvis[i]], s, a[1][i], a[0][i], np.array(v)) for v in verts])
    ctx.close()

# ---- trace of the first 8 objective values (debug build: tools/build_variant.sh trace -DDP_DEBUG_TRACE)
if os.environ.get("DP_REFINE_TRACE"):
    ctx = capi.Context(0, capi.default_params(minimum_visible_image=2))
    ctx.set_views(sc.P, sc.images)
    pos, nrm, ev, xb = ctx.refine(*a, s)
    tr = np.fromfile(os.environ["DP_REFINE_TRACE"], np.float64).reshape(-1, 8)
    verts = [(-0.01, -0.1, -0.1), (0.01, 0.0, 0.0), (0.0, 0.1, 0.0), (0.0, 0.0, 0.1)]
    for i in range(8):
        of = [orc.objective(V, a[2][i], vis[i, :nvis[i]], s, a[1][i], a[0][i], np.array(v)) for v in verts]
        print("patch", i, "gpu first evals", tr[i, :6], "\n     oracle vertices", of)

"""Oracle self-consistency: the parts of the path the reference's own tests do not pin
(Nelder-Mead, filter off-by-one, organizer, FIFO expansion).  No GPU."""
import numpy as np
import pytest

from densepoints_b200 import scenes


# ---- cv::DownhillSolver restatement (parity UNPINNED against upstream OpenCV) -------------------

def test_downhill_quadratic_bowl(orc):
    f = lambda x: float(((x - np.array([0.3, -0.2, 0.1])) ** 2).sum())
    x, res, fc = orc.downhill(f, [0, 0, 0], [0.02, 0.2, 0.2], max_evals=5000, eps=1e-12)
    assert np.abs(x - [0.3, -0.2, 0.1]).max() < 1e-4 and res < 1e-8 and 4 < fc <= 5003


def test_downhill_rosenbrock_2d(orc):
    f = lambda x: float(100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2)
    x, res, fc = orc.downhill(f, [-1.2, 1.0], [0.5, 0.5], max_evals=20000, eps=1e-14)
    assert np.abs(x - [1, 1]).max() < 1e-3


def test_downhill_initial_simplex_and_constant_function(orc):
    seen = []
    f = lambda x: (seen.append(x.copy()), 2.0)[1]
    x, res, fc = orc.downhill(f, [0, 0, 0], [0.02, 0.2, 0.2])
    # createInitialSimplex: v0 = x0 - step/2, v_i = x0 + step_i/2 e_i; all equal -> stop at once
    assert fc == 4 and res == 2.0
    assert np.allclose(seen[0], [-0.01, -0.1, -0.1])
    assert np.allclose(seen[1], [0.01, 0, 0]) and np.allclose(seen[2], [0, 0.1, 0])
    assert np.allclose(seen[3], [0, 0, 0.1])
    assert np.allclose(x, [0, 0, 0.1])          # `yval <= y[ilo]` keeps the LAST tied vertex


def test_downhill_max_evals_cap(orc):
    f = lambda x: float(np.sin(37 * x[0]) + np.cos(23 * x[1]) + x[2] ** 2 + 3)
    x, res, fc = orc.downhill(f, [0, 0, 0], [0.5, 0.5, 0.5], max_evals=50, eps=0.0)
    assert 50 <= fc <= 53                        # the stop test runs once per iteration


def test_downhill_decision_tree_trace(orc):
    """Reflection / expansion / contraction / shrink all occur and every tried point is the
    documented affine combination of the simplex."""
    calls = []
    f = lambda x: (calls.append(x.copy()), float(abs(x[0] - 1) + 3 * abs(x[1] + 2)))[1]
    x, res, fc = orc.downhill(f, [0, 0], [1.0, 1.0], max_evals=400, eps=1e-9)
    assert fc == len(calls)
    assert abs(x[0] - 1) < 1e-3 and abs(x[1] + 2) < 1e-3


# ---- filter off-by-one ---------------------------------------------------------------

def test_filter_off_by_one(orc, golden_scoring, golden_views):
    g = golden_scoring
    s = 7
    ncc = orc.score_batch(golden_views, g["pos"], g["nrm"], g["ref"], g["nvis"], g["vis"], s)
    keep, fnvis, fvis = orc.filter_batch(golden_views, g["pos"], g["nrm"], g["ref"], g["nvis"],
                                         g["vis"], s, 0.6, 2)
    saw_shift = False
    for i in range(len(keep)):
        nv = g["nvis"][i]
        vi = list(g["vis"][i, :nv])
        if nv < 2:
            assert keep[i] == 0 and fnvis[i] == nv
            continue
        # entry k-1 goes iff the score of entry k is low; the last entry always stays
        exp = [vi[k - 1] for k in range(1, nv) if not (ncc[i, k] < 0.6)] + [vi[-1]]
        assert list(fvis[i, :fnvis[i]]) == exp
        assert keep[i] == (len(exp) >= 2)
        if ncc[i, nv - 1] < 0.6 and all(ncc[i, k] >= 0.6 for k in range(1, nv - 1)):
            saw_shift = True                      # the bad view survives, a good one is dropped
    assert saw_shift


# ---- homography modes ----------------------------------------------------------------------------

def test_exact_homography_mode_differs_only_at_ties(orc, golden_scoring, golden_views):
    g = golden_scoring
    for s in (5, 7, 11, 16):
        orc.set_homography_mode(0)
        n0, t0, v0 = orc.score_batch(golden_views, g["pos"], g["nrm"], g["ref"], g["nvis"],
                                     g["vis"], s, want_tex=True)
        orc.set_homography_mode(1)
        try:
            n1, t1, v1 = orc.score_batch(golden_views, g["pos"], g["nrm"], g["ref"], g["nvis"],
                                         g["vis"], s, want_tex=True)
        finally:
            orc.set_homography_mode(0)
        assert np.array_equal(v0, v1)
        diff = (t0 != t1).any(axis=-1)
        assert diff[..., 1:, :].sum() == 0 and diff[..., 0, 1:].sum() == 0   # only texel (0,0)
        assert diff.sum() <= 3


# ---- organizer + expansion -----------------------------------------------------------------------

@pytest.fixture(scope="module")
def small_scene(orc):
    sc = scenes.make_plane_scene(seed=5, n_views=4, width=160, height=120, yaw_spread_deg=14.0)
    seeds = scenes.make_seeds(sc, 40, seed=6, depth_noise=0.004, tilt_deg=4.0)
    V = orc.Views(sc.P, sc.images)
    nvis, vis, _, _ = orc.visibility_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"])
    return sc, seeds, V, nvis, vis


def test_try_insert_semantics(orc, small_scene):
    sc, seeds, V, nvis, vis = small_scene
    org = orc.Organizer(V)
    i = int(np.argmax(nvis))
    idx, cells = org.try_insert(seeds["pos"][i], seeds["nrm"][i], seeds["ref"][i], vis[i, :nvis[i]])
    assert idx == 0 and len(cells) == nvis[i] and org.size() == 1
    for v, r, c in cells:
        assert org.grid(v)[r, c] == 1
        assert org.grid(v).shape == (120 // 8, 160 // 8)
    # the same patch again finds every cell taken -> rejected, store unchanged
    idx2, cells2 = org.try_insert(seeds["pos"][i], seeds["nrm"][i], seeds["ref"][i], vis[i, :nvis[i]])
    assert idx2 == -1 and len(cells2) == 0 and org.size() == 1
    # a patch that wins exactly one cell is rejected but the cell stays consumed
    org2 = orc.Organizer(V)
    one = vis[i, :1]
    idx3, cells3 = org2.try_insert(seeds["pos"][i], seeds["nrm"][i], seeds["ref"][i], one)
    assert idx3 == -1 and len(cells3) == 1 and org2.size() == 0
    v, r, c = cells3[0]
    assert org2.grid(v)[r, c] == 1


def test_level_synchronous_expansion_equals_fifo(orc, small_scene):
    sc, seeds, V, nvis, vis = small_scene
    prm = orc.default_params(minimum_visible_image=2)
    orgs = []
    for mode in ("fifo", "level"):
        org = orc.Organizer(V, prm)
        org.set_seeds(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
        n0 = org.size()
        pops = org.expand_fifo(5) if mode == "fifo" else org.expand(5, -1)
        orgs.append((org, pops, n0))
    (a, pa, n0), (b, pb, _) = orgs
    assert pa == pb and a.size() == b.size() and a.size() > n0     # it did expand
    ea, eb = a.export(), b.export()
    for k in ea:
        assert np.array_equal(ea[k], eb[k]), k
    for v in range(sc.n_views):
        assert np.array_equal(a.grid(v), b.grid(v))


def test_scene_is_deterministic():
    a = scenes.make_plane_scene(seed=3, n_views=2, width=64, height=48)
    b = scenes.make_plane_scene(seed=3, n_views=2, width=64, height=48)
    assert np.array_equal(a.P, b.P) and all(np.array_equal(x, y) for x, y in zip(a.images, b.images))
    sa, sb = scenes.make_seeds(a, 10, seed=1), scenes.make_seeds(b, 10, seed=1)
    assert all(np.array_equal(sa[k], sb[k]) for k in sa)


def test_downhill_frozen_trajectories(orc):
    """Every point the solver evaluates, in order, against the frozen vectors of
    tests/golden/make_golden_downhill.py (drift guard: upstream OpenCV cannot be run here)."""
    import json
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_golden_downhill as mg
    want = json.load(open(os.path.join(here, "golden_downhill.json")))
    for name in mg.CASES:
        got = mg.run(name)
        assert got["fcount"] == want[name]["fcount"], name
        assert got["points"] == want[name]["points"], name      # bit-exact doubles via JSON repr
        assert got["x"] == want[name]["x"] and got["res"] == want[name]["res"]
    # the piecewise-constant case really exercises ties and the shrink step
    q = want["quantised3"]
    vals = [mg.objective("quantised3")(np.array(p)) for p in q["points"]]
    assert len(set(vals)) < len(vals)
    assert want["capped3"]["fcount"] >= 60


def test_downhill_upstream_regression_cases(orc):
    """The two cases of upstream OpenCV's own regression test, AS RECALLED from
    modules/core/test/test_downhill_simplex.cpp (the file is not available offline, so this is a
    recollection, not a copy): SphereF from (1,1) with step (-0.5,-0.5) -> (0,0), RosenbrockF
    from (0,0) with step (0.5,0.5) -> (1,1); default TermCriteria(MAX_ITER+EPS, 5000, 1e-6);
    tolerance 1e-2 on the minimiser and the minimum."""
    sphere = lambda x: float(x[0] * x[0] + x[1] * x[1])
    x, res, fc = orc.downhill(sphere, [1.0, 1.0], [-0.5, -0.5], max_evals=5000, eps=1e-6)
    assert abs(res) < 1e-2 and np.abs(x).max() < 1e-2
    rosen = lambda x: float(100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2)
    x, res, fc = orc.downhill(rosen, [0.0, 0.0], [0.5, 0.5], max_evals=5000, eps=1e-6)
    assert abs(res) < 1e-2 and np.abs(x - 1.0).max() < 1e-2


def test_level_selection_follows_its_definition(orc):
    """Per-(patch, view) pyramid level (dp_set_level_selection): by definition the
    texture of view v read at level k is what the reference path gives on cv2.pyrDown^k of that
    view with P_k = diag(2^-k, 2^-k, 1) P while the patch frame still comes from the base-level
    reference view.  Checked by running the oracle WITHOUT selection on view sets in which every
    view but the reference one is replaced by its level-k version (images from the real
    cv2.pyrDown) and picking, pair by pair, the level the selection rule reports."""
    cv2 = pytest.importorskip("cv2")
    from densepoints_b200 import scenes
    zoom = [1, 2, 4]
    sc = scenes.make_plane_scene(seed=7, n_views=3, width=240, height=180, yaw_spread_deg=14.0)
    seeds = scenes.make_seeds(sc, 150, seed=9, depth_noise=0.004, tilt_deg=5.0)
    P0, I0 = [], []
    for P, im, z in zip(sc.P, sc.images, zoom):
        Pz = P.copy()
        Pz[:2, :] *= float(z)
        P0.append(Pz)
        I0.append(np.ascontiguousarray(np.repeat(np.repeat(im, z, axis=0), z, axis=1)))
    Ps, Is = [np.array(P0)], [I0]
    for _ in (1, 2):
        Ps.append(Ps[-1].copy())
        Ps[-1][:, :2, :] *= 0.5
        Is.append([cv2.pyrDown(im) for im in Is[-1]])
    lv = [orc.Views(P, I) for P, I in zip(Ps, Is)]
    pos, nrm, ref = seeds["pos"], seeds["nrm"], seeds["ref"]
    nvis, vis, _, _ = orc.visibility_batch(lv[0], pos, nrm, ref)
    s = 7
    orc.set_level_selection(lv, 1.5)
    try:
        picked = orc.levels_batch(lv[0], pos, nrm, ref, nvis, vis, s)
        ncc, tex, valid = orc.score_batch(lv[0], pos, nrm, ref, nvis, vis, s, want_tex=True)
    finally:
        orc.set_level_selection(None)
    assert all((picked == l).sum() > 20 for l in range(3))
    checked = 0
    for r in range(3):
        sel = np.where(ref == r)[0]
        if len(sel) == 0:
            continue
        for l in range(3):
            W = orc.Views([Ps[l][v] if v != r else Ps[0][r] for v in range(3)],
                          [Is[l][v] if v != r else Is[0][r] for v in range(3)])
            _, t_l, v_l = orc.score_batch(W, pos[sel], nrm[sel], ref[sel], nvis[sel], vis[sel], s,
                                          want_tex=True)
            for a, i in enumerate(sel):
                for k in range(nvis[i]):
                    if picked[i, k] == l and vis[i, k] != r:
                        assert valid[i, k] == v_l[a, k]
                        assert np.array_equal(tex[i, k], t_l[a, k])
                        checked += 1
    assert checked > 100
    # the reference view of a patch is sampled at ~1 pixel per texel by construction of the
    # patch frame (optimization.cpp:19-30), i.e. always at the base level
    k0 = np.arange(vis.shape[1])[None, :]
    is_ref = (vis == ref[:, None]) & (k0 < nvis[:, None])
    assert (picked[is_ref] == 0).all()


def test_eigen_sum_order_sensitivity(orc):
    """The reference's vector arithmetic is Eigen's, unpinned and absent here.  The oracle and the
    kernels sum small products sequentially (Eigen 3.2); Eigen >= 3.3 halves them,
    (a0+a1)+(a2+a3).  This measures what that choice can change on C1 (3 views 640x480, 2 000
    seeds, mu = 5): projections move by at most a few ulp of fp64, and no ROI, texel, score bit,
    visible set, filter decision or Nelder-Mead evaluation count changes on this sample -- the
    bit-exact claims hold under either Eigen up to events of probability ~1e-9 per coordinate."""
    from densepoints_b200 import scenes
    sc = scenes.make_plane_scene(seed=1, n_views=3, width=640, height=480)
    seeds = scenes.make_seeds(sc, 2000, seed=1)
    pos, nrm, ref = seeds["pos"], seeds["nrm"], seeds["ref"]
    out = {}
    for mode in (False, True):
        orc.set_eigen_pairwise(mode)
        try:
            V = orc.Views(sc.P, sc.images)
            nvis, vis, ncand, cand = orc.visibility_batch(V, pos, nrm, ref)
            ncc, tex, valid = orc.score_batch(V, pos, nrm, ref, nvis, vis, 5, want_tex=True)
            keep, fnv, fvi = orc.filter_batch(V, pos, nrm, ref, nvis, vis, 5, min_visible=2)
            p1, n1, ev, _ = orc.refine_batch(V, pos[:300], nrm[:300], ref[:300], nvis[:300], vis[:300], 5,
                                             orc.default_params(minimum_visible_image=2))
            uv = np.array([orc.project(V, int(r), p.astype(np.float64)) for r, p in zip(ref[:500], pos[:500])])
            out[mode] = dict(nvis=nvis, vis=vis, tex=tex, valid=valid, ncc=ncc, keep=keep, fnv=fnv, fvi=fvi,
                             ev=ev, p1=p1, n1=n1, uv=uv)
        finally:
            orc.set_eigen_pairwise(False)
    a, b = out[False], out[True]
    duv = np.abs(a["uv"] - b["uv"])
    assert duv.max() < 1e-11                                   # a few ulp of a ~500-px coordinate
    for k in ("nvis", "vis", "valid", "tex", "keep", "fnv", "fvi", "ev", "p1", "n1"):
        assert np.array_equal(a[k], b[k]), k
    assert np.abs(a["ncc"] - b["ncc"]).max() == 0
    print(f"Eigen sum order: max |d projection| {duv.max():.2e} px over {len(duv)} points, "
          f"{(duv > 0).sum()} coordinates differ in the last bits; {a['valid'].sum()} textures identical")


@pytest.mark.parametrize("m", [1, 2, 3])
def test_set_seeds_against_a_direct_restatement(orc, m):
    """PatchOrganizer::SetSeeds / TryInsert / PatchGrid::TryInsert (patch_organizer.cpp:15-75)
    restated line by line in Python -- cells as lists with a capacity of max_patches_per_cell,
    a patch kept iff it entered more than one cell, cells consumed either way --
    against the oracle's organizer, for 1, 2 and 3 patches per cell on seeds dense enough to
    fill cells."""
    sc = scenes.make_plane_scene(seed=5, n_views=4, width=160, height=120, yaw_spread_deg=14.0)
    seeds = scenes.make_seeds(sc, 700, seed=6, depth_noise=0.004, tilt_deg=4.0)
    V = orc.Views(sc.P, sc.images)
    nvis, vis, _, _ = orc.visibility_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"])
    prm = orc.default_params(minimum_visible_image=2, max_patches_per_cell=m)
    org = orc.Organizer(V, prm)
    acc = org.set_seeds(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
    gs = 8
    gw, gh = 160 // gs, 120 // gs                                  # AllocateViews, :32-40
    grids = [[[0] * gw for _ in range(gh)] for _ in range(sc.n_views)]
    want = []
    for i in range(len(nvis)):
        cells = 0
        for v in vis[i, :nvis[i]]:
            uv = orc.project(V, int(v), seeds["pos"][i].astype(np.float64))     # :46
            qr, qc = uv[1] / gs, uv[0] / gs
            if not (qr > -1 and qc > -1):                         # static_cast<size_t> of q <= -1: out of bounds
                continue
            row, col = int(qr), int(qc)                           # :47-48 (truncation toward zero)
            if col < gw and row < gh and grids[v][row][col] < m:  # PatchGrid::TryInsert, :18-26
                grids[v][row][col] += 1
                cells += 1
        want.append(1 if cells > 1 else 0)                        # :58
    assert np.array_equal(acc, np.array(want, acc.dtype))
    for v in range(sc.n_views):
        assert np.array_equal(org.grid(v), np.array(grids[v], np.uint8))
    assert max(max(max(r) for r in g) for g in grids) == m        # cells did fill up
    assert org.size() == sum(want)


def test_init_related_images_against_a_direct_restatement(orc):
    """Patch::InitRelatedImages (patch.cpp:19-49) restated in Python doubles -- reference view
    skipped, strict IsPointInside (types.cpp:77-84), angle = acos(n . d / |d|) with d = position -
    camera centre, < 0.78 -> visible, < 1.04 -> candidate, ascending view order -- against the
    oracle, bit for bit, on a 16-view scene (the visible sets are an integer output north_star
    wants bit-exact; the CUDA kernel is compared with the oracle on the GPU)."""
    import math
    sc = scenes.make_sphere_scene(seed=2, n_views=16, width=160, height=120, f=125.0)
    seeds = scenes.make_seeds(sc, 1500, seed=21)
    V = orc.Views(sc.P, sc.images)
    nvis, vis, ncand, cand = orc.visibility_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"])
    C = [V.center(i) for i in range(sc.n_views)]
    seen_vis = seen_cand = 0
    for i in range(len(nvis)):
        n = [float(x) for x in seeds["nrm"][i]]                   # fp32 storage -> double (patch.h:38-53)
        p = [float(x) for x in seeds["pos"][i]]
        wv, wc = [], []
        for v in range(sc.n_views):
            if v == seeds["ref"][i]:
                continue
            uv = orc.project(V, v, np.array(p))
            if not (uv[0] > 0 and uv[0] < 160 and uv[1] > 0 and uv[1] < 120):
                continue
            d = [p[0] - C[v][0], p[1] - C[v][1], p[2] - C[v][2]]
            dot = n[0] * d[0] + n[1] * d[1] + n[2] * d[2]
            nd = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
            q = dot / nd
            angle = math.acos(q) if -1.0 <= q <= 1.0 else float("nan")
            if angle < 0.78:
                wv.append(v)
            elif angle < 1.04:
                wc.append(v)
        assert list(vis[i, :nvis[i]]) == wv and nvis[i] == len(wv), i
        assert list(cand[i, :ncand[i]]) == wc and ncand[i] == len(wc), i
        seen_vis += len(wv)
        seen_cand += len(wc)
    assert seen_vis > 3000 and seen_cand > 100

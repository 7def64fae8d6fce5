"""GPU parity of the integer side of the path: organizer (TryInsert / SetSeeds), the
expansion loop, colours, the pyramid -- bit-exact against the CPU oracle's 1-thread FIFO."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi_mod():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from densepoints_b200 import build as b
    b.build_cuda()
    from densepoints_b200 import capi
    return capi


@pytest.fixture(scope="module")
def exact_orc(orc):
    orc.set_homography_mode(1)
    yield orc
    orc.set_homography_mode(0)


def _setup(capi_mod, orc, n_views, w, h, n_seeds, min_vis, seed=5):
    from densepoints_b200 import scenes
    sc = scenes.make_plane_scene(seed=seed, n_views=n_views, width=w, height=h, yaw_spread_deg=14.0)
    seeds = scenes.make_seeds(sc, n_seeds, seed=seed + 1, depth_noise=0.004, tilt_deg=4.0)
    ctx = capi_mod.Context(0, capi_mod.default_params(minimum_visible_image=min_vis))
    ctx.set_views(sc.P, sc.images)
    V = orc.Views(sc.P, sc.images)
    prm = orc.default_params(minimum_visible_image=min_vis)
    nvis, vis, _, _ = orc.visibility_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"])
    return sc, seeds, ctx, V, prm, nvis, vis


def _same_store(ctx, org, n_views):
    a, b = ctx.organizer_export(), org.export()
    assert ctx.organizer_size() == org.size()
    for k in ("ref", "nvis", "vis", "rgb", "pos", "nrm"):
        assert np.array_equal(a[k], b[k]), k          # incl. fp32 pos/nrm, bit-exact
    for v in range(n_views):
        assert np.array_equal(ctx.organizer_grid(v), org.grid(v)), f"grid {v}"


def test_set_seeds_matches_oracle(capi_mod, exact_orc):
    sc, seeds, ctx, V, prm, nvis, vis = _setup(capi_mod, exact_orc, 4, 320, 240, 3000, 2)
    ctx.organizer_reset()
    acc = ctx.organizer_insert(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
    org = exact_orc.Organizer(V, prm)
    o_acc = org.set_seeds(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
    assert np.array_equal(acc, o_acc)
    assert 0 < acc.sum() < len(acc)                   # collisions happened (3000 seeds, 1200 cells)
    _same_store(ctx, org, sc.n_views)
    # a second batch lands on a partly occupied grid
    acc2 = ctx.organizer_insert(seeds["pos"][::-1].copy(), seeds["nrm"][::-1].copy(),
                                seeds["ref"][::-1].copy(), nvis[::-1].copy(), vis[::-1].copy())
    o_acc2 = org.set_seeds(seeds["pos"][::-1], seeds["nrm"][::-1], seeds["ref"][::-1], nvis[::-1],
                           vis[::-1])
    assert np.array_equal(acc2, o_acc2)
    _same_store(ctx, org, sc.n_views)
    ctx.close()


@pytest.mark.parametrize("levels", [1, 3, -1])
def test_expansion_matches_fifo_oracle(capi_mod, exact_orc, levels):
    sc, seeds, ctx, V, prm, nvis, vis = _setup(capi_mod, exact_orc, 4, 160, 120, 40, 2)
    ctx.organizer_reset()
    ctx.organizer_insert(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
    org = exact_orc.Organizer(V, prm)
    org.set_seeds(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
    n0 = org.size()
    stats = ctx.expand(5, levels)
    pops = org.expand(5, levels)
    assert stats["pops"] == pops
    assert org.size() > n0 and stats["inserted"] == org.size() - n0
    _same_store(ctx, org, sc.n_views)
    ctx.close()


@pytest.mark.parametrize("m", [2, 3])
def test_several_patches_per_cell(capi_mod, exact_orc, m):
    """PatchOrganizerOptions::max_patches_per_cell > 1 (patch_organizer.h:40-47,
    patch_organizer.cpp:21): a cell takes the first m patches that ask for it in insertion
    order.  Seeds dense enough that cells fill up (grid_scale 8 on 160x120 = 300 cells per view),
    then two expansion levels: accept bits, grids (counts up to m) and store against the oracle."""
    from densepoints_b200 import scenes
    sc = scenes.make_plane_scene(seed=5, n_views=4, width=160, height=120, yaw_spread_deg=14.0)
    seeds = scenes.make_seeds(sc, 900, seed=6, depth_noise=0.004, tilt_deg=4.0)
    ctx = capi_mod.Context(0, capi_mod.default_params(minimum_visible_image=2, max_patches_per_cell=m))
    ctx.set_views(sc.P, sc.images)
    V = exact_orc.Views(sc.P, sc.images)
    prm = exact_orc.default_params(minimum_visible_image=2, max_patches_per_cell=m)
    nvis, vis, _, _ = exact_orc.visibility_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"])
    org = exact_orc.Organizer(V, prm)
    acc_o = org.set_seeds(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
    ctx.organizer_reset()
    acc = ctx.organizer_insert(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
    assert np.array_equal(acc, acc_o)
    grids = [ctx.organizer_grid(v) for v in range(sc.n_views)]
    assert max(int(g.max()) for g in grids) == m              # some cell did fill up
    assert 0 < acc.sum() < len(acc)                           # and some seeds were turned away
    _same_store(ctx, org, sc.n_views)
    stats = ctx.expand(5, 2)
    assert stats["pops"] == org.expand(5, 2)
    _same_store(ctx, org, sc.n_views)
    ctx.close()


def test_expansion_default_params_cell11(capi_mod, exact_orc):
    """Reference defaults: cell_size 11 (expand.h:12), minimum_visible_image 3."""
    sc, seeds, ctx, V, prm, nvis, vis = _setup(capi_mod, exact_orc, 6, 240, 180, 60, 3, seed=9)
    ctx.organizer_reset()
    ctx.organizer_insert(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
    org = exact_orc.Organizer(V, prm)
    org.set_seeds(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
    stats = ctx.expand(11, 2)
    assert stats["pops"] == org.expand(11, 2)
    _same_store(ctx, org, sc.n_views)
    ctx.close()


def test_max_pops_cap(capi_mod, exact_orc):
    """expand.cpp:95-97: the loop stops after max_pops pops."""
    sc, seeds, ctx, V, prm, nvis, vis = _setup(capi_mod, exact_orc, 4, 160, 120, 40, 2)
    p = ctx.get_params()
    p.max_pops = 25
    ctx.set_params(p)
    prm.max_pops = 25
    ctx.organizer_reset()
    ctx.organizer_insert(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
    org = exact_orc.Organizer(V, prm)
    org.set_seeds(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
    assert org.size() > 25
    stats = ctx.expand(5, -1)
    assert stats["pops"] == 25 == org.expand(5, -1)
    _same_store(ctx, org, sc.n_views)
    ctx.close()


def test_pyramid_matches_cv2_pyrdown(capi_mod, golden_primitives, exact_orc):
    g = golden_primitives
    ctx = capi_mod.Context(0)
    P = np.hstack([np.eye(3), np.zeros((3, 1))])
    ctx.set_views([P, P], [g["pyr_src"], g["pyr_src2"]])
    ctx.build_pyramid(3)
    assert np.array_equal(ctx.download_level(0, 0), g["pyr_src"])
    assert np.array_equal(ctx.download_level(0, 1), g["pyr_dst"])       # cv2.pyrDown golden
    assert np.array_equal(ctx.download_level(1, 1), g["pyr_dst2"])
    assert np.array_equal(ctx.download_level(0, 2), exact_orc.pyrdown(g["pyr_dst"]))
    ctx.close()


def test_scoring_on_pyramid_level(capi_mod, exact_orc):
    """Level l = the reference path on pyrDown^l images with P_l = diag(2^-l, 2^-l, 1) P."""
    from densepoints_b200 import scenes
    sc = scenes.make_plane_scene(seed=1, n_views=3, width=640, height=480)
    seeds = scenes.make_seeds(sc, 500, seed=1)
    ctx = capi_mod.Context(0, capi_mod.default_params(minimum_visible_image=2))
    ctx.set_views(sc.P, sc.images)
    ctx.build_pyramid(3)
    imgs, Ps = list(sc.images), sc.P.copy()
    for level in (1, 2):
        imgs = [exact_orc.pyrdown(im) for im in imgs]
        Ps = Ps.copy()
        Ps[:, :2, :] *= 0.5
        ctx.set_level(level)
        V = exact_orc.Views(Ps, imgs)
        nvis, vis, _, _ = ctx.visibility(seeds["pos"], seeds["nrm"], seeds["ref"])
        o = exact_orc.visibility_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"])
        assert np.array_equal(nvis, o[0]) and np.array_equal(vis, o[1])
        ncc, tex, valid = ctx.score(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis, 7,
                                    want_tex=True)
        o_ncc, o_tex, o_valid = exact_orc.score_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"],
                                                      nvis, vis, 7, want_tex=True)
        assert np.array_equal(valid, o_valid) and np.array_equal(tex, o_tex)
        assert np.abs(ncc - o_ncc).max() < 1e-6
    ctx.close()


def _zoomed_scene(zoom, n_seeds, width=480, height=360, seed=7):
    """A plane scene whose view i is delivered at `zoom[i]` times the resolution (pixel
    replication, P_i' = diag(z, z, 1) P_i): the same patch then covers 1, 2, 3 or 4 pixels per
    texel depending on the view, which is what per-(patch, view) level selection is for."""
    from densepoints_b200 import scenes
    sc = scenes.make_plane_scene(seed=seed, n_views=len(zoom), width=width, height=height,
                                 yaw_spread_deg=16.0)
    seeds = scenes.make_seeds(sc, n_seeds, seed=seed + 2, depth_noise=0.004, tilt_deg=5.0)
    Ps, imgs = [], []
    for P, im, z in zip(sc.P, sc.images, zoom):
        Pz = P.copy()
        Pz[:2, :] *= float(z)
        Ps.append(Pz)
        imgs.append(np.ascontiguousarray(np.repeat(np.repeat(im, z, axis=0), z, axis=1)))
    return np.array(Ps), imgs, seeds


def _level_views(orc, Ps, imgs, n_levels):
    out = [orc.Views(Ps, imgs)]
    for _ in range(1, n_levels):
        imgs = [orc.pyrdown(im) for im in imgs]            # pinned against cv2.pyrDown
        Ps = Ps.copy()
        Ps[:, :2, :] *= 0.5
        out.append(orc.Views(Ps, imgs))
    return out


@pytest.mark.parametrize("s,env", [(7, {"DP_LANE_MIN_PATCHES": "0"}), (7, {"DP_REFINE_KERNEL": "group", "DP_SCORE_KERNEL": "group"}),
                                   (5, {"DP_LANE_MIN_PATCHES": "0"}), (11, {}), (20, {})])
def test_per_view_level_selection(capi_mod, exact_orc, monkeypatch, s, env):
    """dp_set_level_selection: the texture of view v of a patch comes from the pyramid level its
    projected footprint asks for -- by definition the reference path (optimization.cpp:14-56) on
    pyrDown^k of that view with P_k = diag(2^-k, 2^-k, 1) P.  Textures, scores, the filter and
    the refinement (every kernel family: one patch per lane, four lanes per patch, one warp per
    patch) against the oracle, on a scene whose views deliver 1x, 2x, 3x and 4x the reference
    resolution so that levels 0, 1 and 2 all occur."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    zoom = [1, 2, 1, 4, 3, 2]
    Ps, imgs, seeds = _zoomed_scene(zoom, 900)
    ctx = capi_mod.Context(0, capi_mod.default_params(minimum_visible_image=2))
    ctx.set_views(Ps, imgs)
    ctx.build_pyramid(3)
    ctx.set_level_selection(True, 1.25)
    lv = _level_views(exact_orc, Ps, imgs, 3)
    V = lv[0]
    pos, nrm, ref = seeds["pos"], seeds["nrm"], seeds["ref"]
    nvis, vis, _, _ = ctx.visibility(pos, nrm, ref)           # on the base level
    o = exact_orc.visibility_batch(V, pos, nrm, ref)
    assert np.array_equal(nvis, o[0]) and np.array_equal(vis, o[1])
    exact_orc.set_level_selection(lv, 1.25)
    try:
        picked = exact_orc.levels_batch(V, pos, nrm, ref, nvis, vis, s)
        counts = [int((picked == l).sum()) for l in range(3)]
        assert min(counts) > 50, counts                       # the three levels really mix
        ncc, tex, valid = ctx.score(pos, nrm, ref, nvis, vis, s, want_tex=True)
        o_ncc, o_tex, o_valid = exact_orc.score_batch(V, pos, nrm, ref, nvis, vis, s, want_tex=True)
        assert np.array_equal(valid, o_valid) and np.array_equal(tex, o_tex)
        assert np.abs(ncc - o_ncc).max() < 1e-6
        assert valid.sum() > 0.5 * (nvis.sum())
        keep, fnv, fvi = ctx.filter(pos, nrm, ref, nvis, vis, s)
        o_keep, o_fnv, o_fvi = exact_orc.filter_batch(V, pos, nrm, ref, nvis, vis, s, min_visible=2)
        assert np.array_equal(keep, o_keep) and np.array_equal(fnv, o_fnv) and np.array_equal(fvi, o_fvi)
        m = np.where(keep.astype(bool))[0][:250]
        p1, n1, ev, _ = ctx.refine(pos[m], nrm[m], ref[m], fnv[m], fvi[m], s)
        o_p, o_n, o_ev, _ = exact_orc.refine_batch(V, pos[m], nrm[m], ref[m], fnv[m], fvi[m], s,
                                                   exact_orc.default_params(minimum_visible_image=2))
        assert np.array_equal(ev, o_ev) and np.array_equal(p1, o_p) and np.array_equal(n1, o_n)
        # switched off again: back to the base level for every view
        ctx.set_level_selection(False)
        exact_orc.set_level_selection(None)
        ncc0 = ctx.score(pos, nrm, ref, nvis, vis, s)
        assert np.abs(ncc0 - exact_orc.score_batch(V, pos, nrm, ref, nvis, vis, s)).max() < 1e-6
        assert np.abs(ncc0 - ncc).max() > 1e-3                # and it does make a difference
    finally:
        exact_orc.set_level_selection(None)
        ctx.close()


def test_expansion_with_level_selection(capi_mod, exact_orc):
    """The expansion loop (refine + filter of every candidate) under per-view level selection:
    store and grids against the 1-thread FIFO oracle."""
    zoom = [1, 2, 1, 3]
    Ps, imgs, seeds = _zoomed_scene(zoom, 120, width=320, height=240, seed=11)
    ctx = capi_mod.Context(0, capi_mod.default_params(minimum_visible_image=2))
    ctx.set_views(Ps, imgs)
    ctx.build_pyramid(3)
    ctx.set_level_selection(True, 1.5)
    lv = _level_views(exact_orc, Ps, imgs, 3)
    V = lv[0]
    prm = exact_orc.default_params(minimum_visible_image=2)
    nvis, vis, _, _ = exact_orc.visibility_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"])
    exact_orc.set_level_selection(lv, 1.5)
    try:
        org = exact_orc.Organizer(V, prm)
        acc_o = org.set_seeds(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
        ctx.organizer_reset()
        acc = ctx.organizer_insert(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis)
        assert np.array_equal(acc, acc_o)
        stats = ctx.expand(5, 2)
        assert stats["pops"] == org.expand(5, 2)
        assert ctx.organizer_size() > int(acc.sum())          # something was expanded
        _same_store(ctx, org, len(zoom))
    finally:
        exact_orc.set_level_selection(None)
        ctx.close()


def test_create_patches_and_ply_export(capi_mod, exact_orc, tmp_path):
    """Seed::CreatePatchesFromPoints on the device; PLY export of the store."""
    from densepoints_b200 import scenes
    sc = scenes.make_sphere_scene(seed=2, n_views=16, width=320, height=240, f=250.0)
    rng = np.random.default_rng(3)
    d = rng.normal(size=(5000, 3))
    pts = d / np.linalg.norm(d, axis=1, keepdims=True) * sc.radius * rng.uniform(0.97, 1.03, (5000, 1))
    ctx = capi_mod.Context(0)
    ctx.set_views(sc.P, sc.images)
    V = exact_orc.Views(sc.P, sc.images)
    got = ctx.create_patches(pts)
    want = exact_orc.create_patches(V, pts)
    for k in ("ref", "nvis", "vis", "pos", "nrm"):
        assert np.array_equal(got[k], want[k]), k
    assert len(set(got["ref"])) > 4 and got["nvis"].max() >= 3
    # store -> PLY in the reference's PrintCloud layout
    ctx.organizer_reset()
    acc = ctx.organizer_insert(got["pos"], got["nrm"], got["ref"], got["nvis"], got["vis"])
    path = str(tmp_path / "cloud.ply")
    ctx.export_ply(path)
    lines = open(path).read().split("\n")
    st = ctx.organizer_export()
    assert lines[0] == "ply" and lines[1] == "format ascii 1.0"
    assert lines[2] == f"element vertex {acc.sum()}" and lines[12] == "end_header"
    assert lines[3:12] == ["property float x", "property float y", "property float z",
                           "property uchar red", "property uchar green", "property uchar blue",
                           "property float nx", "property float ny", "property float nz"]
    rows = np.array([[float(t) for t in ln.split()] for ln in lines[13:13 + acc.sum()]])
    assert rows.shape == (acc.sum(), 9)
    assert np.allclose(rows[:, :3], st["pos"], rtol=1e-5) and np.allclose(rows[:, 6:], st["nrm"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(rows[:, 3:6].astype(np.uint8), st["rgb"])
    ctx.close()

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

"""CPU oracle vs. the reference's own KATs and the cv2-made golden vectors
(tests/golden/make_golden.py).  No GPU."""
import numpy as np


def test_reference_kat_ncc(orc):
    # tests/core/test_error_functions.cpp:9-15 (EXPECT_FLOAT_EQ = 4 ulp of float)
    a = np.array([[1, 2, 3], [-1, -2, -3], [1, 2, 3]], float)
    b = np.array([[2, 0, 5], [-4, 5, -2], [-1, 0, -3]], float)
    assert abs(np.float32(orc.ncc_f64(a, b)) - np.float32(0.1005653)) <= 4 * np.spacing(np.float32(0.1))
    assert abs(np.float32(orc.ncc_f64(a, a)) - np.float32(1.0)) <= 4 * np.spacing(np.float32(1.0))
    assert orc.ncc_bgr(None, np.zeros(75, np.uint8)) == -1       # error_measurements.cpp:38-40


def test_reference_kat_projection_decomposition(orc):
    # tests/core/test_projection_matrix_decomposition.cpp:10-36
    P = np.array([[3.53553e2, 3.39645e2, 2.77744e2, -1.44946e6],
                  [-1.03528e2, 2.33212e1, 4.59607e2, -6.32525e5],
                  [7.07107e-1, -3.53553e-1, 6.12372e-1, -9.18559e2]])
    K, R, c = orc.view_decompose(P)
    assert abs(K[0, 0] - 468.2) < 0.1 and abs(K[1, 1] - 427.2) < 0.1
    assert abs(K[0, 2] - 300) < 0.1 and abs(K[1, 2] - 200) < 0.1 and abs(K[2, 2] - 1) < 0.1
    E = np.hstack([R, (-R @ c)[:, None]])
    rec = K @ E * np.linalg.norm(P[2, :3])
    assert np.abs(rec - P).max() < 0.5
    assert np.abs(c - [1000, 2000, 1500]).max() < 0.01
    assert abs(np.linalg.det(R) - 1) < 1e-9 and np.abs(R @ R.T - np.eye(3)).max() < 1e-12


def test_golden_homography_and_warp(orc, golden_primitives):
    g = golden_primitives
    img = g["image"]
    for quad, s, roi, H, tex in zip(g["quad"], g["s"], g["roi"], g["H"], g["tex"]):
        s = int(s)
        cell = np.array([[0, 0], [s, 0], [s, s], [0, s]], np.float32)
        Ho = orc.find_homography4(quad, cell)
        assert Ho is not None
        assert np.abs(Ho - H).max() <= 1e-9 * max(1.0, np.abs(H).max())
        x0, y0, w, h = (int(v) for v in roi)
        sub = img[y0:y0 + h, x0:x0 + w]
        assert np.array_equal(orc.warp_perspective(sub, Ho, s), tex[:s, :s])       # own H
        assert np.array_equal(orc.warp_perspective(sub, H, s), tex[:s, :s])        # cv2's H


def test_golden_homography_and_warp_wide(orc, golden_primitives_wide):
    """cv2 vectors for the cell sizes / ROI shapes the first set does not hold: s in
    {2,3,4,6,8,9,13,20,32}, ROIs up to 48 px, keystone quads, quads mostly outside their ROI
    (BORDER_REPLICATE on most texels), 1-pixel ROIs (tests/golden/make_golden_wide.py)."""
    g = golden_primitives_wide
    img = g["image"]
    assert set(int(k) for k in g["kind"]) == {0, 1, 2, 3} and len(g["s"]) >= 300
    for quad, s, roi, H, tex in zip(g["quad"], g["s"], g["roi"], g["H"], g["tex"]):
        s = int(s)
        cell = np.array([[0, 0], [s, 0], [s, s], [0, s]], np.float32)
        Ho = orc.find_homography4(quad, cell)
        assert Ho is not None
        assert np.abs(Ho - H).max() <= 1e-9 * max(1.0, np.abs(H).max())
        x0, y0, w, h = (int(v) for v in roi)
        sub = img[y0:y0 + h, x0:x0 + w]
        assert np.array_equal(orc.warp_perspective(sub, Ho, s), tex[:s, :s])
        assert np.array_equal(orc.warp_perspective(sub, H, s), tex[:s, :s])


def test_golden_gray(orc, golden_primitives):
    g = golden_primitives
    mine = np.array([orc.gray(*px) for px in g["gray_px"]])
    assert np.array_equal(mine, g["gray"])


def test_golden_pyrdown(orc, golden_primitives):
    g = golden_primitives
    assert np.array_equal(orc.pyrdown(g["pyr_src"]), g["pyr_dst"])      # odd sizes, random noise
    assert np.array_equal(orc.pyrdown(g["pyr_src2"]), g["pyr_dst2"])    # a rendered scene image


def test_golden_textures_ncc_filter(orc, golden_scoring, golden_views):
    g = golden_scoring
    for s in (5, 7, 11, 16):
        ncc, tex, valid = orc.score_batch(golden_views, g["pos"], g["nrm"], g["ref"], g["nvis"],
                                          g["vis"], s, want_tex=True)
        assert np.array_equal(valid, g[f"valid{s}"])
        m = g[f"valid{s}"].astype(bool)
        assert np.array_equal(tex[m], g[f"tex{s}"][m])                       # bit-exact u8
        k = np.arange(g["vis"].shape[1])[None, :]
        sm = (k >= 1) & (k < g["nvis"][:, None])
        assert np.abs(ncc[sm] - g[f"ncc{s}"][sm]).max() < 2e-6               # bar: 1e-4
        keep, fnvis, fvis = orc.filter_batch(golden_views, g["pos"], g["nrm"], g["ref"],
                                             g["nvis"], g["vis"], s, 0.6, 2)
        assert np.array_equal(keep, g[f"keep{s}"])
        assert np.array_equal(fnvis, g[f"fnvis{s}"])
        assert np.array_equal(fvis, g[f"fvis{s}"])
    # the golden decomposition (numpy SVD + QR, like the reference's Eigen calls)
    for v in range(golden_views.n):
        assert np.abs(golden_views.center(v) - g["center"][v]).max() < 1e-9
        assert np.abs(golden_views.xaxis(v) - g["xaxis"][v]).max() < 1e-12


def test_golden_sphere_textures_ncc_filter(orc, golden_scoring_sphere):
    """The same chain against cv2 on a curved scene (sphere, 6 views, patches tilted up to 30
    degrees, default minimum_visible_image = 3) at s = 3, 8, 13, 20
    (tests/golden/make_golden_sphere.py); at s = 3 most ROIs shrink to nothing (empty textures)."""
    g = golden_scoring_sphere
    V = orc.Views(g["P"], list(g["images"]))
    for s in (3, 8, 13, 20):
        ncc, tex, valid = orc.score_batch(V, g["pos"], g["nrm"], g["ref"], g["nvis"], g["vis"], s,
                                          want_tex=True)
        assert np.array_equal(valid, g[f"valid{s}"])
        m = g[f"valid{s}"].astype(bool)
        assert np.array_equal(tex[m], g[f"tex{s}"][m])
        k = np.arange(g["vis"].shape[1])[None, :]
        sm = (k >= 1) & (k < g["nvis"][:, None])
        assert np.abs(ncc[sm] - g[f"ncc{s}"][sm]).max() < 2e-6
        keep, fnvis, fvis = orc.filter_batch(V, g["pos"], g["nrm"], g["ref"], g["nvis"], g["vis"],
                                             s, 0.6, 3)
        assert np.array_equal(keep, g[f"keep{s}"])
        assert np.array_equal(fnvis, g[f"fnvis{s}"])
        assert np.array_equal(fvis, g[f"fvis{s}"])
    assert g["valid3"].sum() < 0.75 * g["valid8"].sum()      # the empty-texture path is exercised


def test_golden_refinement_objective(orc, golden_scoring, golden_views):
    """PatchOptimizationOpenCVFunctor::calc at 16 (depth, roll, pitch) points per patch against
    cv2 (tests/golden/make_golden_objective.py, an independent restatement of
    optimization_opencv.cpp:14-39 / optimization.cpp:14-56,78-96 / patch.cpp:111-164).  The
    pure-depth points pin that the corners are built around the STORED position
    (patch.cpp:119-123) and the trial position only scales the quad: with the corners around
    the trial position instead, these vectors fail at x[4:8]."""
    import os
    from conftest import GOLDEN
    go = dict(np.load(os.path.join(GOLDEN, "golden_objective.npz")))
    g = golden_scoring
    orc.set_homography_mode(0)
    idx = go["patch"]
    ties = 0
    for s in (5, 7, 11):
        for b, x in enumerate(go["x"]):
            # the textures at trial parameters: GetProjectedTextures(normal, position, ...)
            tn, tp = zip(*(orc.unparametrize(golden_views, g["ref"][i], g["nrm"][i], g["pos"][i], x)
                           for i in idx))
            _, tex, valid = orc.score_batch(golden_views, g["pos"][idx], g["nrm"][idx],
                                            g["ref"][idx], g["nvis"][idx], g["vis"][idx], s,
                                            want_tex=True, trial_nrm=np.array(tn),
                                            trial_pos=np.array(tp))
            assert np.array_equal(valid, go[f"valid_{s}"][:, b])
            diff = (tex != go[f"tex_{s}"][:, b]) & valid.astype(bool)[:, :, None, None, None]
            # only texel (0,0) may differ: an exact tie of the 1/32-px rounding, decided inside
            # OpenCV by the sign of its eigen-solver's noise
            assert diff[:, :, 1:].sum() == 0 and diff[:, :, 0, 1:].sum() == 0
            clean = diff.sum(axis=(1, 2, 3, 4)) == 0
            ties += int((~clean).sum())
            for a, i in enumerate(idx):
                if clean[a]:
                    f = orc.objective(golden_views, g["ref"][i], g["vis"][i, :g["nvis"][i]], s,
                                      g["nrm"][i], g["pos"][i], x)
                    assert abs(f - go[f"f_{s}"][a, b]) < 1e-12, (s, a, b)
    assert ties <= 3, ties          # 1 of 3072 (patch, point, cell size) cases today

"""CUDA vs the cv2-made golden vectors of the curved scene (tests/golden/make_golden_sphere.py):
sphere, 6 views, patches tilted up to 30 degrees, s = 3, 8, 13, 20, default parameters."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_golden_sphere_textures_ncc_filter(golden_scoring_sphere):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from densepoints_b200 import build as b
    b.build_cuda()
    from densepoints_b200 import capi
    g = golden_scoring_sphere
    ctx = capi.Context(0)                      # minimum_visible_image = 3, threshold 0.6
    ctx.set_views(g["P"], list(g["images"]), xaxes=g["xaxis"], centers=g["center"])
    for s in (3, 8, 13, 20):
        ncc, tex, valid = ctx.score(g["pos"], g["nrm"], g["ref"], g["nvis"], g["vis"], s,
                                    want_tex=True)
        assert np.array_equal(valid, g[f"valid{s}"])
        diff = (tex != g[f"tex{s}"]).any(axis=-1)
        # only an exact 1/64-px tie of texel (0,0) may differ from OpenCV's noise-dependent result
        assert diff[..., 1:, :].sum() == 0 and diff[..., 0, 1:].sum() == 0
        assert diff.sum() <= 3
        k = np.arange(g["vis"].shape[1])[None, :]
        sm = (k >= 1) & (k < g["nvis"][:, None])
        clean = ~diff.any(axis=(2, 3))
        clean = clean & clean[:, :1]
        assert np.abs(ncc[sm & clean] - g[f"ncc{s}"][sm & clean]).max() < 5e-6   # bar 1e-4
        # the filter for every patch whose textures are all clean (at most 3 are not)
        pclean = ~diff.any(axis=(1, 2, 3))
        assert pclean.sum() >= len(pclean) - 3
        keep, fnvis, fvis = ctx.filter(g["pos"], g["nrm"], g["ref"], g["nvis"], g["vis"], s)
        assert np.array_equal(keep[pclean], g[f"keep{s}"][pclean])
        assert np.array_equal(fnvis[pclean], g[f"fnvis{s}"][pclean])
        assert np.array_equal(fvis[pclean], g[f"fvis{s}"][pclean])
    ctx.close()

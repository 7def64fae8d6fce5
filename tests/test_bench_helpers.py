"""CPU checks of bench.py's host-side helpers (no GPU work)."""
import numpy as np


def test_b_alg_matches_survey_values():
    """B_alg(s, #V) = 3 (2 floor(s/2) + 2)^2 + 4 + (28 + 2 #V) / #V."""
    import bench
    assert bench.b_alg(5, 2) == 108 + 4 + 16
    assert abs(bench.b_alg(7, 8) - (192 + 4 + 5.5)) < 1e-12
    assert abs(bench.b_alg(11, 8) - (432 + 4 + 5.5)) < 1e-12


def test_c3_patches_follow_the_seed_rules():
    """The C3 leg's device-side patch generator, run on the CPU device: reference = nearest camera,
    eight other views in ascending order, points on the cap the cameras face, unit normals."""
    import torch
    import bench
    from densepoints_b200 import scenes
    sc = scenes.make_sphere_scene(seed=2, n_views=16, width=64, height=48, f=50.0)
    pos, nrm, ref, vis = bench.c3_patches(sc.centers, sc.radius, 5000, 3000, torch.device("cpu"))
    pos, nrm, ref, vis = (t.numpy() for t in (pos, nrm, ref, vis))
    assert pos.shape == (5000, 3) and vis.shape == (5000, 8) and ref.shape == (5000,)
    d = np.linalg.norm(pos[:, None, :].astype(np.float64) - sc.centers[None], axis=2)
    assert np.array_equal(ref, d.argmin(1))
    assert (np.diff(vis, axis=1) > 0).all() and vis.min() >= 0 and vis.max() < 16
    assert not (vis == ref[:, None]).any()
    assert np.abs(np.linalg.norm(nrm, axis=1) - 1).max() < 1e-5
    r = np.linalg.norm(pos, axis=1)
    assert np.abs(r / sc.radius - 1).max() <= 0.0101
    mean_dir = sc.centers.mean(0) / np.linalg.norm(sc.centers.mean(0))
    assert ((pos / r[:, None]) @ mean_dir > 0.79).all()
    assert ((nrm * pos).sum(1) < 0).all()                     # inward normals
    # deterministic
    again = bench.c3_patches(sc.centers, sc.radius, 5000, 3000, torch.device("cpu"))
    assert np.array_equal(again[0].numpy(), pos) and np.array_equal(again[3].numpy(), vis)

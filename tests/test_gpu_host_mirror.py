"""The C++ host mirror of the reference's plugin interface (densepoints_b200/host:
View, Patch, Optimization / OptimizationCUDA, the Seed batch drivers, Expand) driven by a
C++ program the way methods/pmvs drives its own classes, checked against the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_driver(tmp, name="host_mirror_test"):
    from densepoints_b200 import build as b
    lib = b.build_cuda()
    exe = os.path.join(tmp, name)
    cmd = ["/usr/bin/g++", "-std=c++14", "-O2", "-Wall", "-I" + os.path.join(ROOT, "densepoints_b200", "host"),
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
           "-o", exe, "-L" + os.path.dirname(lib), "-ldensepoints_cuda",
           "-Wl,-rpath," + os.path.dirname(lib)]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run(cmd, check=True, env=env)
    return exe


def test_host_mirror_compiles(tmp_path):
    """CPU: the mirror headers + driver compile and link against the C ABI."""
    assert os.path.exists(_build_driver(str(tmp_path)))


def test_pmvs_facade_compiles(tmp_path):
    """CPU: the mirrored method facade (PMVS::AddCamera / Run / GetPointCloud, reference
    pmvs.h:14-35) driven like programs/densify compiles, links and starts."""
    exe = _build_driver(str(tmp_path), "pmvs_facade_check")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "usage" in r.stdout


def _read_patches(f, n_views):
    n = struct.unpack("<i", f.read(4))[0]
    out = dict(pos=np.zeros((n, 3), np.float32), nrm=np.zeros((n, 3), np.float32),
               rgb=np.zeros((n, 3), np.uint8), ref=np.zeros(n, np.int32), nvis=np.zeros(n, np.int32),
               vis=np.full((n, n_views), -1, np.int32))
    for i in range(n):
        g = np.frombuffer(f.read(24), np.float32)
        out["pos"][i], out["nrm"][i] = g[:3], g[3:]
        out["rgb"][i] = np.frombuffer(f.read(3), np.uint8)
        out["ref"][i], out["nvis"][i] = struct.unpack("<ii", f.read(8))
        out["vis"][i] = np.frombuffer(f.read(4 * n_views), np.int32)
    return out


@pytest.mark.gpu
def test_host_mirror_pipeline_matches_oracle(tmp_path, orc):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from densepoints_b200 import scenes
    exe = _build_driver(str(tmp_path))
    CELL_SEED, CELL_EXP, MIN_VIS, LEVELS = 7, 5, 2, 2
    sc = scenes.make_plane_scene(seed=5, n_views=4, width=160, height=120, yaw_spread_deg=14.0)
    seeds = scenes.make_seeds(sc, 60, seed=6, depth_noise=0.004, tilt_deg=4.0)
    fin, fout = str(tmp_path / "scene.bin"), str(tmp_path / "result.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<iii", sc.n_views, sc.width, sc.height))
        for P, im in zip(sc.P, sc.images):
            f.write(np.ascontiguousarray(P, np.float64).tobytes())
            f.write(np.ascontiguousarray(im, np.uint8).tobytes())
        f.write(struct.pack("<i", len(seeds["ref"])))
        f.write(seeds["pos"].astype(np.float32).tobytes())
        f.write(seeds["nrm"].astype(np.float32).tobytes())
        f.write(seeds["ref"].astype(np.int32).tobytes())
        f.write(struct.pack("<iiii", CELL_SEED, CELL_EXP, MIN_VIS, LEVELS))
    fply = str(tmp_path / "cloud.ply")
    r = subprocess.run([exe, fin, fout, fply], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "host mirror ok" in r.stdout

    # ---- the same pipeline with the oracle (reference call order) ----------------------------
    orc.set_homography_mode(1)
    try:
        V = orc.Views(sc.P, sc.images)
        prm = orc.default_params(minimum_visible_image=MIN_VIS)
        nvis, vis, _, _ = orc.visibility_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"])
        _, tex0, valid0 = orc.score_batch(V, seeds["pos"][:1], seeds["nrm"][:1], seeds["ref"][:1],
                                          nvis[:1], vis[:1], CELL_SEED, want_tex=True)
        k1, fn1, fv1 = orc.filter_batch(V, seeds["pos"][:1], seeds["nrm"][:1], seeds["ref"][:1],
                                        nvis[:1], vis[:1], CELL_SEED, 0.6, MIN_VIS)
        p1, n1, _, _ = orc.refine_batch(V, seeds["pos"][:1], seeds["nrm"][:1], seeds["ref"][:1],
                                        fn1, fv1, CELL_SEED, prm)
        keep, fnvis, fvis = orc.filter_batch(V, seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis,
                                             CELL_SEED, 0.6, MIN_VIS)
        m = keep.astype(bool)
        rp, rn, _, _ = orc.refine_batch(V, seeds["pos"][m], seeds["nrm"][m], seeds["ref"][m],
                                        fnvis[m], fvis[m], CELL_SEED, prm)
        org = orc.Organizer(V, prm)
        org.set_seeds(rp, rn, seeds["ref"][m], fnvis[m], fvis[m])
        org.expand(CELL_EXP, LEVELS)
        want = org.export()
    finally:
        orc.set_homography_mode(0)

    with open(fout, "rb") as f:
        nt = struct.unpack("<i", f.read(4))[0]
        assert nt == nvis[0]
        for k in range(nt):                                   # GetProjectedTextures
            sz = struct.unpack("<i", f.read(4))[0]
            assert (sz != 0) == bool(valid0[0, k])
            if sz:
                t = np.frombuffer(f.read(sz * sz * 3), np.uint8).reshape(sz, sz, 3)
                assert np.array_equal(t, tex0[0, k])
        assert struct.unpack("<i", f.read(4))[0] == int(k1[0])   # FilterByErrorMeasurement
        one = _read_patches(f, sc.n_views)                       # ... then Optimize
        assert np.array_equal(one["pos"][0], p1[0]) and np.array_equal(one["nrm"][0], n1[0])
        assert one["nvis"][0] == fn1[0] and np.array_equal(one["vis"][0, :fn1[0]], fv1[0, :fn1[0]])
        ref = _read_patches(f, sc.n_views)                       # Seed::OptimizeAndRefinePatches
        assert np.array_equal(ref["pos"], rp) and np.array_equal(ref["nrm"], rn)
        assert np.array_equal(ref["ref"], seeds["ref"][m]) and np.array_equal(ref["nvis"], fnvis[m])
        assert np.array_equal(ref["vis"], fvis[m])
        exp = _read_patches(f, sc.n_views)                       # Expand::SetSeeds
        for k in ("pos", "nrm", "rgb", "ref", "nvis", "vis"):
            assert np.array_equal(exp[k], want[k]), k
        assert len(want["ref"]) > m.sum() > 0
        f.read(32)                                               # Expand::Stats
        cre = _read_patches(f, sc.n_views)                       # Seed::CreatePatchesFromPoints
    orc.set_homography_mode(0)
    oc = orc.create_patches(orc.Views(sc.P, sc.images), seeds["pos"].astype(np.float64))
    for k in ("pos", "nrm", "ref", "nvis", "vis"):
        assert np.array_equal(cre[k], oc[k]), k
    ply = open(fply).read().split("\n")
    assert ply[0] == "ply" and ply[2] == f"element vertex {len(want['ref'])}"

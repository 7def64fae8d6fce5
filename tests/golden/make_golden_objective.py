"""Golden vectors for the refinement objective, made with the REAL OpenCV primitives (cv2 4.13.0)
driven by an independent Python restatement of the reference's glue:

  PatchOptimizationOpenCVFunctor::calc      optimization_opencv.cpp:14-39
  Optimization::UnparametrizePatch          optimization.cpp:78-96
  Optimization::GetProjectedTextures(normal, position, textures)   optimization.cpp:14-56
  Patch::ComputePatchToViewHomography       patch.cpp:111-164

What this pins: the TRIAL (normal, position) only feed GetProjectedXYAxisAndScale
(optimization.cpp:24-26: axes and dx); the four corners are built around the patch's STORED
position, GetPosition() (patch.cpp:119-123).  A trial depth therefore changes the scale of the
quad, never its centre.  The scene and patches are the ones of golden_scoring.npz.

Run from the repo root:  python tests/golden/make_golden_objective.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import inside, ncc_cv2, project  # noqa: E402

cv2.setNumThreads(1)


def unparametrize(center, nrm0, pos0, depth, roll, pitch):
    """optimization.cpp:78-96 (nrm0 / pos0 are the fp32-stored values widened to double)."""
    pos = center + (1 + depth) * (pos0 - center)
    ca, sa, cb, sb = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch)
    R = np.array([[cb, 0, -sb], [sa * sb, ca, cb * sa], [ca * sb, -sa, ca * cb]])
    return R @ nrm0, pos


def textures_at(Ps, xaxes, images, ref, vis, s, normal, position, stored_position):
    """optimization.cpp:14-56 with patch.cpp:111-164 inlined; cv2 does the OpenCV work."""
    xa = xaxes[ref] / np.linalg.norm(xaxes[ref])                 # patch.cpp:95
    ya = np.cross(normal, xa)                                    # patch.cpp:96
    dx = np.linalg.norm(project(Ps[ref], position + xa) - project(Ps[ref], position))
    scale = (s // 2) / dx                                        # optimization.cpp:30
    ax, ay = scale * xa, scale * ya
    c = stored_position                                          # GetPosition(), patch.cpp:120-123
    out = []
    for v in vis:
        img = images[v]
        h, w = img.shape[:2]
        corners = [c - ax - ay, c + ax - ay, c + ax + ay, c - ax + ay]
        tl, br, pts, ok = [w, h], [0, 0], [], True
        for Xc in corners:
            if not inside(Ps[v], Xc, w, h):
                ok = False
                break
            p = project(Ps[v], Xc)
            pts.append([np.float32(p[0]), np.float32(p[1])])
            tl = [min(tl[0], int(np.ceil(p[0]))), min(tl[1], int(np.ceil(p[1])))]
            br = [max(br[0], int(np.floor(p[0]))), max(br[1], int(np.floor(p[1])))]
        if not ok:
            out.append(None)
            continue
        roi = (tl[0], tl[1], br[0] - tl[0], br[1] - tl[1])
        if roi[2] <= 0 or roi[3] <= 0:
            out.append(None)
            continue
        pts = np.array(pts, np.float32)
        pts[:, 0] -= np.float32(roi[0])
        pts[:, 1] -= np.float32(roi[1])
        cell = np.array([[0, 0], [s, 0], [s, s], [0, s]], np.float32)
        H, _ = cv2.findHomography(pts, cell, 0)
        if H is None:
            out.append(None)
            continue
        sub = img[roi[1]:roi[1] + roi[3], roi[0]:roi[0] + roi[2]]
        out.append(cv2.warpPerspective(sub, H, (s, s), flags=cv2.INTER_LINEAR,
                                       borderMode=cv2.BORDER_REPLICATE))
    return out


def calc(Ps, xaxes, centers, images, ref, vis, s, nrm0, pos0, x):
    """optimization_opencv.cpp:14-39"""
    normal, position = unparametrize(centers[ref], nrm0, pos0, *x)
    tex = textures_at(Ps, xaxes, images, ref, vis, s, normal, position, pos0)
    scores = [1.0 - ncc_cv2(tex[0], tex[k]) for k in range(1, len(tex))]
    if not scores:
        return 2.0, tex
    total = 0.0
    for sc in scores:                       # std::accumulate, in order
        total += sc
    return total / len(scores), tex


def main():
    g = dict(np.load(os.path.join(HERE, "golden_scoring.npz")))
    Ps, xaxes, centers, images = g["P"], g["xaxis"], g["center"], list(g["images"])
    rng = np.random.default_rng(20261019)
    pick = [i for i in range(len(g["ref"])) if g["nvis"][i] >= 2][12:76]     # 64 patches
    # (depth, roll, pitch): the initial simplex of Optimize(), pure-depth points (the case that
    # separates "corners around the stored position" from "corners around the trial position"),
    # and random points of the size Nelder-Mead visits
    xs = [(-0.01, -0.1, -0.1), (0.01, 0.0, 0.0), (0.0, 0.1, 0.0), (0.0, 0.0, 0.1),
          (0.03, 0.0, 0.0), (-0.05, 0.0, 0.0), (0.08, 0.0, 0.0), (-0.12, 0.0, 0.0)]
    xs += [tuple(rng.uniform(-1, 1, 3) * (0.06, 0.3, 0.3)) for _ in range(8)]
    xs = np.array(xs, np.float64)
    out = dict(patch=np.array(pick, np.int32), x=xs)
    for s in (5, 7, 11):
        f = np.zeros((len(pick), len(xs)))
        tex = np.zeros((len(pick), len(xs), Ps.shape[0], s, s, 3), np.uint8)
        valid = np.zeros((len(pick), len(xs), Ps.shape[0]), np.uint8)
        for a, i in enumerate(pick):
            vis = [int(v) for v in g["vis"][i, :g["nvis"][i]]]
            for b, x in enumerate(xs):
                f[a, b], t = calc(Ps, xaxes, centers, images, int(g["ref"][i]), vis, s,
                                  g["nrm"][i].astype(np.float64), g["pos"][i].astype(np.float64), x)
                for k, tk in enumerate(t):
                    if tk is not None:
                        tex[a, b, k] = tk
                        valid[a, b, k] = 1
        out[f"f_{s}"] = f
        out[f"tex_{s}"] = tex
        out[f"valid_{s}"] = valid
        print(s, "objective range", f.min(), f.max(), "valid textures", int(valid.sum()))
    np.savez_compressed(os.path.join(HERE, "golden_objective.npz"), **out)


if __name__ == "__main__":
    main()

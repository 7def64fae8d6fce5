"""Generates the golden fixtures under tests/golden/ with the REAL OpenCV
primitives the reference calls (Python cv2 4.13.0 in this image):
cv2.findHomography / cv2.warpPerspective / cv2.cvtColor / cv2.meanStdDev /
cv2.pyrDown, driven in the reference's order (optimization.cpp:14-56,
patch.cpp:111-164, error_measurements.cpp:36-60, optimization.cpp:98-132).

The reference C++ cannot be built here (no OpenCV/Eigen/PCL headers),
so these vectors are what pins the CPU oracle's restatement of the OpenCV
arithmetic.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from densepoints_b200 import scenes  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
cv2.setNumThreads(1)


def decompose(P):
    """View::SetProjectionMatrix (types.cpp:28-68) with numpy: centre = SVD null
    vector, R = RQ of P[:, :3] with positive diagonal of K."""
    _, _, Vt = np.linalg.svd(P)
    c = Vt[-1]
    center = c[:3] / c[3]
    M = P[:, :3]
    Jm = np.flipud(np.eye(3))
    Q, R = np.linalg.qr((Jm @ M).T)
    K = Jm @ R.T @ Jm
    Rot = Jm @ Q.T
    S = np.diag(np.sign(np.diag(K)))
    return K @ S, S @ Rot, center


def project(P, X):
    h = P @ np.append(X, 1.0)
    return h[:2] / h[2]


def inside(P, X, w, h):
    u, v = project(P, X)
    return (u > 0) and (u < w) and (v > 0) and (v < h)


def textures_cv2(Ps, xaxes, images, ref, vis, s, nrm, pos):
    """Optimization::GetProjectedTextures with cv2 doing the OpenCV work."""
    nrm = nrm.astype(np.float64)
    pos = pos.astype(np.float64)
    xa = xaxes[ref] / np.linalg.norm(xaxes[ref])
    ya = np.cross(nrm, xa)
    dx = np.linalg.norm(project(Ps[ref], pos + xa) - project(Ps[ref], pos))
    scale = (s // 2) / dx
    ax, ay = scale * xa, scale * ya
    out, rois, Hs = [], [], []
    for v in vis:
        img = images[v]
        h, w = img.shape[:2]
        corners = [pos - ax - ay, pos + ax - ay, pos + ax + ay, pos - ax + ay]
        tl = [w, h]
        br = [0, 0]
        pts = []
        ok = True
        for Xc in corners:
            if not inside(Ps[v], Xc, w, h):
                ok = False
                break
            p = project(Ps[v], Xc)
            pts.append([np.float32(p[0]), np.float32(p[1])])
            tl[0] = min(tl[0], int(np.ceil(p[0])))
            tl[1] = min(tl[1], int(np.ceil(p[1])))
            br[0] = max(br[0], int(np.floor(p[0])))
            br[1] = max(br[1], int(np.floor(p[1])))
        if not ok:
            out.append(None); rois.append((0, 0, 0, 0)); Hs.append(np.zeros((3, 3)))
            continue
        roi = (tl[0], tl[1], br[0] - tl[0], br[1] - tl[1])
        pts = np.array(pts, np.float32)
        pts[:, 0] -= np.float32(roi[0])
        pts[:, 1] -= np.float32(roi[1])
        cell = np.array([[0, 0], [s, 0], [s, s], [0, s]], np.float32)
        H, _ = cv2.findHomography(pts, cell, 0)
        if H is None or roi[2] <= 0 or roi[3] <= 0:
            out.append(None); rois.append(roi); Hs.append(np.zeros((3, 3)))
            continue
        sub = img[roi[1]:roi[1] + roi[3], roi[0]:roi[0] + roi[2]]     # image(roi): a view
        tex = cv2.warpPerspective(sub, H, (s, s), flags=cv2.INTER_LINEAR,
                                  borderMode=cv2.BORDER_REPLICATE)
        out.append(tex); rois.append(roi); Hs.append(H)
    return out, rois, Hs


def ncc_cv2(ta, tb):
    """NCCScore (error_measurements.cpp:36-60) with cv2 primitives."""
    if ta is None or tb is None:
        return -1.0
    fa = cv2.cvtColor(ta, cv2.COLOR_BGR2GRAY).astype(np.float32)
    fb = cv2.cvtColor(tb, cv2.COLOR_BGR2GRAY).astype(np.float32)
    ma, sa = cv2.meanStdDev(fa)
    mb, sb = cv2.meanStdDev(fb)
    da = fa - np.float32(ma[0, 0])
    db = fb - np.float32(mb[0, 0])
    num = float(np.dot(da.ravel().astype(np.float64), db.ravel().astype(np.float64)))
    den = max(1e-1, float(sa[0, 0] * sb[0, 0]))
    return num / den / fa.size


def filter_ref(scores, vis, thr, min_vis):
    """FilterByErrorMeasurement erase loop incl. the off-by-one (optimization.cpp:112-131)."""
    vis = list(vis)
    if len(scores) == 0:
        return False, vis
    removed = 0
    for i, sc in enumerate(scores):
        if sc < thr:
            del vis[i - removed]
            removed += 1
    return len(vis) >= min_vis, vis


def main():
    rng = np.random.default_rng(20261018)
    # ---- primitives: random quads on random sub-views ----------------------
    img = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    prim = dict(image=img, quad=[], s=[], roi=[], H=[], tex=[])
    for it in range(240):
        s = int(rng.choice([5, 7, 11, 16]))
        w = int(rng.integers(1, 18)); h = int(rng.integers(1, 18))
        x0 = int(rng.integers(0, 128 - w)); y0 = int(rng.integers(0, 96 - h))
        quad = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float32) + \
            rng.uniform(-0.95, 0.95, (4, 2)).astype(np.float32)
        if it % 3 == 0:
            quad = quad + rng.uniform(-3, 3, (4, 2)).astype(np.float32)
        cell = np.array([[0, 0], [s, 0], [s, s], [0, s]], np.float32)
        H, _ = cv2.findHomography(quad, cell, 0)
        tex = cv2.warpPerspective(img[y0:y0 + h, x0:x0 + w], H, (s, s), flags=cv2.INTER_LINEAR,
                                  borderMode=cv2.BORDER_REPLICATE)
        pad = np.zeros((16, 16, 3), np.uint8)
        pad[:s, :s] = tex
        prim["quad"].append(quad); prim["s"].append(s); prim["roi"].append([x0, y0, w, h])
        prim["H"].append(H); prim["tex"].append(pad)
    px = rng.integers(0, 256, (4096, 1, 3), dtype=np.uint8)
    gray = cv2.cvtColor(px, cv2.COLOR_BGR2GRAY)
    pyr_src = rng.integers(0, 256, (61, 83, 3), dtype=np.uint8)
    pyr = cv2.pyrDown(pyr_src)
    pyr2 = cv2.pyrDown(scenes.make_plane_scene(seed=7, width=96, height=64).images[0])
    np.savez_compressed(
        os.path.join(OUT, "golden_primitives.npz"), image=img, quad=np.array(prim["quad"]),
        s=np.array(prim["s"], np.int32), roi=np.array(prim["roi"], np.int32),
        H=np.array(prim["H"]), tex=np.array(prim["tex"]), gray_px=px.reshape(-1, 3),
        gray=gray.reshape(-1), pyr_src=pyr_src, pyr_dst=pyr,
        pyr_src2=scenes.make_plane_scene(seed=7, width=96, height=64).images[0], pyr_dst2=pyr2)

    # ---- scoring + filter on a small plane scene ----------------------------
    sc = scenes.make_plane_scene(seed=11, n_views=4, width=160, height=120, yaw_spread_deg=18.0,
                                 name="golden-plane")
    seeds = scenes.make_seeds(sc, 160, seed=12, depth_noise=0.03, tilt_deg=25.0)
    # push a few seeds to the image border so some corners fall outside
    seeds["pos"][:12, 0] = np.linspace(-9.8, -8.6, 12).astype(np.float32)
    Ps = sc.P
    dec = [decompose(P) for P in Ps]
    xaxes = [d[1][0] for d in dec]
    centers = [d[2] for d in dec]
    n = seeds["pos"].shape[0]
    V = sc.n_views
    # visible = all views but the reference whose centre projects inside (no angle test here:
    # keeps oblique views in the golden); ascending view id (patch.cpp:29-30)
    vis = np.full((n, V), -1, np.int32)
    nvis = np.zeros(n, np.int32)
    for i in range(n):
        k = 0
        for v in range(V):
            if v != seeds["ref"][i] and inside(Ps[v], seeds["pos"][i].astype(np.float64),
                                               sc.width, sc.height):
                vis[i, k] = v
                k += 1
        nvis[i] = k
    out = dict(images=np.array(sc.images), P=Ps, xaxis=np.array(xaxes), center=np.array(centers),
               pos=seeds["pos"], nrm=seeds["nrm"], ref=seeds["ref"], nvis=nvis, vis=vis)
    for s in (5, 7, 11, 16):
        tex = np.zeros((n, V, s, s, 3), np.uint8)
        valid = np.zeros((n, V), np.uint8)
        ncc = np.zeros((n, V), np.float64)
        rois = np.zeros((n, V, 4), np.int32)
        keep = np.zeros(n, np.uint8)
        fvis = np.full((n, V), -1, np.int32)
        fnvis = np.zeros(n, np.int32)
        for i in range(n):
            vi = [int(v) for v in vis[i, :nvis[i]]]
            t, r, _ = textures_cv2(Ps, xaxes, sc.images, int(seeds["ref"][i]), vi, s,
                                   seeds["nrm"][i], seeds["pos"][i])
            scores = []
            for k, tk in enumerate(t):
                rois[i, k] = r[k]
                if tk is not None:
                    tex[i, k] = tk
                    valid[i, k] = 1
                if k > 0:
                    scv = ncc_cv2(t[0], tk)
                    ncc[i, k] = scv
                    scores.append(scv)
            kp, fv = filter_ref(scores, vi, 0.6, 2)
            keep[i] = kp
            fnvis[i] = len(fv)
            fvis[i, :len(fv)] = fv
        out.update({f"tex{s}": tex, f"valid{s}": valid, f"ncc{s}": ncc, f"roi{s}": rois,
                    f"keep{s}": keep, f"fvis{s}": fvis, f"fnvis{s}": fnvis})
        print(f"s={s}: valid {valid.sum()}/{nvis.sum()}  keep {keep.sum()}/{n}  "
              f"ncc median {np.median(ncc[:, 1][nvis >= 2]):.3f}")
    np.savez_compressed(os.path.join(OUT, "golden_scoring.npz"), **out)
    for f in ("golden_primitives.npz", "golden_scoring.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()

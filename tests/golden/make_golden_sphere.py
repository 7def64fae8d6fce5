"""Third golden set: the whole scoring chain (ROI rule, fp32 quad, cv2.findHomography,
cv2.warpPerspective, cv2.cvtColor, cv2.meanStdDev, the filter's erase loop) driven with the
REAL cv2 on a curved scene -- textured sphere, 6 views 160x120, strongly tilted patches -- for
the cell sizes 3, 8, 13, 20 (tests/golden/make_golden.py holds a plane at 5, 7, 11, 16).
Run from the repo root:  python tests/golden/make_golden_sphere.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as mg  # noqa: E402
from densepoints_b200 import scenes  # noqa: E402

CELLS = (3, 8, 13, 20)


def main():
    sc = scenes.make_sphere_scene(seed=21, n_views=6, width=160, height=120, f=336.0, radius=5.0,
                                  distance=20.0, cap_deg=22.0, name="golden-sphere")
    seeds = scenes.make_seeds(sc, 64, seed=22, depth_noise=0.02, tilt_deg=30.0)
    Ps = sc.P
    dec = [mg.decompose(P) for P in Ps]
    xaxes = [d[1][0] for d in dec]
    centers = [d[2] for d in dec]
    n, V = seeds["pos"].shape[0], sc.n_views
    vis = np.full((n, V), -1, np.int32)
    nvis = np.zeros(n, np.int32)
    for i in range(n):
        k = 0
        for v in range(V):
            if v != seeds["ref"][i] and mg.inside(Ps[v], seeds["pos"][i].astype(np.float64),
                                                  sc.width, sc.height):
                vis[i, k] = v
                k += 1
        nvis[i] = k
    out = dict(images=np.array(sc.images), P=Ps, xaxis=np.array(xaxes), center=np.array(centers),
               pos=seeds["pos"], nrm=seeds["nrm"], ref=seeds["ref"], nvis=nvis, vis=vis)
    for s in CELLS:
        tex = np.zeros((n, V, s, s, 3), np.uint8)
        valid = np.zeros((n, V), np.uint8)
        ncc = np.zeros((n, V), np.float64)
        keep = np.zeros(n, np.uint8)
        fvis = np.full((n, V), -1, np.int32)
        fnvis = np.zeros(n, np.int32)
        for i in range(n):
            vi = [int(v) for v in vis[i, :nvis[i]]]
            t, _, _ = mg.textures_cv2(Ps, xaxes, sc.images, int(seeds["ref"][i]), vi, s,
                                      seeds["nrm"][i], seeds["pos"][i])
            scores = []
            for k, tk in enumerate(t):
                if tk is not None:
                    tex[i, k] = tk
                    valid[i, k] = 1
                if k > 0:
                    scv = mg.ncc_cv2(t[0], tk)
                    ncc[i, k] = scv
                    scores.append(scv)
            kp, fv = mg.filter_ref(scores, vi, 0.6, 3)
            keep[i] = kp
            fnvis[i] = len(fv)
            fvis[i, :len(fv)] = fv
        out.update({f"tex{s}": tex, f"valid{s}": valid, f"ncc{s}": ncc, f"keep{s}": keep,
                    f"fvis{s}": fvis, f"fnvis{s}": fnvis})
        print(f"s={s}: valid {valid.sum()}/{nvis.sum()}  keep {keep.sum()}/{n}  "
              f"ncc median {np.median(ncc[:, 1][nvis >= 2]):.3f}")
    p = os.path.join(HERE, "golden_scoring_sphere.npz")
    np.savez_compressed(p, **out)
    print(os.path.getsize(p) // 1024, "KiB")


if __name__ == "__main__":
    main()

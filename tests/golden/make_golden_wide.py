"""Second golden set for the OpenCV primitives of the path, made with the REAL cv2 (4.13.0):
cell sizes the first set does not hold (2 ... 32), ROIs up to 48 px, strongly projective
quads, quads lying largely outside their ROI (BORDER_REPLICATE on most texels) and 1-pixel ROIs.
Pins orc_find_homography4 / orc_warp_perspective beyond tests/golden/make_golden.py.
Run from the repo root:  python tests/golden/make_golden_wide.py
"""
import os

import cv2
import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
cv2.setNumThreads(1)
SMAX = 32


def main():
    rng = np.random.default_rng(20261019)
    img = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    # smooth half: neighbouring pixels differ little, so weight errors of 1/32 still show
    yy, xx = np.mgrid[0:120, 0:80]
    for c in range(3):
        img[:, :80, c] = (128 + 90 * np.sin(0.21 * xx + 0.13 * yy * (c + 1))).astype(np.uint8)
    quads, ss, rois, Hs, texs, kinds = [], [], [], [], [], []
    cells = [2, 3, 4, 6, 8, 9, 13, 20, 32]
    for it in range(360):
        s = int(cells[it % len(cells)])
        kind = it % 4
        if kind == 3:
            w = h = 1                                         # single-pixel ROI: all taps clamp
        else:
            w = int(rng.integers(2, 49)); h = int(rng.integers(2, 49))
        x0 = int(rng.integers(0, 160 - w)); y0 = int(rng.integers(0, 120 - h))
        quad = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float32)
        quad += rng.uniform(-0.95, 0.95, (4, 2)).astype(np.float32)
        if kind == 1:                                         # strongly projective (keystone)
            k = rng.uniform(0.15, 0.4)
            quad[1, 1] += np.float32(k * h); quad[2, 1] -= np.float32(k * h)
            quad[0, 0] += np.float32(0.2 * w)
        if kind == 2:                                         # mostly outside the ROI
            quad += rng.uniform(-1.0, 1.0, 2).astype(np.float32) * np.float32(0.8 * max(w, h))
        cell = np.array([[0, 0], [s, 0], [s, s], [0, s]], np.float32)
        H, _ = cv2.findHomography(quad, cell, 0)
        if H is None:
            continue
        tex = cv2.warpPerspective(img[y0:y0 + h, x0:x0 + w], H, (s, s), flags=cv2.INTER_LINEAR,
                                  borderMode=cv2.BORDER_REPLICATE)
        pad = np.zeros((SMAX, SMAX, 3), np.uint8)
        pad[:s, :s] = tex
        quads.append(quad); ss.append(s); rois.append([x0, y0, w, h]); Hs.append(H)
        texs.append(pad); kinds.append(kind)
    np.savez_compressed(os.path.join(OUT, "golden_primitives_wide.npz"), image=img,
                        quad=np.array(quads), s=np.array(ss, np.int32), roi=np.array(rois, np.int32),
                        H=np.array(Hs), tex=np.array(texs), kind=np.array(kinds, np.int32))
    p = os.path.join(OUT, "golden_primitives_wide.npz")
    print(len(ss), "cases,", os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()

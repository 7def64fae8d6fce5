"""Synthetic multi-view scenes "in the style of" the reference's
tests/test_data_generator (test_data_generator.cpp:4-55): pinhole cameras
P = K [R | t] with K = [[f,0,cx],[0,f,cy],[0,0,1]] looking at an analytic,
procedurally textured surface.  The reference generator only makes projection
matrices and random points, so the images, the surface and the seed
patches are synthesised here, numpy only, from fixed seeds.

Pixel convention = the reference's: image.at(row=(int)v, col=(int)u) for a
projected point (u, v) (patch.cpp:65-66); integer coordinates are pixel centres
(cv::remap).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Scene:
    name: str
    P: np.ndarray                 # (V, 3, 4) float64
    images: list                  # V x (H, W, 3) uint8, BGR
    width: int
    height: int
    surface: str                  # "plane" | "sphere"
    radius: float = 0.0           # sphere radius
    extent: float = 0.0           # plane half-extent
    centers: np.ndarray = field(default=None)   # (V, 3) camera centres (ground truth)

    @property
    def n_views(self):
        return len(self.images)


def look_at(center, target, up=(0.0, -1.0, 0.0)):
    """World->camera rotation R with +z towards the target, +x to the right, +y down."""
    c = np.asarray(center, float)
    z = np.asarray(target, float) - c
    z /= np.linalg.norm(z)
    up = np.asarray(up, float)
    x = np.cross(-up, z)
    if np.linalg.norm(x) < 1e-9:
        x = np.cross(np.array([0.0, 0.0, 1.0]), z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return np.stack([x, y, z])


def projection(f, cx, cy, R, center):
    K = np.array([[f, 0, cx], [0, f, cy], [0, 0, 1.0]])
    t = -R @ np.asarray(center, float)
    return K @ np.hstack([R, t[:, None]])


class Texture3D:
    """Band-limited procedural RGB texture defined on world space: a sum of random
    sinusoids, so every view samples the same surface signal (multi-view consistent).
    `wavelength` = (min, max) world units."""

    def __init__(self, seed, wavelength, n_base=10, n_chan=4):
        rng = np.random.default_rng(seed)

        def waves(n):
            d = rng.normal(size=(n, 3))
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            lam = np.exp(rng.uniform(np.log(wavelength[0]), np.log(wavelength[1]), n))
            return d * (2 * np.pi / lam)[:, None], rng.uniform(0, 2 * np.pi, n)

        self.kb, self.pb = waves(n_base)
        self.kc = []
        for _ in range(3):
            self.kc.append(waves(n_chan))
        self.n_base, self.n_chan = n_base, n_chan

    def __call__(self, X):
        """X: (N, 3) float64 -> (N, 3) uint8 BGR."""
        base = np.sin(X @ self.kb.T + self.pb).sum(1) / np.sqrt(self.n_base / 2.0)
        out = np.empty((X.shape[0], 3), np.uint8)
        for c in range(3):
            k, p = self.kc[c]
            ch = np.sin(X @ k.T + p).sum(1) / np.sqrt(self.n_chan / 2.0)
            v = 128.0 + 48.0 * base + 24.0 * ch
            out[:, c] = np.clip(np.rint(v), 0, 255).astype(np.uint8)
        return out


def _render(P_list, centers, Rs, f, cx, cy, width, height, surface, param, tex, chunk=1 << 18,
            only_views=None):
    """only_views: render just these view ids (the others stay black) -- multi-process drivers
    render a slice per rank and exchange the images (distributed.share_images)."""
    images = []
    jj, ii = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    pix = np.stack([(jj.ravel() - cx) / f, (ii.ravel() - cy) / f, np.ones(width * height)], 1)
    for vid, (C0, R) in enumerate(zip(centers, Rs)):
        img = np.zeros((height * width, 3), np.uint8)
        if only_views is not None and vid not in only_views:
            images.append(img.reshape(height, width, 3))
            continue
        for s in range(0, pix.shape[0], chunk):
            d = pix[s:s + chunk] @ R          # rows: R^T * pix  (camera -> world)
            if surface == "plane":            # z = 0
                t = -C0[2] / d[:, 2]
                hit = t > 0
            else:                              # sphere |X| = r, nearest root
                b = d @ C0
                a = (d * d).sum(1)
                cc = C0 @ C0 - param * param
                disc = b * b - a * cc
                hit = disc > 0
                t = (-b - np.sqrt(np.where(hit, disc, 0.0))) / a
                hit &= t > 0
            X = C0[None, :] + t[:, None] * d
            col = tex(X)
            if surface == "plane":
                hit &= (np.abs(X[:, 0]) <= param) & (np.abs(X[:, 1]) <= param)
            bg = np.array([37, 37, 37], np.uint8)
            img[s:s + chunk] = np.where(hit[:, None], col, bg[None, :])
        images.append(img.reshape(height, width, 3))
    return images


def _render_torch(centers, Rs, f, cx, cy, width, height, surface, param, tex, device,
                  only_views=None):
    """_render on a torch device (fp64): for the large benchmark scenes only (64 x 1920x1080 takes
    minutes in numpy).  Test-data plumbing, not the product; pixel values may differ from the
    numpy renderer in the last bit of sin(), so golden vectors always use _render.
    only_views: as in _render (the other views stay black)."""
    import torch
    dev = torch.device(device)
    t64 = lambda a: torch.as_tensor(np.asarray(a, np.float64), device=dev)
    jj, ii = torch.meshgrid(torch.arange(width, dtype=torch.float64, device=dev),
                            torch.arange(height, dtype=torch.float64, device=dev), indexing="xy")
    pix = torch.stack([(jj.reshape(-1) - cx) / f, (ii.reshape(-1) - cy) / f,
                       torch.ones(width * height, dtype=torch.float64, device=dev)], 1)
    kb, pb = t64(tex.kb), t64(tex.pb)
    kc = [(t64(k), t64(p)) for k, p in tex.kc]
    images = []
    for vid, (C0, R) in enumerate(zip(centers, Rs)):
        if only_views is not None and vid not in only_views:
            images.append(np.zeros((height, width, 3), np.uint8))
            continue
        C0t, Rt = t64(C0), t64(R)
        d = pix @ Rt
        if surface == "plane":
            t = -C0t[2] / d[:, 2]
            hit = t > 0
        else:
            b = d @ C0t
            a = (d * d).sum(1)
            disc = b * b - a * (C0t @ C0t - param * param)
            hit = disc > 0
            t = (-b - torch.sqrt(torch.where(hit, disc, torch.zeros_like(disc)))) / a
            hit &= t > 0
        X = C0t[None, :] + t[:, None] * d
        base = torch.sin(X @ kb.T + pb).sum(1) / np.sqrt(tex.n_base / 2.0)
        col = torch.empty((X.shape[0], 3), dtype=torch.uint8, device=dev)
        for c in range(3):
            ch = torch.sin(X @ kc[c][0].T + kc[c][1]).sum(1) / np.sqrt(tex.n_chan / 2.0)
            v = 128.0 + 48.0 * base + 24.0 * ch
            col[:, c] = torch.clamp(torch.round(v), 0, 255).to(torch.uint8)
        if surface == "plane":
            hit &= (X[:, 0].abs() <= param) & (X[:, 1].abs() <= param)
        col = torch.where(hit[:, None], col, torch.full_like(col, 37))
        images.append(col.reshape(height, width, 3).cpu().numpy())
    return images


def make_plane_scene(seed=1, n_views=3, width=640, height=480, f=None, distance=20.0,
                     yaw_spread_deg=15.0, extent=None, name="C1-plane", only_views=None,
                     device=None):
    """Config C1: textured plane z=0 seen by `n_views` TestScene-style pinholes
    (test_data_generator.cpp:8-13 scaled to the image: f = width/4 * ... ) placed on an
    arc at `distance` with +-yaw_spread around the plane normal."""
    f = f if f is not None else float(width)   # 640 -> f=640: the plane fills the view
    cx, cy = width / 2.0, height / 2.0
    extent = extent if extent is not None else 0.45 * distance * width / f
    rng = np.random.default_rng(seed)
    yaws = np.linspace(-yaw_spread_deg, yaw_spread_deg, n_views) if n_views > 1 else [0.0]
    centers, Rs, Ps = [], [], []
    for k, yaw in enumerate(yaws):
        a = np.deg2rad(yaw)
        pitch = np.deg2rad(rng.uniform(-5, 5))
        c = distance * np.array([np.sin(a) * np.cos(pitch), np.sin(pitch), -np.cos(a) * np.cos(pitch)])
        R = look_at(c, rng.uniform(-0.5, 0.5, 3) * np.array([1, 1, 0]))
        centers.append(c)
        Rs.append(R)
        Ps.append(projection(f, cx, cy, R, c))
    px_world = distance / f
    tex = Texture3D(seed + 1000, (3.0 * px_world, 14.0 * px_world))
    if device is not None:
        images = _render_torch(centers, Rs, f, cx, cy, width, height, "plane", extent, tex, device,
                               only_views=only_views)
    else:
        images = _render(Ps, centers, Rs, f, cx, cy, width, height, "plane", extent, tex,
                         only_views=only_views)
    return Scene(name, np.array(Ps), images, width, height, "plane", extent=extent,
                 centers=np.array(centers))


def lattice_plane_cameras(nx=16, ny=16, spacing=4.0, height=20.0, width=3840, height_px=2160,
                          f=3000.0, tilt_deg=20.0, seed=5):
    """BASELINE configs[4] style cameras: an nx x ny lattice `height` above the textured ground
    plane z = 0 (the cameras sit at z < 0 like everywhere in this module), each looking down at a
    point up to +-tilt_deg off its nadir.  Returns (P (V,3,4), centers, Rs, f, cx, cy, extent,
    texture): everything but the images, which render_plane_view() makes one at a time (256 views
    of 3840x2160 are 6.4 GB as BGR)."""
    cx, cy = width / 2.0, height_px / 2.0
    rng = np.random.default_rng(seed)
    centers, Rs, Ps = [], [], []
    for j in range(ny):
        for i in range(nx):
            c = np.array([(i - (nx - 1) / 2.0) * spacing, (j - (ny - 1) / 2.0) * spacing, -height])
            t = np.tan(np.deg2rad(tilt_deg)) * height
            target = np.array([c[0] + rng.uniform(-t, t) * 0.5, c[1] + rng.uniform(-t, t) * 0.5, 0.0])
            R = look_at(c, target)
            centers.append(c)
            Rs.append(R)
            Ps.append(projection(f, cx, cy, R, c))
    extent = 0.5 * max(nx, ny) * spacing + height * width / f
    px_world = height / f
    tex = Texture3D(seed + 1000, (3.0 * px_world, 14.0 * px_world))
    return np.array(Ps), np.array(centers), Rs, f, cx, cy, extent, tex


def render_plane_view(center, R, f, cx, cy, width, height, extent, tex, device):
    """One view of the plane scene on a torch device (see _render_torch)."""
    return _render_torch([center], [R], f, cx, cy, width, height, "plane", extent, tex, device)[0]


def make_sphere_scene(seed=2, n_views=16, width=1280, height=960, f=1000.0, radius=5.0,
                      distance=20.0, cap_deg=32.0, name="C2-sphere", only_views=None):
    """Configs C2/C3: textured sphere, `n_views` cameras on a spherical cap
    (sqrt(n) x sqrt(n) grid of azimuth/elevation within +-cap_deg) looking at the centre."""
    cx, cy = width / 2.0, height / 2.0
    rng = np.random.default_rng(seed)
    g = int(np.ceil(np.sqrt(n_views)))
    angs = np.linspace(-cap_deg, cap_deg, g) if g > 1 else np.array([0.0])
    centers, Rs, Ps = [], [], []
    for k in range(n_views):
        az = np.deg2rad(angs[k % g] + rng.uniform(-1.5, 1.5))
        el = np.deg2rad(angs[k // g] * 0.75 + rng.uniform(-1.5, 1.5))
        c = distance * np.array([np.sin(az) * np.cos(el), np.sin(el), -np.cos(az) * np.cos(el)])
        R = look_at(c, rng.uniform(-0.3, 0.3, 3))
        centers.append(c)
        Rs.append(R)
        Ps.append(projection(f, cx, cy, R, c))
    px_world = (distance - radius) / f
    tex = Texture3D(seed + 1000, (3.0 * px_world, 14.0 * px_world))
    images = _render(Ps, centers, Rs, f, cx, cy, width, height, "sphere", radius, tex,
                     only_views=only_views)
    return Scene(name, np.array(Ps), images, width, height, "sphere", radius=radius,
                 centers=np.array(centers))


def _tilt(n, rng, max_deg):
    """Rotate unit vectors n (N,3) by a random angle in [0, max_deg] about a random axis."""
    N = n.shape[0]
    ax = rng.normal(size=(N, 3))
    ax -= (ax * n).sum(1, keepdims=True) * n
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    a = np.deg2rad(rng.uniform(0, max_deg, N))[:, None]
    return n * np.cos(a) + np.cross(ax, n) * np.sin(a)


def make_seeds(scene: Scene, n, seed=0, depth_noise=0.01, tilt_deg=10.0):
    """Seed patches on the surface, Seed::CreatePatchesFromPoints style
    (seed.cpp:26-54): ref = nearest camera centre (first minimum wins); the normal
    is the *inward* surface normal (pointing away from the cameras, like the
    reference's unit viewing ray) tilted by <= tilt_deg; the position is displaced
    along the reference ray by U(-depth_noise, depth_noise) relative depth.
    Returns fp32 pos/nrm (PointXYZRGBNormal storage) and int32 ref."""
    rng = np.random.default_rng(seed)
    C = scene.centers
    if scene.surface == "plane":
        e = scene.extent * 0.8
        X = np.stack([rng.uniform(-e, e, n), rng.uniform(-e, e, n), np.zeros(n)], 1)
        nin = np.tile(np.array([0.0, 0.0, 1.0]), (n, 1))   # cameras sit at z < 0
    else:
        mean_dir = C.mean(0)
        mean_dir /= np.linalg.norm(mean_dir)
        X = np.empty((0, 3))
        while X.shape[0] < n:
            d = rng.normal(size=(2 * n, 3))
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            d = d[d @ mean_dir > 0.80]
            X = np.vstack([X, d * scene.radius])
        X = X[:n]
        nin = -X / scene.radius
    ref = np.empty(n, np.int32)
    for s0 in range(0, n, 1 << 18):       # chunked: n x views x 3 doubles at once is GBs for 1e6+
        dist = np.linalg.norm(X[s0:s0 + (1 << 18), None, :] - C[None, :, :], axis=2)
        ref[s0:s0 + (1 << 18)] = np.argmin(dist, axis=1)
    Cr = C[ref]
    depth = rng.uniform(-depth_noise, depth_noise, n)[:, None]
    pos = Cr + (1.0 + depth) * (X - Cr)
    nrm = _tilt(nin, rng, tilt_deg) if tilt_deg > 0 else nin
    return dict(pos=pos.astype(np.float32), nrm=nrm.astype(np.float32), ref=ref)


def force_visible(scene: Scene, seeds, k):
    """C3: the k views (reference excluded) nearest by angle to the patch normal,
    ascending view id, as a dense (n, k) int32 visible table."""
    pos = seeds["pos"].astype(np.float64)
    nrm = seeds["nrm"].astype(np.float64)
    d = pos[:, None, :] - scene.centers[None, :, :]
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    cosang = (d * nrm[:, None, :]).sum(2)
    cosang[np.arange(pos.shape[0]), seeds["ref"]] = -2.0
    idx = np.argsort(-cosang, axis=1, kind="stable")[:, :k]
    vis = np.sort(idx, axis=1).astype(np.int32)
    nvis = np.full(pos.shape[0], k, np.int32)
    return nvis, vis

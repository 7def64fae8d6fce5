"""ctypes binding of the CPU ORACLE (oracle/dp_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke(), never from densepoints_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdp_oracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/dp_oracle.c -> oracle/_build/libdp_oracle.so (gcc, OpenMP)."""
    src = [os.path.join(_HERE, f) for f in ("dp_oracle.c", "dp_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE], check=True, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT)
    return _SO


class OrcView(C.Structure):
    _fields_ = [("P", C.c_double * 12), ("xaxis", C.c_double * 3), ("center", C.c_double * 3),
                ("width", C.c_int), ("height", C.c_int), ("bgr", C.c_void_p),
                ("stride", C.c_size_t)]


class OrcParams(C.Structure):
    _fields_ = [("score_threshold", C.c_double), ("minimum_visible_image", C.c_int),
                ("visible_threshold", C.c_double), ("candidate_threshold", C.c_double),
                ("grid_scale", C.c_int), ("max_patches_per_cell", C.c_int),
                ("nm_step", C.c_double * 3), ("nm_max_evals", C.c_int), ("nm_eps", C.c_double),
                ("max_pops", C.c_longlong)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_ncc_bgr.restype = C.c_double
        _lib.orc_ncc_f64.restype = C.c_double
        _lib.orc_downhill.restype = C.c_double
        _lib.orc_objective.restype = C.c_double
        _lib.orc_organizer_create.restype = C.c_void_p
        _lib.orc_organizer_try_insert.restype = C.c_longlong
        _lib.orc_organizer_size.restype = C.c_longlong
        _lib.orc_organizer_grid.restype = C.POINTER(C.c_uint8)
        _lib.orc_expand_patches.restype = C.c_longlong
        _lib.orc_expand_patches_fifo.restype = C.c_longlong
    return _lib


def _p(a, t=None):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def default_params(**kw) -> OrcParams:
    p = OrcParams()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "nm_step":
            for i in range(3):
                p.nm_step[i] = v[i]
        else:
            setattr(p, k, v)
    return p


def view_decompose(P):
    P = np.ascontiguousarray(P, dtype=np.float64).reshape(12)
    K = np.zeros(9)
    R = np.zeros(9)
    c = np.zeros(3)
    lib().orc_view_decompose(_p(P), _p(K), _p(R), _p(c))
    return K.reshape(3, 3), R.reshape(3, 3), c


class Views:
    """std::vector<View> of the reference: projection matrices + BGR u8 images."""

    def __init__(self, Ps, images):
        self.n = len(images)
        self.images = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        self.arr = (OrcView * self.n)()
        for i, (P, im) in enumerate(zip(Ps, self.images)):
            Pc = np.ascontiguousarray(P, dtype=np.float64).reshape(12)
            h, w = im.shape[:2]
            lib().orc_view_init(C.byref(self.arr[i]), _p(Pc), _p(im), C.c_int(w), C.c_int(h),
                                C.c_size_t(im.strides[0]))

    def xaxis(self, i):
        return np.array(self.arr[i].xaxis[:])

    def center(self, i):
        return np.array(self.arr[i].center[:])


def ncc_f64(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    b = np.ascontiguousarray(b, dtype=np.float64).ravel()
    return lib().orc_ncc_f64(_p(a), _p(b), C.c_int(a.size))


def ncc_bgr(ta, tb):
    n = (ta if ta is not None else tb).size // 3
    ta = None if ta is None else np.ascontiguousarray(ta, dtype=np.uint8)
    tb = None if tb is None else np.ascontiguousarray(tb, dtype=np.uint8)
    return lib().orc_ncc_bgr(_p(ta), _p(tb), C.c_int(n))


def find_homography4(src, dst):
    src = _f32(src).reshape(8)
    dst = _f32(dst).reshape(8)
    H = np.zeros(9)
    ok = lib().orc_find_homography4(_p(src), _p(dst), _p(H))
    return (H.reshape(3, 3) if ok else None)


def warp_perspective(src, H, s):
    """src: h x w x 3 uint8 (may be a strided view of a larger image)."""
    assert src.dtype == np.uint8 and src.strides[2] == 1 and src.strides[1] == 3
    H = np.ascontiguousarray(H, dtype=np.float64).reshape(9)
    out = np.zeros((s, s, 3), np.uint8)
    lib().orc_warp_perspective(C.c_void_p(src.ctypes.data), C.c_size_t(src.strides[0]),
                               C.c_int(src.shape[1]), C.c_int(src.shape[0]), _p(H), C.c_int(s),
                               _p(out))
    return out


def pyrdown(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    out = np.zeros(((h + 1) // 2, (w + 1) // 2, 3), np.uint8)
    lib().orc_pyrdown(_p(img), C.c_size_t(img.strides[0]), C.c_int(w), C.c_int(h), _p(out),
                      C.c_size_t(out.strides[0]))
    return out


def gray(b, g, r):
    return lib().orc_gray(C.c_int(int(b)), C.c_int(int(g)), C.c_int(int(r)))


_FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_void_p)


def downhill(fn, x0, step, max_evals=500, eps=1e-4):
    nd = len(x0)
    cb = _FN(lambda xp, _u: float(fn(np.array([xp[i] for i in range(nd)]))))
    x = np.array(x0, dtype=np.float64)
    st = np.array(step, dtype=np.float64)
    fc = C.c_int(0)
    res = lib().orc_downhill(cb, None, C.c_int(nd), _p(x), _p(st), C.c_int(max_evals),
                             C.c_double(eps), C.byref(fc))
    return x, res, fc.value


def patch_homography(views: Views, view_id, cell_size, pos, ax, ay):
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    ax = np.ascontiguousarray(ax, dtype=np.float64)
    ay = np.ascontiguousarray(ay, dtype=np.float64)
    H = np.zeros(9)
    roi = np.zeros(4, np.int32)
    ok = lib().orc_patch_homography(C.byref(views.arr[view_id]), C.c_int(cell_size), _p(pos),
                                    _p(ax), _p(ay), _p(H), _p(roi))
    return ok, H.reshape(3, 3), roi


def axes_scale(views: Views, ref, nrm, pos):
    nrm = np.ascontiguousarray(nrm, dtype=np.float64)
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    xa = np.zeros(3)
    ya = np.zeros(3)
    dx = C.c_double(0)
    lib().orc_axes_scale(C.byref(views.arr[ref]), _p(nrm), _p(pos), _p(xa), _p(ya), C.byref(dx))
    return xa, ya, dx.value


def score_batch(views: Views, pos, nrm, ref, nvis, vis, cell_size, want_tex=False,
                trial_nrm=None, trial_pos=None):
    """Returns ncc (n, vstride) f32 [, tex (n, vstride, s, s, 3) u8, valid (n, vstride) u8].
    trial_nrm / trial_pos (n, 3) float64: GetProjectedTextures(normal, position, ...) -- the
    stored pos stays the corner centre (patch.cpp:119-123)."""
    pos, nrm, ref, nvis, vis = _f32(pos), _f32(nrm), _i32(ref), _i32(nvis), _i32(vis)
    n, vs = vis.shape
    ncc = np.zeros((n, vs), np.float32)
    tex = np.zeros((n, vs, cell_size, cell_size, 3), np.uint8) if want_tex else None
    valid = np.zeros((n, vs), np.uint8) if want_tex else None
    tn = None if trial_nrm is None else np.ascontiguousarray(trial_nrm, dtype=np.float64)
    tp = None if trial_pos is None else np.ascontiguousarray(trial_pos, dtype=np.float64)
    lib().orc_score_at_batch(views.arr, C.c_int(n), _p(pos), _p(nrm), _p(ref), _p(nvis), _p(vis),
                             C.c_int(vs), C.c_int(cell_size), _p(tn), _p(tp), _p(ncc), _p(tex),
                             _p(valid))
    return (ncc, tex, valid) if want_tex else ncc


def unparametrize(views: Views, ref, nrm0, pos0, x):
    """Optimization::UnparametrizePatch (optimization.cpp:78-96) -> (normal, position) fp64."""
    nrm0, pos0 = _f32(nrm0), _f32(pos0)
    n, p = np.zeros(3), np.zeros(3)
    lib().orc_unparametrize(C.byref(views.arr[int(ref)]), _p(nrm0), _p(pos0), C.c_double(x[0]),
                            C.c_double(x[1]), C.c_double(x[2]), _p(n), _p(p))
    return n, p


def filter_batch(views: Views, pos, nrm, ref, nvis, vis, cell_size, thr=0.6, min_visible=3):
    pos, nrm, ref = _f32(pos), _f32(nrm), _i32(ref)
    nvis = _i32(nvis).copy()
    vis = _i32(vis).copy()
    n, vs = vis.shape
    keep = np.zeros(n, np.uint8)
    lib().orc_filter_batch(views.arr, C.c_int(n), _p(pos), _p(nrm), _p(ref), _p(nvis), _p(vis),
                           C.c_int(vs), C.c_int(cell_size), C.c_double(thr), C.c_int(min_visible),
                           _p(keep))
    return keep, nvis, vis


def refine_batch(views: Views, pos, nrm, ref, nvis, vis, cell_size, params=None):
    pos = _f32(pos).copy()
    nrm = _f32(nrm).copy()
    ref, nvis, vis = _i32(ref), _i32(nvis), _i32(vis)
    n, vs = vis.shape
    prm = params or default_params()
    fc = np.zeros(n, np.int32)
    xb = np.zeros((n, 3), np.float64)
    lib().orc_refine_batch(views.arr, C.c_int(n), _p(pos), _p(nrm), _p(ref), _p(nvis), _p(vis),
                           C.c_int(vs), C.c_int(cell_size), C.byref(prm), _p(fc), _p(xb))
    return pos, nrm, fc, xb


def objective(views: Views, ref, vis, cell_size, nrm0, pos0, x):
    vis = _i32(vis)
    nrm0, pos0 = _f32(nrm0), _f32(pos0)
    x = np.ascontiguousarray(x, dtype=np.float64)
    return lib().orc_objective(views.arr, C.c_int(int(ref)), _p(vis), C.c_int(vis.size),
                               C.c_int(cell_size), _p(nrm0), _p(pos0), _p(x))


def visibility_batch(views: Views, pos, nrm, ref, t_vis=0.78, t_cand=1.04, vstride=None):
    pos, nrm, ref = _f32(pos), _f32(nrm), _i32(ref)
    n = pos.shape[0]
    vs = vstride or views.n
    nvis = np.zeros(n, np.int32)
    ncand = np.zeros(n, np.int32)
    vis = np.full((n, vs), -1, np.int32)
    cand = np.full((n, vs), -1, np.int32)
    lib().orc_visibility_batch(views.arr, C.c_int(views.n), C.c_int(n), _p(pos), _p(nrm), _p(ref),
                               C.c_double(t_vis), C.c_double(t_cand), _p(nvis), _p(vis),
                               _p(ncand), _p(cand), C.c_int(vs))
    return nvis, vis, ncand, cand


def create_patches(views: Views, points, t_vis=0.78, t_cand=1.04):
    points = np.ascontiguousarray(points, dtype=np.float64)
    n = points.shape[0]
    out = dict(pos=np.zeros((n, 3), np.float32), nrm=np.zeros((n, 3), np.float32),
               ref=np.zeros(n, np.int32), nvis=np.zeros(n, np.int32),
               vis=np.full((n, views.n), -1, np.int32))
    lib().orc_create_patches(views.arr, C.c_int(views.n), C.c_int(n), _p(points), C.c_double(t_vis),
                             C.c_double(t_cand), _p(out["pos"]), _p(out["nrm"]), _p(out["ref"]),
                             _p(out["nvis"]), _p(out["vis"]), C.c_int(views.n))
    return out


def compute_color(views: Views, pos):
    pos = _f32(pos)
    out = np.zeros((pos.shape[0], 3), np.uint8)
    for i in range(pos.shape[0]):
        lib().orc_compute_color(views.arr, C.c_int(views.n), _p(pos[i]), _p(out[i]))
    return out


def expand_patch(views: Views, prm, cell_size, pos, nrm, ref, pvis):
    """Expand::ExpandPatch (expand.cpp:103-143): accepted children of one parent as a list of
    (direction, pos f32[3], nrm f32[3], vis i32[nvis])."""
    pos, nrm, pvis = _f32(pos), _f32(nrm), _i32(pvis)
    out_pos = np.zeros((4, 3), np.float32)
    out_nrm = np.zeros((4, 3), np.float32)
    out_nvis = np.zeros(4, np.int32)
    out_vis = np.full((4, views.n), -1, np.int32)
    dirs = np.zeros(4, np.int32)
    cnt = lib().orc_expand_patch(views.arr, C.c_int(views.n), C.byref(prm), C.c_int(cell_size),
                                 _p(pos), _p(nrm), C.c_int(int(ref)), _p(pvis),
                                 C.c_int(pvis.size), _p(out_pos), _p(out_nrm), _p(out_nvis),
                                 _p(out_vis), _p(dirs))
    return [(int(dirs[k]), out_pos[k].copy(), out_nrm[k].copy(), out_vis[k, :out_nvis[k]].copy())
            for k in range(cnt)]


class Organizer:
    """PatchOrganizer of the reference (patch_organizer.cpp) + Expand driver (expand.cpp)."""

    def __init__(self, views: Views, params=None):
        self.views = views
        self.prm = params or default_params()
        self.h = C.c_void_p(lib().orc_organizer_create(views.arr, C.c_int(views.n),
                                                       C.byref(self.prm)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_organizer_destroy(self.h)
            self.h = None

    def try_insert(self, pos, nrm, ref, vis):
        vis = _i32(vis)
        pos, nrm = _f32(pos), _f32(nrm)
        nc = C.c_int(0)
        cells = np.zeros((max(vis.size, 1), 3), np.int32)
        idx = lib().orc_organizer_try_insert(self.h, _p(pos), _p(nrm), C.c_int(int(ref)), _p(vis),
                                             C.c_int(vis.size), C.byref(nc), _p(cells))
        return idx, cells[:nc.value]

    def set_seeds(self, pos, nrm, ref, nvis, vis):
        """PatchOrganizer::SetSeeds (patch_organizer.cpp:70-75)."""
        acc = np.zeros(len(ref), np.uint8)
        for i in range(len(ref)):
            idx, _ = self.try_insert(pos[i], nrm[i], ref[i], vis[i][:nvis[i]])
            acc[i] = idx >= 0
        return acc

    def size(self):
        return lib().orc_organizer_size(self.h)

    def grid(self, view):
        gw, gh = C.c_int(0), C.c_int(0)
        ptr = lib().orc_organizer_grid(self.h, C.c_int(view), C.byref(gw), C.byref(gh))
        return np.ctypeslib.as_array(ptr, shape=(gh.value, gw.value)).copy()

    def export(self, vstride=None):
        n = self.size()
        vs = vstride or self.views.n
        pos = np.zeros((n, 3), np.float32)
        nrm = np.zeros((n, 3), np.float32)
        rgb = np.zeros((n, 3), np.uint8)
        ref = np.zeros(n, np.int32)
        nvis = np.zeros(n, np.int32)
        vis = np.full((n, vs), -1, np.int32)
        lib().orc_organizer_export(self.h, _p(pos), _p(nrm), _p(rgb), _p(ref), _p(nvis), _p(vis),
                                   C.c_int(vs))
        return dict(pos=pos, nrm=nrm, rgb=rgb, ref=ref, nvis=nvis, vis=vis)

    def expand(self, cell_size=11, max_levels=-1):
        return lib().orc_expand_patches(self.h, C.c_int(cell_size), C.c_int(max_levels))

    def expand_fifo(self, cell_size=11, max_pops=-1):
        return lib().orc_expand_patches_fifo(self.h, C.c_int(cell_size), C.c_longlong(max_pops))


_level_tables = None      # keeps the registered level table (and its Views) alive


def set_level_selection(levels, px_per_cell=1.5):
    """Per-(patch, view) pyramid level (dp_oracle.c): levels = [Views of level 0 (the base: the
    object later calls pass as `views`), Views of level 1, ...] or None to switch it off."""
    global _level_tables
    if not levels or len(levels) < 2:
        lib().orc_set_level_selection(None, C.c_int(0), C.c_int(0), C.c_double(px_per_cell))
        _level_tables = None
        return
    nv = levels[0].n
    tab = (OrcView * (nv * len(levels)))()
    tab_p = C.cast(tab, C.POINTER(OrcView))
    for l, V in enumerate(levels):
        assert V.n == nv
        for i in range(nv):
            C.memmove(C.byref(tab, (l * nv + i) * C.sizeof(OrcView)), C.byref(V.arr[i]), C.sizeof(OrcView))
    # level 0 of the table must BE the base array (the C side indexes the table by view id and
    # takes the frame from the `views` argument): same contents, so either works
    lib().orc_set_level_selection(tab_p, C.c_int(len(levels)), C.c_int(nv), C.c_double(px_per_cell))
    _level_tables = (tab, list(levels))


def levels_batch(views, pos, nrm, ref, nvis, vis, cell_size):
    """The level each (patch, visible view) pair is read at under the registered selection."""
    pos, nrm, ref, nvis, vis = _f32(pos), _f32(nrm), _i32(ref), _i32(nvis), _i32(vis)
    n, vs = vis.shape
    out = np.full((n, vs), -1, np.int32)
    lib().orc_levels_batch(views.arr, _p(pos), _p(nrm), _p(ref), _p(nvis), _p(vis), C.c_int(vs),
                           C.c_int(n), C.c_int(cell_size), _p(out))
    return out


def project(views: Views, view_id, X):
    """View::ProjectPoint (types.cpp:70-75)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    uv = np.zeros(2)
    lib().orc_project(C.byref(views.arr[int(view_id)]), _p(X), _p(uv))
    return uv


def set_eigen_pairwise(on: bool):
    """Sum order of Eigen's small fixed-size products: False = sequential (Eigen 3.2, the
    default everywhere), True = the halving order of Eigen >= 3.3 (a measuring device)."""
    lib().orc_set_eigen_pairwise(C.c_int(1 if on else 0))


def set_homography_mode(mode: int):
    """0 = OpenCV's DLT + eigen-solve + inversion (pinned against cv2); 1 = exact closed form
    (deterministic at ties; what the CUDA path is checked against).  See dp_oracle.h."""
    lib().orc_set_homography_mode(C.c_int(mode))


def get_homography_mode() -> int:
    return lib().orc_get_homography_mode()


def use_all_cores():
    """Use every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it
    to 1 for its workers)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(C.c_int(n))
    return num_threads()


def num_threads():
    return lib().orc_num_threads()

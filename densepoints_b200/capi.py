"""ctypes binding of the C ABI (include/densepoints_cuda.h).

There is NO CPU fallback: if the CUDA library is missing this module raises at
load time, and every call raises DpError on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "_build", "libdensepoints_cuda.so")
HEADER = os.path.join(ROOT, "include", "densepoints_cuda.h")


class DpError(RuntimeError):
    pass


class DpParams(C.Structure):
    _fields_ = [("score_threshold", C.c_double), ("minimum_visible_image", C.c_int32),
                ("visible_threshold", C.c_double), ("candidate_threshold", C.c_double),
                ("grid_scale", C.c_int32), ("max_patches_per_cell", C.c_int32),
                ("nm_step", C.c_double * 3), ("nm_max_evals", C.c_int32), ("nm_eps", C.c_double),
                ("max_pops", C.c_int64)]


class DpPatchSoa(C.Structure):
    _fields_ = [("n", C.c_int32), ("vstride", C.c_int32), ("pos", C.c_void_p),
                ("nrm", C.c_void_p), ("ref", C.c_void_p), ("nvis", C.c_void_p),
                ("vis", C.c_void_p), ("rgb", C.c_void_p)]


class DpPatchDev(C.Structure):
    _fields_ = DpPatchSoa._fields_


def declared_symbols():
    """Every function the header declares (used by the CPU-side export test)."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dp_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("DENSEPOINTS_CUDA_LIB", LIB_PATH)   # tuning builds only
        if not os.path.exists(path):
            raise DpError(f"{path} is missing: run __graft_entry__.build() "
                          "(there is no CPU fallback for the CUDA path)")
        L = C.CDLL(path)
        L.dp_last_error.restype = C.c_char_p
        L.dp_last_error.argtypes = [C.c_void_p]
        L.dp_launch_count.restype = C.c_int64
        L.dp_launch_count.argtypes = [C.c_void_p]
        L.dp_organizer_size.restype = C.c_int64
        L.dp_organizer_size.argtypes = [C.c_void_p]
        L.dp_expand_last_candidates.restype = C.c_int64
        L.dp_expand_last_candidates.argtypes = [C.c_void_p]
        L.dp_record_bytes.restype = C.c_size_t
        L.dp_record_bytes.argtypes = [C.c_void_p]
        L.dp_destroy.restype = None
        L.dp_destroy.argtypes = [C.c_void_p]
        L.dp_default_params.restype = None
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(int(a))          # raw device pointer (torch .data_ptr())


def default_params(**kw) -> DpParams:
    p = DpParams()
    lib().dp_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "nm_step":
            for i in range(3):
                p.nm_step[i] = v[i]
        else:
            setattr(p, k, v)
    return p


def _soa(pos, nrm, ref, nvis, vis, rgb=None, cls=DpPatchSoa):
    s = cls()
    if isinstance(vis, np.ndarray):
        s.n, s.vstride = vis.shape
    else:
        raise TypeError("vis must be a numpy array here")
    s.pos, s.nrm, s.ref, s.nvis, s.vis = (_ptr(x) for x in (pos, nrm, ref, nvis, vis))
    s.rgb = _ptr(rgb)
    return s


def dev_batch(n, vstride, pos, nrm, ref, nvis, vis, rgb=0) -> DpPatchDev:
    """dp_patch_dev from raw device pointers (ints, e.g. torch tensor.data_ptr())."""
    d = DpPatchDev()
    d.n, d.vstride = int(n), int(vstride)
    d.pos, d.nrm, d.ref, d.nvis, d.vis, d.rgb = (C.c_void_p(int(x)) if x else None
                                                 for x in (pos, nrm, ref, nvis, vis, rgb))
    return d


class Context:
    """dp_context handle.  One per GPU / host thread."""

    def __init__(self, device=-1, params: DpParams | None = None):
        self._h = C.c_void_p()
        rc = lib().dp_create(C.byref(self._h), C.c_int(device),
                             C.byref(params) if params is not None else None)
        if rc != 0:
            raise DpError(f"dp_create failed with status {rc} "
                          "(no CUDA device? the CUDA path has no CPU fallback)")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().dp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise DpError(f"{what} -> {rc}: {lib().dp_last_error(self._h).decode()}")

    # ---- params / views --------------------------------------------------------------
    def set_params(self, p: DpParams):
        self._ck(lib().dp_set_params(self._h, C.byref(p)), "dp_set_params")

    def get_params(self) -> DpParams:
        p = DpParams()
        self._ck(lib().dp_get_params(self._h, C.byref(p)), "dp_get_params")
        return p

    def set_views(self, Ps, images, xaxes=None, centers=None):
        self._ck(lib().dp_set_num_views(self._h, C.c_int(len(images))), "dp_set_num_views")
        for i, (P, im) in enumerate(zip(Ps, images)):
            P = np.ascontiguousarray(P, dtype=np.float64).reshape(12)
            im = np.ascontiguousarray(im, dtype=np.uint8)
            h, w = im.shape[:2]
            xa = None if xaxes is None else np.ascontiguousarray(xaxes[i], dtype=np.float64)
            ce = None if centers is None else np.ascontiguousarray(centers[i], dtype=np.float64)
            self._ck(lib().dp_upload_view(self._h, C.c_int(i), _ptr(P), _ptr(xa), _ptr(ce),
                                          _ptr(im), C.c_int(w), C.c_int(h),
                                          C.c_size_t(im.strides[0])), "dp_upload_view")

    def set_num_views(self, n):
        self._ck(lib().dp_set_num_views(self._h, C.c_int(n)), "dp_set_num_views")

    def upload_view(self, i, P, image):
        """One view at a time (after set_num_views): large image sets need not sit on the host."""
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(12)
        im = np.ascontiguousarray(image, dtype=np.uint8)
        h, w = im.shape[:2]
        self._ck(lib().dp_upload_view(self._h, C.c_int(i), _ptr(P), None, None, _ptr(im), C.c_int(w),
                                      C.c_int(h), C.c_size_t(im.strides[0])), "dp_upload_view")

    def num_views(self):
        return lib().dp_num_views(self._h)

    def get_view(self, i):
        xa, ce = np.zeros(3), np.zeros(3)
        w, h = C.c_int(0), C.c_int(0)
        self._ck(lib().dp_get_view(self._h, C.c_int(i), _ptr(xa), _ptr(ce), C.byref(w),
                                   C.byref(h)), "dp_get_view")
        return xa, ce, w.value, h.value

    def sync(self):
        self._ck(lib().dp_sync(self._h), "dp_sync")

    def launch_count(self):
        return int(lib().dp_launch_count(self._h))

    # ---- host-buffer layer ---------------------------------------------------------------
    def score(self, pos, nrm, ref, nvis, vis, cell_size, want_tex=False):
        pos, nrm = (np.ascontiguousarray(a, dtype=np.float32) for a in (pos, nrm))
        ref, nvis, vis = (np.ascontiguousarray(a, dtype=np.int32) for a in (ref, nvis, vis))
        n, vs = vis.shape
        ncc = np.zeros((n, vs), np.float32)
        tex = np.zeros((n, vs, cell_size, cell_size, 3), np.uint8) if want_tex else None
        valid = np.zeros((n, vs), np.uint8) if want_tex else None
        s = _soa(pos, nrm, ref, nvis, vis)
        self._ck(lib().dp_score(self._h, C.byref(s), C.c_int(cell_size), _ptr(ncc), _ptr(tex),
                                _ptr(valid)), "dp_score")
        return (ncc, tex, valid) if want_tex else ncc

    def score_at(self, pos, nrm, ref, nvis, vis, cell_size, normal=None, position=None):
        """Optimization::GetProjectedTextures(normal, position, textures) + the NCC loop: trial
        (normal, position) as float64 (n, 3); returns (ncc, tex, valid)."""
        pos, nrm = (np.ascontiguousarray(a, dtype=np.float32) for a in (pos, nrm))
        ref, nvis, vis = (np.ascontiguousarray(a, dtype=np.int32) for a in (ref, nvis, vis))
        n, vs = vis.shape
        tn = None if normal is None else np.ascontiguousarray(normal, dtype=np.float64)
        tp = None if position is None else np.ascontiguousarray(position, dtype=np.float64)
        ncc = np.zeros((n, vs), np.float32)
        tex = np.zeros((n, vs, cell_size, cell_size, 3), np.uint8)
        valid = np.zeros((n, vs), np.uint8)
        s = _soa(pos, nrm, ref, nvis, vis)
        self._ck(lib().dp_score_at(self._h, C.byref(s), C.c_int(cell_size), _ptr(tn), _ptr(tp),
                                   _ptr(ncc), _ptr(tex), _ptr(valid)), "dp_score_at")
        return ncc, tex, valid

    def filter(self, pos, nrm, ref, nvis, vis, cell_size):
        pos, nrm = (np.ascontiguousarray(a, dtype=np.float32) for a in (pos, nrm))
        ref = np.ascontiguousarray(ref, dtype=np.int32)
        nvis = np.array(nvis, dtype=np.int32, copy=True)
        vis = np.array(vis, dtype=np.int32, copy=True, order="C")
        keep = np.zeros(vis.shape[0], np.uint8)
        s = _soa(pos, nrm, ref, nvis, vis)
        self._ck(lib().dp_filter(self._h, C.byref(s), C.c_int(cell_size), _ptr(keep)), "dp_filter")
        return keep, nvis, vis

    def refine(self, pos, nrm, ref, nvis, vis, cell_size, mask=None):
        pos = np.array(pos, dtype=np.float32, copy=True, order="C")
        nrm = np.array(nrm, dtype=np.float32, copy=True, order="C")
        ref, nvis, vis = (np.ascontiguousarray(a, dtype=np.int32) for a in (ref, nvis, vis))
        n = vis.shape[0]
        evals = np.zeros(n, np.int32)
        xbest = np.zeros((n, 3), np.float64)
        s = _soa(pos, nrm, ref, nvis, vis)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._ck(lib().dp_refine(self._h, C.byref(s), C.c_int(cell_size), _ptr(mask), _ptr(evals),
                                 _ptr(xbest)), "dp_refine")
        return pos, nrm, evals, xbest

    def filter_refine(self, pos, nrm, ref, nvis, vis, cell_size):
        """Seed::OptimizeAndRefinePatches in one call (copies its inputs)."""
        pos = np.array(pos, dtype=np.float32, copy=True, order="C")
        nrm = np.array(nrm, dtype=np.float32, copy=True, order="C")
        ref = np.ascontiguousarray(ref, dtype=np.int32)
        nvis = np.array(nvis, dtype=np.int32, copy=True)
        vis = np.array(vis, dtype=np.int32, copy=True, order="C")
        keep = np.zeros(vis.shape[0], np.uint8)
        evals = np.zeros(vis.shape[0], np.int32)
        self.filter_refine_inplace(pos, nrm, ref, nvis, vis, cell_size, keep, evals)
        return keep, nvis, vis, pos, nrm, evals

    def filter_refine_inplace(self, pos, nrm, ref, nvis, vis, cell_size, keep, evals=None):
        """dp_filter_refine on the caller's own (e.g. pinned) arrays, edited in place."""
        s = _soa(pos, nrm, ref, nvis, vis)
        self._ck(lib().dp_filter_refine(self._h, C.byref(s), C.c_int(cell_size), _ptr(keep),
                                        _ptr(evals)), "dp_filter_refine")

    def visibility(self, pos, nrm, ref, vstride=None):
        pos, nrm = (np.ascontiguousarray(a, dtype=np.float32) for a in (pos, nrm))
        ref = np.ascontiguousarray(ref, dtype=np.int32)
        n = pos.shape[0]
        vs = vstride or self.num_views()
        nvis, ncand = np.zeros(n, np.int32), np.zeros(n, np.int32)
        vis, cand = np.full((n, vs), -1, np.int32), np.full((n, vs), -1, np.int32)
        s = _soa(pos, nrm, ref, nvis, vis)
        self._ck(lib().dp_visibility(self._h, C.byref(s), _ptr(ncand), _ptr(cand)),
                 "dp_visibility")
        return nvis, vis, ncand, cand

    def color(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        rgb = np.zeros((pos.shape[0], 3), np.uint8)
        s = DpPatchSoa()
        s.n, s.vstride = pos.shape[0], 1
        s.pos, s.rgb = _ptr(pos), _ptr(rgb)
        self._ck(lib().dp_color(self._h, C.byref(s)), "dp_color")
        return rgb

    def create_patches(self, points, vstride=None):
        """Seed::CreatePatchesFromPoints: points (n,3) float64 -> dict(pos,nrm,ref,nvis,vis)."""
        points = np.ascontiguousarray(points, dtype=np.float64)
        n = points.shape[0]
        vs = vstride or self.num_views()
        out = dict(pos=np.zeros((n, 3), np.float32), nrm=np.zeros((n, 3), np.float32),
                   ref=np.zeros(n, np.int32), nvis=np.zeros(n, np.int32),
                   vis=np.full((max(n, 1), vs), -1, np.int32)[:n])
        s = DpPatchSoa()
        s.n, s.vstride = n, vs
        s.pos, s.nrm, s.ref, s.nvis, s.vis = (_ptr(out[k]) for k in ("pos", "nrm", "ref", "nvis", "vis"))
        self._ck(lib().dp_create_patches(self._h, _ptr(points), C.c_int(n), C.byref(s), None, None),
                 "dp_create_patches")
        return out

    def export_ply(self, path):
        self._ck(lib().dp_export_ply(self._h, C.c_char_p(path.encode())), "dp_export_ply")

    # ---- organizer / expansion ---------------------------------------------------------------
    def organizer_reset(self):
        self._ck(lib().dp_organizer_reset(self._h), "dp_organizer_reset")

    def organizer_insert(self, pos, nrm, ref, nvis, vis):
        pos, nrm = (np.ascontiguousarray(a, dtype=np.float32) for a in (pos, nrm))
        ref, nvis, vis = (np.ascontiguousarray(a, dtype=np.int32) for a in (ref, nvis, vis))
        acc = np.zeros(vis.shape[0], np.uint8)
        s = _soa(pos, nrm, ref, nvis, vis)
        self._ck(lib().dp_organizer_insert(self._h, C.byref(s), _ptr(acc)), "dp_organizer_insert")
        return acc

    def organizer_size(self):
        return int(lib().dp_organizer_size(self._h))

    def organizer_export(self, vstride=None):
        n = self.organizer_size()
        vs = vstride or self.num_views()
        cap = max(n, 1)
        out = dict(pos=np.zeros((cap, 3), np.float32), nrm=np.zeros((cap, 3), np.float32),
                   rgb=np.zeros((cap, 3), np.uint8), ref=np.zeros(cap, np.int32),
                   nvis=np.zeros(cap, np.int32), vis=np.full((cap, vs), -1, np.int32))
        s = _soa(out["pos"], out["nrm"], out["ref"], out["nvis"], out["vis"], out["rgb"])
        self._ck(lib().dp_organizer_export(self._h, C.byref(s)), "dp_organizer_export")
        return {k: v[:n] for k, v in out.items()}

    def organizer_grid(self, view):
        gw, gh = C.c_int(0), C.c_int(0)
        self._ck(lib().dp_organizer_grid(self._h, C.c_int(view), None, C.c_size_t(0),
                                         C.byref(gw), C.byref(gh)), "dp_organizer_grid")
        out = np.zeros((gh.value, gw.value), np.uint8)
        self._ck(lib().dp_organizer_grid(self._h, C.c_int(view), _ptr(out), C.c_size_t(out.size),
                                         C.byref(gw), C.byref(gh)), "dp_organizer_grid")
        return out

    def organizer_grids(self):
        """All occupancy grids concatenated in view order (uint8)."""
        n = C.c_int64(0)
        self._ck(lib().dp_organizer_grids(self._h, None, C.c_size_t(0), C.byref(n)),
                 "dp_organizer_grids")
        out = np.zeros(max(n.value, 1), np.uint8)
        self._ck(lib().dp_organizer_grids(self._h, _ptr(out), C.c_size_t(out.size), C.byref(n)),
                 "dp_organizer_grids")
        return out[:n.value]

    def expand(self, cell_size=11, max_levels=-1):
        stats = np.zeros(4, np.int64)
        self._ck(lib().dp_expand(self._h, C.c_int(cell_size), C.c_int(max_levels), _ptr(stats)),
                 "dp_expand")
        return dict(pops=int(stats[0]), candidates=int(stats[1]), passed=int(stats[2]),
                    inserted=int(stats[3]))

    # multi-GPU level steps (device pointers are ints)
    def record_bytes(self):
        return int(lib().dp_record_bytes(self._h))

    def expand_frontier(self):
        b, e = C.c_int64(0), C.c_int64(0)
        self._ck(lib().dp_expand_frontier(self._h, C.byref(b), C.byref(e)), "dp_expand_frontier")
        return b.value, e.value

    def expand_last_candidates(self):
        return int(lib().dp_expand_last_candidates(self._h))

    def expand_frontier_weights(self):
        w = np.zeros(max(self.num_views(), 1), np.int64)
        self._ck(lib().dp_expand_frontier_weights(self._h, _ptr(w)), "dp_expand_frontier_weights")
        return w[:self.num_views()]

    def expand_level_commit_gathered(self, gathered_ptr, world, segment_capacity, counts, stream=0):
        counts = np.ascontiguousarray(counts, dtype=np.int64)
        n = C.c_int64(0)
        self._ck(lib().dp_expand_level_commit_gathered(
            self._h, C.c_void_p(gathered_ptr), C.c_int(world), C.c_int64(segment_capacity),
            _ptr(counts), C.byref(n), C.c_void_p(stream)), "dp_expand_level_commit_gathered")
        return n.value

    def expand_level_local(self, cell_size, rank, world, rank_of_view, records_ptr, max_records,
                           stream=0):
        rov = None if rank_of_view is None else np.ascontiguousarray(rank_of_view, dtype=np.int32)
        n = C.c_int64(0)
        self._ck(lib().dp_expand_level_local(self._h, C.c_int(cell_size), C.c_int(rank),
                                             C.c_int(world), _ptr(rov), C.c_void_p(records_ptr),
                                             C.c_int64(max_records), C.byref(n),
                                             C.c_void_p(stream)), "dp_expand_level_local")
        return n.value

    def expand_level_commit(self, records_ptr, n_records, stream=0):
        n = C.c_int64(0)
        self._ck(lib().dp_expand_level_commit(self._h, C.c_void_p(records_ptr),
                                              C.c_int64(n_records), C.byref(n),
                                              C.c_void_p(stream)), "dp_expand_level_commit")
        return n.value

    # ---- device-pointer layer (async) ---------------------------------------------------------
    def score_dev(self, batch: DpPatchDev, cell_size, ncc_ptr, tex_ptr=0, valid_ptr=0, stream=0):
        self._ck(lib().dp_score_dev(self._h, C.byref(batch), C.c_int(cell_size),
                                    C.c_void_p(ncc_ptr), C.c_void_p(tex_ptr) if tex_ptr else None,
                                    C.c_void_p(valid_ptr) if valid_ptr else None,
                                    C.c_void_p(stream)), "dp_score_dev")

    def filter_dev(self, batch: DpPatchDev, cell_size, keep_ptr, stream=0):
        self._ck(lib().dp_filter_dev(self._h, C.byref(batch), C.c_int(cell_size),
                                     C.c_void_p(keep_ptr), C.c_void_p(stream)), "dp_filter_dev")

    def refine_dev(self, batch: DpPatchDev, cell_size, mask_ptr=0, evals_ptr=0, xbest_ptr=0,
                   stream=0):
        self._ck(lib().dp_refine_dev(self._h, C.byref(batch), C.c_int(cell_size),
                                     C.c_void_p(mask_ptr) if mask_ptr else None,
                                     C.c_void_p(evals_ptr) if evals_ptr else None,
                                     C.c_void_p(xbest_ptr) if xbest_ptr else None,
                                     C.c_void_p(stream)), "dp_refine_dev")

    def visibility_dev(self, batch: DpPatchDev, ncand_ptr=0, cand_ptr=0, stream=0):
        self._ck(lib().dp_visibility_dev(self._h, C.byref(batch),
                                         C.c_void_p(ncand_ptr) if ncand_ptr else None,
                                         C.c_void_p(cand_ptr) if cand_ptr else None,
                                         C.c_void_p(stream)), "dp_visibility_dev")

    def color_dev(self, batch: DpPatchDev, stream=0):
        self._ck(lib().dp_color_dev(self._h, C.byref(batch), C.c_void_p(stream)), "dp_color_dev")

    # ---- pyramid ---------------------------------------------------------------------------------
    def build_pyramid(self, n_levels):
        self._ck(lib().dp_build_pyramid(self._h, C.c_int(n_levels)), "dp_build_pyramid")

    def set_level(self, level):
        self._ck(lib().dp_set_level(self._h, C.c_int(level)), "dp_set_level")

    def set_level_selection(self, enable, px_per_cell=1.5):
        """Per-(patch, view) pyramid level (dp_set_level_selection)."""
        self._ck(lib().dp_set_level_selection(self._h, C.c_int(1 if enable else 0),
                                              C.c_double(px_per_cell)), "dp_set_level_selection")

    def download_level(self, view, level):
        w, h = C.c_int(0), C.c_int(0)
        self._ck(lib().dp_download_level(self._h, C.c_int(view), C.c_int(level), None,
                                         C.c_size_t(0), C.byref(w), C.byref(h)),
                 "dp_download_level")
        out = np.zeros((h.value, w.value, 3), np.uint8)
        self._ck(lib().dp_download_level(self._h, C.c_int(view), C.c_int(level), _ptr(out),
                                         C.c_size_t(out.size), C.byref(w), C.byref(h)),
                 "dp_download_level")
        return out

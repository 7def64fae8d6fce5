"""Larger BASELINE configurations as one-off measurements (not the default bench):
  c3   NCC scoring microbench: N patches x 8 forced-visible views, mu = 7
  c4   64-view 1920x1080 plane scene ("room wall"), seeds -> filter -> refine -> expand loop
  c4score  scoring + seed refinement on that scene with the image set (531 MB packed) far
       beyond L2: the HBM-bound regime.  --order random | spatial (patches sorted by reference
       view and Morton code of their pixel in it, a caller-side ordering) shows what locality
       is worth; reports evals/s and algorithmic GB/s against the measured HBM peak.
usage: python tools/scale_cases.py c3 [--patches 10000000] | c4 [--seeds 50000] [--levels -1]
       | c4score [--patches 2000000] [--order spatial] [--cell 7]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from densepoints_b200 import capi, scenes  # noqa: E402


def c3(a):
    import torch
    dev = torch.device("cuda", 0)
    sc = scenes.make_sphere_scene(seed=2, n_views=16, width=1280, height=960, f=1000.0)
    seeds = scenes.make_seeds(sc, a.patches, seed=3)
    nvis, vis = scenes.force_visible(sc, seeds, 8)
    ctx = capi.Context(0)
    ctx.set_views(sc.P, sc.images)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    pos, nrm, ref, nv, vi = t(seeds["pos"]), t(seeds["nrm"]), t(seeds["ref"].astype(np.int32)), t(nvis), t(vis)
    ncc = torch.zeros((a.patches, 8), dtype=torch.float32, device=dev)
    b = capi.dev_batch(a.patches, 8, pos.data_ptr(), nrm.data_ptr(), ref.data_ptr(), nv.data_ptr(), vi.data_ptr())
    st = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ctx.score_dev(b, 7, ncc.data_ptr(), stream=st)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        ctx.score_dev(b, 7, ncc.data_ptr(), stream=st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    evals = a.patches * 8
    print(json.dumps(dict(case="c3", patches=a.patches, views=8, cell=7, ms=ms, evals_per_s=evals / ms * 1e3,
                          out_bytes=a.patches * 8 * 4, ncc_mean=float(ncc[:, 1:].mean().item()))))
    ctx.close()


def c4(a):
    sc = scenes.make_plane_scene(seed=4, n_views=a.views, width=a.width, height=a.width * 9 // 16,
                                 yaw_spread_deg=25.0, name="C4")
    seeds = scenes.make_seeds(sc, a.seeds, seed=40, depth_noise=0.003, tilt_deg=5.0)
    ctx = capi.Context(0)
    t0 = time.perf_counter()
    ctx.set_views(sc.P, sc.images)
    t_up = time.perf_counter() - t0
    t0 = time.perf_counter()
    nvis, vis, _, _ = ctx.visibility(seeds["pos"], seeds["nrm"], seeds["ref"])
    keep, fnvis, fvis, pos, nrm, evals = ctx.filter_refine(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis, 16)
    t_seed = time.perf_counter() - t0
    m = keep.astype(bool)
    ctx.organizer_reset()
    t0 = time.perf_counter()
    acc = ctx.organizer_insert(pos[m], nrm[m], seeds["ref"][m], fnvis[m], fvis[m])
    st = ctx.expand(11, a.levels)
    t_exp = time.perf_counter() - t0
    print(json.dumps(dict(case="c4", views=a.views, width=a.width, seeds=a.seeds, upload_s=t_up,
                          seed_filter_refine_s=t_seed, kept=int(m.sum()), seeded=int(acc.sum()),
                          expand_s=t_exp, expand=st, patches=ctx.organizer_size(),
                          mean_nvis=float(nvis.mean()))))
    ctx.close()


def _morton(x, y):
    def spread(v):
        v = v.astype(np.uint64) & 0xFFFF
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return spread(x) | (spread(y) << 1)


def c4score(a):
    import torch
    dev = torch.device("cuda", 0)
    t0 = time.perf_counter()
    sc = scenes.make_plane_scene(seed=4, n_views=a.views, width=a.width, height=a.width * 9 // 16,
                                 yaw_spread_deg=25.0, name="C4", device="cuda:0")
    t_scene = time.perf_counter() - t0
    seeds = scenes.make_seeds(sc, a.patches, seed=41, depth_noise=0.003, tilt_deg=5.0)
    pos, nrm, ref = seeds["pos"], seeds["nrm"], seeds["ref"].astype(np.int32)
    if a.order == "spatial":
        P = sc.P[ref]
        X = np.concatenate([pos.astype(np.float64), np.ones((len(pos), 1))], 1)
        uvw = np.einsum("nij,nj->ni", P, X)
        u = np.clip(uvw[:, 0] / uvw[:, 2], 0, sc.width - 1).astype(np.int64) >> 3
        v = np.clip(uvw[:, 1] / uvw[:, 2], 0, sc.height - 1).astype(np.int64) >> 3
        key = (ref.astype(np.uint64) << np.uint64(32)) | _morton(u, v)
        o = np.argsort(key, kind="stable")
        pos, nrm, ref = pos[o], nrm[o], ref[o]
    ctx = capi.Context(0)
    ctx.set_views(sc.P, sc.images)
    n, V = a.patches, sc.n_views
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    d_pos0, d_nrm0, d_ref = t(pos), t(nrm), t(ref)
    nvis0 = torch.zeros(n, dtype=torch.int32, device=dev)
    vis0 = torch.full((n, V), -1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ctx.visibility_dev(capi.dev_batch(n, V, d_pos0.data_ptr(), d_nrm0.data_ptr(), d_ref.data_ptr(),
                                      nvis0.data_ptr(), vis0.data_ptr()), stream=st)
    d_pos, d_nrm, nvis, vis = (torch.empty_like(x) for x in (d_pos0, d_nrm0, nvis0, vis0))
    ncc = torch.zeros((n, V), dtype=torch.float32, device=dev)
    keep = torch.zeros(n, dtype=torch.uint8, device=dev)
    evals = torch.zeros(n, dtype=torch.int32, device=dev)
    wb = capi.dev_batch(n, V, d_pos.data_ptr(), d_nrm.data_ptr(), d_ref.data_ptr(), nvis.data_ptr(),
                        vis.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    best = [1e30] * 3
    for rep in range(a.reps):
        flush.zero_()
        d_pos.copy_(d_pos0); d_nrm.copy_(d_nrm0); nvis.copy_(nvis0); vis.copy_(vis0)
        e[0].record()
        ctx.score_dev(wb, a.cell, ncc.data_ptr(), stream=st)
        e[1].record()
        ctx.filter_dev(wb, a.cell, keep.data_ptr(), stream=st)
        e[2].record()
        ctx.refine_dev(wb, a.cell, mask_ptr=keep.data_ptr(), evals_ptr=evals.data_ptr(), stream=st)
        e[3].record()
        torch.cuda.synchronize()
        best = [min(b, e[i].elapsed_time(e[i + 1])) for i, b in enumerate(best)]
    ev_score = int(nvis0.sum().item())
    ev_ref = int((evals.long() * nvis.long() * keep.long()).sum().item())
    mean_nv = ev_score / max(n, 1)
    b_alg = 3.0 * (2 * (a.cell // 2) + 2) ** 2 + 4.0 + (28.0 + 2.0 * mean_nv) / max(mean_nv, 1)
    peak = 6549.4
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(
            os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gbs = lambda ev, ms: ev * b_alg / (ms * 1e-3) / 1e9
    print(json.dumps(dict(case="c4score", order=a.order, views=V, width=sc.width, height=sc.height,
                          image_set_mb=V * sc.height * ((sc.width + 31) // 32 * 32) * 4 / 1e6,
                          patches=n, cell=a.cell, mean_visible=mean_nv, scene_s=t_scene,
                          score_ms=best[0], score_evals_per_s=ev_score / best[0] * 1e3,
                          score_alg_gbs=gbs(ev_score, best[0]), score_frac_hbm=gbs(ev_score, best[0]) / peak,
                          filter_ms=best[1], refined=int(keep.sum().item()), refine_ms=best[2],
                          refine_evals_per_s=ev_ref / best[2] * 1e3,
                          refine_alg_gbs=gbs(ev_ref, best[2]), refine_frac_hbm=gbs(ev_ref, best[2]) / peak,
                          alg_bytes_per_eval=b_alg, hbm_peak_gbs=peak)))
    ctx.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("case", choices=["c3", "c4", "c4score"])
    ap.add_argument("--patches", type=int, default=10_000_000)
    ap.add_argument("--seeds", type=int, default=50_000)
    ap.add_argument("--levels", type=int, default=-1)
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--order", choices=["random", "spatial"], default="random")
    ap.add_argument("--cell", type=int, default=7)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    if a.case == "c4score" and a.patches == 10_000_000:
        a.patches = 2_000_000
    {"c3": c3, "c4": c4, "c4score": c4score}[a.case](a)

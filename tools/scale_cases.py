"""Larger BASELINE configurations as one-off measurements (not the default bench):
  c3   NCC scoring microbench: N patches x 8 forced-visible views, mu = 7
  c4   64-view 1920x1080 plane scene ("room wall"), seeds -> filter -> refine -> expand loop
usage: python tools/scale_cases.py c3 [--patches 10000000] | c4 [--seeds 50000] [--levels -1]"""
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


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("case", choices=["c3", "c4"])
    ap.add_argument("--patches", type=int, default=10_000_000)
    ap.add_argument("--seeds", type=int, default=50_000)
    ap.add_argument("--levels", type=int, default=-1)
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--width", type=int, default=1920)
    a = ap.parse_args()
    {"c3": c3, "c4": c4}[a.case](a)

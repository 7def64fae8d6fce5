"""Evaluation-count distribution of the bench workload and refine throughput of the lane / group
kernels as a function of the evaluation cap (how much of the refine time is tail)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densepoints_b200 import capi, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
full = "--full-res" in sys.argv
dev = torch.device("cuda", 0)
sc = scenes.make_sphere_scene(seed=2, n_views=16, width=1280 if full else 640, height=960 if full else 480,
                              f=1000.0 if full else 500.0)
seeds = scenes.make_seeds(sc, n, seed=200)
V = sc.n_views
st = torch.cuda.current_stream().cuda_stream
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
pos0, nrm0, ref = t(seeds["pos"]), t(seeds["nrm"]), t(seeds["ref"].astype(np.int32))
for cap in (500, 250, 128, 64):
    ctx = capi.Context(0, capi.default_params(nm_max_evals=cap))
    ctx.set_views(sc.P, sc.images)
    nvis0 = torch.zeros(n, dtype=torch.int32, device=dev)
    vis0 = torch.full((n, V), -1, dtype=torch.int32, device=dev)
    ctx.visibility_dev(capi.dev_batch(n, V, pos0.data_ptr(), nrm0.data_ptr(), ref.data_ptr(),
                                      nvis0.data_ptr(), vis0.data_ptr()), stream=st)
    pos, nrm, nvis, vis = (torch.empty_like(x) for x in (pos0, nrm0, nvis0, vis0))
    keep = torch.zeros(n, dtype=torch.uint8, device=dev)
    evals = torch.zeros(n, dtype=torch.int32, device=dev)
    wb = capi.dev_batch(n, V, pos.data_ptr(), nrm.data_ptr(), ref.data_ptr(), nvis.data_ptr(), vis.data_ptr())
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e30
    for rep in range(3):
        pos.copy_(pos0); nrm.copy_(nrm0); nvis.copy_(nvis0); vis.copy_(vis0)
        ctx.filter_dev(wb, 7, keep.data_ptr(), stream=st)
        e[0].record()
        ctx.refine_dev(wb, 7, mask_ptr=keep.data_ptr(), evals_ptr=evals.data_ptr(), stream=st)
        e[1].record()
        torch.cuda.synchronize()
        best = min(best, e[0].elapsed_time(e[1]))
    ev = evals.cpu().numpy()[keep.cpu().numpy().astype(bool)]
    nv = nvis.cpu().numpy()[keep.cpu().numpy().astype(bool)]
    tot = int((ev.astype(np.int64) * nv).sum())
    print(f"{os.environ.get('DP_REFINE_KERNEL', 'lane'):6s} cap {cap:3d}: refine {best:8.3f} ms, {tot / best / 1e6:6.3f} Gev/s, "
          f"{len(ev)} patches, evals mean {ev.mean():.1f} pct50/90/99/99.9 "
          f"{np.percentile(ev, [50, 90, 99, 99.9]).tolist()} max {ev.max()}, "
          f"share of evals beyond 64/128/250: {[(np.maximum(ev - k, 0) * nv).sum() / tot for k in (64, 128, 250)]}",
          flush=True)
    ctx.close()

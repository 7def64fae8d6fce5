"""Time tuning builds of the CUDA library against each other in one process (one scene, one
GPU): filter + refine on the bench workload shape.  The first library is the reference for a
bit-equality check of the outputs (ablation builds are expected to differ).

With --check-cells the host-buffer entry points are also run on a small batch for other cell
sizes (textures, scores, filter results, refined geometry where s <= 8) and compared with the
first library: that covers the other template instantiations a change of the shared device
code touches.

usage: python tools/variant_sweep.py [--seeds N] [--full-res] [--reps R] [--cell S]
                                     [--check-cells 3,5,11,16] LIB.so [LIB.so ...]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from densepoints_b200 import capi, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=262144)
    ap.add_argument("--cell", type=int, default=7)
    ap.add_argument("--full-res", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check-cells", default="")
    ap.add_argument("--check-seeds", type=int, default=4000)
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    if a.full_res:
        sc = scenes.make_sphere_scene(seed=2, n_views=16, width=1280, height=960, f=1000.0)
    else:
        sc = scenes.make_sphere_scene(seed=2, n_views=16, width=640, height=480, f=500.0)
    seeds = scenes.make_seeds(sc, a.seeds, seed=200)
    n, V = a.seeds, sc.n_views
    st = torch.cuda.current_stream().cuda_stream
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    pos0, nrm0, ref = t(seeds["pos"]), t(seeds["nrm"]), t(seeds["ref"].astype(np.int32))
    want = None
    check_cells = [int(x) for x in a.check_cells.split(",") if x]
    cseeds = scenes.make_seeds(sc, a.check_seeds, seed=201)
    cwant = {}
    for path in a.libs:
        if not os.path.exists(path):
            print(f"{os.path.basename(path):24s} missing, skipped", flush=True)
            continue
        os.environ["DENSEPOINTS_CUDA_LIB"] = os.path.abspath(path)
        capi._lib = None
        ctx = capi.Context(0)
        ctx.set_views(sc.P, sc.images)
        nvis0 = torch.zeros(n, dtype=torch.int32, device=dev)
        vis0 = torch.full((n, V), -1, dtype=torch.int32, device=dev)
        ctx.visibility_dev(capi.dev_batch(n, V, pos0.data_ptr(), nrm0.data_ptr(), ref.data_ptr(),
                                          nvis0.data_ptr(), vis0.data_ptr()), stream=st)
        pos, nrm, nvis, vis = (torch.empty_like(x) for x in (pos0, nrm0, nvis0, vis0))
        keep = torch.zeros(n, dtype=torch.uint8, device=dev)
        evals = torch.zeros(n, dtype=torch.int32, device=dev)
        ncc = torch.zeros((n, V), dtype=torch.float32, device=dev)
        wb = capi.dev_batch(n, V, pos.data_ptr(), nrm.data_ptr(), ref.data_ptr(), nvis.data_ptr(),
                            vis.data_ptr())
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        best = [1e30, 1e30, 1e30]
        for rep in range(a.reps):
            pos.copy_(pos0); nrm.copy_(nrm0); nvis.copy_(nvis0); vis.copy_(vis0)
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
        got = [x.cpu().numpy().copy() for x in (ncc, keep, nvis, vis, pos, nrm, evals)]
        same = "ref"
        if want is None:
            want = got
        else:
            same = "EQUAL" if all(np.array_equal(g, w) for g, w in zip(got, want)) else "differs"
        print(f"{os.path.basename(path):24s} score {best[0]:7.3f} ms {ev_score / best[0] / 1e6:6.3f} Gev/s | "
              f"filter {best[1]:7.3f} ms | refine {best[2]:8.3f} ms {ev_ref / best[2] / 1e6:6.3f} Gev/s "
              f"({ev_ref} evals) | outputs {same}", flush=True)
        for cs in check_cells:
          try:
            cnv, cvi, _, _ = ctx.visibility(cseeds["pos"], cseeds["nrm"], cseeds["ref"])
            r_ncc, r_tex, r_valid = ctx.score(cseeds["pos"], cseeds["nrm"], cseeds["ref"], cnv, cvi, cs,
                                              want_tex=True)
            r_keep, r_fnv, r_fvi = ctx.filter(cseeds["pos"], cseeds["nrm"], cseeds["ref"], cnv, cvi, cs)
            res = [r_ncc, r_tex, r_valid, r_keep, r_fnv, r_fvi]
            if cs <= 8:
                m = r_keep.astype(bool)
                res += list(ctx.refine(cseeds["pos"][m], cseeds["nrm"][m], cseeds["ref"][m], r_fnv[m],
                                       r_fvi[m], cs))
            if cs not in cwant:
                cwant[cs] = res
                st_ = "ref"
            else:
                st_ = "EQUAL" if all(np.array_equal(g, w) for g, w in zip(res, cwant[cs])) else "DIFFERS"
            print(f"    cell {cs:2d}: {len(res)} arrays {st_}", flush=True)
          except Exception as ex:  # keep the sweep alive: the timing lines matter most
            import traceback
            traceback.print_exc()
            print(f"    cell {cs:2d}: check failed: {ex}", flush=True)
        ctx.close()


if __name__ == "__main__":
    main()

"""Multi-GPU expansion check (run under torchrun, one rank per GPU): every rank holds all
views and the same seeds, expansion is sharded by reference image with one NCCL allgather
per BFS level (densepoints_b200/distributed.py).  Verifies that every rank ends with the
store / grids of the single-GPU dp_expand, and reports timing.

torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
    tools/multigpu_expand_check.py [--seeds 4000] [--levels 3] [--views 8] [--width 640]
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from densepoints_b200 import capi, scenes  # noqa: E402
from densepoints_b200 import distributed as dd  # noqa: E402


def digest(ctx, n_views):
    ex = ctx.organizer_export()
    h = hashlib.sha256()
    for k in ("pos", "nrm", "rgb", "ref", "nvis", "vis"):
        h.update(np.ascontiguousarray(ex[k]).tobytes())
    for v in range(n_views):
        h.update(ctx.organizer_grid(v).tobytes())
    return h.hexdigest(), len(ex["ref"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=4000)
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--cell", type=int, default=11)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h = a.width * 3 // 4
    # each rank renders a slice of the views; the images are then exchanged (replicated views)
    sc = scenes.make_plane_scene(seed=4, n_views=a.views, width=a.width, height=h,
                                 yaw_spread_deg=20.0,
                                 only_views=dd.views_of_rank(a.views, rank, world))
    dd.share_images(sc.images, rank, world, dev)
    seeds = scenes.make_seeds(sc, a.seeds, seed=40, depth_noise=0.003, tilt_deg=4.0)
    ctx = capi.Context(local)
    ctx.set_views(sc.P, sc.images)
    nvis, vis, _, _ = ctx.visibility(seeds["pos"], seeds["nrm"], seeds["ref"])
    pos, nrm, _, _ = ctx.refine(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis, 7)

    def seed_organizer():
        ctx.organizer_reset()
        ctx.organizer_insert(pos, nrm, seeds["ref"], nvis, vis)

    # single-GPU reference on this rank
    seed_organizer()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st1 = ctx.expand(a.cell, a.levels)
    torch.cuda.synchronize()
    t_single = time.perf_counter() - t0
    want, n_want = digest(ctx, sc.n_views)

    # sharded run
    seed_organizer()
    rov = dd.partition_views(seeds["ref"], sc.n_views, world)
    be = dd.CudaLevelBackend(ctx, dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = dd.expand_distributed(be, a.cell, a.levels, rank, world, rov)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_multi = time.perf_counter() - t0
    got, n_got = digest(ctx, sc.n_views)
    ok = (got == want)
    flags = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    res = dict(world=world, ok_all_ranks=bool(flags.item()), patches=n_got, patches_single=n_want,
               seeds=a.seeds, levels=st["levels"], pops=st["pops"], passed=st["passed"],
               inserted=st["inserted"], local_records_rank0=st["local_records"],
               single_gpu_s=t_single, sharded_s=t_multi, rank_of_view=rov.tolist(),
               single_stats=st1)
    if rank == 0:
        print(json.dumps(res), flush=True)
        if a.out:
            json.dump(res, open(a.out, "w"))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if flags.item() else 1)


if __name__ == "__main__":
    main()

"""Where does a sharded expansion level lose time?  Replays bench.py's 64-view expansion with
the WORLD ranks of the sharded run emulated one after the other on ONE GPU (dp_expand_level_local
needs no collective: the store is replicated, the cut is computed from it), timing every rank's
local step of every level, and prints per level: the N = 1 time, each virtual rank's time,
candidates and records.  The store hash must equal the bench's.

usage: python tools/shard_balance.py [--world 8] [--levels 12] [--out FILE.json]"""
import argparse, hashlib, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from densepoints_b200 import capi, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--levels", type=int, default=12)
ap.add_argument("--out", default="")
ap.add_argument("--c5", action="store_true", help="the C5 scene (256 views 3840x2160, 200 000 seeds) instead")
a = ap.parse_args()
dev = torch.device("cuda", 0)
if a.c5:
    W, H = 3840, 2160
    P, centers, Rs, f, cx, cy, extent, tex = scenes.lattice_plane_cameras(nx=16, ny=16, width=W, height_px=H, f=3000.0)
    ctx = capi.Context(0)
    ctx.set_num_views(len(P))
    for v in range(len(P)):
        img = scenes.render_plane_view(centers[v], Rs[v], f, cx, cy, W, H, extent, tex, dev)
        ctx.upload_view(v, P[v], img)
    sc = scenes.Scene("C5", P, [np.zeros((1, 1, 3), np.uint8)] * len(P), W, H, "plane", extent=extent, centers=centers)
    sc.extent = 0.5 * 15 * 4.0 / 0.8
    seeds = scenes.make_seeds(sc, 200_000, seed=50, depth_noise=0.003, tilt_deg=5.0)
else:
    sc = scenes.make_plane_scene(seed=4, n_views=64, width=1920, height=1080, yaw_spread_deg=25.0,
                                 name="C4", device="cuda:0")
    seeds = scenes.make_seeds(sc, 50_000, seed=40, depth_noise=0.003, tilt_deg=5.0)
    ctx = capi.Context(0)
    ctx.set_views(sc.P, sc.images)
nvis, vis, _, _ = ctx.visibility(seeds["pos"], seeds["nrm"], seeds["ref"])
keep, fnvis, fvis, pos, nrm, evs = ctx.filter_refine(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis, 16)
m = keep.astype(bool)
words = ctx.record_bytes() // 4
st = torch.cuda.current_stream().cuda_stream


def run(world):
    ctx.organizer_reset()
    ctx.organizer_insert(pos[m], nrm[m], seeds["ref"][m], fnvis[m], fvis[m])
    rows = []
    for lvl in range(a.levels):
        fb, fe = ctx.expand_frontier()
        nf = fe - fb
        if nf <= 0:
            break
        bufs, counts, ms, cands = [], [], [], []
        for r in range(world):
            buf = torch.zeros((max(4 * nf, 1), words), dtype=torch.int32, device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = ctx.expand_level_local(11, r, world, None, buf.data_ptr(), 4 * nf, stream=st)
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
            counts.append(n)
            cands.append(ctx.expand_last_candidates())
            bufs.append(buf)
        cap = max(max(counts), 1)
        recv = torch.cat([b[:cap] for b in bufs]).contiguous()
        ctx.expand_level_commit_gathered(recv.data_ptr(), world, cap, counts, stream=st)
        torch.cuda.synchronize()
        rows.append(dict(level=lvl, frontier=nf, ms=ms, candidates=cands, records=counts))
    ex = ctx.organizer_export()
    h = hashlib.sha256()
    for k in ("pos", "nrm", "rgb", "ref", "nvis", "vis"):
        h.update(np.ascontiguousarray(ex[k]).tobytes())
    h.update(ctx.organizer_grids().tobytes())
    return rows, h.hexdigest()


run(1)
one, h1 = run(1)
run(a.world)
many, hw = run(a.world)
print("store hash N=1", h1[:16], f"N={a.world}", hw[:16], "equal" if h1 == hw else "DIFFER")
tot1 = totw = ideal = 0.0
for r1, rw in zip(one, many):
    t1 = r1["ms"][0]
    tw = rw["ms"]
    tot1 += t1; totw += max(tw); ideal += t1 / a.world
    print(f"level {r1['level']:2d} frontier {r1['frontier']:6d}  N=1 {t1:7.2f} ms ({r1['candidates'][0]} cand)  "
          f"ideal {t1 / a.world:6.2f}  ranks max {max(tw):6.2f} min {min(tw):6.2f} sum {sum(tw):7.2f} | "
          + " ".join(f"{x:5.2f}" for x in tw) + " | cand " + " ".join(str(c) for c in rw["candidates"]))
print(f"sum of local steps: N=1 {tot1:.1f} ms, ideal/{a.world} {ideal:.1f}, slowest virtual rank per level {totw:.1f} "
      f"-> {tot1 / totw:.2f}x")
if a.out:
    json.dump(dict(world=a.world, one=one, many=many, hash1=h1, hashw=hw), open(a.out, "w"))

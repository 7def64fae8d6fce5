"""Debug harness: run the level loop once with world=1 semantics and once sharded, compare the
store after every level, report the first divergence."""
import argparse, hashlib, json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from densepoints_b200 import capi, scenes
from densepoints_b200 import distributed as dd

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", type=int, default=60000); ap.add_argument("--levels", type=int, default=5)
ap.add_argument("--views", type=int, default=32); ap.add_argument("--width", type=int, default=640)
ap.add_argument("--cell", type=int, default=11)
ap.add_argument("--full-texture", action="store_true")
ap.add_argument("--dist-render", action="store_true")
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
_T = scenes.Texture3D
class _LightTex(_T):          # cheaper texture: this harness is about bookkeeping, not photometry
    def __init__(self, seed, wavelength, n_base=3, n_chan=1):
        super().__init__(seed, wavelength, n_base=n_base, n_chan=n_chan)
if not a.full_texture:
    scenes.Texture3D = _LightTex
if a.dist_render and world > 1:
    # every rank renders views r::world (and view 0, to compare renderings across ranks), then
    # the images are exchanged, so all ranks hold bit-identical images by construction
    orig_render = scenes._render
    mine = {}
    def part_render(Ps, centers, Rs, f, cx, cy, w, h, surface, param, tex, chunk=1 << 18):
        idx = sorted(set(list(range(rank, len(Ps), world)) + [0]))
        imgs = orig_render([Ps[i] for i in idx], [centers[i] for i in idx], [Rs[i] for i in idx], f, cx, cy, w, h, surface, param, tex, chunk)
        out = [np.zeros((h, w, 3), np.uint8) for _ in Ps]
        for i, im in zip(idx, imgs): out[i] = im
        return out
    scenes._render = part_render
sc = scenes.make_plane_scene(seed=4, n_views=a.views, width=a.width, height=a.width * 3 // 4, yaw_spread_deg=20.0)
if a.dist_render and world > 1:
    h0 = hashlib.sha256(sc.images[0].tobytes()).hexdigest()[:16]
    hs = [None] * world
    dist.all_gather_object(hs, h0)
    if rank == 0: print("view-0 rendering digests per rank:", hs, flush=True)
    for v in range(a.views):
        t = torch.from_numpy(sc.images[v]).to(dev)
        dist.broadcast(t, src=v % world)
        sc.images[v] = t.cpu().numpy()
seeds = scenes.make_seeds(sc, a.seeds, seed=40, depth_noise=0.003, tilt_deg=4.0)
ctx = capi.Context(local); ctx.set_views(sc.P, sc.images)
nvis, vis, _, _ = ctx.visibility(seeds["pos"], seeds["nrm"], seeds["ref"])
pos, nrm, _, _ = ctx.refine(seeds["pos"], seeds["nrm"], seeds["ref"], nvis, vis, 7)
rov = dd.partition_views(seeds["ref"], sc.n_views, world)
be = dd.CudaLevelBackend(ctx, dev)

def run(w, r, rv):
    ctx.organizer_reset(); ctx.organizer_insert(pos, nrm, seeds["ref"], nvis, vis)
    out = []
    for lvl in range(a.levels):
        fb, fe = be.frontier(); nf = fe - fb
        if nf <= 0: break
        buf, n_local = be.local(a.cell, r, w, rv, 4 * nf)
        if w > 1:
            rec, total = dd.gather_records(buf, n_local, w)
        else:
            rec, total = buf[:n_local], n_local
        recs = rec[:total].cpu().numpy().copy()
        ins = be.commit(rec, total)
        out.append((nf, total, ins, recs, ctx.organizer_export()))
    return out

single = run(1, 0, None)
shard = run(world, rank, rov)
if rank == 0:
    for lvl, (s, m) in enumerate(zip(single, shard)):
        rs = s[3][np.argsort(s[3][:, 0])]; rm = m[3][np.argsort(m[3][:, 0])]
        same_rec = rs.shape == rm.shape and np.array_equal(rs, rm)
        same_store = all(np.array_equal(s[4][k], m[4][k]) for k in s[4])
        print(f"level {lvl}: nf {s[0]}/{m[0]} records {s[1]}/{m[1]} inserted {s[2]}/{m[2]} records_equal {same_rec} store_equal {same_store}", flush=True)
        if not same_rec:
            ss, sm_ = set(rs[:, 0].tolist()), set(rm[:, 0].tolist())
            print("  seq only in single:", sorted(ss - sm_)[:10], len(ss - sm_), " only in sharded:", sorted(sm_ - ss)[:10], len(sm_ - ss))
            common = sorted(ss & sm_)
            ds = {int(r[0]): r for r in rs}; dm = {int(r[0]): r for r in rm}
            diff = [q for q in common if not np.array_equal(ds[q], dm[q])]
            print("  common seq with different payload:", len(diff), diff[:5])
            for q in (sorted(ss - sm_)[:3] + diff[:3]):
                par = q // 4
                print("   seq", q, "parent ref", int(s[4]["ref"][0]) if False else "", "single rec", ds.get(q, None)[:12] if q in ds else None, "sharded", dm.get(q, None)[:12] if q in dm else None)
            break
ctx.close()
if world > 1: dist.destroy_process_group()

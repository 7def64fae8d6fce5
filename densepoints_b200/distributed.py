"""Multi-GPU expansion driver: one process per GPU, patches sharded by reference image,
one allgather of candidate records per BFS level.

torch.distributed is plumbing only (NCCL over NVLink on GPUs, gloo in the CPU tests); all
the work of a level happens behind the C ABI:

    dp_expand_level_local   refine + visibility + filter of the candidates this rank owns,
                            survivors written as records into the send buffer
    all_gather              counts, then the (padded) record buffers
    dp_expand_level_commit_gathered
                            every rank compacts the gathered segments on the device and replays
                            TryInsert over all records -> identical grids and stores everywhere
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partition_views(ref, n_views: int, world: int) -> np.ndarray:
    """rank_of_view: contiguous blocks of reference-image ids balanced by patch count
    (children inherit the parent's reference image, reference expand.cpp:126, so ownership
    is stable across levels)."""
    counts = np.bincount(np.asarray(ref, dtype=np.int64), minlength=n_views).astype(np.float64)
    total = counts.sum()
    rov = np.zeros(n_views, np.int32)
    if world <= 1 or total == 0:
        if world > 1:
            rov[:] = (np.arange(n_views) * world // max(n_views, 1)).astype(np.int32)
        return rov
    cum = np.cumsum(counts) - counts / 2.0          # centre of mass of each view's block
    rov[:] = np.minimum((cum / total * world).astype(np.int32), world - 1)
    rov = np.maximum.accumulate(rov)                 # contiguous, non-decreasing
    return rov


def views_of_rank(n_views: int, rank: int, world: int):
    """The views rank `rank` renders / loads before share_images: v % world == rank."""
    return list(range(rank, n_views, world))


def share_images(images, rank: int, world: int, device=None):
    """Every rank loaded (or rendered) only views_of_rank(...); broadcast each image from its
    owner so that all ranks hold the full, bit-identical image set (views are replicated,
    SURVEY 8e).  In place; returns the list."""
    if world <= 1:
        return images
    for v in range(len(images)):
        t = torch.from_numpy(np.ascontiguousarray(images[v]))
        if device is not None:
            t = t.to(device)
        dist.broadcast(t, src=v % world)
        images[v] = t.cpu().numpy()
    return images


def assign_views(weights, world: int) -> np.ndarray:
    """rank_of_view for ONE level from the frontier's work per reference view
    (dp_expand_frontier_weights): longest-processing-time-first -- views by descending weight
    (ties: ascending view id), each to the least loaded rank so far (ties: lowest rank).  Every
    rank computes the same table from the replicated store, so no exchange is needed.  Patches
    still shard by reference image (children inherit it, reference expand.cpp:126); only which
    rank serves a view may change from level to level, which costs nothing because views, grids
    and store are replicated."""
    w = np.asarray(weights, dtype=np.int64)
    rov = np.zeros(len(w), np.int32)
    if world <= 1:
        return rov
    load = np.zeros(world, np.int64)
    for v in sorted(range(len(w)), key=lambda i: (-int(w[i]), i)):
        r = int(np.argmin(load))
        rov[v] = r
        load[r] += int(w[v])
    return rov


class CudaLevelBackend:
    """The per-level steps on a dp_context (device records = torch int32 tensors)."""

    def __init__(self, ctx, device):
        self.ctx = ctx
        self.device = device
        self.words = ctx.record_bytes() // 4
        self._buf = None

    def frontier(self):
        return self.ctx.expand_frontier()

    def frontier_weights(self):
        return self.ctx.expand_frontier_weights()

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def local(self, cell_size, rank, world, rank_of_view, max_records):
        if self._buf is None or self._buf.shape[0] < max(max_records, 1):
            self._buf = torch.empty((max(max_records, 1), self.words), dtype=torch.int32,
                                    device=self.device)
        n = self.ctx.expand_level_local(cell_size, rank, world, rank_of_view, self._buf.data_ptr(),
                                        max_records, stream=self._stream())
        return self._buf, n

    def last_candidates(self):
        return self.ctx.expand_last_candidates()

    def commit_gathered(self, recv, world, capacity, counts):
        return self.ctx.expand_level_commit_gathered(recv.data_ptr(), world, capacity, counts,
                                                     stream=self._stream())

    def mark(self):
        """A timing mark on the stream the level's work is launched on."""
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    @staticmethod
    def elapsed_ms(a, b):
        b.synchronize()
        return a.elapsed_time(b)


def gather_records(buf: torch.Tensor, n_local: int, world: int):
    """allgather of ragged record lists: the counts (one tiny collective, read on the host
    because the commit sizes its launches from them), then the buffers padded to the level's
    maximum.  Returns (gathered [world * capacity, words], counts list, capacity); the ragged
    segments are compacted on the device by dp_expand_level_commit_gathered."""
    if world <= 1:
        return buf, [n_local], max(int(buf.shape[0]), 1)
    cnt = torch.tensor([n_local], dtype=torch.int64, device=buf.device)
    counts = torch.empty(world, dtype=torch.int64, device=buf.device)
    dist.all_gather_into_tensor(counts, cnt)
    counts = counts.cpu().tolist()
    cap = max(max(counts), 1)
    recv = torch.empty((world * cap, buf.shape[1]), dtype=buf.dtype, device=buf.device)
    send = buf[:cap]                                    # rows beyond n_local are padding
    if send.shape[0] < cap:                             # (a backend with an exact-size buffer)
        send = torch.cat([send, send.new_zeros((cap - send.shape[0], send.shape[1]))])
    dist.all_gather_into_tensor(recv, send.contiguous())
    return recv, counts, cap


def expand_distributed(backend, cell_size: int, max_levels: int, rank: int, world: int,
                       rank_of_view=None, timings=None, ownership="views") -> dict:
    """Expand::ExpandPatches (reference expand.cpp:34-101) across `world` ranks.  The
    organizer must hold the same seeds on every rank (dp_organizer_insert on each).
    Ownership of a level's parents:
      rank_of_view given      a fixed table by reference image (partition_views);
      ownership == "views"    by reference image, re-balanced every level from the frontier's
                              work per view (assign_views) -- bounded by the heaviest view;
      ownership == "ranges"   `world` contiguous ranges of the frontier with equal work, cut
                              inside the library (dp_expand_level_local with a NULL table).
    timings (optional dict): per level lists local_ms / allgather_ms / commit_ms and the record
    counts, when the backend can time its stream."""
    stats = dict(levels=0, pops=0, passed=0, inserted=0, local_records=0)
    level = 0
    can_time = timings is not None and hasattr(backend, "mark")
    if timings is not None:
        for k in ("local_ms", "allgather_ms", "commit_ms", "frontier", "records", "local_records",
                  "local_candidates"):
            timings.setdefault(k, [])
    while max_levels < 0 or level < max_levels:
        fb, fe = backend.frontier()
        nf = fe - fb
        if nf <= 0:
            break
        rov = rank_of_view
        if rov is None and world > 1 and ownership == "views":
            rov = assign_views(backend.frontier_weights(), world)
        t0 = backend.mark() if can_time else None
        buf, n_local = backend.local(cell_size, rank, world, rov, 4 * nf)
        t1 = backend.mark() if can_time else None
        recv, counts, cap = gather_records(buf, n_local, world)
        t2 = backend.mark() if can_time else None
        inserted = backend.commit_gathered(recv, world, cap, counts)
        t3 = backend.mark() if can_time else None
        total = int(sum(counts))
        stats["levels"] += 1
        stats["pops"] += nf
        stats["passed"] += total
        stats["inserted"] += inserted
        stats["local_records"] += n_local
        if timings is not None:
            timings["frontier"].append(int(nf))
            timings["records"].append(total)
            timings["local_records"].append(int(n_local))
            if hasattr(backend, "last_candidates"):
                timings["local_candidates"].append(int(backend.last_candidates()))
            if can_time:
                timings["local_ms"].append(backend.elapsed_ms(t0, t1))
                timings["allgather_ms"].append(backend.elapsed_ms(t1, t2))
                timings["commit_ms"].append(backend.elapsed_ms(t2, t3))
        level += 1
    return stats

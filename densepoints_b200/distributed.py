"""Multi-GPU expansion driver: one process per GPU, patches sharded by reference image,
one allgather of candidate records per BFS level.

torch.distributed is plumbing only (NCCL over NVLink on GPUs, gloo in the CPU tests); all
the work of a level happens behind the C ABI:

    dp_expand_level_local   refine + visibility + filter of the candidates this rank owns,
                            survivors written as records into the send buffer
    all_gather              counts, then the (padded) record buffers
    dp_expand_level_commit  every rank replays TryInsert over all records in sequence order
                            -> identical grids and patch stores on every rank
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


class CudaLevelBackend:
    """The three per-level steps on a dp_context (device records = torch int32 tensors)."""

    def __init__(self, ctx, device):
        self.ctx = ctx
        self.device = device
        self.words = ctx.record_bytes() // 4

    def frontier(self):
        return self.ctx.expand_frontier()

    def local(self, cell_size, rank, world, rank_of_view, max_records):
        buf = torch.empty((max(max_records, 1), self.words), dtype=torch.int32, device=self.device)
        n = self.ctx.expand_level_local(cell_size, rank, world, rank_of_view, buf.data_ptr(),
                                        max_records,
                                        stream=torch.cuda.current_stream().cuda_stream)
        return buf, n

    def commit(self, records, n_records):
        return self.ctx.expand_level_commit(records.data_ptr() if n_records else 0, n_records,
                                            stream=torch.cuda.current_stream().cuda_stream)


def gather_records(buf: torch.Tensor, n_local: int, world: int):
    """allgather of ragged record lists: counts first, then buffers padded to the level's
    maximum; returns (records [total, words] in rank order, total)."""
    if world <= 1:
        return buf[:n_local], n_local
    cnt = torch.tensor([n_local], dtype=torch.int64, device=buf.device)
    counts = torch.empty(world, dtype=torch.int64, device=buf.device)
    dist.all_gather_into_tensor(counts, cnt)
    counts = counts.cpu().tolist()
    mx = max(counts)
    if mx == 0:
        return buf[:0], 0
    send = buf[:mx] if buf.shape[0] >= mx else torch.cat(
        [buf, buf.new_zeros((mx - buf.shape[0], buf.shape[1]))])
    send = send.contiguous()
    recv = torch.empty((world * mx, buf.shape[1]), dtype=buf.dtype, device=buf.device)
    dist.all_gather_into_tensor(recv, send)
    parts = [recv[r * mx: r * mx + c] for r, c in enumerate(counts) if c > 0]
    out = torch.cat(parts).contiguous() if len(parts) > 1 else parts[0].contiguous()
    return out, int(sum(counts))


def expand_distributed(backend, cell_size: int, max_levels: int, rank: int, world: int,
                       rank_of_view) -> dict:
    """Expand::ExpandPatches (reference expand.cpp:34-101) across `world` ranks.  The
    organizer must hold the same seeds on every rank (dp_organizer_insert on each)."""
    stats = dict(levels=0, pops=0, passed=0, inserted=0, local_records=0)
    level = 0
    while max_levels < 0 or level < max_levels:
        fb, fe = backend.frontier()
        nf = fe - fb
        if nf <= 0:
            break
        buf, n_local = backend.local(cell_size, rank, world, rank_of_view, 4 * nf)
        records, total = gather_records(buf, n_local, world)
        inserted = backend.commit(records, total)
        stats["levels"] += 1
        stats["pops"] += nf
        stats["passed"] += total
        stats["inserted"] += inserted
        stats["local_records"] += n_local
        level += 1
    return stats

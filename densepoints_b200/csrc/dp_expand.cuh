// dp_expand.cuh -- device-resident PatchOrganizer and the expansion loop (K5, K6).
//
// Reference: PatchOrganizer / PatchGrid (methods/pmvs/patch_organizer.cpp:15-75) and
// Expand (methods/pmvs/expand.cpp:34-143).  All of this is integer work and must be
// bit-exact against the reference's single-thread FIFO order:
//
//  * A BFS level is processed at once.  ExpandPatch never reads the grids, so the children
//    of a level can be refined in any order; only TryInsert is order dependent.
//  * Every candidate carries its canonical sequence id
//        seq = (index of the parent in the level's frontier) * 4 + direction,
//    the order in which the 1-thread FIFO would call TryInsert.
//  * With max_patches_per_cell == 1 a free cell goes to the lowest sequence id that asks
//    for it, whether or not that candidate is finally kept (cells stay consumed,
//    patch_organizer.cpp:47-57) => atomicMin(seq) per cell reproduces the sequential
//    result; a candidate is kept iff it won > 1 cells; new patches are appended in
//    ascending seq.
//  * Multi-GPU: candidates are refined by the rank that owns their reference image; the
//    survivors travel as fixed-size records through one allgather; every rank then replays
//    this same commit => identical grids and stores everywhere.
#pragma once
#include "dp_context.h"

// ---- candidate record (what the allgather moves) and the store's visible sets ------------------
// A visible set is always in ascending view order (Patch::InitRelatedImages pushes view ids in
// ascending order, patch.cpp:29-47, and FilterByErrorMeasurement's erase keeps the order), so it
// is stored as a bit mask over the views: MW = ceil(n_views / 32) words.  A record is
//   [0] seq   [1] ref | nvis << 16   [2..4] pos (f32 bits)   [5..7] nrm   [8 .. 8+MW) mask
// = 40 bytes at 64 views, 64 bytes at 256 views (round 1: 9 + n_views words = 1 060 bytes).
#define DP_REC_HDR 8
__host__ __device__ static inline int dp_mask_words(int n_views) { return (n_views + 31) >> 5; }
__host__ __device__ static inline size_t rec_words(int n_views) {
  return (size_t)(DP_REC_HDR + dp_mask_words(n_views));
}

// (row, col) = ((size_t)(v / grid_scale), (size_t)(u / grid_scale)) with bounds test
// (patch_organizer.cpp:47-54, 15-30).  static_cast<size_t> truncates toward zero, so a quotient
// in (-1, 0) is cell 0 (defined behaviour; reachable for a refined seed that kept a view from
// the pre-refinement InitRelatedImages); q <= -1, NaN and huge q (UB in the reference; 2^63
// from x86-64's cvttsd2si) are out of bounds.
__device__ __forceinline__ long long dp_cell_of(const DpViewDev *__restrict__ V, double p0,
                                                double p1, double p2, double grid_scale) {
  double u, v;
  dp_project(V->P, p0, p1, p2, u, v);
  const double qr = v / grid_scale, qc = u / grid_scale;
  if (!(qr > -1.0) || !(qc > -1.0) || !(qr < 2147483647.0) || !(qc < 2147483647.0)) return -1;
  const long long row = (long long)qr, col = (long long)qc;  // truncation toward zero
  if (col >= V->gw || row >= V->gh) return -1;
  return V->grid_off + row * V->gw + col;
}

// K6a / K6b / K6c share one shape: a warp per record, a lane per view (stride 32), the lane's
// view is in the record's visible set iff its mask bit is set.
//   MODE 0 (claim)   every (record, visible view) asks for its cell with atomicMin(seq)
//   MODE 1 (resolve) count the cells the record won; accepted iff > 1 (patch_organizer.cpp:58)
//   MODE 2 (occupy)  winners occupy their cells (kept or not); claims are reset
template <int MODE>
__global__ void __launch_bounds__(256)
dp_cells_kernel(const DpViewDev *__restrict__ views, int n_views, const uint32_t *__restrict__ rec,
                long long n_rec, double grid_scale, uint8_t *__restrict__ grid,
                unsigned int *__restrict__ claim, unsigned int *__restrict__ flags,
                uint8_t *__restrict__ accepted) {
  const int lane = threadIdx.x & 31;
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rec) return;
  const uint32_t *R = rec + (size_t)r * rec_words(n_views);
  const uint32_t seq = R[0];
  const double p0 = (double)__uint_as_float(R[2]), p1 = (double)__uint_as_float(R[3]),
               p2 = (double)__uint_as_float(R[4]);
  unsigned wins = 0;
  for (int v = lane; v < n_views; v += 32) {
    if (!((R[DP_REC_HDR + (v >> 5)] >> lane) & 1u)) continue;
    const long long cell = dp_cell_of(views + v, p0, p1, p2, grid_scale);
    if (cell < 0) continue;
    if (MODE == 0) {
      if (grid[cell] == 0) atomicMin(claim + cell, seq);  // free before this level
    } else if (MODE == 1) {
      if (claim[cell] == seq) ++wins;
    } else {
      if (claim[cell] == seq) {
        grid[cell] = 1;
        claim[cell] = 0xffffffffu;
      }
    }
  }
  if (MODE == 1) {
    wins = __reduce_add_sync(DP_FULL, wins);
    if (lane == 0) {
      const bool acc = wins > 1;
      if (acc) flags[seq] = 1u;
      if (accepted) accepted[r] = acc ? 1 : 0;
    }
  }
}

// max_patches_per_cell = m > 1 (patch_organizer.cpp:21): a cell takes the first m patches that
// ask for it, in sequence order, over all levels.  The same claim / take pair runs m times: in
// round j the lowest sequence id still asking for a cell that is not full takes a slot of it and
// stops asking (its bit in `won`, a view mask per record); after m rounds nobody can win any more.
// That is exactly the sequential outcome: the requesters of a cell are served in ascending
// sequence id until it holds m patches.
//   MODE 0 (claim)  (record, visible view) pairs that have not won ask with atomicMin(seq)
//   MODE 1 (count)  accepted iff the record won > 1 cells (patch_organizer.cpp:58)
//   MODE 2 (take)   the winner of a cell occupies one more slot; the claim is reset
template <int MODE>
__global__ void __launch_bounds__(256)
dp_cells_multi_kernel(const DpViewDev *__restrict__ views, int n_views,
                      const uint32_t *__restrict__ rec, long long n_rec, double grid_scale, int m,
                      uint8_t *__restrict__ grid, unsigned int *__restrict__ claim,
                      unsigned int *__restrict__ won, unsigned int *__restrict__ flags,
                      uint8_t *__restrict__ accepted) {
  const int lane = threadIdx.x & 31;
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rec) return;
  const int mw = dp_mask_words(n_views);
  const uint32_t *R = rec + (size_t)r * rec_words(n_views);
  unsigned int *Wn = won + (size_t)r * mw;
  const uint32_t seq = R[0];
  if (MODE == 1) {
    unsigned wins = 0;
    for (int w = lane; w < mw; w += 32) wins += __popc(Wn[w]);
    wins = __reduce_add_sync(DP_FULL, wins);
    if (lane == 0) {
      const bool acc = wins > 1;
      if (acc) flags[seq] = 1u;
      if (accepted) accepted[r] = acc ? 1 : 0;
    }
    return;
  }
  const double p0 = (double)__uint_as_float(R[2]), p1 = (double)__uint_as_float(R[3]),
               p2 = (double)__uint_as_float(R[4]);
  for (int v = lane; v < n_views; v += 32) {
    if (!((R[DP_REC_HDR + (v >> 5)] >> lane) & 1u)) continue;
    if ((Wn[v >> 5] >> lane) & 1u) continue;  // has its slot already
    const long long cell = dp_cell_of(views + v, p0, p1, p2, grid_scale);
    if (cell < 0) continue;
    if (MODE == 0) {
      if ((int)grid[cell] < m) atomicMin(claim + cell, seq);
    } else {
      if (claim[cell] == seq) {  // one winner per cell and round
        grid[cell] = (uint8_t)(grid[cell] + 1);
        claim[cell] = 0xffffffffu;
        atomicOr(Wn + (v >> 5), 1u << lane);
      }
    }
  }
}

// K6d: append accepted records to the store at n0 + rank(seq).
__global__ void __launch_bounds__(256)
dp_append_kernel(const uint32_t *__restrict__ rec, long long n_rec, int n_views,
                 const unsigned int *__restrict__ flags, const unsigned int *__restrict__ offs,
                 long long n0, float *__restrict__ pos, float *__restrict__ nrm,
                 int32_t *__restrict__ ref, int32_t *__restrict__ nvis,
                 uint32_t *__restrict__ vmask) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int rw = (int)rec_words(n_views), mw = dp_mask_words(n_views);
  const long long r = t / rw;
  const int w = (int)(t - r * rw);
  if (r >= n_rec) return;
  const uint32_t *R = rec + (size_t)r * rw;
  const uint32_t seq = R[0];
  if (flags[seq] == 0) return;
  const long long d = n0 + offs[seq];
  const uint32_t x = R[w];
  if (w == 0) return;
  if (w == 1) {
    ref[d] = (int32_t)(x & 0xffffu);
    nvis[d] = (int32_t)(x >> 16);
  } else if (w < 5) {
    pos[3 * d + (w - 2)] = __uint_as_float(x);
  } else if (w < 8) {
    nrm[3 * d + (w - 5)] = __uint_as_float(x);
  } else {
    vmask[(size_t)d * mw + (w - DP_REC_HDR)] = x;
  }
}

// SoA batch (visible sets as ascending id lists) -> records, seq = index (seeds) or seq[]
// (expansion), only where keep != 0; slot[i] = output position (exclusive scan of keep) or
// identity when slot == null.  One warp per patch.
__global__ void __launch_bounds__(256)
dp_pack_records_kernel(int n, int vstride, int n_views, const float *__restrict__ pos,
                       const float *__restrict__ nrm, const int32_t *__restrict__ ref,
                       const int32_t *__restrict__ nvis, const int32_t *__restrict__ vis,
                       const unsigned int *__restrict__ seq, const uint8_t *__restrict__ keep,
                       const unsigned int *__restrict__ slot, uint32_t *__restrict__ rec) {
  const int lane = threadIdx.x & 31;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  if (keep && keep[i] == 0) return;
  uint32_t *R = rec + (size_t)(slot ? slot[i] : (unsigned)i) * rec_words(n_views);
  const int nv = min(max(nvis[i], 0), vstride);
  const int mw = dp_mask_words(n_views);
  if (lane == 0) {
    R[0] = seq ? seq[i] : (uint32_t)i;
    R[1] = ((uint32_t)ref[i] & 0xffffu) | ((uint32_t)nv << 16);
  }
  if (lane < 3) {
    R[2 + lane] = __float_as_uint(pos[3 * i + lane]);
    R[5 + lane] = __float_as_uint(nrm[3 * i + lane]);
  }
  // mask word w = OR of the bits of the listed views in [32 w, 32 w + 32)
  const int32_t *vi = vis + (size_t)i * vstride;
  for (int w = 0; w < mw; ++w) {
    unsigned m = 0;
    for (int k = lane; k < nv; k += 32) {
      const int v = vi[k];
      if (v >= 0 && v < n_views && (v >> 5) == w) m |= 1u << (v & 31);
    }
    m = __reduce_or_sync(DP_FULL, m);
    if (lane == 0) R[DP_REC_HDR + w] = m;
  }
}

// Store masks -> ascending id lists (dp_organizer_export and the children's copy of the
// parent's visible set).  One warp per patch.
__device__ __forceinline__ void dp_mask_to_list(const uint32_t *__restrict__ m, int mw, int lane,
                                                int32_t *__restrict__ out, int vstride) {
  int base = 0;
  for (int w = 0; w < mw; ++w) {
    const uint32_t bits = m[w];
    if ((bits >> lane) & 1u) {
      const int k = base + __popc(bits & ((1u << lane) - 1u));
      if (k < vstride) out[k] = 32 * w + lane;
    }
    base += __popc(bits);
  }
  for (int k = base + lane; k < vstride; k += 32) out[k] = -1;
}

__global__ void __launch_bounds__(256)
dp_unpack_masks_kernel(const uint32_t *__restrict__ vmask, long long n, int n_views,
                       int32_t *__restrict__ vis, int vstride) {
  const int lane = threadIdx.x & 31;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int mw = dp_mask_words(n_views);
  dp_mask_to_list(vmask + (size_t)i * mw, mw, lane, vis + (size_t)i * vstride, vstride);
}

__global__ void dp_u8_to_u32_kernel(const uint8_t *__restrict__ in, unsigned int *__restrict__ out,
                                    long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] ? 1u : 0u;
}

// K5a: which frontier parents expand: >= 2 visible views (expand.cpp:69) and owned by `rank`.
// Ownership: rank_of_view[reference image] when a table is given; else (world > 1) the frontier
// is cut into `world` contiguous ranges of equal work -- wscan = exclusive scan of the parents'
// weights (their visible-view counts, 0 if they do not expand), wtotal its total.  The frontier
// is cut into world * DP_RANGE_INTERLEAVE pieces of equal work dealt out round-robin (parents
// that are neighbours in the store are neighbours on the surface and tend to be equally hard, so
// one long range per rank left the ranks up to 60 % apart; eight interleaved pieces average that
// out): parent i belongs to rank floor(wscan[i] * world * 8 / wtotal) mod world.  Every rank
// computes the same cut from the replicated store.
#define DP_RANGE_INTERLEAVE 8
__global__ void dp_parent_flags_kernel(const int32_t *__restrict__ nvis,
                                       const int32_t *__restrict__ ref, long long begin,
                                       long long n_f, const int32_t *__restrict__ rank_of_view,
                                       const unsigned int *__restrict__ wscan,
                                       unsigned long long wtotal, int world, int rank, int n_views,
                                       unsigned int *__restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_f) return;
  const long long p = begin + i;
  bool f = nvis[p] >= 2;
  if (f && rank_of_view) {
    const int r = ref[p];
    f = (r >= 0 && r < n_views) && rank_of_view[r] == rank;
  } else if (f && wscan) {
    f = (int)((((unsigned long long)wscan[i] * (unsigned long long)(world * DP_RANGE_INTERLEAVE)) / wtotal) %
              (unsigned long long)world) == rank;
  }
  flags[i] = f ? 1u : 0u;
}
__global__ void dp_parent_weights_kernel(const int32_t *__restrict__ nvis, long long begin,
                                         long long n_f, unsigned int *__restrict__ w) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_f) return;
  const int nv = nvis[begin + i];
  w[i] = nv >= 2 ? (unsigned int)nv : 0u;
}

// Work of the frontier per reference view: sum of the visible-view counts of the parents that
// expand (the cost of a candidate is evaluations x views).  Feeds the per-level ownership table
// of the multi-GPU driver.
__global__ void dp_frontier_weights_kernel(const int32_t *__restrict__ nvis,
                                           const int32_t *__restrict__ ref, long long begin,
                                           long long n_f, int n_views,
                                           unsigned long long *__restrict__ weight) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_f) return;
  const long long p = begin + i;
  const int nv = nvis[p], r = ref[p];
  if (nv >= 2 && r >= 0 && r < n_views) atomicAdd(weight + r, (unsigned long long)nv);
}

// K5b: Expand::ExpandPatch's proposals (expand.cpp:106-127): 4 children per expanding
// parent at +-(grid_scale / dx) along the patch x / y axes; children copy the parent.  Only the
// parents whose compacted index lies in [slot0, slot1) are proposed (one chunk of a level).
__global__ void __launch_bounds__(128)
dp_propose_kernel(const DpViewDev *__restrict__ views, int n_views, long long begin, long long n_f,
                  const unsigned int *__restrict__ pflags, const unsigned int *__restrict__ pslot,
                  unsigned int slot0, unsigned int slot1,
                  const float *__restrict__ spos, const float *__restrict__ snrm,
                  const int32_t *__restrict__ sref, const int32_t *__restrict__ snvis,
                  const uint32_t *__restrict__ smask, int vstride, double grid_scale,
                  float *__restrict__ cpos, float *__restrict__ cnrm, int32_t *__restrict__ cref,
                  int32_t *__restrict__ cnvis, int32_t *__restrict__ cvis,
                  unsigned int *__restrict__ cseq) {
  const int lane = threadIdx.x & 31;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // warp / parent
  if (i >= n_f || pflags[i] == 0) return;
  const unsigned int ps = pslot[i];
  if (ps < slot0 || ps >= slot1) return;
  const long long p = begin + i;
  const int r = sref[p];
  if (r < 0 || r >= n_views) return;
  const DpViewDev *V = views + r;
  const double p0 = spos[3 * p], p1 = spos[3 * p + 1], p2 = spos[3 * p + 2];
  const double n0 = snrm[3 * p], n1 = snrm[3 * p + 1], n2 = snrm[3 * p + 2];
  const double xa0 = V->xa[0], xa1 = V->xa[1], xa2 = V->xa[2];
  const double ya0 = xsub(xmul(n1, xa2), xmul(n2, xa1));
  const double ya1 = xsub(xmul(n2, xa0), xmul(n0, xa2));
  const double ya2 = xsub(xmul(n0, xa1), xmul(n1, xa0));
  double cu, cv, qu, qv;
  dp_project(V->P, p0, p1, p2, cu, cv);
  dp_project(V->P, xadd(p0, xa0), xadd(p1, xa1), xadd(p2, xa2), qu, qv);
  const double du = xsub(qu, cu), dv = xsub(qv, cv);
  const double dx = sqrt(xadd(xmul(du, du), xmul(dv, dv)));
  const double scale = grid_scale / dx;  // expand.cpp:112
  const long long c0 = (long long)(ps - slot0) * 4;
  const int nv = min(max(snvis[p], 0), vstride);
  if (lane < 4) {
    const int d = lane;  // directions x, -x, y, -y (expand.cpp:114-116)
    const double sg = (d & 1) ? -1.0 : 1.0;
    const double d0 = sg * (d < 2 ? xa0 : ya0), d1 = sg * (d < 2 ? xa1 : ya1),
                 d2 = sg * (d < 2 ? xa2 : ya2);
    const long long c = c0 + d;
    cpos[3 * c + 0] = (float)xadd(p0, xmul(scale, d0));  // SetPosition: fp32
    cpos[3 * c + 1] = (float)xadd(p1, xmul(scale, d1));
    cpos[3 * c + 2] = (float)xadd(p2, xmul(scale, d2));
    cnrm[3 * c + 0] = snrm[3 * p];
    cnrm[3 * c + 1] = snrm[3 * p + 1];
    cnrm[3 * c + 2] = snrm[3 * p + 2];
    cref[c] = r;
    cnvis[c] = nv;
    cseq[c] = (unsigned int)(i * 4 + d);
  }
  const int mw = dp_mask_words(n_views);
  for (int d = 0; d < 4; ++d)  // Patch new_patch = *patch: the parent's visible set
    dp_mask_to_list(smask + (size_t)p * mw, mw, lane, cvis + (size_t)(c0 + d) * vstride, vstride);
}

// ---- exclusive scan of u32 (multi-level, 1024 items per block) ---------------------------
__global__ void __launch_bounds__(256)
dp_scan_block_kernel(const unsigned int *__restrict__ in, unsigned int *__restrict__ out,
                     unsigned int *__restrict__ block_sums, long long n) {
  __shared__ unsigned int warp_sums[8];
  const long long base = (long long)blockIdx.x * 1024 + threadIdx.x * 4;
  unsigned int v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = (base + j < n) ? in[base + j] : 0u;
  unsigned int tsum = v[0] + v[1] + v[2] + v[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned int t = __shfl_up_sync(DP_FULL, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  unsigned int woff = 0;
  for (int w = 0; w < warp; ++w) woff += warp_sums[w];
  unsigned int excl = woff + inc - tsum;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (base + j < n) out[base + j] = excl;
    excl += v[j];
  }
  if (threadIdx.x == 255 && block_sums) block_sums[blockIdx.x] = woff + inc;
}

__global__ void dp_scan_add_kernel(unsigned int *__restrict__ out,
                                   const unsigned int *__restrict__ block_offs, long long n) {
  const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long k = i + (long long)j * 256;
    if (k < n) out[k] += block_offs[blockIdx.x];
  }
}

// out[i] = sum in[0..i) for i < n (n > 0).  Scratch (per-level block sums) grows on demand.
static int dp_exclusive_scan(dp_context *ctx, const unsigned int *in, unsigned int *out, long long n,
                             cudaStream_t st) {
  if (n <= 0) return DP_OK;
  std::vector<long long> cnt;  // items per level
  cnt.push_back(n);
  while (cnt.back() > 1024) cnt.push_back((cnt.back() + 1023) / 1024);
  size_t scratch = 0;
  for (size_t l = 1; l < cnt.size(); ++l) scratch += 2 * (size_t)cnt[l] * sizeof(unsigned int);
  DP_CUDA(ctx, ctx->e_scan.ensure(scratch + 16));
  std::vector<unsigned int *> sums(cnt.size() + 1, nullptr), offs(cnt.size() + 1, nullptr);
  unsigned int *sp = ctx->e_scan.as<unsigned int>();
  for (size_t l = 1; l < cnt.size(); ++l) {
    sums[l] = sp;
    sp += cnt[l];
    offs[l] = sp;
    sp += cnt[l];
  }
  for (size_t l = 0; l < cnt.size(); ++l) {  // up-sweep: local scans + block sums
    const unsigned int *src = (l == 0) ? in : sums[l];
    unsigned int *dst = (l == 0) ? out : offs[l];
    const unsigned blocks = (unsigned)((cnt[l] + 1023) / 1024);
    dp_scan_block_kernel<<<blocks, 256, 0, st>>>(src, dst, l + 1 < cnt.size() ? sums[l + 1] : nullptr,
                                                 cnt[l]);
    ++ctx->launches;
  }
  for (int l = (int)cnt.size() - 2; l >= 0; --l) {  // down-sweep: add the scanned block sums
    unsigned int *dst = (l == 0) ? out : offs[l];
    const unsigned blocks = (unsigned)((cnt[l] + 1023) / 1024);
    dp_scan_add_kernel<<<blocks, 256, 0, st>>>(dst, offs[l + 1], cnt[l]);
    ++ctx->launches;
  }
  DP_CUDA(ctx, cudaGetLastError());
  return DP_OK;
}

// total = out[n-1] + in[n-1], read back on the host
static int dp_scan_total(dp_context *ctx, const unsigned int *in, const unsigned int *out, long long n,
                         cudaStream_t st, long long *total) {
  if (n == 0) {
    *total = 0;
    return DP_OK;
  }
  unsigned int a = 0, b = 0;
  DP_CUDA(ctx, cudaMemcpyAsync(&a, out + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaMemcpyAsync(&b, in + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  *total = (long long)a + b;
  return DP_OK;
}

// ---- organizer --------------------------------------------------------------------------

static int org_reserve(dp_context *ctx, long long need, cudaStream_t st) {
  DpOrganizer &o = ctx->org;
  if (need <= o.cap) return DP_OK;
  long long cap = std::max<long long>(need + need / 2, 1 << 16);
  const size_t mw = (size_t)dp_mask_words(o.vstride);
  struct Item {
    DpDevBuf *b;
    size_t elem;
  } items[] = {{&o.pos, 12}, {&o.nrm, 12}, {&o.rgb, 3}, {&o.ref, 4}, {&o.nvis, 4}, {&o.vis, 4 * mw}};
  for (auto &it : items) {
    void *np = nullptr;
    DP_CUDA(ctx, cudaMalloc(&np, it.elem * (size_t)cap));
    cudaError_t e = cudaSuccess;
    if (o.n > 0 && it.b->ptr)
      e = cudaMemcpyAsync(np, it.b->ptr, it.elem * (size_t)o.n, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      cudaFree(np);
      return dp_fail(ctx, DP_ERR_CUDA, "org_reserve", e);
    }
    if (it.b->ptr) cudaFree(it.b->ptr);
    it.b->ptr = np;
    it.b->cap = it.elem * (size_t)cap;
  }
  o.cap = cap;
  return DP_OK;
}

// PatchOrganizer::AllocateViews (patch_organizer.cpp:32-40)
extern "C" int dp_organizer_reset(dp_context *ctx) {
  if (!ctx) return DP_ERR_INVALID_ARG;
  DpDeviceGuard guard__(ctx->device);
  int rc = dp_sync_views(ctx);
  if (rc != DP_OK) return rc;
  DpOrganizer &o = ctx->org;
  cudaStream_t st = ctx->stream;
  DP_CUDA(ctx, cudaDeviceSynchronize());  // level calls may have run on caller streams
  ctx->scratch_busy = false;
  DP_CUDA(ctx, o.grid.ensure((size_t)o.n_cells + 1));
  DP_CUDA(ctx, o.claim.ensure(((size_t)o.n_cells + 1) * 4));
  DP_CUDA(ctx, cudaMemsetAsync(o.grid.ptr, 0, (size_t)o.n_cells + 1, st));
  DP_CUDA(ctx, cudaMemsetAsync(o.claim.ptr, 0xff, ((size_t)o.n_cells + 1) * 4, st));
  o.n = 0;
  o.vstride = (int)ctx->views.size();
  o.frontier_begin = 0;
  o.pops = 0;
  o.level = ctx->level;
  o.ready = true;
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  return DP_OK;
}

// TryInsert over `n_rec` records (any order) whose seq ids lie in [0, seq_space).
static int org_commit(dp_context *ctx, const uint32_t *rec, long long n_rec, long long seq_space,
                      uint8_t *accepted_dev, long long *n_inserted, cudaStream_t st) {
  DpOrganizer &o = ctx->org;
  *n_inserted = 0;
  if (n_rec == 0 || seq_space == 0) return DP_OK;
  const int nviews = (int)ctx->views.size();
  const DpViewDev *views = ctx->d_views.as<DpViewDev>();
  const double gs = (double)ctx->prm.grid_scale;
  DP_CUDA(ctx, ctx->e_flags.ensure(((size_t)seq_space + 1) * 8));
  unsigned int *flags = ctx->e_flags.as<unsigned int>();
  unsigned int *offs = flags + (seq_space + 1);
  DP_CUDA(ctx, cudaMemsetAsync(flags, 0, ((size_t)seq_space + 1) * 4, st));
  const long long tw = n_rec * 32;
  const unsigned gw = (unsigned)((tw + 255) / 256);
  uint8_t *grid = o.grid.as<uint8_t>();
  unsigned int *claim = o.claim.as<unsigned int>();
  const int m = ctx->prm.max_patches_per_cell;
  if (m == 1) {
    dp_cells_kernel<0><<<gw, 256, 0, st>>>(views, nviews, rec, n_rec, gs, grid, claim, flags, nullptr);
    dp_cells_kernel<1><<<gw, 256, 0, st>>>(views, nviews, rec, n_rec, gs, grid, claim, flags,
                                           accepted_dev);
    dp_cells_kernel<2><<<gw, 256, 0, st>>>(views, nviews, rec, n_rec, gs, grid, claim, flags, nullptr);
    ctx->launches += 3;
  } else {
    const size_t won_bytes = (size_t)n_rec * dp_mask_words(nviews) * 4;
    DP_CUDA(ctx, ctx->e_won.ensure(won_bytes));
    unsigned int *won = ctx->e_won.as<unsigned int>();
    DP_CUDA(ctx, cudaMemsetAsync(won, 0, won_bytes, st));
    for (int round = 0; round < m; ++round) {
      dp_cells_multi_kernel<0><<<gw, 256, 0, st>>>(views, nviews, rec, n_rec, gs, m, grid, claim, won,
                                                   flags, nullptr);
      dp_cells_multi_kernel<2><<<gw, 256, 0, st>>>(views, nviews, rec, n_rec, gs, m, grid, claim, won,
                                                   flags, nullptr);
    }
    dp_cells_multi_kernel<1><<<gw, 256, 0, st>>>(views, nviews, rec, n_rec, gs, m, grid, claim, won,
                                                 flags, accepted_dev);
    ctx->launches += 2 * m + 1;
  }
  DP_CUDA(ctx, cudaGetLastError());
  int rc = dp_exclusive_scan(ctx, flags, offs, seq_space, st);
  if (rc != DP_OK) return rc;
  long long total = 0;
  rc = dp_scan_total(ctx, flags, offs, seq_space, st, &total);
  if (rc != DP_OK) return rc;
  if (total > 0) {
    rc = org_reserve(ctx, o.n + total, st);
    if (rc != DP_OK) return rc;
    const long long ta = n_rec * (long long)rec_words(nviews);
    dp_append_kernel<<<(unsigned)((ta + 255) / 256), 256, 0, st>>>(
        rec, n_rec, nviews, flags, offs, o.n, o.pos.as<float>(), o.nrm.as<float>(),
        o.ref.as<int32_t>(), o.nvis.as<int32_t>(), o.vis.as<uint32_t>());
    // Patch::ComputeColor for the new patches (patch_organizer.cpp:60)
    const long long t3 = total * 32;
    dp_color_kernel<<<(unsigned)((t3 + 255) / 256), 256, 0, st>>>(
        views, nviews, (int)total, o.pos.as<float>() + 3 * o.n, o.rgb.as<uint8_t>() + 3 * o.n);
    ctx->launches += 2;
    DP_CUDA(ctx, cudaGetLastError());
    o.n += total;
  }
  *n_inserted = total;
  return DP_OK;
}

static int org_check(dp_context *ctx) {
  if (!ctx) return DP_ERR_INVALID_ARG;
  if (!ctx->org.ready) return dp_fail(ctx, DP_ERR_STATE, "call dp_organizer_reset first");
  if (ctx->org.level != ctx->level || ctx->views_dirty)
    return dp_fail(ctx, DP_ERR_STATE, "views / level changed since dp_organizer_reset");
  return DP_OK;
}

// PatchOrganizer::SetSeeds (patch_organizer.cpp:70-75)
extern "C" int dp_organizer_insert(dp_context *ctx, const dp_patch_soa *h, uint8_t *accepted) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  if (!h) return DP_ERR_INVALID_ARG;
  if (h->n == 0) return DP_OK;
  DpDeviceGuard guard__(ctx->device);
  DpOrganizer &o = ctx->org;
  // The store keeps a visible set as a bit mask: the ids must be valid and strictly ascending
  // (what Patch::InitRelatedImages and FilterByErrorMeasurement produce).
  for (int i = 0; i < h->n; ++i) {
    const int nv = h->nvis ? h->nvis[i] : 0;
    if (nv < 0 || nv > h->vstride || nv > o.vstride)
      return dp_fail(ctx, DP_ERR_INVALID_ARG, "nvis out of range");
    const int32_t *vi = h->vis + (size_t)i * h->vstride;
    for (int k = 0; k < nv; ++k)
      if (vi[k] < 0 || vi[k] >= o.vstride || (k > 0 && vi[k] <= vi[k - 1]))
        return dp_fail(ctx, DP_ERR_INVALID_ARG,
                       "visible ids must be valid and strictly ascending (Patch::InitRelatedImages order)");
    if (h->ref[i] < 0 || h->ref[i] >= o.vstride)
      return dp_fail(ctx, DP_ERR_INVALID_ARG, "reference image out of range");
  }
  dp_patch_dev d;
  rc = upload_patches(ctx, h, &d, true);
  if (rc != DP_OK) return rc;
  cudaStream_t st = ctx->stream;
  const size_t rw = rec_words(o.vstride);
  DP_CUDA(ctx, ctx->e_cells.ensure((size_t)h->n * rw * 4));
  DP_CUDA(ctx, ctx->e_keep.ensure((size_t)h->n));
  uint32_t *rec = ctx->e_cells.as<uint32_t>();
  const long long t = (long long)h->n * 32;
  dp_pack_records_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(
      h->n, h->vstride, o.vstride, d.pos, d.nrm, d.ref, d.nvis, d.vis, nullptr, nullptr, nullptr, rec);
  ++ctx->launches;
  DP_CUDA(ctx, cudaGetLastError());
  long long ins = 0;
  rc = org_commit(ctx, rec, h->n, h->n, ctx->e_keep.as<uint8_t>(), &ins, st);
  if (rc != DP_OK) return rc;
  if (accepted)
    DP_CUDA(ctx, cudaMemcpyAsync(accepted, ctx->e_keep.ptr, (size_t)h->n, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  return DP_OK;
}

extern "C" int64_t dp_organizer_size(const dp_context *ctx) { return ctx ? ctx->org.n : 0; }

extern "C" int dp_organizer_export(dp_context *ctx, dp_patch_soa *out) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  if (!out) return DP_ERR_INVALID_ARG;
  DpDeviceGuard guard__(ctx->device);
  DpOrganizer &o = ctx->org;
  if (out->n < o.n || out->vstride < 1) return dp_fail(ctx, DP_ERR_INVALID_ARG, "export capacity");
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)o.n;
  if (n > 0) {
    DP_CUDA(ctx, cudaMemcpyAsync(out->pos, o.pos.ptr, n * 12, cudaMemcpyDeviceToHost, st));
    DP_CUDA(ctx, cudaMemcpyAsync(out->nrm, o.nrm.ptr, n * 12, cudaMemcpyDeviceToHost, st));
    DP_CUDA(ctx, cudaMemcpyAsync(out->ref, o.ref.ptr, n * 4, cudaMemcpyDeviceToHost, st));
    DP_CUDA(ctx, cudaMemcpyAsync(out->nvis, o.nvis.ptr, n * 4, cudaMemcpyDeviceToHost, st));
    if (out->rgb) DP_CUDA(ctx, cudaMemcpyAsync(out->rgb, o.rgb.ptr, n * 3, cudaMemcpyDeviceToHost, st));
    // masks -> ascending id lists in the caller's row length, in slabs of bounded size
    const size_t vs = (size_t)out->vstride;
    const size_t slab = std::max<size_t>(1, std::min<size_t>(n, ((size_t)256 << 20) / (vs * 4)));
    DP_CUDA(ctx, ctx->e_vis.ensure(slab * vs * 4));
    for (size_t o0 = 0; o0 < n; o0 += slab) {
      const size_t m = std::min(slab, n - o0);
      const long long t = (long long)m * 32;
      dp_unpack_masks_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(
          o.vis.as<uint32_t>() + o0 * dp_mask_words(o.vstride), (long long)m, o.vstride,
          ctx->e_vis.as<int32_t>(), out->vstride);
      ++ctx->launches;
      DP_CUDA(ctx, cudaGetLastError());
      DP_CUDA(ctx, cudaMemcpyAsync(out->vis + o0 * vs, ctx->e_vis.ptr, m * vs * 4,
                                   cudaMemcpyDeviceToHost, st));
      DP_CUDA(ctx, cudaStreamSynchronize(st));
    }
  }
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  out->n = (int32_t)o.n;
  return DP_OK;
}

extern "C" int dp_organizer_grid(dp_context *ctx, int view_id, uint8_t *out, size_t capacity,
                                 int *gw, int *gh) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  if (view_id < 0 || view_id >= (int)ctx->views.size()) return dp_fail(ctx, DP_ERR_INVALID_ARG, "view_id");
  DpDeviceGuard guard__(ctx->device);
  const DpLevel &l = ctx->views[view_id].levels[ctx->level];
  const int w = l.width / ctx->prm.grid_scale, h = l.height / ctx->prm.grid_scale;
  if (gw) *gw = w;
  if (gh) *gh = h;
  if (!out) return DP_OK;
  if (capacity < (size_t)w * h) return dp_fail(ctx, DP_ERR_INVALID_ARG, "grid capacity");
  long long off = 0;
  for (int v = 0; v < view_id; ++v) {
    const DpLevel &lv = ctx->views[v].levels[ctx->level];
    off += (long long)(lv.width / ctx->prm.grid_scale) * (lv.height / ctx->prm.grid_scale);
  }
  DP_CUDA(ctx, cudaMemcpyAsync(out, ctx->org.grid.as<uint8_t>() + off, (size_t)w * h,
                               cudaMemcpyDeviceToHost, ctx->stream));
  DP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return DP_OK;
}

// All occupancy grids, concatenated in view order (what dp_organizer_grid returns view by view).
extern "C" int dp_organizer_grids(dp_context *ctx, uint8_t *out, size_t capacity, int64_t *n_cells) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  DpDeviceGuard guard__(ctx->device);
  if (n_cells) *n_cells = ctx->org.n_cells;
  if (!out) return DP_OK;
  if (capacity < (size_t)ctx->org.n_cells) return dp_fail(ctx, DP_ERR_INVALID_ARG, "grid capacity");
  DP_CUDA(ctx, cudaMemcpyAsync(out, ctx->org.grid.ptr, (size_t)ctx->org.n_cells, cudaMemcpyDeviceToHost,
                               ctx->stream));
  DP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return DP_OK;
}

// ---- expansion ----------------------------------------------------------------------------

extern "C" size_t dp_record_bytes(const dp_context *ctx) {
  return ctx ? rec_words((int)ctx->views.size()) * 4 : 0;
}

extern "C" int dp_expand_frontier(dp_context *ctx, int64_t *begin, int64_t *end) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  long long nf = ctx->org.n - ctx->org.frontier_begin;
  const long long room = ctx->prm.max_pops - ctx->org.pops;  // expand.cpp:95
  if (nf > room) nf = room > 0 ? room : 0;
  if (begin) *begin = ctx->org.frontier_begin;
  if (end) *end = ctx->org.frontier_begin + nf;
  return DP_OK;
}

extern "C" int64_t dp_expand_last_candidates(const dp_context *ctx) {
  return ctx ? ctx->org_last_candidates : 0;
}

// Per reference view: sum of the visible-view counts of the frontier parents that expand -- the
// work a rank takes on by owning that view in this level (weights: n_views entries, host).
extern "C" int dp_expand_frontier_weights(dp_context *ctx, int64_t *weights) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  if (!weights) return DP_ERR_INVALID_ARG;
  DpDeviceGuard guard__(ctx->device);
  const int nviews = (int)ctx->views.size();
  for (int v = 0; v < nviews; ++v) weights[v] = 0;
  int64_t fb = 0, fe = 0;
  dp_expand_frontier(ctx, &fb, &fe);
  if (fe <= fb) return DP_OK;
  cudaStream_t st = ctx->stream;
  DP_CUDA(ctx, ctx->s_misc.ensure((size_t)nviews * 8));
  DP_CUDA(ctx, cudaMemsetAsync(ctx->s_misc.ptr, 0, (size_t)nviews * 8, st));
  const long long nf = fe - fb;
  dp_frontier_weights_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(
      ctx->org.nvis.as<int32_t>(), ctx->org.ref.as<int32_t>(), fb, nf, nviews,
      ctx->s_misc.as<unsigned long long>());
  ++ctx->launches;
  DP_CUDA(ctx, cudaGetLastError());
  DP_CUDA(ctx, cudaMemcpyAsync(weights, ctx->s_misc.ptr, (size_t)nviews * 8, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  return DP_OK;
}

// Step 1 of a level: propose + refine + visibility + filter for the parents this rank owns;
// survivors are packed as records (ascending seq) straight into records_dev.  The candidates are
// processed in chunks so that their visible-id tables (4 bytes x n_views per candidate) stay
// below DP_EXPAND_CHUNK_BYTES whatever the frontier size (256 views: 1 KB per candidate).
#ifndef DP_EXPAND_CHUNK_BYTES
#define DP_EXPAND_CHUNK_BYTES ((size_t)1 << 30)
#endif
extern "C" int dp_expand_level_local(dp_context *ctx, int cell_size, int rank, int world,
                                     const int32_t *rank_of_view, void *records_dev,
                                     int64_t max_records, int64_t *n_records, void *stream) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  if (!n_records) return DP_ERR_INVALID_ARG;
  *n_records = 0;
  DpDeviceGuard guard__(ctx->device);
  DpOrganizer &o = ctx->org;
  cudaStream_t st = (cudaStream_t)stream;  // as given: 0 is the legacy default stream (torch's)
  int64_t fb = 0, fe = 0;
  dp_expand_frontier(ctx, &fb, &fe);
  const long long nf = fe - fb;
  ctx->org_last_candidates = 0;
  if (nf <= 0) return DP_OK;
  if ((rc = dp_scratch_acquire(ctx, st)) != DP_OK) return rc;
  const int vs = o.vstride, nviews = (int)ctx->views.size();
  const DpViewDev *views = ctx->d_views.as<DpViewDev>();
  // ownership table on the device
  const int32_t *d_rov = nullptr;
  if (world > 1 && rank_of_view) {
    DP_CUDA(ctx, ctx->s_misc.ensure((size_t)nviews * 8));
    DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_misc.ptr, rank_of_view, (size_t)nviews * 4,
                                 cudaMemcpyHostToDevice, st));
    d_rov = ctx->s_misc.as<int32_t>();
  }
  // K5a + scan: compact the expanding parents
  DP_CUDA(ctx, ctx->e_count.ensure(((size_t)nf + 1) * 16));
  unsigned int *pflags = ctx->e_count.as<unsigned int>();
  unsigned int *pslot = pflags + (nf + 1);
  const unsigned int *wscan = nullptr;
  long long wtotal = 0;
  if (world > 1 && !rank_of_view) {  // equal-work contiguous ranges of the frontier
    unsigned int *w = pslot + (nf + 1), *ws = w + (nf + 1);
    dp_parent_weights_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(o.nvis.as<int32_t>(), fb, nf, w);
    ++ctx->launches;
    if ((rc = dp_exclusive_scan(ctx, w, ws, nf, st)) != DP_OK) return rc;
    if ((rc = dp_scan_total(ctx, w, ws, nf, st, &wtotal)) != DP_OK) return rc;
    if (wtotal == 0) return dp_scratch_release(ctx, st);
    wscan = ws;
  }
  dp_parent_flags_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(
      o.nvis.as<int32_t>(), o.ref.as<int32_t>(), fb, nf, d_rov, wscan, (unsigned long long)wtotal,
      world, rank, nviews, pflags);
  ++ctx->launches;
  rc = dp_exclusive_scan(ctx, pflags, pslot, nf, st);
  if (rc != DP_OK) return rc;
  long long n_par = 0;
  rc = dp_scan_total(ctx, pflags, pslot, nf, st, &n_par);
  if (rc != DP_OK) return rc;
  if (n_par == 0) return dp_scratch_release(ctx, st);
  if (n_par * 4 > 0x7fffffffLL) return dp_fail(ctx, DP_ERR_INVALID_ARG, "level too large");
  const long long par_chunk =
      std::max<long long>(1024, (long long)(DP_EXPAND_CHUNK_BYTES / ((size_t)vs * 16)));
  const long long ncmax = std::min(n_par, par_chunk) * 4;
  DP_CUDA(ctx, ctx->e_pos.ensure((size_t)ncmax * 12));
  DP_CUDA(ctx, ctx->e_nrm.ensure((size_t)ncmax * 12));
  DP_CUDA(ctx, ctx->e_ref.ensure((size_t)ncmax * 4));
  DP_CUDA(ctx, ctx->e_nvis.ensure((size_t)ncmax * 4));
  DP_CUDA(ctx, ctx->e_vis.ensure((size_t)ncmax * vs * 4));
  DP_CUDA(ctx, ctx->e_seq.ensure((size_t)ncmax * 4));
  DP_CUDA(ctx, ctx->e_keep.ensure((size_t)ncmax));
  DP_CUDA(ctx, ctx->e_flags.ensure(((size_t)ncmax + 1) * 8));
  long long n_out = 0;
  for (long long s0 = 0; s0 < n_par; s0 += par_chunk) {
    const long long s1 = std::min(n_par, s0 + par_chunk);
    const long long nc = (s1 - s0) * 4;
    const long long tp = nf * 32;
    dp_propose_kernel<<<(unsigned)((tp + 127) / 128), 128, 0, st>>>(
        views, nviews, fb, nf, pflags, pslot, (unsigned)s0, (unsigned)s1, o.pos.as<float>(),
        o.nrm.as<float>(), o.ref.as<int32_t>(), o.nvis.as<int32_t>(), o.vis.as<uint32_t>(), vs,
        (double)ctx->prm.grid_scale, ctx->e_pos.as<float>(), ctx->e_nrm.as<float>(),
        ctx->e_ref.as<int32_t>(), ctx->e_nvis.as<int32_t>(), ctx->e_vis.as<int32_t>(),
        ctx->e_seq.as<unsigned int>());
    ++ctx->launches;
    DP_CUDA(ctx, cudaGetLastError());
    dp_patch_dev c;
    c.n = (int32_t)nc;
    c.vstride = vs;
    c.pos = ctx->e_pos.as<float>();
    c.nrm = ctx->e_nrm.as<float>();
    c.ref = ctx->e_ref.as<int32_t>();
    c.nvis = ctx->e_nvis.as<int32_t>();
    c.vis = ctx->e_vis.as<int32_t>();
    c.rgb = nullptr;
    // Optimize (expand.cpp:129-130) -> InitRelatedImages (:132) -> FilterByErrorMeasurement (:133)
    if ((rc = dp_refine_dev(ctx, &c, cell_size, nullptr, nullptr, nullptr, st)) != DP_OK) return rc;
    if ((rc = dp_visibility_dev(ctx, &c, nullptr, nullptr, st)) != DP_OK) return rc;
    if ((rc = dp_filter_dev(ctx, &c, cell_size, ctx->e_keep.as<uint8_t>(), st)) != DP_OK) return rc;
    // compact survivors into records, ascending seq (candidate order is already ascending)
    unsigned int *kflags = ctx->e_flags.as<unsigned int>();
    unsigned int *kslot = kflags + (nc + 1);
    dp_u8_to_u32_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(ctx->e_keep.as<uint8_t>(), kflags, nc);
    ++ctx->launches;
    rc = dp_exclusive_scan(ctx, kflags, kslot, nc, st);
    if (rc != DP_OK) return rc;
    long long n_keep = 0;
    rc = dp_scan_total(ctx, kflags, kslot, nc, st, &n_keep);
    if (rc != DP_OK) return rc;
    if (n_out + n_keep > max_records) return dp_fail(ctx, DP_ERR_INVALID_ARG, "records buffer too small");
    if (n_keep > 0) {
      const long long t = nc * 32;
      dp_pack_records_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(
          (int)nc, vs, nviews, c.pos, c.nrm, c.ref, c.nvis, c.vis, ctx->e_seq.as<unsigned int>(),
          ctx->e_keep.as<uint8_t>(), kslot,
          (uint32_t *)records_dev + (size_t)n_out * rec_words(nviews));
      ++ctx->launches;
      DP_CUDA(ctx, cudaGetLastError());
    }
    n_out += n_keep;
    ctx->org_last_candidates += nc;  // stats: candidates refined this level
  }
  *n_records = n_out;
  return dp_scratch_release(ctx, st);
}

// Step 3 of a level: TryInsert replay over the gathered records; advances the frontier.
extern "C" int dp_expand_level_commit(dp_context *ctx, const void *records_dev, int64_t n_records,
                                      int64_t *n_inserted, void *stream) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  DpDeviceGuard guard__(ctx->device);
  DpOrganizer &o = ctx->org;
  cudaStream_t st = (cudaStream_t)stream;  // as given: 0 is the legacy default stream (torch's)
  int64_t fb = 0, fe = 0;
  dp_expand_frontier(ctx, &fb, &fe);
  const long long nf = fe - fb;
  const long long n_before = o.n;
  long long ins = 0;
  if ((rc = dp_scratch_acquire(ctx, st)) != DP_OK) return rc;
  rc = org_commit(ctx, (const uint32_t *)records_dev, n_records, nf * 4, nullptr, &ins, st);
  if (rc != DP_OK) return rc;
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->scratch_busy = false;  // the stream is drained: the scratch is free
  o.pops += nf;
  o.frontier_begin = n_before;  // the next level = the patches appended by this one
  if (o.pops >= ctx->prm.max_pops) o.frontier_begin = o.n;  // expand.cpp:95-97: stop
  if (n_inserted) *n_inserted = ins;
  return DP_OK;
}

// The same over the output of an allgather of padded per-rank buffers: `world` segments of
// `segment_capacity` records each, of which the first counts[r] are valid (rank order = seq
// order is irrelevant: the commit is order independent).  The valid records are compacted on the
// device (no host-side concatenation) and committed.
__global__ void dp_compact_segments_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                           long long seg_words, long long valid_words,
                                           long long out_off_words) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < valid_words) out[out_off_words + t] = in[t];
  (void)seg_words;
}
extern "C" int dp_expand_level_commit_gathered(dp_context *ctx, const void *gathered_dev, int world,
                                               int64_t segment_capacity, const int64_t *counts,
                                               int64_t *n_inserted, void *stream) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  if (world < 1 || !counts) return dp_fail(ctx, DP_ERR_INVALID_ARG, "dp_expand_level_commit_gathered");
  DpDeviceGuard guard__(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rw = rec_words((int)ctx->views.size());
  long long total = 0;
  for (int r = 0; r < world; ++r) {
    if (counts[r] < 0 || counts[r] > segment_capacity)
      return dp_fail(ctx, DP_ERR_INVALID_ARG, "segment count out of range");
    total += counts[r];
  }
  if (total == 0) return dp_expand_level_commit(ctx, nullptr, 0, n_inserted, stream);
  if ((rc = dp_scratch_acquire(ctx, st)) != DP_OK) return rc;
  DP_CUDA(ctx, ctx->e_cells.ensure((size_t)total * rw * 4));
  long long off = 0;
  for (int r = 0; r < world; ++r) {
    if (counts[r] == 0) continue;
    const long long vw = counts[r] * (long long)rw;
    dp_compact_segments_kernel<<<(unsigned)((vw + 255) / 256), 256, 0, st>>>(
        (const uint32_t *)gathered_dev + (size_t)r * segment_capacity * rw, ctx->e_cells.as<uint32_t>(),
        segment_capacity * (long long)rw, vw, off);
    ++ctx->launches;
    off += vw;
  }
  DP_CUDA(ctx, cudaGetLastError());
  if ((rc = dp_scratch_release(ctx, st)) != DP_OK) return rc;
  return dp_expand_level_commit(ctx, ctx->e_cells.ptr, total, n_inserted, stream);
}

// Expand::ExpandPatches (expand.cpp:34-101) on one GPU.
extern "C" int dp_expand(dp_context *ctx, int cell_size, int max_levels, int64_t *stats) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  DpDeviceGuard guard__(ctx->device);
  DpOrganizer &o = ctx->org;
  o.frontier_begin = 0;  // queue <- all patches in the organizer (expand.cpp:45-48)
  int64_t st_pops = 0, st_cand = 0, st_pass = 0, st_ins = 0;
  const size_t rb = dp_record_bytes(ctx);
  for (int level = 0; max_levels < 0 || level < max_levels; ++level) {
    int64_t fb = 0, fe = 0;
    dp_expand_frontier(ctx, &fb, &fe);
    const long long nf = fe - fb;
    if (nf <= 0) break;
    // the survivors of a level are committed from a private buffer (e_cells is the commit's)
    DP_CUDA(ctx, ctx->e_recs.ensure((size_t)nf * 4 * rb));
    int64_t nrec = 0, ins = 0;
    rc = dp_expand_level_local(ctx, cell_size, 0, 1, nullptr, ctx->e_recs.ptr, nf * 4, &nrec,
                               ctx->stream);
    if (rc != DP_OK) return rc;
    const long long cand = ctx->org_last_candidates;
    rc = dp_expand_level_commit(ctx, ctx->e_recs.ptr, nrec, &ins, ctx->stream);
    if (rc != DP_OK) return rc;
    st_pops += nf;
    st_cand += cand;
    st_pass += nrec;
    st_ins += ins;
  }
  if (stats) {
    stats[0] = st_pops;
    stats[1] = st_cand;
    stats[2] = st_pass;
    stats[3] = st_ins;
  }
  return DP_OK;
}

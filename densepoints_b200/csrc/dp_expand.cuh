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

// ---- candidate record (what the allgather moves) ----------------------------------------
// u32 words: [0] seq, [1] ref, [2] nvis, [3..5] pos (f32 bits), [6..8] nrm, [9..9+vstride) vis
#define DP_REC_HDR 9
__host__ __device__ static inline size_t rec_words(int vstride) { return (size_t)(DP_REC_HDR + vstride); }

// (row, col) = ((size_t)(v / grid_scale), (size_t)(u / grid_scale)) with bounds test
// (patch_organizer.cpp:47-54, 15-30); negative / NaN quotients are out of bounds.
__device__ __forceinline__ long long dp_cell_of(const DpViewDev *__restrict__ V, double p0,
                                                double p1, double p2, double grid_scale) {
  double u, v;
  dp_project(V->P, p0, p1, p2, u, v);
  const double qr = v / grid_scale, qc = u / grid_scale;
  if (!(qr >= 0.0) || !(qc >= 0.0) || !(qr < 2147483647.0) || !(qc < 2147483647.0)) return -1;
  const long long row = (long long)qr, col = (long long)qc;
  if (col >= V->gw || row >= V->gh) return -1;
  return V->grid_off + row * V->gw + col;
}

// K6a: every (record, visible view) asks for its cell with atomicMin(seq).
__global__ void __launch_bounds__(256)
dp_claim_kernel(const DpViewDev *__restrict__ views, int n_views, const uint32_t *__restrict__ rec,
                long long n_rec, int vstride, double grid_scale,
                const uint8_t *__restrict__ grid, unsigned int *__restrict__ claim) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long r = t / vstride;
  const int k = (int)(t - r * vstride);
  if (r >= n_rec) return;
  const uint32_t *R = rec + (size_t)r * rec_words(vstride);
  if (k >= (int)R[2]) return;
  const int v = (int)R[DP_REC_HDR + k];
  if (v < 0 || v >= n_views) return;
  const long long cell = dp_cell_of(views + v, (double)__uint_as_float(R[3]),
                                    (double)__uint_as_float(R[4]), (double)__uint_as_float(R[5]),
                                    grid_scale);
  if (cell < 0 || grid[cell] != 0) return;  // occupied before this level
  atomicMin(claim + cell, R[0]);
}

// K6b: count the cells each record won; accepted iff > 1 (patch_organizer.cpp:58).
// One warp per record; flags[seq] = 1 for accepted records.
__global__ void __launch_bounds__(256)
dp_resolve_kernel(const DpViewDev *__restrict__ views, int n_views, const uint32_t *__restrict__ rec,
                  long long n_rec, int vstride, double grid_scale,
                  const unsigned int *__restrict__ claim, unsigned int *__restrict__ flags,
                  uint8_t *__restrict__ accepted) {
  const int lane = threadIdx.x & 31;
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rec) return;
  const uint32_t *R = rec + (size_t)r * rec_words(vstride);
  const int nv = min((int)R[2], vstride);
  const double p0 = (double)__uint_as_float(R[3]), p1 = (double)__uint_as_float(R[4]),
               p2 = (double)__uint_as_float(R[5]);
  unsigned wins = 0;
  for (int k = lane; k < nv; k += 32) {
    const int v = (int)R[DP_REC_HDR + k];
    if (v < 0 || v >= n_views) continue;
    const long long cell = dp_cell_of(views + v, p0, p1, p2, grid_scale);
    if (cell >= 0 && claim[cell] == R[0]) ++wins;
  }
  wins = __reduce_add_sync(DP_FULL, wins);
  if (lane == 0) {
    const bool acc = wins > 1;
    if (acc) flags[R[0]] = 1u;
    if (accepted) accepted[r] = acc ? 1 : 0;
  }
}

// K6c: winners occupy their cells (kept or not); every touched claim is reset.
__global__ void __launch_bounds__(256)
dp_occupy_kernel(const DpViewDev *__restrict__ views, int n_views, const uint32_t *__restrict__ rec,
                 long long n_rec, int vstride, double grid_scale, uint8_t *__restrict__ grid,
                 unsigned int *__restrict__ claim) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long r = t / vstride;
  const int k = (int)(t - r * vstride);
  if (r >= n_rec) return;
  const uint32_t *R = rec + (size_t)r * rec_words(vstride);
  if (k >= (int)R[2]) return;
  const int v = (int)R[DP_REC_HDR + k];
  if (v < 0 || v >= n_views) return;
  const long long cell = dp_cell_of(views + v, (double)__uint_as_float(R[3]),
                                    (double)__uint_as_float(R[4]), (double)__uint_as_float(R[5]),
                                    grid_scale);
  if (cell < 0) return;
  if (claim[cell] == R[0]) {
    grid[cell] = 1;
    claim[cell] = 0xffffffffu;
  }
}

// K6d: append accepted records to the store at n0 + rank(seq).
__global__ void __launch_bounds__(256)
dp_append_kernel(const uint32_t *__restrict__ rec, long long n_rec, int vstride,
                 const unsigned int *__restrict__ flags, const unsigned int *__restrict__ offs,
                 long long n0, float *__restrict__ pos, float *__restrict__ nrm,
                 int32_t *__restrict__ ref, int32_t *__restrict__ nvis, int32_t *__restrict__ vis) {
  const int lane = threadIdx.x & 31;
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rec) return;
  const uint32_t *R = rec + (size_t)r * rec_words(vstride);
  const uint32_t seq = R[0];
  if (flags[seq] == 0) return;
  const long long d = n0 + offs[seq];
  if (lane < 3) {
    pos[3 * d + lane] = __uint_as_float(R[3 + lane]);
    nrm[3 * d + lane] = __uint_as_float(R[6 + lane]);
  }
  if (lane == 0) {
    ref[d] = (int32_t)R[1];
    nvis[d] = (int32_t)R[2];
  }
  for (int k = lane; k < vstride; k += 32)
    vis[(size_t)d * vstride + k] = k < (int)R[2] ? (int32_t)R[DP_REC_HDR + k] : -1;
}

// SoA batch -> records, seq = seq0 + index (seeds) or seq[] (expansion), only where keep != 0.
// slot[i] = output position (exclusive scan of keep), or identity when slot == null.
__global__ void __launch_bounds__(256)
dp_pack_records_kernel(int n, int vstride, const float *__restrict__ pos,
                       const float *__restrict__ nrm, const int32_t *__restrict__ ref,
                       const int32_t *__restrict__ nvis, const int32_t *__restrict__ vis,
                       const unsigned int *__restrict__ seq, const uint8_t *__restrict__ keep,
                       const unsigned int *__restrict__ slot, uint32_t *__restrict__ rec) {
  const int lane = threadIdx.x & 31;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  if (keep && keep[i] == 0) return;
  uint32_t *R = rec + (size_t)(slot ? slot[i] : (unsigned)i) * rec_words(vstride);
  const int nv = min(max(nvis[i], 0), vstride);
  if (lane == 0) {
    R[0] = seq ? seq[i] : (uint32_t)i;
    R[1] = (uint32_t)ref[i];
    R[2] = (uint32_t)nv;
  }
  if (lane < 3) {
    R[3 + lane] = __float_as_uint(pos[3 * i + lane]);
    R[6 + lane] = __float_as_uint(nrm[3 * i + lane]);
  }
  for (int k = lane; k < vstride; k += 32)
    R[DP_REC_HDR + k] = (uint32_t)(k < nv ? vis[(size_t)i * vstride + k] : -1);
}

__global__ void dp_u8_to_u32_kernel(const uint8_t *__restrict__ in, unsigned int *__restrict__ out,
                                    long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] ? 1u : 0u;
}

// K5a: which frontier parents expand: >= 2 visible views (expand.cpp:69) and owned by `rank`.
__global__ void dp_parent_flags_kernel(const int32_t *__restrict__ nvis,
                                       const int32_t *__restrict__ ref, long long begin,
                                       long long n_f, const int32_t *__restrict__ rank_of_view,
                                       int rank, int n_views, unsigned int *__restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_f) return;
  const long long p = begin + i;
  bool f = nvis[p] >= 2;
  if (f && rank_of_view) {
    const int r = ref[p];
    f = (r >= 0 && r < n_views) && rank_of_view[r] == rank;
  }
  flags[i] = f ? 1u : 0u;
}

// K5b: Expand::ExpandPatch's proposals (expand.cpp:106-127): 4 children per expanding
// parent at +-(grid_scale / dx) along the patch x / y axes; children copy the parent.
__global__ void __launch_bounds__(128)
dp_propose_kernel(const DpViewDev *__restrict__ views, int n_views, long long begin, long long n_f,
                  const unsigned int *__restrict__ pflags, const unsigned int *__restrict__ pslot,
                  const float *__restrict__ spos, const float *__restrict__ snrm,
                  const int32_t *__restrict__ sref, const int32_t *__restrict__ snvis,
                  const int32_t *__restrict__ svis, int vstride, double grid_scale,
                  float *__restrict__ cpos, float *__restrict__ cnrm, int32_t *__restrict__ cref,
                  int32_t *__restrict__ cnvis, int32_t *__restrict__ cvis,
                  unsigned int *__restrict__ cseq) {
  const int lane = threadIdx.x & 31;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // warp / parent
  if (i >= n_f || pflags[i] == 0) return;
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
  const long long c0 = (long long)pslot[i] * 4;
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
  for (int t = lane; t < 4 * vstride; t += 32) {
    const int d = t / vstride, k = t - d * vstride;
    cvis[(size_t)(c0 + d) * vstride + k] = k < nv ? svis[(size_t)p * vstride + k] : -1;
  }
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
  const size_t vs = (size_t)o.vstride;
  struct Item {
    DpDevBuf *b;
    size_t elem;
  } items[] = {{&o.pos, 12}, {&o.nrm, 12}, {&o.rgb, 3}, {&o.ref, 4}, {&o.nvis, 4}, {&o.vis, 4 * vs}};
  for (auto &it : items) {
    void *np = nullptr;
    DP_CUDA(ctx, cudaMalloc(&np, it.elem * (size_t)cap));
    if (o.n > 0 && it.b->ptr)
      DP_CUDA(ctx, cudaMemcpyAsync(np, it.b->ptr, it.elem * (size_t)o.n, cudaMemcpyDeviceToDevice, st));
    DP_CUDA(ctx, cudaStreamSynchronize(st));
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
  cudaSetDevice(ctx->device);
  int rc = dp_sync_views(ctx);
  if (rc != DP_OK) return rc;
  DpOrganizer &o = ctx->org;
  cudaStream_t st = ctx->stream;
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
  const int vs = o.vstride, nviews = (int)ctx->views.size();
  const DpViewDev *views = ctx->d_views.as<DpViewDev>();
  const double gs = (double)ctx->prm.grid_scale;
  DP_CUDA(ctx, ctx->e_flags.ensure(((size_t)seq_space + 1) * 8));
  unsigned int *flags = ctx->e_flags.as<unsigned int>();
  unsigned int *offs = flags + (seq_space + 1);
  DP_CUDA(ctx, cudaMemsetAsync(flags, 0, ((size_t)seq_space + 1) * 4, st));
  const long long t1 = n_rec * vs;
  dp_claim_kernel<<<(unsigned)((t1 + 255) / 256), 256, 0, st>>>(
      views, nviews, rec, n_rec, vs, gs, o.grid.as<uint8_t>(), o.claim.as<unsigned int>());
  const long long t2 = n_rec * 32;
  dp_resolve_kernel<<<(unsigned)((t2 + 255) / 256), 256, 0, st>>>(
      views, nviews, rec, n_rec, vs, gs, o.claim.as<unsigned int>(), flags, accepted_dev);
  dp_occupy_kernel<<<(unsigned)((t1 + 255) / 256), 256, 0, st>>>(
      views, nviews, rec, n_rec, vs, gs, o.grid.as<uint8_t>(), o.claim.as<unsigned int>());
  ctx->launches += 3;
  DP_CUDA(ctx, cudaGetLastError());
  int rc = dp_exclusive_scan(ctx, flags, offs, seq_space, st);
  if (rc != DP_OK) return rc;
  long long total = 0;
  rc = dp_scan_total(ctx, flags, offs, seq_space, st, &total);
  if (rc != DP_OK) return rc;
  if (total > 0) {
    rc = org_reserve(ctx, o.n + total, st);
    if (rc != DP_OK) return rc;
    dp_append_kernel<<<(unsigned)((t2 + 255) / 256), 256, 0, st>>>(
        rec, n_rec, vs, flags, offs, o.n, o.pos.as<float>(), o.nrm.as<float>(),
        o.ref.as<int32_t>(), o.nvis.as<int32_t>(), o.vis.as<int32_t>());
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
  dp_patch_dev d;
  rc = upload_patches(ctx, h, &d, true);
  if (rc != DP_OK) return rc;
  DpOrganizer &o = ctx->org;
  cudaStream_t st = ctx->stream;
  const size_t rw = rec_words(o.vstride);
  DP_CUDA(ctx, ctx->e_cells.ensure((size_t)h->n * rw * 4));
  DP_CUDA(ctx, ctx->e_keep.ensure((size_t)h->n));
  uint32_t *rec = ctx->e_cells.as<uint32_t>();
  // repack to the organizer's vstride
  if (h->vstride > o.vstride) {
    // visible lists longer than n_views cannot be valid
    for (int i = 0; i < h->n; ++i)
      if (h->nvis[i] > o.vstride) return dp_fail(ctx, DP_ERR_INVALID_ARG, "nvis > number of views");
  }
  {
    // pack with the batch's own vstride into records of the organizer's vstride
    const long long t = (long long)h->n * 32;
    if (h->vstride == o.vstride) {
      dp_pack_records_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(
          h->n, o.vstride, d.pos, d.nrm, d.ref, d.nvis, d.vis, nullptr, nullptr, nullptr, rec);
    } else {
      // re-stride the visible table on the device: copy row by row with a 2D memcpy
      DP_CUDA(ctx, ctx->e_vis.ensure((size_t)h->n * o.vstride * 4));
      DP_CUDA(ctx, cudaMemsetAsync(ctx->e_vis.ptr, 0xff, (size_t)h->n * o.vstride * 4, st));
      const size_t wbytes = (size_t)std::min(h->vstride, o.vstride) * 4;
      DP_CUDA(ctx, cudaMemcpy2DAsync(ctx->e_vis.ptr, (size_t)o.vstride * 4, d.vis,
                                     (size_t)h->vstride * 4, wbytes, h->n,
                                     cudaMemcpyDeviceToDevice, st));
      dp_pack_records_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(
          h->n, o.vstride, d.pos, d.nrm, d.ref, d.nvis, ctx->e_vis.as<int32_t>(), nullptr, nullptr,
          nullptr, rec);
    }
    ++ctx->launches;
    DP_CUDA(ctx, cudaGetLastError());
  }
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
    const size_t wbytes = (size_t)std::min(out->vstride, o.vstride) * 4;
    if (out->vstride > o.vstride) memset(out->vis, 0xff, n * (size_t)out->vstride * 4);
    DP_CUDA(ctx, cudaMemcpy2DAsync(out->vis, (size_t)out->vstride * 4, o.vis.ptr,
                                   (size_t)o.vstride * 4, wbytes, n, cudaMemcpyDeviceToHost, st));
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

// Step 1 of a level: propose + refine + visibility + filter for the parents this rank owns;
// survivors are packed as records (ascending seq) straight into records_dev.
extern "C" int dp_expand_level_local(dp_context *ctx, int cell_size, int rank, int world,
                                     const int32_t *rank_of_view, void *records_dev,
                                     int64_t max_records, int64_t *n_records, void *stream) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  if (!n_records) return DP_ERR_INVALID_ARG;
  *n_records = 0;
  DpOrganizer &o = ctx->org;
  cudaStream_t st = (cudaStream_t)stream;  // as given: 0 is the legacy default stream (torch's)
  int64_t fb = 0, fe = 0;
  dp_expand_frontier(ctx, &fb, &fe);
  const long long nf = fe - fb;
  if (nf <= 0) return DP_OK;
  const int vs = o.vstride, nviews = (int)ctx->views.size();
  const DpViewDev *views = ctx->d_views.as<DpViewDev>();
  // ownership table on the device
  const int32_t *d_rov = nullptr;
  if (world > 1) {
    if (!rank_of_view) return dp_fail(ctx, DP_ERR_INVALID_ARG, "rank_of_view is null");
    DP_CUDA(ctx, ctx->s_misc.ensure((size_t)nviews * 4));
    DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_misc.ptr, rank_of_view, (size_t)nviews * 4,
                                 cudaMemcpyHostToDevice, st));
    d_rov = ctx->s_misc.as<int32_t>();
  }
  // K5a + scan: compact the expanding parents
  DP_CUDA(ctx, ctx->e_count.ensure(((size_t)nf + 1) * 8));
  unsigned int *pflags = ctx->e_count.as<unsigned int>();
  unsigned int *pslot = pflags + (nf + 1);
  dp_parent_flags_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(
      o.nvis.as<int32_t>(), o.ref.as<int32_t>(), fb, nf, d_rov, rank, nviews, pflags);
  ++ctx->launches;
  rc = dp_exclusive_scan(ctx, pflags, pslot, nf, st);
  if (rc != DP_OK) return rc;
  long long n_par = 0;
  rc = dp_scan_total(ctx, pflags, pslot, nf, st, &n_par);
  if (rc != DP_OK) return rc;
  if (n_par == 0) return DP_OK;
  const long long nc = n_par * 4;
  if (nc > 0x7fffffffLL) return dp_fail(ctx, DP_ERR_INVALID_ARG, "level too large");
  DP_CUDA(ctx, ctx->e_pos.ensure((size_t)nc * 12));
  DP_CUDA(ctx, ctx->e_nrm.ensure((size_t)nc * 12));
  DP_CUDA(ctx, ctx->e_ref.ensure((size_t)nc * 4));
  DP_CUDA(ctx, ctx->e_nvis.ensure((size_t)nc * 4));
  DP_CUDA(ctx, ctx->e_vis.ensure((size_t)nc * vs * 4));
  DP_CUDA(ctx, ctx->e_seq.ensure((size_t)nc * 4));
  DP_CUDA(ctx, ctx->e_keep.ensure((size_t)nc));
  const long long tp = nf * 32;
  dp_propose_kernel<<<(unsigned)((tp + 127) / 128), 128, 0, st>>>(
      views, nviews, fb, nf, pflags, pslot, o.pos.as<float>(), o.nrm.as<float>(),
      o.ref.as<int32_t>(), o.nvis.as<int32_t>(), o.vis.as<int32_t>(), vs,
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
  DP_CUDA(ctx, ctx->e_flags.ensure(((size_t)nc + 1) * 8));
  unsigned int *kflags = ctx->e_flags.as<unsigned int>();
  unsigned int *kslot = kflags + (nc + 1);
  dp_u8_to_u32_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(ctx->e_keep.as<uint8_t>(), kflags, nc);
  ++ctx->launches;
  rc = dp_exclusive_scan(ctx, kflags, kslot, nc, st);
  if (rc != DP_OK) return rc;
  long long n_keep = 0;
  rc = dp_scan_total(ctx, kflags, kslot, nc, st, &n_keep);
  if (rc != DP_OK) return rc;
  if (n_keep > max_records) return dp_fail(ctx, DP_ERR_INVALID_ARG, "records buffer too small");
  if (n_keep > 0) {
    const long long t = nc * 32;
    dp_pack_records_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(
        (int)nc, vs, c.pos, c.nrm, c.ref, c.nvis, c.vis, ctx->e_seq.as<unsigned int>(),
        ctx->e_keep.as<uint8_t>(), kslot, (uint32_t *)records_dev);
    ++ctx->launches;
    DP_CUDA(ctx, cudaGetLastError());
  }
  // stats: candidates refined this level (for dp_expand's counters)
  ctx->org_last_candidates = nc;
  *n_records = n_keep;
  return DP_OK;
}

// Step 3 of a level: TryInsert replay over the gathered records; advances the frontier.
extern "C" int dp_expand_level_commit(dp_context *ctx, const void *records_dev, int64_t n_records,
                                      int64_t *n_inserted, void *stream) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  DpOrganizer &o = ctx->org;
  cudaStream_t st = (cudaStream_t)stream;  // as given: 0 is the legacy default stream (torch's)
  int64_t fb = 0, fe = 0;
  dp_expand_frontier(ctx, &fb, &fe);
  const long long nf = fe - fb;
  const long long n_before = o.n;
  long long ins = 0;
  rc = org_commit(ctx, (const uint32_t *)records_dev, n_records, nf * 4, nullptr, &ins, st);
  if (rc != DP_OK) return rc;
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  o.pops += nf;
  o.frontier_begin = n_before;  // the next level = the patches appended by this one
  if (o.pops >= ctx->prm.max_pops) o.frontier_begin = o.n;  // expand.cpp:95-97: stop
  if (n_inserted) *n_inserted = ins;
  return DP_OK;
}

// Expand::ExpandPatches (expand.cpp:34-101) on one GPU.
extern "C" int dp_expand(dp_context *ctx, int cell_size, int max_levels, int64_t *stats) {
  int rc = org_check(ctx);
  if (rc != DP_OK) return rc;
  DpOrganizer &o = ctx->org;
  o.frontier_begin = 0;  // queue <- all patches in the organizer (expand.cpp:45-48)
  int64_t st_pops = 0, st_cand = 0, st_pass = 0, st_ins = 0;
  const size_t rb = dp_record_bytes(ctx);
  for (int level = 0; max_levels < 0 || level < max_levels; ++level) {
    int64_t fb = 0, fe = 0;
    dp_expand_frontier(ctx, &fb, &fe);
    const long long nf = fe - fb;
    if (nf <= 0) break;
    DP_CUDA(ctx, ctx->e_cells.ensure((size_t)nf * 4 * rb));
    int64_t nrec = 0, ins = 0;
    ctx->org_last_candidates = 0;
    rc = dp_expand_level_local(ctx, cell_size, 0, 1, nullptr, ctx->e_cells.ptr, nf * 4, &nrec,
                               ctx->stream);
    if (rc != DP_OK) return rc;
    rc = dp_expand_level_commit(ctx, ctx->e_cells.ptr, nrec, &ins, ctx->stream);
    if (rc != DP_OK) return rc;
    st_pops += nf;
    st_cand += ctx->org_last_candidates;
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

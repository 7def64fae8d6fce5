// dp_seed.cuh -- the steps immediately before and after the photometric path:
//   dp_create_patches  Seed::CreatePatchesFromPoints (reference methods/pmvs/seed.cpp:26-54):
//                      reference image = nearest camera centre (first minimum wins), normal =
//                      unit viewing ray, then Patch::InitRelatedImages.  Patch order = point
//                      order (the reference's `omp critical` push_back order is racy).
//   dp_export_ply      the patch store as an ASCII PLY in the layout of the reference's debug
//                      writer PMVS::PrintCloud (methods/pmvs/utils.cpp:9-50): x y z float,
//                      red green blue uchar, nx ny nz float.  Fills the declared-but-undefined
//                      PMVS::GetPointCloud (pmvs.h:21).
#pragma once
#include <stdio.h>

#include "dp_context.h"

__global__ void __launch_bounds__(256)
dp_create_patches_kernel(const DpViewDev *__restrict__ views, int n_views,
                         const double *__restrict__ points, int n, float *__restrict__ pos,
                         float *__restrict__ nrm, int32_t *__restrict__ ref) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = points[3 * i], y = points[3 * i + 1], z = points[3 * i + 2];
  int best = 0;
  double best_d = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0;
  for (int v = 0; v < n_views; ++v) {
    const double d0 = xsub(x, views[v].center[0]), d1 = xsub(y, views[v].center[1]),
                 d2 = xsub(z, views[v].center[2]);
    const double d = sqrt(xadd(xadd(xmul(d0, d0), xmul(d1, d1)), xmul(d2, d2)));
    if (v == 0 || d < best_d) {  // strict <: the first minimum wins (seed.cpp:36)
      best = v; best_d = d; b0 = d0; b1 = d1; b2 = d2;
    }
  }
  pos[3 * i] = (float)x; pos[3 * i + 1] = (float)y; pos[3 * i + 2] = (float)z;  // SetPosition
  nrm[3 * i] = (float)(b0 / best_d);  // patch_to_center / norm, SetNormal -> fp32
  nrm[3 * i + 1] = (float)(b1 / best_d);
  nrm[3 * i + 2] = (float)(b2 / best_d);
  ref[i] = best;
}

extern "C" int dp_create_patches(dp_context *ctx, const double *points, int n, dp_patch_soa *out,
                                 int32_t *ncand, int32_t *cand) {
  if (!ctx || !out || n < 0 || (n > 0 && !points)) return dp_fail(ctx, DP_ERR_INVALID_ARG, "dp_create_patches");
  if (out->n < n || out->vstride <= 0 || (n > 0 && (!out->pos || !out->nrm || !out->ref || !out->nvis || !out->vis)))
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "dp_create_patches: output capacity");
  out->n = n;
  if (n == 0) return DP_OK;
  DpDeviceGuard guard__(ctx->device);
  int rc = dp_sync_views(ctx);
  if (rc != DP_OK) return rc;
  cudaStream_t st = ctx->stream;
  const size_t N = (size_t)n, vs = (size_t)out->vstride;
  DP_CUDA(ctx, ctx->s_misc.ensure(N * 24));
  DP_CUDA(ctx, ctx->s_pos.ensure(N * 12));
  DP_CUDA(ctx, ctx->s_nrm.ensure(N * 12));
  DP_CUDA(ctx, ctx->s_ref.ensure(N * 4));
  DP_CUDA(ctx, ctx->s_nvis.ensure(N * 4));
  DP_CUDA(ctx, ctx->s_vis.ensure(N * vs * 4));
  DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_misc.ptr, points, N * 24, cudaMemcpyHostToDevice, st));
  dp_create_patches_kernel<<<(n + 255) / 256, 256, 0, st>>>(
      ctx->d_views.as<DpViewDev>(), (int)ctx->views.size(), ctx->s_misc.as<double>(), n,
      ctx->s_pos.as<float>(), ctx->s_nrm.as<float>(), ctx->s_ref.as<int32_t>());
  ++ctx->launches;
  DP_CUDA(ctx, cudaGetLastError());
  dp_patch_dev d;
  d.n = n; d.vstride = out->vstride;
  d.pos = ctx->s_pos.as<float>(); d.nrm = ctx->s_nrm.as<float>(); d.ref = ctx->s_ref.as<int32_t>();
  d.nvis = ctx->s_nvis.as<int32_t>(); d.vis = ctx->s_vis.as<int32_t>(); d.rgb = nullptr;
  if (ncand) DP_CUDA(ctx, ctx->s_ncand.ensure(N * 4));
  if (cand) DP_CUDA(ctx, ctx->s_cand.ensure(N * vs * 4));
  rc = dp_visibility_dev(ctx, &d, ncand ? ctx->s_ncand.as<int32_t>() : nullptr,
                         cand ? ctx->s_cand.as<int32_t>() : nullptr, st);
  if (rc != DP_OK) return rc;
  DP_CUDA(ctx, cudaMemcpyAsync(out->pos, d.pos, N * 12, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaMemcpyAsync(out->nrm, d.nrm, N * 12, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaMemcpyAsync(out->ref, d.ref, N * 4, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaMemcpyAsync(out->nvis, d.nvis, N * 4, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaMemcpyAsync(out->vis, d.vis, N * vs * 4, cudaMemcpyDeviceToHost, st));
  if (ncand) DP_CUDA(ctx, cudaMemcpyAsync(ncand, ctx->s_ncand.ptr, N * 4, cudaMemcpyDeviceToHost, st));
  if (cand) DP_CUDA(ctx, cudaMemcpyAsync(cand, ctx->s_cand.ptr, N * vs * 4, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  return DP_OK;
}

extern "C" int dp_export_ply(dp_context *ctx, const char *path) {
  if (!ctx || !path) return dp_fail(ctx, DP_ERR_INVALID_ARG, "dp_export_ply");
  if (!ctx->org.ready) return dp_fail(ctx, DP_ERR_STATE, "call dp_organizer_reset first");
  DpDeviceGuard guard__(ctx->device);
  const DpOrganizer &o = ctx->org;
  const size_t n = (size_t)o.n;
  std::vector<float> pos(n * 3), nrm(n * 3);
  std::vector<uint8_t> rgb(n * 3);
  if (n > 0) {
    DP_CUDA(ctx, cudaMemcpyAsync(pos.data(), o.pos.ptr, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    DP_CUDA(ctx, cudaMemcpyAsync(nrm.data(), o.nrm.ptr, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    DP_CUDA(ctx, cudaMemcpyAsync(rgb.data(), o.rgb.ptr, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
    DP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  FILE *f = fopen(path, "w");
  if (!f) return dp_fail(ctx, DP_ERR_INVALID_ARG, "dp_export_ply: cannot open file");
  fprintf(f, "ply\nformat ascii 1.0\nelement vertex %zu\n", n);
  fprintf(f, "property float x\nproperty float y\nproperty float z\n");
  fprintf(f, "property uchar red\nproperty uchar green\nproperty uchar blue\n");
  fprintf(f, "property float nx\nproperty float ny\nproperty float nz\nend_header\n");
  for (size_t i = 0; i < n; ++i)  // rply's ASCII writer: %g for float32, %d for uchar
    fprintf(f, "%g %g %g %d %d %d %g %g %g\n", pos[3 * i], pos[3 * i + 1], pos[3 * i + 2],
            (int)rgb[3 * i], (int)rgb[3 * i + 1], (int)rgb[3 * i + 2], nrm[3 * i], nrm[3 * i + 1],
            nrm[3 * i + 2]);
  const bool bad = ferror(f) != 0;
  fclose(f);
  if (bad) return dp_fail(ctx, DP_ERR_INVALID_ARG, "dp_export_ply: write failed");
  return DP_OK;
}

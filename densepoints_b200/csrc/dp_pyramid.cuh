// dp_pyramid.cuh -- K8: device-resident Gaussian pyramid (new; the reference's
// modules/image is an empty placeholder, modules/image/Image.h:1-7).  Level l+1 =
// cv::pyrDown(level l) on 8UC3: separable [1 4 6 4 1]/16 kernel, BORDER_REFLECT_101,
// output ((w+1)/2, (h+1)/2), exact integer sum then (sum + 128) >> 8.
#pragma once
#include "dp_context.h"

__device__ __forceinline__ int dp_reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
  return i;
}

__global__ void __launch_bounds__(256)
dp_pyrdown_kernel(const uint32_t *__restrict__ src, int sw, int sh, int spitch,
                  uint32_t *__restrict__ dst, int dw, int dh, int dpitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= dpitch || y >= dh) return;
  uint32_t outv = 0;
  if (x < dw) {
    const int wts[5] = {1, 4, 6, 4, 1};
    unsigned sb = 0, sg = 0, sr = 0;
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy) {
      const int yy = dp_reflect101(2 * y + dy, sh);
      unsigned rb = 0, rg = 0, rr = 0;
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        const int xx = dp_reflect101(2 * x + dx, sw);
        const uint32_t p = __ldg(src + (size_t)yy * spitch + xx);
        const unsigned w = wts[dx + 2];
        rb += w * (p & 0xff);
        rg += w * ((p >> 8) & 0xff);
        rr += w * ((p >> 16) & 0xff);
      }
      const unsigned wy = wts[dy + 2];
      sb += wy * rb;
      sg += wy * rg;
      sr += wy * rr;
    }
    outv = ((sb + 128) >> 8) | (((sg + 128) >> 8) << 8) | (((sr + 128) >> 8) << 16);
  }
  dst[(size_t)y * dpitch + x] = outv;
}

extern "C" int dp_build_pyramid(dp_context *ctx, int n_levels) {
  if (!ctx || n_levels < 1 || n_levels > 16) return dp_fail(ctx, DP_ERR_INVALID_ARG, "n_levels");
  DpDeviceGuard guard__(ctx->device);
  cudaStream_t st = ctx->stream;
  for (auto &v : ctx->views) {
    if (!v.set) return dp_fail(ctx, DP_ERR_STATE, "a view was not uploaded");
    while ((int)v.levels.size() > n_levels) {
      if (v.levels.back().img) cudaFree(v.levels.back().img);
      v.levels.pop_back();
    }
    while ((int)v.levels.size() < n_levels) {
      const DpLevel s = v.levels.back();
      DpLevel d;
      d.width = (s.width + 1) / 2;
      d.height = (s.height + 1) / 2;
      d.pitch_px = (d.width + 31) & ~31;
      DP_CUDA(ctx, cudaMalloc(&d.img, (size_t)d.pitch_px * (d.height + 1) * sizeof(uint32_t)));  // + spare row
      v.levels.push_back(d);  // owned by the view from here on (no leak on a later failure)
      DP_CUDA(ctx, cudaMemsetAsync(d.img + (size_t)d.pitch_px * d.height, 0, (size_t)d.pitch_px * 4, st));
      dim3 grid((d.pitch_px + 255) / 256, d.height);
      dp_pyrdown_kernel<<<grid, 256, 0, st>>>(s.img, s.width, s.height, s.pitch_px, d.img, d.width,
                                              d.height, d.pitch_px);
      ++ctx->launches;
    }
  }
  DP_CUDA(ctx, cudaGetLastError());
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->n_levels = n_levels;
  if (ctx->level >= n_levels) {
    ctx->level = 0;
    ctx->views_dirty = true;
    ctx->org.ready = false;
  }
  if (ctx->auto_level) ctx->views_dirty = true;  // the table of the levels above the base changed
  return DP_OK;
}

// Per-(patch, view) level selection.  Off: every view is read at the level
// dp_set_level chose.  On: that level is the BASE -- patch frame, visibility, proposals, grids and
// colours stay on it -- and GetProjectedTextures reads view v of a patch at base + k(p, v), k =
// the number of halvings that bring the patch's projected quad in v below px_per_cell pixels per
// texel (dp_pick_level), capped at the coarsest level built.
extern "C" int dp_set_level_selection(dp_context *ctx, int enable, double px_per_cell) {
  if (!ctx) return DP_ERR_INVALID_ARG;
  if (enable && !(px_per_cell > 0.0 && px_per_cell < 1e6))
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "px_per_cell");
  ctx->auto_level = enable != 0;
  if (enable) ctx->level_px = px_per_cell;
  ctx->views_dirty = true;
  return DP_OK;
}

extern "C" int dp_set_level(dp_context *ctx, int level) {
  if (!ctx || level < 0 || level >= ctx->n_levels) return dp_fail(ctx, DP_ERR_INVALID_ARG, "level");
  if (level != ctx->level) {
    ctx->level = level;
    ctx->views_dirty = true;
    ctx->org.ready = false;
  }
  return DP_OK;
}

extern "C" int dp_download_level(dp_context *ctx, int view_id, int level, uint8_t *bgr,
                                 size_t capacity, int *width, int *height) {
  if (!ctx || view_id < 0 || view_id >= (int)ctx->views.size() || !ctx->views[view_id].set)
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "view_id");
  const DpViewHost &v = ctx->views[view_id];
  if (level < 0 || level >= (int)v.levels.size()) return dp_fail(ctx, DP_ERR_INVALID_ARG, "level");
  const DpLevel &l = v.levels[level];
  if (width) *width = l.width;
  if (height) *height = l.height;
  if (!bgr) return DP_OK;
  const size_t bytes = (size_t)l.width * l.height * 3;
  if (capacity < bytes) return dp_fail(ctx, DP_ERR_INVALID_ARG, "capacity");
  DpDeviceGuard guard__(ctx->device);
  DP_CUDA(ctx, ctx->s_img.ensure(bytes));
  dim3 grid((l.width + 255) / 256, l.height);
  dp_unpack_bgrx_kernel<<<grid, 256, 0, ctx->stream>>>(l.img, l.pitch_px, l.width, l.height,
                                                        ctx->s_img.as<uint8_t>());
  ++ctx->launches;
  DP_CUDA(ctx, cudaGetLastError());
  DP_CUDA(ctx, cudaMemcpyAsync(bgr, ctx->s_img.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  DP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return DP_OK;
}

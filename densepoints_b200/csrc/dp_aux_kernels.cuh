// dp_aux_kernels.cuh -- the small integer / elementwise kernels around the photometric
// core: image packing, visibility (K4), colour (K7).
#pragma once
#include "dp_device.cuh"

// cv::imread layout (BGR u8, `stride` bytes per row) -> packed BGRx, pitch_px per row.
__global__ void dp_pack_bgrx_kernel(const uint8_t *__restrict__ bgr, size_t stride, int width,
                                    int height, uint32_t *__restrict__ out, int pitch_px) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= pitch_px || y >= height) return;
  uint32_t v = 0;
  if (x < width) {
    const uint8_t *p = bgr + (size_t)y * stride + 3 * (size_t)x;
    v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
  }
  out[(size_t)y * pitch_px + x] = v;
}

__global__ void dp_unpack_bgrx_kernel(const uint32_t *__restrict__ in, int pitch_px, int width,
                                      int height, uint8_t *__restrict__ bgr) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= width || y >= height) return;
  const uint32_t v = in[(size_t)y * pitch_px + x];
  uint8_t *p = bgr + ((size_t)y * width + x) * 3;
  p[0] = (uint8_t)(v & 0xff);
  p[1] = (uint8_t)((v >> 8) & 0xff);
  p[2] = (uint8_t)((v >> 16) & 0xff);
}

// K4: Patch::InitRelatedImages (patch.cpp:19-49).  One warp per patch, one view per lane
// per step; ballots keep the ascending-view-id order of the reference's push_back loop.
__global__ void __launch_bounds__(256)
dp_visibility_kernel(const DpViewDev *__restrict__ views, int n_views, int n,
                     const float *__restrict__ pos, const float *__restrict__ nrm,
                     const int32_t *__restrict__ ref, double t_vis, double t_cand,
                     int32_t *__restrict__ nvis, int32_t *__restrict__ vis,
                     int32_t *__restrict__ ncand, int32_t *__restrict__ cand, int vstride) {
  const int lane = threadIdx.x & 31;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const double p0 = pos[3 * i], p1 = pos[3 * i + 1], p2 = pos[3 * i + 2];
  const double n0 = nrm[3 * i], n1 = nrm[3 * i + 1], n2 = nrm[3 * i + 2];
  const int r = ref[i];
  int32_t *vi = vis + (size_t)i * vstride;
  int32_t *ci = cand ? cand + (size_t)i * vstride : nullptr;
  int nv = 0, nc = 0;
  const unsigned lt = (1u << lane) - 1u;
  for (int base = 0; base < n_views; base += 32) {
    const int v = base + lane;
    bool isv = false, isc = false;
    if (v < n_views && v != r) {
      const DpViewDev *V = views + v;
      double u, w;
      dp_project(V->P, p0, p1, p2, u, w);
      if (u > 0 && u < (double)V->width && w > 0 && w < (double)V->height) {  // IsPointInside
        const double d0 = xsub(p0, V->center[0]), d1 = xsub(p1, V->center[1]),
                     d2 = xsub(p2, V->center[2]);
        const double dot = xadd(xadd(xmul(n0, d0), xmul(n1, d1)), xmul(n2, d2));
        const double nn = sqrt(xadd(xadd(xmul(d0, d0), xmul(d1, d1)), xmul(d2, d2)));
        const double angle = acos(dot / nn);  // NaN (|arg| > 1) compares false twice
        if (angle < t_vis) isv = true;
        else if (angle < t_cand) isc = true;
      }
    }
    const unsigned mv = __ballot_sync(DP_FULL, isv), mc = __ballot_sync(DP_FULL, isc);
    if (isv) {
      const int k = nv + __popc(mv & lt);
      if (k < vstride) vi[k] = v;
    }
    if (isc && ci) {
      const int k = nc + __popc(mc & lt);
      if (k < vstride) ci[k] = v;
    }
    nv += __popc(mv);
    nc += __popc(mc);
  }
  nv = min(nv, vstride);
  for (int k = nv + lane; k < vstride; k += 32) vi[k] = -1;
  if (ci)
    for (int k = min(nc, vstride) + lane; k < vstride; k += 32) ci[k] = -1;
  if (lane == 0) {
    nvis[i] = nv;
    if (ncand) ncand[i] = nc;
  }
}

// K7: Patch::ComputeColor (patch.cpp:51-73): mean BGR of the pixel under the patch centre
// over ALL views that contain it; integer sums are exact, so the fp64 mean and its
// truncation to u8 do not depend on summation order.  No containing view -> 0 (0/0 in the
// reference).
__global__ void __launch_bounds__(256)
dp_color_kernel(const DpViewDev *__restrict__ views, int n_views, int n,
                const float *__restrict__ pos, uint8_t *__restrict__ rgb) {
  const int lane = threadIdx.x & 31;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const double p0 = pos[3 * i], p1 = pos[3 * i + 1], p2 = pos[3 * i + 2];
  unsigned sb = 0, sg = 0, sr = 0, cnt = 0;
  for (int v = lane; v < n_views; v += 32) {
    const DpViewDev *V = views + v;
    double u, w;
    dp_project(V->P, p0, p1, p2, u, w);
    if (u > 0 && u < (double)V->width && w > 0 && w < (double)V->height) {
      const uint32_t px = V->img[(size_t)((int)w) * V->pitch_px + (int)u];
      sb += px & 0xff;
      sg += (px >> 8) & 0xff;
      sr += (px >> 16) & 0xff;
      ++cnt;
    }
  }
  sb = __reduce_add_sync(DP_FULL, sb);
  sg = __reduce_add_sync(DP_FULL, sg);
  sr = __reduce_add_sync(DP_FULL, sr);
  cnt = __reduce_add_sync(DP_FULL, cnt);
  if (lane == 0) {
    uint8_t r = 0, g = 0, b = 0;
    if (cnt > 0) {
      r = (uint8_t)((double)sr / (double)cnt);
      g = (uint8_t)((double)sg / (double)cnt);
      b = (uint8_t)((double)sb / (double)cnt);
    }
    rgb[3 * i] = r;
    rgb[3 * i + 1] = g;
    rgb[3 * i + 2] = b;
  }
}

// densepoints/pmvs/batch.h -- std::vector<Patch> <-> dp_patch_soa marshalling.
// The loops over the patches are OpenMP-parallel when the including program is compiled with
// -fopenmp (the reference already links OpenMP, CMakeLists.txt:37-40): walking a million Patch
// objects and their heap-allocated index vectors is pointer chasing, ~200 ms single-threaded
// against ~27 ms of GPU time for the same batch (bench.py, e2e.cxx_mirror_pageable).
#ifndef DENSEPOINTS_B200_PMVS_BATCH
#define DENSEPOINTS_B200_PMVS_BATCH

#include <algorithm>
#include <vector>

#include "densepoints/pmvs/patch.h"
#include "densepoints_cuda.h"

namespace DensePoints {
namespace PMVS {

struct PatchBatch {
  std::vector<float> pos, nrm;
  std::vector<int32_t> ref, nvis, vis;
  std::vector<uint8_t> rgb;
  dp_patch_soa soa;

  template <typename PatchPtr>
  PatchBatch(PatchPtr const *patches, size_t n, int min_vstride = 1) {
    size_t vs = (size_t)std::max(min_vstride, 1);
#pragma omp parallel for reduction(max : vs) schedule(static)
    for (long long i = 0; i < (long long)n; ++i) vs = std::max(vs, patches[i]->GetTrullyVisibleImages().size());
    pos.resize(n * 3); nrm.resize(n * 3); ref.resize(n); nvis.resize(n); rgb.resize(n * 3);
    vis.resize(n * vs);
#pragma omp parallel for schedule(static)
    for (long long ii = 0; ii < (long long)n; ++ii) {
      const size_t i = (size_t)ii;
      const PointXYZRGBNormal p = patches[i]->GetPoint();
      pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z;
      nrm[3 * i] = p.normal_x; nrm[3 * i + 1] = p.normal_y; nrm[3 * i + 2] = p.normal_z;
      ref[i] = (int32_t)patches[i]->GetReferenceImage();
      const ImagesIndices &v = patches[i]->GetTrullyVisibleImages();
      nvis[i] = (int32_t)v.size();
      for (size_t k = 0; k < v.size(); ++k) vis[i * vs + k] = (int32_t)v[k];
      for (size_t k = v.size(); k < vs; ++k) vis[i * vs + k] = -1;
    }
    soa.n = (int32_t)n; soa.vstride = (int32_t)vs;
    soa.pos = pos.data(); soa.nrm = nrm.data(); soa.ref = ref.data();
    soa.nvis = nvis.data(); soa.vis = vis.data(); soa.rgb = rgb.data();
  }
  void StoreGeometry(Patch *const *patches) const {  // SetNormal / SetPosition (fp32 already)
#pragma omp parallel for schedule(static)
    for (int i = 0; i < soa.n; ++i) {
      PointXYZRGBNormal &p = patches[i]->Point();
      p.x = pos[3 * i]; p.y = pos[3 * i + 1]; p.z = pos[3 * i + 2];
      p.normal_x = nrm[3 * i]; p.normal_y = nrm[3 * i + 1]; p.normal_z = nrm[3 * i + 2];
    }
  }
  void StoreVisible(Patch *const *patches) const {
#pragma omp parallel
    {
    ImagesIndices v;  // one buffer per thread: the assignment below reuses the patch's storage
#pragma omp for schedule(static)
    for (int i = 0; i < soa.n; ++i) {
      v.clear();
      for (int k = 0; k < nvis[i]; ++k) v.push_back((size_t)vis[(size_t)i * soa.vstride + k]);
      if (v != patches[i]->GetTrullyVisibleImages()) patches[i]->SetTrullyVisibleImages(v);
    }
    }
  }
  void StoreColor(Patch *const *patches) const {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < soa.n; ++i) {
      PointXYZRGBNormal &p = patches[i]->Point();
      p.r = rgb[3 * i]; p.g = rgb[3 * i + 1]; p.b = rgb[3 * i + 2];
    }
  }
};

inline std::vector<Patch *> Pointers(Patches &patches) {
  std::vector<Patch *> out(patches.size());
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)patches.size(); ++i) out[(size_t)i] = &patches[(size_t)i];
  return out;
}

}  // namespace PMVS
}  // namespace DensePoints
#endif

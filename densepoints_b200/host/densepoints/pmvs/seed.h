// densepoints/pmvs/seed.h -- the batched bodies of Seed::FilterPatches / OptimizePatches /
// RemovePatches (reference methods/pmvs/seed.cpp:110-156) and of the patch-creation loop's
// visibility step (seed.cpp:26-54): one C-ABI call per loop instead of one Optimization
// object per patch.
#ifndef DENSEPOINTS_B200_PMVS_SEED
#define DENSEPOINTS_B200_PMVS_SEED

#include <chrono>
#include <utility>
#include <vector>

#include "densepoints/pmvs/optimization.h"

namespace DensePoints {
namespace PMVS {

class SeedCUDA {
 public:
  SeedCUDA(Session session, size_t cell_size = 16 /* MatcherOptions::cell_size, matcher.h:25 */,
           double score_threshold = 0.6, size_t minimum_visible_image = 3)
      : session_(session), cell_size_(cell_size), thr_(score_threshold), min_vis_(minimum_visible_image) {}

  void SetPatches(const Patches &p) { patches_ = p; }
  void GetPatches(Patches &p) const { p = patches_; }
  Patches &patches() { return patches_; }

  // Seed::CreatePatchesFromPoints (seed.cpp:26-54) for triangulated points: reference image =
  // nearest camera centre, normal = unit viewing ray, InitRelatedImages; patch order = point
  // order (the reference's omp-critical push_back order is racy).
  void CreatePatchesFromPoints(const std::vector<Vector3> &points) {
    const size_t n = points.size();
    const int vs = (int)session_->views()->size();
    std::vector<double> xyz(n * 3);
    for (size_t i = 0; i < n; ++i)
      for (int j = 0; j < 3; ++j) xyz[3 * i + j] = points[i][j];
    std::vector<float> pos(n * 3), nrm(n * 3);
    std::vector<int32_t> ref(n), nvis(n), vis(n * (size_t)vs, -1), ncand(n), cand(n * (size_t)vs, -1);
    dp_patch_soa s;
    s.n = (int32_t)n; s.vstride = vs;
    s.pos = pos.data(); s.nrm = nrm.data(); s.ref = ref.data(); s.nvis = nvis.data();
    s.vis = vis.data(); s.rgb = nullptr;
    session_->Check(dp_create_patches(session_->ctx(), xyz.data(), (int)n, &s, ncand.data(), cand.data()),
                    "dp_create_patches");
    patches_.assign(n, Patch());
    for (size_t i = 0; i < n; ++i) {
      Patch &p = patches_[i];
      p.SetReferenceImage((size_t)ref[i]);
      PointXYZRGBNormal &q = p.Point();
      q.x = pos[3 * i]; q.y = pos[3 * i + 1]; q.z = pos[3 * i + 2];
      q.normal_x = nrm[3 * i]; q.normal_y = nrm[3 * i + 1]; q.normal_z = nrm[3 * i + 2];
      ImagesIndices v, c;
      for (int k = 0; k < nvis[i]; ++k) v.push_back((size_t)vis[i * vs + k]);
      for (int k = 0; k < ncand[i] && k < vs; ++k) c.push_back((size_t)cand[i * vs + k]);
      p.SetTrullyVisibleImages(v);
      p.SetPotentiallyVisibleImages(c);
    }
  }

  // Patch::InitRelatedImages for every patch (seed.cpp:47)
  void InitRelatedImages() {
    if (patches_.empty()) return;
    std::vector<Patch *> ptr = Pointers(patches_);
    PatchBatch b(ptr.data(), ptr.size(), (int)session_->views()->size());
    std::vector<int32_t> ncand(ptr.size()), cand(ptr.size() * (size_t)b.soa.vstride, -1);
    session_->Check(dp_visibility(session_->ctx(), &b.soa, ncand.data(), cand.data()), "dp_visibility");
    b.StoreVisible(ptr.data());
    for (size_t i = 0; i < ptr.size(); ++i) {
      ImagesIndices c;
      for (int k = 0; k < ncand[i] && k < b.soa.vstride; ++k) c.push_back((size_t)cand[i * b.soa.vstride + k]);
      ptr[i]->SetPotentiallyVisibleImages(c);
    }
  }
  // Wall time of the stages of the last OptimizeAndRefinePatches() call, seconds: Patch -> SoA,
  // the C-ABI call (H2D, kernels, D2H), SoA -> Patch, RemovePatches.
  struct StageSeconds { double marshal = 0, call = 0, store = 0, remove = 0; };
  const StageSeconds &LastStageSeconds() const { return stage_; }

  void OptimizeAndRefinePatches() {  // seed.cpp:88-108: FilterPatches(); OptimizePatches();
    if (patches_.empty()) return;     // one upload / download for both stages
    typedef std::chrono::steady_clock clk;
    const clk::time_point t0 = clk::now();
    std::vector<Patch *> ptr = Pointers(patches_);
    PatchBatch b(ptr.data(), ptr.size());
    OptimizationCUDA::WithThresholds guard(*session_, thr_, min_vis_);
    std::vector<uint8_t> keep(ptr.size());
    evals_.assign(ptr.size(), 0);
    const clk::time_point t1 = clk::now();
    session_->Check(dp_filter_refine(session_->ctx(), &b.soa, (int)cell_size_, keep.data(), evals_.data()),
                    "dp_filter_refine");
    const clk::time_point t2 = clk::now();
    b.StoreVisible(ptr.data());
    b.StoreGeometry(ptr.data());
    const clk::time_point t3 = clk::now();
    std::vector<size_t> to_remove;
    for (size_t i = 0; i < keep.size(); ++i)
      if (!keep[i]) to_remove.push_back(i);
    RemovePatches(to_remove);
    const clk::time_point t4 = clk::now();
    stage_.marshal = std::chrono::duration<double>(t1 - t0).count();
    stage_.call = std::chrono::duration<double>(t2 - t1).count();
    stage_.store = std::chrono::duration<double>(t3 - t2).count();
    stage_.remove = std::chrono::duration<double>(t4 - t3).count();
  }
  void FilterPatches() {  // seed.cpp:110-126
    if (patches_.empty()) return;
    std::vector<Patch *> ptr = Pointers(patches_);
    PatchBatch b(ptr.data(), ptr.size());
    OptimizationCUDA::WithThresholds guard(*session_, thr_, min_vis_);
    std::vector<uint8_t> keep(ptr.size());
    session_->Check(dp_filter(session_->ctx(), &b.soa, (int)cell_size_, keep.data()), "dp_filter");
    b.StoreVisible(ptr.data());
    std::vector<size_t> to_remove;
    for (size_t i = 0; i < keep.size(); ++i)
      if (!keep[i]) to_remove.push_back(i);
    RemovePatches(to_remove);
  }
  void OptimizePatches() {  // seed.cpp:128-144 (Optimize always returns true: nothing removed)
    if (patches_.empty()) return;
    std::vector<Patch *> ptr = Pointers(patches_);
    PatchBatch b(ptr.data(), ptr.size());
    evals_.assign(ptr.size(), 0);
    session_->Check(dp_refine(session_->ctx(), &b.soa, (int)cell_size_, nullptr, evals_.data(), nullptr), "dp_refine");
    b.StoreGeometry(ptr.data());
  }
  // seed.cpp:146-156 erases the listed patches one by one (ascending indices, order kept) --
  // quadratic in the batch size.  Same result, one stable pass.
  void RemovePatches(const std::vector<size_t> &patch_indices) {
    if (patch_indices.empty()) return;
    size_t w = 0, k = 0;
    for (size_t i = 0; i < patches_.size(); ++i) {
      if (k < patch_indices.size() && patch_indices[k] == i) {
        ++k;
        continue;
      }
      if (w != i) patches_[w] = std::move(patches_[i]);
      ++w;
    }
    patches_.resize(w);
  }
  const std::vector<int32_t> &LastEvals() const { return evals_; }

 private:
  Session session_;
  size_t cell_size_;
  double thr_;
  size_t min_vis_;
  Patches patches_;
  std::vector<int32_t> evals_;
  StageSeconds stage_;
};

}  // namespace PMVS
}  // namespace DensePoints
#endif

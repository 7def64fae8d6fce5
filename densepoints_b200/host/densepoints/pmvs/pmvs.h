// densepoints/pmvs/pmvs.h -- mirror of the method facade (reference methods/pmvs/pmvs.h:14-35,
// pmvs.cpp:11-43): AddCamera collects the views, Run = InsertSeeds + ExpandSeeds.  The views
// are uploaded to the GPU once, when the first stage needs them (the reference loads each image
// in AddCamera, pmvs.cpp:13); seed *generation* (Matcher::GenerateSeeds, modules/features) is
// outside the photometric path, so the triangulated seed points are handed in with SetSeedPoints.
#ifndef DENSEPOINTS_B200_PMVS_PMVS
#define DENSEPOINTS_B200_PMVS_PMVS

#include <memory>
#include <string>
#include <vector>

#include "densepoints/pmvs/expand.h"
#include "densepoints/pmvs/seed.h"

namespace DensePoints {
namespace PMVS {

// options.h:8-21 -- declared by the reference and never read by it; `expansions`
// is honoured here as the cap on BFS levels (BASELINE's "3 expansion rounds"), -1 = until empty.
class Options {
 public:
  Options(const int scale = 1, const int cell_size = 4, const int expansions = -1)
      : scale_(scale), cell_size_(cell_size), expansions_(expansions) {}
  int expansions() const { return expansions_; }

 protected:
  int scale_;
  int cell_size_;
  int expansions_;
};

class PMVS {
 public:
  PMVS(const Options &options = Options(), int device = -1) : options_(options), device_(device) {
    views_ = std::make_shared<std::vector<View>>();
  }
  // pmvs.cpp:11-20: views without an image are discarded
  void AddCamera(View view) {
    if (view.ImageLoaded()) views_->push_back(view);
  }
  void SetSeedPoints(const std::vector<Vector3> &points) { seed_points_ = points; }
  bool Run() {  // pmvs.cpp:22-27
    InsertSeeds();
    ExpandSeeds();
    return true;
  }
  // pmvs.h:21 declares GetPointCloud and never defines it; the patches of the organizer
  // (position, normal, colour = PointXYZRGBNormal) are the cloud
  std::shared_ptr<std::vector<PointXYZRGBNormal>> GetPointCloud() {
    auto cloud = std::make_shared<std::vector<PointXYZRGBNormal>>();
    if (expand_)
      for (const Patch &p : expand_->GetPatches()) cloud->push_back(p.GetPoint());
    return cloud;
  }
  void WritePly(const std::string &path) {
    if (expand_) expand_->WritePly(path);
  }
  Views views() const { return views_; }

 protected:
  void InsertSeeds() {  // pmvs.cpp:29-34 (GenerateSeeds replaced by SetSeedPoints)
    if (!session_) session_ = std::make_shared<CudaSession>(views_, device_);
    seeds_ = std::make_shared<SeedCUDA>(session_);
    seeds_->CreatePatchesFromPoints(seed_points_);  // Seed::ConvertSeedsToPatches, seed.cpp:20-24
    seeds_->OptimizeAndRefinePatches();
  }
  void ExpandSeeds() {  // pmvs.cpp:36-43
    Patches seeds;
    seeds_->GetPatches(seeds);
    expand_ = std::make_shared<Expand>(session_);
    expand_->SetSeeds(seeds, options_.expansions());
  }

  Options options_;
  int device_;
  Views views_;
  Session session_;
  std::vector<Vector3> seed_points_;
  std::shared_ptr<SeedCUDA> seeds_;
  std::shared_ptr<Expand> expand_;
};

}  // namespace PMVS
}  // namespace DensePoints
#endif

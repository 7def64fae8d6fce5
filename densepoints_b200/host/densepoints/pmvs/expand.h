// densepoints/pmvs/expand.h -- mirror of Expand (reference methods/pmvs/expand.h:10-32,
// expand.cpp:13-143) and PatchOrganizer (patch_organizer.h:49-78) with the store and the
// occupancy grids resident on the GPU.  Order = the reference's 1-thread FIFO order.
#ifndef DENSEPOINTS_B200_PMVS_EXPAND
#define DENSEPOINTS_B200_PMVS_EXPAND

#include <string>
#include <vector>

#include "densepoints/pmvs/optimization.h"

namespace DensePoints {
namespace PMVS {

struct ExpandOptions {
  size_t cell_size;
  ExpandOptions(size_t cell_size = 11) : cell_size(cell_size) {}
};

// patch_organizer.h:40-47
struct PatchOrganizerOptions {
  size_t max_patches_per_cell;
  size_t grid_scale;
  PatchOrganizerOptions(size_t max_patches_per_cell = 1, size_t grid_scale = 8)
      : max_patches_per_cell(max_patches_per_cell), grid_scale(grid_scale) {}
};

class Expand {
 public:
  Expand(Session session, ExpandOptions options = ExpandOptions()) : session_(session), options_(options) {}
  void SetOptions(const ExpandOptions expand_options) { options_ = expand_options; }
  // The reference's Expand::SetSeeds builds its PatchOrganizer with default options
  // (expand.cpp:16); PatchOrganizer::SetOptions (patch_organizer.h:56-58) is how a caller changes
  // them.  Here they travel in dp_params; call before SetSeeds.
  void SetOrganizerOptions(const PatchOrganizerOptions o) {
    dp_params p = session_->Params();
    p.max_patches_per_cell = (int32_t)o.max_patches_per_cell;
    p.grid_scale = (int32_t)o.grid_scale;
    session_->SetParams(p);
  }

  // expand.cpp:13-32: new organizer, AllocateViews, SetSeeds, ExpandPatches
  void SetSeeds(const Patches seeds, int max_levels = -1) {
    session_->Check(dp_organizer_reset(session_->ctx()), "dp_organizer_reset");
    if (!seeds.empty()) {
      std::vector<const Patch *> ptr;
      for (const Patch &p : seeds) ptr.push_back(&p);
      PatchBatch b(ptr.data(), ptr.size(), (int)session_->views()->size());
      session_->Check(dp_organizer_insert(session_->ctx(), &b.soa, nullptr), "dp_organizer_insert");
    }
    ExpandPatches(max_levels);
  }
  void ExpandPatches(int max_levels = -1) {  // expand.cpp:34-101
    session_->Check(dp_expand(session_->ctx(), (int)options_.cell_size, max_levels, stats_), "dp_expand");
  }
  // PatchOrganizer::GetPatches (patch_organizer.h:62)
  Patches GetPatches() {
    const int64_t n = dp_organizer_size(session_->ctx());
    const int vs = (int)session_->views()->size();
    Patches out((size_t)n);
    if (n == 0) return out;
    std::vector<float> pos(n * 3), nrm(n * 3);
    std::vector<int32_t> ref(n), nvis(n), vis((size_t)n * vs);
    std::vector<uint8_t> rgb(n * 3);
    dp_patch_soa s;
    s.n = (int32_t)n; s.vstride = vs;
    s.pos = pos.data(); s.nrm = nrm.data(); s.ref = ref.data(); s.nvis = nvis.data();
    s.vis = vis.data(); s.rgb = rgb.data();
    session_->Check(dp_organizer_export(session_->ctx(), &s), "dp_organizer_export");
    for (int64_t i = 0; i < n; ++i) {
      PointXYZRGBNormal &p = out[i].Point();
      p.x = pos[3 * i]; p.y = pos[3 * i + 1]; p.z = pos[3 * i + 2];
      p.normal_x = nrm[3 * i]; p.normal_y = nrm[3 * i + 1]; p.normal_z = nrm[3 * i + 2];
      p.r = rgb[3 * i]; p.g = rgb[3 * i + 1]; p.b = rgb[3 * i + 2];
      out[i].SetReferenceImage((size_t)ref[i]);
      ImagesIndices v;
      for (int k = 0; k < nvis[i]; ++k) v.push_back((size_t)vis[(size_t)i * vs + k]);
      out[i].SetTrullyVisibleImages(v);
    }
    return out;
  }
  // the point cloud as an ASCII PLY in the layout of PMVS::PrintCloud (utils.cpp:9-50); stands
  // in for the declared-but-undefined PMVS::GetPointCloud (pmvs.h:21)
  void WritePly(const std::string &path) {
    session_->Check(dp_export_ply(session_->ctx(), path.c_str()), "dp_export_ply");
  }
  const int64_t *Stats() const { return stats_; }  // pops, candidates, passed, inserted

 private:
  Session session_;
  ExpandOptions options_;
  int64_t stats_[4] = {0, 0, 0, 0};
};

}  // namespace PMVS
}  // namespace DensePoints
#endif

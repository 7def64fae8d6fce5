// densepoints/pmvs/patch.h -- mirror of the reference's methods/pmvs/patch.h:21-101.
// Position / normal are stored as fp32 like pcl::PointXYZRGBNormal (core/types.h:27).
#ifndef DENSEPOINTS_B200_PMVS_PATCH
#define DENSEPOINTS_B200_PMVS_PATCH

#include <map>
#include <vector>

#include "densepoints/core/types.h"

namespace DensePoints {
namespace PMVS {

typedef std::vector<size_t> ImagesIndices;
typedef std::map<size_t, std::pair<size_t, size_t>> PatchCells;

struct PointXYZRGBNormal {
  float x = 0, y = 0, z = 0, normal_x = 0, normal_y = 0, normal_z = 0;
  uint8_t r = 0, g = 0, b = 0;
};

class Patch {
 public:
  void SetReferenceImage(size_t image_index) { reference_image_ = image_index; }
  size_t GetReferenceImage() const { return reference_image_; }
  void SetNormal(Vector3 n) { point_.normal_x = (float)n[0]; point_.normal_y = (float)n[1]; point_.normal_z = (float)n[2]; }
  Vector3 GetNormal() const { return Vector3(point_.normal_x, point_.normal_y, point_.normal_z); }
  void SetPosition(Vector3 p) { point_.x = (float)p[0]; point_.y = (float)p[1]; point_.z = (float)p[2]; }
  Vector3 GetPosition() const { return Vector3(point_.x, point_.y, point_.z); }
  const ImagesIndices &GetTrullyVisibleImages() const { return visible_images_; }
  const ImagesIndices &GetPotentiallyVisibleImages() const { return candidate_images_; }
  void SetTrullyVisibleImages(const ImagesIndices &v) { visible_images_ = v; }
  void SetPotentiallyVisibleImages(const ImagesIndices &v) { candidate_images_ = v; }
  void RemoveTrullyVisibleImage(size_t index) { visible_images_.erase(visible_images_.begin() + index); }
  void SetPatchCells(const PatchCells &c) { patch_cells_ = c; }
  const PointXYZRGBNormal GetPoint() const { return point_; }
  PointXYZRGBNormal &Point() { return point_; }

 private:
  PointXYZRGBNormal point_;
  size_t reference_image_ = 0;
  ImagesIndices visible_images_, candidate_images_;
  PatchCells patch_cells_;
};
typedef std::vector<Patch> Patches;

}  // namespace PMVS
}  // namespace DensePoints
#endif

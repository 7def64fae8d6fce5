// densepoints/pmvs/optimization.h -- mirror of the reference's plugin interface for the
// path: class Optimization (methods/pmvs/optimization.h:11-45) and its CUDA back end
// OptimizationCUDA (the counterpart of OptimizationOpenCV, optimization_opencv.h:10-14).
#ifndef DENSEPOINTS_B200_PMVS_OPTIMIZATION
#define DENSEPOINTS_B200_PMVS_OPTIMIZATION

#include <cmath>
#include <vector>

#include "densepoints/pmvs/batch.h"
#include "densepoints/pmvs/cuda_session.h"

namespace DensePoints {
namespace PMVS {

// cv::Mat stand-in for a projected texture: empty() where the reference pushes cv::Mat().
struct Texture {
  int size = 0;
  std::vector<uint8_t> bgr;  // size*size*3
  bool empty() const { return bgr.empty(); }
};

class Optimization {
 public:
  Optimization(Patch &patch, Views views, size_t cell_size, double score_threshold = 0.6,
               size_t minimum_visible_image = 3)
      : patch_(patch), views_(views), cell_size_(cell_size),
        minimum_visible_image_(minimum_visible_image), score_threshold_(score_threshold) {}
  virtual ~Optimization() {}
  virtual bool Optimize() = 0;
  virtual void GetProjectedTextures(std::vector<Texture> &textures) = 0;
  virtual bool FilterByErrorMeasurement() = 0;
  // optimization.cpp:78-96 (pure host arithmetic, identical to the reference)
  void UnparametrizePatch(double depth, double roll, double pitch, Vector3 &normal, Vector3 &position) {
    const Vector3 c = (*views_)[patch_.GetReferenceImage()].GetCameraCenter();
    const Vector3 cur = patch_.GetPosition();
    position = c + (cur - c) * (1 + depth);
    double ca = std::cos(roll), sa = std::sin(roll), cb = std::cos(pitch), sb = std::sin(pitch);
    const Vector3 n = patch_.GetNormal();
    normal = Vector3(cb * n[0] + 0 * n[1] + -sb * n[2], sa * sb * n[0] + ca * n[1] + cb * sa * n[2],
                     ca * sb * n[0] + -sa * n[1] + ca * cb * n[2]);
  }

 protected:
  Patch &patch_;
  Views views_;
  size_t cell_size_;
  size_t minimum_visible_image_;
  double score_threshold_;
};

// Batch-of-1 adapter: keeps per-patch call sites (e.g. Seed::PrintTextures, seed.cpp:194-199)
// working; the batched drivers in seed.h / expand.h are what production code calls.
class OptimizationCUDA : public Optimization {
 public:
  OptimizationCUDA(Session session, Patch &patch, size_t cell_size, double score_threshold = 0.6,
                   size_t minimum_visible_image = 3)
      : Optimization(patch, session->views(), cell_size, score_threshold, minimum_visible_image),
        session_(session) {}

  bool Optimize() override {  // optimization_opencv.cpp:44-78; always true
    Patch *p = &patch_;
    PatchBatch b(&p, 1);
    session_->Check(dp_refine(session_->ctx(), &b.soa, (int)cell_size_, nullptr, nullptr, nullptr), "dp_refine");
    b.StoreGeometry(&p);
    return true;
  }
  void GetProjectedTextures(std::vector<Texture> &textures) override {  // optimization.cpp:9-56
    Patch *p = &patch_;
    PatchBatch b(&p, 1);
    const int s = (int)cell_size_, nv = b.nvis[0];
    std::vector<float> ncc(b.soa.vstride);
    std::vector<uint8_t> tex((size_t)b.soa.vstride * s * s * 3), valid(b.soa.vstride);
    session_->Check(dp_score(session_->ctx(), &b.soa, s, ncc.data(), tex.data(), valid.data()), "dp_score");
    for (int k = 0; k < nv; ++k) {
      Texture t;
      if (valid[k]) {
        t.size = s;
        t.bgr.assign(tex.begin() + (size_t)k * s * s * 3, tex.begin() + (size_t)(k + 1) * s * s * 3);
      }
      textures.push_back(t);
    }
  }
  bool FilterByErrorMeasurement() override {  // optimization.cpp:98-132
    Patch *p = &patch_;
    PatchBatch b(&p, 1);
    WithThresholds guard(*session_, score_threshold_, minimum_visible_image_);
    uint8_t keep = 0;
    session_->Check(dp_filter(session_->ctx(), &b.soa, (int)cell_size_, &keep), "dp_filter");
    b.StoreVisible(&p);
    return keep != 0;
  }

  // per-call ctor arguments of Optimization -> library parameters, restored on exit
  struct WithThresholds {
    CudaSession &s;
    dp_params saved;
    WithThresholds(CudaSession &sess, double thr, size_t min_vis) : s(sess), saved(sess.Params()) {
      dp_params p = saved;
      p.score_threshold = thr;
      p.minimum_visible_image = (int32_t)min_vis;
      s.SetParams(p);
    }
    ~WithThresholds() { dp_set_params(s.ctx(), &saved); }
  };

 private:
  Session session_;
};

}  // namespace PMVS
}  // namespace DensePoints
#endif

// densepoints/pmvs/cuda_session.h -- owns the dp_context for a set of Views: images and
// cameras are uploaded once (PMVS::AddCamera, reference pmvs.cpp:11-20) and stay resident.
#ifndef DENSEPOINTS_B200_PMVS_CUDA_SESSION
#define DENSEPOINTS_B200_PMVS_CUDA_SESSION

#include <memory>
#include <stdexcept>
#include <string>

#include "densepoints/core/types.h"
#include "densepoints_cuda.h"

namespace DensePoints {
namespace PMVS {

class CudaSession {
 public:
  explicit CudaSession(Views views, int device = -1, const dp_params *params = nullptr) : views_(views) {
    Check(dp_create(&ctx_, device, params), "dp_create");
    Check(dp_set_num_views(ctx_, (int)views->size()), "dp_set_num_views");
    for (size_t i = 0; i < views->size(); ++i) {
      const View &v = (*views)[i];
      ProjectionMatrix P = v.GetProjectionMatrix();
      Vector3 xa = v.GetXAxis(), c = v.GetCameraCenter();
      const Image &im = v.GetImage();
      Check(dp_upload_view(ctx_, (int)i, P.data(), xa.v, c.v, im.data(), im.cols, im.rows, im.step),
            "dp_upload_view");
    }
  }
  ~CudaSession() { dp_destroy(ctx_); }
  CudaSession(const CudaSession &) = delete;
  CudaSession &operator=(const CudaSession &) = delete;
  dp_context *ctx() const { return ctx_; }
  Views views() const { return views_; }
  // The library reports errors as status codes; the mirror turns them into exceptions the
  // way OpenCV calls inside the reference would (cv::Exception).
  void Check(int rc, const char *what) const {
    if (rc != DP_OK)
      throw std::runtime_error(std::string(what) + ": " + (ctx_ ? dp_last_error(ctx_) : "no context"));
  }
  dp_params Params() const { dp_params p; dp_get_params(ctx_, &p); return p; }
  void SetParams(const dp_params &p) { Check(dp_set_params(ctx_, &p), "dp_set_params"); }

 private:
  Views views_;
  dp_context *ctx_ = nullptr;
};
typedef std::shared_ptr<CudaSession> Session;

}  // namespace PMVS
}  // namespace DensePoints
#endif

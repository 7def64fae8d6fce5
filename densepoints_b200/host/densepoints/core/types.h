// densepoints/core/types.h -- host-side mirror of the reference's modules/core/types.h for
// the photometric path, without OpenCV / Eigen / PCL (none of them exist in this image).
// Same names and argument meaning as the reference (View, Views, Vector3, ProjectionMatrix);
// the image is a plain BGR u8 buffer with cv::Mat's row layout.
#ifndef DENSEPOINTS_B200_CORE_TYPES
#define DENSEPOINTS_B200_CORE_TYPES

#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <vector>

namespace DensePoints {

struct Vector2 {
  double v[2];
  double operator[](int i) const { return v[i]; }
  double &operator[](int i) { return v[i]; }
};

struct Vector3 {
  double v[3];
  Vector3() : v{0, 0, 0} {}
  Vector3(double x, double y, double z) : v{x, y, z} {}
  double operator[](int i) const { return v[i]; }
  double &operator[](int i) { return v[i]; }
  Vector3 operator+(const Vector3 &o) const { return {v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]}; }
  Vector3 operator-(const Vector3 &o) const { return {v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]}; }
  Vector3 operator*(double s) const { return {v[0] * s, v[1] * s, v[2] * s}; }
  double dot(const Vector3 &o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
  Vector3 cross(const Vector3 &o) const {
    return {v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]};
  }
  double norm() const { return std::sqrt(dot(*this)); }
  Vector3 normalized() const { double n = norm(); return {v[0] / n, v[1] / n, v[2] / n}; }
};

// 3x4, row-major (Eigen::Matrix<double,3,4> of the reference)
typedef std::array<double, 12> ProjectionMatrix;

// cv::Mat stand-in for CV_8UC3: rows x cols BGR, `step` bytes per row.  Shares its buffer.
struct Image {
  int rows = 0, cols = 0;
  size_t step = 0;
  std::shared_ptr<std::vector<uint8_t>> buf;
  const uint8_t *data() const { return buf ? buf->data() : nullptr; }
  bool empty() const { return !buf || rows == 0 || cols == 0; }
  static Image Create(int rows, int cols) {
    Image m;
    m.rows = rows;
    m.cols = cols;
    m.step = (size_t)cols * 3;
    m.buf = std::make_shared<std::vector<uint8_t>>((size_t)rows * m.step);
    return m;
  }
};

// reference: modules/core/types.h:37-75, types.cpp:28-89
class View {
 public:
  View(const ProjectionMatrix &projection_matrix) { SetProjectionMatrix(projection_matrix); }
  View(const ProjectionMatrix &projection_matrix, const Image &image) : image_(image), image_loaded_(true) {
    SetProjectionMatrix(projection_matrix);
  }
  const Image &GetImage() const { return image_; }
  void SetImage(const Image &image) { image_ = image; image_loaded_ = !image.empty(); }
  bool ImageLoaded() const { return image_loaded_; }

  // types.cpp:28-68: centre = null vector of P; x axis = row 0 of the orthogonal factor of
  // the RQ decomposition with a positive-diagonal K.
  void SetProjectionMatrix(const ProjectionMatrix &P);
  ProjectionMatrix GetProjectionMatrix() const { return projection_matrix_; }
  Vector3 GetCameraCenter() const { return camera_center_; }
  Vector3 GetXAxis() const { return x_axis_; }
  Vector2 ProjectPoint(const Vector3 &p) const {  // types.cpp:70-75
    const ProjectionMatrix &P = projection_matrix_;
    double x = P[0] * p[0] + P[1] * p[1] + P[2] * p[2] + P[3];
    double y = P[4] * p[0] + P[5] * p[1] + P[6] * p[2] + P[7];
    double w = P[8] * p[0] + P[9] * p[1] + P[10] * p[2] + P[11];
    return Vector2{{x / w, y / w}};
  }
  bool IsPointInside(const Vector3 &p) const {  // types.cpp:77-84
    Vector2 q = ProjectPoint(p);
    return q[0] > 0 && q[0] < image_.cols && q[1] > 0 && q[1] < image_.rows;
  }
  size_t Height() const { return image_.rows; }
  size_t Width() const { return image_.cols; }

 private:
  ProjectionMatrix projection_matrix_;
  Vector3 camera_center_, x_axis_;
  Image image_;
  bool image_loaded_ = false;
};
typedef std::shared_ptr<std::vector<View>> Views;

inline void View::SetProjectionMatrix(const ProjectionMatrix &P) {
  projection_matrix_ = P;
  Vector3 m0(P[0], P[1], P[2]), m1(P[4], P[5], P[6]), m2(P[8], P[9], P[10]);
  Vector3 b(-P[3], -P[7], -P[11]);
  Vector3 c0 = m1.cross(m2), c1 = m2.cross(m0), c2 = m0.cross(m1);
  double det = m0.dot(c0);
  camera_center_ = (c0 * b[0] + c1 * b[1] + c2 * b[2]) * (1.0 / det);
  Vector3 r2 = m2.normalized();
  Vector3 r1 = (m1 - r2 * m1.dot(r2)).normalized();
  x_axis_ = (m0 - r2 * m0.dot(r2) - r1 * m0.dot(r1)).normalized();
}

}  // namespace DensePoints
#endif

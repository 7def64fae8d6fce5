// pmvs_facade_check.cpp -- programs/densify (reference main.cpp:29-36) against the mirrored
// facade: AddCamera per view, Run, GetPointCloud.  Built by tests/test_gpu_host_mirror.py to
// prove that the facade compiles and links against the C ABI; with a scene dump as argument it
// also runs (seed points instead of Matcher::GenerateSeeds, which is outside the path).
#include <cstdio>
#include <stdexcept>
#include <vector>

#include "densepoints/pmvs/pmvs.h"

using namespace DensePoints;

int main(int argc, char **argv) {
  if (argc < 2) {
    std::printf("usage: %s scene.bin [out.ply]\n", argv[0]);
    return 0;
  }
  try {
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) throw std::runtime_error("cannot open input");
    int32_t hdr[3];
    if (std::fread(hdr, 4, 3, f) != 3) throw std::runtime_error("short read");
    PMVS::PMVS pmvs(PMVS::Options(1, 4, 3));
    for (int v = 0; v < hdr[0]; ++v) {
      ProjectionMatrix P;
      Image im = Image::Create(hdr[2], hdr[1]);
      if (std::fread(P.data(), 8, 12, f) != 12 ||
          std::fread(im.buf->data(), 1, im.buf->size(), f) != im.buf->size())
        throw std::runtime_error("short read");
      pmvs.AddCamera(View(P, im));
    }
    int32_t n = 0;
    if (std::fread(&n, 4, 1, f) != 1) throw std::runtime_error("short read");
    std::vector<Vector3> pts(n);
    for (auto &p : pts)
      if (std::fread(p.v, 8, 3, f) != 3) throw std::runtime_error("short read");
    std::fclose(f);
    pmvs.SetSeedPoints(pts);
    pmvs.Run();
    std::printf("%zu views, %zu seed points, %zu points in the cloud\n", pmvs.views()->size(), pts.size(),
                pmvs.GetPointCloud()->size());
    if (argc > 2) pmvs.WritePly(argv[2]);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}

// host_mirror_test.cpp -- drives the C++ host mirror (densepoints_b200/host) the way
// methods/pmvs drives its own classes: seeds -> InitRelatedImages -> FilterPatches ->
// OptimizePatches (Seed, reference seed.cpp:88-144) -> Expand::SetSeeds (expand.cpp:13-32),
// plus the per-patch OptimizationCUDA adapter.  Reads a scene dump, writes the results;
// tests/test_gpu_host_mirror.py compares them with the CPU oracle.
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "densepoints/pmvs/expand.h"
#include "densepoints/pmvs/seed.h"

using namespace DensePoints;
using namespace DensePoints::PMVS;

template <typename T>
static void rd(FILE *f, T *p, size_t n) {
  if (fread(p, sizeof(T), n, f) != n) throw std::runtime_error("short read");
}
template <typename T>
static void wr(FILE *f, const T *p, size_t n) {
  if (fwrite(p, sizeof(T), n, f) != n) throw std::runtime_error("short write");
}

static void dump(FILE *f, const Patches &ps, int n_views) {
  int32_t n = (int32_t)ps.size();
  wr(f, &n, 1);
  for (const Patch &p : ps) {
    const PointXYZRGBNormal q = p.GetPoint();
    float g[6] = {q.x, q.y, q.z, q.normal_x, q.normal_y, q.normal_z};
    wr(f, g, 6);
    uint8_t c[3] = {q.r, q.g, q.b};
    wr(f, c, 3);
    int32_t ref = (int32_t)p.GetReferenceImage(), nv = (int32_t)p.GetTrullyVisibleImages().size();
    wr(f, &ref, 1);
    wr(f, &nv, 1);
    std::vector<int32_t> vis(n_views, -1);
    for (int k = 0; k < nv; ++k) vis[k] = (int32_t)p.GetTrullyVisibleImages()[k];
    wr(f, vis.data(), n_views);
  }
}

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  try {
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) throw std::runtime_error("cannot open input");
    int32_t hdr[3];
    rd(fi, hdr, 3);
    const int n_views = hdr[0], W = hdr[1], H = hdr[2];
    Views views = std::make_shared<std::vector<View>>();
    for (int v = 0; v < n_views; ++v) {
      ProjectionMatrix P;
      rd(fi, P.data(), 12);
      Image im = Image::Create(H, W);
      rd(fi, im.buf->data(), (size_t)H * W * 3);
      views->push_back(View(P, im));
    }
    int32_t n;
    rd(fi, &n, 1);
    std::vector<float> pos(n * 3), nrm(n * 3);
    std::vector<int32_t> ref(n);
    rd(fi, pos.data(), n * 3);
    rd(fi, nrm.data(), n * 3);
    rd(fi, ref.data(), n);
    int32_t cfg[4];  // seed cell size, expand cell size, minimum_visible_image, max_levels
    rd(fi, cfg, 4);
    fclose(fi);

    Session session = std::make_shared<CudaSession>(views, 0);
    // Seed::CreatePatchesFromPoints (seed.cpp:40-47) with the given reference images
    Patches seeds(n);
    for (int i = 0; i < n; ++i) {
      seeds[i].SetReferenceImage(ref[i]);
      seeds[i].SetPosition(Vector3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
      seeds[i].SetNormal(Vector3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
    }
    SeedCUDA seed(session, cfg[0], 0.6, cfg[2]);
    seed.SetPatches(seeds);
    seed.InitRelatedImages();

    FILE *fo = fopen(argv[2], "wb");
    if (!fo) throw std::runtime_error("cannot open output");
    // per-patch adapter on the first seed: textures, then a copy filtered + optimised
    {
      Patch p0 = seed.patches()[0];
      OptimizationCUDA opt(session, p0, cfg[0], 0.6, cfg[2]);
      std::vector<Texture> tex;
      opt.GetProjectedTextures(tex);
      int32_t nt = (int32_t)tex.size();
      wr(fo, &nt, 1);
      for (const Texture &t : tex) {
        int32_t sz = t.empty() ? 0 : t.size;
        wr(fo, &sz, 1);
        if (sz) wr(fo, t.bgr.data(), t.bgr.size());
      }
      int32_t keep = opt.FilterByErrorMeasurement() ? 1 : 0;
      wr(fo, &keep, 1);
      opt.Optimize();
      Patches one(1, p0);
      dump(fo, one, n_views);
    }
    seed.OptimizeAndRefinePatches();
    Patches refined;
    seed.GetPatches(refined);
    dump(fo, refined, n_views);

    dp_params prm = session->Params();
    prm.minimum_visible_image = cfg[2];
    session->SetParams(prm);
    Expand expand(session, ExpandOptions(cfg[1]));
    expand.SetSeeds(refined, cfg[3]);
    dump(fo, expand.GetPatches(), n_views);
    wr(fo, expand.Stats(), 4);
    // Seed::CreatePatchesFromPoints on the raw seed positions
    {
      std::vector<Vector3> pts;
      for (int i = 0; i < n; ++i) pts.push_back(Vector3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
      SeedCUDA created(session, cfg[0], 0.6, cfg[2]);
      created.CreatePatchesFromPoints(pts);
      dump(fo, created.patches(), n_views);
    }
    fclose(fo);
    if (argc > 3) expand.WritePly(argv[3]);
    std::printf("host mirror ok: %d seeds -> %zu refined -> %zu patches after expansion\n", n,
                refined.size(), expand.GetPatches().size());
  } catch (const std::exception &e) {
    std::fprintf(stderr, "host_mirror_test: %s\n", e.what());
    return 1;
  }
  return 0;
}

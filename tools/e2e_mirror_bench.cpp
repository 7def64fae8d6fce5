// e2e_mirror_bench.cpp -- end-to-end time of the call a maintainer's code would make:
// SeedCUDA::OptimizeAndRefinePatches() (= Seed::OptimizeAndRefinePatches, reference
// methods/pmvs/seed.cpp:88-108) on a std::vector<Patch> in ordinary (pageable) memory, through the
// C++ host mirror of the reference's classes (densepoints_b200/host) and the C ABI:
// Patch objects -> PatchBatch marshalling -> dp_filter_refine (H2D, kernels, D2H) -> visible sets
// and geometry written back into the Patch objects -> RemovePatches.
// Input: the scene dump bench.py writes; output: one JSON line.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "densepoints/pmvs/seed.h"

using namespace DensePoints;
using namespace DensePoints::PMVS;

template <typename T>
static void rd(FILE *f, T *p, size_t n) {
  if (fread(p, sizeof(T), n, f) != n) throw std::runtime_error("short read");
}

int main(int argc, char **argv) {
  if (argc < 4) {
    std::printf("usage: e2e_mirror_bench scene.bin cell_size steps\n");
    return 2;
  }
  try {
    const int cell = atoi(argv[2]), steps = atoi(argv[3]);
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
    std::vector<float> pos((size_t)n * 3), nrm((size_t)n * 3);
    std::vector<int32_t> ref(n);
    rd(fi, pos.data(), (size_t)n * 3);
    rd(fi, nrm.data(), (size_t)n * 3);
    rd(fi, ref.data(), n);
    fclose(fi);
    Session session = std::make_shared<CudaSession>(views, 0);
    Patches seeds(n);
    for (int i = 0; i < n; ++i) {
      seeds[i].SetReferenceImage(ref[i]);
      seeds[i].SetPosition(Vector3(pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2]));
      seeds[i].SetNormal(Vector3(nrm[3 * (size_t)i], nrm[3 * (size_t)i + 1], nrm[3 * (size_t)i + 2]));
    }
    SeedCUDA seed(session, (size_t)cell);
    seed.SetPatches(seeds);
    seed.InitRelatedImages();  // Patch::InitRelatedImages (set-up, untimed)
    Patches ready;
    seed.GetPatches(ready);
    long long visible = 0;
    for (const Patch &p : ready) visible += (long long)p.GetTrullyVisibleImages().size();
    double total_s = 0.0, st_marshal = 0, st_call = 0, st_store = 0, st_remove = 0;
    long long evals = 0, refined = 0;
    for (int it = -1; it < steps; ++it) {  // it = -1: warm-up
      seed.SetPatches(ready);               // a fresh std::vector<Patch> (untimed)
      const auto t0 = std::chrono::steady_clock::now();
      seed.OptimizeAndRefinePatches();
      const auto t1 = std::chrono::steady_clock::now();
      if (it < 0) continue;
      total_s += std::chrono::duration<double>(t1 - t0).count();
      st_marshal += seed.LastStageSeconds().marshal; st_call += seed.LastStageSeconds().call;
      st_store += seed.LastStageSeconds().store; st_remove += seed.LastStageSeconds().remove;
      // evaluations of the step: every visible view once in the filter + evals x views of the
      // survivors (LastEvals is indexed by the patches before removal; survivors keep their order)
      const std::vector<int32_t> &ev = seed.LastEvals();
      const Patches &out = seed.patches();
      size_t k = 0;
      long long e = visible;
      for (size_t i = 0; i < ev.size(); ++i)
        if (ev[i] > 0) {
          e += (long long)ev[i] * (long long)out[k].GetTrullyVisibleImages().size();
          ++k;
        }
      if (k != out.size()) throw std::runtime_error("survivor bookkeeping");
      evals += e;
      refined += (long long)out.size();
    }
    std::printf("{\"steps\": %d, \"seconds\": %.6f, \"evals\": %lld, \"refined\": %lld, \"patches\": %d, "
                "\"marshal_s\": %.6f, \"call_s\": %.6f, \"store_s\": %.6f, \"remove_s\": %.6f}\n",
                steps, total_s, evals, refined, n, st_marshal, st_call, st_store, st_remove);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "e2e_mirror_bench: %s\n", e.what());
    return 1;
  }
  return 0;
}

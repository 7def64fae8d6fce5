// host_marshal_bench.cpp -- what the std::vector<Patch> <-> SoA marshalling of the C++ host mirror
// costs for a 1 M-patch batch (no GPU needed): Pointers, PatchBatch, StoreVisible, StoreGeometry,
// KeepPatches, and how it scales with OpenMP threads.
//   g++ -O2 -std=c++14 -fopenmp -Idensepoints_b200/host -Iinclude tools/host_marshal_bench.cpp -o hmb
//   for t in 1 4 16; do OMP_NUM_THREADS=$t ./hmb; done
#include <chrono>
#include <cstdio>
#include <random>
#include "densepoints/pmvs/batch.h"
using namespace DensePoints; using namespace DensePoints::PMVS;
static double now(){ return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(){
  const int n=1<<20; std::mt19937 g(1);
  Patches ps(n);
  for(int i=0;i<n;++i){ ps[i].SetReferenceImage(i%16); ImagesIndices v; int nv=2+g()%7; for(int k=0;k<nv;++k) v.push_back((i+k)%16); ps[i].SetTrullyVisibleImages(v);}  
  Patches ref = ps;
  double t0=now(); std::vector<Patch*> ptr=Pointers(ps); double t1=now();
  PatchBatch b(ptr.data(), ptr.size()); double t2=now();
  for(int i=0;i<n;++i) if(i%3==0 && b.nvis[i]>2) b.nvis[i]--;
  double t3=now(); b.StoreVisible(ptr.data()); double t4=now(); b.StoreGeometry(ptr.data()); double t5=now();
  std::vector<uint8_t> keep(n); for(int i=0;i<n;++i) keep[i]= (i%100)>=72;
  KeepPatches(ps, keep); double t6=now();
  // check
  size_t k=0; bool ok=true; for(int i=0;i<n;++i) if(keep[i]){ const auto&a=ps[k].GetTrullyVisibleImages(); auto e=ref[i].GetTrullyVisibleImages(); if(i%3==0&&e.size()>2) e.pop_back(); if(a!=e||ps[k].GetReferenceImage()!=ref[i].GetReferenceImage()) ok=false; ++k;}
  printf("pointers %.1f ms, batch ctor %.1f ms (vs=%d), store visible %.1f ms, store geometry %.1f ms, keep %.1f ms -> %zu patches, %s\n",(t1-t0)*1e3,(t2-t1)*1e3,b.soa.vstride,(t4-t3)*1e3,(t5-t4)*1e3,(t6-t5)*1e3,ps.size(), ok&&k==ps.size()?"OK":"MISMATCH");
}

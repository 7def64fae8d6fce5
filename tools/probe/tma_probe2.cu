#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, int x, int y, uint32_t *out, int stage) {
  __shared__ __align__(128) uint32_t tile[256];
  __shared__ __align__(8) uint64_t bar;
  const int lane = threadIdx.x;
  for (int t = lane; t < 256; t += 32) tile[t] = 0xdeadbeef;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    if (stage >= 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0) {
    if (stage <= 1) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    } else if (stage == 2) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    } else {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1024) : "memory");
      if (stage == 3)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(tile)), "l"(&tm), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
      else
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(tile)), "l"(&tm), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    }
  }
  uint32_t done; int spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  } while (!done && ++spins < 2000000);
  for (int t = lane; t < 256; t += 32) out[t] = tile[t];
  if (lane == 0) out[256] = spins;
}
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                           const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
  int stage = argc > 1 ? atoi(argv[1]) : 0;
  int variant = argc > 2 ? atoi(argv[2]) : 0;
  const int H = 480, pitch = 640;
  std::vector<uint32_t> img((size_t)pitch * H);
  for (int yy = 0; yy < H; ++yy) for (int xx = 0; xx < pitch; ++xx) img[(size_t)yy * pitch + xx] = (yy << 16) | xx;
  uint32_t *d_img, *d_out;
  cudaMalloc(&d_img, img.size() * 4);
  cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, 2048);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  alignas(64) CUtensorMap tm;
  const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)H};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {16, 16};
  if (variant == 1) { box[0] = 32; box[1] = 8; }
  const cuuint32_t es[2] = {1, 1};
  CUresult r = ((enc_fn)p)(&tm, variant == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_img, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           variant == 3 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("stage %d variant %d encode -> %d; ", stage, variant, (int)r);
  probe<<<1, 32>>>(tm, argc > 3 ? atoi(argv[3]) : 37, 101, d_out, stage);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run: %s; ", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    uint32_t o[257];
    cudaMemcpy(o, d_out, 257 * 4, cudaMemcpyDeviceToHost);
    printf("spins %u o[0]=%08x o[17]=%08x", o[256], o[0], o[17]);
  }
  printf("\n");
  return 0;
}

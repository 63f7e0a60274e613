// Does a 1-D tiled TMA load accept start coordinates that are not 16-byte aligned?
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, int coord, int box, float* out) {
  __shared__ alignas(128) float buf[256];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(box * 4));
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];"
                 ::"r"(smem_u32(buf)), "l"(&map), "r"(smem_u32(&bar)), "r"(coord) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  if (threadIdx.x < box) out[threadIdx.x] = buf[threadIdx.x];
}
int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int n = 4096; float *d, *o; cudaMalloc(&d, n * 4); cudaMalloc(&o, 1024);
  float h[n]; for (int i = 0; i < n; ++i) h[i] = (float)i; cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
  for (int box : {100, 108}) {
    CUtensorMap m; cuuint64_t dims[1] = {(cuuint64_t)n}; cuuint64_t st[1] = {0}; cuuint32_t bx[1] = {(cuuint32_t)box}; cuuint32_t es[1] = {1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, d, dims, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode box %d -> %d\n", box, (int)r);
    for (int coord : {0, 4, 196, -4, -200, 4092, 5000}) {
      k<<<1, 128>>>(m, coord, box, o);
      cudaError_t e = cudaDeviceSynchronize();
      float ho[4] = {0}; if (e == cudaSuccess) cudaMemcpy(ho, o, 16, cudaMemcpyDeviceToHost);
      printf("  coord %5d: %s  first %g %g %g %g\n", coord, cudaGetErrorString(e), ho[0], ho[1], ho[2], ho[3]);
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}

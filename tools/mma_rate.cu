// Microbenchmark: issue rate of tcgen05.mma kind::tf32 / kind::f16 (M=128, N variable, K=32 bytes)
// with the A operand in TMEM (.ts) or in shared memory (.ss), no other traffic on the SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
template <int KIND, int TS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_t, uint64_t a_d, uint64_t b_d,
                                    uint32_t idesc, uint32_t acc) {
  if (TS) {
    if (KIND == 0)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                   "r"(a_t), "l"(b_d), "r"(idesc), "r"(acc) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                   "r"(a_t), "l"(b_d), "r"(idesc), "r"(acc) : "memory");
  } else {
    if (KIND == 0)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                   "l"(a_d), "l"(b_d), "r"(idesc), "r"(acc) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                   "l"(a_d), "l"(b_d), "r"(idesc), "r"(acc) : "memory");
  }
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {   // SW128 K-major, 128B rows
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

// nacc: number of distinct accumulators cycled through (1 = one dependent chain)
template <int KIND, int TS>
__global__ void __launch_bounds__(512, 1) rate_kernel(int N, int iters, int nacc, long long* out, int agg,
                                                      int pace, long long* agg_ops) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  if (threadIdx.x == 0) done = 0;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    const uint32_t fmt = (KIND == 0) ? 2u : 1u;   // tf32 : bf16
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t tmu = __shfl_sync(0xffffffffu, tm, 0);
    const uint64_t a_d = make_desc(sb), b_d = make_desc(sb + 16384);
    const uint32_t d1 = tmu + (nacc > 1 ? (uint32_t)N : 0u);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t pred;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
      if (pred) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mma<KIND, TS>((j & 1) ? d1 : tmu, tmu + 384 + 8 * j, a_d + 2 * j, b_d + 2 * j, idesc, 1u);
      }
      __syncwarp();
    }
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      mbar_wait(&bar, 0);
      out[blockIdx.x] = clock64() - t0;
      done = 1;
    }
    __syncwarp();
  }
  else if (threadIdx.x >= 128 && threadIdx.x < 256 && (agg & 1)) {
    // aggressor 1: tcgen05.st 32 columns (x 128 lanes = 16 KB) per op into columns 256..
    const uint32_t w = (threadIdx.x >> 5) & 3;
    uint32_t v[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = threadIdx.x + e;
    long long n = 0;
    while (!done) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
          "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(tm + ((w * 32) << 16) + 256 + (uint32_t)((n & 1) * 32)),
          "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      ++n;
      if (pace) __nanosleep(pace);
    }
    if (threadIdx.x == 128) agg_ops[blockIdx.x * 4 + 0] = n;
  } else if (threadIdx.x >= 256 && threadIdx.x < 384 && (agg & 2)) {
    // aggressor 2: tcgen05.ld 32 columns of the accumulator (x 128 lanes = 16 KB) per op
    const uint32_t w = (threadIdx.x >> 5) & 3;
    long long n = 0;
    uint32_t sink = 0;
    while (!done) {
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
          "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tm + ((w * 32) << 16) + 128 + (uint32_t)((n & 1) * 32)) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int e = 0; e < 32; ++e) sink ^= v[e];
      ++n;
      if (pace) __nanosleep(pace);
    }
    if (sink == 0x12345u) agg_ops[blockIdx.x * 4 + 3] = sink;
    if (threadIdx.x == 256) agg_ops[blockIdx.x * 4 + 1] = n;
  } else if (threadIdx.x >= 384 && (agg & 4)) {
    // aggressor 3: 128-bit shared loads, 2 KB per warp op (conflict-free), of the A region
    long long n = 0;
    float acc = 0.f;
    const int t = threadIdx.x - 384;
    while (!done) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(smem_u32(smem + ((t * 16 + k * 2048) & 16383))));
        acc += v.x + v.y + v.z + v.w;
      }
      ++n;
      if (pace) __nanosleep(pace);
    }
    if (acc == 1.2345f) agg_ops[blockIdx.x * 4 + 3] = 1;
    if (threadIdx.x == 384) agg_ops[blockIdx.x * 4 + 2] = n;
  }
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

template <int KIND, int TS>
void run(const char* tag, int N, int nacc, int grid, int agg = 0, int pace = 0) {
  long long *d, *a; cudaMalloc(&d, sizeof(long long) * grid); cudaMalloc(&a, sizeof(long long) * grid * 4);
  cudaMemset(a, 0, sizeof(long long) * grid * 4);
  const int iters = 2000;
  cudaFuncSetAttribute(rate_kernel<KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  rate_kernel<KIND, TS><<<grid, 512, 60000>>>(N, iters, nacc, d, agg, pace, a);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148], ha[148 * 4]; cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  cudaMemcpy(ha, a, sizeof(long long) * grid * 4, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  const double clk = (double)h[0];
  printf("%-14s N=%3d agg=%d pace=%4d: %6.1f clk/MMA | st %.1f B/clk  ld %.1f B/clk  lds %.1f B/clk (%s)\n", tag, N,
         agg, pace, avg / (iters * 4.0), ha[0] * 16384.0 / clk, ha[1] * 16384.0 / clk,
         ha[2] * 4 * 8 * 512.0 / clk, cudaGetErrorString(e));
  cudaFree(d); cudaFree(a);
}

int main() {
  const int grid = 148;
  for (int N : {64, 128, 256}) {
    run<0, 1>("tf32 A=TMEM", N, 1, grid);
    run<0, 0>("tf32 A=smem", N, 1, grid);
  }
  for (int agg : {1, 2, 4, 3, 7})
    for (int pace : {0, 200, 1000}) run<0, 1>("tf32 A=TMEM", 128, 2, grid, agg, pace);
  for (int agg : {1, 2, 4}) run<0, 0>("tf32 A=smem", 128, 2, grid, agg, 0);
  return 0;
}

// Measures the accumulation bias of tcgen05.mma kind::tf32 (fp32 accumulator in TMEM): a chain of m
// MMAs (M=128, N=64, K=8 each) over exactly representable TF32 operands against the fp64 sum of the
// same products.  The tensor core adds into its accumulator with truncation, so a one-signed chain
// comes out SHORT by a relative amount that grows linearly with m; conv_tc.cu / first_tc.cu /
// fire_tc.cu cut chains into segments of <= 36 MMAs and scale each segment sum by
// 1 + bias_comp * m.  This tool is where bias_comp comes from (profiles/r2_mma_bias.txt):
//   per distribution and m:  mean signed (D_gpu - D_exact) / D_exact, its slope per MMA, and the
//   residual after the compensation.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bias mma_bias.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <random>
#include <vector>
#include "../squeezedet_b200/csrc/tc_ptx.cuh"

using namespace sqdet;

constexpr int N = 64, P = 32;       // P distinct K steps (8 wide) cycled through by the chain
constexpr int A_LBO = 128 * 16, B_LBO = N * 16;   // bytes between 16-byte K chunks (no swizzle)

__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc) : "memory");
}

// smem A: [P*2 chunks][128 rows][4 floats], B: [P*2 chunks][N rows][4 floats]  (K-major, SWIZZLE_NONE:
// rows 16 B apart, 8-row groups 128 B apart, chunks LBO apart)
__global__ void __launch_bounds__(128, 1) k(const float* a, const float* b, float* out, int m) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  float* sa = (float*)smem;
  float* sb = (float*)(smem + P * 2 * A_LBO);
  for (int i = threadIdx.x; i < P * 2 * 128 * 4; i += 128) sa[i] = a[i];
  for (int i = threadIdx.x; i < P * 2 * N * 4; i += 128) sb[i] = b[i];
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 64);
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    for (int j = 0; j < m; ++j) {
      const int ks = j % P;
      const uint64_t da = (uint64_t)(((smem_u32(sa) + ks * 2 * A_LBO) & 0x3FFFFu) >> 4) |
                          ((uint64_t)(A_LBO >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
      const uint64_t db = (uint64_t)(((smem_u32(sb) + ks * 2 * B_LBO) & 0x3FFFFu) >> 4) |
                          ((uint64_t)(B_LBO >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
      umma_ss(tm, da, db, idesc, j ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16_nowait(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    for (int e = 0; e < 16; ++e) out[(warp * 32 + lane) * N + c0 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 64);
}

static float tf32(float x) {
  uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r;
}

int main(int argc, char** argv) {
  const double comp = argc > 1 ? atof(argv[1]) : 1.4e-8;
  const size_t smem = (size_t)P * 2 * (A_LBO + B_LBO) + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  float *da, *db, *dout;
  cudaMalloc(&da, P * 2 * 128 * 16); cudaMalloc(&db, P * 2 * N * 16); cudaMalloc(&dout, 128 * N * 4);
  std::vector<float> A(128 * P * 8), B(N * P * 8), sa(P * 2 * 128 * 4), sb(P * 2 * N * 4), out(128 * N);
  const char* names[] = {"uniform[0.5,1.5] x uniform[0.5,1.5]", "|normal| x |normal|",
                         "lognormal(0,1) x lognormal(0,1)", "relu(normal) x normal (mixed sign)"};
  printf("# mma_bias: M=128 N=%d K=8 chains; compensation tested: %.2e per MMA\n", N, comp);
  for (int dist = 0; dist < 4; ++dist) {
    std::mt19937 rng(7 + dist);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::uniform_real_distribution<float> ud(0.5f, 1.5f);
    auto draw = [&](bool is_b) -> float {
      switch (dist) {
        case 0: return ud(rng);
        case 1: return fabsf(nd(rng)) + 1e-3f;
        case 2: return expf(nd(rng));
        default: return is_b ? nd(rng) : fmaxf(nd(rng), 0.f);
      }
    };
    for (auto& v : A) v = tf32(draw(false));
    for (auto& v : B) v = tf32(draw(true));
    for (int kk = 0; kk < P * 8; ++kk) {
      for (int r = 0; r < 128; ++r) sa[((kk / 4) * 128 + r) * 4 + kk % 4] = A[r * P * 8 + kk];
      for (int n = 0; n < N; ++n) sb[((kk / 4) * N + n) * 4 + kk % 4] = B[n * P * 8 + kk];
    }
    cudaMemcpy(da, sa.data(), sa.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, sb.data(), sb.size() * 4, cudaMemcpyHostToDevice);
    printf("%s\n", names[dist]);
    const int ms[] = {3, 6, 12, 24, 36, 48, 72, 96, 144, 288};
    for (int m : ms) {
      k<<<1, 128, smem>>>(da, db, dout, m);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      double s_rel = 0, s_abs = 0, s_comp = 0, worst = 0, denom = 0;
      for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) {
        double ex = 0, mag = 0;
        for (int j = 0; j < m; ++j) for (int e = 0; e < 8; ++e) {
          const double pr = (double)A[r * P * 8 + (j % P) * 8 + e] * B[n * P * 8 + (j % P) * 8 + e];
          ex += pr; mag += fabs(pr);
        }
        const double got = out[r * N + n];
        s_rel += (got - ex) / mag;                       // signed, relative to sum |products|
        s_comp += (got * (1.0 + comp * m) - ex) / mag;
        s_abs += fabs(got - ex) / mag;
        worst = fmax(worst, fabs(got - ex) / mag);
        denom += 1;
      }
      printf("  m=%3d  mean signed err/sum|p| %+.3e  (per MMA %+.3e)  after comp %+.3e  mean|err| %.3e  max %.3e\n",
             m, s_rel / denom, s_rel / denom / m, s_comp / denom, s_abs / denom, worst);
    }
  }
  return 0;
}

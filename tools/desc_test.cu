// Unit test of the shared-memory matrix descriptor the fused fire kernel relies on (fire_tc.cu):
// A operand K-major, SWIZZLE_NONE, rows of one 8-row core matrix 16 bytes apart, 8-row groups
// SBO bytes apart, 16-byte K chunks LBO bytes apart - so that a 3x3 tap (dy, dx) over a halo tile
// stored [k chunk][h 18][w 10][4 floats] is only a different START ADDRESS (dy*160 + dx*16).
// B operand: the SWIZZLE_64B K-major tile the conv kernels already use.
// Prints the max error of D = A_tap * B^T against the host for every tap, for (LBO, SBO) as assumed
// and swapped.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o desc_test desc_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../squeezedet_b200/csrc/tc_ptx.cuh"

using namespace sqdet;

constexpr int HH = 18, HW = 10, K = 16, N = 64;
constexpr int CH_STRIDE = HH * HW * 16;   // bytes between 16-byte K chunks
constexpr int ROW_STRIDE = HW * 16;       // bytes between halo rows

__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(const float* q, const float* w, float* out, int dy, int dx,
                                           int swapped) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  float* sq = (float*)smem;                         // [K/4][18][10][4]
  float* sw = (float*)(smem + 16384);               // [N][16] floats, SW64
  for (int i = threadIdx.x; i < (K / 4) * HH * HW * 4; i += 128) sq[i] = q[i];
  for (int i = threadIdx.x; i < N * K; i += 128) {
    const int n = i / K, kk = i % K;
    const int chunk = kk / 4, e = kk % 4;
    sw[n * 16 + ((chunk ^ ((n >> 1) & 3)) << 2) + e] = w[i];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 64);
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t qa = smem_u32(sq) + dy * ROW_STRIDE + dx * 16;
    const uint64_t lbo = swapped ? ROW_STRIDE : CH_STRIDE, sbo = swapped ? CH_STRIDE : ROW_STRIDE;
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t da = (uint64_t)(((qa + ks * 2 * CH_STRIDE) & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16) |
                          ((sbo >> 4) << 32) | (1ull << 46);
      const uint64_t db = make_desc<16>(smem_u32(sw)) + (uint64_t)(2 * ks);
      umma_ss(tm, da, db, idesc, ks ? 1u : 0u);
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


// Issue rate of the same MMA (M=128, N, K=8, A from shared memory through the SWIZZLE_NONE
// descriptor) as a function of the A tile's row pitch and of the tap offset: does a core matrix that
// straddles a 128-byte line (row pitch 160 B, dx != 0) cost extra shared-memory cycles?
__global__ void __launch_bounds__(128, 1) k_rate(long long* out, int n, int row_stride, int ch_stride,
                                                  int dy, int dx, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((float*)smem)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 256);
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t qa = smem_u32(smem) + dy * row_stride + dx * 16;
    const uint64_t da = (uint64_t)((qa & 0x3FFFFu) >> 4) | ((uint64_t)(ch_stride >> 4) << 16) |
                        ((uint64_t)(row_stride >> 4) << 32) | (1ull << 46);
    const uint64_t db = make_desc<16>(smem_u32(smem + 24576));
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) umma_ss(tm, da, db, idesc, 1u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[0] = clock64() - t0;
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}


// Cost of the generic->async proxy fence a thread pays before a TMA store / an MMA may read what it
// wrote with st.shared: 8 x st.shared.v4 followed by (mode 0) nothing, (1) fence.proxy.async,
// (2) fence.proxy.async + __syncwarp, per iteration, one warp per SM sub-partition.
__global__ void __launch_bounds__(128, 1) k_fence(long long* out, int mode, int iters) {
  __shared__ __align__(1024) float buf[128 * 32];
  const uint32_t a = smem_u32(buf) + threadIdx.x * 128;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) sts128(a + ((k ^ (threadIdx.x & 7)) << 4), make_float4(i, k, 0.f, 1.f));
    if (mode >= 1) fence_async_proxy();
    if (mode >= 2) __syncwarp();
  }
  if (threadIdx.x == 0) out[0] = clock64() - t0;
}

static float tf32(float x) {
  uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r;
}

int main() {
  std::vector<float> q((K / 4) * HH * HW * 4), w(N * K), pix(HH * HW * K);
  srand(1);
  for (auto& v : pix) v = tf32((float)rand() / RAND_MAX - 0.5f);
  for (auto& v : w) v = tf32((float)rand() / RAND_MAX - 0.5f);
  for (int h = 0; h < HH; ++h) for (int x = 0; x < HW; ++x) for (int c = 0; c < K; ++c)
    q[(((c / 4) * HH + h) * HW + x) * 4 + c % 4] = pix[(h * HW + x) * K + c];
  float *dq, *dw, *dout;
  cudaMalloc(&dq, q.size() * 4); cudaMalloc(&dw, w.size() * 4); cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  std::vector<float> out(128 * N);
  for (int swapped = 0; swapped < 1; ++swapped)
    for (int dy = 0; dy < 3; ++dy) for (int dx = 0; dx < 3; ++dx) {
      cudaMemset(dout, 0, 128 * N * 4);
      k<<<1, 128, 40000>>>(dq, dw, dout, dy, dx, swapped);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("swapped=%d tap(%d,%d): %s\n", swapped, dy, dx, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      double err = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        const int hh = m / 8, ww = m % 8;      // tile 16 h x 8 w, TMEM lane = hh*8 + ww
        double s = 0;
        for (int c = 0; c < K; ++c) s += (double)pix[((hh + dy) * HW + ww + dx) * K + c] * w[n * K + c];
        err = fmax(err, fabs(s - out[m * N + n]));
      }
      printf("desc_test swapped=%d tap(%d,%d) max_err %.3e %s\n", swapped, dy, dx, err, err < 1e-4 ? "OK" : "MISMATCH");
    }

  {
    long long* dt;
    cudaMalloc(&dt, 8);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    const int cfgs[][4] = {{128, 2304, 0, 0}, {128, 2304, 1, 1}, {160, 2880, 0, 0}, {160, 2880, 0, 1},
                           {160, 2880, 1, 0}, {160, 2880, 1, 1}, {160, 2880, 2, 2}, {256, 4608, 0, 0}, {256, 4608, 1, 1}};
    for (int n : {16, 32, 64, 128})
      for (auto& c : cfgs) {
        long long h = 0;
        k_rate<<<1, 128, 60000>>>(dt, n, c[0], c[1], c[2], c[3], 2000);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("rate launch failed\n"); return 1; }
        cudaMemcpy(&h, dt, 8, cudaMemcpyDeviceToHost);
        printf("mma_rate_ss N=%3d row_pitch %3d B chunk_pitch %4d B tap(%d,%d): %.1f clk/MMA\n", n, c[0], c[1],
               c[2], c[3], (double)h / 2000);
      }
  }

  {
    long long* dt;
    cudaMalloc(&dt, 8);
    for (int mode = 0; mode < 3; ++mode) {
      long long h = 0;
      k_fence<<<1, 128>>>(dt, mode, 1000);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, dt, 8, cudaMemcpyDeviceToHost);
      printf("fence_cost mode %d (0 = 8 x st.shared.v4, 1 = + fence.proxy.async, 2 = + __syncwarp): %.1f clk/iteration\n",
             mode, (double)h / 1000);
    }
  }
  return 0;
}

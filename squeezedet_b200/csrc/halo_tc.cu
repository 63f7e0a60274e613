// Halo-tile 3x3 convolution on tcgen05 (SQDET_MATH_TF32X3_TC), stride 1, SAME, Cin % 16 == 0.
//
// Replaces tf.nn.conv2d + bias_add [+ batch_normalization] + relu of the reference
// (src/nn_skeleton.py:374-586) for the ConvDet head (src/nets/squeezeDet.py:73-78: 768 -> 72
// channels on the 24 x 78 grid) and other 3x3 layers whose operand staging, not the tensor pipe,
// bounded conv_tc.cu: there every filter tap re-fetched its shifted input tile through TMA and
// re-split it into tensor memory (ConvDet: 2.39 GB of TMA loads for 128 MB of algorithmic bytes, the
// splitter at ~650 clocks per 480-clock MMA stage, split-K over filter rows re-reading the input
// three times from DRAM).
//
// Here (same operand trick as fire_tc.cu, checked by tools/desc_test.cu):
//   item     = one 128-pixel output tile (16h x 8w, or 8h x 16w when that wastes fewer rows) x one
//              chunk of <= 128 output channels x one range of input-channel chunks (split-K over
//              CHANNELS: every byte of the input is read once);
//   Q chunk  = the 18 x 10 halo of the tile for 16 input channels: ONE TMA box -> 4 splitter warps
//              split it hi/lo ONCE into shared memory as [16-byte channel chunk][outer 18][inner 10]
//              - a K-major SWIZZLE_NONE operand whose tap (dy, dx) is a start-address offset;
//   MMAs     = 9 taps x 2 K steps x 3 (a_lo*b_hi, a_hi*b_lo, a_hi*b_hi) from shared-memory
//              descriptors (.ss form); weights stream through a ring of [N][16] hi/lo tiles;
//   segments = 36 chained MMAs, summed in fp32 registers by two drain warpgroups (conv_tc.cu 4.1);
//   epilogue = + bias [*scale + shift], ReLU, TMA store; split-K partials go to an L2-resident scratch
//              tensor and a deterministic reduction kernel finishes them.
// Roles (512 threads): warpgroups 0,1 drains (32-channel groups jg = g mod 2 of every segment),
//   warpgroup 2 splitter, warp 12 TMA producer (raw halo ring + weight ring, non-blocking), warp 13
//   TMEM owner + MMA issuer.
// Roofline: ConvDet is tensor-bound (AI 290 FLOP/B); the MMA floor is 46 clocks per N=80 MMA
//   (tools/desc_test.cu), the weight stream (1.3 GB per launch from L2) sits next to it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "conv_tc.cuh"
#include "halo_tc.cuh"
#include "tc_ptx.cuh"

namespace sqdet {
namespace {

constexpr int HC_THREADS = 512;
constexpr int HC_KC = 16;                          // input channels per Q chunk / weight tile
constexpr int HC_QROW = 10 * 16;                   // bytes between outer rows of one channel chunk
constexpr int HC_QCH = 18 * 10 * 16;               // bytes between 16-byte channel chunks (2880)
constexpr int HC_ROWS = 180;                       // halo pixels
constexpr int HC_QBYTES = 24576;                   // one Q chunk: [hi 4 x 2880 | lo 4 x 2880], padded
constexpr int HC_QHALF = 4 * HC_QCH;
constexpr int HC_RAW = 12288;                      // raw halo box 180 x 64 B, padded
constexpr int HC_SEG = 6;                          // weight tiles (6 MMAs each) per accumulation segment
constexpr int HC_MAX_RING = 8;
constexpr int HC_MAX_C = 512;

struct HaloParams {
  CUtensorMap tmX;     // input [B,H,W,Cin], box {16 ch, 10|18 w, 18|10 h, 1}, SWIZZLE_64B
  CUtensorMap tmW;     // weight tiles, rows [tile][Ne][16], box {16, Ne}
  CUtensorMap tmY;     // output (or the split-K scratch), box {32 ch, 8 w, 4 h} or {32 ch, 1 w, 8 h}
  const float* bias;   // null for split-K partials
  const float* scale;
  const float* shift;
  int B, H, W, Cin, Cout;
  int orient;          // 0: tile 16h x 8w (outer = h, inner = w); 1: tile 8h x 16w (outer = w, inner = h)
  int tiles_h, tiles_w, ntiles;
  int Ne, nchunks_n, kchunks, ksplit, kper, nitems;
  int nq, nr, nw, store_ring, tmem_cols;
  int relu, y_coff, part_pitch;
  int lo_rows;         // rows between the hi and the lo copy of the packed weights
  float bias_comp;
  int off_q, off_raw, off_w, w_tile, off_out, off_par, off_bar;
  long long* dbg;
};

#define HC_WAIT(counter, bar, parity)                       \
  do {                                                      \
    if (p.dbg) {                                            \
      const uint32_t _t0 = (uint32_t)clock();               \
      mbar_wait(bar, parity);                               \
      counter += (uint32_t)clock() - _t0;                   \
    } else {                                                \
      mbar_wait(bar, parity);                               \
    }                                                       \
  } while (0)

__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_store_4d_nc(uint32_t src, const CUtensorMap* map, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

struct Ring {
  int s;
  uint32_t ph;
  __device__ __forceinline__ void next(int n) {
    if (++s == n) { s = 0; ph ^= 1u; }
  }
};

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HC_THREADS, 1)
halo_conv_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~uintptr_t(1023));
  const uint32_t smem_b = smem_u32(smem);
  const uint32_t q_b = smem_b + (uint32_t)p.off_q;
  const uint32_t raw_b = smem_b + (uint32_t)p.off_raw;
  const uint32_t w_b = smem_b + (uint32_t)p.off_w;
  const uint32_t out_b = smem_b + (uint32_t)p.off_out;
  const uint32_t par_b = smem_b + (uint32_t)p.off_par;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* rfull = bars;                        // [nr] TMA -> splitter
  uint64_t* rempty = bars + HC_MAX_RING;         // [nr] splitter -> TMA
  uint64_t* qfull = bars + 2 * HC_MAX_RING;      // [nq] splitter -> MMA
  uint64_t* qempty = bars + 3 * HC_MAX_RING;     // [nq] MMA commit -> splitter
  uint64_t* wfull = bars + 4 * HC_MAX_RING;      // [nw] TMA -> MMA
  uint64_t* wempty = bars + 5 * HC_MAX_RING;     // [nw] MMA commit -> TMA
  uint64_t* tfull = bars + 6 * HC_MAX_RING;      // [2]  MMA commit -> drains
  uint64_t* tempty = tfull + 2;                  // [2]  drains -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 4);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ne = p.Ne;
  const int my_items = ((int)p.nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < HC_MAX_RING; ++s) {
      mbar_init(&rfull[s], 1);
      mbar_init(&rempty[s], 128);
      mbar_init(&qfull[s], 128);
      mbar_init(&qempty[s], 1);
      mbar_init(&wfull[s], 1);
      mbar_init(&wempty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 13) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // item -> (tile, output-channel chunk, input-channel range); tile -> (image, th, tw)
#define HC_ITEM_DECODE(k_)                                              \
  int it_ = (int)blockIdx.x + (k_) * (int)gridDim.x;                    \
  const int ks = it_ % p.ksplit;                                        \
  it_ /= p.ksplit;                                                      \
  const int nc = it_ % p.nchunks_n;                                     \
  it_ /= p.nchunks_n;                                                   \
  const int tw = it_ % p.tiles_w;                                       \
  it_ /= p.tiles_w;                                                     \
  const int th = it_ % p.tiles_h;                                       \
  const int img = it_ / p.tiles_h;                                      \
  const int h0 = th * (p.orient ? 8 : 16), w0 = tw * (p.orient ? 16 : 8); \
  const int c0 = ks * p.kper, c1 = (c0 + p.kper < p.kchunks) ? c0 + p.kper : p.kchunks;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (warp == 12) {
      // ================================ TMA producer ==========================================
      // raw halo ring and weight ring, each FIFO in consumption order; never blocks on either
      if (lane == 0) {
        Ring rr{0, 0u}, rw{0, 0u};
        const uint32_t half = (uint32_t)(Ne * HC_KC * 4);
        int xk = 0, xc = 0, xc1 = 0, x_h = 0, x_w = 0, x_img = 0;   // next raw chunk
        int wk = 0, wt = 0, wt1 = 0;                                // next weight tile (global index)
        bool x_ready = false, w_ready = false;
        bool x_left = my_items > 0, w_left = my_items > 0;
        while (x_left || w_left) {
          if (x_left) {
            if (!x_ready) {
              HC_ITEM_DECODE(xk);
              (void)nc;
              xc = c0; xc1 = c1; x_h = h0 - 1; x_w = w0 - 1; x_img = img;
              x_ready = true;
            }
            if (mbar_test(&rempty[rr.s], rr.ph ^ 1u)) {
              mbar_expect_tx(&rfull[rr.s], (uint32_t)(HC_ROWS * 64));
              tma_load_4d(smem + p.off_raw + (size_t)rr.s * HC_RAW, &p.tmX, &rfull[rr.s], xc * HC_KC,
                          x_w, x_h, x_img);
              rr.next(p.nr);
              if (++xc == xc1) {
                x_ready = false;
                if (++xk == my_items) x_left = false;
              }
            }
          }
          if (w_left) {
            if (!w_ready) {
              HC_ITEM_DECODE(wk);
              (void)h0; (void)w0; (void)img;
              wt = (nc * p.kchunks + c0) * 9;
              wt1 = (nc * p.kchunks + c1) * 9;
              w_ready = true;
            }
            if (mbar_test(&wempty[rw.s], rw.ph ^ 1u)) {
              uint8_t* dst = smem + p.off_w + (size_t)rw.s * p.w_tile;
              mbar_expect_tx(&wfull[rw.s], 2u * half);
              tma_load_2d(dst, &p.tmW, &wfull[rw.s], 0, wt * Ne);
              tma_load_2d(dst + half, &p.tmW, &wfull[rw.s], 0, wt * Ne + p.lo_rows);
              rw.next(p.nw);
              if (++wt == wt1) {
                w_ready = false;
                if (++wk == my_items) w_left = false;
              }
            }
          }
        }
      }
    } else if (warp == 13) {
      // ================================ MMA issuer ===========================================
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Ne >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t dw_hi = make_desc<HC_KC>(0) & 0xFFFFFFFF00000000ull;
      const uint32_t dw_lo = (uint32_t)(make_desc<HC_KC>(0) & 0xFFFFFFFFull);
      // A: the Q chunk, SWIZZLE_NONE, LBO = channel-chunk stride, SBO = outer-row stride
      const uint64_t dq_hi = ((uint64_t)(HC_QROW >> 4) << 32) | (1ull << 46);
      const uint32_t dq_lo = (uint32_t)(HC_QCH >> 4) << 16;
      // 16-byte units: step of the tap offset along the filter's dx and dy
      const uint32_t step_dx = p.orient ? (uint32_t)(HC_QROW >> 4) : 1u;
      const uint32_t step_dy = p.orient ? 1u : (uint32_t)(HC_QROW >> 4);
      Ring rq{0, 0u}, rw{0, 0u};
      uint32_t eg = 0u;
      uint32_t w_qfull = 0, w_wfull = 0, w_tempty = 0;
      const uint32_t t_begin = (uint32_t)clock();
      for (int k = 0; k < my_items; ++k) {
        HC_ITEM_DECODE(k);
        (void)nc; (void)h0; (void)w0; (void)img;
        const int nt = (c1 - c0) * 9;
        int tcount = 0;
        uint32_t d_tmem = 0u, buf = 0u;
        for (int kc = c0; kc < c1; ++kc) {
          HC_WAIT(w_qfull, &qfull[rq.s], rq.ph);
          tc_fence_after();
          const uint32_t qa = ((q_b + (uint32_t)(rq.s * HC_QBYTES)) & 0x3FFFFu) >> 4;
          uint32_t tap_off = 0u;
          int dx = 0;
          for (int tap = 0; tap < 9; ++tap) {
            const bool seg_start = (tcount % HC_SEG) == 0;
            if (seg_start) {
              buf = eg & 1u;
              HC_WAIT(w_tempty, &tempty[buf], ((eg >> 1) & 1u) ^ 1u);
              tc_fence_after();
              d_tmem = tmem_u + buf * (uint32_t)Ne;
            }
            HC_WAIT(w_wfull, &wfull[rw.s], rw.ph);
            tc_fence_after();
            const uint32_t wb = w_b + (uint32_t)(rw.s * p.w_tile);
            const uint32_t a_hi = dq_lo | (qa + tap_off);
            const uint32_t a_lo = a_hi + (uint32_t)(HC_QHALF >> 4);
            const uint32_t b_hi = dw_lo | ((wb & 0x3FFFFu) >> 4);
            const uint32_t b_lo = b_hi + (uint32_t)((Ne * HC_KC * 4) >> 4);
            ++tcount;
            const bool seg_end = (tcount % HC_SEG) == 0 || tcount == nt;
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint32_t ao = (uint32_t)(j * ((2 * HC_QCH) >> 4)), bo = (uint32_t)(2 * j);
                umma_tf32_ss(d_tmem, dq_hi | (uint64_t)(a_lo + ao), dw_hi | (uint64_t)(b_hi + bo), idesc,
                             (!seg_start || j != 0) ? 1u : 0u);
                umma_tf32_ss(d_tmem, dq_hi | (uint64_t)(a_hi + ao), dw_hi | (uint64_t)(b_lo + bo), idesc, 1u);
                umma_tf32_ss(d_tmem, dq_hi | (uint64_t)(a_hi + ao), dw_hi | (uint64_t)(b_hi + bo), idesc, 1u);
              }
              umma_commit(&wempty[rw.s]);
              if (seg_end) umma_commit(&tfull[buf]);
              if (tap == 8) umma_commit(&qempty[rq.s]);
            }
            __syncwarp();
            rw.next(p.nw);
            if (seg_end) ++eg;
            if (++dx == 3) { dx = 0; tap_off += step_dy - 2u * step_dx; }
            else tap_off += step_dx;
          }
          rq.next(p.nq);
        }
      }
      if (p.dbg && lane == 0) {
        p.dbg[blockIdx.x * 16 + 0] = (uint32_t)clock() - t_begin;
        p.dbg[blockIdx.x * 16 + 1] = w_qfull;
        p.dbg[blockIdx.x * 16 + 2] = w_wfull;
        p.dbg[blockIdx.x * 16 + 3] = w_tempty;
      }
    } else if (warp < 12) {
      // ================================ splitter ==============================================
      // raw halo rows (64 B = 16 channels, SWIZZLE_64B) -> q = q_hi + q_lo -> Q chunk
      const int t = threadIdx.x - 256;             // 0..127
      uint32_t qoff[2], roff[2], rsw[2];
      bool rok[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = t + 128 * i;
        rok[i] = r < HC_ROWS;
        const int rr = rok[i] ? r : 0;
        const int outer = p.orient ? rr % 18 : rr / 10;
        const int inner = p.orient ? rr / 18 : rr % 10;
        qoff[i] = (uint32_t)(outer * HC_QROW + inner * 16);
        roff[i] = (uint32_t)(rr * 64);
        rsw[i] = (uint32_t)((rr >> 1) & 3);
      }
      Ring rr{0, 0u}, rq{0, 0u};
      uint32_t w_rfull = 0, w_qempty = 0;
      for (int k = 0; k < my_items; ++k) {
        HC_ITEM_DECODE(k);
        (void)nc; (void)h0; (void)w0; (void)img;
        for (int kc = c0; kc < c1; ++kc) {
          HC_WAIT(w_rfull, &rfull[rr.s], rr.ph);
          HC_WAIT(w_qempty, &qempty[rq.s], rq.ph ^ 1u);
          const uint32_t raw = raw_b + (uint32_t)(rr.s * HC_RAW);
          const uint32_t q = q_b + (uint32_t)(rq.s * HC_QBYTES);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (rok[i]) {
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                const float4 v = lds128(raw + roff[i] + (((uint32_t)c4 ^ rsw[i]) << 4));
                float4 hi, lo;
                hi.x = rn_tf32(v.x); hi.y = rn_tf32(v.y); hi.z = rn_tf32(v.z); hi.w = rn_tf32(v.w);
                lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
                sts128(q + qoff[i] + (uint32_t)(c4 * HC_QCH), hi);
                sts128(q + qoff[i] + (uint32_t)(HC_QHALF + c4 * HC_QCH), lo);
              }
            }
          }
          fence_async_proxy();                     // the tensor core reads Q through the async proxy
          mbar_arrive(&qfull[rq.s]);
          mbar_arrive(&rempty[rr.s]);
          rr.next(p.nr);
          rq.next(p.nq);
        }
      }
      if (p.dbg && t == 0) {
        p.dbg[blockIdx.x * 16 + 4] = w_rfull;
        p.dbg[blockIdx.x * 16 + 5] = w_qempty;
      }
    }
  } else {
    // ================================ drains + epilogue ======================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    const int g = warp >> 2;
    const int q = warp & 3;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    // epilogue parameters of the whole conv: bias | scale | shift
    for (int i = threadIdx.x; i < 3 * HC_MAX_C; i += 256) {
      const int c = i % HC_MAX_C, which = i / HC_MAX_C;
      float v = which == 1 ? 1.f : 0.f;
      if (c < p.Cout) {
        if (which == 0 && p.bias) v = __ldg(p.bias + c);
        if (which == 1 && p.scale) v = __ldg(p.scale + c);
        if (which == 2 && p.scale) v = __ldg(p.shift + c);
      }
      sts32(par_b + 4u * (uint32_t)i, v);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const bool affine = p.scale != nullptr;
    const float lo_clip = p.relu ? 0.f : -CUDART_INF_F;
    uint32_t eg = 0u;
    int n_store = 0;
    uint32_t w_tfull = 0, c_epi = 0;
    for (int k = 0; k < my_items; ++k) {
      HC_ITEM_DECODE(k);
      const int nt = (c1 - c0) * 9;
      float acc[64];                               // 32-channel groups jg = g and g + 2 of the chunk
      for (int t0 = 0; t0 < nt; t0 += HC_SEG, ++eg) {
        const uint32_t buf = eg & 1u;
        HC_WAIT(w_tfull, &tfull[buf], (eg >> 1) & 1u);
        tc_fence_after();
        const int ntl = (nt - t0) < HC_SEG ? (nt - t0) : HC_SEG;
        const float gain = 1.f + p.bias_comp * (float)(6 * ntl);
        const uint32_t trow = tmem_base + lane_sel + buf * (uint32_t)Ne;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int jg = g + 2 * jj;
          if (jg * 32 < Ne) {                      // warp-uniform
            uint32_t v0[16], v1[16];
            tmem_ld16_nowait(trow + (uint32_t)(jg * 32), v0);
            if (jg * 32 + 16 < Ne) tmem_ld16_nowait(trow + (uint32_t)(jg * 32 + 16), v1);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              acc[jj * 32 + e] = fmaf(__uint_as_float(v0[e]), gain, t0 == 0 ? 0.f : acc[jj * 32 + e]);
              acc[jj * 32 + 16 + e] =
                  (jg * 32 + 16 < Ne)
                      ? fmaf(__uint_as_float(v1[e]), gain, t0 == 0 ? 0.f : acc[jj * 32 + 16 + e])
                      : 0.f;
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty[buf]);
      }
      // ---- epilogue: one 32-pixel x 32-channel tile per warp and channel group -> TMA store -------
      const uint32_t t0c = p.dbg ? (uint32_t)clock() : 0u;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int jg = g + 2 * jj;
        if (jg * 32 < Ne && nc * Ne + jg * 32 < p.Cout) {
          if (lane == 0) {
            if (p.store_ring == 2) tma_store_wait_read_le1();
            else tma_store_wait_read_all();
          }
          __syncwarp();
          const uint32_t tile_w = out_b + (uint32_t)((g * 4 + q) * (4096 * p.store_ring) +
                                                     (p.store_ring == 2 ? (n_store & 1) * 4096 : 0));
          const int cbase = nc * Ne + jg * 32;     // channel of this group within the conv
#pragma unroll
          for (int kq = 0; kq < 8; ++kq) {
            float o[4] = {acc[jj * 32 + kq * 4], acc[jj * 32 + kq * 4 + 1], acc[jj * 32 + kq * 4 + 2],
                          acc[jj * 32 + kq * 4 + 3]};
            if (p.part_pitch == 0) {
              const float4 b = lds128(par_b + 4u * (uint32_t)(cbase + kq * 4));
              o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
              if (affine) {
                const float4 s0 = lds128(par_b + 4u * (uint32_t)(HC_MAX_C + cbase + kq * 4));
                const float4 h0v = lds128(par_b + 4u * (uint32_t)(2 * HC_MAX_C + cbase + kq * 4));
                o[0] = o[0] * s0.x + h0v.x; o[1] = o[1] * s0.y + h0v.y;
                o[2] = o[2] * s0.z + h0v.z; o[3] = o[3] * s0.w + h0v.w;
              }
              o[0] = fmaxf(o[0], lo_clip); o[1] = fmaxf(o[1], lo_clip);
              o[2] = fmaxf(o[2], lo_clip); o[3] = fmaxf(o[3], lo_clip);
            }
            sts128(tile_w + (uint32_t)(lane * 128 + ((kq ^ (lane & 7)) << 4)),
                   make_float4(o[0], o[1], o[2], o[3]));
          }
          fence_async_proxy();
          __syncwarp();
          if (lane == 0) {
            const int cdst = (p.part_pitch ? ks * p.part_pitch : p.y_coff) + cbase;
            if (p.orient == 0) {
              // TMEM lane 32q + l = pixel (h0 + 4q + l / 8, w0 + l % 8) = row l of a {32, 8 w, 4 h} box
              tma_store_4d_nc(tile_w, &p.tmY, cdst, w0, h0 + 4 * q, img);
            } else {
              // TMEM lane 32q + l = pixel (h0 + l % 8, w0 + 4q + l / 8): four {32, 1 w, 8 h} boxes
#pragma unroll
              for (int wl = 0; wl < 4; ++wl)
                tma_store_4d_nc(tile_w + (uint32_t)(wl * 1024), &p.tmY, cdst, w0 + 4 * q + wl, h0, img);
            }
            tma_store_commit();
          }
          ++n_store;
        }
      }
      if (p.dbg) c_epi += (uint32_t)clock() - t0c;
    }
    if (lane == 0) tma_store_wait_all();
    if (p.dbg && g == 0 && (threadIdx.x & 127) == 0) {
      p.dbg[blockIdx.x * 16 + 6] = w_tfull;
      p.dbg[blockIdx.x * 16 + 7] = c_epi;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
struct HaloImpl {
  HaloParams prm;
  size_t smem_bytes = 0;
  dim3 grid;
  float* d_w = nullptr;        // packed weight tiles [N chunk][K chunk][tap][Ne][16], hi rows then lo rows
  float* d_bias = nullptr;
  float* d_scale = nullptr;
  float* d_shift = nullptr;
  float* d_scratch = nullptr;  // split-K partials [pixels][ksplit * part_pitch]
  float* y_final = nullptr;
  int y_cstride = 0, y_coff = 0, relu = 0;
  long long npix = 0;
};

static inline float hc_rn_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

static void release_halo(void** impl) {
  if (!*impl) return;
  HaloImpl* im = static_cast<HaloImpl*>(*impl);
  cudaFree(im->d_w);
  cudaFree(im->d_bias);
  cudaFree(im->d_scale);
  cudaFree(im->d_shift);
  cudaFree(im->d_scratch);
  delete im;
  *impl = nullptr;
}

static int hc_env(const char* name, int dflt) {
  const char* a = getenv(name);
  return a ? atoi(a) : dflt;
}

// x box for the two tile orientations: SWIZZLE_64B rows of 16 channels
static int encode_halo_x(CUtensorMap* map, const float* x, int B, int H, int W, int C, int orient) {
  return tc_encode_act_map(map, x, B, H, W, C, HC_KC, orient ? 18 : 10, orient ? 10 : 18);
}

}  // namespace

int halo_conv_plan(HaloConvPlan* plan, int B, int H, int W, int Cin, int Cout, int relu,
                   bool has_affine, int y_cstride, int y_coff, const float* x_dev, float* y_dev) {
  plan->enabled = false;
  plan->impl = nullptr;
  if (Cin % HC_KC != 0 || Cin < HC_KC || Cout < 8 || Cout > HC_MAX_C || (Cout % 4) || (y_cstride % 4) ||
      (y_coff % 4))
    return 0;
  // output-channel chunks: uniform N <= 128; a chunk must be whole 32-channel groups unless it
  // ends the output tensor (TMA clips the tail there)
  const int nsplit = (Cout + 127) / 128;
  const int Ne = ((Cout + nsplit - 1) / nsplit + 15) / 16 * 16;
  if (nsplit > 1 && (Cout % nsplit != 0 || (Cout / nsplit) % 32 != 0)) return 0;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  HaloImpl* im = new HaloImpl();
  HaloParams& P = im->prm;
  memset(&P, 0, sizeof P);
  P.B = B; P.H = H; P.W = W; P.Cin = Cin; P.Cout = Cout; P.relu = relu;
  // tile orientation: the one that covers the image with fewer 128-pixel tiles
  const long long t0 = (long long)((H + 15) / 16) * ((W + 7) / 8), t1 = (long long)((H + 7) / 8) * ((W + 15) / 16);
  P.orient = hc_env("SQDET_HALO_ORIENT", t1 < t0 ? 1 : 0);
  P.tiles_h = P.orient ? (H + 7) / 8 : (H + 15) / 16;
  P.tiles_w = P.orient ? (W + 15) / 16 : (W + 7) / 8;
  P.ntiles = B * P.tiles_h * P.tiles_w;
  P.Ne = Ne;
  P.nchunks_n = nsplit;
  P.kchunks = Cin / HC_KC;
  // split-K over input-channel ranges when the persistent grid would run few, long rounds: cost model
  // rounds(items) x (chunks per item + ~1.5 chunks of pipeline fill / epilogue)
  int ksplit = 1;
  {
    const long long base = (long long)P.ntiles * nsplit;
    const int force = hc_env("SQDET_HALO_KSPLIT", 0);
    if (force > 0) {
      ksplit = force;
    } else if (base < 4LL * sms && nsplit == 1) {
      double best = 1e30;
      for (int k = 1; k <= 8; ++k) {
        const int kper = (P.kchunks + k - 1) / k;
        if (kper < 4 && k > 1) break;
        const double rounds = (double)((base * k + sms - 1) / sms);
        const double cost = rounds * (kper + 1.5);
        if (cost < best * 0.97) { best = cost; ksplit = k; }
      }
    }
    if (ksplit > P.kchunks) ksplit = P.kchunks;
  }
  P.ksplit = ksplit;
  P.kper = (P.kchunks + ksplit - 1) / ksplit;
  P.ksplit = (P.kchunks + P.kper - 1) / P.kper;        // no empty ranges
  ksplit = P.ksplit;
  P.nitems = P.ntiles * nsplit * ksplit;
  if (ksplit == 1) {
    // direct epilogue: the last (partial) 32-channel group must end the tensor
    if ((Cout % 32) != 0 && (y_coff + Cout != y_cstride)) { delete im; return 0; }
  }
  P.w_tile = 2 * Ne * HC_KC * 4;
  P.bias_comp = 1.4e-8f;
  {
    const char* a = getenv("SQDET_TC_BIAS_COMP");
    if (a) P.bias_comp = (float)atof(a);
  }
  // shared-memory plan
  P.nq = hc_env("SQDET_HALO_NQ", 3);
  P.nr = hc_env("SQDET_HALO_NR", 2);
  P.store_ring = 1;
  {
    const long long fixed = (long long)P.nq * HC_QBYTES + (long long)P.nr * HC_RAW + 8LL * 4096 * P.store_ring +
                            3LL * HC_MAX_C * 4 + 1024 + 2048;
    int nw = (int)((232448 - fixed) / P.w_tile);
    if (nw > HC_MAX_RING) nw = HC_MAX_RING;
    const int force = hc_env("SQDET_HALO_NW", 0);
    if (force > 0 && force < nw) nw = force;
    if (nw < 3) { delete im; return 0; }
    P.nw = nw;
  }
  {
    auto up = [](int v) { return (v + 1023) & ~1023; };
    int off = 0;
    P.off_q = off;   off = up(off + P.nq * HC_QBYTES);
    P.off_raw = off; off = up(off + P.nr * HC_RAW);
    P.off_w = off;   off = up(off + P.nw * P.w_tile);
    P.off_out = off; off = up(off + 8 * 4096 * P.store_ring);
    P.off_par = off; off += 3 * HC_MAX_C * 4;
    P.off_bar = off; off += 1024;
    im->smem_bytes = (size_t)off + 1024;
    if (im->smem_bytes > 232448) { delete im; return 0; }
  }
  {
    int cols = 32;
    while (cols < 2 * Ne) cols <<= 1;
    P.tmem_cols = cols;
  }
  im->grid = dim3((unsigned)(P.nitems < sms ? P.nitems : sms));
  im->y_final = y_dev;
  im->y_cstride = y_cstride;
  im->y_coff = y_coff;
  im->relu = relu;
  im->npix = (long long)B * H * W;
  void* pim = im;
  const size_t tiles = (size_t)nsplit * P.kchunks * 9;
  P.lo_rows = (int)(tiles * Ne);
  const size_t w_floats = tiles * Ne * HC_KC * 2;
  bool ok = cudaMalloc(&im->d_w, sizeof(float) * w_floats) == cudaSuccess &&
            cudaMalloc(&im->d_bias, sizeof(float) * Cout) == cudaSuccess;
  if (ok && has_affine)
    ok = cudaMalloc(&im->d_scale, sizeof(float) * Cout) == cudaSuccess &&
         cudaMalloc(&im->d_shift, sizeof(float) * Cout) == cudaSuccess;
  float* y_target = y_dev;
  int y_channels = y_cstride;
  if (ok && ksplit > 1) {
    P.part_pitch = (Ne + 31) / 32 * 32;
    ok = cudaMalloc(&im->d_scratch, sizeof(float) * (size_t)im->npix * ksplit * P.part_pitch) == cudaSuccess;
    y_target = im->d_scratch;
    y_channels = ksplit * P.part_pitch;
  }
  if (!ok) {
    release_halo(&pim);
    return fail(SQDET_ERR_CUDA, "halo_conv_plan: cudaMalloc failed");
  }
  cudaMemset(im->d_w, 0, sizeof(float) * w_floats);
  cudaMemset(im->d_bias, 0, sizeof(float) * Cout);
  if (ksplit == 1) {
    P.bias = im->d_bias;
    P.scale = im->d_scale;
    P.shift = im->d_shift;
    P.y_coff = y_coff;
  }
  int rc = encode_halo_x(&P.tmX, x_dev, B, H, W, Cin, P.orient);
  if (!rc) rc = tc_encode_w_map(&P.tmW, im->d_w, 2 * P.lo_rows, HC_KC, Ne);
  if (!rc)
    rc = tc_encode_act_map(&P.tmY, y_target, B, H, W, y_channels, 32, P.orient ? 1 : 8, P.orient ? 8 : 4);
  if (rc) {
    release_halo(&pim);
    return rc;
  }
  cudaError_t ce = cudaFuncSetAttribute(halo_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        232448);
  if (ce != cudaSuccess) {
    release_halo(&pim);
    return cuda_fail(ce, "cudaFuncSetAttribute(halo_conv_kernel)");
  }
  plan->enabled = true;
  plan->B = B; plan->H = H; plan->W = W; plan->Cin = Cin; plan->Cout = Cout; plan->relu = relu;
  plan->launches = ksplit > 1 ? 2 : 1;
  plan->impl = im;
  return 1;
}

int halo_conv_pack_weights(HaloConvPlan* plan, const float* w_hwio, const float* bias) {
  HaloImpl* im = static_cast<HaloImpl*>(plan->impl);
  const HaloParams& P = im->prm;
  const int Cin = P.Cin, Cout = P.Cout, Ne = P.Ne;
  std::vector<float> packed((size_t)2 * P.lo_rows * HC_KC, 0.f);
  for (int nc = 0; nc < P.nchunks_n; ++nc)
    for (int kc = 0; kc < P.kchunks; ++kc)
      for (int tap = 0; tap < 9; ++tap)
        for (int n = 0; n < Ne; ++n) {
          const int co = nc * Ne + n;
          if (co >= Cout) continue;
          const size_t row = (((size_t)nc * P.kchunks + kc) * 9 + tap) * Ne + n;
          for (int k = 0; k < HC_KC; ++k) {
            const float v = w_hwio[((size_t)tap * Cin + (size_t)kc * HC_KC + k) * Cout + co];
            const float hi = hc_rn_tf32(v);
            packed[row * HC_KC + k] = hi;
            packed[((size_t)P.lo_rows + row) * HC_KC + k] = hc_rn_tf32(v - hi);
          }
        }
  SQ_CUDA(cudaMemcpy(im->d_w, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (bias) SQ_CUDA(cudaMemcpy(im->d_bias, bias, sizeof(float) * Cout, cudaMemcpyHostToDevice));
  return SQDET_OK;
}

int halo_conv_set_affine(HaloConvPlan* plan, const float* scale, const float* shift) {
  HaloImpl* im = static_cast<HaloImpl*>(plan->impl);
  if (!im->d_scale) return fail(SQDET_ERR_STATE, "halo conv planned without an affine epilogue");
  SQ_CUDA(cudaMemcpy(im->d_scale, scale, sizeof(float) * plan->Cout, cudaMemcpyHostToDevice));
  SQ_CUDA(cudaMemcpy(im->d_shift, shift, sizeof(float) * plan->Cout, cudaMemcpyHostToDevice));
  return SQDET_OK;
}

int launch_halo_conv(const HaloConvPlan& plan, cudaStream_t stream) {
  const HaloImpl* im = static_cast<const HaloImpl*>(plan.impl);
  if (!im) return fail(SQDET_ERR_STATE, "no halo conv plan");
  HaloParams prm = im->prm;
  static int debug = -1;
  if (debug < 0) debug = hc_env("SQDET_TC_DEBUG", 0);
  long long* dbg = nullptr;
  const int nb = (int)im->grid.x;
  if (debug) {
    SQ_CUDA(cudaMalloc(&dbg, sizeof(long long) * 16 * nb));
    SQ_CUDA(cudaMemsetAsync(dbg, 0, sizeof(long long) * 16 * nb, stream));
    prm.dbg = dbg;
  }
  SQ_CUDA(launch_kernel(halo_conv_kernel, im->grid, dim3(HC_THREADS), im->smem_bytes, stream, prm));
  if (prm.ksplit > 1) {
    int rc = launch_splitk_reduce(im->d_scratch, im->y_final, im->d_bias, im->d_scale, im->d_shift, im->npix,
                                  prm.Cout, prm.part_pitch, prm.ksplit, im->y_cstride, im->y_coff, im->relu,
                                  stream);
    if (rc) return rc;
  }
  if (debug) {
    std::vector<long long> h((size_t)16 * nb);
    SQ_CUDA(cudaStreamSynchronize(stream));
    SQ_CUDA(cudaMemcpy(h.data(), dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    cudaFree(dbg);
    double a[16] = {0};
    for (int b = 0; b < nb; ++b)
      for (int k = 0; k < 16; ++k) a[k] += (double)h[(size_t)b * 16 + k] / nb;
    fprintf(stderr,
            "[halo_tc] grid %d items %d (tiles %d %s, N chunks %d x %d, k-split %d x %d chunks) | nq %d nr %d "
            "nw %d smem %zu | per-CTA avg cycles: mma total %.0f waits: qfull %.0f wfull %.0f tempty %.0f | "
            "splitter waits: raw %.0f qempty %.0f | drain(g0): wait-tfull %.0f epilogue %.0f\n",
            nb, prm.nitems, prm.ntiles, prm.orient ? "8h x 16w" : "16h x 8w", prm.nchunks_n, prm.Ne, prm.ksplit,
            prm.kper, prm.nq, prm.nr, prm.nw, im->smem_bytes, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
  }
  return SQDET_OK;
}

void halo_conv_release(HaloConvPlan* plan) {
  release_halo(&plan->impl);
  plan->enabled = false;
}

}  // namespace sqdet

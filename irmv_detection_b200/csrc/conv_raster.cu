// Raster (halo-tile) implicit-GEMM convolution on tcgen05 + TMEM, sm_100a only.
// Covers the 3x3 / stride-1 / pad-1 and the 1x1 convolutions of YOLOv8n (about 85 % of its FLOPs),
// i.e. the bulk of what the reference runs inside its TensorRT engine
// (reference src/yolo_engine.cpp:100-105).
//
// Activations live in the zero-padded raster layout (common.cuh): pixel q of the flat raster has
// its 3x3 neighbours at q + (ky-1)*Wp + (kx-1), with the padding pixels really being zero.  A tile
// of TM = 128*R consecutive raster pixels therefore needs ONE contiguous range of input pixels
// (TM + 2*Wp + 2 of them), and all nine taps are the same shared-memory tile read at nine
// different row offsets:
//
//   warp 8 (one lane)  TMA producer: the activation planes are [pixel][8 channels] arrays, so the
//                      halo range of one plane is ONE contiguous byte range and one
//                      cp.async.bulk per plane fills the tile  [cin/8][PP pixels][8 halfs]  -- the
//                      K-major, unswizzled UMMA operand layout, in which a tap shift is just
//                      +16 bytes per pixel on the descriptor's start address.  One mbarrier
//                      (expect_tx) hand-off per TILE; the weights are loaded once the same way.
//   warp 9             issues R * taps * cin/16 tcgen05.mma per tile back to back.
//   warps 0-7          epilogue out of double-buffered TMEM: bias, SiLU, residual, FP16; a warp
//                      stores 32 consecutive pixels of one plane = 512 contiguous bytes; padding
//                      pixels are skipped so the zero border is preserved.
//
// Against the per-tap gather kernel this cuts L2->SM traffic and load instructions ~4x for 3x3
// layers and the number of producer/consumer hand-offs 9x.
#include <cstdlib>

#include "common.cuh"

namespace irmv {
namespace {

constexpr int NEPI_WARPS = 8;               // warps 0-7 (TMEM lane quarter = warp % 4)
constexpr int TMA_WARP = 8;
constexpr int MMA_WARP = 9;
constexpr int NTHREADS = 10 * 32;
constexpr int MAX_STAGES = 4;
constexpr int SMEM_BUDGET = 226 * 1024;     // 232448 B is the opt-in maximum per CTA

struct Bars {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t bfull;
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 %%rx;\n"
      ".reg .pred %%px;\n"
      "elect.sync %%rx|%%px, %1;\n"
      "@%%px mov.s32 %0, 1;\n"
      "}" : "+r"(pred) : "r"(0xffffffffu));
  return pred;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand without swizzle: 8x(16 B) core matrices; SBO = bytes between 8-row groups
// (128, rows are 16 B apart), LBO = bytes between the two 16-byte K halves of one MMA.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}

__device__ __forceinline__ float silu(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

struct RArgs {
  ConvParams p;
  int R, TM, PP, halo_front, npix_need;   // rows = 128*R; PP = plane pitch (pixels); halo before q0
  int NCH, taps, Wp, Hp1;                 // cin/8, 1 or 9, W+1, H+1
  int q_begin, q_end, num_tiles, stages, tmem_cols, ctas_per_sm;
  long long npix;                         // raster pixels of the tensors for this batch
  uint32_t idesc, mul_wp, mul_hp1;        // magic dividers (q / Wp, row / (H+1)), >> 34
  uint32_t a_stage_bytes, b_bytes, off_b, off_bias, off_bars;
};

__global__ void __launch_bounds__(NTHREADS) conv_raster_kernel(const __grid_constant__ RArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvParams &p = a.p;
  uint8_t *sA = smem;
  uint8_t *sB = smem + a.off_b;
  float *s_bias = reinterpret_cast<float *>(smem + a.off_bias);
  Bars *bars = reinterpret_cast<Bars *>(smem + a.off_bars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int npad = p.npad;

  for (int i = tid; i < npad; i += NTHREADS) s_bias[i] = p.bias[i];
  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], NEPI_WARPS);
    }
    mbar_init(&bars->bfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&bars->tmem_base)), "r"((uint32_t)a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp < NEPI_WARPS) {
    // ===================================================================== epilogue
    // warp w reads TMEM lanes 32*(w%4)..+31 (hardware rule); the two warps of a lane quarter split
    // the (accumulator, 16-column chunk) work items between them.
    const int ew = warp & 3, half = warp >> 2;
    const int chunks = npad >> 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      const bool tr = p.trace && blockIdx.x == 0 && tid == 0 && it < p.trace_cap;
      mbar_wait_warp(&bars->tmem_full[buf], aph, lane);
      tc_fence_after();
      if (tr) p.trace[it * 8 + 6] = clock64();
      const int items = a.R * chunks;
      for (int w = half; w < items; w += 2) {
        const int r = w / chunks, c0 = (w - r * chunks) << 4;
        const int q = a.q_begin + tile * a.TM + r * 128 + ew * 32 + lane;
        // real pixel? (not the zero column x == W, not a zero row, inside the batch)
        const uint32_t row = (uint32_t)(((uint64_t)(uint32_t)q * a.mul_wp) >> 34);
        const int x = q - (int)row * a.Wp;
        const uint32_t img = (uint32_t)(((uint64_t)row * a.mul_hp1) >> 34);
        const int yrow = (int)row - (int)img * a.Hp1;
        const bool ok = q < a.q_end && x < p.W && yrow != 0 && c0 < p.cout;
        const int pl = c0 >> 3;
        uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
        if (ok && p.res) {                                  // residual: issue the loads first
          q0 = *reinterpret_cast<const uint4 *>(p.res + (size_t)pl * p.res_pstride + (size_t)q * 8);
          q1 = *reinterpret_cast<const uint4 *>(p.res + (size_t)(pl + 1) * p.res_pstride + (size_t)q * 8);
        }
        uint32_t v32[16];
        tc_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((buf * a.R + r) * npad + c0), v32);
        tc_ld_wait();
        if (!ok) continue;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float xv = __uint_as_float(v32[j]) + s_bias[c0 + j];
          v[j] = p.act ? silu(xv) : xv;
        }
        if (p.res) {
          const __half2 *h0 = reinterpret_cast<const __half2 *>(&q0);
          const __half2 *h1 = reinterpret_cast<const __half2 *>(&q1);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float2 f0 = __half22float2(h0[t]), f1 = __half22float2(h1[t]);
            v[2 * t] += f0.x; v[2 * t + 1] += f0.y;
            v[8 + 2 * t] += f1.x; v[8 + 2 * t + 1] += f1.y;
          }
        }
        __half2 hv[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) hv[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
        *reinterpret_cast<uint4 *>(p.out + (size_t)pl * p.out_pstride + (size_t)q * 8) = *reinterpret_cast<uint4 *>(&hv[0]);
        *reinterpret_cast<uint4 *>(p.out + (size_t)(pl + 1) * p.out_pstride + (size_t)q * 8) = *reinterpret_cast<uint4 *>(&hv[4]);
      }
      // all of this warp's TMEM reads of the buffer are done: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
      if (tr) p.trace[it * 8 + 7] = clock64();
    }
  } else if (warp == TMA_WARP) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      mbar_expect_tx(&bars->bfull, a.b_bytes);
      for (uint32_t off = 0; off < a.b_bytes; off += 65536u) {
        const uint32_t n = a.b_bytes - off < 65536u ? a.b_bytes - off : 65536u;
        bulk_g2s(sB + off, reinterpret_cast<const uint8_t *>(p.w_raster) + off, n, &bars->bfull);
      }
      int s = 0, itl = 0;
      uint32_t ph = 0;
      const uint32_t plane_bytes = (uint32_t)a.npix_need * 16u;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++itl) {
        const bool tr = p.trace && blockIdx.x == 0 && itl < p.trace_cap;
        if (tr) p.trace[itl * 8 + 0] = clock64();
        const long long qlo = (long long)a.q_begin + (long long)tile * a.TM - a.halo_front;
        mbar_wait(&bars->empty[s], ph ^ 1u);
        if (tr) p.trace[itl * 8 + 1] = clock64();
        uint8_t *stage = sA + (size_t)s * a.a_stage_bytes;
        mbar_expect_tx(&bars->full[s], plane_bytes * (uint32_t)a.NCH);
        int plane = 0;
        for (int sg = 0; sg < p.nseg; ++sg) {
          const __half *src = p.seg[sg].ptr + qlo * 8;
          const long long ps = p.seg[sg].pstride;
          for (int c = 0; c < (p.seg[sg].c >> 3); ++c, ++plane)
            bulk_g2s(stage + (size_t)plane * a.PP * 16, src + (long long)c * ps, plane_bytes, &bars->full[s]);
        }
        if (tr) p.trace[itl * 8 + 2] = clock64();
        if (++s == a.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===================================================================== MMA issuer
    mbar_wait_warp(&bars->bfull, 0, lane);
    int it = 0, s = 0;
    uint32_t ph = 0;
    const uint32_t lbo_a = (uint32_t)a.PP * 16u, lbo_b = (uint32_t)npad * 16u;
    const int pairs = a.NCH >> 1;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      const bool tr = p.trace && blockIdx.x == 0 && lane == 0 && it < p.trace_cap;
      if (tr) p.trace[it * 8 + 3] = clock64();
      mbar_wait_warp(&bars->tmem_empty[buf], aph ^ 1u, lane);
      mbar_wait_warp(&bars->full[s], ph, lane);
      tc_fence_after();
      if (tr) p.trace[it * 8 + 4] = clock64();
      if (elect_one()) {
        const uint32_t abase = smem_u32(sA + (size_t)s * a.a_stage_bytes);
        const uint32_t bbase = smem_u32(sB);
        // Loop order: accumulator index r innermost.  MMAs into the same TMEM tile form a dependent
        // chain (measured ~260-320 cycles per MMA at N = 16/32, i.e. pipeline latency, not
        // throughput), so consecutive issues go to the R independent accumulators of the tile.
        const uint32_t tmem_d0 = tmem_base + (uint32_t)(buf * a.R * npad);
        for (int t = 0; t < a.taps; ++t) {
          const int shift = a.taps == 9 ? (t / 3) * a.Wp + (t % 3) : 0;   // tap shift in pixels
          const uint32_t arow = abase + (uint32_t)(shift << 4);
          const uint32_t brow = bbase + (uint32_t)(t * a.NCH * npad) * 16u;
          for (int j = 0; j < pairs; ++j) {
            const uint64_t bdesc = make_desc(brow + (uint32_t)(2 * j) * lbo_b, lbo_b);
            const uint32_t acol = arow + (uint32_t)(2 * j) * lbo_a;
            const uint32_t acc = (uint32_t)((t | j) != 0);
            for (int r = 0; r < a.R; ++r)
              tc_mma_f16(tmem_d0 + (uint32_t)(r * npad), make_desc(acol + (uint32_t)(r * 128 * 16), lbo_a), bdesc,
                         a.idesc, acc);
          }
        }
        tc_commit(&bars->empty[s]);
        tc_commit(&bars->tmem_full[buf]);
      }
      __syncwarp();
      if (tr) p.trace[it * 8 + 5] = clock64();
      if (++s == a.stages) { s = 0; ph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
  }
}

bool plan(const ConvParams &p, int num_sms, RArgs &a) {
  if (p.stride != 1 || !(p.k == 1 || p.k == 3) || (p.k == 3 && p.pad != 1) || (p.k == 1 && p.pad != 0)) return false;
  if (p.cin % 16 != 0 || p.npad % 16 != 0 || p.npad > 256) return false;
  for (int i = 0; i < p.nseg; ++i) if (p.seg[i].up || p.seg[i].c % 8) return false;
  if (p.W + 2 > kGuardFront || p.cout % 16 != 0) return false;
  a.p = p;
  a.NCH = p.cin / 8;
  a.taps = p.k * p.k;
  a.Wp = p.W + 1;
  a.Hp1 = p.H + 1;
  a.npix = pr_pixels(p.B, p.H, p.W);
  a.q_begin = a.Wp;                                  // first pixel of raster row 1
  a.q_end = (p.B * (p.H + 1)) * a.Wp;               // end of the last image row
  const long long Mr = (long long)a.q_end - a.q_begin;
  a.b_bytes = (uint32_t)((size_t)a.taps * p.cin * p.npad * 2);
  const size_t misc = (size_t)p.npad * 4 + sizeof(Bars) + 1024 + 256;
  const int halo = p.k == 3 ? 2 * a.Wp + 2 : 0;
  int best_R = 0;
  for (int R = 4; R >= 1; R >>= 1) {
    if (2 * R * p.npad > 512) continue;
    int PP = 128 * R + halo;
    PP += (9 - (PP & 7)) & 7;                        // PP = 1 (mod 8): conflict-free plane writes
    size_t stage = (((size_t)a.NCH * PP * 16) + 127) & ~(size_t)127;
    if (a.b_bytes + 2 * stage + misc > (size_t)SMEM_BUDGET) continue;
    long long tiles = (Mr + 128 * R - 1) / (128 * R);
    if (R > 1 && tiles < 2LL * num_sms) continue;    // keep every SM busy at small batch
    best_R = R;
    break;
  }
  if (!best_R) return false;
  a.R = best_R;
  a.TM = 128 * a.R;
  a.halo_front = p.k == 3 ? a.Wp + 1 : 0;
  a.npix_need = a.TM + halo;
  a.PP = a.npix_need + ((9 - (a.npix_need & 7)) & 7);
  a.a_stage_bytes = (uint32_t)((((size_t)a.NCH * a.PP * 16) + 127) & ~(size_t)127);
  int cols = 2 * a.R * p.npad, alloc = 32;
  while (alloc < cols) alloc <<= 1;
  a.tmem_cols = alloc;
  // The per-tile chain load -> MMA -> epilogue is latency-bound (profiles: warps mostly asleep on
  // barriers), so when two CTAs fit on an SM (shared memory and TMEM halves) run two: their
  // phases interleave and hide each other's latencies.
  const size_t half_budget = 112 * 1024;
  const bool two = !getenv("IRMV_ONE_CTA") && alloc <= 256 && a.b_bytes + 2 * (size_t)a.a_stage_bytes + misc <= half_budget;
  a.ctas_per_sm = two ? 2 : 1;
  const size_t budget = two ? half_budget : (size_t)SMEM_BUDGET;
  a.stages = (int)((budget - a.b_bytes - misc) / a.a_stage_bytes);
  if (a.stages > MAX_STAGES) a.stages = MAX_STAGES;
  a.num_tiles = (int)((Mr + a.TM - 1) / a.TM);
  a.idesc = (1u << 4) | ((uint32_t)(p.npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  a.mul_wp = (uint32_t)(((1ull << 34) + (uint64_t)a.Wp - 1) / (uint64_t)a.Wp);
  a.mul_hp1 = (uint32_t)(((1ull << 34) + (uint64_t)a.Hp1 - 1) / (uint64_t)a.Hp1);
  a.off_b = (uint32_t)((size_t)a.stages * a.a_stage_bytes);
  a.off_bias = (a.off_b + a.b_bytes + 127u) & ~127u;
  a.off_bars = (a.off_bias + (uint32_t)p.npad * 4 + 15u) & ~15u;
  return true;
}

}  // namespace

bool conv_raster_fits(const ConvParams &p) {
  RArgs a;
  return p.w_raster != nullptr && plan(p, 148, a);
}

cudaError_t launch_conv_raster(const ConvParams &p, int num_sms, cudaStream_t s) {
  RArgs a;
  if (!plan(p, num_sms, a)) return cudaErrorInvalidValue;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_raster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  size_t smem = (size_t)a.off_bars + sizeof(Bars) + 64;
  const int slots = num_sms * a.ctas_per_sm;
  int grid = a.num_tiles < slots ? a.num_tiles : slots;
  conv_raster_kernel<<<grid, NTHREADS, smem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace irmv

// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
// Replaces the TensorRT FP16 engine body of the reference (reference src/yolo_engine.cpp:100-105).
//
//   D[M = B*OH*OW pixels][N = cout] = im2col(A)[M][K = k*k*cin] * W[N][K]^T   (FP16 in, FP32 accum)
//
// Persistent, warp-specialised CTA (one per SM), 128-row tiles:
//   warps 0-7   A producers: gather im2col rows straight from the NHWC activation(s) with 16-byte
//               cp.async into a 128B-swizzled K-major smem ring.  The gather does the padding
//               (zero fill), the stride, concat-on-read (two segments) and nearest-2x
//               upsample-on-read, so no im2col / concat / upsample tensor ever touches HBM.
//   warp  13    weight loader: the host pre-swizzles W into the exact smem image, so whole
//               k-blocks arrive with TMA bulk copies (cp.async.bulk + mbarrier complete_tx);
//               resident for the CTA's lifetime when they fit, streamed per stage otherwise.
//   warp  12    MMA issuer: one lane issues tcgen05.mma (M=128, N=npad, K=16) from smem
//               descriptors; tcgen05.commit releases ring slots and publishes accumulators.
//   warps 8-11  epilogue: tcgen05.ld accumulator rows out of TMEM (double-buffered, 2*N columns),
//               + bias, SiLU, residual, FP16 pack, channel-slice store.
#include "common.cuh"

namespace irmv {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE = BM * BK * 2;          // 16 KB
constexpr int NPROD = 256;                    // warps 0-7
constexpr int EPI_WARP0 = 8;
constexpr int MMA_WARP = 12;
constexpr int BLD_WARP = 13;
constexpr int NTHREADS = 14 * 32;
constexpr int MAX_STAGES = 6;
constexpr int SMEM_BUDGET = 160 * 1024;       // leave L1 room for the 3x3 tap re-reads

struct Bars {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t bfull;
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
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
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand: rows are 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address
  d |= (uint64_t)1 << 16;                          // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // SBO
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

struct TcArgs {
  ConvParams p;
  int M, num_tiles, KB, ksteps_last, stages, b_resident, tmem_cols;
  uint32_t idesc;
  uint32_t mul_ow, mul_oh;      // ceil(2^34 / OW), ceil(2^34 / OH): q = (n * mul) >> 34, exact for n < 2^25
  uint32_t off_b, off_ktab, off_bias, off_bars;
};

__global__ void __launch_bounds__(NTHREADS, 1) conv_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvParams &p = a.p;
  uint8_t *sA = smem;
  uint8_t *sB = smem + a.off_b;
  int32_t *s_ktab = reinterpret_cast<int32_t *>(smem + a.off_ktab);
  float *s_bias = reinterpret_cast<float *>(smem + a.off_bias);
  Bars *bars = reinterpret_cast<Bars *>(smem + a.off_bars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int npad = p.npad;
  const uint32_t b_block_bytes = (uint32_t)npad * 128u;

  for (int i = tid; i < a.KB * 16; i += NTHREADS) s_ktab[i] = p.ktab[i];
  for (int i = tid; i < npad; i += NTHREADS) s_bias[i] = p.bias[i];
  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&bars->full[s], NPROD + (a.b_resident ? 0 : 1));
      mbar_init(&bars->empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], 128);
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

  if (warp < EPI_WARP0) {
    // ===================================================================== A producers
    // Per tile: decode the thread's 4 rows once (magic-number division), keep one 32-bit element
    // offset per (segment, row) and a bitmask of the taps that fall inside the image.  Per
    // k-block: one table lookup gives the tap's element offset, so a row costs a bit test, one
    // add and the cp.async.
    const int chunk = tid & 7;
    const int rbase = tid >> 3;                       // 0..31, rows rbase + 32*i
    const uint32_t dst_off = (uint32_t)rbase * 128u + (uint32_t)((chunk ^ (rbase & 7)) << 4);
    const int H = p.H, W = p.W, ksz = p.k;
    const __half *sp0 = p.seg[0].ptr, *sp1 = p.seg[1].ptr;
    const int cs0 = p.seg[0].cstride, cs1 = p.seg[1].cstride;
    const int up0 = p.seg[0].up, up1 = p.seg[1].up;
    const int two = p.nseg > 1;
    int g = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      int off0[4], off1[4];
      uint32_t mask[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = tile * BM + rbase + 32 * i;
        off0[i] = 0; off1[i] = 0; mask[i] = 0u;
        if (m < a.M) {
          const uint32_t t = (uint32_t)(((uint64_t)(uint32_t)m * a.mul_ow) >> 34);      // m / OW
          const int ox = m - (int)t * p.OW;
          const uint32_t bimg = (uint32_t)(((uint64_t)t * a.mul_oh) >> 34);             // t / OH
          const int oy = (int)t - (int)bimg * p.OH;
          const int iy0 = oy * p.stride, ix0 = ox * p.stride;
          off0[i] = up0 ? (((int)bimg * (H >> 1) + (iy0 >> 1)) * (W >> 1) + (ix0 >> 1)) * cs0
                        : (((int)bimg * H + iy0) * W + ix0) * cs0;
          if (two)
            off1[i] = up1 ? (((int)bimg * (H >> 1) + (iy0 >> 1)) * (W >> 1) + (ix0 >> 1)) * cs1
                          : (((int)bimg * H + iy0) * W + ix0) * cs1;
          uint32_t mk = 0u;
          for (int ky = 0; ky < ksz; ++ky) {
            const int iy = iy0 - p.pad + ky;
            const bool yok = iy >= 0 && iy < H;
            for (int kx = 0; kx < ksz; ++kx) {
              const int ix = ix0 - p.pad + kx;
              if (yok && ix >= 0 && ix < W) mk |= 1u << (ky * ksz + kx);
            }
          }
          mask[i] = mk;
        }
      }
      for (int kb = 0; kb < a.KB; ++kb, ++g) {
        const int s = g % a.stages;
        const uint32_t ph = (uint32_t)(g / a.stages) & 1u;
        const int2 e = *reinterpret_cast<const int2 *>(&s_ktab[(kb * 8 + chunk) * 2]);
        const int tapbit = e.y & 15, sg = (e.y >> 4) & 1;
        const uint32_t kvalid = (uint32_t)(e.y >> 5) & 1u;
        const __half *sbase = (sg ? sp1 : sp0) + e.x;
        const uint32_t dst0 = smem_u32(sA + (size_t)s * A_STAGE) + dst_off;
        mbar_wait(&bars->empty[s], ph ^ 1u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t ok = kvalid & (mask[i] >> tapbit);
          const __half *src = ok ? sbase + (sg ? off1[i] : off0[i]) : sp0;
          cp_async16(dst0 + (uint32_t)i * (32u * 128u), src, ok ? 16u : 0u);
        }
        if (p.sync_mode == 0) {
          cp_async_mbar_arrive(&bars->full[s]);
          mbar_arrive(&bars->full[s]);
        } else {
          asm volatile("cp.async.commit_group;" ::: "memory");
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(&bars->full[s]);
        }
      }
    }
  } else if (warp < MMA_WARP) {
    // ===================================================================== epilogue
    const int ew = warp - EPI_WARP0;
    const int row = ew * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&bars->tmem_full[acc], aph);
      tc_fence_after();
      const int m = tile * BM + row;
      const bool mok = m < a.M;
      __half *orow = p.out + (size_t)(mok ? m : 0) * p.out_cstride + p.out_coff;
      const __half *rrow = p.res ? p.res + (size_t)(mok ? m : 0) * p.res_cstride + p.res_coff : nullptr;
      const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * npad);
      for (int c0 = 0; c0 < npad; c0 += 16) {
        uint32_t r[16];
        tc_ld16(tbase + (uint32_t)c0, r);
        tc_ld_wait();
        if (c0 + 16 >= npad) {           // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(&bars->tmem_empty[acc]);
        }
        if (!mok || c0 >= p.cout) continue;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = __uint_as_float(r[j]) + s_bias[c0 + j];
          v[j] = p.act ? silu(x) : x;
        }
        if (rrow) {
          uint4 q0 = *reinterpret_cast<const uint4 *>(rrow + c0);
          uint4 q1 = *reinterpret_cast<const uint4 *>(rrow + c0 + 8);
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
        *reinterpret_cast<uint4 *>(orow + c0) = *reinterpret_cast<uint4 *>(&hv[0]);
        if (c0 + 8 < p.cout) *reinterpret_cast<uint4 *>(orow + c0 + 8) = *reinterpret_cast<uint4 *>(&hv[4]);
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================================================================== MMA issuer
    if (a.b_resident) mbar_wait(&bars->bfull, 0);
    int g = 0, it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&bars->tmem_empty[acc], aph ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * npad);
      for (int kb = 0; kb < a.KB; ++kb, ++g) {
        const int s = g % a.stages;
        const uint32_t ph = (uint32_t)(g / a.stages) & 1u;
        mbar_wait(&bars->full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = make_desc(smem_u32(sA + (size_t)s * A_STAGE));
          const uint64_t bdesc = make_desc(smem_u32(sB + (size_t)(a.b_resident ? kb : s) * b_block_bytes));
          const int nks = (kb == a.KB - 1) ? a.ksteps_last : 4;
          for (int ks = 0; ks < nks; ++ks)
            tc_mma_f16(tmem_d, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), a.idesc,
                       (uint32_t)((kb | ks) != 0));
          tc_commit(&bars->empty[s]);
          if (kb == a.KB - 1) tc_commit(&bars->tmem_full[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================================== weight loader
    if (lane == 0) {
      if (a.b_resident) {
        mbar_expect_tx(&bars->bfull, (uint32_t)a.KB * b_block_bytes);
        for (int kb = 0; kb < a.KB; ++kb)
          bulk_g2s(sB + (size_t)kb * b_block_bytes, p.w_tiled + (size_t)kb * npad * BK,
                   b_block_bytes, &bars->bfull);
      } else {
        int g = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
          for (int kb = 0; kb < a.KB; ++kb, ++g) {
            const int s = g % a.stages;
            const uint32_t ph = (uint32_t)(g / a.stages) & 1u;
            mbar_wait(&bars->empty[s], ph ^ 1u);
            mbar_expect_tx(&bars->full[s], b_block_bytes);
            bulk_g2s(sB + (size_t)s * b_block_bytes, p.w_tiled + (size_t)kb * npad * BK,
                     b_block_bytes, &bars->full[s]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
  }
}

}  // namespace

size_t conv_tc_smem_bytes(const ConvParams &p, int *stages_out, int *b_resident_out) {
  const int KB = p.kpad / BK;
  const size_t b_block = (size_t)p.npad * 128;
  const size_t misc = (size_t)KB * 16 * 4 + (size_t)p.npad * 4 + sizeof(Bars) + 256;
  int resident = 0, stages = 0;
  if ((size_t)KB * b_block + 3 * (size_t)A_STAGE + misc <= (size_t)SMEM_BUDGET) {
    resident = 1;
    stages = (int)(((size_t)SMEM_BUDGET - (size_t)KB * b_block - misc) / A_STAGE);
    if (stages > 4) stages = 4;
  } else {
    size_t budget = 200 * 1024;
    stages = (int)((budget - misc) / (A_STAGE + b_block));
    if (stages > 4) stages = 4;
  }
  if (stages > KB + 1) stages = KB + 1;   // no point in more slots than one tile can fill twice
  if (stages < 2) stages = 2;
  if (stages_out) *stages_out = stages;
  if (b_resident_out) *b_resident_out = resident;
  size_t b_bytes = resident ? (size_t)KB * b_block : (size_t)stages * b_block;
  return (size_t)stages * A_STAGE + b_bytes + misc + 1024;
}

cudaError_t launch_conv_tc(const ConvParams &p, int num_sms, cudaStream_t s) {
  TcArgs a;
  a.p = p;
  a.M = p.B * p.OH * p.OW;
  a.num_tiles = (a.M + BM - 1) / BM;
  a.KB = p.kpad / BK;
  const int ksteps_total = (p.K + 15) / 16;
  a.ksteps_last = ksteps_total - (a.KB - 1) * 4;
  size_t smem = conv_tc_smem_bytes(p, &a.stages, &a.b_resident);
  int cols = 2 * p.npad;
  int alloc = 32;
  while (alloc < cols) alloc <<= 1;
  a.tmem_cols = alloc;
  // instruction descriptor: D=F32, A=B=F16, both K-major, N = npad, M = 128
  a.idesc = (1u << 4) | ((uint32_t)(p.npad >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const size_t b_block = (size_t)p.npad * 128;
  a.off_b = (uint32_t)((size_t)a.stages * A_STAGE);
  size_t b_bytes = a.b_resident ? (size_t)a.KB * b_block : (size_t)a.stages * b_block;
  a.off_ktab = (uint32_t)(a.off_b + b_bytes);
  a.off_bias = a.off_ktab + (uint32_t)a.KB * 16 * 4;
  a.mul_ow = (uint32_t)(((1ull << 34) + (uint64_t)p.OW - 1) / (uint64_t)p.OW);
  a.mul_oh = (uint32_t)(((1ull << 34) + (uint64_t)p.OH - 1) / (uint64_t)p.OH);
  a.off_bars = (a.off_bias + (uint32_t)p.npad * 4 + 15u) & ~15u;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(227 * 1024));
    if (e != cudaSuccess) return e;
    configured = 227 * 1024;
  }
  int grid = a.num_tiles < num_sms ? a.num_tiles : num_sms;
  conv_tc_kernel<<<grid, NTHREADS, smem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace irmv

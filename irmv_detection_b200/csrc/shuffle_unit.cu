// Fused ShuffleNetV2 units (sm_100a): one kernel per unit instead of three to five launches that
// each stream the tensors through HBM / L2.  BASELINE.json configs[2] (the keypoint detector on a
// ShuffleNetV2-style backbone; the reference names the model, README.md:12,16) and the north_star's
// "depthwise and elementwise layers are fused bandwidth-bound kernels".
//
//   basic unit  (Ma et al. 2018, fig. 3c):  x1 | x2 -> x1 | pw2(dw3x3(pw1(x2)))        -> shuffle
//   down unit   (fig. 3d):  pwA(dwA_s2(x)) | pw2(dwB_s2(pw1(x)))                       -> shuffle
//
// Persistent CTAs: the unit's weights (one packed blob, shuffle_blob_layout) are fetched once per CTA with
// one bulk copy and stay in shared memory; the CTA then walks tiles of TH full-width output rows of one
// image.  A tile's input is ONE contiguous run of raster pixels per plane (the zero-padded raster layout,
// common.cuh: row pitch W + 1, so the left / right / top / bottom padding the depthwise conv wants is
// already in the run) -> one cp.async.bulk per plane onto an mbarrier, and the next tile's copy is issued
// as soon as the last reader of the input tile is done, i.e. it lands under the second half of the current
// tile.  Everything runs on the tensor cores with mma.sync.m16n8k16:
//   * the 1x1 convolutions straight from the tile (rows = pixels, [plane][pixel][8 channels] -> every A
//     fragment register is one conflict-free 32-bit load);
//   * the depthwise 3x3 convs as a block-diagonal GEMM per 8-channel plane (K = 9 taps x 8 channels, see
//     dw_frags), whose accumulator fragment IS the A fragment of the 1x1 that follows -- the depthwise output
//     never leaves registers.  (The CUDA-core form, FP16 -> FP32 convert + FMA per tap, cost 17 issue slots
//     per output value and made the unit issue bound at half the speed.)
// Per tile: phase 1 = first 1x1 over the input tile -> T1 in shared memory as FP16 (exactly the values the
// unfused path stores in HBM; zero outside the image, which is the depthwise conv's padding), and for down
// units branch 1 (depthwise s2 on the input -> 1x1 -> output); barrier; phase 2 = depthwise on T1 -> second
// 1x1 -> output; barrier.  HBM traffic of a unit = its input + its output (a basic unit also copies its
// pass-through half into the other buffer of the stage's ping-pong pair: reading a halo that another CTA may
// already have rewritten rules out updating in place).
// Channel split / concat / shuffle stay the logical -> physical channel map folded into the weights
// (engine.cu), so the unit reads and writes plane runs of the stage buffer (ConvSeg::runs).
//
// Bound: HBM by bytes (2 B in + 2 B out per value against ~3 h MACs), issue / shared-memory pipe in practice
// (profiles/r2_summary.md section 12); used for the stages with h <= 64, whose 1x1 GEMMs (K, N <= 64) are far
// too small for a tcgen05 tile to pay for its TMEM round trip, hence mma.sync.
// The last stage (h = 128, 20 x 20 maps) is weight-heavy and stays one tcgen05 launch per convolution.
#include "common.cuh"

namespace irmv {
namespace {

constexpr int NT = 256;

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
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// floor(m / d) for m < 65536 through the reciprocal inv = recip_u16(d) (a 32-bit integer division costs ~25 issue slots)
__device__ __forceinline__ uint32_t recip_u16(int d) { return 0xFFFFFFFFu / (uint32_t)d + 1u; }
__device__ __forceinline__ int fast_div(int m, uint32_t inv) { return (int)__umulhi((uint32_t)m, inv); }

// SiLU(x) from h = x / 2 (the 1/2 is folded into the 1x1 weights and biases of the blob): h + h * tanh(h), one MUFU
__device__ __forceinline__ float silu_f(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&h);
}

// Pointwise conv of one 16-pixel m-tile: out[m][n] = silu(2 * (sum_k A[m][k] * W[n][k] + bias[n])) with W, bias = half the
// layer's (blob), K = N = C (every unit of
// the backbone has cin == h on the fused stages).  a: the m-tile's A fragments (a[ks] = k-step ks: planes 2ks, 2ks+1);
// sW: [C][C + 8] halves; bias: [C] floats.  All C/8 n-tile accumulators are live at once (C/8 independent HMMA
// chains, the B fragments stream from shared memory).  epi(cookie, nt, packed): row cookie r0 = row m0 + g,
// r1 = row m0 + g + 8; packed = the two activated FP16 values of channels nt*8 + 2t, + 1.
template <int C, typename Cookie, typename Epi>
__device__ __forceinline__ void pw_tile(const uint32_t (&a)[C / 16][4], const __half *sW, const float *bias, int g, int t,
                                        Cookie r0, Cookie r1, Epi epi) {
  constexpr int KS = C / 16, NTL = C / 8, WP2 = (C + 8) / 2;                   // k-steps, n-tiles, weight row pitch in 32-bit words
  const uint32_t *W32 = reinterpret_cast<const uint32_t *>(sW) + g * WP2 + t;
  const float2 *b2 = reinterpret_cast<const float2 *>(bias) + t;
  float c[NTL][4];
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt) {
    const float2 bv = b2[nt * 4];
    c[nt][0] = bv.x; c[nt][1] = bv.y; c[nt][2] = bv.x; c[nt][3] = bv.y;
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt)
      mma16816(c[nt], a[ks][0], a[ks][1], a[ks][2], a[ks][3], W32[nt * 8 * WP2 + ks * 8], W32[nt * 8 * WP2 + ks * 8 + 4]);
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt) {
    epi(r0, nt, pack_h2(silu_f(c[nt][0]), silu_f(c[nt][1])));
    epi(r1, nt, pack_h2(silu_f(c[nt][2]), silu_f(c[nt][3])));
  }
}

// A fragments of the m-tile starting at pixel m0 of a shared-memory tile [C/8][Mp][8] halves (plane-major): every
// register is one conflict-free 32-bit load (plane 2ks / 2ks+1, pixel m0+g / m0+g+8, channels 2t, 2t+1).
template <int C>
__device__ __forceinline__ void load_frags(const __half *sA, int Mp, int m0, int g, int t, uint32_t (&a)[C / 16][4]) {
  const uint32_t *ap = reinterpret_cast<const uint32_t *>(sA) + (m0 + g) * 4 + t;
#pragma unroll
  for (int ks = 0; ks < C / 16; ++ks) {
    a[ks][0] = ap[(2 * ks) * Mp * 4]; a[ks][1] = ap[(2 * ks) * Mp * 4 + 32];
    a[ks][2] = ap[(2 * ks + 1) * Mp * 4]; a[ks][3] = ap[(2 * ks + 1) * Mp * 4 + 32];
  }
}

// Depthwise 3x3 (+ bias) of one 16-pixel m-tile ON THE TENSOR CORES, result delivered as the A fragments of the 1x1 conv
// that follows.  Per 8-channel plane the depthwise conv is the GEMM
//     out[px][n] = sum_{k = (tap, c')} in[px + off(tap)][c'] * (c' == n ? w[tap][n] : 0),
// K = 9 taps x 8 channels padded to 80 = five m16n8k16 steps of two taps each: an A register is one 32-bit load from the
// tile at the tap's pixel offset, a B register is the lane's table word (blob layout) or zero.  Seven eighths of the MACs
// multiply by zero, which the tensor pipe has to spare; the CUDA-core form costs ~17 issue slots per output value
// (convert + FMA per tap), this one 0.3.  The accumulator fragment of plane p (rows g / g+8, channels 2t, 2t+1) IS the A
// fragment register pair of plane p in the following 1x1 GEMM, so the depthwise output never touches shared memory.
//   sIn: [PL][inMp][8] halves, pixels in rows of pitch inW; base0 / base1: tile pixel of the window's corner for rows
//   g / g+8; tab: [PL][10][8] words; bias: [PL * 8]; keep: all-ones in the lanes with t == g / 2, else zero.
template <int PL>
__device__ __forceinline__ void dw_frags(const __half *sIn, int inMp, int inW, int base0, int base1, const uint32_t *tab,
                                         const float *bias, int g, int t, uint32_t keep, uint32_t (&a)[PL / 2][4]) {
  const uint32_t *i0 = reinterpret_cast<const uint32_t *>(sIn) + base0 * 4 + t;
  const uint32_t *i1 = reinterpret_cast<const uint32_t *>(sIn) + base1 * 4 + t;
  const float2 *b2 = reinterpret_cast<const float2 *>(bias) + t;
  const uint32_t *tp = tab + g;
  const int rw = inW * 4, pw = inMp * 4;                                       // words per tile row / per plane
#pragma unroll
  for (int p = 0; p < PL; ++p, i0 += pw, i1 += pw, tp += 80) {
    const float2 bv = b2[p * 4];
    float c[4] = {bv.x, bv.y, bv.x, bv.y};
    // one pointer per window row: every tap is then [row pointer + constant]
    const uint32_t *q0[3] = {i0, i0 + rw, i0 + 2 * rw}, *q1[3] = {i1, i1 + rw, i1 + 2 * rw};
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) {
      const int ta = 2 * ks, tb = ks < 4 ? 2 * ks + 1 : 8;                     // the pad tap re-reads tap 8's pixel (finite) against zero weights
      mma16816(c, q0[ta / 3][(ta % 3) * 4], q1[ta / 3][(ta % 3) * 4], q0[tb / 3][(tb % 3) * 4], q1[tb / 3][(tb % 3) * 4],
               tp[(2 * ks) * 8] & keep, tp[(2 * ks + 1) * 8] & keep);
    }
    a[p >> 1][(p & 1) * 2] = pack_h2(c[0], c[1]);
    a[p >> 1][(p & 1) * 2 + 1] = pack_h2(c[2], c[3]);
  }
}

// C = the unit's half width h = its input width cin (16 / 32 / 64), DOWN = down unit
template <int C, bool DOWN>
__global__ void __launch_bounds__(NT, C == 16 ? 3 : 2) shuffle_unit_kernel(const __grid_constant__ ShuffleUnitParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int PL = C / 8, S = DOWN ? 2 : 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int H = p.H, W = p.W, TH = p.TH;
  const int Hin = S * H, Win = S * W;
  // input tile of one plane: rows S*y0 - 1 .. S*(y0 + TH - 1) + 1 as one run of raster pixels starting at
  // column -1 (pitch WC = Win + 1), plus the pixel after it (the right neighbour of the last pixel)
  const int HR = S * TH + (DOWN ? 1 : 2), WC = Win + 1;
  const int HP = HR * WC + 1, HPp = (HP + 15) & ~15;
  const int TP = TH * W;
  const ShuffleBlobLayout L = shuffle_blob_layout(DOWN, C, C);
  // shared memory: blob | sX [PL][HPp][8] | sT1 [PL][HPp][8] | output plane offsets | 2 mbarriers
  const __half *sW1 = reinterpret_cast<const __half *>(smem + L.w1), *sW2 = reinterpret_cast<const __half *>(smem + L.w2);
  const __half *sWa = reinterpret_cast<const __half *>(smem + L.wa);
  const uint32_t *sDw = reinterpret_cast<const uint32_t *>(smem + L.dw), *sDwa = reinterpret_cast<const uint32_t *>(smem + L.dwa);
  const float *sB = reinterpret_cast<const float *>(smem + L.bias);         // b1[C] b2[C] ba[C] dwb[C] dwab[C]
  __half *sX = reinterpret_cast<__half *>(smem + L.bytes);
  __half *sT1 = sX + (size_t)PL * HPp * 8;
  long long *sOff = reinterpret_cast<long long *>(sT1 + (size_t)PL * HPp * 8);   // [PL]: half offset of the plane the nt-th n-tile of the second 1x1 writes
  uint64_t *bars = reinterpret_cast<uint64_t *>(sOff + PL);
  const int tiles_per_img = H / TH, tiles = p.B * tiles_per_img;
  const uint32_t plane_bytes = (uint32_t)HP * 16;
  auto issue = [&](int tile) {                                              // one thread
    const int tt = p.rev ? tiles - 1 - tile : tile;
    const int b = tt / tiles_per_img, y0 = (tt - b * tiles_per_img) * TH;
    const __half *src = p.in + pr_index(b, S * y0 - 1, -1, Hin, Win) * 8;
    mbar_expect_tx(&bars[0], plane_bytes * PL);
#pragma unroll 1
    for (int pl = 0; pl < PL; ++pl) {
      const int plane = DOWN ? pl : p.first_plane + run_plane(pl, p.runs);
      bulk_g2s(sX + (size_t)pl * HPp * 8, src + (long long)plane * p.in_ps, plane_bytes, &bars[0]);
    }
  };
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bars[1], (uint32_t)L.bytes);
    bulk_g2s(smem, p.blob, (uint32_t)L.bytes, &bars[1]);
    if ((int)blockIdx.x < tiles) issue(blockIdx.x);
  }
  if (tid < PL) sOff[tid] = (long long)(DOWN ? PL + tid : p.first_plane + run_plane(tid, p.runs)) * p.out_ps;
  __syncthreads();
  mbar_wait(&bars[1], 0);
  const int t2 = 2 * t;
  const uint32_t keep = (t == (g >> 1)) ? 0xffffffffu : 0u;
  const uint32_t invW = recip_u16(W), invWC = recip_u16(WC);
  const int n_pw1 = (HP + 15) >> 4, n_out = (TP + 15) >> 4;                 // m-tiles of the first 1x1 / of the unit's output
  // corner (tile pixel index) of the depthwise window of output pixel m; pixels past the tile re-read its last pixel
  auto window = [&](int m) -> int {
    m = m < TP ? m : TP - 1;
    const int r = fast_div(m, invW), c = m - r * W;
    return (S * r) * WC + S * c;
  };
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
    const int tt = p.rev ? tiles - 1 - tile : tile;
    const int b = tt / tiles_per_img, y0 = (tt - b * tiles_per_img) * TH;
    const long long out0 = pr_index(b, y0, 0, H, W) * 8;                    // output pixel (y0, 0); the TH rows follow with pitch W + 1
    if (!DOWN) {
      // pass-through half: the planes one run before the unit's, copied to the other buffer of the ping-pong pair
      const int rp = p.runs ? (1 << (p.runs - 1)) : PL;
      const int run_px = TP + TH - 1;                                       // the TH rows of a plane are one contiguous run (pads included)
      // four loads in flight per thread before the first store: the copy is latency, not bandwidth
      const int total = PL * run_px;
      const uint32_t inv_run = recip_u16(run_px);
      for (int i0 = tid; i0 < total; i0 += 4 * NT) {
        uint4 v[4];
        long long off[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k * NT;
          off[k] = -1;
          if (i < total) {
            const int pl = fast_div(i, inv_run), px = i - pl * run_px;
            const int plane = p.first_plane + run_plane(pl, p.runs) - rp;
            off[k] = out0 + (long long)px * 8;
            v[k] = __ldg(reinterpret_cast<const uint4 *>(p.in + (long long)plane * p.in_ps + off[k]));
            off[k] += (long long)plane * p.out_ps;
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (off[k] >= 0) *reinterpret_cast<uint4 *>(p.out + off[k]) = v[k];
      }
    }
    mbar_wait(&bars[0], it & 1);
    const int ylo = 1 - S * y0, yhi = Hin + 1 - S * y0;                     // tile rows [ylo, yhi) are image rows
    // row cookie of an output pixel: where its channels 2t, 2t+1 of plane 0 go (null: past the tile)
    auto out_row = [&](int m) -> __half * {
      if (m >= TP) return nullptr;
      return p.out + out0 + (long long)(m + fast_div(m, invW)) * 8 + t2;    // pixel (r, c) sits r pad pixels further along the raster
    };
    // ---- phase 1: first 1x1 over the whole input tile -> T1; down units also run their branch 1 (depthwise s2 on the
    //      input, 1x1 into output planes [0, C/8)), which reads the input tile only
    for (int item = warp; item < n_pw1 + (DOWN ? n_out : 0); item += NT / 32) {
      uint32_t a[C / 16][4];
      if (item < n_pw1) {
        const int m0 = item * 16;
        load_frags<C>(sX, HPp, m0, g, t, a);
        // pixels outside the image must be ZERO in T1 (the depthwise conv pads T1, not the first 1x1's input)
        auto t1_row = [&](int m) -> int {
          const int r = fast_div(m, invWC), c = m - r * WC;
          return (c != 0 && r >= ylo && r < yhi) ? m : -1 - m;
        };
        pw_tile<C>(a, sW1, sB, g, t, t1_row(m0 + g), t1_row(m0 + g + 8), [&](int ck, int nt, uint32_t v) {
          const bool in = ck >= 0;
          const int m = in ? ck : -1 - ck;
          *reinterpret_cast<uint32_t *>(sT1 + ((size_t)nt * HPp + m) * 8 + t2) = in ? v : 0u;
        });
      } else {
        const int m0 = (item - n_pw1) * 16;
        dw_frags<PL>(sX, HPp, WC, window(m0 + g), window(m0 + g + 8), sDwa, sB + 4 * C, g, t, keep, a);
        pw_tile<C>(a, sWa, sB + 2 * C, g, t, out_row(m0 + g), out_row(m0 + g + 8), [&](__half *o, int nt, uint32_t v) {
          if (o) *reinterpret_cast<uint32_t *>(o + (long long)nt * p.out_ps) = v;
        });
      }
    }
    __syncthreads();                                                        // sX is free, sT1 complete
    if (tid == 0 && tile + (int)gridDim.x < tiles) issue(tile + gridDim.x);
    // ---- phase 2: depthwise on T1 -> second 1x1 -> the unit's planes of the output
    for (int item = warp; item < n_out; item += NT / 32) {
      const int m0 = item * 16;
      uint32_t a[C / 16][4];
      dw_frags<PL>(sT1, HPp, WC, window(m0 + g), window(m0 + g + 8), sDw, sB + 3 * C, g, t, keep, a);
      pw_tile<C>(a, sW2, sB + C, g, t, out_row(m0 + g), out_row(m0 + g + 8), [&](__half *o, int nt, uint32_t v) {
        if (o) *reinterpret_cast<uint32_t *>(o + sOff[nt]) = v;
      });
    }
    __syncthreads();                                                        // sT1 is rewritten by the next tile
  }
}

}  // namespace

size_t shuffle_unit_smem(const ShuffleUnitParams &p) {
  const int s = p.down ? 2 : 1;
  const int HR = s * p.TH + (p.down ? 1 : 2), WC = s * p.W + 1;
  const size_t HPp = ((size_t)HR * WC + 1 + 15) & ~(size_t)15;
  const size_t pl = p.h / 8;
  return (size_t)shuffle_blob_layout(p.down, p.cin, p.h).bytes + 2 * pl * HPp * 16 + pl * 8 + 16;
}

int shuffle_unit_ctas_per_sm(const ShuffleUnitParams &p) {
  const size_t smem = shuffle_unit_smem(p) + 1024;                          // + the per-CTA reservation
  const int by_smem = (int)((size_t)228 * 1024 / smem);
  const int by_regs = p.h == 16 ? 3 : 2;                                    // __launch_bounds__(256, 3 / 2): 85 / 128 registers per thread
  return by_smem < 1 ? 0 : (by_smem > by_regs ? by_regs : by_smem);
}

cudaError_t launch_shuffle_unit(const ShuffleUnitParams &p, cudaStream_t s) {
  if (p.B <= 0) return cudaSuccess;
  if (p.TH < 1 || p.H % p.TH || p.cin != p.h || !(p.h == 16 || p.h == 32 || p.h == 64)) return cudaErrorInvalidValue;
  const size_t smem = shuffle_unit_smem(p);
  const int per_sm = shuffle_unit_ctas_per_sm(p);
  if (per_sm < 1) return cudaErrorInvalidValue;
  void (*k)(const ShuffleUnitParams) = nullptr;
  if (p.down) k = p.h == 16 ? shuffle_unit_kernel<16, true> : (p.h == 32 ? shuffle_unit_kernel<32, true> : shuffle_unit_kernel<64, true>);
  else k = p.h == 16 ? shuffle_unit_kernel<16, false> : (p.h == 32 ? shuffle_unit_kernel<32, false> : shuffle_unit_kernel<64, false>);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int tiles = p.B * (p.H / p.TH);
  const int grid = std::min(tiles, (p.num_sms > 0 ? p.num_sms : 148) * per_sm);
  k<<<grid, NT, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace irmv

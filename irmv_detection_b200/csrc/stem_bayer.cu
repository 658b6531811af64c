// Fused stem for the camera case: 8-bit Bayer mosaic, source width exactly 2 x 640, corner-aligned
// stretch (the reference's NPP map, src/yolo_engine.cpp:186-190), 8-bit intermediate.
//   demosaic (camera ISP step, reference src/mv_camera.cpp:96) -> rot180 (nppiMirror :182-184) ->
//   bilinear resize to 640 x 640, u8 (nppiResize :186-190) -> /255 (nppiScale :192-194) ->
//   conv0 3x3 s2 3->16 + bias + SiLU (first layer of the TensorRT engine, :105)
// in one kernel; the 640 x 640 x 3 network input only ever exists as a shared-memory tile.
//
// Why a second stem kernel: the generic one (preprocess.cu) is instruction-issue bound
// (profiles/r1_summary.md section 4: 630 M warp instructions per 128 frames, 12 % of the HBM
// roofline).  Three properties of the camera case remove most of that work:
//   * x scale is exactly 2 and the map is corner-aligned, so the second x tap has weight 0 and a
//     network-input pixel needs ONE source column; every needed column has the same Bayer column
//     parity, so the colour-site case is uniform per source row.
//   * two neighbouring needed columns live in one 32-bit word of the raw row: the demosaic runs on
//     2 x 16-bit lanes packed in a register (sums of <= 4 bytes + rounding never carry between
//     lanes), one thread produces two network-input pixels.
//   * y scale = P/Q (1024/640 = 8/5): the vertical weight is k/Q, k integer, so
//     floor(a*(1-f) + b*f + 0.5) == (2(Q-k)*a + 2k*b + Q) / 2Q exactly (checked exhaustively over
//     all rows and byte pairs against the oracle's FP32 arithmetic: tests/test_oracle_cpu.py
//     test_stem_integer_lerp_equals_float_oracle).  The numerator (< 2Q*255 + Q + 1) indexes a
//     shared-memory table of the FP16 values of q/255: lerp + quantise + scale = 2 IMAD + 1 LDS.
// A CTA owns a full-width strip of OROWS conv0 output rows: staging is whole 1280-byte rows
// (16-byte cp.async, always aligned), the tile is [2*OROWS+1][641 px][3] halves, conv0 runs as
// mma.sync m16n8k16 with K ordered (ky, 3 px x 3 ch + 1 pad) so that every A fragment register is one
// aligned 32-bit shared-memory load.
#include "common.cuh"

namespace irmv {
namespace {

constexpr int SW = 1280;                 // source width handled by this kernel (2 * kNet)
constexpr int RP = 16 + SW + 16;         // staged raw row: [16 B apron | row | 16 B apron]
constexpr int PITCHW = 965;              // tile row pitch in 32-bit words (641 px * 6 B = 3846 B -> 962 words + skew)
constexpr int NTHREADS = 512;
constexpr int OWID = kNet / 2;           // conv0 output side, 320

struct StemBayerArgs {
  const uint8_t *src;                    // frames [..][H][1280]; used when src_indirect == null
  const uint8_t *const *src_indirect;
  int n, H, frame0;                      // frame0: first source frame of this launch
  int P, Q;                              // H / 640 in lowest terms
  int red_y, red_x;                      // position of the red sample in the 2x2 Bayer tile
  int nr_max;                            // staged-row capacity (rows of RP bytes)
  int lut_n;                             // 2*Q*255 + Q + 1
  const float *w, *bias;                 // conv0: [16][9 taps][3] FP32, [16]
  __half *out; long long out_ps;         // normal layout (may be null)
  __half *out2; long long out2_ps;       // parity-split twin (may be null)
};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

struct Px3 { uint32_t w, c, e; };        // west / centre / east neighbours of the two needed columns, 2 x 16-bit lanes

// Needed columns of unit j.  ROT (needed columns are odd): 4j+1 (low lane), 4j+3 (high lane);
// otherwise (even): 4j (low), 4j+2 (high).
template <bool ROT>
__device__ __forceinline__ Px3 unpack(const uint8_t *row, int j) {
  const uint32_t *wp = reinterpret_cast<const uint32_t *>(row);
  constexpr uint32_t M = 0x00ff00ffu;
  const uint32_t C = wp[j];
  Px3 o;
  if (ROT) {
    const uint32_t N = wp[j + 1];
    o.c = (C >> 8) & M;
    o.w = C & M;
    o.e = __funnelshift_r(C, N, 16) & M;
  } else {
    const uint32_t P = wp[j - 1];
    o.c = C & M;
    o.e = (C >> 8) & M;
    o.w = __funnelshift_r(P, C, 24) & M;
  }
  return o;
}

template <bool ROT, bool RED_COL, int OROWS>
__global__ void __launch_bounds__(NTHREADS) stem_bayer2x_kernel(const __grid_constant__ StemBayerArgs a) {
  constexpr int NIR = 2 * OROWS + 1;     // network-input rows of the strip (one halo row on top)
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t *tile_w = reinterpret_cast<uint32_t *>(smem);                       // [NIR][PITCHW]
  __half *lut = reinterpret_cast<__half *>(smem + (size_t)NIR * PITCHW * 4);   // [lut_n]
  uint8_t *raw = smem + (((size_t)NIR * PITCHW * 4 + (size_t)a.lut_n * 2 + 15) & ~(size_t)15);   // [nr_max][RP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = a.H, P = a.P, Q = a.Q;
  const int n = blockIdx.y, oy0 = blockIdx.x * OROWS, iy0 = 2 * oy0 - 1;
  const uint8_t *base = a.src_indirect ? *a.src_indirect : a.src;
  const uint8_t *frame = base + (size_t)(a.frame0 + n) * ((size_t)H * SW);

  // ---- source rows of the strip: virtual rows [vlo, vlo + nr), slot = v - vlo, content = row reflect101(v)
  const int ry_min = (max(iy0, 0) * P) / Q;
  const int ry_max = min(((iy0 + NIR - 1) * P) / Q + 1, H - 1);
  const int vlo = (ROT ? H - 1 - ry_max : ry_min) - 1;
  const int nr = ry_max - ry_min + 4;
  if (((size_t)frame & 15) == 0) {
    for (int i = tid; i < nr * (SW / 16); i += NTHREADS) {
      const int r = i / (SW / 16), c = i - r * (SW / 16);
      const int sy = min(reflect101(vlo + r, H), H - 1);
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(raw + (size_t)r * RP + 16 + c * 16);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(frame + (size_t)sy * SW + c * 16) : "memory");
    }
  } else {                                // caller's frames are not 16-byte aligned: plain byte copies
    for (int i = tid; i < nr * SW; i += NTHREADS) {
      const int r = i / SW, c = i - r * SW;
      raw[(size_t)r * RP + 16 + c] = frame[(size_t)min(reflect101(vlo + r, H), H - 1) * SW + c];
    }
  }
  if (tid < nr) {                         // mirrored columns -1 and W (reflect-101), read straight from global
    const uint8_t *g = frame + (size_t)min(reflect101(vlo + tid, H), H - 1) * SW;
    raw[(size_t)tid * RP + 15] = g[1];
    raw[(size_t)tid * RP + 16 + SW] = g[SW - 2];
  }
  // table: numerator of the vertical lerp -> FP16(q / 255), q = x / 2Q (the 8-bit intermediate of the reference)
  for (int x = tid; x < a.lut_n; x += NTHREADS) lut[x] = __float2half_rn(__fdiv_rn((float)(x / (2 * Q)), 255.0f));
  // conv padding of the tile: column 0 (ix = -1) and the pad half behind the last pixel
  if (tid < NIR) {
    uint16_t *row16 = reinterpret_cast<uint16_t *>(tile_w + (size_t)tid * PITCHW);
    row16[0] = row16[1] = row16[2] = 0;
    row16[3 * 641] = 0; row16[3 * 641 + 1] = 0;
  }
  // conv0 B fragments (constants): k = ky*10 + (kx*3 + c), entry 9 of every ky group and k >= 30 are zero
  const int g = lane >> 2, t = lane & 3;
  uint32_t bfrag[2][2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 16 * s2 + 2 * t + 8 * h + e, ky = k / 10, j = k - ky * 10;
          v[e] = (ky < 3 && j < 9) ? a.w[(nt * 8 + g) * 27 + ky * 9 + j] : 0.f;
        }
        const __half2 hv = __floats2half2_rn(v[0], v[1]);
        bfrag[nt][s2][h] = *reinterpret_cast<const uint32_t *>(&hv);
      }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  // ---- sampling: warp task = (input row r, 32 units); unit = two network-input pixels
  constexpr uint32_t M = 0x00ff00ffu;
  const uint32_t K1 = 0x00010001u, K2 = 0x00020002u;
  for (int task = warp; task < NIR * 10; task += NTHREADS / 32) {
    const int r = task / 10, j = (task - r * 10) * 32 + lane;
    const int iy = iy0 + r;
    // first / second pixel of the unit in ascending ix; tile column = ix + 1
    const int ix_first = ROT ? 638 - 2 * j : 2 * j;
    uint8_t *dst = reinterpret_cast<uint8_t *>(tile_w + (size_t)r * PITCHW) + 6 * (ix_first + 1);
    if (iy < 0) {                         // conv padding row above the image
      *reinterpret_cast<uint16_t *>(dst) = 0;
      *reinterpret_cast<uint32_t *>(dst + 2) = 0;
      *reinterpret_cast<uint32_t *>(dst + 6) = 0;
      *reinterpret_cast<uint16_t *>(dst + 10) = 0;
      continue;
    }
    const int tt = iy * P, i0 = tt / Q, k = tt - i0 * Q, i1 = min(i0 + 1, H - 1);
    int wA = 2 * (Q - k), wB = 2 * k;
    if (i1 == i0) { wA += wB; wB = 0; }
    const int sy0 = ROT ? H - 1 - i0 : i0, sy1 = ROT ? H - 1 - i1 : i1;
    const int lo = min(sy0, sy1);
    const uint32_t w_lo = (uint32_t)(sy0 <= sy1 ? wA : wB), w_hi = (uint32_t)(sy0 <= sy1 ? wB : wA);
    // a "site" row holds the red or blue sample at the needed columns, a "green" row the green one
    const bool lo_site = (((lo & 1) == a.red_y) == RED_COL);
    const uint8_t *rp = raw + (size_t)(lo - 1 - vlo) * RP + 16;
    const Px3 r0 = unpack<ROT>(rp, j), r1 = unpack<ROT>(rp + RP, j), r2 = unpack<ROT>(rp + 2 * RP, j),
              r3 = unpack<ROT>(rp + 3 * RP, j);
    uint32_t cross, diag, cS, horiz, vert, cG, wS, wG;
    if (lo_site) {                        // site row = lo (r0 r1 r2), green row = lo + 1 (r1 r2 r3)
      cross = ((r0.c + r2.c + r1.w + r1.e + K2) >> 2) & M;
      diag = ((r0.w + r0.e + r2.w + r2.e + K2) >> 2) & M;
      cS = r1.c;
      horiz = ((r2.w + r2.e + K1) >> 1) & M;
      vert = ((r1.c + r3.c + K1) >> 1) & M;
      cG = r2.c;
      wS = w_lo; wG = w_hi;
    } else {                              // green row = lo, site row = lo + 1
      horiz = ((r1.w + r1.e + K1) >> 1) & M;
      vert = ((r0.c + r2.c + K1) >> 1) & M;
      cG = r1.c;
      cross = ((r1.c + r3.c + r2.w + r2.e + K2) >> 2) & M;
      diag = ((r1.w + r1.e + r3.w + r3.e + K2) >> 2) & M;
      cS = r2.c;
      wG = w_lo; wS = w_hi;
    }
    // site row: own colour at the centre, green = cross, the other colour = diagonal; green row on a
    // red row (RED_COL == false): R horizontal, B vertical; on a blue row: R vertical, B horizontal
    const uint32_t RS = RED_COL ? cS : diag, BS = RED_COL ? diag : cS;
    const uint32_t RG = RED_COL ? vert : horiz, BG = RED_COL ? horiz : vert;
    const uint32_t QQ = (uint32_t)Q * 0x00010001u;     // rounding term in both lanes
    const uint32_t xr = RS * wS + RG * wG + QQ;
    const uint32_t xg = cross * wS + cG * wG + QQ;
    const uint32_t xb = BS * wS + BG * wG + QQ;
    const uint16_t *lut16 = reinterpret_cast<const uint16_t *>(lut);
    const uint32_t r_l = lut16[xr & 0xffffu], r_h = lut16[xr >> 16];
    const uint32_t g_l = lut16[xg & 0xffffu], g_h = lut16[xg >> 16];
    const uint32_t b_l = lut16[xb & 0xffffu], b_h = lut16[xb >> 16];
    // ROT: the high lane (column 4j+3) is the smaller ix
    const uint32_t R1 = ROT ? r_h : r_l, G1 = ROT ? g_h : g_l, B1 = ROT ? b_h : b_l;
    const uint32_t R2 = ROT ? r_l : r_h, G2 = ROT ? g_l : g_h, B2 = ROT ? b_l : b_h;
    *reinterpret_cast<uint16_t *>(dst) = (uint16_t)R1;
    *reinterpret_cast<uint32_t *>(dst + 2) = G1 | (B1 << 16);
    *reinterpret_cast<uint32_t *>(dst + 6) = R2 | (G2 << 16);
    *reinterpret_cast<uint16_t *>(dst + 10) = (uint16_t)B2;
  }
  __syncthreads();

  // ---- conv0: implicit GEMM on mma.sync m16n8k16 (FP16 operands, FP32 accumulate).  An m-tile is 16
  // consecutive pixels of one output row; A[px][k] = tile[2*orow + ky][2*px + kx][c] = the 10 halves
  // starting at word 3*px of tile row 2*orow + ky.
  int off[2][2];
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kw = 8 * s2 + 4 * h + t, ky = kw / 5, tp = kw - ky * 5;
      off[s2][h] = ky < 3 ? ky * PITCHW + tp : 0;
    }
  float bias_r[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) { bias_r[nt][0] = a.bias[nt * 8 + 2 * t]; bias_r[nt][1] = a.bias[nt * 8 + 2 * t + 1]; }
  for (int mt = warp; mt < OROWS * (OWID / 16); mt += NTHREADS / 32) {
    const int orow = mt / (OWID / 16), xb = (mt - orow * (OWID / 16)) * 16;
    const uint32_t *tw = tile_w + (size_t)(2 * orow) * PITCHW + 3 * (xb + g);
    float acc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      acc[nt][0] = acc[nt][2] = bias_r[nt][0];
      acc[nt][1] = acc[nt][3] = bias_r[nt][1];
    }
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      const uint32_t a0 = tw[off[s2][0]], a1 = tw[24 + off[s2][0]], a2 = tw[off[s2][1]], a3 = tw[24 + off[s2][1]];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bfrag[nt][s2][0]), "r"(bfrag[nt][s2][1]));
    }
    // lane holds channels {2t, 2t+1} of plane 0 (nt = 0) and plane 1 for pixels xb + g and xb + g + 8
    const int y = oy0 + orow;
    const long long rb = pr_index(n, y, 0, OWID, OWID), rb2 = pr_index(n, y >> 1, 0, OWID / 2, OWID / 2);
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int x = xb + g + 8 * hr;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const __half2 hv = __floats2half2_rn(silu_fast(acc[nt][2 * hr]), silu_fast(acc[nt][2 * hr + 1]));
        if (a.out) {
          const long long pix = rb + x;
          *reinterpret_cast<__half2 *>(a.out + (long long)nt * a.out_ps + pix * 8 + 2 * t) = hv;
        }
        if (a.out2) {                     // parity-split twin for the stride-2 consumer (common.cuh, ConvParams)
          const long long pix2 = rb2 + (x >> 1);
          *reinterpret_cast<__half2 *>(a.out2 + (long long)(((y & 1) * 2 + (x & 1)) * 2 + nt) * a.out2_ps + pix2 * 8 + 2 * t) = hv;
        }
      }
    }
  }
}

int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

template <bool ROT, bool RED_COL, int OROWS>
cudaError_t launch_t(const StemBayerArgs &a, size_t smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(stem_bayer2x_kernel<ROT, RED_COL, OROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid(OWID / OROWS, a.n);
  stem_bayer2x_kernel<ROT, RED_COL, OROWS><<<grid, NTHREADS, smem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace

// The fast path applies to: Bayer source, width 1280, height >= 640 with H/640 = P/Q, Q <= 16,
// reference resize (corner-aligned stretch) and the 8-bit intermediate.
bool stem_bayer2x_applies(const PreprocessParams &p) {
  if (p.chan_order < 2 || p.src_w != SW || p.resize_mode != 0 || !p.quantize_u8 || p.src_h < kNet || p.src_h > 4096) return false;
  const int g = gcd_i(p.src_h, kNet);
  return kNet / g <= 16;
}

cudaError_t launch_stem_bayer2x(const PreprocessParams &p, int frame0, const float *w, const float *bias, __half *out,
                                long long out_ps, __half *out2, long long out2_ps, cudaStream_t s) {
  constexpr int OROWS = 8;
  StemBayerArgs a{};
  a.src = p.src; a.src_indirect = p.src_indirect; a.n = p.n; a.H = p.src_h; a.frame0 = frame0;
  const int g = gcd_i(p.src_h, kNet);
  a.P = p.src_h / g; a.Q = kNet / g;
  a.red_y = (p.chan_order == 3 || p.chan_order == 5) ? 1 : 0;      // BGGR, GBRG: red on odd rows
  a.red_x = (p.chan_order == 3 || p.chan_order == 4) ? 1 : 0;      // BGGR, GRBG: red on odd columns
  a.nr_max = (2 * OROWS * a.P + a.Q - 1) / a.Q + 6;
  a.lut_n = 2 * a.Q * 255 + a.Q + 1;
  a.w = w; a.bias = bias; a.out = out; a.out_ps = out_ps; a.out2 = out2; a.out2_ps = out2_ps;
  const size_t smem = (((size_t)(2 * OROWS + 1) * PITCHW * 4 + (size_t)a.lut_n * 2 + 15) & ~(size_t)15) + (size_t)a.nr_max * RP;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  // needed columns: W - 1 - 2*ix (odd) with rot180, 2*ix (even) without
  const bool rot = p.rotate180 != 0;
  const bool red_col = ((rot ? 1 : 0) == a.red_x);
  if (rot) return red_col ? launch_t<true, true, OROWS>(a, smem, s) : launch_t<true, false, OROWS>(a, smem, s);
  return red_col ? launch_t<false, true, OROWS>(a, smem, s) : launch_t<false, false, OROWS>(a, smem, s);
}

}  // namespace irmv

// CUDA-core direct convolution (bring-up / cross-check kernel, IRMV_CONV_DIRECT).
// Same ConvParams contract as the tcgen05 kernel: NHWC FP16 in, concat-on-read over up to two
// segments, nearest-2x upsample-on-read, bias + SiLU + residual epilogue, channel-slice store.
// Not the production path: it exists so that a wrong tensor-core tile can be localised on the
// GPU against an independent kernel, layer by layer.
#include "common.cuh"

namespace irmv {
namespace {

__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

// one thread: one output pixel x 8 output channels
__global__ void __launch_bounds__(128) conv_direct_kernel(ConvParams p) {
  const int M = p.B * p.OH * p.OW;
  const int ngroups = p.cout / 8;
  long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // pixel fastest so a warp shares one weight group (broadcast loads)
  int m = (int)(gid % M);
  int g = (int)(gid / M);
  if (g >= ngroups) return;
  const int ox = m % p.OW, oy = (m / p.OW) % p.OH, b = m / (p.OW * p.OH);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = p.bias[g * 8 + j];

  for (int ky = 0; ky < p.k; ++ky) {
    int iy = oy * p.stride - p.pad + ky;
    if (iy < 0 || iy >= p.H) continue;
    for (int kx = 0; kx < p.k; ++kx) {
      int ix = ox * p.stride - p.pad + kx;
      if (ix < 0 || ix >= p.W) continue;
      int kbase = (ky * p.k + kx) * p.cin;
      int cdone = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const ConvSeg sg = p.seg[s];
        int hs = sg.up ? p.H >> 1 : p.H, ws = sg.up ? p.W >> 1 : p.W;
        int sy = sg.up ? iy >> 1 : iy, sx = sg.up ? ix >> 1 : ix;
        const __half *ip = sg.ptr + (size_t)pr_index(b, sy, sx, hs, ws) * 8;
        for (int c = 0; c < sg.c; c += 8) {
          uint4 iv = *reinterpret_cast<const uint4 *>(ip + (size_t)(c >> 3) * sg.pstride);
          const __half2 *ih = reinterpret_cast<const __half2 *>(&iv);
          float xin[8];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float2 f = __half22float2(ih[t]);
            xin[2 * t] = f.x; xin[2 * t + 1] = f.y;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const __half *wp = p.w_plain + (size_t)(g * 8 + j) * p.kpad + kbase + cdone + c;
            uint4 wv = __ldg(reinterpret_cast<const uint4 *>(wp));
            const __half2 *wh = reinterpret_cast<const __half2 *>(&wv);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float2 f = __half22float2(wh[t]);
              acc[j] = fmaf(xin[2 * t], f.x, acc[j]);
              acc[j] = fmaf(xin[2 * t + 1], f.y, acc[j]);
            }
          }
        }
        cdone += sg.c;
      }
    }
  }
  __half outv[8];
  float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const size_t opix = (size_t)pr_index(b, oy, ox, p.OH, p.OW);
  if (p.res) {
    uint4 rv = *reinterpret_cast<const uint4 *>(p.res + (size_t)g * p.res_pstride + opix * 8);
    const __half2 *rh = reinterpret_cast<const __half2 *>(&rv);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 f = __half22float2(rh[t]);
      r[2 * t] = f.x; r[2 * t + 1] = f.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = p.act ? silu(acc[j]) : acc[j];
    outv[j] = __float2half_rn(v + r[j]);
  }
  *reinterpret_cast<uint4 *>(p.out + (size_t)g * p.out_pstride + opix * 8) =
      *reinterpret_cast<uint4 *>(outv);
}

}  // namespace

cudaError_t launch_conv_direct(const ConvParams &p, cudaStream_t s) {
  long long total = (long long)p.B * p.OH * p.OW * (p.cout / 8);
  int blocks = (int)((total + 127) / 128);
  conv_direct_kernel<<<blocks, 128, 0, s>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------- SPPF pooling
namespace {
// Three chained MaxPool2d(5,1,2) == windows of radius 2, 4, 6 clipped to the image (padding is
// -inf).  One thread: one pixel x 8 channels; separable max would cut reads further, but the
// whole tensor is 20x20x128 per frame (0.1 MB) and lives in L2.
__global__ void __launch_bounds__(128) sppf_pool_kernel(__half *buf, int B, int H, int W,
                                                        long long pstride, int c) {
  int groups = c / 8;
  long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)B * H * W * groups;
  if (gid >= total) return;
  int g = (int)(gid % groups);
  long long pix = gid / groups;
  int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
  __half2 m1[4], m2[4], m3[4];
  const __half2 ninf = __float2half2_rn(-65504.0f);
#pragma unroll
  for (int t = 0; t < 4; ++t) m1[t] = m2[t] = m3[t] = ninf;
  for (int dy = -6; dy <= 6; ++dy) {
    int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -6; dx <= 6; ++dx) {
      int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      uint4 v = *reinterpret_cast<const uint4 *>(buf + (size_t)g * pstride + (size_t)pr_index(b, yy, xx, H, W) * 8);
      const __half2 *h = reinterpret_cast<const __half2 *>(&v);
      int r = max(abs(dy), abs(dx));
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        m3[t] = __hmax2(m3[t], h[t]);
        if (r <= 4) m2[t] = __hmax2(m2[t], h[t]);
        if (r <= 2) m1[t] = __hmax2(m1[t], h[t]);
      }
    }
  }
  __half *o = buf + (size_t)g * pstride + (size_t)pr_index(b, y, x, H, W) * 8;
  const size_t grp = (size_t)(c / 8) * pstride;
  *reinterpret_cast<uint4 *>(o + grp) = *reinterpret_cast<uint4 *>(m1);
  *reinterpret_cast<uint4 *>(o + 2 * grp) = *reinterpret_cast<uint4 *>(m2);
  *reinterpret_cast<uint4 *>(o + 3 * grp) = *reinterpret_cast<uint4 *>(m3);
}
}  // namespace

cudaError_t launch_sppf_pool(__half *buf, int B, int H, int W, long long pstride, int c, cudaStream_t s) {
  long long total = (long long)B * H * W * (c / 8);
  int blocks = (int)((total + 127) / 128);
  sppf_pool_kernel<<<blocks, 128, 0, s>>>(buf, B, H, W, pstride, c);
  return cudaGetLastError();
}

}  // namespace irmv

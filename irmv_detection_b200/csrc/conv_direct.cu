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
// Three chained MaxPool2d(5,1,2) (ultralytics SPPF; padding is -inf, i.e. windows clipped to the
// image).  One CTA per (image, plane of 8 channels): the 20x20 plane is staged in shared memory
// and each 5x5 max is done separably (row max, then column max), three times in a row.
constexpr int POOL_MAX_HW = 20 * 20;
__global__ void __launch_bounds__(256) sppf_pool_kernel(__half *buf, int H, int W, long long pstride, int c) {
  __shared__ uint4 cur[POOL_MAX_HW], tmp[POOL_MAX_HW];
  const int planes = c / 8;
  const int b = blockIdx.x / planes, g = blockIdx.x - b * planes;
  const int n = H * W;
  __half *plane = buf + (size_t)g * pstride;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    cur[i] = *reinterpret_cast<const uint4 *>(plane + (size_t)pr_index(b, i / W, i % W, H, W) * 8);
  __syncthreads();
  for (int pass = 1; pass <= 3; ++pass) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {          // row max over x-2..x+2
      const int y = i / W, x = i - y * W;
      uint4 m = cur[i];
      __half2 *mh = reinterpret_cast<__half2 *>(&m);
      for (int dx = -2; dx <= 2; ++dx) {
        const int xx = x + dx;
        if (dx == 0 || xx < 0 || xx >= W) continue;
        const uint4 v = cur[y * W + xx];
        const __half2 *vh = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
        for (int t = 0; t < 4; ++t) mh[t] = __hmax2(mh[t], vh[t]);
      }
      tmp[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {          // column max over y-2..y+2
      const int y = i / W, x = i - y * W;
      uint4 m = tmp[i];
      __half2 *mh = reinterpret_cast<__half2 *>(&m);
      for (int dy = -2; dy <= 2; ++dy) {
        const int yy = y + dy;
        if (dy == 0 || yy < 0 || yy >= H) continue;
        const uint4 v = tmp[yy * W + x];
        const __half2 *vh = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
        for (int t = 0; t < 4; ++t) mh[t] = __hmax2(mh[t], vh[t]);
      }
      *reinterpret_cast<uint4 *>(plane + (size_t)pass * planes * pstride + (size_t)pr_index(b, y, x, H, W) * 8) = m;
      cur[i] = m;   // each thread rewrites only its own pixels; readers of cur[] are behind the next barrier
    }
    __syncthreads();
  }
}
}  // namespace

cudaError_t launch_sppf_pool(__half *buf, int B, int H, int W, long long pstride, int c, cudaStream_t s) {
  if (H * W > POOL_MAX_HW) return cudaErrorInvalidValue;
  sppf_pool_kernel<<<B * (c / 8), 256, 0, s>>>(buf, H, W, pstride, c);
  return cudaGetLastError();
}

}  // namespace irmv

// Depthwise 3x3 convolution (stride 1 or 2, pad 1) + bias over the channel-blocked planar layout,
// sm_100a.  The ShuffleNetV2-backbone keypoint detector (BASELINE.json configs[2]; the reference lists
// the model in its benchmark table, README.md:12,16) is the only user: its units are
// 1x1 -> depthwise 3x3 -> 1x1, and the 1x1 convolutions run on the tcgen05 raster kernel.
//
// Bandwidth bound: 9 MACs per output value against 4 bytes moved (2 in + 2 out, FP16), so the kernel
// is written for the memory system.  A plane is [pixel][8 channels], i.e. a depthwise conv of one
// plane touches exactly one 16-byte vector per pixel; the zero-padded raster (common.cuh) makes the
// 3x3 neighbourhood nine unconditional 16-byte loads (the left / top / right / bottom neighbours of
// border pixels are the layout's zero pixels), consecutive threads read consecutive pixels (512
// contiguous bytes per warp and tap), and the 9x reuse is served by L1.  FP32 accumulation like every
// other layer; output FP16.  A thread produces one pixel of one plane.
#include "common.cuh"

namespace irmv {
namespace {

__global__ void __launch_bounds__(256) dwconv3x3_kernel(const DwParams p) {
  __shared__ float sw[9][8];
  __shared__ float sb[8];
  const int plane = p.rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  if (threadIdx.x < 72) sw[threadIdx.x >> 3][threadIdx.x & 7] = __half2float(p.w[(size_t)plane * 72 + threadIdx.x]);
  if (threadIdx.x < 8) sb[threadIdx.x] = p.bias[plane * 8 + threadIdx.x];
  __syncthreads();
  const int OH = p.H / p.stride, OW = p.W / p.stride;
  const int total = p.B * OH * OW;
  const unsigned blk = p.rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int i = (int)(blk * blockDim.x + threadIdx.x);
  if (i >= total) return;
  const int ox = i % OW, row = i / OW, oy = row % OH, b = row / OH;
  const int Wp = p.W + 1;
  // centre input pixel of the window and its raster index
  const int iy = oy * p.stride, ix = ox * p.stride;
  const __half *in = p.in + (long long)plane * p.in_ps + pr_index(b, iy, ix, p.H, p.W) * 8;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = sb[c];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + ((ky - 1) * Wp + (kx - 1)) * 8));
      const __half2 *h = reinterpret_cast<const __half2 *>(&v);
      const float *w = sw[ky * 3 + kx];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float2 f = __half22float2(h[c]);
        acc[2 * c] = fmaf(f.x, w[2 * c], acc[2 * c]);
        acc[2 * c + 1] = fmaf(f.y, w[2 * c + 1], acc[2 * c + 1]);
      }
    }
  __half2 o[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) o[c] = __floats2half2_rn(acc[2 * c], acc[2 * c + 1]);
  *reinterpret_cast<uint4 *>(p.out + (long long)plane * p.out_ps + pr_index(b, oy, ox, OH, OW) * 8) = *reinterpret_cast<uint4 *>(o);
}

}  // namespace

cudaError_t launch_dwconv3x3(const DwParams &p, cudaStream_t s) {
  if (p.planes <= 0 || p.B <= 0) return cudaSuccess;
  if (!(p.stride == 1 || p.stride == 2) || (p.H % p.stride) || (p.W % p.stride)) return cudaErrorInvalidValue;
  const long long total = (long long)p.B * (p.H / p.stride) * (p.W / p.stride);
  dim3 grid((unsigned)((total + 255) / 256), (unsigned)p.planes);
  dwconv3x3_kernel<<<grid, 256, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace irmv

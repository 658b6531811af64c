// DFL box decode + class-score filter + top-k + bitmask NMS for sm_100a.
// Replaces the EfficientNMS_TRT plugin the reference runs inside its TensorRT engine
// (reference src/yolo_engine.cpp:33,53-57; README.md:25).  Contract = oracle/nms_ref.py:
//   candidates (anchor a, class c) with score > thr, ordered by (score desc, a*nc+c asc),
//   first kMaxCand enter NMS, same-class IoU > iou_thr suppresses, stop at max_det.
// IoU uses explicitly rounded FP32 ops so kept indices are bit-exact against the oracle.
#include <math_constants.h>

#include "common.cuh"

namespace irmv {
namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ unsigned long long make_key(float score, uint32_t flat) {
  // positive floats order like their bit patterns; ties -> lower flat index first
  return ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
}

// ---- decode: 4 lanes per anchor (one box side each, 16 DFL bins = 32 bytes per lane) --------
// A quad handles DEC_K anchors (64 apart, so a warp-level load still covers 8 consecutive anchors): the DEC_K
// class-logit loads are issued back to back before any of them is used.  The kernel is latency bound (a few
// loads and a compare per anchor), so the loads in flight per thread, not the bytes, set its duration.
constexpr int DEC_K = 4;
struct AnchorPos { int scale, hw, gy, gx; float stride; size_t pix; };
__device__ __forceinline__ AnchorPos anchor_pos(int a, int b, int padded) {
  AnchorPos p;
  int local;
  if (a < 6400) { p.scale = 0; local = a; p.hw = 80; p.stride = 8.0f; p.gy = local / 80; }
  else if (a < 8000) { p.scale = 1; local = a - 6400; p.hw = 40; p.stride = 16.0f; p.gy = local / 40; }
  else { p.scale = 2; local = a - 8000; p.hw = 20; p.stride = 32.0f; p.gy = local / 20; }
  p.gx = local - p.gy * p.hw;
  p.pix = padded ? (size_t)pr_index(b, p.gy, p.gx, p.hw, p.hw) : (size_t)b * p.hw * p.hw + local;
  return p;
}

__global__ void __launch_bounds__(256) decode_kernel(HeadPtrs h, int B, int nc, float score_thr,
                                                     NmsScratch sc, float *scores_out) {
  // grid.y = frame, 64 * DEC_K anchors per CTA: no 64-bit or variable-divisor divisions on the index path
  const int lane4 = threadIdx.x & 3;
  const int b = blockIdx.y;
  const int a0 = blockIdx.x * (64 * DEC_K) + (threadIdx.x >> 2);
  const int qbase = (threadIdx.x & 31) & ~3;
  const unsigned qmask = 0xFu << qbase;
  const float lthr = !(score_thr > 0.f) ? -CUDART_INF_F : (score_thr >= 1.f ? CUDART_INF_F : __logf(score_thr / (1.f - score_thr)) - 0.05f);

  // class logits first: lane handles classes lane4*4 .. +3.  A logit clearly below the threshold's
  // logit cannot pass (guard band 0.05 in logit space = 0.009 in score, far above the rounding of
  // the sigmoid below), and an anchor without a candidate needs no box: the quad skips the
  // DFL, so the 64 box logits of almost every anchor are never read.
  uint2 cv[DEC_K];
#pragma unroll
  for (int k = 0; k < DEC_K; ++k) {
    const int a = a0 + 64 * k;
    cv[k] = make_uint2(0xfc00fc00u, 0xfc00fc00u);          // -inf logits: anchors past the end never pass
    if (a < kNumAnchors) {
      const AnchorPos ap = anchor_pos(a, b, h.padded);
      const __half *cp = h.padded ? h.cls[ap.scale] + (size_t)(lane4 >> 1) * h.cls_ps[ap.scale] + ap.pix * 8 + (lane4 & 1) * 4
                                  : h.cls[ap.scale] + ap.pix * kClsPad + lane4 * 4;
      cv[k] = __ldg(reinterpret_cast<const uint2 *>(cp));
    }
  }
#pragma unroll
  for (int k = 0; k < DEC_K; ++k) {
    const int a = a0 + 64 * k;
    if (a >= kNumAnchors) break;                            // whole quads leave together
    float lg[4];
    {
      const __half2 *ch = reinterpret_cast<const __half2 *>(&cv[k]);
      float2 f = __half22float2(ch[0]); lg[0] = f.x; lg[1] = f.y;
      float2 g = __half22float2(ch[1]); lg[2] = g.x; lg[3] = g.y;
    }
    bool maybe = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) maybe |= (lane4 * 4 + j < nc) && lg[j] > lthr;
    if (!scores_out && __ballot_sync(qmask, maybe) == 0) continue;
    const AnchorPos ap = anchor_pos(a, b, h.padded);
    const int scale = ap.scale, gx = ap.gx, gy = ap.gy;
    const float stride = ap.stride;
    const size_t pix = ap.pix;

    // DFL expectation of this lane's side (16 bins = planes 2*side, 2*side+1 in the planar layout)
    uint4 v0, v1;
    if (h.padded) {
      const __half *bp = h.box[scale] + (size_t)(2 * lane4) * h.box_ps[scale] + pix * 8;
      v0 = __ldg(reinterpret_cast<const uint4 *>(bp));
      v1 = __ldg(reinterpret_cast<const uint4 *>(bp + h.box_ps[scale]));
    } else {
      const uint4 *bp = reinterpret_cast<const uint4 *>(h.box[scale] + pix * 64 + lane4 * 16);
      v0 = __ldg(bp); v1 = __ldg(bp + 1);
    }
    float x[16];
    {
      const __half2 *p0 = reinterpret_cast<const __half2 *>(&v0);
      const __half2 *p1 = reinterpret_cast<const __half2 *>(&v1);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = __half22float2(p0[t]); x[2 * t] = f.x; x[2 * t + 1] = f.y;
        float2 g = __half22float2(p1[t]); x[8 + 2 * t] = g.x; x[8 + 2 * t + 1] = g.y;
      }
    }
    float mx = x[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) mx = fmaxf(mx, x[i]);
    float se = 0.f, sw = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float e = __expf(x[i] - mx);     // ex2.approx: 2 ulp, the expectation moves by < 1e-4 px
      se += e;
      sw += e * (float)i;
    }
    float dist = sw / se;
    // gather l,t,r,b in every lane of the quad
    float dl = __shfl_sync(qmask, dist, qbase + 0);
    float dt = __shfl_sync(qmask, dist, qbase + 1);
    float dr = __shfl_sync(qmask, dist, qbase + 2);
    float db = __shfl_sync(qmask, dist, qbase + 3);
    float ax = (float)gx + 0.5f, ay = (float)gy + 0.5f;
    if (lane4 == 0) {
      float4 bx;
      bx.x = (ax - dl) * stride; bx.y = (ay - dt) * stride;
      bx.z = (ax + dr) * stride; bx.w = (ay + db) * stride;
      *reinterpret_cast<float4 *>(sc.boxes + ((size_t)b * kNumAnchors + a) * 4) = bx;
    }
    // class scores
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = lane4 * 4 + j;
      if (c >= nc) break;
      if (!scores_out && !(lg[j] > lthr)) continue;
      float s = sigmoidf_(lg[j]);
      if (scores_out) scores_out[((size_t)b * kNumAnchors + a) * nc + c] = s;
      if (s > score_thr) {
        uint32_t flat = (uint32_t)(a * nc + c);
        int slot = atomicAdd(sc.counts + b, 1);
        sc.keys[(size_t)b * kNumAnchors * nc + slot] = make_key(s, flat);
      }
    }
  }
}

// ---- candidate filter from FP32 scores (isolated NMS parity path) ---------------------------
__global__ void __launch_bounds__(256) score_filter_kernel(const float *scores, int B, int A, int nc,
                                                           float score_thr, NmsScratch sc) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long per = (long long)A * nc;
  if (i >= per * B) return;
  int b = (int)(i / per);
  uint32_t flat = (uint32_t)(i - (long long)b * per);
  float s = scores[i];
  if (s > score_thr) {
    int slot = atomicAdd(sc.counts + b, 1);
    sc.keys[(size_t)b * per + slot] = make_key(s, flat);
  }
}

// ---- NMS: one CTA per frame -----------------------------------------------------------------
constexpr int NMS_T = 1024;   // one CTA per frame (fewer frames than SMs per replay): the whole SM works on the frame

__device__ __forceinline__ float iou_rn(const float4 &a, const float4 &b) {
  float ix1 = fmaxf(a.x, b.x), iy1 = fmaxf(a.y, b.y);
  float ix2 = fminf(a.z, b.z), iy2 = fminf(a.w, b.w);
  float iw = fmaxf(__fsub_rn(ix2, ix1), 0.f), ih = fmaxf(__fsub_rn(iy2, iy1), 0.f);
  float inter = __fmul_rn(iw, ih);
  float aa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  float ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  float uni = __fsub_rn(__fadd_rn(aa, ab), inter);
  if (!(uni > 0.f)) return 0.f;
  return __fdiv_rn(inter, uni);
}

__global__ void __launch_bounds__(NMS_T) nms_kernel(NmsScratch sc, int A, int nc, float iou_thr,
                                                    int max_det, DetOut out) {
  extern __shared__ __align__(16) uint8_t nms_smem[];
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(nms_smem);     // [kMaxCand]
  float4 *cbox = reinterpret_cast<float4 *>(nms_smem + (size_t)kMaxCand * 8);      // [kMaxCand]
  uint32_t *alive = reinterpret_cast<uint32_t *>(nms_smem + (size_t)kMaxCand * 24); // [kMaxCand/32]
  uint16_t *ccls = reinterpret_cast<uint16_t *>(nms_smem + (size_t)kMaxCand * 24 + (kMaxCand / 32) * 4);   // [kMaxCand] class of each candidate
  __shared__ int s_cnt, s_cur;
  __shared__ unsigned long long s_pivot;
  __shared__ int s_red[NMS_T / 32];

  const int b = blockIdx.x, tid = threadIdx.x;
  const size_t per = (size_t)A * nc;
  const unsigned long long *gk = sc.keys + (size_t)b * per;
  int total = sc.counts[b];
  int n = min(total, kMaxCand);

  if (total > kMaxCand) {
    // exact top-kMaxCand: bitwise search for the pivot key T with count(key >= T) == kMaxCand
    unsigned long long prefix = 0ull;
    for (int bit = 63; bit >= 0; --bit) {
      unsigned long long trial = prefix | (1ull << bit);
      int c = 0;
      for (int i = tid; i < total; i += NMS_T) c += (gk[i] >= trial) ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if ((tid & 31) == 0) s_red[tid >> 5] = c;
      __syncthreads();
      if (tid == 0) {
        int t = 0;
        for (int w = 0; w < NMS_T / 32; ++w) t += s_red[w];
        s_cnt = t;
      }
      __syncthreads();
      if (s_cnt >= kMaxCand) prefix = trial;   // pivot can be at least `trial`
      __syncthreads();
    }
    if (tid == 0) { s_pivot = prefix; s_cur = 0; }
    __syncthreads();
    for (int i = tid; i < total; i += NMS_T) {
      unsigned long long k = gk[i];
      if (k >= s_pivot) {
        int slot = atomicAdd(&s_cur, 1);
        if (slot < kMaxCand) keys[slot] = k;
      }
    }
    __syncthreads();
  } else {
    for (int i = tid; i < n; i += NMS_T) keys[i] = gk[i];
  }
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = n + tid; i < np2; i += NMS_T) keys[i] = 0ull;
  __syncthreads();

  // bitonic sort, descending
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += NMS_T) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long x = keys[i], y = keys[ixj];
          bool desc = ((i & k) == 0);
          if (desc ? (x < y) : (x > y)) { keys[i] = y; keys[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }

  const float4 *gbox = reinterpret_cast<const float4 *>(sc.boxes) + (size_t)b * A;
  for (int i = tid; i < n; i += NMS_T) {
    uint32_t flat = 0xFFFFFFFFu - (uint32_t)(keys[i] & 0xFFFFFFFFull);
    const uint32_t anchor = flat / nc;
    cbox[i] = gbox[anchor];
    ccls[i] = (uint16_t)(flat - anchor * nc);
  }
  for (int w = tid; w < (kMaxCand >> 5); w += NMS_T) {
    int lo = w << 5;
    alive[w] = (lo + 32 <= n) ? 0xFFFFFFFFu : (lo >= n ? 0u : ((1u << (n - lo)) - 1u));
  }
  __syncthreads();

  // Greedy pass.  Every warp finds the next alive candidate by itself (the scan only reads `alive`, and the
  // suppression pass of the same round never clears the bit it is looking for: it touches j > i only), so a
  // round costs ONE block barrier -- after the suppression pass -- instead of three (scan result broadcast
  // through shared memory + counters).
  const int nwords = (n + 31) >> 5;
  const int lane = tid & 31;
  int cur = 0, kept = 0;
  while (true) {
    int found = -1;
    for (int w0 = (cur >> 5); w0 < nwords && found < 0; w0 += 32) {
      int w = w0 + lane;
      uint32_t bits = (w < nwords) ? alive[w] : 0u;
      if (w == (cur >> 5)) bits &= ~((1u << (cur & 31)) - 1u);
      unsigned m = __ballot_sync(0xffffffffu, bits != 0u);
      if (m) {
        int src = __ffs(m) - 1;
        uint32_t bb = __shfl_sync(0xffffffffu, bits, src);
        found = ((w0 + src) << 5) + (__ffs(bb) - 1);
      }
    }
    const int i = found;
    if (i < 0) break;
    const unsigned long long ki = keys[i];
    const uint32_t flat_i = 0xFFFFFFFFu - (uint32_t)(ki & 0xFFFFFFFFull);
    const int cls_i = flat_i % nc;
    const float4 bi = cbox[i];
    if (tid == 0) {
      size_t o = (size_t)b * max_det + kept;
      reinterpret_cast<float4 *>(out.boxes)[o] = bi;
      out.scores[o] = __uint_as_float((uint32_t)(ki >> 32));
      out.classes[o] = cls_i;
      out.index[o] = (int32_t)flat_i;
    }
    ++kept;
    if (kept >= max_det) break;
    // suppress later same-class candidates
    for (int j = i + 1 + tid; j < n; j += NMS_T) {
      if ((int)ccls[j] != cls_i || !((alive[j >> 5] >> (j & 31)) & 1u)) continue;
      if (iou_rn(bi, cbox[j]) > iou_thr) atomicAnd(&alive[j >> 5], ~(1u << (j & 31)));
    }
    cur = i + 1;
    __syncthreads();
  }
  if (tid == 0) {
    out.num_dets[b] = kept;
    sc.counts[b] = 0;   // re-arm the candidate counter for the next replay
  }
}

}  // namespace

cudaError_t launch_decode(const HeadPtrs &h, int B, int nc, float score_thr, NmsScratch sc,
                          float *scores_or_null, cudaStream_t s) {
  if (B < 1) return cudaSuccess;
  dim3 grid((kNumAnchors + 64 * DEC_K - 1) / (64 * DEC_K), B);
  decode_kernel<<<grid, 256, 0, s>>>(h, B, nc, score_thr, sc, scores_or_null);
  return cudaGetLastError();
}

cudaError_t launch_score_filter(const float *scores, int B, int A, int nc, float score_thr,
                                NmsScratch sc, cudaStream_t s) {
  long long total = (long long)B * A * nc;
  int blocks = (int)((total + 255) / 256);
  score_filter_kernel<<<blocks, 256, 0, s>>>(scores, B, A, nc, score_thr, sc);
  return cudaGetLastError();
}

cudaError_t launch_nms(NmsScratch sc, int B, int A, int nc, float iou_thr, int max_det, DetOut out,
                       cudaStream_t s) {
  size_t smem = (size_t)kMaxCand * 24 + (kMaxCand / 32) * 4 + (size_t)kMaxCand * 2;
  cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  nms_kernel<<<B, NMS_T, smem, s>>>(sc, A, nc, iou_thr, max_det, out);
  return cudaGetLastError();
}

}  // namespace irmv

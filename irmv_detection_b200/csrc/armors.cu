// Light bars -> armors on the GPU: the B200 form of IrmDetector::extract_armors
// (reference src/irm_detector.cpp:292-355) and the Light / Armor constructors
// (reference include/irmv_detection/armor.hpp:11-77).  Oracle: oracle/armor_ref.py (pinned against
// cv2.cvtColor / findContours / minAreaRect).
//
// The reference walks the detections on the CPU: ROI of the rotated frame -> cvtColor(BGR2GRAY) ->
// threshold -> findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) -> per contour with >= 5 points
// minAreaRect -> Light -> is_light -> the first two lights make the armor.  Here one CTA owns one
// detection and never materialises the rotated frame, the gray ROI or a contour:
//   1. the ROI is thresholded straight from the camera frame (rot180 and, for Bayer sources, the
//      demosaic folded into the fetch) into a 1-bit-per-pixel bitmap in shared memory, one
//      __ballot per 32 pixels;
//   2. the exterior background (4-connected to the ROI frame) is flood-filled on the bitmap with
//      carry-propagation word fills: a component is top level (RETR_EXTERNAL) iff the background
//      left of its first pixel is exterior;
//   3. every start candidate (foreground pixel, exterior background to the W, background to the NW,
//      N, NE) is followed with the Suzuki-Abe border walk by one lane; the walk is dropped as soon
//      as it meets a pixel that precedes the candidate in raster order (then it is not the
//      component's first pixel, or it is a hole border), and it counts the CHAIN_APPROX_SIMPLE
//      vertices (direction changes) instead of storing them;
//   4. for a kept border the per-row extreme columns give the convex hull (monotone chain), the
//      hull edges are scanned in parallel for the minimum-area rectangle (same rectangle as
//      rotating calipers), and lane 0 applies the Light constructor and the light filter;
//   5. OpenCV returns contours in reverse raster order of their first pixel, so "the first two
//      lights" are the two valid lights with the largest start index.
// Bitmaps of ROIs larger than the shared-memory budget live in a per-CTA global scratch slot.
#include <math_constants.h>

#include "common.cuh"

namespace irmv {
namespace {

constexpr int kThreads = 128, kWarps = kThreads / 32;
constexpr int kSmemWords = 4096;       // per bitmap: ROIs up to ~128 K padded pixels stay in shared memory
constexpr int kMaxRows = 1088;         // rows of the per-warp row-extreme arrays (ROI height limit)
constexpr int kHullCap = 768;          // a convex lattice polygon in a 1280 x 1088 box has < 400 vertices
constexpr unsigned kFull = 0xffffffffu;

struct LightRec {
  float cx, cy, tx, ty, bx, by;        // centre, top, bottom (source pixels, already offset by the ROI origin)
  double length;
  unsigned key;                        // raster index of the contour's first pixel inside the padded ROI
};

struct WarpBuf {
  short rmin[kMaxRows], rmax[kMaxRows];
  int hull[kHullCap];                  // x | y << 16, ROI coordinates
};

struct Shared {
  uint32_t fg[kSmemWords];
  uint32_t ext[kSmemWords];
  WarpBuf wb[kWarps];
  LightRec best[2];
  int nbest;
  int lock;
};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// Gray value of pixel (X, Y) of the rotated frame as cv::cvtColor(COLOR_BGR2GRAY) computes it on the
// buffer the reference holds (memory channel 0 takes the blue weight, whatever the channel really
// is: src/irm_detector.cpp:310; a Bayer source is demosaiced to RGB like the vendor ISP does).
__device__ __forceinline__ int gray_at(const uint8_t *__restrict__ frame, int W, int H, int chan, int rot, int X, int Y) {
  const int sx = rot ? W - 1 - X : X, sy = rot ? H - 1 - Y : Y;
  int c0, c1, c2;
  if (chan >= 2) {
    int red_y = 0, red_x = 0;
    if (chan == 3) { red_y = 1; red_x = 1; }
    else if (chan == 4) { red_y = 0; red_x = 1; }
    else if (chan == 5) { red_y = 1; red_x = 0; }
    const int ym = reflect101(sy - 1, H), yp = reflect101(sy + 1, H);
    const int xm = reflect101(sx - 1, W), xp = reflect101(sx + 1, W);
    const uint8_t *r0 = frame + (size_t)ym * W, *r1 = frame + (size_t)sy * W, *r2 = frame + (size_t)yp * W;
    const int c = r1[sx];
    const bool red_row = ((sy & 1) == red_y), red_col = ((sx & 1) == red_x);
    if (red_row == red_col) {            // red or blue site
      const int cross = (r0[sx] + r2[sx] + r1[xm] + r1[xp] + 2) >> 2;
      const int diag = (r0[xm] + r0[xp] + r2[xm] + r2[xp] + 2) >> 2;
      c1 = cross;
      c0 = red_row ? c : diag;
      c2 = red_row ? diag : c;
    } else {                             // green site
      const int horiz = (r1[xm] + r1[xp] + 1) >> 1, vert = (r0[sx] + r2[sx] + 1) >> 1;
      c1 = c;
      c0 = red_row ? horiz : vert;
      c2 = red_row ? vert : horiz;
    }
  } else {
    const uint8_t *s = frame + ((size_t)sy * W + sx) * 3;
    c0 = s[0]; c1 = s[1]; c2 = s[2];
  }
  return (c0 * 3735 + c1 * 19235 + c2 * 9798 + (1 << 14)) >> 15;     // OpenCV BY15, GY15, RY15
}

// Seeds `s` (subset of mask `m`) spread along the runs of ones of `m` inside one word.
__device__ __forceinline__ uint32_t fill_runs(uint32_t s, uint32_t m) {
  const uint32_t up = (m & ~(m + s)) | s;
  const uint32_t mr = __brev(m), sr = __brev(s);
  const uint32_t dn = __brev((mr & ~(mr + sr)) | sr);
  return up | dn;
}

// chain code of OpenCV: 0 E, 1 NE, 2 N, 3 NW, 4 W, 5 SW, 6 S, 7 SE (y down)
__device__ __forceinline__ int dir_dx(int s) { return (int)((0x21000122u >> (4 * s)) & 3u) - 1; }
__device__ __forceinline__ int dir_dy(int s) { return (int)((0x22210001u >> (4 * s)) & 3u) - 1; }

// Suzuki-Abe outer-border walk from (x0, y0) (padded ROI coordinates).  Returns false when the
// border holds a pixel that precedes (x0, y0) in raster order.  npts = CHAIN_APPROX_SIMPLE vertex
// count, ymax = last row of the border.  RECORD: per-row extreme columns into rmin / rmax (rows
// relative to y0, which is the component's first row when the walk is the owner's).
template <bool RECORD>
__device__ bool walk_border(const uint32_t *fg, int wpr, int x0, int y0, long long step_cap, int &npts, int &ymax,
                            short *rmin, short *rmax) {
  auto bit = [&](int x, int y) -> bool { return (fg[y * wpr + (x >> 5)] >> (x & 31)) & 1u; };
  int s = 4;
  do { s = (s - 1) & 7; } while (s != 4 && !bit(x0 + dir_dx(s), y0 + dir_dy(s)));
  ymax = y0;
  if (s == 4) {                          // isolated pixel
    npts = 1;
    if (RECORD) { rmin[0] = (short)x0; rmax[0] = (short)x0; }
    return true;
  }
  const int x1 = x0 + dir_dx(s), y1 = y0 + dir_dy(s);
  int x3 = x0, y3 = y0, prev_s = s ^ 4, n = 0;
  for (long long it = 0; it < step_cap; ++it) {
    int x4, y4;
    do { s = (s + 1) & 7; x4 = x3 + dir_dx(s); y4 = y3 + dir_dy(s); } while (!bit(x4, y4));
    if (y4 < y0 || (y4 == y0 && x4 < x0)) return false;
    if (s != prev_s) ++n;
    if (RECORD) {
      const int r = y3 - y0;
      if (x3 < rmin[r]) rmin[r] = (short)x3;
      if (x3 > rmax[r]) rmax[r] = (short)x3;
    }
    ymax = max(ymax, y3);
    prev_s = s;
    if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) { npts = n; return true; }
    x3 = x4; y3 = y4; s = (s + 4) & 7;
  }
  return false;                          // safety net: never reached on a consistent bitmap
}

__device__ __forceinline__ int cross3(int a, int b, int cx, int cy) {
  const int ax = a & 0xffff, ay = a >> 16, bx = b & 0xffff, by = b >> 16;
  return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}

// Kept border -> Light (armor.hpp:15-29) -> filter (armor.hpp:31-38) -> best-two list.
// Warp-cooperative; (x0, y0) padded ROI coordinates of the first pixel, ymax its last row.
__device__ void process_border(Shared &sh, WarpBuf &wb, const uint32_t *fg, int wpr, int x0, int y0, int ymax,
                               long long step_cap, float min_x, float min_y, const ArmorParams &p, int lane) {
  const int rows = ymax - y0 + 1;
  for (int r = lane; r < rows; r += 32) { wb.rmin[r] = 32767; wb.rmax[r] = -1; }
  __syncwarp();
  int nh = 0;
  if (lane == 0) {
    int npts, ym;
    walk_border<true>(fg, wpr, x0, y0, step_cap, npts, ym, wb.rmin, wb.rmax);
    // convex hull from the row extremes: right side downwards, then left side upwards; both chains
    // turn the same way, end points are hull vertices (extreme rows)
    int n = 0;
    bool overflow = false;
    for (int r = 0; r < rows && !overflow; ++r) {
      const int px = wb.rmax[r] - 1, py = y0 + r - 1;
      while (n >= 2 && cross3(wb.hull[n - 2], wb.hull[n - 1], px, py) <= 0) --n;
      if (n >= kHullCap) { overflow = true; break; }
      wb.hull[n++] = px | (py << 16);
    }
    const int base = n;
    for (int r = rows - 1; r >= 0 && !overflow; --r) {
      const int px = wb.rmin[r] - 1, py = y0 + r - 1;
      while (n - base >= 2 && cross3(wb.hull[n - 2], wb.hull[n - 1], px, py) <= 0) --n;
      if (n >= kHullCap) { overflow = true; break; }
      wb.hull[n++] = px | (py << 16);
    }
    if (!overflow) {
      // junction duplicates: bottom (single-pixel last row) and top (single-pixel first row)
      if (n > base && base > 0 && wb.hull[base] == wb.hull[base - 1]) {
        for (int i = base; i + 1 < n; ++i) wb.hull[i] = wb.hull[i + 1];
        --n;
      }
      if (n > 1 && wb.hull[n - 1] == wb.hull[0]) --n;
      nh = n;
    }
  }
  nh = __shfl_sync(kFull, nh, 0);
  __syncwarp();
  if (nh < 2) return;
  // minimum-area enclosing rectangle: one hull edge per lane; projections relative to hull[0]
  // (small integers: the FP32 products are exact or nearly so whatever the ROI size)
  const int org_x = wb.hull[0] & 0xffff, org_y = wb.hull[0] >> 16;
  float best_area = CUDART_INF_F;
  int best_i = 0x7fffffff;
  for (int i = lane; i < nh; i += 32) {
    const int a = wb.hull[i], b = wb.hull[i + 1 == nh ? 0 : i + 1];
    const int ex = (b & 0xffff) - (a & 0xffff), ey = (b >> 16) - (a >> 16);
    const float ln = __fsqrt_rn((float)(ex * ex + ey * ey));
    const float ux = __fdiv_rn((float)ex, ln), uy = __fdiv_rn((float)ey, ln), vx = -uy, vy = ux;
    float u0 = CUDART_INF_F, u1 = -CUDART_INF_F, v0 = CUDART_INF_F, v1 = -CUDART_INF_F;
    for (int j = 0; j < nh; ++j) {
      const int h = wb.hull[j];
      const float hx = (float)((h & 0xffff) - org_x), hy = (float)((h >> 16) - org_y);
      const float pu = __fadd_rn(__fmul_rn(hx, ux), __fmul_rn(hy, uy));
      const float pv = __fadd_rn(__fmul_rn(hx, vx), __fmul_rn(hy, vy));
      u0 = fminf(u0, pu); u1 = fmaxf(u1, pu); v0 = fminf(v0, pv); v1 = fmaxf(v1, pv);
    }
    const float area = __fmul_rn(__fsub_rn(u1, u0), __fsub_rn(v1, v0));
    if (area < best_area) { best_area = area; best_i = i; }
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    const float oa = __shfl_xor_sync(kFull, best_area, off);
    const int oi = __shfl_xor_sync(kFull, best_i, off);
    if (oa < best_area || (oa == best_area && oi < best_i)) { best_area = oa; best_i = oi; }
  }
  if (lane != 0 || best_i == 0x7fffffff) return;
  // corners of the chosen rectangle
  float cxs[4], cys[4];
  {
    const int i = best_i;
    const int a = wb.hull[i], b = wb.hull[i + 1 == nh ? 0 : i + 1];
    const int ex = (b & 0xffff) - (a & 0xffff), ey = (b >> 16) - (a >> 16);
    const float ln = __fsqrt_rn((float)(ex * ex + ey * ey));
    const float ux = __fdiv_rn((float)ex, ln), uy = __fdiv_rn((float)ey, ln), vx = -uy, vy = ux;
    float u0 = CUDART_INF_F, u1 = -CUDART_INF_F, v0 = CUDART_INF_F, v1 = -CUDART_INF_F;
    for (int j = 0; j < nh; ++j) {
      const int h = wb.hull[j];
      const float hx = (float)((h & 0xffff) - org_x), hy = (float)((h >> 16) - org_y);
      const float pu = __fadd_rn(__fmul_rn(hx, ux), __fmul_rn(hy, uy));
      const float pv = __fadd_rn(__fmul_rn(hx, vx), __fmul_rn(hy, vy));
      u0 = fminf(u0, pu); u1 = fmaxf(u1, pu); v0 = fminf(v0, pv); v1 = fmaxf(v1, pv);
    }
    const float us[4] = {u0, u1, u1, u0}, vs[4] = {v0, v0, v1, v1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cxs[k] = __fadd_rn(__fadd_rn(__fmul_rn(ux, us[k]), __fmul_rn(vx, vs[k])), (float)org_x);
      cys[k] = __fadd_rn(__fadd_rn(__fmul_rn(uy, us[k]), __fmul_rn(vy, vs[k])), (float)org_y);
    }
  }
  const float cen_x = __fmul_rn(__fadd_rn(cxs[0], cxs[2]), 0.5f), cen_y = __fmul_rn(__fadd_rn(cys[0], cys[2]), 0.5f);
  // Light::Light: corners by ascending y (stable), top / bottom = midpoints of the upper / lower pair
  int o0 = 0, o1 = 1, o2 = 2, o3 = 3;
  auto cswap = [&](int &a, int &b) { if (cys[b] < cys[a]) { const int t = a; a = b; b = t; } };
  cswap(o0, o1); cswap(o2, o3); cswap(o0, o2); cswap(o1, o3); cswap(o1, o2);
  // (the 5-comparator network is not stable by itself; equal y only happens for axis-parallel
  // rectangles, where either order gives the same midpoints and the same width)
  const float tx = __fdiv_rn(__fadd_rn(cxs[o0], cxs[o1]), 2.f), ty = __fdiv_rn(__fadd_rn(cys[o0], cys[o1]), 2.f);
  const float bx = __fdiv_rn(__fadd_rn(cxs[o2], cxs[o3]), 2.f), by = __fdiv_rn(__fadd_rn(cys[o2], cys[o3]), 2.f);
  const float dx = __fsub_rn(tx, bx), dy = __fsub_rn(ty, by);
  const double length = sqrt((double)dx * dx + (double)dy * dy);
  const float wx = __fsub_rn(cxs[o0], cxs[o1]), wy = __fsub_rn(cys[o0], cys[o1]);
  const double width = sqrt((double)wx * wx + (double)wy * wy);
  const double tilt = (double)atan2f(fabsf(dx), fabsf(dy)) / 3.1415926535897932384626433832795 * 180.0;
  const double ratio = width / length;                 // 0/0 = NaN compares false like the reference
  if (!((double)p.min_ratio < ratio && ratio < (double)p.max_ratio && tilt < (double)p.max_angle)) return;
  LightRec rec;
  rec.cx = __fadd_rn(cen_x, min_x); rec.cy = __fadd_rn(cen_y, min_y);     // Light::offset_bbox
  rec.tx = __fadd_rn(tx, min_x); rec.ty = __fadd_rn(ty, min_y);
  rec.bx = __fadd_rn(bx, min_x); rec.by = __fadd_rn(by, min_y);
  rec.length = length;
  rec.key = (unsigned)(y0 * wpr * 32 + x0);
  while (atomicCAS(&sh.lock, 0, 1) != 0) {}
  __threadfence_block();
  {
    volatile int *nb = &sh.nbest;
    const int n = *nb;
    LightRec b0 = sh.best[0], b1 = sh.best[1];
    if (n == 0 || rec.key > b0.key) { sh.best[1] = b0; sh.best[0] = rec; }
    else if (n == 1 || rec.key > b1.key) sh.best[1] = rec;
    *nb = n < 2 ? n + 1 : 2;
  }
  __threadfence_block();
  atomicExch(&sh.lock, 0);
}

__global__ void __launch_bounds__(kThreads) extract_armors_kernel(ArmorParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Shared &sh = *reinterpret_cast<Shared *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = p.src_w, H = p.src_h;
  const bool bayer = p.chan_order >= 2;
  const size_t frame_bytes = (size_t)W * H * (bayer ? 1 : 3);
  const uint8_t *base = p.src_indirect ? *p.src_indirect : p.src;
  const int total = p.n * p.max_det;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int f = item / p.max_det, k = item - f * p.max_det;
    ArmorOut *out = p.out + item;
    if (k >= p.num[f]) {
      if (tid == 0) out->valid = 0;
      continue;
    }
    // ROI (src/irm_detector.cpp:299-304): float clamp, cv::Rect truncation
    const float4 b = reinterpret_cast<const float4 *>(p.boxes)[item];
    const float x1 = __fmul_rn(__fsub_rn(b.x, p.box_px), p.box_sx), y1 = __fmul_rn(__fsub_rn(b.y, p.box_py), p.box_sy);
    const float x2 = __fmul_rn(__fsub_rn(b.z, p.box_px), p.box_sx), y2 = __fmul_rn(__fsub_rn(b.w, p.box_py), p.box_sy);
    const float min_x = x1 > 0.f ? x1 : 0.f, min_y = y1 > 0.f ? y1 : 0.f;
    const float max_x = x2 < (float)W ? x2 : (float)W, max_y = y2 < (float)H ? y2 : (float)H;
    const int rx = (int)min_x, ry = (int)min_y;
    const int rw = (int)__fsub_rn(max_x, min_x), rh = (int)__fsub_rn(max_y, min_y);
    if (!(min_x < max_x && min_y < max_y) || rw <= 0 || rh <= 0 || rh > kMaxRows - 2) {
      if (tid == 0) out->valid = 0;
      continue;
    }
    const int PW = rw + 2, PH = rh + 2, wpr = (PW + 31) >> 5, words = PH * wpr;
    uint32_t *fg = sh.fg, *ext = sh.ext;
    if (words > kSmemWords) {
      fg = p.scratch + (size_t)blockIdx.x * p.scratch_words_per_cta;
      ext = fg + p.scratch_words_per_cta / 2;
    }
    if (tid == 0) { sh.nbest = 0; sh.lock = 0; }
    const uint8_t *frame = base + (size_t)f * frame_bytes;
    // 1. threshold bitmap, padded by one background pixel all round: a warp builds one word per ballot
    for (int w = warp; w < words; w += kWarps) {
      const int row = w / wpr, wi = w - row * wpr;
      const int x = wi * 32 + lane - 1, y = row - 1;
      bool on = false;
      if (x >= 0 && x < rw && y >= 0 && y < rh) on = gray_at(frame, W, H, p.chan_order, p.rotate180, rx + x, ry + y) > p.binary_threshold;
      const uint32_t m = __ballot_sync(kFull, on);
      if (lane == 0) fg[w] = m;
    }
    __syncthreads();
    // 2. exterior background: seeds on the padding, spread along runs inside each word ...
    for (int w = tid; w < words; w += kThreads) {
      const int row = w / wpr, wi = w - row * wpr;
      const uint32_t m = ~fg[w];
      uint32_t s = (row == 0 || row == PH - 1) ? kFull : 0u;
      if (wi == 0) s |= 1u;
      if (wi == wpr - 1) s |= kFull << ((PW - 1) & 31);
      ext[w] = fill_runs(s & m, m);
    }
    __syncthreads();
    // ... then row sweeps (down, up) by one warp until nothing changes
    if (warp == 0) {
      for (;;) {
        bool changed = false;
        for (int pass = 0; pass < 2; ++pass) {
          for (int i = 1; i <= PH - 2; ++i) {
            const int r = pass == 0 ? i : PH - 1 - i;
            uint32_t *er = ext + r * wpr;
            const uint32_t *fr = fg + r * wpr;
            bool ch = false;
            for (int wi = lane; wi < wpr; wi += 32) {
              const uint32_t m = ~fr[wi], e = er[wi];
              const uint32_t s = e | (m & (er[wi - wpr] | er[wi + wpr]));
              if (s != e) { er[wi] = fill_runs(s, m); ch = true; }
            }
            ch = __any_sync(kFull, ch);
            if (wpr > 1) {
              for (;;) {                          // runs that cross a word boundary
                __syncwarp();
                bool c2 = false;
                for (int wi = lane; wi < wpr; wi += 32) {
                  const uint32_t m = ~fr[wi], e = er[wi];
                  const uint32_t lft = wi > 0 ? er[wi - 1] >> 31 : 0u, rgt = wi + 1 < wpr ? er[wi + 1] & 1u : 0u;
                  const uint32_t s = e | (m & (lft | (rgt << 31)));
                  if (s != e) { er[wi] = fill_runs(s, m); c2 = true; }
                }
                c2 = __any_sync(kFull, c2);
                if (!c2) break;
                ch = true;
              }
            }
            __syncwarp();
            changed |= ch;
          }
        }
        if (!changed) break;
      }
    }
    __syncthreads();
    // 3. border walks from the start candidates, 4. lights
    const long long step_cap = 8ll * PW * PH;
    const int nchunks = (words + 31) >> 5;
    for (int chunk = warp; chunk < nchunks; chunk += kWarps) {
      const int w = chunk * 32 + lane;
      uint32_t cand = 0;
      int row = 0, wi = 0;
      if (w < words) {
        row = w / wpr; wi = w - row * wpr;
        if (row >= 1 && row <= PH - 2) {
          const uint32_t f0 = fg[w], e0 = ext[w];
          const uint32_t ep = wi > 0 ? ext[w - 1] >> 31 : 0u;
          const uint32_t un = fg[w - wpr];
          const uint32_t up = wi > 0 ? fg[w - wpr - 1] >> 31 : 0u, ux = wi + 1 < wpr ? fg[w - wpr + 1] << 31 : 0u;
          const uint32_t ext_w = (e0 << 1) | ep;
          const uint32_t above = un | (un << 1) | up | (un >> 1) | ux;
          cand = f0 & ext_w & ~above;
        }
      }
      while (__any_sync(kFull, cand != 0)) {
        bool acc = false;
        int x0 = 0, y0 = 0, ymax = 0;
        if (cand) {
          const int bpos = __ffs(cand) - 1;
          cand &= cand - 1;
          x0 = wi * 32 + bpos; y0 = row;
          int npts = 0;
          acc = walk_border<false>(fg, wpr, x0, y0, step_cap, npts, ymax, nullptr, nullptr) && npts >= 5;
        }
        uint32_t m = __ballot_sync(kFull, acc);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const int X0 = __shfl_sync(kFull, x0, src), Y0 = __shfl_sync(kFull, y0, src), YM = __shfl_sync(kFull, ymax, src);
          process_border(sh, sh.wb[warp], fg, wpr, X0, Y0, YM, step_cap, min_x, min_y, p, lane);
          __syncwarp();
        }
      }
    }
    __syncthreads();
    // 5. Armor::Armor (armor.hpp:58-68) and the centre-distance filter (src/irm_detector.cpp:333-350)
    if (tid == 0) {
      int valid = 0;
      if (sh.nbest >= 2) {
        const LightRec l0 = sh.best[0], l1 = sh.best[1];
        const bool first_left = l0.cx < l1.cx;
        const LightRec L = first_left ? l0 : l1, R = first_left ? l1 : l0;
        const double avg_len = (l0.length + l1.length) / 2;
        const float ddx = __fsub_rn(L.cx, R.cx), ddy = __fsub_rn(L.cy, R.cy);
        const double cd = sqrt((double)ddx * ddx + (double)ddy * ddy) / avg_len;
        const int size = cd > p.min_large ? 1 : 0;
        bool ok = true;
        if (size == 0 && (p.min_small > cd || p.max_small < cd)) ok = false;
        if (size == 1 && (p.min_large > cd || p.max_large < cd)) ok = false;
        if (ok) {
          out->pts[0] = L.bx; out->pts[1] = L.by; out->pts[2] = L.tx; out->pts[3] = L.ty;
          out->pts[4] = R.tx; out->pts[5] = R.ty; out->pts[6] = R.bx; out->pts[7] = R.by;
          out->center[0] = __fdiv_rn(__fadd_rn(L.cx, R.cx), 2.f);
          out->center[1] = __fdiv_rn(__fadd_rn(L.cy, R.cy), 2.f);
          out->score = p.scores[item];
          const int c = p.classes[item];
          out->class_id = (c >= 0 && c < 14) ? c : 14;
          out->size = size;
          valid = 1;
        }
      }
      out->valid = valid;
    }
    __syncthreads();
  }
}

// Armor corners {left.bottom, left.top, right.top, right.bottom} -> PnP quads in the calibration
// frame (src/pnp_solver.cpp:41-44); slots without an armor get a fixed valid quad and are masked after.
__global__ void quads_from_armors_kernel(const ArmorOut *armors, int total, float sx, float sy, float *pts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const ArmorOut &a = armors[i];
  float q[8] = {150.f, 160.f, 150.f, 140.f, 200.f, 140.f, 200.f, 160.f};
  if (a.valid) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { q[2 * k] = a.pts[2 * k] * sx; q[2 * k + 1] = a.pts[2 * k + 1] * sy; }
  }
  float4 *o = reinterpret_cast<float4 *>(pts + (size_t)i * 8);
  o[0] = make_float4(q[0], q[1], q[2], q[3]);
  o[1] = make_float4(q[4], q[5], q[6], q[7]);
}

__global__ void mask_pose_ok_kernel(const ArmorOut *armors, int total, uint8_t *ok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total && !armors[i].valid) ok[i] = 0;
}

}  // namespace

int armors_grid(int num_sms) { return num_sms * 3; }

size_t armors_scratch_words_per_cta(int src_w, int src_h) {
  const size_t wpr = ((size_t)src_w + 2 + 31) / 32;
  return 2 * wpr * ((size_t)src_h + 2);
}

cudaError_t launch_extract_armors(const ArmorParams &p, cudaStream_t s) {
  if (p.src_h > kMaxRows - 2 || p.src_w > 32766) return cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(extract_armors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Shared));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int total = p.n * p.max_det;
  if (total <= 0) return cudaSuccess;
  const int grid = total < p.grid ? total : p.grid;
  extract_armors_kernel<<<grid, kThreads, sizeof(Shared), s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_quads_from_armors(const ArmorOut *armors, int total, float sx, float sy, float *pts, cudaStream_t s) {
  if (total <= 0) return cudaSuccess;
  quads_from_armors_kernel<<<(total + 127) / 128, 128, 0, s>>>(armors, total, sx, sy, pts);
  return cudaGetLastError();
}

cudaError_t launch_mask_pose_ok(const ArmorOut *armors, int total, uint8_t *ok, cudaStream_t s) {
  if (total <= 0) return cudaSuccess;
  mask_pose_ok_kernel<<<(total + 127) / 128, 128, 0, s>>>(armors, total, ok);
  return cudaGetLastError();
}

}  // namespace irmv

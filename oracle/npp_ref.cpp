// TEST INFRASTRUCTURE ONLY -- the reference's literal preprocessing chain, for pinning the oracle.
//
// Restates the four NPP calls of YoloEngine::preprocess (reference src/yolo_engine.cpp:179-200)
// with the same arguments, so that NPP's undocumented bilinear convention can be measured on a
// GPU box.  This is our own harness around the NPP library the
// reference links (CMakeLists.txt:62-67); no reference source is copied.  Built by
// oracle/Makefile into oracle/_ref/npp_ref; needs a GPU to run.
//
// usage: npp_ref <in.raw u8 HxWx3> <W> <H> <out.f32 3x640x640> [<out_rotated.raw>]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>
#include <npp.h>

#define CK(x) do { auto e_ = (x); if (e_ != 0) { fprintf(stderr, "%s failed: %d\n", #x, (int)e_); return 2; } } while (0)

int main(int argc, char **argv) {
  if (argc < 5) { fprintf(stderr, "usage: npp_ref in.raw W H out.f32 [rot.raw]\n"); return 1; }
  const int W = atoi(argv[2]), H = atoi(argv[3]);
  std::vector<unsigned char> img((size_t)W * H * 3);
  FILE *f = fopen(argv[1], "rb");
  if (!f || fread(img.data(), 1, img.size(), f) != img.size()) { fprintf(stderr, "bad input\n"); return 1; }
  fclose(f);
  unsigned char *src, *resized;
  float *hwc, *chw;
  CK(cudaMalloc(&src, img.size()));
  CK(cudaMalloc(&resized, 640 * 640 * 3));
  CK(cudaMalloc(&hwc, 640 * 640 * 3 * sizeof(float)));
  CK(cudaMalloc(&chw, 640 * 640 * 3 * sizeof(float)));
  CK(cudaMemcpy(src, img.data(), img.size(), cudaMemcpyHostToDevice));
  cudaStream_t stream;
  CK(cudaStreamCreate(&stream));
  NppStreamContext ctx;
  CK(nppGetStreamContext(&ctx));
  ctx.hStream = stream;
  Npp32f *planes[3] = {chw, chw + 640 * 640, chw + 640 * 640 * 2};
  // K1 rotate 180 degrees in place
  CK(nppiMirror_8u_C3IR_Ctx(src, W * 3, NppiSize{W, H}, NPP_BOTH_AXIS, ctx));
  // K2 stretch to 640x640, bilinear
  CK(nppiResize_8u_C3R_Ctx(src, W * 3, NppiSize{W, H}, NppiRect{0, 0, W, H}, resized, 640 * 3, NppiSize{640, 640},
                           NppiRect{0, 0, 640, 640}, NPPI_INTER_LINEAR, ctx));
  // K3 u8 -> f32 in [0,1]
  CK(nppiScale_8u32f_C3R_Ctx(resized, 640 * 3, hwc, 640 * 3 * sizeof(float), NppiSize{640, 640}, 0.0, 1.0, ctx));
  // K4 packed -> planar
  CK(nppiCopy_32f_C3P3R_Ctx(hwc, 640 * 3 * sizeof(float), planes, 640 * sizeof(float), NppiSize{640, 640}, ctx));
  CK(cudaStreamSynchronize(stream));
  std::vector<float> out(640 * 640 * 3);
  CK(cudaMemcpy(out.data(), chw, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
  f = fopen(argv[4], "wb");
  fwrite(out.data(), sizeof(float), out.size(), f);
  fclose(f);
  if (argc > 5) {
    CK(cudaMemcpy(img.data(), src, img.size(), cudaMemcpyDeviceToHost));
    f = fopen(argv[5], "wb");
    fwrite(img.data(), 1, img.size(), f);
    fclose(f);
  }
  int v = 0;
  const NppLibraryVersion *lv = nppGetLibVersion();
  if (lv) v = lv->major * 1000 + lv->minor * 10 + lv->build;
  printf("npp_ref ok, NPP %d\n", v);
  return 0;
}

// tcgen05.mma micro-probe (sm_100a): what does one CTA sustain for M128 x N x K16 f16 MMAs as a
// function of operand layout (no swizzle / 128B swizzle), start-address alignment, N, and the number
// of independent accumulator chains?  Also checks that a ROW-SHIFTED start address works with the
// 128B swizzle (the raster convolution reads its 9 taps as shifted views of one halo tile).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -cudart shared -o scripts/_bin/mma_probe scripts/mma_probe.cu
// (-cudart shared: a statically linked runtime would carry the names of API calls this repo must not ship)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %1;\n@%%px mov.s32 %0, 1;\n}" : "+r"(pred) : "r"(0xffffffffu));
  return pred;
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// layout: 0 none, 2 SW128, 4 SW64, 6 SW32
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)layout << 61;
  return d;
}

struct Cfg {
  int mode;      // 0 no-swizzle aligned, 1 no-swizzle start+16B & LBO = 16 (mod 128), 2 SW128, 3 SW128 row-shifted (base_off set),
                 // 4 SW128 row-shifted (base_off 0), 5 SW64, 6 SW32
  int N, chains, iters, shift;
};

__global__ void __launch_bounds__(128) probe(Cfg c, long long *cycles, float *dout) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t *sA = smem;                    // 96 KB
  uint8_t *sB = smem + 96 * 1024;        // 64 KB
  const int tid = threadIdx.x, warp = tid >> 5;
  // ---- fill operands.  Logical A[r][k] (r < 512, k < 64) = ((r * 7 + k * 3) % 17 - 8) / 8 ; B[n][k] = (n == k)
  const bool sw = c.mode >= 2;
  const int rowbytes = c.mode == 5 ? 64 : c.mode == 6 ? 32 : 128;     // swizzle span
  for (int i = tid; i < 96 * 1024 / 2; i += 128) reinterpret_cast<__half *>(sA)[i] = __float2half(0.f);
  for (int i = tid; i < 64 * 1024 / 2; i += 128) reinterpret_cast<__half *>(sB)[i] = __float2half(0.f);
  __syncthreads();
  const int ROWS = 512;
  const int lboA_rows = c.mode == 1 ? ROWS + 1 : ROWS;     // no-swizzle plane pitch in rows
  for (int i = tid; i < ROWS * 64; i += 128) {
    const int r = i >> 6, k = i & 63;
    const float v = (float)(((r * 7 + k * 3) % 17) - 8) * 0.125f;
    size_t off;
    if (!sw) off = ((size_t)(k >> 3) * lboA_rows + r) * 16 + (k & 7) * 2 + (c.mode == 1 ? 16 : 0);
    else if (c.mode <= 4) off = (size_t)r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1));
    else if (c.mode == 5) {   // SW64: [k/32][r][32 halfs], 16B chunk ^= (r>>1)&3
      off = (size_t)(k >> 5) * ROWS * 64 + (size_t)r * 64 + (((((k & 31) >> 3) ^ ((r >> 1) & 3)) << 4) | ((k & 7) << 1));
    } else {                  // SW32: [k/16][r][16 halfs], 16B chunk ^= (r>>2)&1
      off = (size_t)(k >> 4) * ROWS * 32 + (size_t)r * 32 + (((((k & 15) >> 3) ^ ((r >> 2) & 1)) << 4) | ((k & 7) << 1));
    }
    if (off + 2 <= 96 * 1024) *reinterpret_cast<__half *>(sA + off) = __float2half(v);
  }
  for (int i = tid; i < 256 * 64; i += 128) {
    const int n = i >> 6, k = i & 63;
    const float v = (n == k) ? 1.f : 0.f;
    size_t off;
    if (!sw) off = ((size_t)(k >> 3) * 256 + n) * 16 + (k & 7) * 2;
    else if (c.mode <= 4) off = (size_t)n * 128 + ((((k >> 3) ^ (n & 7)) << 4) | ((k & 7) << 1));
    else if (c.mode == 5) off = (size_t)(k >> 5) * 256 * 64 + (size_t)n * 64 + (((((k & 31) >> 3) ^ ((n >> 1) & 3)) << 4) | ((k & 7) << 1));
    else off = (size_t)(k >> 4) * 256 * 32 + (size_t)n * 32 + (((((k & 15) >> 3) ^ ((n >> 2) & 1)) << 4) | ((k & 7) << 1));
    *reinterpret_cast<__half *>(sB + off) = __float2half(v);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  auto descA = [&](int row0, int ks) -> uint64_t {   // ks = K16 step 0..3
    if (c.mode == 0) return make_desc(smem_u32(sA) + (uint32_t)(2 * ks) * ROWS * 16 + row0 * 16, ROWS * 16, 128, 0, 0);
    if (c.mode == 1) return make_desc(smem_u32(sA) + 16 + (uint32_t)(2 * ks) * lboA_rows * 16 + row0 * 16, lboA_rows * 16, 128, 0, 0);
    if (c.mode <= 4) {
      const uint32_t a = smem_u32(sA) + row0 * 128 + ks * 32;
      return make_desc(a, 16, 1024, 2, c.mode == 3 ? (a >> 7) & 7 : 0);
    }
    if (c.mode == 5) {
      const uint32_t a = smem_u32(sA) + (uint32_t)(ks >> 1) * ROWS * 64 + row0 * 64 + (ks & 1) * 32;
      return make_desc(a, 16, 512, 4, 0);
    }
    const uint32_t a = smem_u32(sA) + (uint32_t)ks * ROWS * 32 + row0 * 32;
    return make_desc(a, 16, 256, 6, 0);
  };
  auto descB = [&](int ks) -> uint64_t {
    if (!sw) return make_desc(smem_u32(sB) + (uint32_t)(2 * ks) * 256 * 16, 256 * 16, 128, 0, 0);
    if (c.mode <= 4) return make_desc(smem_u32(sB) + ks * 32, 16, 1024, 2, 0);
    if (c.mode == 5) return make_desc(smem_u32(sB) + (uint32_t)(ks >> 1) * 256 * 64 + (ks & 1) * 32, 16, 512, 4, 0);
    return make_desc(smem_u32(sB) + (uint32_t)ks * 256 * 32, 16, 256, 6, 0);
  };

  // ---- correctness: D[m][n] = A[m + shift][n] for n < 64 (B = identity), one chain
  if (tid == 0) {
    for (int ks = 0; ks < 4; ++ks) tc_mma_f16(tmem, descA(c.shift, ks), descB(ks), idesc, ks != 0);
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (blockIdx.x == 0 && dout) {
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t v[16];
      tc_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) dout[(size_t)tid * 64 + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- throughput: 9 taps x 4 K-steps x chains per "tile", descriptors advanced by uniform adds
  if (warp == 0 && elect_one()) {
    const uint64_t a0 = descA(c.shift, 0), b0 = descB(0);
    const uint32_t a_hi = (uint32_t)(a0 >> 32), b_hi = (uint32_t)(b0 >> 32);
    const uint32_t a_lo0 = (uint32_t)a0, b_lo0 = (uint32_t)b0;
    const uint32_t a_k = (uint32_t)descA(c.shift, 1) - a_lo0, b_k = (uint32_t)descB(1) - b_lo0;   // per K16 step
    const uint32_t a_row = (uint32_t)descA(c.shift + 1, 0) - a_lo0;                              // per pixel row
    const uint32_t a_ch = a_row * 128;
    const int tiles = c.iters / (36 * c.chains);
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      uint32_t a_ty = a_lo0;
      for (int ty = 0; ty < 3; ++ty, a_ty += a_row * (c.shift ? 83 : 0)) {
        uint32_t a_tx = a_ty;
#pragma unroll
        for (int tx = 0; tx < 3; ++tx, a_tx += c.shift ? a_row : 0) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t al = a_tx + ks * a_k, bl = b_lo0 + ks * b_k;
            if (c.chains == 1) {
              mma2(tmem, al, a_hi, bl, b_hi, idesc);
            } else if (c.chains == 2) {
              mma2(tmem, al, a_hi, bl, b_hi, idesc);
              mma2(tmem + c.N, al + a_ch, a_hi, bl, b_hi, idesc);
            } else {
              mma2(tmem, al, a_hi, bl, b_hi, idesc);
              mma2(tmem + c.N, al + a_ch, a_hi, bl, b_hi, idesc);
              mma2(tmem + 2 * c.N, al + 2 * a_ch, a_hi, bl, b_hi, idesc);
              mma2(tmem + 3 * c.N, al + 3 * a_ch, a_hi, bl, b_hi, idesc);
            }
          }
        }
      }
    }
    const long long t1 = clock64();
    tc_commit(&bar);
    mbar_wait(&bar, 1);
    const long long t2 = clock64();
    cycles[blockIdx.x * 2 + 0] = (t1 - t0) * c.iters / (tiles * 36 * c.chains);
    cycles[blockIdx.x * 2 + 1] = (t2 - t0) * c.iters / (tiles * 36 * c.chains);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main(int argc, char **argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  long long *dc; float *dd;
  cudaMalloc(&dc, grid * 16); cudaMalloc(&dd, 128 * 64 * 4);
  std::vector<long long> hc(grid * 2); std::vector<float> hd(128 * 64);
  const char *names[] = {"noswz-aligned", "noswz-mis16", "sw128", "sw128-shift-bo", "sw128-shift-bo0", "sw64", "sw32"};
  printf("%-16s %4s %6s %5s | %9s %9s | %s\n", "mode", "N", "chains", "shift", "cyc/mma", "issue/mma", "mismatch(of 8192)");
  for (int mode = 0; mode < 7; ++mode)
    for (int N : {16, 32, 64, 128, 256})
      for (int chains : {1, 2, 4}) {
        if (chains * N > 512) continue;
        for (int shift : {0, 83}) {
          if (mode == 2 && shift) continue;
          if ((mode == 3 || mode == 4) && !shift) continue;
          Cfg c{mode, N, chains, 1152, shift};
          cudaMemset(dd, 0, 128 * 64 * 4);
          probe<<<grid, 128, 160 * 1024>>>(c, dc, dd);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%-16s N=%d failed: %s\n", names[mode], N, cudaGetErrorString(e)); return 1; }
          cudaMemcpy(hc.data(), dc, grid * 16, cudaMemcpyDeviceToHost);
          cudaMemcpy(hd.data(), dd, 128 * 64 * 4, cudaMemcpyDeviceToHost);
          int bad = 0;
          const int ncheck = N < 64 ? N : 64;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < ncheck; ++n) {
              const int r = m + shift;
              const float ref = (float)(((r * 7 + n * 3) % 17) - 8) * 0.125f;
              if (hd[m * 64 + n] != ref) ++bad;
            }
          double tot = 0, iss = 0;
          for (int b = 0; b < grid; ++b) { iss += hc[b * 2]; tot += hc[b * 2 + 1]; }
          printf("%-16s %4d %6d %5d | %9.1f %9.1f | %d\n", names[mode], N, chains, shift, tot / grid / c.iters, iss / grid / c.iters, bad);
        }
      }
  return 0;
}

// Micro-benchmark (B200): cost of the softmax row math on a 128 x 208 fp32 tile held in tensor memory, per variant.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I dinov3-jax_b200/csrc -o tools/microbench/bin/rowmath tools/microbench/rowmath.cu
#include <cstdio>
#include "ptx.cuh"
using namespace d3;

__device__ __forceinline__ float max16(const uint32_t* v) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 8]));
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = fmaxf(a[i], a[i + 4]);
  return fmaxf(fmaxf(a[0], a[2]), fmaxf(a[1], a[3]));
}
// bf16 round-to-nearest-even of two fp32 with integer ops only (no F2FP)
__device__ __forceinline__ uint32_t pack_int(float a, float b) {
  uint32_t x = __float_as_uint(a), y = __float_as_uint(b);
  x += 0x7fffu + ((x >> 16) & 1u);
  y += 0x7fffu + ((y >> 16) & 1u);
  return __byte_perm(x, y, 0x7632);
}
// truncating pack (1 PRMT)
__device__ __forceinline__ uint32_t pack_trunc(float a, float b) { return __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632); }

// mode: 0 max pass only; 1 exp pass F2FP pack; 2 exp pass integer pack; 3 exp pass truncating pack; 4 exp pass, no pack / no store
//       5 exp pass F2FP without tmem store; 6 ld only (sum); 7 max + exp(F2FP) both passes
// SPIN: 0 none; 1: three extra warps whose lane 0 spins in mbarrier try_wait (like the TMA / MMA issuer lanes of the
// attention kernel while they wait for the softmax warps); 2: the same with the other 31 lanes parked at __syncwarp
template <int MODE, int SPIN>
__global__ void bench(long long* out, int tiles, float* sink, const uint8_t* gsrc) {
  __shared__ uint32_t slot;
  __shared__ uint64_t never;
  __shared__ volatile int done;
  const int nwork = blockDim.x / 32 - (SPIN ? 3 : 0);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&never, 1); done = 0; fence_mbar_init(); }
  if (SPIN == 3 && warp >= nwork) {
    extern __shared__ __align__(1024) uint8_t ring[];
    __shared__ uint64_t cbar;
    __syncthreads(); __syncthreads();
    if (warp == nwork && elect_one()) {
      mbar_init(&cbar, 1); fence_mbar_init();
      uint32_t ph = 0;
      int i = 0;
      while (!done) {
        mbar_expect_tx(&cbar, 4 * 16384);
        for (int k = 0; k < 4; ++k)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(ring + k * 16384)), "l"(gsrc + ((size_t)(blockIdx.x * 64 + (i++ & 63)) << 14)), "r"(16384), "r"(smem_u32(&cbar)) : "memory");
        mbar_wait(&cbar, ph); ph ^= 1;
      }
    }
    __syncwarp();
    __syncthreads(); __syncthreads();
    return;
  }
  if (SPIN && warp >= nwork) {
    __syncthreads(); __syncthreads();
    if (SPIN == 1 ? (threadIdx.x & 31) == 0 : elect_one()) {
      while (!done) { if (mbar_try_wait(&never, 0)) break; }
    }
    __syncwarp();
    __syncthreads(); __syncthreads();
    return;
  }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  const int NK = 208;
  const float cs = 0.18f;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    float mx = -3e38f, sum = 0.f;
    uint32_t va[32], vb[32];
    if (MODE == 0 || MODE == 7) {
      tmem_ld32(base, va); tmem_ld_wait();
      for (int c0 = 0; c0 < NK; c0 += 64) {
        const bool hasB = c0 + 32 < NK;
        if (hasB) { if (c0 + 64 <= NK) tmem_ld32(base + c0 + 32, vb); else tmem_ld16(base + c0 + 32, vb); }
        mx = fmaxf(mx, max16(va)); if (c0 + 16 < NK) mx = fmaxf(mx, max16(va + 16));
        tmem_ld_wait();
        if (hasB) {
          if (c0 + 64 < NK) { if (c0 + 96 <= NK) tmem_ld32(base + c0 + 64, va); else tmem_ld16(base + c0 + 64, va); }
          mx = fmaxf(mx, max16(vb)); if (c0 + 48 < NK) mx = fmaxf(mx, max16(vb + 16));
          tmem_ld_wait();
        }
      }
    } else mx = 1.f;
    const float mxs = mx * cs;
    if (MODE != 0) {
      auto emit = [&](const uint32_t* v, int c0) {
        uint32_t pw[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float p0, p1;
          if (MODE == 6) { p0 = __uint_as_float(v[i]); p1 = __uint_as_float(v[i + 1]); }
          else { p0 = ex2_approx(fmaf(__uint_as_float(v[i]), cs, -mxs)); p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), cs, -mxs)); }
          sum += p0 + p1;
          if (MODE == 1 || MODE == 5 || MODE == 7) pw[i >> 1] = pack_bf16(p0, p1);
          else if (MODE == 2) pw[i >> 1] = pack_int(p0, p1);
          else if (MODE == 3) pw[i >> 1] = pack_trunc(p0, p1);
          else pw[i >> 1] = 0;
        }
        if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 7) tmem_st8(base + (c0 >> 1), pw);
        else if (MODE == 5) { acc += __uint_as_float(pw[0] ^ pw[3] ^ pw[5] ^ pw[7] ^ pw[1] ^ pw[2] ^ pw[4] ^ pw[6]); }
      };
      tmem_ld32(base, va); tmem_ld_wait();
      for (int c0 = 0; c0 < NK; c0 += 64) {
        const bool hasB = c0 + 32 < NK;
        if (hasB) { if (c0 + 64 <= NK) tmem_ld32(base + c0 + 32, vb); else tmem_ld16(base + c0 + 32, vb); }
        tmem_ld_wait();
        emit(va, c0); if (c0 + 16 < NK) emit(va + 16, c0 + 16);
        if (hasB) {
          if (c0 + 64 < NK) { if (c0 + 96 <= NK) tmem_ld32(base + c0 + 64, va); else tmem_ld16(base + c0 + 64, va); }
          tmem_ld_wait();
          emit(vb, c0 + 32); if (c0 + 48 < NK) emit(vb + 16, c0 + 48);
        }
      }
      tmem_st_wait();
    }
    acc += sum + mx;
  }
  const long long t1 = clock64();
  if (SPIN && threadIdx.x == 0) { done = 1; mbar_arrive(&never); }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free<512>(slot);
}

template <int MODE, int SPIN = 0>
void run(const char* name, int warps) {
  long long* out; float* sink; uint8_t* gsrc;
  cudaMalloc(&out, 8 * 148); cudaMalloc(&sink, 4); cudaMalloc(&gsrc, (size_t)148 * 64 * 16384);
  cudaFuncSetAttribute(bench<MODE, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int tiles = 200;
  bench<MODE, SPIN><<<148, (warps + (SPIN ? 3 : 0)) * 32, SPIN == 3 ? 80 * 1024 : 0>>>(out, tiles, sink, gsrc);
  bench<MODE, SPIN><<<148, (warps + (SPIN ? 3 : 0)) * 32, SPIN == 3 ? 80 * 1024 : 0>>>(out, tiles, sink, gsrc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-44s warps=%d: %8.0f cycles per 128x208 tile%s (%s)\n", name, warps, (double)h[0] / tiles, warps == 8 ? "-pair" : "", cudaGetErrorString(e));
  cudaFree(out); cudaFree(sink); cudaFree(gsrc);
}

int main() {
  for (int w = 4; w <= 8; w += 4) {
    run<6>("ld only (+add)", w);
    run<0>("max pass", w);
    run<4>("exp pass: fma+ex2+add, no pack, no store", w);
    run<5>("exp pass: + F2FP pack, no store", w);
    run<1>("exp pass: F2FP pack + tcgen05.st", w);
    run<2>("exp pass: integer RNE pack + tcgen05.st", w);
    run<3>("exp pass: truncating pack + tcgen05.st", w);
    run<7>("max + exp(F2FP) passes", w);
    run<0, 1>("max pass + 3 lanes spinning in try_wait", w);
    run<1, 1>("exp pass (F2FP+st) + 3 lanes spinning", w);
    run<0, 2>("max pass + 3 elected lanes spinning", w);
    run<0, 3>("max pass + bulk-copy stream into smem", w);
    run<1, 3>("exp pass (F2FP+st) + bulk-copy stream", w);
  }
  return 0;
}

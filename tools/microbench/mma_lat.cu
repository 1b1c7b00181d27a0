// Micro-benchmark (B200): latency / throughput of the tcgen05.mma groups the attention kernels issue, alone on an SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I dinov3-jax_b200/csrc -o tools/microbench/bin/mma_lat tools/microbench/mma_lat.cu
#include <cstdio>
#include "ptx.cuh"
using namespace d3;

// mode 0: S = Q K^T, SS, M128 x N x K64 (4 MMAs);  mode 1: O = P V, TS, M128 x 64 x Nk (Nk/16 MMAs), V MN-major
// mode 2: SS with MN-major A and B, M128 x 64 x 128 (8 MMAs)  (backward dV / dK shape)
template <bool ELECT>
__global__ void bench(long long* out, int mode, int N, int reps, int groups_in_flight) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (ELECT ? threadIdx.x < 32 : threadIdx.x == 0) {
    uint32_t phase = 0;
    long long best = 1ll << 60, total = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (!ELECT || elect_one()) {
      for (int g = 0; g < groups_in_flight; ++g) {
        const uint32_t d = tmem + (g & 1) * 256;
        if (mode == 0) {
          const uint64_t qd = umma_desc_sw128(smem_u32(smem), 16, 1024), kd = umma_desc_sw128(smem_u32(smem + 32768), 16, 1024);
          const uint32_t id = umma_idesc_bf16(128, N, 0, 0);
          for (int k = 0; k < 4; ++k) umma_bf16(d, qd + 2 * k, kd + 2 * k, id, k > 0);
        } else if (mode == 1) {
          const uint64_t vd = umma_desc_sw128(smem_u32(smem + 65536), 8192, 1024);
          const uint32_t id = umma_idesc_bf16(128, 64, 0, 1);
          for (int k = 0; k < N / 16; ++k) umma_bf16_ts(d + 128, d + k * 8, vd + (uint64_t)(k * 128), id, k > 0);
        } else {
          const uint64_t ad = umma_desc_sw128(smem_u32(smem), 16384, 1024), bd = umma_desc_sw128(smem_u32(smem + 65536), 8192, 1024);
          const uint32_t id = umma_idesc_bf16(128, 64, 1, 1);
          for (int k = 0; k < 8; ++k) umma_bf16(d, ad + (uint64_t)(k * 128), bd + (uint64_t)(k * 128), id, k > 0);
        }
      }
      umma_commit(&bar);
      }
      if (ELECT) __syncwarp();
      const long long t1 = clock64();
      mbar_wait(&bar, phase);
      phase ^= 1;
      const long long t2 = clock64();
      if (r > 0) { total += t2 - t0; if (t2 - t0 < best) best = t2 - t0; }
      if (r == reps - 1 && threadIdx.x == 0) { out[0] = best; out[1] = total / (reps - 1); out[2] = t1 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_free<512>(tmem);
}

int main() {
  long long* out; cudaMalloc(&out, 64);
  struct { int mode, N, groups; const char* name; } cases[] = {
      {0, 208, 1, "S  SS M128 N208 K64 (4 MMA)"}, {0, 208, 2, "S  x2 groups"}, {0, 208, 8, "S  x8 groups"},
      {0, 128, 1, "S  SS M128 N128 K64"}, {0, 128, 8, "S  N128 x8 groups"}, {0, 256, 8, "S  N256 x8 groups"}, {0, 112, 8, "S  N112 x8 groups"},
      {1, 208, 1, "PV TS M128 N64 K208 (13 MMA)"}, {1, 208, 2, "PV x2 groups"}, {1, 208, 8, "PV x8 groups"}, {1, 112, 8, "PV K112 x8 groups"},
      {2, 0, 1, "dV SS MNxMN M128 N64 K128 (8 MMA)"}, {2, 0, 8, "dV x8 groups"}};
  cudaFuncSetAttribute(bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bench<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int el = 0; el < 2; ++el)
  for (auto& c : cases) {
    if (el) bench<true><<<1, 128, 180 * 1024>>>(out, c.mode, c.N, 20, c.groups);
    else bench<false><<<1, 128, 180 * 1024>>>(out, c.mode, c.N, 20, c.groups);
    if (c.mode == 0 && c.N == 208 && c.groups == 1) printf(el ? "---- elect.sync in converged warp\n" : "---- if (threadIdx.x == 0)\n");
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-36s best %6lld cyc  mean %6lld  (issue %5lld)  per group %7.1f  %s\n", c.name, h[0], h[1], h[2], (double)h[0] / c.groups, cudaGetErrorString(e));
  }
  return 0;
}

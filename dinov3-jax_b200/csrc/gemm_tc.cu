// Persistent, warp-specialised bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA
// operand staging).  One kernel family covers every dense contraction on the DINOv3 training path:
//   forward  Y = X W        (A K-major, B MN-major: reference kernels are stored [in,out])
//   dgrad    dX = dY W^T    (A K-major, B K-major)
//   wgrad    dW = X^T dY    (A MN-major, B MN-major, fp32 output, optional accumulate)
// replacing nn.Dense / nn.Conv(stride=kernel) call sites of the reference:
//   dinov3_jax/layers/attention.py:63-65,94,101   dinov3_jax/layers/ffn_layers.py:36-47
//   dinov3_jax/layers/patch_embed.py:38-51        dinov3_jax/layers/dino_head.py:20-43,65-85
// The epilogue fuses bias, tanh-GELU, GELU', LayerScale (gamma) and the residual add
// (dinov3_jax/layers/block.py:198-199, dinov3_jax/layers/layer_scale.py:17-21).
//
// Roles (256 threads): warp0 = TMA producer, warp1 = MMA issuer (one elected lane), warp2 = TMEM allocator,
// warps4-7 = epilogue (TMEM -> registers -> global).  Two accumulator stages in TMEM (2 x BN columns) let the
// epilogue of tile i overlap the main loop of tile i+1.
#include "ptx.cuh"
#include "d3_internal.h"

namespace d3 {

constexpr int BM = 128;
constexpr int BK = 64;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(256, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmEpilogue ep, int M, int N, int K) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + Cfg::STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * Cfg::STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + BK - 1) / BK;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile % num_m) * BM;
        const int n0 = (tile / num_m) * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(&tmA, &full_bar[stage], sa + j * 8192, m0 + j * 64, kb * BK);
          } else {
            tma_load_2d(&tmA, &full_bar[stage], sa, kb * BK, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(&tmB, &full_bar[stage], sb + j * 8192, n0 + j * 64, kb * BK);
          } else {
            tma_load_2d(&tmB, &full_bar[stage], sb, kb * BK, n0);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      // per-UMMA_K (16 elements) advance of the descriptor start address, in 16-byte units
      constexpr uint32_t a_adv = A_MN ? (2 * 1024 >> 4) : (32 >> 4);
      constexpr uint32_t b_adv = B_MN ? (2 * 1024 >> 4) : (32 >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = A_MN ? umma_desc_sw128(sa, 8192, 1024) : umma_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = B_MN ? umma_desc_sw128(sb, 8192, 1024) : umma_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16(d_tmem, adesc + (uint64_t)(k * a_adv), bdesc + (uint64_t)(k * b_adv), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp - 4;  // TMEM lane quarter == warp % 4
    int local = 0;
    const bool fast = (ep.flags & EP_SLOW) == 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int m0 = (tile % num_m) * BM;
      const int n0 = (tile / num_m) * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int nc = n0 + c * 32;
        if (nc >= N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(t_addr + c * 32, r);
        tmem_ld_wait();
        if (!row_ok) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * ep.alpha;
        const bool full = fast && (nc + 32 <= N);
        const size_t ro = (size_t)row;
        if (full) {
          if (ep.flags & EP_BIAS) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 b = *reinterpret_cast<const float4*>(ep.bias + nc + j);
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (ep.flags & EP_STORE_PRE) {
            uint4* dst = reinterpret_cast<uint4*>(ep.aux_out + ro * ep.ld_aux + nc);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                  pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
          }
          if (ep.flags & EP_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(v[j]);
          }
          if (ep.flags & EP_MUL_DGELU) {
            const uint4* src = reinterpret_cast<const uint4*>(ep.aux_in + ro * ep.ld_aux + nc);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u = src[j];
              float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
              v[8 * j] *= gelu_tanh_grad(a.x); v[8 * j + 1] *= gelu_tanh_grad(a.y);
              v[8 * j + 2] *= gelu_tanh_grad(b.x); v[8 * j + 3] *= gelu_tanh_grad(b.y);
              v[8 * j + 4] *= gelu_tanh_grad(cc.x); v[8 * j + 5] *= gelu_tanh_grad(cc.y);
              v[8 * j + 6] *= gelu_tanh_grad(d.x); v[8 * j + 7] *= gelu_tanh_grad(d.y);
            }
          }
          if (ep.flags & EP_GAMMA) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 g = *reinterpret_cast<const float4*>(ep.gamma + nc + j);
              v[j] *= g.x; v[j + 1] *= g.y; v[j + 2] *= g.z; v[j + 3] *= g.w;
            }
          }
          if (ep.flags & EP_RESID) {
            const float4* rs = reinterpret_cast<const float4*>(ep.resid + ro * ep.ld_resid + nc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 x = rs[j];
              v[4 * j] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
            }
          }
          if (ep.flags & EP_OUT_F32) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + ro * ep.ld_out + nc);
            if (ep.flags & EP_ACCUM) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 x = dst[j];
                v[4 * j] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + ro * ep.ld_out + nc);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                  pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
          }
        } else {
          // ragged / unaligned edge: scalar path with bounds checks
#pragma unroll 1
          for (int j = 0; j < 32; ++j) {
            const int n = nc + j;
            if (n >= N) break;
            float x = v[j];
            if (ep.flags & EP_BIAS) x += ep.bias[n];
            if (ep.flags & EP_STORE_PRE) ep.aux_out[ro * ep.ld_aux + n] = __float2bfloat16(x);
            if (ep.flags & EP_GELU) x = gelu_tanh(x);
            if (ep.flags & EP_MUL_DGELU) x *= gelu_tanh_grad(__bfloat162float(ep.aux_in[ro * ep.ld_aux + n]));
            if (ep.flags & EP_GAMMA) x *= ep.gamma[n];
            if (ep.flags & EP_RESID) x += ep.resid[ro * ep.ld_resid + n];
            if (ep.flags & EP_OUT_F32) {
              float* o = reinterpret_cast<float*>(ep.out) + ro * ep.ld_out + n;
              if (ep.flags & EP_ACCUM) x += *o;
              *o = x;
            } else {
              reinterpret_cast<__nv_bfloat16*>(ep.out)[ro * ep.ld_out + n] = __float2bfloat16(x);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_free<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host side
static int make_operand_map(CUtensorMap* map, const void* ptr, int mn, int k, int ld, int is_mn_major, int box_mn) {
  // K-major : memory [mn][k], row stride ld   -> dims {k, mn}, box {64, box_mn}
  // MN-major: memory [k][mn], row stride ld   -> dims {mn, k}, box {64, 64}
  cuuint64_t dims[2], strides[1];
  cuuint32_t box[2], estr[2] = {1, 1};
  if (is_mn_major) {
    dims[0] = (cuuint64_t)mn; dims[1] = (cuuint64_t)k; box[0] = 64; box[1] = 64;
  } else {
    dims[0] = (cuuint64_t)k; dims[1] = (cuuint64_t)mn; box[0] = 64; box[1] = (cuuint32_t)box_mn;
  }
  strides[0] = (cuuint64_t)ld * 2;
  return encode_tensor_map_2d_bf16(map, ptr, dims, strides, box, estr);
}

template <int BN, int A_MN, int B_MN>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N, int K,
                      cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_tc_kernel<BN, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(D3_ERR_CUDA, cudaGetErrorString(e));
    configured = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, 256, Cfg::SMEM_BYTES, stream>>>(ta, tb, ep, M, N, K);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return set_error(D3_ERR_CUDA, cudaGetErrorString(e));
  count_launch();
  return D3_OK;
}

template <int BN>
static int launch_major(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep,
                        int M, int N, int K, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_cfg<BN, 0, 0>(ta, tb, ep, M, N, K, s);
  if (!a_mn && b_mn) return launch_cfg<BN, 0, 1>(ta, tb, ep, M, N, K, s);
  if (a_mn && !b_mn) return launch_cfg<BN, 1, 0>(ta, tb, ep, M, N, K, s);
  return launch_cfg<BN, 1, 1>(ta, tb, ep, M, N, K, s);
}

int gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int N, int K,
              GemmEpilogue ep, int force_bn, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(D3_ERR_ARG, "gemm: empty problem");
  if ((lda % 8) || (ldb % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
    return set_error(D3_ERR_ARG, "gemm: operands must be 16-byte aligned with ld % 8 == 0");
  // tile-N choice: minimise (waves x tile cost)
  int bn = force_bn;
  if (bn == 0) {
    const int cand[3] = {256, 128, 64};
    long best = -1;
    for (int i = 0; i < 3; ++i) {
      long tiles = (long)((M + BM - 1) / BM) * ((N + cand[i] - 1) / cand[i]);
      long waves = (tiles + sm_count() - 1) / sm_count();
      long cost = waves * (cand[i] + 24);
      if (best < 0 || cost < best) { best = cost; bn = cand[i]; }
    }
  }
  // fast (vectorised) epilogue needs 16-byte aligned rows everywhere
  const int out_elt = (ep.flags & EP_OUT_F32) ? 4 : 2;
  bool aligned = ((uintptr_t)ep.out % 16 == 0) && ((ep.ld_out * out_elt) % 16 == 0);
  if (ep.flags & EP_BIAS) aligned = aligned && ((uintptr_t)ep.bias % 16 == 0);
  if (ep.flags & EP_GAMMA) aligned = aligned && ((uintptr_t)ep.gamma % 16 == 0);
  if (ep.flags & EP_RESID) aligned = aligned && ((uintptr_t)ep.resid % 16 == 0) && (ep.ld_resid % 4 == 0);
  if (ep.flags & EP_STORE_PRE) aligned = aligned && ((uintptr_t)ep.aux_out % 16 == 0) && (ep.ld_aux % 8 == 0);
  if (ep.flags & EP_MUL_DGELU) aligned = aligned && ((uintptr_t)ep.aux_in % 16 == 0) && (ep.ld_aux % 8 == 0);
  if (!aligned) ep.flags |= EP_SLOW;

  CUtensorMap ta, tb;
  int rc = make_operand_map(&ta, A, M, K, lda, a_mn, BM);
  if (rc) return rc;
  rc = make_operand_map(&tb, B, N, K, ldb, b_mn, bn);
  if (rc) return rc;
  switch (bn) {
    case 256: return launch_major<256>(a_mn, b_mn, ta, tb, ep, M, N, K, stream);
    case 128: return launch_major<128>(a_mn, b_mn, ta, tb, ep, M, N, K, stream);
    case 64: return launch_major<64>(a_mn, b_mn, ta, tb, ep, M, N, K, stream);
  }
  return set_error(D3_ERR_ARG, "gemm: bad tile N");
}

}  // namespace d3

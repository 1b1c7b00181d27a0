// Internal declarations shared by the kernels behind the C ABI in include/dinov3_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/dinov3_b200.h"

namespace d3 {

// epilogue flag bits mirror D3_EP_* in the public header
enum : int {
  EP_BIAS = D3_EP_BIAS,
  EP_GELU = D3_EP_GELU,
  EP_STORE_PRE = D3_EP_STORE_PRE,
  EP_MUL_DGELU = D3_EP_MUL_DGELU,
  EP_GAMMA = D3_EP_GAMMA,
  EP_RESID = D3_EP_RESID,
  EP_OUT_F32 = D3_EP_OUT_F32,
  EP_ACCUM = D3_EP_ACCUM,
  EP_SCATTER = D3_EP_SCATTER,
  EP_SLOW = 1 << 30,      // internal: force the bounds-checked scalar epilogue
  EP_ATOMIC = 1 << 29,    // internal: split-K partial sums, fp32 atomic reduction into out
  EP_DEBUG_NOSTORE = 1 << 27,  // internal (D3_GEMM_DEBUG=1): skip epilogue global traffic
  EP_M_FASTEST = 1 << 26,      // internal (D3_GEMM_DEBUG=2): M-fastest tile order
  EP_TMA_EPI = 1 << 25,        // internal: epilogue tiles move through swizzled smem staging + TMA store / load
  EP_FAST_ACT = 1 << 28,  // internal: hardware tanh in GELU / GELU' (bf16-rounded outputs)
};

struct GemmEpilogue {
  const float* bias;             // [N] fp32
  const float* gamma;            // [N] fp32 (LayerScale)
  const float* resid;            // [M, ld_resid] fp32 residual stream (may alias out)
  const __nv_bfloat16* aux_in;   // [M, ld_aux] bf16 pre-activation for GELU'
  __nv_bfloat16* aux_out;        // [M, ld_aux] bf16 pre-activation stash
  void* out;                     // [M, ld_out] bf16 or fp32
  int ld_out, ld_aux, ld_resid;
  int flags;
  float alpha;
  // EP_SCATTER: the fp32 result is not stored at `out` but added (red.add over NVLink peer mappings) into the rank
  // that owns it: element e of the output ([row*ld_out + col]) has global index g = sc_off + e inside a flat range cut
  // into sc_world slices of sc_shard elements; it goes to sc_peer[g / sc_shard][g % sc_shard].
  float* sc_peer[8];
  long long sc_off;
  int sc_shard, sc_world;
  int sc_sys;      // 1: scalar system-scope atomics (atomicAdd_system) instead of one device-scope vector red
};

int set_error(int code, const char* msg);
int scatter_mode();            // 0: one vector device-scope red per float4 (default); 1: four scalar system-scope atomics
void set_scatter_mode(int mode);
int sm_count();
void count_launch(int n = 1);
int encode_tensor_map_2d_bf16(CUtensorMap* map, const void* ptr, const cuuint64_t dims[2],
                              const cuuint64_t strides[1], const cuuint32_t box[2], const cuuint32_t estr[2]);
// generic 2-D map: elt_bytes 2 (bf16) or 4 (fp32); swizzle_bytes 0 / 64 / 128
int encode_tensor_map_2d(CUtensorMap* map, const void* ptr, int elt_bytes, cuuint64_t cols, cuuint64_t rows,
                         cuuint64_t row_stride_bytes, cuuint32_t box_cols, cuuint32_t box_rows, int swizzle_bytes);

int gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int N, int K,
              GemmEpilogue ep, int tile_n, int split_k, cudaStream_t stream);

// warp-specialised persistent attention (attention_ws.cu); *handled = 0 when the shape must take the single-pass kernels
int attn_fwd_ws(const void* qkv, void* o, float* lse, int n_crops, int N, int D, int H, cudaStream_t st, int* handled);

void attn_ws_set_trace(long long* buf);

#define D3_CHECK_LAUNCH()                                               \
  do {                                                                  \
    cudaError_t e__ = cudaPeekAtLastError();                            \
    if (e__ != cudaSuccess) return d3::set_error(D3_ERR_CUDA, cudaGetErrorString(e__)); \
    d3::count_launch();                                                 \
  } while (0)

}  // namespace d3

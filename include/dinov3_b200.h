/* dinov3_b200.h — C ABI of libdinov3_b200.so: the B200 (sm_100a) kernels behind the DINOv3 SSL training hot path.
 *
 * The reference (Dhia-naouali/dinov3-jax) has no FFI: its hot path is Flax modules traced by jax.jit.  Each entry
 * point below therefore cites the reference *module call site* (path:line under dinov3_jax/) whose arithmetic it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer (torch.Tensor.data_ptr()); the library allocates nothing
 *     persistent; `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream.
 *   - return value: D3_OK (0) or a negative d3_status; d3_last_error() gives the message for the calling thread.
 *     Launch-configuration errors are reported synchronously, asynchronous faults surface at the next sync.
 *   - one process per GPU, calls come from that process' single training thread.
 *   - there is NO CPU fallback: without a Blackwell GPU d3_init() fails.
 */
#ifndef DINOV3_B200_H
#define DINOV3_B200_H
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  D3_OK = 0,
  D3_ERR_ARG = -1,     /* bad shape / alignment / null pointer */
  D3_ERR_CUDA = -2,    /* CUDA runtime or driver error (message has the CUDA string) */
  D3_ERR_DEVICE = -3,  /* not an sm_100 device */
} d3_status;

/* ---- library ------------------------------------------------------------------------------------------------ */
int d3_init(int device);                /* bind to device, cache SM count, resolve cuTensorMapEncodeTiled */
const char* d3_last_error(void);
int d3_abi_version(void);
long long d3_launch_count(void);        /* kernels launched by this library since the last reset (bench: gpu_launches) */
void d3_reset_launch_count(void);

/* ---- dense contraction (tcgen05 / TMEM / TMA) ------------------------------------------------------------------
 * D[M,N] = epilogue( alpha * A[M,K] . B[K,N] ), bf16 operands, fp32 accumulation in tensor memory.
 *   a_major = 0: A stored [M][K] (row stride lda)      a_major = 1: A stored [K][M]
 *   b_major = 0: B stored [N][K] (row stride ldb)      b_major = 1: B stored [K][N]   (reference kernel layout [in,out])
 * Replaces nn.Dense / nn.Conv(stride=kernel) at layers/attention.py:63-65,94,101, layers/ffn_layers.py:36-47,
 * layers/patch_embed.py:38-51, layers/dino_head.py:20-43,65-85, and their jax.grad transposes (train/train.py:504-513).
 * Epilogue order: +bias -> [store bf16 pre-activation] -> [tanh-GELU] -> [* GELU'(aux_in)] -> [* gamma] -> [+ resid]
 *                 -> [+= out] -> store (bf16 or fp32).   (layers/block.py:198-199, layers/layer_scale.py:17-21)     */
enum {
  D3_EP_BIAS = 1,       /* v += bias[n]                       (fp32 [N]) */
  D3_EP_GELU = 2,       /* v = gelu_tanh(v)                   flax nn.gelu, approximate=True */
  D3_EP_STORE_PRE = 4,  /* aux_out[m,n] = bf16(v) before the activation (stash for backward) */
  D3_EP_MUL_DGELU = 8,  /* v *= gelu_tanh'(aux_in[m,n])       backward through an activation */
  D3_EP_GAMMA = 16,     /* v *= gamma[n]                      LayerScale */
  D3_EP_RESID = 32,     /* v += resid[m,n]                    fp32 residual stream */
  D3_EP_OUT_F32 = 64,   /* out is fp32 (default bf16) */
  D3_EP_ACCUM = 128,    /* out += v (fp32 out only; weight-gradient accumulation over crop sets) */
};
typedef struct {
  const float* bias;
  const float* gamma;
  const float* resid;
  const void* aux_in;   /* bf16 [M, ld_aux] */
  void* aux_out;        /* bf16 [M, ld_aux] */
  void* out;            /* bf16 or fp32 [M, ld_out] */
  int ld_out, ld_aux, ld_resid;
  int flags;
  float alpha;
} d3_gemm_epilogue;
int d3_gemm_bf16(const void* A, int lda, int a_major, const void* B, int ldb, int b_major, int M, int N, int K,
                 const d3_gemm_epilogue* ep, int tile_n /*0 = auto; 64/128/256*/, void* stream);

#ifdef __cplusplus
}
#endif
#endif

// On-GPU DINO multi-crop augmentation: the step BEFORE the training hot path.  Replaces the per-sample
// torchvision / PIL host pipeline of dinov3_jax/data/augmentations.py:23-230 (RandomResizedCrop(bicubic) + flip,
// ColorJitter(0.4, 0.4, 0.2, 0.1) in random order, RandomGrayscale, GaussianBlur(9, sigma 0.1..2), RandomSolarize(128),
// ToTensor + Normalize) for a whole batch of decoded uint8 images resident in HBM; the random parameters are drawn on the
// host (a few scalars per crop, dinov3_jax/data/gpu_augment.py) so that the kernels are deterministic functions that can
// be checked against torchvision's float implementations.  Output: crop-major NHWC bf16, i.e. exactly the
// `collated_global_crops` / `collated_local_crops` tensors of data/collate.py:72-93.
//
// All kernels are HBM / L2 streaming work (one thread per output pixel, channels innermost); arithmetic is fp32 on
// [0, 1] images like torchvision.transforms.v2.functional on float tensors (PIL's per-op uint8 re-quantisation is not
// reproduced).
#include "ptx.cuh"
#include "d3_internal.h"

namespace d3 {

// one record per output crop (host-filled, 64 bytes)
struct AugCrop {
  int img;                 // source image index
  int x0, y0, w, h;        // crop box in the source (RandomResizedCrop.get_params)
  int flip;                // horizontal flip
  int order[4];            // ColorJitter op order: 0 brightness, 1 contrast, 2 saturation, 3 hue; -1 = jitter not applied
  float fb, fc, fs, fh;    // factors
  int gray;                // RandomGrayscale applied
  int solarize;            // RandomSolarize applied (threshold 128/255)
};
struct AugBlur { float sigma; };   // <= 0: no blur

__device__ __forceinline__ float cubic_aa(float x) {           // Keys cubic, a = -0.5 (PIL / torch antialias bicubic)
  x = fabsf(x);
  if (x < 1.f) return ((1.5f * x - 2.5f) * x) * x + 1.f;
  if (x < 2.f) return ((-0.5f * x + 2.5f) * x - 4.f) * x + 2.f;
  return 0.f;
}

// out[n, S, S, 3] (fp32, [0,1]) = antialiased bicubic resize of src[img, y0:y0+h, x0:x0+w] (+ horizontal flip).
// Same definition as torch's _upsample_bicubic2d_aa (align_corners = False): per axis, scale = in/out,
// support = 2 * max(scale, 1), taps j in [floor(center - support + 0.5), ...), weight cubic((j + 0.5 - center) / max(scale, 1)),
// normalised; taps are clipped to the crop box.
__global__ void aug_resized_crop_kernel(const uint8_t* __restrict__ src, int H, int W, const AugCrop* __restrict__ crops,
                                        float* __restrict__ out, int S) {
  const int n = blockIdx.z;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ox >= S || oy >= S) return;
  const AugCrop c = crops[n];
  const int sx_out = c.flip ? (S - 1 - ox) : ox;        // flip after resize == sample the mirrored column
  const float scx = (float)c.w / S, scy = (float)c.h / S;
  const float isx = 1.f / fmaxf(scx, 1.f), isy = 1.f / fmaxf(scy, 1.f);
  const float supx = 2.f * fmaxf(scx, 1.f), supy = 2.f * fmaxf(scy, 1.f);
  const float cx = scx * (sx_out + 0.5f), cy = scy * (oy + 0.5f);
  const int xmin = max((int)(cx - supx + 0.5f), 0), xmax = min((int)(cx + supx + 0.5f), c.w);
  const int ymin = max((int)(cy - supy + 0.5f), 0), ymax = min((int)(cy + supy + 0.5f), c.h);
  float wxs = 0.f, wys = 0.f;
  for (int x = xmin; x < xmax; ++x) wxs += cubic_aa((x - cx + 0.5f) * isx);
  for (int y = ymin; y < ymax; ++y) wys += cubic_aa((y - cy + 0.5f) * isy);
  const uint8_t* base = src + ((size_t)c.img * H + c.y0) * W * 3 + (size_t)c.x0 * 3;
  float r = 0.f, g = 0.f, b = 0.f;
  for (int y = ymin; y < ymax; ++y) {
    const float wy = cubic_aa((y - cy + 0.5f) * isy);
    const uint8_t* row = base + (size_t)y * W * 3;
    float rr = 0.f, gg = 0.f, bb = 0.f;
    for (int x = xmin; x < xmax; ++x) {
      const float wx = cubic_aa((x - cx + 0.5f) * isx);
      rr += wx * row[3 * x]; gg += wx * row[3 * x + 1]; bb += wx * row[3 * x + 2];
    }
    r += wy * rr; g += wy * gg; b += wy * bb;
  }
  const float norm = 1.f / (255.f * wxs * wys);
  float* o = out + (((size_t)n * S + oy) * S + ox) * 3;
  // bicubic overshoots are clamped like a uint8 image would (PIL result is uint8)
  o[0] = fminf(fmaxf(r * norm, 0.f), 1.f); o[1] = fminf(fmaxf(g * norm, 0.f), 1.f); o[2] = fminf(fmaxf(b * norm, 0.f), 1.f);
}

// torchvision _rgb_to_grayscale_image weights
__device__ __forceinline__ float gray_of(float r, float g, float b) { return 0.2989f * r + 0.587f * g + 0.114f * b; }
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

// torchvision _rgb2hsv / _hsv2rgb on one pixel, hue shifted by fh (adjust_hue)
__device__ __forceinline__ void hue_shift(float& r, float& g, float& b, float fh) {
  const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
  const bool eqc = maxc == minc;
  const float cr = maxc - minc;
  const float s = cr / (eqc ? 1.f : maxc);
  const float crd = eqc ? 1.f : cr;
  const float rc = (maxc - r) / crd, gc = (maxc - g) / crd, bc = (maxc - b) / crd;
  float h = 0.f;
  if (maxc == r) h = bc - gc;
  else if (maxc == g) h = 2.f + rc - bc;
  else h = 4.f + gc - rc;
  h = h / 6.f + 1.f;
  h = h - floorf(h);
  h = h + fh;
  h = h - floorf(h);
  const float v = maxc;
  const float i = floorf(h * 6.f);
  const float f = h * 6.f - i;
  const int ii = ((int)i) % 6;
  const float p = clamp01(v * (1.f - s)), q = clamp01(v * (1.f - s * f)), t = clamp01(v * (1.f - s * (1.f - f)));
  switch (ii) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

__device__ __forceinline__ void jitter_op(int op, const AugCrop& c, float mean_gray, float& r, float& g, float& b) {
  if (op == 0) { r = clamp01(r * c.fb); g = clamp01(g * c.fb); b = clamp01(b * c.fb); }
  else if (op == 1) { const float m = (1.f - c.fc) * mean_gray; r = clamp01(c.fc * r + m); g = clamp01(c.fc * g + m); b = clamp01(c.fc * b + m); }
  else if (op == 2) { const float m = (1.f - c.fs) * gray_of(r, g, b); r = clamp01(c.fs * r + m); g = clamp01(c.fs * g + m); b = clamp01(c.fs * b + m); }
  else if (op == 3) hue_shift(r, g, b, c.fh);
}

// Pass A: the jitter ops that come BEFORE contrast in this crop's order, plus the sum of the gray values of the result
// (contrast blends with the mean gray of the image it is applied to).  Pass B: contrast and what follows, then
// RandomGrayscale.  A crop without jitter / without a pending contrast simply passes through pass A untouched.
__global__ void aug_color_a_kernel(float* __restrict__ x, const AugCrop* __restrict__ crops, float* __restrict__ gray_sum,
                                   int S) {
  const int n = blockIdx.y;
  const AugCrop c = crops[n];
  const int npix = S * S;
  float local = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    float* px = x + ((size_t)n * npix + p) * 3;
    float r = px[0], g = px[1], b = px[2];
    if (c.order[0] >= 0)
      for (int k = 0; k < 4 && c.order[k] != 1; ++k) jitter_op(c.order[k], c, 0.f, r, g, b);
    px[0] = r; px[1] = g; px[2] = b;
    local += gray_of(r, g, b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&gray_sum[n], local);
}
__global__ void aug_color_b_kernel(float* __restrict__ x, const AugCrop* __restrict__ crops,
                                   const float* __restrict__ gray_sum, int S) {
  const int n = blockIdx.y;
  const AugCrop c = crops[n];
  const int npix = S * S;
  const float mean_gray = gray_sum[n] / npix;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    float* px = x + ((size_t)n * npix + p) * 3;
    float r = px[0], g = px[1], b = px[2];
    if (c.order[0] >= 0) {
      int k = 0;
      while (k < 4 && c.order[k] != 1) ++k;
      for (; k < 4; ++k) jitter_op(c.order[k], c, mean_gray, r, g, b);
    }
    if (c.gray) { const float y = gray_of(r, g, b); r = g = b = y; }
    px[0] = r; px[1] = g; px[2] = b;
  }
}

// separable 9-tap Gaussian (torchvision gaussian_blur: kernel_size 9, reflect padding), one axis per launch
__global__ void aug_blur_kernel(const float* __restrict__ x, float* __restrict__ y, const AugBlur* __restrict__ blur, int S,
                                int vertical) {
  const int n = blockIdx.z;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ox >= S || oy >= S) return;
  const float sigma = blur[n].sigma;
  const float* xi = x + (size_t)n * S * S * 3;
  float* yo = y + (((size_t)n * S + oy) * S + ox) * 3;
  if (sigma <= 0.f) {
    const float* p = xi + ((size_t)oy * S + ox) * 3;
    yo[0] = p[0]; yo[1] = p[1]; yo[2] = p[2];
    return;
  }
  float w[9], ws = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { const float d = (float)(k - 4) / sigma; w[k] = __expf(-0.5f * d * d); ws += w[k]; }
  float r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    int t = (vertical ? oy : ox) + k - 4;
    t = t < 0 ? -t : (t >= S ? 2 * S - 2 - t : t);           // reflect (no edge repeat), S >= 5
    const float* p = xi + (vertical ? ((size_t)t * S + ox) : ((size_t)oy * S + t)) * 3;
    r += w[k] * p[0]; g += w[k] * p[1]; b += w[k] * p[2];
  }
  const float inv = 1.f / ws;
  yo[0] = r * inv; yo[1] = g * inv; yo[2] = b * inv;
}

// RandomSolarize (pixels >= 128/255 inverted), Normalize(mean, std), cast to bf16 NHWC
__global__ void aug_finish_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                  const AugCrop* __restrict__ crops, int S, float m0, float m1, float m2, float is0,
                                  float is1, float is2) {
  const int n = blockIdx.y;
  const int sol = crops[n].solarize;
  const int npix = S * S;
  const float thr = 128.f / 255.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const float* px = x + ((size_t)n * npix + p) * 3;
    float r = px[0], g = px[1], b = px[2];
    if (sol) { r = r >= thr ? 1.f - r : r; g = g >= thr ? 1.f - g : g; b = b >= thr ? 1.f - b : b; }
    __nv_bfloat16* o = out + ((size_t)n * npix + p) * 3;
    o[0] = __float2bfloat16((r - m0) * is0); o[1] = __float2bfloat16((g - m1) * is1); o[2] = __float2bfloat16((b - m2) * is2);
  }
}

}  // namespace d3

using namespace d3;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int d3_aug_resized_crop(const void* src_u8, int n_img, int H, int W, const void* crops, int n_crops, float* out, int S,
                        void* stream) {
  if (n_crops <= 0) return D3_OK;
  if (S < 5 || H <= 0 || W <= 0 || n_img <= 0) return set_error(D3_ERR_ARG, "d3_aug_resized_crop: bad geometry");
  dim3 block(32, 8), grid((S + 31) / 32, (S + 7) / 8, n_crops);
  aug_resized_crop_kernel<<<grid, block, 0, STREAM(stream)>>>((const uint8_t*)src_u8, H, W, (const AugCrop*)crops, out, S);
  D3_CHECK_LAUNCH();
  return D3_OK;
}

int d3_aug_color(float* x, const void* crops, int n_crops, int S, float* gray_sum /* [n_crops] zeroed */, void* stream) {
  if (n_crops <= 0) return D3_OK;
  dim3 grid(min((S * S + 255) / 256, 64), n_crops);
  aug_color_a_kernel<<<grid, 256, 0, STREAM(stream)>>>(x, (const AugCrop*)crops, gray_sum, S);
  aug_color_b_kernel<<<grid, 256, 0, STREAM(stream)>>>(x, (const AugCrop*)crops, gray_sum, S);
  D3_CHECK_LAUNCH();
  count_launch(1);
  return D3_OK;
}

int d3_aug_blur(const float* x, float* tmp, float* y, const void* blur, int n_crops, int S, void* stream) {
  if (n_crops <= 0) return D3_OK;
  if (S < 5) return set_error(D3_ERR_ARG, "d3_aug_blur: S < 5");
  dim3 block(32, 8), grid((S + 31) / 32, (S + 7) / 8, n_crops);
  aug_blur_kernel<<<grid, block, 0, STREAM(stream)>>>(x, tmp, (const AugBlur*)blur, S, 0);
  aug_blur_kernel<<<grid, block, 0, STREAM(stream)>>>(tmp, y, (const AugBlur*)blur, S, 1);
  D3_CHECK_LAUNCH();
  count_launch(1);
  return D3_OK;
}

int d3_aug_finish(const float* x, void* out_bf16, const void* crops, int n_crops, int S, const float* mean3,
                  const float* std3, void* stream) {
  if (n_crops <= 0) return D3_OK;
  dim3 grid(min((S * S + 255) / 256, 64), n_crops);
  aug_finish_kernel<<<grid, 256, 0, STREAM(stream)>>>(x, (__nv_bfloat16*)out_bf16, (const AugCrop*)crops, S, mean3[0], mean3[1],
                                                     mean3[2], 1.f / std3[0], 1.f / std3[1], 1.f / std3[2]);
  D3_CHECK_LAUNCH();
  return D3_OK;
}

}  // extern "C"

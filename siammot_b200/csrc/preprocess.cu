// Test-time frame preprocessing on the device: the reference resizes every
// decoded frame on the CPU with PIL (torchvision F.resize, image_augmentation.py:44-46), converts it to float
// (ToTensor) and normalises it (maskrcnn_benchmark Normalize, build_augmentation.py:52-66) before a 10.8 MB
// pageable host->device copy.  Here the uint8 frame (2.8 MB at 720p) is copied and everything else runs on the GPU,
// bit-exactly like the CPU chain:
//   * Pillow's 8-bit resampling (libImaging/Resample.c) is integer arithmetic: antialiased triangle filter,
//     coefficients rounded to 22-bit fixed point, a horizontal pass and a vertical pass with a uint8 image in
//     between, each pass only when that dimension changes -> reproduced with the same integers;
//   * ToTensor / Normalize are single IEEE fp32 operations (x/255, [*255], -mean, /std) -> reproduced with
//     __fdiv_rn / __fmul_rn / __fsub_rn in the same order.
// Both kernels are HBM-bound byte work (one read of the frame, one write of the tensor).
#include <math.h>

#include "common.cuh"

namespace smot {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ int clip8(int v) {
  v >>= RS_PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: out[y][ox][c] from in[y][xmin .. xmin+n)[c]
__global__ void __launch_bounds__(256) resample_h_u8_kernel(const uint8_t* __restrict__ in, int in_pitch, int H, int OW,
                                                            const int* __restrict__ bounds, const int* __restrict__ kk,
                                                            int ksize, uint8_t* __restrict__ out, int out_pitch) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (ox >= OW) return;
  const int x0 = bounds[2 * ox], n = bounds[2 * ox + 1];
  const int* k = kk + (size_t)ox * ksize;
  const uint8_t* src = in + (size_t)y * in_pitch + (size_t)x0 * 3;
  int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int t = 0; t < n; ++t) {
    const int w = k[t];
    s0 += (int)src[3 * t + 0] * w;
    s1 += (int)src[3 * t + 1] * w;
    s2 += (int)src[3 * t + 2] * w;
  }
  uint8_t* dst = out + (size_t)y * out_pitch + (size_t)ox * 3;
  dst[0] = (uint8_t)clip8(s0), dst[1] = (uint8_t)clip8(s1), dst[2] = (uint8_t)clip8(s2);
}

struct NormArgs {
  float mean[3], std[3];
  int to_bgr255;
};

// vertical pass (identity table when the height does not change) fused with ToTensor + Normalize:
// out[c'][oy][x] (float32 planes) from in[ymin .. ymin+n)[x][c]
__global__ void __launch_bounds__(256) resample_v_normalize_kernel(const uint8_t* __restrict__ in, int in_pitch, int W, int OH,
                                                                   const int* __restrict__ bounds, const int* __restrict__ kk,
                                                                   int ksize, NormArgs na, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
  if (x >= W) return;
  int v[3];
  if (bounds) {
    const int y0 = bounds[2 * oy], n = bounds[2 * oy + 1];
    const int* k = kk + (size_t)oy * ksize;
    int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int t = 0; t < n; ++t) {
      const uint8_t* src = in + (size_t)(y0 + t) * in_pitch + (size_t)x * 3;
      const int w = k[t];
      s0 += (int)src[0] * w;
      s1 += (int)src[1] * w;
      s2 += (int)src[2] * w;
    }
    v[0] = clip8(s0), v[1] = clip8(s1), v[2] = clip8(s2);
  } else {
    const uint8_t* src = in + (size_t)oy * in_pitch + (size_t)x * 3;
    v[0] = src[0], v[1] = src[1], v[2] = src[2];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float f = __fdiv_rn((float)v[na.to_bgr255 ? 2 - c : c], 255.f);   // ToTensor
    if (na.to_bgr255) f = __fmul_rn(f, 255.f);                        // image[[2, 1, 0]] * 255
    out[((size_t)c * OH + oy) * W + x] = __fdiv_rn(__fsub_rn(f, na.mean[c]), na.std[c]);
  }
}

static double bilinear_filter(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? 1.0 - x : 0.0;
}

}  // namespace smot

using namespace smot;

// Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc (bilinear, box = whole image), in double like the original
extern "C" int smot_resample_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return 0;
  double filterscale = (double)in_size / (double)out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  return (int)ceil(1.0 * filterscale) * 2 + 1;
}

extern "C" int smot_resample_coeffs(int in_size, int out_size, int* bounds, int* kk) {
  SMOT_CHECK_ARG(in_size > 0 && out_size > 0 && bounds && kk, "smot_resample_coeffs: bad arguments");
  const double scale = (double)in_size / (double)out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double w[64];
    SMOT_CHECK_ARG(ksize <= 64, "smot_resample_coeffs: reduction factor too large (ksize %d)", ksize);
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bilinear_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    int* k = kk + (size_t)xx * ksize;
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      k[x] = v < 0 ? (int)(-0.5 + v * (1 << RS_PRECISION_BITS)) : (int)(0.5 + v * (1 << RS_PRECISION_BITS));
    }
    for (int x = xmax; x < ksize; ++x) k[x] = 0;
    bounds[2 * xx] = xmin, bounds[2 * xx + 1] = xmax;
  }
  return SMOT_OK;
}

extern "C" int smot_resample_h_u8(const void* in, int in_pitch, int H, int W, const int* bounds, const int* kk, int ksize,
                                  int OW, void* out, int out_pitch, void* stream) {
  SMOT_CHECK_ARG(in && out && bounds && kk && H > 0 && W > 0 && OW > 0 && ksize > 0 && in_pitch >= 3 * W && out_pitch >= 3 * OW,
                 "smot_resample_h_u8: bad arguments");
  SMOT_CHECK_ARG(H <= 65535, "smot_resample_h_u8: H %d too large", H);
  resample_h_u8_kernel<<<dim3((OW + 255) / 256, H), 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)in, in_pitch, H, OW, bounds, kk, ksize, (uint8_t*)out, out_pitch);
  SMOT_CHECK_LAUNCH("smot_resample_h_u8");
  return SMOT_OK;
}

extern "C" int smot_resample_v_normalize(const void* in, int in_pitch, int H, int W, const int* bounds, const int* kk,
                                         int ksize, int OH, const float* mean3, const float* std3, int to_bgr255,
                                         float* out_chw, void* stream) {
  SMOT_CHECK_ARG(in && out_chw && mean3 && std3 && H > 0 && W > 0 && OH > 0 && in_pitch >= 3 * W,
                 "smot_resample_v_normalize: bad arguments");
  SMOT_CHECK_ARG((bounds && kk && ksize > 0) || (!bounds && OH == H), "smot_resample_v_normalize: no table given but OH != H");
  SMOT_CHECK_ARG(OH <= 65535, "smot_resample_v_normalize: OH %d too large", OH);
  NormArgs na;
  for (int c = 0; c < 3; ++c) na.mean[c] = mean3[c], na.std[c] = std3[c];
  na.to_bgr255 = to_bgr255;
  resample_v_normalize_kernel<<<dim3((W + 255) / 256, OH), 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)in, in_pitch, W, OH, bounds, kk, ksize, na, out_chw);
  SMOT_CHECK_LAUNCH("smot_resample_v_normalize");
  return SMOT_OK;
}

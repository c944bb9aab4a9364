// Small NHWC tensor kernels: layout change, 2x2 max-pool, FPN top-down merge, P6 subsample,
// GroupNorm+ReLU.  All are HBM/latency bound: one thread per 4 channels, 16B/8B vector accesses.
#include "common.cuh"

namespace smot {

// ---- CHW fp32 image -> NHWC ---------------------------------------------------------------
template <typename T>
__global__ void image_to_nhwc_kernel(const float* __restrict__ chw, T* __restrict__ out, int C, int HW, int ld) {
  pdl_launch_dependents();
  pdl_wait();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  for (int c = 0; c < C; ++c) out[(size_t)p * ld + c] = from_f<T>(chw[(size_t)c * HW + p]);
  for (int c = C; c < ld; ++c) out[(size_t)p * ld + c] = from_f<T>(0.f);
}

// ---- MaxPool2d(2,2) --------------------------------------------------------------------------
template <typename T>
__global__ void maxpool2x2_kernel(const T* __restrict__ in, T* __restrict__ out, int batch, int H, int W, int C,
                                  int in_ld, int out_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int OH = H / 2, OW = W / 2, C4 = C / 4;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)batch * OH * OW * C4;
  if (idx >= total) return;
  int c = (int)(idx % C4) * 4;
  size_t pix = idx / C4;
  int ow = (int)(pix % OW);
  int oh = (int)((pix / OW) % OH);
  int n = (int)(pix / ((size_t)OW * OH));
  const T* base = in + (((size_t)n * H + 2 * oh) * W + 2 * ow) * in_ld + c;
  float4 a = ld4(base), b = ld4(base + in_ld), d = ld4(base + (size_t)W * in_ld), e = ld4(base + (size_t)(W + 1) * in_ld);
  float4 r = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y)),
                         fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w)));
  st4(out + pix * out_ld + c, r);
}

// fp16, 8 channels (16 bytes) per thread: the max of fp16 values is exact in fp16, so no conversion at all
__global__ void maxpool2x2_h8_kernel(const __half* __restrict__ in, __half* __restrict__ out, int batch, int H, int W, int C,
                                     int in_ld, int out_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int OH = H / 2, OW = W / 2, C8 = C / 8;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)batch * OH * OW * C8) return;
  const int c = (int)(idx % C8) * 8;
  const size_t pix = idx / C8;
  const int ow = (int)(pix % OW), oh = (int)((pix / OW) % OH), n = (int)(pix / ((size_t)OW * OH));
  const __half* base = in + (((size_t)n * H + 2 * oh) * W + 2 * ow) * in_ld + c;
  uint4 a = *reinterpret_cast<const uint4*>(base), b = *reinterpret_cast<const uint4*>(base + in_ld);
  uint4 d = *reinterpret_cast<const uint4*>(base + (size_t)W * in_ld), e = *reinterpret_cast<const uint4*>(base + (size_t)(W + 1) * in_ld);
  uint4 r;
  const __half2 *pa = reinterpret_cast<const __half2*>(&a), *pb = reinterpret_cast<const __half2*>(&b);
  const __half2 *pd = reinterpret_cast<const __half2*>(&d), *pe = reinterpret_cast<const __half2*>(&e);
  __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(__hmax2(pa[i], pb[i]), __hmax2(pd[i], pe[i]));
  *reinterpret_cast<uint4*>(out + pix * out_ld + c) = r;
}

// ---- max_pool2d(kernel 3, stride 2, padding 1): the ResNet stem pool (upstream resnet.py BaseStem.forward) --------
// Padding is "-inf" (taps outside the map do not take part), OH = (H - 1) / 2 + 1.  One thread = 4 channels of one
// output pixel; the max of storage-type values is exact, so fp16 maps go through float only as a container.
template <typename T>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ in, T* __restrict__ out, int batch, int H, int W, int C,
                                    int in_ld, int out_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1, C4 = C / 4;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)batch * OH * OW * C4) return;
  const int c = (int)(idx % C4) * 4;
  const size_t pix = idx / C4;
  const int ow = (int)(pix % OW), oh = (int)((pix / OW) % OH), n = (int)(pix / ((size_t)OW * OH));
  const float ninf = -__int_as_float(0x7f800000);
  float4 r = make_float4(ninf, ninf, ninf, ninf);
  for (int dy = -1; dy <= 1; ++dy) {
    const int y = 2 * oh + dy;
    if (y < 0 || y >= H) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int x = 2 * ow + dx;
      if (x < 0 || x >= W) continue;
      const float4 v = ld4(in + (((size_t)n * H + y) * W + x) * in_ld + c);
      r.x = fmaxf(r.x, v.x), r.y = fmaxf(r.y, v.y), r.z = fmaxf(r.z, v.z), r.w = fmaxf(r.w, v.w);
    }
  }
  st4(out + pix * out_ld + c, r);
}

// ---- deformable 3x3 sampling (DCN v1: upstream layers/dcn deform_im2col, reached through DFConv2d at dla.py:74-78) ----
// cols[oy][ox][k*C + c] = bilinear sample of channel c at (oy*stride - 1 + i + dy, ox*stride - 1 + j + dx), k = 3i + j,
// (dy, dx) = offsets[oy][ox][2k], [2k+1]; a sample position outside (-1, H) x (-1, W) is zero and so is every corner outside
// the map.  The deformable convolution is then a plain GEMM over the 9*C "channels" of `cols` (smot_conv2d, 1x1) with the
// 3x3 weight viewed as [Cout][9*C] -- the gather is the only part that is not a dense contraction.
// One thread = 4 channels of one (pixel, tap); corner reads of neighbouring threads are contiguous.
template <typename T>
__global__ void deform_im2col3x3_kernel(const T* __restrict__ in, const float* __restrict__ off, T* __restrict__ cols, int H, int W,
                                        int C, int in_ld, int off_ld, int OH, int OW, int out_ld, int stride) {
  pdl_launch_dependents();
  pdl_wait();
  const int C4 = C / 4;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)OH * OW * 9 * C4) return;
  const int c = (int)(idx % C4) * 4;
  size_t r = idx / C4;
  const int k = (int)(r % 9);
  const size_t pix = r / 9;
  const int ox = (int)(pix % OW), oy = (int)(pix / OW);
  const float* o = off + pix * off_ld + 2 * k;
  const float y = (float)(oy * stride - 1 + k / 3) + o[0];
  const float x = (float)(ox * stride - 1 + k % 3) + o[1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (y > -1.f && y < (float)H && x > -1.f && x < (float)W) {
    const float yf = floorf(y), xf = floorf(x);
    const int y0 = (int)yf, x0 = (int)xf;
    const float ly = y - yf, lx = x - xf, hy = 1.f - ly, hx = 1.f - lx;
    const float w[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yi = y0 + (q >> 1), xi = x0 + (q & 1);
      if (yi >= 0 && yi < H && xi >= 0 && xi < W) {
        const float4 v = ld4(in + ((size_t)yi * W + xi) * in_ld + c);
        acc.x += w[q] * v.x, acc.y += w[q] * v.y, acc.z += w[q] * v.z, acc.w += w[q] * v.w;
      }
    }
  }
  st4(cols + pix * out_ld + (size_t)k * C + c, acc);
}

// ---- lateral += bilinear(top), align_corners=False (ATen upsample_bilinear2d index rule) -----
template <typename T>
__global__ void upsample_add_kernel(const T* __restrict__ top, int Ht, int Wt, int top_ld, T* __restrict__ lat, int H,
                                    int W, int lat_ld, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int C4 = C / 4;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)H * W * C4) return;
  int c = (int)(idx % C4) * 4;
  size_t pix = idx / C4;
  int x = (int)(pix % W), y = (int)(pix / W);
  const float sh = (float)Ht / (float)H, sw = (float)Wt / (float)W;
  float fy = fmaxf(sh * ((float)y + 0.5f) - 0.5f, 0.f);
  float fx = fmaxf(sw * ((float)x + 0.5f) - 0.5f, 0.f);
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = y0 + (y0 < Ht - 1 ? 1 : 0), x1 = x0 + (x0 < Wt - 1 ? 1 : 0);
  float ly = fy - (float)y0, lx = fx - (float)x0;
  float hy = 1.f - ly, hx = 1.f - lx;
  float4 v00 = ld4(top + ((size_t)y0 * Wt + x0) * top_ld + c), v01 = ld4(top + ((size_t)y0 * Wt + x1) * top_ld + c);
  float4 v10 = ld4(top + ((size_t)y1 * Wt + x0) * top_ld + c), v11 = ld4(top + ((size_t)y1 * Wt + x1) * top_ld + c);
  T* dst = lat + pix * lat_ld + c;
  float4 l = ld4(dst);
  // ATen: w_y0 * (w_x0 * v00 + w_x1 * v01) + w_y1 * (w_x0 * v10 + w_x1 * v11)
  l.x += hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
  l.y += hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
  l.z += hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
  l.w += hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
  st4(dst, l);
}

// fp16, 8 channels per thread; same arithmetic per channel as upsample_add_kernel
__global__ void upsample_add_h8_kernel(const __half* __restrict__ top, int Ht, int Wt, int top_ld, __half* __restrict__ lat, int H,
                                       int W, int lat_ld, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int C8 = C / 8;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)H * W * C8) return;
  const int c = (int)(idx % C8) * 8;
  const size_t pix = idx / C8;
  const int x = (int)(pix % W), y = (int)(pix / W);
  const float sh = (float)Ht / (float)H, sw = (float)Wt / (float)W;
  const float fy = fmaxf(sh * ((float)y + 0.5f) - 0.5f, 0.f);
  const float fx = fmaxf(sw * ((float)x + 0.5f) - 0.5f, 0.f);
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < Ht - 1 ? 1 : 0), x1 = x0 + (x0 < Wt - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const uint4 q00 = *reinterpret_cast<const uint4*>(top + ((size_t)y0 * Wt + x0) * top_ld + c);
  const uint4 q01 = *reinterpret_cast<const uint4*>(top + ((size_t)y0 * Wt + x1) * top_ld + c);
  const uint4 q10 = *reinterpret_cast<const uint4*>(top + ((size_t)y1 * Wt + x0) * top_ld + c);
  const uint4 q11 = *reinterpret_cast<const uint4*>(top + ((size_t)y1 * Wt + x1) * top_ld + c);
  __half* dst = lat + pix * lat_ld + c;
  uint4 ql = *reinterpret_cast<const uint4*>(dst);
  const __half2 *p00 = reinterpret_cast<const __half2*>(&q00), *p01 = reinterpret_cast<const __half2*>(&q01);
  const __half2 *p10 = reinterpret_cast<const __half2*>(&q10), *p11 = reinterpret_cast<const __half2*>(&q11);
  __half2* pl = reinterpret_cast<__half2*>(&ql);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v00 = __half22float2(p00[i]), v01 = __half22float2(p01[i]), v10 = __half22float2(p10[i]), v11 = __half22float2(p11[i]);
    float2 l = __half22float2(pl[i]);
    // ATen: w_y0 * (w_x0 * v00 + w_x1 * v01) + w_y1 * (w_x0 * v10 + w_x1 * v11)
    l.x += hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
    l.y += hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
    pl[i] = __floats2half2_rn(l.x, l.y);
  }
  *reinterpret_cast<uint4*>(dst) = ql;
}

// ---- out[y][x] = in[2y][2x] ----------------------------------------------------------------
template <typename T>
__global__ void subsample2_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W, int C, int in_ld,
                                  int out_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1, C4 = C / 4;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)OH * OW * C4) return;
  int c = (int)(idx % C4) * 4;
  size_t pix = idx / C4;
  int x = (int)(pix % OW), y = (int)(pix / OW);
  st4(out + pix * out_ld + c, ld4(in + ((size_t)(2 * y) * W + 2 * x) * in_ld + c));
}

// ---- GroupNorm (+ReLU), in place; one CTA per (sample, group) --------------------------------
template <typename T>
__global__ void groupnorm_relu_kernel(T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      int HW, int C, int ld, int groups, float eps, int relu) {
  const int n = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = C / groups;
  T* base = x + (size_t)n * HW * ld + g * cpg;
  const int cnt = HW * cpg;
  __shared__ float s_sum[32], s_sq[32];
  __shared__ float s_mean, s_rstd;
  // pass 1: mean
  float sum = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) sum += to_f(base[(size_t)(i / cpg) * ld + (i % cpg)]);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s_sum[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) s_mean = v / (float)cnt;
  }
  __syncthreads();
  const float mean = s_mean;
  // pass 2: biased variance around the mean
  float sq = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    float d = to_f(base[(size_t)(i / cpg) * ld + (i % cpg)]) - mean;
    sq += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((threadIdx.x & 31) == 0) s_sq[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s_sq[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) s_rstd = rsqrtf(v / (float)cnt + eps);
  }
  __syncthreads();
  const float rstd = s_rstd;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int c = i % cpg;
    T* ptr = base + (size_t)(i / cpg) * ld + c;
    float y = (to_f(*ptr) - mean) * rstd * gamma[g * cpg + c] + beta[g * cpg + c];
    if (relu) y = fmaxf(y, 0.f);
    *ptr = from_f<T>(y);
  }
}

// Register-resident variant for 4 channels per group and HW <= 256 (the EMM towers: 16x16 maps, 32 groups of 4):
// one CTA per (sample, slab of 32 channels = 8 groups); thread = (pixel phase tid/8, group tid%8) keeps its
// <= 8 pixels x 4 channels in registers, so the tensor is read once and written once; the two-pass variance
// (mean first, then squared deviations) is kept.
template <typename T>
__global__ void __launch_bounds__(256) groupnorm4_relu_kernel(T* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, int HW, int C, int ld,
                                                             float eps, int relu) {
  constexpr int ITERS = 8;
  const int slabs = C / 32;
  const int n = blockIdx.x / slabs, c0 = (blockIdx.x % slabs) * 32;
  const int q = threadIdx.x & 7, pr = threadIdx.x >> 3, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  T* base = x + (size_t)n * HW * ld + c0 + q * 4;
  __shared__ float part[8][8];
  float4 v[ITERS];
  float sum = 0.f;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int px = pr + it * 32;
    v[it] = px < HW ? ld4(base + (size_t)px * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
    sum += (v[it].x + v[it].y) + (v[it].z + v[it].w);
  }
  const float inv = 1.f / (float)(HW * 4);
  // group total: lanes q, q+8, q+16, q+24 of every warp, then the 8 warps
  auto group_total = [&](float t) -> float {
    t += __shfl_xor_sync(0xffffffffu, t, 8);
    t += __shfl_xor_sync(0xffffffffu, t, 16);
    __syncthreads();  // previous use of part[] is over
    if (lane < 8) part[warp][lane] = t;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += part[w][q];
    return r;
  };
  const float mean = group_total(sum) * inv;
  float sq = 0.f;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    if (pr + it * 32 < HW) {
      const float dx = v[it].x - mean, dy = v[it].y - mean, dz = v[it].z - mean, dw = v[it].w - mean;
      sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float rstd = rsqrtf(group_total(sq) * inv + eps);
  const float4 gm = *reinterpret_cast<const float4*>(gamma + c0 + q * 4), bt = *reinterpret_cast<const float4*>(beta + c0 + q * 4);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int px = pr + it * 32;
    if (px < HW) {
      float4 y;
      y.x = (v[it].x - mean) * rstd * gm.x + bt.x;
      y.y = (v[it].y - mean) * rstd * gm.y + bt.y;
      y.z = (v[it].z - mean) * rstd * gm.z + bt.z;
      y.w = (v[it].w - mean) * rstd * gm.w + bt.w;
      if (relu) y.x = fmaxf(y.x, 0.f), y.y = fmaxf(y.y, 0.f), y.z = fmaxf(y.z, 0.f), y.w = fmaxf(y.w, 0.f);
      st4(base + (size_t)px * ld, y);
    }
  }
}

static inline unsigned blocks_for(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace smot

using namespace smot;

extern "C" int smot_image_to_nhwc(const float* chw, void* out, int C, int H, int W, int out_ld, int dtype, void* stream) {
  SMOT_CHECK_ARG(chw && out && C > 0 && H > 0 && W > 0 && out_ld >= C, "smot_image_to_nhwc: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int HW = H * W;
  if (dtype == SMOT_F32)
    launch_pdl(image_to_nhwc_kernel<float>, dim3(blocks_for(HW, 256)), dim3(256), 0, st, chw, (float*)out, C, HW, out_ld);
  else if (dtype == SMOT_F16)
    launch_pdl(image_to_nhwc_kernel<__half>, dim3(blocks_for(HW, 256)), dim3(256), 0, st, chw, (__half*)out, C, HW, out_ld);
  else
    SMOT_CHECK_ARG(false, "smot_image_to_nhwc: bad dtype %d", dtype);
  SMOT_CHECK_LAUNCH("smot_image_to_nhwc");
  return SMOT_OK;
}

extern "C" int smot_maxpool2x2(const void* in, void* out, int batch, int H, int W, int C, int in_ld, int out_ld,
                               int dtype, void* stream) {
  SMOT_CHECK_ARG(in && out && batch > 0 && H >= 2 && W >= 2 && C % 4 == 0 && in_ld % 4 == 0 && out_ld % 4 == 0,
                 "smot_maxpool2x2: bad arguments (C, in_ld, out_ld must be multiples of 4)");
  cudaStream_t st = (cudaStream_t)stream;
  size_t total = (size_t)batch * (H / 2) * (W / 2) * (C / 4);
  if (dtype == SMOT_F32)
    launch_pdl(maxpool2x2_kernel<float>, dim3(blocks_for(total, 256)), dim3(256), 0, st, (const float*)in, (float*)out, batch, H, W, C, in_ld, out_ld);
  else if (dtype == SMOT_F16 && C % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0)
    launch_pdl(maxpool2x2_h8_kernel, dim3(blocks_for(total / 2, 256)), dim3(256), 0, st, (const __half*)in, (__half*)out, batch, H, W, C, in_ld, out_ld);
  else if (dtype == SMOT_F16)
    launch_pdl(maxpool2x2_kernel<__half>, dim3(blocks_for(total, 256)), dim3(256), 0, st, (const __half*)in, (__half*)out, batch, H, W, C, in_ld, out_ld);
  else
    SMOT_CHECK_ARG(false, "smot_maxpool2x2: bad dtype %d", dtype);
  SMOT_CHECK_LAUNCH("smot_maxpool2x2");
  return SMOT_OK;
}

extern "C" int smot_maxpool3x3s2(const void* in, void* out, int batch, int H, int W, int C, int in_ld, int out_ld,
                                 int dtype, void* stream) {
  SMOT_CHECK_ARG(in && out && batch > 0 && H >= 1 && W >= 1 && C > 0 && C % 4 == 0 && in_ld % 4 == 0 && out_ld % 4 == 0,
                 "smot_maxpool3x3s2: bad arguments (C, in_ld, out_ld must be multiples of 4)");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)batch * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1) * (C / 4);
  if (dtype == SMOT_F32)
    launch_pdl(maxpool3x3s2_kernel<float>, dim3(blocks_for(total, 256)), dim3(256), 0, st, (const float*)in, (float*)out, batch, H, W, C, in_ld, out_ld);
  else if (dtype == SMOT_F16)
    launch_pdl(maxpool3x3s2_kernel<__half>, dim3(blocks_for(total, 256)), dim3(256), 0, st, (const __half*)in, (__half*)out, batch, H, W, C, in_ld, out_ld);
  else
    SMOT_CHECK_ARG(false, "smot_maxpool3x3s2: bad dtype %d", dtype);
  SMOT_CHECK_LAUNCH("smot_maxpool3x3s2");
  return SMOT_OK;
}

extern "C" int smot_deform_im2col3x3(const void* in, const float* offsets, void* cols, int H, int W, int C, int in_ld, int off_ld,
                                     int OH, int OW, int out_ld, int stride, int dtype, void* stream) {
  SMOT_CHECK_ARG(in && offsets && cols && H > 0 && W > 0 && C > 0 && C % 4 == 0 && in_ld % 4 == 0 && out_ld % 4 == 0 && off_ld >= 18,
                 "smot_deform_im2col3x3: bad arguments (C, in_ld, out_ld multiples of 4; off_ld >= 18)");
  SMOT_CHECK_ARG((stride == 1 || stride == 2) && OH == (H - 1) / stride + 1 && OW == (W - 1) / stride + 1 && out_ld >= 9 * C,
                 "smot_deform_im2col3x3: geometry (3x3, pad 1, stride %d: %dx%d -> %dx%d, out_ld %d)", stride, H, W, OH, OW, out_ld);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)OH * OW * 9 * (C / 4);
  if (dtype == SMOT_F32)
    launch_pdl(deform_im2col3x3_kernel<float>, dim3(blocks_for(total, 256)), dim3(256), 0, st, (const float*)in, offsets, (float*)cols, H, W, C, in_ld, off_ld, OH, OW, out_ld, stride);
  else if (dtype == SMOT_F16)
    launch_pdl(deform_im2col3x3_kernel<__half>, dim3(blocks_for(total, 256)), dim3(256), 0, st, (const __half*)in, offsets, (__half*)cols, H, W, C, in_ld, off_ld, OH, OW, out_ld, stride);
  else
    SMOT_CHECK_ARG(false, "smot_deform_im2col3x3: bad dtype %d", dtype);
  SMOT_CHECK_LAUNCH("smot_deform_im2col3x3");
  return SMOT_OK;
}

extern "C" int smot_upsample_add(const void* top, int Ht, int Wt, int top_ld, void* lateral, int H, int W, int lat_ld,
                                 int C, int dtype, void* stream) {
  SMOT_CHECK_ARG(top && lateral && Ht > 0 && Wt > 0 && H > 0 && W > 0 && C % 4 == 0 && top_ld % 4 == 0 && lat_ld % 4 == 0,
                 "smot_upsample_add: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  size_t total = (size_t)H * W * (C / 4);
  if (dtype == SMOT_F32)
    launch_pdl(upsample_add_kernel<float>, dim3(blocks_for(total, 256)), dim3(256), 0, st, (const float*)top, Ht, Wt, top_ld, (float*)lateral, H, W, lat_ld, C);
  else if (dtype == SMOT_F16 && C % 8 == 0 && top_ld % 8 == 0 && lat_ld % 8 == 0 && (((uintptr_t)top | (uintptr_t)lateral) & 15) == 0)
    launch_pdl(upsample_add_h8_kernel, dim3(blocks_for(total / 2, 256)), dim3(256), 0, st, (const __half*)top, Ht, Wt, top_ld, (__half*)lateral, H, W, lat_ld, C);
  else if (dtype == SMOT_F16)
    launch_pdl(upsample_add_kernel<__half>, dim3(blocks_for(total, 256)), dim3(256), 0, st, (const __half*)top, Ht, Wt, top_ld, (__half*)lateral, H, W, lat_ld, C);
  else
    SMOT_CHECK_ARG(false, "smot_upsample_add: bad dtype %d", dtype);
  SMOT_CHECK_LAUNCH("smot_upsample_add");
  return SMOT_OK;
}

extern "C" int smot_subsample2(const void* in, void* out, int H, int W, int C, int in_ld, int out_ld, int dtype,
                               void* stream) {
  SMOT_CHECK_ARG(in && out && H > 0 && W > 0 && C % 4 == 0 && in_ld % 4 == 0 && out_ld % 4 == 0, "smot_subsample2: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  size_t total = (size_t)((H - 1) / 2 + 1) * ((W - 1) / 2 + 1) * (C / 4);
  if (dtype == SMOT_F32)
    launch_pdl(subsample2_kernel<float>, dim3(blocks_for(total, 256)), dim3(256), 0, st, (const float*)in, (float*)out, H, W, C, in_ld, out_ld);
  else if (dtype == SMOT_F16)
    launch_pdl(subsample2_kernel<__half>, dim3(blocks_for(total, 256)), dim3(256), 0, st, (const __half*)in, (__half*)out, H, W, C, in_ld, out_ld);
  else
    SMOT_CHECK_ARG(false, "smot_subsample2: bad dtype %d", dtype);
  SMOT_CHECK_LAUNCH("smot_subsample2");
  return SMOT_OK;
}

extern "C" int smot_groupnorm_relu(void* x, const float* gamma, const float* beta, int batch, int HW, int C, int ld,
                                   int groups, float eps, int relu, int dtype, void* stream) {
  SMOT_CHECK_ARG(x && gamma && beta && batch >= 0 && HW > 0 && groups > 0 && C % groups == 0 && ld >= C,
                 "smot_groupnorm_relu: bad arguments");
  if (batch == 0) return SMOT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = dtype == SMOT_F16 ? 2 : 4;
  const bool fast = C == groups * 4 && C % 32 == 0 && HW <= 256 && ((uintptr_t)x % (4 * esz)) == 0 && ld % 4 == 0 &&
                    ((uintptr_t)gamma & 15) == 0 && ((uintptr_t)beta & 15) == 0;
  if (fast && dtype == SMOT_F32)
    groupnorm4_relu_kernel<float><<<batch * (C / 32), 256, 0, st>>>((float*)x, gamma, beta, HW, C, ld, eps, relu);
  else if (fast && dtype == SMOT_F16)
    groupnorm4_relu_kernel<__half><<<batch * (C / 32), 256, 0, st>>>((__half*)x, gamma, beta, HW, C, ld, eps, relu);
  else if (dtype == SMOT_F32)
    groupnorm_relu_kernel<float><<<batch * groups, 256, 0, st>>>((float*)x, gamma, beta, HW, C, ld, groups, eps, relu);
  else if (dtype == SMOT_F16)
    groupnorm_relu_kernel<__half><<<batch * groups, 256, 0, st>>>((__half*)x, gamma, beta, HW, C, ld, groups, eps, relu);
  else
    SMOT_CHECK_ARG(false, "smot_groupnorm_relu: bad dtype %d", dtype);
  SMOT_CHECK_LAUNCH("smot_groupnorm_relu");
  return SMOT_OK;
}

// Convolution / fully-connected layers with very few output channels (Cout <= 16): the RPN
// objectness+delta predictor (15), the EMM cls/center/reg heads (3 / 4) and the box predictor (10).
// They are bandwidth/latency bound (the weight is tiny, every input element is used Cout times), so a
// GEMM tile is the wrong shape: here ONE WARP produces PIX consecutive output pixels; its lanes split
// the input channels (4 per lane, contiguous 8/16-byte loads, 256/512 B per warp per pixel and tap),
// the weights sit in shared memory as [Cout][K] and are read conflict-free as float4, and the partial
// sums are combined with a butterfly reduction.  Epilogue: bias / scale, optional ReLU, fp32 or fp16 out.
#include "common.cuh"

namespace smot {

struct SmallNArgs {
  const void* in;
  const void* wt;
  const float* scale;
  const float* bias;
  void* out;
  int batch, H, W, Cin, in_ld, OH, OW, Cout, out_ld, KH, KW, pad, relu, M, K;
};

constexpr int SN_PIX = 4;
constexpr int SN_WARPS = 8;

template <typename TI, typename TO, int COUT>
__global__ void __launch_bounds__(SN_WARPS * 32) conv_smalln_kernel(const SmallNArgs a) {
  extern __shared__ __align__(16) float sn_w[];  // [COUT][K] (rows >= a.Cout are zero)
  const TI* __restrict__ wt = reinterpret_cast<const TI*>(a.wt);
  for (int i = threadIdx.x; i < COUT * a.K; i += blockDim.x) {
    const int co = i / a.K;
    sn_w[i] = co < a.Cout ? to_f(wt[(size_t)co * a.K + (i - co * a.K)]) : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m0 = (blockIdx.x * SN_WARPS + warp) * SN_PIX;
  if (m0 >= a.M) return;
  const TI* __restrict__ in = reinterpret_cast<const TI*>(a.in);
  const TI* base[SN_PIX];
  int ih0[SN_PIX], iw0[SN_PIX];
  bool ok[SN_PIX];
#pragma unroll
  for (int p = 0; p < SN_PIX; ++p) {
    const int m = m0 + p;
    ok[p] = m < a.M;
    const int mm = ok[p] ? m : m0;
    const int img = mm / (a.OH * a.OW);
    const int rem = mm - img * (a.OH * a.OW);
    const int oh = rem / a.OW;
    ih0[p] = oh - a.pad;
    iw0[p] = rem - oh * a.OW - a.pad;
    base[p] = in + (size_t)img * a.H * a.W * a.in_ld;
  }
  float acc[SN_PIX][COUT];
#pragma unroll
  for (int p = 0; p < SN_PIX; ++p)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[p][c] = 0.f;

  for (int r = 0; r < a.KH; ++r)
    for (int s = 0; s < a.KW; ++s) {
      const int kbase = (r * a.KW + s) * a.Cin;
      for (int c0 = lane * 4; c0 < a.Cin; c0 += 128) {
        float4 x[SN_PIX];
#pragma unroll
        for (int p = 0; p < SN_PIX; ++p) {
          const int ih = ih0[p] + r, iw = iw0[p] + s;
          const bool v = ok[p] && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W;
          x[p] = v ? ld4(base[p] + ((size_t)ih * a.W + iw) * a.in_ld + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          const float4 w = *reinterpret_cast<const float4*>(sn_w + (size_t)c * a.K + kbase + c0);
#pragma unroll
          for (int p = 0; p < SN_PIX; ++p) {
            acc[p][c] = fmaf(x[p].x, w.x, acc[p][c]);
            acc[p][c] = fmaf(x[p].y, w.y, acc[p][c]);
            acc[p][c] = fmaf(x[p].z, w.z, acc[p][c]);
            acc[p][c] = fmaf(x[p].w, w.w, acc[p][c]);
          }
        }
      }
    }
  // butterfly reduction: afterwards every lane holds the full sums
#pragma unroll
  for (int p = 0; p < SN_PIX; ++p)
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      float v = acc[p][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[p][c] = v;
    }
  TO* __restrict__ out = reinterpret_cast<TO*>(a.out);
#pragma unroll
  for (int p = 0; p < SN_PIX; ++p) {
    if (!ok[p]) continue;
    // lane c writes channel c (select without dynamic register indexing)
    float y = 0.f;
#pragma unroll
    for (int c = 0; c < COUT; ++c)
      if (lane == c) y = acc[p][c];
    if (lane < a.Cout) {
      if (a.scale) y = __fmul_rn(y, a.scale[lane]);
      if (a.bias) y = __fadd_rn(y, a.bias[lane]);
      if (a.relu) y = fmaxf(y, 0.f);
      out[(size_t)(m0 + p) * a.out_ld + lane] = from_f<TO>(y);
    }
  }
}

// Few output pixels (box predictor on <= 300 ROIs, refinement on the tracks): staging the whole weight in
// shared memory per CTA would dominate, so each warp streams the weight rows it needs straight from L2
// (one warp per output pixel, lanes split the input channels, Cout accumulators per lane).
template <typename TI, typename TO, int COUT>
__global__ void __launch_bounds__(128) conv_smalln_direct_kernel(const SmallNArgs a) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (m >= a.M) return;
  const TI* __restrict__ in = reinterpret_cast<const TI*>(a.in);
  const TI* __restrict__ wt = reinterpret_cast<const TI*>(a.wt);
  const int img = m / (a.OH * a.OW);
  const int rem = m - img * (a.OH * a.OW);
  const int oh = rem / a.OW, ow = rem - oh * a.OW;
  const TI* base = in + (size_t)img * a.H * a.W * a.in_ld;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
  for (int r = 0; r < a.KH; ++r)
    for (int s = 0; s < a.KW; ++s) {
      const int ih = oh - a.pad + r, iw = ow - a.pad + s;
      if (ih < 0 || ih >= a.H || iw < 0 || iw >= a.W) continue;
      const TI* px = base + ((size_t)ih * a.W + iw) * a.in_ld;
      const int kbase = (r * a.KW + s) * a.Cin;
      for (int c0 = lane * 4; c0 < a.Cin; c0 += 128) {
        const float4 x = ld4(px + c0);
        float4 w[COUT];  // unconditional loads (rows past Cout re-read the last row) so all are in flight together
#pragma unroll
        for (int c = 0; c < COUT; ++c) w[c] = ld4(wt + (size_t)min(c, a.Cout - 1) * a.K + kbase + c0);
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          acc[c] = fmaf(x.x, w[c].x, acc[c]);
          acc[c] = fmaf(x.y, w[c].y, acc[c]);
          acc[c] = fmaf(x.z, w[c].z, acc[c]);
          acc[c] = fmaf(x.w, w[c].w, acc[c]);
        }
      }
    }
#pragma unroll
  for (int c = 0; c < COUT; ++c) {
    float v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    acc[c] = v;
  }
  float y = 0.f;
#pragma unroll
  for (int c = 0; c < COUT; ++c)
    if (lane == c) y = acc[c];
  if (lane < a.Cout) {
    if (a.scale) y = __fmul_rn(y, a.scale[lane]);
    if (a.bias) y = __fadd_rn(y, a.bias[lane]);
    if (a.relu) y = fmaxf(y, 0.f);
    reinterpret_cast<TO*>(a.out)[(size_t)m * a.out_ld + lane] = from_f<TO>(y);
  }
}

// ---------------------------------------------------------------------------------------------
// fp16 storage: the same layers on the tensor cores (mma.sync.m16n8k16, fp32 accumulation), WITHOUT shared
// memory or barriers.  A warp owns 16 output pixels x 8*NT output channels.  The MMA contraction index is
// permuted so that every fragment register pair is part of ONE 16-byte global load: within a 32-channel chunk
// thread (g, t) loads channels [8t, 8t+8) of pixel g (and g+8) and of weight row n = g (and g+8), and uses
// halves {0,1 | 2,3} as the (k = 2t.. | k = 2t+8..) slots of a first k-step and halves {4,5 | 6,7} of a second:
// A and B agree on the permutation, so the sum is unchanged.  Zero padding = predicated loads.
// The 8 warps of a CTA are WM pixel groups x WK slices of the K loop (partial sums meet in shared memory), which
// keeps >= ~1000 warps busy from the 56k-pixel RPN map down to the 30-row refinement FC.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sn_mma(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <typename TO, int NT>
__global__ void __launch_bounds__(256) conv_smalln_mma_kernel(const SmallNArgs a, int WM, int WK) {
  __shared__ float red[8][32][NT * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int wm = warp % WM, wk = warp / WM;
  const int m0 = (blockIdx.x * WM + wm) * 16;
  pdl_launch_dependents();
  pdl_wait();
  const __half* __restrict__ in = reinterpret_cast<const __half*>(a.in);
  const __half* __restrict__ wt = reinterpret_cast<const __half*>(a.wt);
  // the two pixel rows of this thread's A fragments
  const __half* base[2];
  int ih0[2], iw0[2];
  bool ok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int m = m0 + g + 8 * h;
    ok[h] = m < a.M;
    const int mm = ok[h] ? m : 0;
    const int img = mm / (a.OH * a.OW), rem = mm - img * (a.OH * a.OW), oh = rem / a.OW;
    ih0[h] = oh - a.pad, iw0[h] = rem - oh * a.OW - a.pad;
    base[h] = in + (size_t)img * a.H * a.W * a.in_ld + t * 8;
  }
  const __half* wrow[NT];
  bool wok[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    wok[nt] = nt * 8 + g < a.Cout;
    wrow[nt] = wt + (size_t)(wok[nt] ? nt * 8 + g : 0) * a.K + t * 8;
  }
  float acc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
  const int cpc = a.Cin >> 5, chunks = a.KH * a.KW * cpc;   // 32-channel chunks per tap / in total
  const int per = (chunks + WK - 1) / WK;
  const int q0 = wk * per, q1 = min(chunks, q0 + per);
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  constexpr int UN = 4;  // chunks whose loads are in flight together
  for (int qb = q0; qb < q1; qb += UN) {
    uint4 av[UN][2], bv[UN][NT];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int q = qb + u;
      const bool live = q < q1;
      const int tap = live ? q / cpc : 0, cc = live ? q - tap * cpc : 0;
      const int r = tap / a.KW, sx = tap - r * a.KW;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ih = ih0[h] + r, iw = iw0[h] + sx;
        const bool v = live && ok[h] && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W;
        av[u][h] = v ? *reinterpret_cast<const uint4*>(base[h] + ((size_t)ih * a.W + iw) * a.in_ld + cc * 32) : zero;
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        bv[u][nt] = (live && wok[nt]) ? *reinterpret_cast<const uint4*>(wrow[nt] + (size_t)tap * a.Cin + cc * 32) : zero;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        sn_mma(acc[nt], av[u][0].x, av[u][1].x, av[u][0].y, av[u][1].y, bv[u][nt].x, bv[u][nt].y);
        sn_mma(acc[nt], av[u][0].z, av[u][1].z, av[u][0].w, av[u][1].w, bv[u][nt].z, bv[u][nt].w);
      }
  }
  if (WK > 1) {  // partial sums of the K slices meet in shared memory; slice 0 finishes the tile
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) red[warp][lane][nt * 4 + e] = acc[nt][e];
    __syncthreads();
    if (wk != 0) return;
    for (int k = 1; k < WK; ++k)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] += red[k * WM + wm][lane][nt * 4 + e];
  }
  TO* __restrict__ out = reinterpret_cast<TO*>(a.out);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = nt * 8 + 2 * t + (e & 1), h = e >> 1;   // D fragment: rows g / g+8, columns 2t / 2t+1
      if (ok[h] && n < a.Cout) {
        float y = acc[nt][e];
        if (a.scale) y = __fmul_rn(y, a.scale[n]);
        if (a.bias) y = __fadd_rn(y, a.bias[n]);
        if (a.relu) y = fmaxf(y, 0.f);
        out[(size_t)(m0 + g + 8 * h) * a.out_ld + n] = from_f<TO>(y);
      }
    }
}

static bool smalln_mma_ok(const smot_conv_desc* d) {
  return d->in_dtype == SMOT_F16 && d->Cin % 32 == 0 && d->in_ld % 8 == 0 && (((uintptr_t)d->in | (uintptr_t)d->weight) & 15) == 0 &&
         ((size_t)d->KH * d->KW * d->Cin) % 8 == 0;
}

template <typename TO>
static int launch_smalln_mma(const SmallNArgs& a, cudaStream_t st) {
  const int mtiles = ceil_div(a.M, 16), chunks = a.KH * a.KW * (a.Cin / 32);
  // the K slicing (= the summation order) is chosen from the pixels of ONE image, so a batched pass rounds exactly like
  // per-image passes (Engine.pair_plan: clip results must equal frame-by-frame results bit for bit)
  const int mtiles_img = ceil_div(a.M / (a.batch > 0 ? a.batch : 1), 16);
  int WK = 1;
  while (WK < 8 && (long long)mtiles_img * WK < 1184 && chunks / (WK * 2) >= 2) WK *= 2;   // ~8 warps per SM, >= 2 chunks per slice
  const int WM = 8 / WK;
  const unsigned grid = (unsigned)ceil_div(mtiles, WM);
  if (a.Cout <= 8)
    launch_pdl(conv_smalln_mma_kernel<TO, 1>, dim3(grid), dim3(256), 0, st, a, WM, WK);
  else
    launch_pdl(conv_smalln_mma_kernel<TO, 2>, dim3(grid), dim3(256), 0, st, a, WM, WK);
  SMOT_CHECK_LAUNCH("smot_conv2d(smalln mma)");
  return SMOT_OK;
}

bool conv2d_smalln_supported(const smot_conv_desc* d) {
  if (d->Cout > 16 || d->stride != 1 || d->residual) return false;
  if (d->Cin < 64 || d->Cin % 4 != 0 || d->in_ld % 4 != 0 || ((uintptr_t)d->in & 15)) return false;  // lanes split channels
  if ((size_t)d->KH * d->KW * d->Cin * 16 * sizeof(float) > 96 * 1024) return false;
  if (d->OH != d->H + 2 * d->pad - d->KH + 1 || d->OW != d->W + 2 * d->pad - d->KW + 1) return false;
  return true;
}

template <typename TI, typename TO>
static int launch_smalln(const SmallNArgs& a, cudaStream_t st) {
  const int cout_pad = a.Cout <= 4 ? 4 : (a.Cout <= 8 ? 8 : 16);
  if (a.M / (a.batch > 0 ? a.batch : 1) <= 1024 && ((uintptr_t)a.wt & 15) == 0 && a.K % 4 == 0) {   // per image (see launch_smalln_mma)
    const unsigned g = (unsigned)ceil_div(a.M, 4);
    if (cout_pad == 4)
      conv_smalln_direct_kernel<TI, TO, 4><<<g, 128, 0, st>>>(a);
    else if (cout_pad == 8)
      conv_smalln_direct_kernel<TI, TO, 8><<<g, 128, 0, st>>>(a);
    else
      conv_smalln_direct_kernel<TI, TO, 16><<<g, 128, 0, st>>>(a);
    SMOT_CHECK_LAUNCH("smot_conv2d(smalln direct)");
    return SMOT_OK;
  }
  const size_t smem = (size_t)cout_pad * a.K * sizeof(float);
  const unsigned grid = (unsigned)ceil_div(a.M, SN_PIX * SN_WARPS);
#define SMOT_SN_LAUNCH(CO)                                                                       \
  do {                                                                                           \
    SMOT_ENSURE_SMEM((conv_smalln_kernel<TI, TO, CO>), smem, "smot_conv2d(smalln)");             \
    conv_smalln_kernel<TI, TO, CO><<<grid, SN_WARPS * 32, smem, st>>>(a);                        \
  } while (0)
  if (cout_pad == 4)
    SMOT_SN_LAUNCH(4);
  else if (cout_pad == 8)
    SMOT_SN_LAUNCH(8);
  else
    SMOT_SN_LAUNCH(16);
#undef SMOT_SN_LAUNCH
  SMOT_CHECK_LAUNCH("smot_conv2d(smalln)");
  return SMOT_OK;
}

int conv2d_smalln(const smot_conv_desc* d, cudaStream_t st) {
  SmallNArgs a;
  a.in = d->in, a.wt = d->weight, a.scale = d->scale, a.bias = d->bias, a.out = d->out;
  a.batch = d->batch, a.H = d->H, a.W = d->W, a.Cin = d->Cin, a.in_ld = d->in_ld, a.OH = d->OH, a.OW = d->OW;
  a.Cout = d->Cout, a.out_ld = d->out_ld, a.KH = d->KH, a.KW = d->KW, a.pad = d->pad, a.relu = d->relu;
  a.M = d->batch * d->OH * d->OW;
  a.K = d->KH * d->KW * d->Cin;
  if (a.M == 0) return SMOT_OK;
  if (smalln_mma_ok(d)) return d->out_dtype == SMOT_F16 ? launch_smalln_mma<__half>(a, st) : launch_smalln_mma<float>(a, st);
  if (d->in_dtype == SMOT_F32 && d->out_dtype == SMOT_F32) return launch_smalln<float, float>(a, st);
  if (d->in_dtype == SMOT_F16 && d->out_dtype == SMOT_F16) return launch_smalln<__half, __half>(a, st);
  if (d->in_dtype == SMOT_F16 && d->out_dtype == SMOT_F32) return launch_smalln<__half, float>(a, st);
  set_error("smot_conv2d(smalln): unsupported dtype combination");
  return SMOT_ERR_UNSUPPORTED;
}

}  // namespace smot

// Device-resident proposal selection, sorting and NMS.
//
// The reference does these steps with ~10 tiny ATen kernels plus a device->host copy of the NMS mask
// and a serial host loop per call (upstream csrc/cuda/nms.cu), seven times per frame.  Here each is
// one kernel with on-device reduction, so the whole detection stage stays capturable in a CUDA graph.
//
//   sort_nms_kernel : 64-bit composite keys (score key << 32 | ~index) -> bitonic sort in shared
//                     memory -> 64-wide bitmask IoU(+1) matrix -> chunked warp reduction.
//   rpn_topk_kernel : per FPN level, radix select of the top-k objectness logits, sort, anchor
//                     synthesis + BoxCoder decode + clip  (rpn_patch.py:15-52).
//   box_decode_kernel: softmax + per-class decode + clip + track-row rule (inference.py:58-110).
#include "common.cuh"

namespace smot {

constexpr int SN_THREADS = 1024;
constexpr int SN_MAX = 4096;
constexpr float BBOX_XFORM_CLIP = 4.135166556742356f;  // log(1000/16)

struct SortNmsArgs {
  const float* boxes;
  int box_stride;
  const float* scores;
  int score_stride;
  const int* count;
  int n_max, np;  // np = power of two >= n_max
  float min_score, thresh;
  int max_keep, tag, append, fill_tail;
  int* out_index;
  float* out_boxes;
  float* out_scores;
  int* out_tag;
  int* out_count;
  unsigned long long* mask;  // [n_max][ceil(n_max/64)]
  // batching (blockIdx.x = problem): element offsets added per problem
  int in_step, out_step, mask_step;
};

__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int np) {
  for (int k = 2; k <= np; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = keys[i], b = keys[ixj];
          bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(SN_THREADS) sort_nms_kernel(SortNmsArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  float4* sbox = reinterpret_cast<float4*>(smem_raw + (size_t)a.np * 8);
  int* kept_all = reinterpret_cast<int*>(smem_raw + (size_t)a.np * 24);  // sorted row of the k-th survivor
  __shared__ unsigned long long removed[SN_MAX / 64];
  __shared__ unsigned long long diag[64];
  __shared__ int kept_rows[64];
  __shared__ int s_m, s_kept, s_nk, s_done;

  const int prob = blockIdx.x;
  const float* boxes = a.boxes + (size_t)prob * a.in_step * a.box_stride;
  const float* scores = a.scores + (size_t)prob * a.in_step * a.score_stride;
  unsigned long long* mask = a.mask + (size_t)prob * a.mask_step;
  int* out_count = a.out_count + prob;
  const int n = a.count ? min(a.count[prob], a.n_max) : a.n_max;

  if (threadIdx.x == 0) s_m = 0, s_kept = 0, s_done = 0;
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < a.np; i += blockDim.x) {
    unsigned long long key = 0ull;
    if (i < n) {
      float s = scores[(size_t)i * a.score_stride];
      if (s > a.min_score) {
        key = ((unsigned long long)float_key(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
        ++local;
      }
    }
    keys[i] = key;
  }
  if (local) atomicAdd(&s_m, local);
  __syncthreads();
  bitonic_sort_desc(keys, a.np);
  const int m = s_m;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const unsigned idx = 0xFFFFFFFFu - (unsigned)(keys[i] & 0xFFFFFFFFull);
    const float* b = boxes + (size_t)idx * a.box_stride;
    sbox[i] = make_float4(b[0], b[1], b[2], b[3]);
  }
  for (int i = threadIdx.x; i < SN_MAX / 64; i += blockDim.x) removed[i] = 0ull;
  __syncthreads();

  const int words = (m + 63) >> 6;
  int kept_total;
  if (a.thresh <= 0.f || a.max_keep <= 0) {
    kept_total = max(min(m, a.max_keep), 0);
    for (int i = threadIdx.x; i < kept_total; i += blockDim.x) kept_all[i] = i;
  } else {
    // ---- suppression bitmask: mask[i][w] bit b set <=> j = 64w+b > i and IoU(i,j) > thresh
    for (int item = threadIdx.x; item < m * words; item += blockDim.x) {
      const int i = item / words, w = item - i * words;
      unsigned long long bits = 0ull;
      if (w >= (i >> 6)) {
        const float4 bi = sbox[i];
        const int j0 = w << 6, j1 = min(m, j0 + 64);
        for (int j = max(j0, i + 1); j < j1; ++j) {
          const float4 bj = sbox[j];
          // disjoint boxes have IoU 0: skip the division (same result, most pairs are disjoint)
          if (fminf(bi.z, bj.z) - fmaxf(bi.x, bj.x) + 1.f > 0.f && fminf(bi.w, bj.w) - fmaxf(bi.y, bj.y) + 1.f > 0.f)
            if (iou_plus1(bi, bj) > a.thresh) bits |= 1ull << (j - j0);
        }
      }
      mask[(size_t)i * words + w] = bits;
    }
    __syncthreads();
    // ---- reduction, 64 sorted rows per round: (A) stage the diagonal words, (B) one thread resolves the
    //      round serially from registers, (C) the whole CTA ORs the survivors' rows into `removed`
    for (int c = 0; c < words; ++c) {
      if (threadIdx.x < 64) {
        const int row = (c << 6) + threadIdx.x;
        diag[threadIdx.x] = row < m ? mask[(size_t)row * words + c] : 0ull;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned long long cur = removed[c];
        int kept = s_kept, nk = 0;
        const int rows_here = min(64, m - (c << 6));
        for (int b = 0; b < rows_here && kept + nk < a.max_keep; ++b) {
          if (!((cur >> b) & 1ull)) {
            kept_rows[nk] = (c << 6) + b;
            kept_all[kept + nk] = (c << 6) + b;
            ++nk;
            cur |= diag[b];
          }
        }
        s_nk = nk;
        s_kept = kept + nk;
        if (kept + nk >= a.max_keep) s_done = 1;
      }
      __syncthreads();
      if (s_done) break;
      const int nk = s_nk, nw = words - c - 1;
      for (int item = threadIdx.x; item < nk * nw; item += blockDim.x) {
        const int qq = item / nw, w = c + 1 + (item - qq * nw);
        const unsigned long long v = mask[(size_t)kept_rows[qq] * words + w];
        if (v) atomicOr(&removed[w], v);
      }
      __syncthreads();
    }
    kept_total = s_kept;
  }
  __syncthreads();
  // ---- outputs, all threads
  const int base = a.append ? *out_count : 0;
  const int out_off = prob * a.out_step + base;
  for (int k = threadIdx.x; k < kept_total; k += blockDim.x) {
    const int row = kept_all[k];
    const unsigned idx = 0xFFFFFFFFu - (unsigned)(keys[row] & 0xFFFFFFFFull);
    if (a.out_index) a.out_index[out_off + k] = (int)idx;
    if (a.out_boxes) reinterpret_cast<float4*>(a.out_boxes)[out_off + k] = sbox[row];
    if (a.out_scores) a.out_scores[out_off + k] = scores[(size_t)idx * a.score_stride];
    if (a.out_tag) a.out_tag[out_off + k] = a.tag;
  }
  // rows [kept, fill_tail) of this problem's output slot are marked invalid (score -1)
  if (a.fill_tail > 0 && a.out_scores)
    for (int i = kept_total + threadIdx.x; i < a.fill_tail - base; i += blockDim.x) a.out_scores[out_off + i] = -1.f;
  __syncthreads();
  if (threadIdx.x == 0) *out_count = base + kept_total;
}

static int next_pow2(int n) {
  int p = 64;
  while (p < n) p <<= 1;
  return p;
}

static int launch_sort_nms(SortNmsArgs& a, int problems, cudaStream_t st) {
  a.np = next_pow2(a.n_max);
  const size_t smem = (size_t)a.np * 28;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sort_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SN_MAX * 28);
    if (e != cudaSuccess) {
      set_error("sort_nms: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return SMOT_ERR_CUDA;
    }
    attr_set = true;
  }
  sort_nms_kernel<<<problems, SN_THREADS, smem, st>>>(a);
  SMOT_CHECK_LAUNCH("sort_nms");
  return SMOT_OK;
}

// ---------------------------------------------------------------------------------------------
// RPN: per-level top-k by objectness logit + decode
// ---------------------------------------------------------------------------------------------
struct RpnArgs {
  smot_rpn_level lv[SMOT_MAX_LEVELS];
  int pre_nms_top_n;  // <= 1024
  float min_size;
  int img_w, img_h, amodal;
  float* cand_boxes;   // [levels][pre_nms_top_n][4]
  float* cand_scores;  // [levels][pre_nms_top_n]
  int* cand_count;     // [levels]
};

__global__ void __launch_bounds__(1024) rpn_topk_kernel(const RpnArgs a) {
  const smot_rpn_level& L = a.lv[blockIdx.x];
  const int n = L.H * L.W * L.A;
  const int k = min(a.pre_nms_top_n, n);
  __shared__ unsigned long long cand[1024];
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_need, s_cnt, s_eqbase, s_numeq;
  __shared__ unsigned warp_tot[32];
  const float* __restrict__ head = L.head;
  auto logit_at = [&](int i) -> float { return head[(size_t)(i / L.A) * L.head_ld + (i % L.A)]; };

  // ---- radix select (MSB first, 8 bits per pass) of the k-th largest key
  if (threadIdx.x == 0) s_prefix = 0u, s_need = (unsigned)k;
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned pmask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i0 = threadIdx.x; i0 < n; i0 += 8 * blockDim.x) {  // 8 independent loads in flight per thread
      unsigned key[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        key[u] = i < n ? float_key(logit_at(i)) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * (int)blockDim.x < n && (key[u] & pmask) == prefix) atomicAdd(&hist[(key[u] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned need = s_need, cum = 0u;
      int b = 255;
      for (; b > 0; --b) {
        if (cum + hist[b] >= need) break;
        cum += hist[b];
      }
      s_need = need - cum;  // still needed among keys sharing the extended prefix
      s_prefix = prefix | ((unsigned)b << shift);
      s_numeq = hist[b];    // after the last pass: how many keys equal the k-th key
    }
    __syncthreads();
  }
  const unsigned T = s_prefix;      // k-th largest key
  const unsigned need_eq = s_need;  // how many keys == T belong to the top-k
  const bool all_eq = s_numeq == need_eq;  // no surplus ties: every key == T is taken, order irrelevant
  // ---- collect: keys > T in any order, keys == T lowest index first
  if (threadIdx.x == 0) s_cnt = 0u, s_eqbase = 0u;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) cand[i] = 0ull;
  __syncthreads();
  for (int i0 = threadIdx.x; i0 < n; i0 += 8 * blockDim.x) {
    unsigned key[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      key[u] = i < n ? float_key(logit_at(i)) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < n && (key[u] > T || (all_eq && key[u] == T))) {
        unsigned pos = atomicAdd(&s_cnt, 1u);
        cand[pos] = ((unsigned long long)key[u] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
      }
    }
  }
  __syncthreads();
  const unsigned n_gt = s_cnt;
  for (int i0 = 0; i0 < n && !all_eq; i0 += blockDim.x) {  // ordered pass over surplus ties (block scan per chunk)
    const int i = i0 + threadIdx.x;
    const bool eq = i < n && float_key(logit_at(i)) == T;
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    unsigned before = s_eqbase;
    for (int w = 0; w < wid; ++w) before += warp_tot[w];
    const unsigned pos = before + __popc(bal & ((1u << lane) - 1u));
    if (eq && pos < need_eq)
      cand[n_gt + pos] = ((unsigned long long)T << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned tot = 0u;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_tot[w];
      s_eqbase += tot;
    }
    __syncthreads();
    if (s_eqbase >= need_eq) break;
  }
  __syncthreads();
  bitonic_sort_desc(cand, 1024);
  // ---- decode the k candidates in sorted order
  float* cb = a.cand_boxes + (size_t)blockIdx.x * a.pre_nms_top_n * 4;
  float* cs = a.cand_scores + (size_t)blockIdx.x * a.pre_nms_top_n;
  for (int j = threadIdx.x; j < a.pre_nms_top_n; j += blockDim.x) {
    if (j >= k) {
      cs[j] = -1.f;
      continue;
    }
    const int i = (int)(0xFFFFFFFFu - (unsigned)(cand[j] & 0xFFFFFFFFull));
    const int cell = i / L.A, an = i - cell * L.A;
    const int y = cell / L.W, x = cell - y * L.W;
    const float* row = head + (size_t)cell * L.head_ld;
    const float logit = row[an];
    const float score = __fdiv_rn(1.f, 1.f + expf(-logit));
    const float sx = (float)(x * L.stride), sy = (float)(y * L.stride);
    const float ax1 = L.cell_anchors[an * 4 + 0] + sx, ay1 = L.cell_anchors[an * 4 + 1] + sy;
    const float ax2 = L.cell_anchors[an * 4 + 2] + sx, ay2 = L.cell_anchors[an * 4 + 3] + sy;
    const float* d = row + L.A + 4 * an;
    // BoxCoder(1,1,1,1).decode (TO_REMOVE = 1)
    const float w = ax2 - ax1 + 1.f, h = ay2 - ay1 + 1.f;
    const float cx = ax1 + 0.5f * w, cy = ay1 + 0.5f * h;
    const float dw = fminf(d[2], BBOX_XFORM_CLIP), dh = fminf(d[3], BBOX_XFORM_CLIP);
    const float pcx = d[0] * w + cx, pcy = d[1] * h + cy;
    const float pw = expf(dw) * w, phh = expf(dh) * h;
    float x1 = pcx - 0.5f * pw, y1 = pcy - 0.5f * phh;
    float x2 = pcx + 0.5f * pw - 1.f, y2 = pcy + 0.5f * phh - 1.f;
    if (!a.amodal) {
      x1 = fminf(fmaxf(x1, 0.f), (float)a.img_w - 1.f), y1 = fminf(fmaxf(y1, 0.f), (float)a.img_h - 1.f);
      x2 = fminf(fmaxf(x2, 0.f), (float)a.img_w - 1.f), y2 = fminf(fmaxf(y2, 0.f), (float)a.img_h - 1.f);
    }
    const bool big = (x2 - x1 + 1.f) >= a.min_size && (y2 - y1 + 1.f) >= a.min_size;
    reinterpret_cast<float4*>(cb)[j] = make_float4(x1, y1, x2, y2);
    cs[j] = big ? score : -1.f;
  }
  if (threadIdx.x == 0) a.cand_count[blockIdx.x] = k;
}

// ---------------------------------------------------------------------------------------------
// box head: softmax + per-class decode
// ---------------------------------------------------------------------------------------------
__global__ void box_decode_kernel(const float* __restrict__ head, int head_ld, const float* __restrict__ rois,
                                  const int* count, int n_max, int ncls, float wx, float wy, float ww, float wh,
                                  int img_w, int img_h, int amodal, const int* track_labels,
                                  float* __restrict__ out_boxes, float* __restrict__ out_scores) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_max) return;
  const int n = count ? min(*count, n_max) : n_max;
  if (r >= n) {
    for (int j = 0; j < ncls; ++j) {
      out_scores[(size_t)r * ncls + j] = -1.f;
      reinterpret_cast<float4*>(out_boxes)[(size_t)r * ncls + j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  const float* row = head + (size_t)r * head_ld;
  float mx = row[0];
  for (int j = 1; j < ncls; ++j) mx = fmaxf(mx, row[j]);
  float den = 0.f;
  for (int j = 0; j < ncls; ++j) den += expf(row[j] - mx);
  const float4 b = reinterpret_cast<const float4*>(rois)[r];
  const float w = b.z - b.x + 1.f, h = b.w - b.y + 1.f;
  const float cx = b.x + 0.5f * w, cy = b.y + 0.5f * h;
  const int label = track_labels ? track_labels[r] : -1;
  for (int j = 0; j < ncls; ++j) {
    float prob = __fdiv_rn(expf(row[j] - mx), den);
    if (track_labels) prob = (j == label) ? prob + 1.f : 0.f;
    const float* d = row + ncls + 4 * j;
    const float dx = __fdiv_rn(d[0], wx), dy = __fdiv_rn(d[1], wy);
    const float dw = fminf(__fdiv_rn(d[2], ww), BBOX_XFORM_CLIP), dh = fminf(__fdiv_rn(d[3], wh), BBOX_XFORM_CLIP);
    const float pcx = dx * w + cx, pcy = dy * h + cy;
    const float pw = expf(dw) * w, phh = expf(dh) * h;
    float x1 = pcx - 0.5f * pw, y1 = pcy - 0.5f * phh;
    float x2 = pcx + 0.5f * pw - 1.f, y2 = pcy + 0.5f * phh - 1.f;
    if (!amodal) {
      x1 = fminf(fmaxf(x1, 0.f), (float)img_w - 1.f), y1 = fminf(fmaxf(y1, 0.f), (float)img_h - 1.f);
      x2 = fminf(fmaxf(x2, 0.f), (float)img_w - 1.f), y2 = fminf(fmaxf(y2, 0.f), (float)img_h - 1.f);
    }
    out_scores[(size_t)r * ncls + j] = prob;
    reinterpret_cast<float4*>(out_boxes)[(size_t)r * ncls + j] = make_float4(x1, y1, x2, y2);
  }
}

// ---------------------------------------------------------------------------------------------
// candidate assembly for the solver: detections ++ refined tracks (roi_heads.py:60-84,44; track_solver.py:69)
// ---------------------------------------------------------------------------------------------
__global__ void track_combine_kernel(const float* __restrict__ det_boxes, const float* __restrict__ det_scores, int ncap,
                                     const float* __restrict__ dec_boxes, const float* __restrict__ dec_scores, int ncls,
                                     const int* __restrict__ labels, const float* __restrict__ conf,
                                     const int* __restrict__ valid, const float* __restrict__ active, int n, int tracktor,
                                     float* __restrict__ cat_boxes, float* __restrict__ cat_scores, int* zero_me) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && zero_me) *zero_me = 0;
  if (i < ncap) {
    reinterpret_cast<float4*>(cat_boxes)[i] = reinterpret_cast<const float4*>(det_boxes)[i];
    cat_scores[i] = det_scores[i];
  } else if (i < ncap + n) {
    const int r = i - ncap;
    const int lab = labels[r];
    const float det_part = dec_scores[(size_t)r * ncls + lab];  // p + 1 (inference.py:103)
    float s = det_part;
    if (!tracktor) s = __fdiv_rn(det_part + (conf[r] + 1.f), 2.f);  // roi_heads.py:67,76
    s = s + active[r];                                              // track_solver.py:69 (active rows +1)
    reinterpret_cast<float4*>(cat_boxes)[i] = reinterpret_cast<const float4*>(dec_boxes)[(size_t)r * ncls + lab];
    cat_scores[i] = valid[r] ? s : -1.f;
  }
}

}  // namespace smot

using namespace smot;

extern "C" int smot_track_combine(const float* det_boxes, const float* det_scores, int ncap, const float* dec_boxes,
                                  const float* dec_scores, int ncls, const int* labels, const float* conf, const int* valid,
                                  const float* active, int n, int tracktor, float* cat_boxes, float* cat_scores,
                                  int* zero_count, void* stream) {
  SMOT_CHECK_ARG(ncap >= 0 && n >= 0 && cat_boxes && cat_scores, "smot_track_combine: bad arguments");
  SMOT_CHECK_ARG(ncap == 0 || (det_boxes && det_scores), "smot_track_combine: null detections");
  SMOT_CHECK_ARG(n == 0 || (dec_boxes && dec_scores && labels && conf && valid && active && ncls >= 2),
                 "smot_track_combine: null track arrays");
  const int total = ncap + n;
  track_combine_kernel<<<(total + 1 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      det_boxes, det_scores, ncap, dec_boxes, dec_scores, ncls, labels, conf, valid, active, n, tracktor, cat_boxes,
      cat_scores, zero_count);
  SMOT_CHECK_LAUNCH("smot_track_combine");
  return SMOT_OK;
}

extern "C" size_t smot_sort_nms_workspace(int n_max) {
  if (n_max <= 0) return 0;
  return (size_t)n_max * ((n_max + 63) / 64) * sizeof(unsigned long long);
}

extern "C" int smot_sort_nms(const float* boxes, int box_stride, const float* scores, int score_stride, const int* count,
                             int n_max, float min_score, float thresh, int max_keep, int tag, int* out_index,
                             float* out_boxes, float* out_scores, int* out_tag, int* out_count, void* workspace,
                             size_t workspace_bytes, void* stream) {
  SMOT_CHECK_ARG(out_count, "smot_sort_nms: out_count is required");
  SMOT_CHECK_ARG(n_max >= 0 && n_max <= SN_MAX, "smot_sort_nms: n_max %d out of range [0,%d]", n_max, SN_MAX);
  if (n_max == 0) return SMOT_OK;
  SMOT_CHECK_ARG(boxes && scores && box_stride >= 4 && score_stride >= 1 && max_keep >= 0, "smot_sort_nms: bad arguments");
  SMOT_CHECK_ARG(thresh <= 0.f || (workspace && workspace_bytes >= smot_sort_nms_workspace(n_max)),
                 "smot_sort_nms: workspace too small (%zu < %zu)", workspace_bytes, smot_sort_nms_workspace(n_max));
  SortNmsArgs a;
  a.boxes = boxes, a.box_stride = box_stride, a.scores = scores, a.score_stride = score_stride, a.count = count;
  a.n_max = n_max, a.min_score = min_score, a.thresh = thresh, a.max_keep = max_keep, a.tag = tag;
  a.append = 1, a.fill_tail = 0;
  a.out_index = out_index, a.out_boxes = out_boxes, a.out_scores = out_scores, a.out_tag = out_tag, a.out_count = out_count;
  a.mask = (unsigned long long*)workspace;
  a.in_step = 0, a.out_step = 0, a.mask_step = 0;
  return launch_sort_nms(a, 1, (cudaStream_t)stream);
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t smot_rpn_select_workspace(int num_levels, int pre_nms_top_n) {
  if (num_levels <= 0 || pre_nms_top_n <= 0) return 0;
  const size_t L = (size_t)num_levels, P = (size_t)pre_nms_top_n;
  size_t b = 0;
  b += align256(L * P * 16);                 // cand_boxes
  b += align256(L * P * 4);                  // cand_scores
  b += align256(L * 4);                      // cand_count
  b += align256(L * P * 16);                 // kept boxes per level
  b += align256(L * P * 4);                  // kept scores per level
  b += align256(L * 4);                      // kept count per level
  b += align256(L * P * ((P + 63) / 64) * 8);  // nms masks
  return b;
}

extern "C" int smot_rpn_select(const smot_rpn_level* levels, int num_levels, int pre_nms_top_n, int post_nms_top_n,
                               float nms_thresh, float min_size, int fpn_post_nms_top_n, int img_w, int img_h,
                               int amodal, float* out_boxes, float* out_scores, int* out_count, void* workspace,
                               size_t workspace_bytes, void* stream) {
  SMOT_CHECK_ARG(levels && out_boxes && out_scores && out_count && workspace, "smot_rpn_select: null argument");
  SMOT_CHECK_ARG(num_levels >= 1 && num_levels <= SMOT_MAX_LEVELS, "smot_rpn_select: num_levels %d", num_levels);
  SMOT_CHECK_ARG(pre_nms_top_n >= 1 && pre_nms_top_n <= 1024, "smot_rpn_select: pre_nms_top_n %d not in [1,1024]", pre_nms_top_n);
  SMOT_CHECK_ARG(post_nms_top_n >= 1 && post_nms_top_n <= pre_nms_top_n, "smot_rpn_select: post_nms_top_n %d", post_nms_top_n);
  SMOT_CHECK_ARG(num_levels * post_nms_top_n <= SN_MAX, "smot_rpn_select: levels*post_nms_top_n > %d", SN_MAX);
  SMOT_CHECK_ARG(workspace_bytes >= smot_rpn_select_workspace(num_levels, pre_nms_top_n), "smot_rpn_select: workspace too small");
  for (int l = 0; l < num_levels; ++l)
    SMOT_CHECK_ARG(levels[l].head && levels[l].A >= 1 && levels[l].A <= SMOT_MAX_ANCHORS && levels[l].H > 0 && levels[l].W > 0 &&
                       levels[l].head_ld >= 5 * levels[l].A,
                   "smot_rpn_select: bad level %d", l);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t L = (size_t)num_levels, P = (size_t)pre_nms_top_n;
  unsigned char* w = (unsigned char*)workspace;
  float* cand_boxes = (float*)w;   w += align256(L * P * 16);
  float* cand_scores = (float*)w;  w += align256(L * P * 4);
  int* cand_count = (int*)w;       w += align256(L * 4);
  float* kept_boxes = (float*)w;   w += align256(L * P * 16);
  float* kept_scores = (float*)w;  w += align256(L * P * 4);
  int* kept_count = (int*)w;       w += align256(L * 4);
  unsigned long long* mask = (unsigned long long*)w;

  RpnArgs ra;
  for (int l = 0; l < num_levels; ++l) ra.lv[l] = levels[l];
  ra.pre_nms_top_n = pre_nms_top_n, ra.min_size = min_size, ra.img_w = img_w, ra.img_h = img_h, ra.amodal = amodal;
  ra.cand_boxes = cand_boxes, ra.cand_scores = cand_scores, ra.cand_count = cand_count;
  rpn_topk_kernel<<<num_levels, 1024, 0, st>>>(ra);
  SMOT_CHECK_LAUNCH("smot_rpn_select(topk)");

  // per-level NMS (one CTA per level), survivors into slots of post_nms_top_n rows, tails marked -1
  SortNmsArgs a;
  a.boxes = cand_boxes, a.box_stride = 4, a.scores = cand_scores, a.score_stride = 1, a.count = cand_count;
  a.n_max = pre_nms_top_n, a.min_score = -0.5f, a.thresh = nms_thresh, a.max_keep = post_nms_top_n, a.tag = 0;
  a.append = 0, a.fill_tail = post_nms_top_n;
  a.out_index = nullptr, a.out_boxes = kept_boxes, a.out_scores = kept_scores, a.out_tag = nullptr, a.out_count = kept_count;
  a.mask = mask;
  a.in_step = pre_nms_top_n, a.out_step = post_nms_top_n, a.mask_step = (int)(P * ((P + 63) / 64));
  int rc = launch_sort_nms(a, num_levels, st);
  if (rc) return rc;

  // cross-level top-k (sort only), level-major order among equal scores
  cudaError_t e = cudaMemsetAsync(out_count, 0, sizeof(int), st);
  if (e != cudaSuccess) {
    set_error("smot_rpn_select: memset failed: %s", cudaGetErrorString(e));
    return SMOT_ERR_CUDA;
  }
  SortNmsArgs m;
  m.boxes = kept_boxes, m.box_stride = 4, m.scores = kept_scores, m.score_stride = 1, m.count = nullptr;
  m.n_max = num_levels * post_nms_top_n, m.min_score = -0.5f, m.thresh = 0.f, m.max_keep = fpn_post_nms_top_n, m.tag = 0;
  m.append = 1, m.fill_tail = 0;
  m.out_index = nullptr, m.out_boxes = out_boxes, m.out_scores = out_scores, m.out_tag = nullptr, m.out_count = out_count;
  m.mask = nullptr, m.in_step = 0, m.out_step = 0, m.mask_step = 0;
  return launch_sort_nms(m, 1, st);
}

extern "C" int smot_box_decode(const float* head, int head_ld, const float* rois, const int* count, int n_max, int ncls,
                               const float* weights4, int img_w, int img_h, int amodal, const int* track_labels,
                               float* out_boxes, float* out_scores, void* stream) {
  SMOT_CHECK_ARG(n_max >= 0 && ncls >= 2 && head_ld >= 5 * ncls && weights4, "smot_box_decode: bad arguments");
  if (n_max == 0) return SMOT_OK;
  SMOT_CHECK_ARG(head && rois && out_boxes && out_scores, "smot_box_decode: null argument");
  box_decode_kernel<<<(n_max + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      head, head_ld, rois, count, n_max, ncls, weights4[0], weights4[1], weights4[2], weights4[3], img_w, img_h, amodal,
      track_labels, out_boxes, out_scores);
  SMOT_CHECK_LAUNCH("smot_box_decode");
  return SMOT_OK;
}

// tcgen05 / TMA implicit-GEMM convolution for sm_100a (fp16 operands, fp32 accumulation in TMEM).
//
// GEMM view as in conv_simt.cu: M = output pixels, N = Cout, K = taps * Cin.  One CTA produces a
// 128 (pixels) x BN (channels) output tile; the 128 pixels are a TILE_W x TILE_H patch of one image
// (16x8 for feature maps, 128x1 for matrices), so that for filter tap (r,s) the A operand of the
// tile is ONE 4-D TMA box of the NHWC input at offset (s-pad, r-pad): im2col is never built, and
// the zero padding of the convolution is TMA's out-of-bounds fill.  K advances over
// taps x (Cin/64): each step stages a 128x64 A box and a BNx64 weight box (both 128B-swizzled,
// K-major) in a shared-memory ring fed by one TMA-producer thread; one MMA thread issues
// 4 x tcgen05.mma (M=128, N=BN, K=16) per step into a TMEM accumulator and releases the ring slot
// with tcgen05.commit; four epilogue warps read the accumulator with tcgen05.ld, apply
// scale/bias (+residual) (+ReLU) and store fp16 NHWC with an arbitrary channel pitch (so DLA roots
// still read their children without a concat).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>
#include <mutex>
#include <string>
#include <unordered_map>

#include "common.cuh"

namespace smot {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;  // fp16 elements = 128 bytes = one swizzle row
constexpr int TC_THREADS = 192;

struct TcArgs {
  const float* scale;
  const float* bias;
  const __half* res;
  __half* out;
  int H, W;          // output (= input, stride 1) spatial size
  int Cin, Cout, out_ld, res_ld, relu;
  int taps, KW, pad, cin_chunks, stride;
  int tiles_w, tiles_h, tile_w, tile_h;
  // split-K (gridDim.z > 1): fp32 partial tiles [z][tile][128][Cout], summed by splitk_reduce_kernel
  int splits, chunks_per_split, num_tiles;
  int cluster_reduce;   // 1: the `splits` CTAs of a tile are one thread-block cluster and finish the tile themselves (no reduce kernel)
  // in-CTA K slices (splits == 1, slices > 1): ONE CTA runs all `slices` K ranges of `chunks_per_split` chunks, each into its own
  // TMEM accumulator (BN columns apart), and sums them in slice order in the epilogue -- the bits of the split-K path without
  // its partial tiles, its reduce kernel and its 8x CTA count
  int slices, tmem_cols;
  float* partial;
  unsigned long long* dbg;  // developer timing probe (SMOT_TC_DEBUG): 8 timestamps of CTA (0,0,0), else null
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a mis-programmed pipeline must fail the launch, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 26)) {
      printf("smot conv_tc: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(i)                                                                                   \
  do {                                                                                                \
    if (a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) a.dbg[i] = gtimer();          \
  } while (0)
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// SM100 shared-memory matrix descriptor: K-major, 128B swizzle, 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)0 << 16;                            // leading byte offset (unused: one atom along K)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset
  d |= (uint64_t)1 << 46;                            // descriptor version (SM100)
  d |= (uint64_t)2 << 61;                            // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: D=f32, A=B=f16, both K-major, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  // the registers are only valid after wait::ld: tie them to the wait so nothing is scheduled across it
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- epilogue warps 2..5: TMEM -> scale / bias (+residual) (+ReLU) -> fp16 NHWC, or the raw fp32 partial tile when K is split
template <int BN, bool SLICES_OK>
__device__ __forceinline__ void tc_epilogue(const TcArgs& a, uint32_t tmem_base, float* s_scale, float* s_bias, uint64_t* tmem_full,
                                            int img, int h0, int w0, int n0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int oh = h0 + row / a.tile_w, ow = w0 + row % a.tile_w;
    const bool valid = oh < a.H && ow < a.W;
    const size_t pix = ((size_t)img * a.H + oh) * a.W + ow;
    // while the main loop runs: park this N-tile's scale / bias in shared memory (read as float4 broadcasts later)
    for (int i = threadIdx.x - 64; i < BN; i += 128) {
      s_scale[i] = a.scale ? __ldg(a.scale + n0 + i) : 1.f;
      s_bias[i] = a.bias ? __ldg(a.bias + n0 + i) : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    auto finish = [&](float* v, int c0) {  // scale / bias / residual / ReLU / fp16 store of 32 channels
      const int n = n0 + c0;
      if (a.scale) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 sc = *reinterpret_cast<const float4*>(s_scale + c0 + 4 * g);
          v[4 * g] = __fmul_rn(v[4 * g], sc.x), v[4 * g + 1] = __fmul_rn(v[4 * g + 1], sc.y);
          v[4 * g + 2] = __fmul_rn(v[4 * g + 2], sc.z), v[4 * g + 3] = __fmul_rn(v[4 * g + 3], sc.w);
        }
      }
      if (a.bias) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 bi = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * g);
          v[4 * g] = __fadd_rn(v[4 * g], bi.x), v[4 * g + 1] = __fadd_rn(v[4 * g + 1], bi.y);
          v[4 * g + 2] = __fadd_rn(v[4 * g + 2], bi.z), v[4 * g + 3] = __fadd_rn(v[4 * g + 3], bi.w);
        }
      }
      if (a.res) {
        const uint4* rp = reinterpret_cast<const uint4*>(a.res + pix * a.res_ld + n);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 rv = __ldg(rp + g);
          const __half2* h2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 f = __half22float2(h2[e]);
            v[g * 8 + 2 * e] += f.x;
            v[g * 8 + 2 * e + 1] += f.y;
          }
        }
      }
      if (a.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      uint4* op = reinterpret_cast<uint4*>(a.out + pix * a.out_ld + n);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 ov;
        __half2* h2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
        op[g] = ov;
      }
    };
    mbar_wait(tmem_full, 0u);
    if (warp == 2 && lane == 0) TC_STAMP(3);
    tc_fence_after();
    if (SLICES_OK && a.slices > 1) {
      // the accumulators of the K slices, summed in slice order from 0.f: splitk_reduce_kernel's operations on the same values
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
#pragma unroll 1
        for (int z = 0; z < a.slices; ++z) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(z * BN + c0), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += v[j];
        }
        if (valid) finish(acc, c0);
      }
    } else if (a.splits == 1) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (valid) finish(v, c0);
      }
    } else {
      // ---- split-K: park the raw fp32 partial tile [split][tile][128][Cout]; splitk_reduce_kernel finishes the layer
      const int tile_id = (int)blockIdx.x;
      float* mine = a.partial + (((size_t)blockIdx.z * a.num_tiles + tile_id) * TC_BM + row) * a.Cout + n0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int g = 0; g < 8; ++g) reinterpret_cast<float4*>(mine + c0)[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
      }
    }
  }

// ---- split-K finish inside the cluster (replaces splitk_reduce_kernel: 22 launches, ~6 us each, 10 % of a frame in round 1).
// The `splits` CTAs of an output tile are launched as ONE thread-block cluster (1,1,splits): they are co-scheduled by the
// hardware, so after parking their raw fp32 partial tiles in the L2-resident workspace they can meet at a cluster barrier and
// each finish 128/splits rows of the tile: sum the partials in split order (deterministic, the same order and the same
// fp32 operations as splitk_reduce_kernel, so the results are bit-identical to it), scale / bias / residual / ReLU, fp16 store.
// Per CTA: 128 x BN x 4 B of partials read back from L2 (.cg: written by other SMs), ~1 us, instead of a kernel boundary.
template <int BN>
__device__ __forceinline__ void tc_splitk_finish(const TcArgs& a, const float* s_scale, const float* s_bias, int img, int h0, int w0,
                                                 int n0) {
  // all six warps take part; every thread has FIN_ITEMS float4 positions x up to 8 partials in flight before it sums
  // (the slice is 128 x BN x 4 B per CTA read back from L2 at ~1 us latency: bytes in flight are what bounds it)
  constexpr int FIN_ITEMS = 3, MAXS = 8;
  const int tid = (int)threadIdx.x;
  const int rows_per = (TC_BM + a.splits - 1) / a.splits;
  const int r_lo = (int)blockIdx.z * rows_per;
  const int r_hi = min(TC_BM, r_lo + rows_per);
  constexpr int C4 = BN / 4;
  const int n_items = (r_hi - r_lo) * C4;
  const size_t zstride = (size_t)a.num_tiles * TC_BM * a.Cout;
  const float* tile = a.partial + (size_t)blockIdx.x * TC_BM * a.Cout + n0;
  for (int base = tid; base < n_items; base += TC_THREADS * FIN_ITEMS) {
    float4 v[FIN_ITEMS][MAXS];
    int row[FIN_ITEMS], col[FIN_ITEMS];
    bool ok[FIN_ITEMS];
#pragma unroll
    for (int it = 0; it < FIN_ITEMS; ++it) {
      const int idx = base + it * TC_THREADS;
      row[it] = r_lo + idx / C4, col[it] = (idx % C4) * 4;
      const int oh = h0 + row[it] / a.tile_w, ow = w0 + row[it] % a.tile_w;
      ok[it] = idx < n_items && oh < a.H && ow < a.W;
      const float* p = tile + (size_t)row[it] * a.Cout + col[it];
#pragma unroll
      for (int z = 0; z < MAXS; ++z)
        if (ok[it] && z < a.splits) v[it][z] = __ldcg(reinterpret_cast<const float4*>(p + (size_t)z * zstride));
    }
#pragma unroll
    for (int it = 0; it < FIN_ITEMS; ++it) {
      if (!ok[it]) continue;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int z = 0; z < MAXS; ++z)
        if (z < a.splits) acc.x += v[it][z].x, acc.y += v[it][z].y, acc.z += v[it][z].z, acc.w += v[it][z].w;   // split order
      float o[4] = {acc.x, acc.y, acc.z, acc.w};
      const int c = col[it];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (a.scale) o[j] = __fmul_rn(o[j], s_scale[c + j]);
        if (a.bias) o[j] = __fadd_rn(o[j], s_bias[c + j]);
      }
      const int oh = h0 + row[it] / a.tile_w, ow = w0 + row[it] % a.tile_w;
      const size_t pix = ((size_t)img * a.H + oh) * a.W + ow;
      if (a.res) {
        const float4 r = ld4(a.res + pix * a.res_ld + n0 + c);
        o[0] += r.x, o[1] += r.y, o[2] += r.z, o[3] += r.w;
      }
      if (a.relu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], 0.f);
      }
      st4(a.out + pix * a.out_ld + n0 + c, make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

template <int BN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + BN * 8 /*scale, bias*/;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  using S = TcSmem<BN, STAGES>;
  extern __shared__ uint8_t tc_smem_raw[];
  // 128B-swizzled tiles need 1024B-aligned bases
  uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * S::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  float* s_scale = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + 256);  // [BN] scale, then [BN] bias
  float* s_bias = s_scale + BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) TC_STAMP(0);
  // tile coordinates
  int t = blockIdx.x;
  const int tw = t % a.tiles_w;
  t /= a.tiles_w;
  const int th = t % a.tiles_h;
  const int img = t / a.tiles_h;
  const int w0 = tw * a.tile_w, h0 = th * a.tile_h;
  const int n0 = blockIdx.y * BN;
  const int all_chunks = a.taps * a.cin_chunks;
  const bool sliced = STAGES > 2 && a.slices > 1;                        // all K slices in this CTA, one accumulator each
  const int it0 = sliced ? 0 : (int)blockIdx.z * a.chunks_per_split;    // this CTA's K range [it0, it0 + total)
  const int total = sliced ? all_chunks : min(a.chunks_per_split, all_chunks - it0);
  const uint32_t tmem_cols = sliced ? (uint32_t)a.tmem_cols : (uint32_t)BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation (whole warp), BN fp32 columns per accumulator
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel; its outputs are read (and buffers rewritten) from here on
  if (threadIdx.x == 0) TC_STAMP(1);

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      for (int it = 0; it < total; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], (uint32_t)S::STAGE_BYTES);
        const int tap = (it0 + it) / a.cin_chunks, cc = (it0 + it) - tap * a.cin_chunks;
        const int r = tap / a.KW, sx = tap - r * a.KW;
        tma_load_4d(sA + s * S::A_BYTES, &tmA, &full[s], cc * TC_BK, w0 * a.stride + sx - a.pad, h0 * a.stride + r - a.pad, img);
        tma_load_2d(sB + s * S::B_BYTES, &tmB, &full[s], tap * a.Cin + cc * TC_BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = umma_idesc_f16(TC_BM, BN);
      for (int it = 0; it < total; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        if (it == 0) TC_STAMP(2);
        tc_fence_after();
        const uint64_t ad = umma_desc_sw128(sA + s * S::A_BYTES);
        const uint64_t bd = umma_desc_sw128(sB + s * S::B_BYTES);
        const int sl = sliced ? it / a.chunks_per_split : 0, itl = it - sl * a.chunks_per_split;   // slice, chunk inside it
        const uint32_t d_tmem = tmem_base + (uint32_t)(sl * BN);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k)  // +32 bytes (2 x 16B units) per K=16 step inside the swizzle atom
          umma_f16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (uint32_t)((itl | k) != 0));
        umma_commit(&empty[s]);  // slot reusable once these MMAs have read it
      }
      umma_commit(tmem_full);    // accumulator(s) complete
    }
  } else {  // ===== epilogue warps 2..5 =====
    tc_epilogue<BN, (STAGES > 2)>(a, tmem_base, s_scale, s_bias, tmem_full, img, h0, w0, n0);
  }
  if (warp == 2 && lane == 0) TC_STAMP(4);
  if constexpr (STAGES > 2)   // the 2-stage variants run 4 CTAs per SM on full-GPU layers: they never split K, keep their registers low
  if (a.cluster_reduce) {
    // every CTA of the tile's cluster has parked its partial: publish (gpu scope), meet, finish 128 / splits rows each
    __threadfence();
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_splitk_finish<BN>(a, s_scale, s_bias, img, h0, w0, n0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TC_STAMP(5);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// 3x3 / stride 1 with a shared-memory resident input halo.
// conv_tc_kernel re-reads the A operand once per filter tap: 9 x 16 KB per 64-channel chunk and CTA.  Measured,
// those layers sit at the L2 -> SM throughput cap (~11.5 TB/s over all SMs, profiles/conv_layers_r01_*.txt), not at
// the tensor pipe.  Here the (8+2) x (16+2) pixel halo of an 8 x 16 output tile is loaded ONCE per channel chunk
// (one 4-D TMA box, PW pixels per patch row, 128B-swizzled, zero fill = conv padding) and the nine taps are nine
// shifted views of it: row m = y*8 + x of tap (r, s) is patch pixel (y + r, x + s), i.e. the UMMA descriptor starts
// (r*PW + s) * 128 B into the patch and steps PW * 128 B between its 8-row groups.  Only the weights still stream per
// tap (their own ring).  K order: channel chunk outermost, taps inside.
// ---------------------------------------------------------------------------------------------
constexpr int HALO_TW = 8, HALO_TH = 16, HALO_SA = 2;

template <int BN, int PW, int SB>
struct HaloSmem {
  static constexpr int A_BOX_BYTES = (HALO_TH + 2) * PW * TC_BK * 2;              // bytes one TMA patch delivers
  static constexpr int A_BYTES = ((A_BOX_BYTES + 1023) / 1024) * 1024;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int RING = HALO_SA * A_BYTES + SB * B_BYTES;
  static constexpr int TOTAL = RING + 1024 /*alignment slack*/ + 256 /*barriers*/ + BN * 8 /*scale, bias*/;
  static_assert((2 * HALO_SA + 2 * SB + 2) * 8 <= 256, "barrier block");
};

__device__ __forceinline__ uint64_t umma_desc_sw128_rows(const void* smem, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7u) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN, int PW, int SB>
__global__ void __launch_bounds__(TC_THREADS) conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB, const TcArgs a,
                                                                  const int bo_mode) {
  using S = HaloSmem<BN, PW, SB>;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + HALO_SA * S::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::RING);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + HALO_SA;
  uint64_t* fullB = emptyA + HALO_SA;
  uint64_t* emptyB = fullB + SB;
  uint64_t* tmem_full = emptyB + SB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* s_scale = reinterpret_cast<float*>(smem + S::RING + 256);
  float* s_bias = s_scale + BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  int t = blockIdx.x;
  const int tw = t % a.tiles_w;
  t /= a.tiles_w;
  const int th = t % a.tiles_h;
  const int img = t / a.tiles_h;
  const int w0 = tw * HALO_TW, h0 = th * HALO_TH;
  const int n0 = blockIdx.y * BN;
  const int cc0 = (int)blockIdx.z * a.chunks_per_split;                  // this CTA's channel chunks [cc0, cc0 + ncc)
  const int ncc = min(a.chunks_per_split, a.cin_chunks - cc0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < HALO_SA; ++i) mbar_init(&fullA[i], 1), mbar_init(&emptyA[i], 1);
    for (int i = 0; i < SB; ++i) mbar_init(&fullB[i], 1), mbar_init(&emptyB[i], 1);
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // developer probes (timing only, results are garbage): 256 = no MMAs, 512 = no weight loads, 1024 = no patch loads
  const bool no_mma = bo_mode & 256, no_b = bo_mode & 512, no_a = bo_mode & 1024;
  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: one halo patch per channel chunk, one weight tile per tap =====
      int ib = 0;
      for (int ic = 0; ic < ncc; ++ic) {
        const int sa = ic % HALO_SA;
        if (!no_a) {
          mbar_wait(&emptyA[sa], ((uint32_t)(ic / HALO_SA) & 1u) ^ 1u);
          mbar_expect_tx(&fullA[sa], (uint32_t)S::A_BOX_BYTES);
          tma_load_4d(sA + sa * S::A_BYTES, &tmA, &fullA[sa], (cc0 + ic) * TC_BK, w0 - 1, h0 - 1, img);
        }
        for (int tap = 0; tap < 9 && !no_b; ++tap, ++ib) {
          const int sb = ib % SB;
          mbar_wait(&emptyB[sb], ((uint32_t)(ib / SB) & 1u) ^ 1u);
          mbar_expect_tx(&fullB[sb], (uint32_t)S::B_BYTES);
          tma_load_2d(sB + sb * S::B_BYTES, &tmB, &fullB[sb], tap * a.Cin + (cc0 + ic) * TC_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = umma_idesc_f16(TC_BM, BN);
      int ib = 0;
      for (int ic = 0; ic < ncc; ++ic) {
        const int sa = ic % HALO_SA;
        if (!no_a) mbar_wait(&fullA[sa], (uint32_t)(ic / HALO_SA) & 1u);
        const uint8_t* patch = sA + sa * S::A_BYTES;
        for (int tap = 0; tap < 9; ++tap, ++ib) {
          const int sb = ib % SB;
          if (!no_b) mbar_wait(&fullB[sb], (uint32_t)(ib / SB) & 1u);
          tc_fence_after();
          const int r = tap / 3, sx = tap - 3 * r;
          const uint32_t first_row = (uint32_t)(r * PW + sx);   // patch pixel of output row 0 for this tap
          const uint64_t ad = umma_desc_sw128_rows(patch + first_row * 128u, (uint32_t)PW * 128u, (bo_mode & 1) ? first_row : 0u);
          const uint64_t bd = umma_desc_sw128(sB + sb * S::B_BYTES);
          if (!no_mma) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              umma_f16(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (uint32_t)((ic | tap | k) != 0));
          }
          if (!no_b) umma_commit(&emptyB[sb]);
        }
        if (!no_a) umma_commit(&emptyA[sa]);   // all nine taps of this patch have been read
      }
      umma_commit(tmem_full);
    }
  } else {
    tc_epilogue<BN, false>(a, tmem_base, s_scale, s_bias, tmem_full, img, h0, w0, n0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// Deterministic split-K finish: out[pixel][n] = act((sum_z partial[z]) * scale + bias + residual), summed in split order.
// One thread per 4 output channels of one pixel; tiles map back to pixels exactly as in conv_tc_kernel.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const TcArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4 = a.Cout / 4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)a.num_tiles * TC_BM * c4;
  if (idx >= total) return;
  const int n = (int)(idx % c4) * 4;
  const size_t rowg = idx / c4;  // tile * 128 + row
  const int row = (int)(rowg % TC_BM);
  int t = (int)(rowg / TC_BM);
  const int tw = t % a.tiles_w;
  t /= a.tiles_w;
  const int th = t % a.tiles_h;
  const int img = t / a.tiles_h;
  const int oh = th * a.tile_h + row / a.tile_w, ow = tw * a.tile_w + row % a.tile_w;
  if (oh >= a.H || ow >= a.W) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int z = 0; z < a.splits; ++z) {
    const float4 p = *reinterpret_cast<const float4*>(a.partial + ((size_t)z * a.num_tiles * TC_BM + rowg) * a.Cout + n);
    acc.x += p.x, acc.y += p.y, acc.z += p.z, acc.w += p.w;
  }
  float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (a.scale) v[j] = __fmul_rn(v[j], a.scale[n + j]);
    if (a.bias) v[j] = __fadd_rn(v[j], a.bias[n + j]);
  }
  const size_t pix = ((size_t)img * a.H + oh) * a.W + ow;
  if (a.res) {
    const float4 r = ld4(a.res + pix * a.res_ld + n);
    v[0] += r.x, v[1] += r.y, v[2] += r.z, v[3] += r.w;
  }
  if (a.relu) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  st4(a.out + pix * a.out_ld + n, make_float4(v[0], v[1], v[2], v[3]));
}

// ---- host side ------------------------------------------------------------------------------
constexpr int SMOT_TC_DEFAULT_MAXSPLIT = 8;
static PFN_cuTensorMapEncodeTiled get_encode() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  }
  return fn;
}

static bool encode_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, const uint32_t* estr_in = nullptr) {
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  if (!enc) {
    set_error("smot_conv2d(tcgen05): cuTensorMapEncodeTiled entry point not available");
    return false;
  }
  uint32_t estr[4] = {1, 1, 1, 1};
  if (estr_in)
    for (int i = 0; i < rank; ++i) estr[i] = estr_in[i];
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("smot_conv2d(tcgen05): cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return false;
  }
  return true;
}

// Widest split of the K loop of a few-tile layer (levels 4 / 5, FC layers), finished by splitk_reduce_kernel.
// SMOT_TC_MAXSPLIT=1..8 (SMOT_TC_NOSPLIT=1 is MAXSPLIT=1).  Measured on B200 in the clip pipeline (profiles/bench_r02m_*.json):
// splitting 8 ways makes the static stage ALONE faster (0.646 vs 0.658 ms) but the pipeline slower (1228 vs 1277 frames/s):
// 8 CTAs per tile plus a 960-CTA reduce kernel per layer occupy SMs that the detection tail and the track stage of the
// neighbouring frames would otherwise use.
static int tc_max_split() {
  static const int v = [] {
    if (getenv("SMOT_TC_NOSPLIT")) return 1;
    const char* e = getenv("SMOT_TC_MAXSPLIT");
    const int m = e ? atoi(e) : SMOT_TC_DEFAULT_MAXSPLIT;
    return m < 1 ? 1 : (m > 8 ? 8 : m);
  }();
  return v;
}

bool conv2d_tc_supported(const smot_conv_desc* d) {
  if (d->in_dtype != SMOT_F16 || d->out_dtype != SMOT_F16) return false;
  if (d->KH != d->KW || (d->KH != 1 && d->KH != 3) || d->pad != d->KH / 2) return false;
  if (d->stride != 1 && !(d->stride == 2 && d->KH == 3 && d->H % 2 == 0 && d->W % 2 == 0 && d->H > 1)) return false;
  if (d->Cin % TC_BK != 0 || d->Cout % 64 != 0) return false;
  if (d->in_ld % 8 != 0 || d->out_ld % 8 != 0 || (d->residual && d->res_ld % 8 != 0)) return false;
  if (((uintptr_t)d->in | (uintptr_t)d->weight | (uintptr_t)d->out | (uintptr_t)d->residual) & 15) return false;
  if (d->batch < 1 || d->OH != d->H / d->stride || d->OW != d->W / d->stride) return false;
  if ((long long)d->batch * d->OH * d->OW < 16) return false;  // not worth a 128-row tile
  return true;
}

// launch attributes: programmatic dependent launch always; a (1,1,cluster_z) thread-block cluster when the CTAs of a tile
// finish their split-K sum themselves
template <int BN, int STAGES>
static cudaError_t launch_tc_ex(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, dim3 grid, int cluster_z,
                                cudaStream_t st) {
  using S = TcSmem<BN, STAGES>;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = dim3(TC_THREADS), cfg.dynamicSmemBytes = S::TOTAL, cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr, cfg.numAttrs = 1;
  if (cluster_z > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 1, attr[1].val.clusterDim.y = 1, attr[1].val.clusterDim.z = (unsigned)cluster_z;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, STAGES>, tmA, tmB, a);
}

template <int BN, int STAGES>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, dim3 grid, cudaStream_t st) {
  using S = TcSmem<BN, STAGES>;
  SMOT_ENSURE_SMEM((conv_tc_kernel<BN, STAGES>), S::TOTAL, "smot_conv2d(tcgen05)");
  cudaError_t e = launch_tc_ex<BN, STAGES>(tmA, tmB, a, grid, a.cluster_reduce ? a.splits : 1, st);
  if (e != cudaSuccess) {
    set_error("smot_conv2d(tcgen05): launch failed: %s", cudaGetErrorString(e));
    return SMOT_ERR_CUDA;
  }
  SMOT_CHECK_LAUNCH("smot_conv2d(tcgen05)");
  return SMOT_OK;
}

// How many (1,1,cz) clusters of this kernel variant the device can hold at once (a cluster lives inside one GPC, so this is
// NOT 148 / cz: 8-CTA clusters fit twice into a 16..20-SM GPC).  Queried once per (variant, cz); 0 = query failed.
template <int BN, int STAGES>
static int tc_max_clusters(int cz) {
  static std::mutex mu;
  static int cache[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};
  if (cz < 2 || cz > 8) return 0;
  std::lock_guard<std::mutex> lock(mu);
  if (cache[cz] >= 0) return cache[cz];
  using S = TcSmem<BN, STAGES>;
  int n = 0;
  if (cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1, 1, (unsigned)cz), cfg.blockDim = dim3(TC_THREADS), cfg.dynamicSmemBytes = S::TOTAL;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = (unsigned)cz;
    cfg.attrs = attr, cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<BN, STAGES>, &cfg) != cudaSuccess) n = 0;
  }
  (void)cudaGetLastError();
  cache[cz] = n;
  return n;
}

// ring depth of the generic kernel for a given tile width / CTA population / K length (measured choices, see conv2d_tc)
static int tc_stages(int BN, bool solo, bool shallow, int my_chunks, int forced) {
  if (BN == 256) return (forced ? forced >= 4 : (solo && my_chunks >= 4)) ? 4 : 3;
  const bool deep = forced ? forced >= 6 : (solo && my_chunks >= 6);
  if (BN == 128) return shallow ? 2 : (deep ? 6 : 3);
  return deep ? 8 : (shallow ? 2 : 4);
}

#define SMOT_TC_DISPATCH(BN_, ST_, EXPR)                                                       \
  ((BN_) == 256 ? ((ST_) == 4 ? EXPR(256, 4) : EXPR(256, 3))                                    \
   : (BN_) == 128 ? ((ST_) == 2 ? EXPR(128, 2) : ((ST_) == 6 ? EXPR(128, 6) : EXPR(128, 3)))   \
                  : ((ST_) == 8 ? EXPR(64, 8) : ((ST_) == 2 ? EXPR(64, 2) : EXPR(64, 4))))

template <int BN, int PW, int SB>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, dim3 grid, int bo_mode, cudaStream_t st) {
  using S = HaloSmem<BN, PW, SB>;
  SMOT_ENSURE_SMEM((conv3x3_halo_kernel<BN, PW, SB>), S::TOTAL, "smot_conv2d(tcgen05 halo)");
  launch_pdl(conv3x3_halo_kernel<BN, PW, SB>, grid, dim3(TC_THREADS), S::TOTAL, st, tmA, tmB, a, bo_mode);
  SMOT_CHECK_LAUNCH("smot_conv2d(tcgen05 halo)");
  return SMOT_OK;
}

static int halo_mode() {  // developer switch SMOT_TC_HALO: 0 = off, 16 / 10 = patch width, +100 = descriptor base offset mode
  const char* e = getenv("SMOT_TC_HALO");
  return e ? atoi(e) : 0;
}

static int conv2d_tc_halo(const smot_conv_desc* d, int mode, cudaStream_t st) {
  TcArgs a;
  a.scale = d->scale, a.bias = d->bias, a.res = (const __half*)d->residual, a.out = (__half*)d->out;
  a.H = d->OH, a.W = d->OW, a.stride = 1, a.Cin = d->Cin, a.Cout = d->Cout, a.out_ld = d->out_ld, a.res_ld = d->res_ld, a.relu = d->relu;
  a.taps = 9, a.KW = 3, a.pad = 1, a.cin_chunks = d->Cin / TC_BK;
  a.tile_w = HALO_TW, a.tile_h = HALO_TH;
  a.tiles_w = ceil_div(d->OW, HALO_TW), a.tiles_h = ceil_div(d->OH, HALO_TH);
  const long long tiles = (long long)a.tiles_w * a.tiles_h * d->batch;
  const int pw = mode % 100;
  int bo_mode = (mode / 100) & 1;
  if (const char* e = getenv("SMOT_TC_PROBE")) bo_mode |= atoi(e);
  int BN = 64;
  if (d->Cout % 256 == 0 && tiles * (d->Cout / 256) >= 96) BN = 256;
  else if (d->Cout % 128 == 0 && tiles * (d->Cout / 128) >= 96) BN = 128;
  int splits = 1;
  {
    const int bw = d->Cout % 256 == 0 ? 256 : (d->Cout % 128 == 0 ? 128 : 64);
    const long long cw = tiles * (d->Cout / bw);
    if (d->workspace && cw <= 40 && a.cin_chunks >= 2 && tc_max_split() > 1) {
      int sp = (int)(148 / cw);
      if (sp > tc_max_split()) sp = tc_max_split();
      if (sp > a.cin_chunks) sp = a.cin_chunks;
      const size_t need = (size_t)SMOT_CONV_WS_COUNTER_BYTES + (size_t)sp * tiles * TC_BM * d->Cout * sizeof(float);
      if (sp >= 2 && need <= d->workspace_bytes) splits = sp, BN = bw;
    }
  }
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->batch};
    uint64_t str[3] = {(uint64_t)d->in_ld * 2, (uint64_t)d->W * d->in_ld * 2, (uint64_t)d->H * d->W * d->in_ld * 2};
    uint32_t box[4] = {(uint32_t)TC_BK, (uint32_t)pw, (uint32_t)(HALO_TH + 2), 1u};
    if (!encode_map(&tmA, d->in, 4, dims, str, box)) return SMOT_ERR_CUDA;
  }
  {
    const uint64_t K = (uint64_t)9 * d->Cin;
    uint64_t dims[2] = {K, (uint64_t)d->Cout};
    uint64_t str[1] = {K * 2};
    uint32_t box[2] = {(uint32_t)TC_BK, (uint32_t)BN};
    if (!encode_map(&tmB, d->weight, 2, dims, str, box)) return SMOT_ERR_CUDA;
  }
  a.chunks_per_split = (a.cin_chunks + splits - 1) / splits;
  a.splits = (a.cin_chunks + a.chunks_per_split - 1) / a.chunks_per_split;
  a.num_tiles = (int)tiles;
  a.cluster_reduce = 0;
  a.slices = 1, a.tmem_cols = 0;
  a.dbg = nullptr;
  a.partial = d->workspace ? (float*)((char*)d->workspace + SMOT_CONV_WS_COUNTER_BYTES) : nullptr;
  dim3 grid((unsigned)tiles, (unsigned)(d->Cout / BN), (unsigned)a.splits);
  const bool crowded = (long long)grid.x * grid.y * grid.z > 148;   // several CTAs per SM: shallow weight ring, 2 CTAs / SM
  int rc;
#define SMOT_HALO(BN_, SB_) (pw == 10 ? launch_halo<BN_, 10, SB_>(tmA, tmB, a, grid, bo_mode, st) : launch_halo<BN_, 16, SB_>(tmA, tmB, a, grid, bo_mode, st))
  if (BN == 256) rc = SMOT_HALO(256, 4);
  else if (BN == 128) rc = crowded ? SMOT_HALO(128, 2) : SMOT_HALO(128, 6);
  else rc = crowded ? SMOT_HALO(64, 4) : SMOT_HALO(64, 8);
#undef SMOT_HALO
  if (rc != SMOT_OK || a.splits == 1) return rc;
  const size_t total_out = (size_t)a.num_tiles * TC_BM * (d->Cout / 4);
  launch_pdl(splitk_reduce_kernel, dim3((unsigned)((total_out + 255) / 256)), dim3(256), 0, st, a);
  SMOT_CHECK_LAUNCH("smot_conv2d(split-K reduce)");
  return SMOT_OK;
}

int conv2d_tc(const smot_conv_desc* d, cudaStream_t st) {
  if (d->KH == 3 && d->stride == 1 && d->H > 1 && (halo_mode() % 100 == 16 || halo_mode() % 100 == 10))
    return conv2d_tc_halo(d, halo_mode(), st);
  TcArgs a;
  a.scale = d->scale, a.bias = d->bias, a.res = (const __half*)d->residual, a.out = (__half*)d->out;
  a.H = d->OH, a.W = d->OW, a.stride = d->stride, a.Cin = d->Cin, a.Cout = d->Cout, a.out_ld = d->out_ld, a.res_ld = d->res_ld, a.relu = d->relu;
  a.taps = d->KH * d->KW, a.KW = d->KW, a.pad = d->pad, a.cin_chunks = d->Cin / TC_BK;
  if (d->H == 1) {
    a.tile_w = 128, a.tile_h = 1;
  } else {
    a.tile_w = 16, a.tile_h = 8;
  }
  a.tiles_w = ceil_div(d->OW, a.tile_w), a.tiles_h = ceil_div(d->OH, a.tile_h);
  const long long tiles = (long long)a.tiles_w * a.tiles_h * d->batch;
  // Tile shape.  Measured (profiles/conv_probe_r01.txt): the k-steps issue at the tensor-pipe floor, but two thirds of a layer
  // is per-kernel / per-CTA fixed cost (launch, TMEM allocation, epilogue).  Hence: (a) the widest N that still leaves ~100
  // CTAs; (b) layers with only a handful of output tiles (levels 4-5, FC layers) take the widest BN AND split K over up
  // to 8 CTAs, finished by splitk_reduce_kernel (not splitting, or splitting less, measured slower).
  const int all_chunks = a.taps * a.cin_chunks;
  static const int min_ctas = getenv("SMOT_TC_MINCTAS") ? atoi(getenv("SMOT_TC_MINCTAS")) : 96;   // developer override
  int BN = 64;
  if (d->Cout % 256 == 0 && tiles * (d->Cout / 256) >= min_ctas) BN = 256;
  else if (d->Cout % 128 == 0 && tiles * (d->Cout / 128) >= min_ctas) BN = 128;
  const char* force = getenv("SMOT_TC_STAGES");  // developer override: ring depth
  const int fs = force ? atoi(force) : 0;
  // developer switch SMOT_TC_CLUSTER=1: the split CTAs of a tile form a cluster and finish the tile themselves (tc_splitk_finish).
  // Measured on B200 (profiles/bench_r02c_*): correct (bit-identical to the reduce kernel) but SLOWER -- 18.7-25.9 us per
  // level-4 layer against 13.7 us + a 6 us reduce launch that overlaps the next layer's prologue through PDL; the serial
  // store -> fence -> cluster barrier -> 128 KB read-back of one CTA cannot match a reduce grid that spreads over all SMs.
  static const bool cluster_ok = getenv("SMOT_TC_CLUSTER") && atoi(getenv("SMOT_TC_CLUSTER")) == 1;
  int splits = 1, cluster_reduce = 0;
  {
    const int bw = d->Cout % 256 == 0 ? 256 : (d->Cout % 128 == 0 ? 128 : 64);
    // The split factor is a function of the tiles of ONE image: a batch-2 backbone pass (Engine.pair_plan) must sum every
    // output element in exactly the order a single-frame pass does, so that clip results equal frame-by-frame results bit
    // for bit.  (With two images the split layers then run two half-length waves instead of one: the same time.)
    const long long tiles_img = tiles / (d->batch > 0 ? d->batch : 1);
    const long long cw = tiles_img * (d->Cout / bw);
    if (d->workspace && cw <= 40 && all_chunks >= 16 && tc_max_split() > 1) {
      int sp = (int)(148 / cw);
      if (sp > tc_max_split()) sp = tc_max_split();
      if (sp > all_chunks / 4) sp = all_chunks / 4;
      const size_t need = (size_t)SMOT_CONV_WS_COUNTER_BYTES + (size_t)sp * tiles * TC_BM * d->Cout * sizeof(float);
      if (sp >= 2 && need <= d->workspace_bytes) {
        // preferred: the split CTAs of a tile form a cluster and finish the tile themselves; take the widest split whose
        // clusters are all resident at once (the cluster barrier needs co-residency only inside a cluster, but a second wave
        // would double the layer's time)
        for (int cz = cluster_ok ? sp : 0; cz >= 2; --cz) {
          const int cps = (all_chunks + cz - 1) / cz;
          const int real = (all_chunks + cps - 1) / cps;                       // no empty split
          if (real != cz) continue;
          const int stg = tc_stages(bw, true, false, cps, fs);
#define SMOT_TC_MAXC(BN_, ST_) tc_max_clusters<BN_, ST_>(cz)
          const int fit = SMOT_TC_DISPATCH(bw, stg, SMOT_TC_MAXC);
#undef SMOT_TC_MAXC
          if (fit >= cw) {   // (per image: a batch of two takes two waves of clusters)
            splits = cz, BN = bw, cluster_reduce = 1;
            break;
          }
        }
        if (!cluster_reduce) splits = sp, BN = bw;                             // fallback: partials + splitk_reduce_kernel
      }
    }
  }
  // In-CTA K slices (developer switch SMOT_TC_SLICED=1 | 128 | 256 = N tile; default off): the SAME K ranges, summed in the SAME
  // order -- bit-identical results (test_conv2d_tcgen05_k_slices_equal_split_k) -- but by one CTA with one TMEM accumulator per
  // range (BN x slices <= 512 columns), no partial tiles, no reduce kernel.  Measured pairwise on the same boxes
  // (profiles/bench_r02{r,s,t}_*.json; value / e2e frames/s, split -> sliced): 720p30 1243 / 1221 -> 1261 / 1262 (+1.5 / +3.4 %),
  // but 1080p80 801 / 785 -> 785 / 770, R-50 570 / 608 -> 572 / 582, model(frame) 687 -> 663; slicing only the 8-wide splits
  // (SMOT_TC_SLICED_MIN=8) lost on all three.  The split's extra CTAs fill the GPU where the neighbouring stages leave room and
  // crowd it where they do not; with no rule that holds across the three workloads, the validated split path stays the default.
  bool sliced = false;
  static const int slice_min = getenv("SMOT_TC_SLICED_MIN") ? atoi(getenv("SMOT_TC_SLICED_MIN")) : 2;
  if (splits >= slice_min && splits > 1 && !cluster_reduce) {
    const char* e = getenv("SMOT_TC_SLICED");
    if (e && e[0] != '0') {
      int want = atoi(e);
      if (want != 128 && want != 256) want = 64;
      int bn = 512 / splits;                       // splits <= 8 -> >= 64
      if (bn > BN) bn = BN;
      if (bn > want) bn = want;
      if (bn >= 64) sliced = true, BN = bn >= 256 ? 256 : (bn >= 128 ? 128 : 64);
    }
  }
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->batch};
    uint64_t str[3] = {(uint64_t)d->in_ld * 2, (uint64_t)d->W * d->in_ld * 2, (uint64_t)d->H * d->W * d->in_ld * 2};
    // stride-2 convs: the box is traversed with element stride 2 and still lands as tile_w x tile_h pixels
    uint32_t box[4] = {(uint32_t)TC_BK, (uint32_t)(a.tile_w * d->stride), (uint32_t)(a.tile_h * d->stride), 1u};
    uint32_t estr[4] = {1u, (uint32_t)d->stride, (uint32_t)d->stride, 1u};
    if (!encode_map(&tmA, d->in, 4, dims, str, box, estr)) return SMOT_ERR_CUDA;
  }
  {
    const uint64_t K = (uint64_t)a.taps * d->Cin;
    uint64_t dims[2] = {K, (uint64_t)d->Cout};
    uint64_t str[1] = {K * 2};
    uint32_t box[2] = {(uint32_t)TC_BK, (uint32_t)BN};
    if (!encode_map(&tmB, d->weight, 2, dims, str, box)) return SMOT_ERR_CUDA;
  }
  a.splits = splits;
  a.chunks_per_split = (all_chunks + splits - 1) / splits;
  a.splits = (all_chunks + a.chunks_per_split - 1) / a.chunks_per_split;  // no empty split
  a.cluster_reduce = (cluster_reduce && a.splits == splits) ? 1 : 0;
  a.slices = 1, a.tmem_cols = 0;
  if (sliced) {
    a.slices = a.splits, a.splits = 1;
    int cols = 32;
    while (cols < a.slices * BN) cols <<= 1;
    a.tmem_cols = cols;
  }
  a.num_tiles = (int)tiles;
  {
    const char* dbg = getenv("SMOT_TC_DEBUG");  // hex device pointer to 8 x u64
    a.dbg = dbg ? (unsigned long long*)strtoull(dbg, nullptr, 16) : nullptr;
  }
  a.partial = d->workspace ? (float*)((char*)d->workspace + SMOT_CONV_WS_COUNTER_BYTES) : nullptr;
  dim3 grid((unsigned)tiles, (unsigned)(d->Cout / BN), (unsigned)a.splits);
  // many short tiles: 2-stage rings let 4 CTAs share an SM, so one CTA's prologue / epilogue overlaps the
  // main loops of the others (same bytes in flight per SM as 2 CTAs x 4 stages)
  const bool shallow = fs ? fs == 2 : (tiles * (d->Cout / BN) >= 296 && all_chunks <= 36);
  // at most one CTA per SM: nothing else hides the ~1 us TMA round trip (the loop then advances `ring depth` chunks
  // per round trip), so use the whole shared memory for the ring: 8 x 24 KB, 6 x 32 KB, 4 x 48 KB
  const bool solo = (long long)grid.x * grid.y * grid.z <= 148;
  int stages = tc_stages(BN, solo, shallow, sliced ? all_chunks : a.chunks_per_split, fs);
  if (sliced && stages == 2) stages = BN == 128 ? 3 : 4;   // the 2-stage variants carry no slice code
#define SMOT_TC_LAUNCH(BN_, ST_) launch_tc<BN_, ST_>(tmA, tmB, a, grid, st)
  const int rc = SMOT_TC_DISPATCH(BN, stages, SMOT_TC_LAUNCH);
#undef SMOT_TC_LAUNCH
  if (rc != SMOT_OK || a.splits == 1 || a.cluster_reduce) return rc;
  const size_t total_out = (size_t)a.num_tiles * TC_BM * (d->Cout / 4);
  launch_pdl(splitk_reduce_kernel, dim3((unsigned)((total_out + 255) / 256)), dim3(256), 0, st, a);
  SMOT_CHECK_LAUNCH("smot_conv2d(split-K reduce)");
  return SMOT_OK;
}

}  // namespace smot

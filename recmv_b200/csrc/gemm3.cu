// Training-path GEMMs of the three MLPs on tcgen05 (sm_100a): backward-data  G_{l-1} = (G_l W_l) * act'(.)  and the
// weight gradient  dW_l = G_l^T X_l  (reduction over the samples), both in the engine's fp32-grade arithmetic:
// every fp32 operand is split a = hi + lo into two fp16 values and a product is 3 MMAs (hi*hi + lo*hi + hi*lo) with
// fp32 accumulation in TMEM.  Replaces what autograd runs for `loss.backward()` (train.py:325) through
// model/network.py:89-119 / Deformer.py:171-206 / RenderNet.py:59-96 -- cuBLAS SGEMMs plus separate softplus / ReLU
// backward kernels -- with ONE kernel template:
//
//      D[M x N] = A[M x K] . B[N x K]^T      (one 128 x 128 output tile per CTA, K in blocks of 64)
//
//  * operands stay in their natural fp32 row-major buffers in HBM; eight producer warps read them (coalesced),
//    scale by an exact power of two, split to fp16 hi/lo and write the UMMA K-major / 128-byte-swizzle tiles into a
//    3-stage shared-memory ring (64 KB per stage).  Two read patterns:
//      direct      element (row, k) at base[row * ld + k]   (K contiguous:  G as the A operand of backward-data)
//      transposed  element (row, k) at base[k * ld + row]   (row contiguous: W^T, and G^T / X^T of the weight gradient)
//    so no transposed copy of weights, activations or gradients ever exists in memory.
//  * one thread issues tcgen05.mma.cta_group::1.kind::f16 (M 128, N 128, K 16), 12 per K block; tcgen05.commit frees
//    the stage.  TWO 128-column accumulators in TMEM alternate every `chunk_kb` K blocks: the tensor core accumulates
//    with truncation (~2^-24 of the accumulator per MMA), so a reduction over 10^5 samples is cut
//    into chunks of 2048 whose partial sums are added in fp32 round-to-nearest by the epilogue while the next chunk's
//    MMAs run.
//  * four epilogue warps (one per TMEM lane quadrant): tcgen05.ld -> unscale -> activation derivative from the SAVED
//    layer input (softplus beta=100: 1 - exp(-100 a); ReLU: a > 0) -> fp32 store; for the weight gradient the tile of
//    dW itself (L2 resident, owned by this CTA: deterministic, no atomics) is the running sum.  The producers of the
//    weight gradient also accumulate the bias gradient (column sums of G) on the way.
// Every mbarrier wait is bounded (status code instead of a hung GPU), as in the forward engine.
#include "sdf_mlp.cuh"
#include "tc_common.cuh"

namespace recmv {
using namespace tc;

namespace {

constexpr int kBM = 128, kBN = 128, kBK = 64;
constexpr int kStages = 3;
constexpr uint32_t kPlane = 16384;                 // one [128 rows][64 fp16] tile
constexpr uint32_t kStageBytes = 4 * kPlane;       // A hi | A lo | B hi | B lo
constexpr int kProdWarps = 8, kEpiWarps = 4;
constexpr int kThreads = 32 * (2 + kProdWarps + kEpiWarps);   // warp 0 MMA issuer, warp 1 TMEM allocator
constexpr uint32_t kOffBar = kStages * kStageBytes;           // 196608
constexpr int kBarFull = 0, kBarEmpty = kStages, kBarAccFull = 2 * kStages, kBarAccEmpty = 2 * kStages + 2;
constexpr int kNumBars = 2 * kStages + 4;
constexpr uint32_t kOffMisc = kOffBar + kNumBars * 8;
constexpr uint32_t kSmemBytes = kOffMisc + 64 + 1024 /* alignment slack */;

enum { EPI_STORE = 0, EPI_SOFTPLUS100 = 1, EPI_RELU = 2, EPI_ACCUM = 3,          // backward-data (0-2), weight gradient (3)
       EPI_FWD_NONE = 4, EPI_FWD_SOFTPLUS100 = 5, EPI_FWD_RELU = 6 };                // forward layer: act(acc + bias)

struct G3Task {
  const float* A; long long lda; int a_transposed;
  const float* B; long long ldb; int b_transposed;  // transposed: element (n, k) at B[k * ldb + n]; direct: B[n * ldb + k]
  const float* bias;                                // forward layers: added after unscaling, before the activation
  float* D; long long ldd;                          // D[m * ldd + n], n < split
  float* D2; long long ldd2;                        // columns n >= split: D2[m * ldd2 + (n - split)], plain store
  const float* E; long long lde;                    // saved layer input for the activation derivative, E[m * lde + n]
  float* colsum;                                    // EPI_ACCUM: sum over k of A[k][m] (bias gradient), zero-initialised
  int M, N, split;
  long long K;
  int tiles_m, tiles_n, tile_base;
  float a_scale, b_scale, d_scale;
  int epi;
};
constexpr int kMaxTasks = 10;
struct G3Params {
  G3Task t[kMaxTasks];
  int ntasks;
  int chunk_kb;               // K blocks per accumulator chunk
  const float* dyn_scale;     // optional device scalar multiplied into a_scale (power of two)
  DevStatus* status;
};

__device__ __forceinline__ void umma_f16_1cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_1cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1) gemm3_kernel(const __grid_constant__ G3Params prm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kOffBar;
  volatile int* abort_flag = reinterpret_cast<volatile int*>(gbase + kOffMisc + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kOffMisc);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- which task / tile -----------------------------------------------------------------------------------------
  int ti = 0;
#pragma unroll 1
  for (int i = 1; i < prm.ntasks; ++i)
    if ((int)blockIdx.x >= prm.t[i].tile_base) ti = i;
  const G3Task& T = prm.t[ti];
  const int tile = (int)blockIdx.x - T.tile_base;
  const int tm = tile / T.tiles_n, tn = tile - tm * T.tiles_n;
  const int m0 = tm * kBM, n0 = tn * kBN;
  const long long nkb = (T.K + kBK - 1) / kBK;
  const int chunk_kb = prm.chunk_kb > 0 ? prm.chunk_kb : 32;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < kStages; ++s) { mbar_init(BAR(kBarFull + s), kProdWarps); mbar_init(BAR(kBarEmpty + s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(BAR(kBarAccFull + b), 1); mbar_init(BAR(kBarAccEmpty + b), kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================================== MMA issuer ========================================================
    if (lane == 0) {
      const uint32_t idesc = idesc_f16(kBM, kBN);
      for (long long kb = 0; kb < nkb; ++kb) {
        const int s = (int)(kb % kStages);
        const uint32_t par = (uint32_t)((kb / kStages) & 1);
        const long long chunk = kb / chunk_kb;
        const int buf = (int)(chunk & 1);
        const bool chunk_first = (kb % chunk_kb) == 0;
        if (chunk_first) {   // the accumulator must have been drained by the epilogue (two chunks ago)
          mbar_wait(BAR(kBarAccEmpty + buf), (uint32_t)(((chunk >> 1) & 1) ^ 1), abort_flag, prm.status, 2200 + buf);
          tc_fence_after();
        }
        mbar_wait(BAR(kBarFull + s), par, abort_flag, prm.status, 2100 + s);
        tc_fence_after();
        const uint32_t st = base + (uint32_t)s * kStageBytes;
        const uint64_t a_hi = smem_desc_sw128(st), a_lo = smem_desc_sw128(st + kPlane);
        const uint64_t b_hi = smem_desc_sw128(st + 2 * kPlane), b_lo = smem_desc_sw128(st + 3 * kPlane);
        const uint32_t dcol = tmem_base + (uint32_t)(buf * kBN);
        // correction products first (they are 2^-11 of the result: added while the K block's contribution is still
        // small they cost no accumulator precision), then hi*hi
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_1cta(dcol, a_lo + 2 * k, b_hi + 2 * k, idesc, (chunk_first && k == 0) ? 0u : 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_1cta(dcol, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_1cta(dcol, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
        umma_commit_1cta(BAR(kBarEmpty + s));
        if (kb == nkb - 1 || ((kb + 1) % chunk_kb) == 0) umma_commit_1cta(BAR(kBarAccFull + buf));
      }
    }
  } else if (warp >= 2 && warp < 2 + kProdWarps) {
    // =========================================== producers ==========================================================
    const int t = threadIdx.x - 64;            // 0 .. 255
    const float dyn = prm.dyn_scale ? __ldg(prm.dyn_scale) : 1.f;
    const float sa = T.a_scale * dyn, sb = T.b_scale;
    float colsum = 0.f;                        // EPI_ACCUM: sum over k of A[k][m0 + (t & 127)], unscaled
    // Task geometry of this thread (4 tasks of 8 K-consecutive elements per operand and K block):
    //   transposed read: row r = t & 127, chunks c = (t >> 7) + 2 i  -- lanes walk consecutive rows (coalesced)
    //   direct read    : chunk c = t & 7, rows r = (t >> 3) + 32 i   -- 8 lanes cover one row's 64 floats
    // ALL loads of a K block are issued before any of them is consumed, and the loads of K block kb + 1 are issued
    // right after K block kb has been written to shared memory: the global-load latency (~0.7 us, against 0.4 us of MMAs
    // per K block) overlaps the slot wait and the other warps' conversions instead of serialising 8 round trips.
    float va[4][8], vb[4][8];
    auto load_block = [&](long long kb) {
      const long long k0 = kb * kBK;
      if (T.a_transposed) {
        const int r = t & 127;
        const bool rok = m0 + r < T.M;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = (t >> 7) + 2 * i;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const long long k = k0 + c * 8 + j;
            va[i][j] = (rok && k < T.K) ? __ldg(T.A + k * T.lda + (m0 + r)) : 0.f;
          }
        }
      } else {
        const int c = t & 7;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = (t >> 3) + 32 * i;
          const long long k = k0 + c * 8;
          const bool rok = m0 + r < T.M;
          const float* src = T.A + (long long)(m0 + r) * T.lda + k;
          if (rok && k + 8 <= T.K && ((T.lda & 3) == 0)) {
            const float4 x0 = __ldg(reinterpret_cast<const float4*>(src)), x1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
            va[i][0] = x0.x; va[i][1] = x0.y; va[i][2] = x0.z; va[i][3] = x0.w;
            va[i][4] = x1.x; va[i][5] = x1.y; va[i][6] = x1.z; va[i][7] = x1.w;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) va[i][j] = (rok && k + j < T.K) ? __ldg(src + j) : 0.f;
          }
        }
      }
      if (T.b_transposed) {
        const int r = t & 127;
        const bool rok = n0 + r < T.N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = (t >> 7) + 2 * i;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const long long k = k0 + c * 8 + j;
            vb[i][j] = (rok && k < T.K) ? __ldg(T.B + k * T.ldb + (n0 + r)) : 0.f;
          }
        }
      } else {
        const int c = t & 7;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = (t >> 3) + 32 * i;
          const long long k = k0 + c * 8;
          const bool rok = n0 + r < T.N;
          const float* src = T.B + (long long)(n0 + r) * T.ldb + k;
          if (rok && k + 8 <= T.K && ((T.ldb & 3) == 0)) {
            const float4 x0 = __ldg(reinterpret_cast<const float4*>(src)), x1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
            vb[i][0] = x0.x; vb[i][1] = x0.y; vb[i][2] = x0.z; vb[i][3] = x0.w;
            vb[i][4] = x1.x; vb[i][5] = x1.y; vb[i][6] = x1.z; vb[i][7] = x1.w;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) vb[i][j] = (rok && k + j < T.K) ? __ldg(src + j) : 0.f;
          }
        }
      }
    };
    if (nkb > 0) load_block(0);
    for (long long kb = 0; kb < nkb; ++kb) {
      const int s = (int)(kb % kStages);
      const uint32_t par = (uint32_t)((kb / kStages) & 1);
      mbar_wait(BAR(kBarEmpty + s), par ^ 1u, abort_flag, prm.status, 2000 + s);
      const uint32_t st = base + (uint32_t)s * kStageBytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // ---- A ----
        {
          const int r = T.a_transposed ? (t & 127) : (t >> 3) + 32 * i;
          const int c = T.a_transposed ? (t >> 7) + 2 * i : (t & 7);
          if (T.colsum) {
#pragma unroll
            for (int j = 0; j < 8; ++j) colsum += va[i][j];
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) va[i][j] *= sa;
          range_check8(va[i], prm.status, 2300);
          uint4 hi, lo;
          split8(va[i], hi, lo);
          const uint32_t off = sw128_offset(r, c);
          st_shared_v4(st + off, hi);
          st_shared_v4(st + kPlane + off, lo);
        }
        // ---- B ----
        {
          const int r = T.b_transposed ? (t & 127) : (t >> 3) + 32 * i;
          const int c = T.b_transposed ? (t >> 7) + 2 * i : (t & 7);
#pragma unroll
          for (int j = 0; j < 8; ++j) vb[i][j] *= sb;
          range_check8(vb[i], prm.status, 2302);
          uint4 hi, lo;
          split8(vb[i], hi, lo);
          const uint32_t off = sw128_offset(r, c);
          st_shared_v4(st + 2 * kPlane + off, hi);
          st_shared_v4(st + 3 * kPlane + off, lo);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_local(BAR(kBarFull + s));
      if (kb + 1 < nkb) load_block(kb + 1);
    }
    // bias gradient: two producer threads hold the two halves of every row's column sum (commutative: deterministic)
    if (T.colsum && tn == 0 && T.a_transposed && m0 + (t & 127) < T.M) atomicAdd(T.colsum + m0 + (t & 127), colsum);
    // the last `kStages` stage releases are asynchronous tcgen05.commit arrivals nobody waits for any more: collect
    // them before this CTA may exit (they would otherwise land in the next CTA's shared memory)
    if (t == 0) {
      for (long long kb = nkb; kb < nkb + kStages; ++kb) {
        if (kb - kStages < 0) continue;
        const int s = (int)(kb % kStages);
        mbar_wait(BAR(kBarEmpty + s), (uint32_t)(((kb / kStages) & 1) ^ 1), abort_flag, prm.status, 2050 + s);
      }
    }
  } else if (warp >= 2 + kProdWarps) {
    // =========================================== epilogue ===========================================================
    const int q = warp & 3;                    // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const float dyn = prm.dyn_scale ? __ldg(prm.dyn_scale) : 1.f;
    const float unscale0 = T.d_scale / (T.a_scale * dyn * T.b_scale);
    const long long nchunks = (nkb + chunk_kb - 1) / chunk_kb;
    for (long long ch = 0; ch < nchunks; ++ch) {
      const int buf = (int)(ch & 1);
      // compensation of the tensor core's truncating fp32 accumulation: every 64-wide K block adds 4 hi*hi MMAs into
      // the full-size accumulator, each dropping a fraction of 2^-24 of it on average (acc_trunc_gain, tc_common.cuh)
      const long long kbs = (ch + 1) * chunk_kb <= nkb ? chunk_kb : nkb - ch * chunk_kb;
      const float unscale = unscale0 * acc_trunc_gain((int)kbs);
      mbar_wait(BAR(kBarAccFull + buf), (uint32_t)((ch >> 1) & 1), abort_flag, prm.status, 2400 + buf);
      tc_fence_after();
#pragma unroll 1
      for (int cg = 0; cg < kBN / 32; ++cg) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kBN + cg * 32), r);
        tmem_ld_wait();
        const int nb = n0 + cg * 32;
        if (m < T.M && nb < T.N) {
          if (T.epi == EPI_ACCUM) {
            float* d = T.D + (long long)m * T.ldd + nb;
            const bool vec = ((T.ldd & 3) == 0) && nb + 32 <= T.N;
            if (vec) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 o = make_float4(__uint_as_float(r[4 * j]) * unscale, __uint_as_float(r[4 * j + 1]) * unscale,
                                       __uint_as_float(r[4 * j + 2]) * unscale, __uint_as_float(r[4 * j + 3]) * unscale);
                if (ch > 0) {
                  const float4 p = *reinterpret_cast<const float4*>(d + 4 * j);
                  o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
                }
                *reinterpret_cast<float4*>(d + 4 * j) = o;
              }
            } else {
#pragma unroll 1
              for (int j = 0; j < 32; ++j)
                if (nb + j < T.N) {
                  const float o = __uint_as_float(r[j]) * unscale;
                  d[j] = ch > 0 ? d[j] + o : o;
                }
            }
          } else {
            // backward-data / forward layer: one chunk (K = layer width).  Columns < split go to D (backward-data: times
            // the activation derivative from the saved input E), columns >= split to D2.
            const bool fwd = T.epi >= EPI_FWD_NONE;
#pragma unroll 1
            for (int j4 = 0; j4 < 8; ++j4) {
              const int n = nb + 4 * j4;
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = __uint_as_float(r[4 * j4 + e]) * unscale;
              if (fwd) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (n + e >= T.N) continue;
                  float v = o[e] + (T.bias ? __ldg(T.bias + n + e) : 0.f);
                  if (T.epi == EPI_FWD_SOFTPLUS100) v = softplus100(v);
                  else if (T.epi == EPI_FWD_RELU) v = fmaxf(v, 0.f);
                  o[e] = v;
                }
              }
              const bool vec = n + 4 <= T.split && (T.ldd & 3) == 0 && (fwd || T.E == nullptr || (T.lde & 3) == 0);
              if (vec) {
                if (T.epi == EPI_SOFTPLUS100) {
                  const float4 a = __ldg(reinterpret_cast<const float4*>(T.E + (long long)m * T.lde + n));
                  // d softplus_100(z) / dz = sigmoid(100 z) = 1 - exp(-100 a),  a = softplus_100(z)  (exact also on
                  // torch's threshold branch a = z > 0.2, where it is 1 - 2e-9)
                  o[0] *= 1.f - __expf(-100.f * a.x); o[1] *= 1.f - __expf(-100.f * a.y);
                  o[2] *= 1.f - __expf(-100.f * a.z); o[3] *= 1.f - __expf(-100.f * a.w);
                } else if (T.epi == EPI_RELU) {
                  const float4 a = __ldg(reinterpret_cast<const float4*>(T.E + (long long)m * T.lde + n));
                  o[0] = a.x > 0.f ? o[0] : 0.f; o[1] = a.y > 0.f ? o[1] : 0.f;
                  o[2] = a.z > 0.f ? o[2] : 0.f; o[3] = a.w > 0.f ? o[3] : 0.f;
                }
                *reinterpret_cast<float4*>(T.D + (long long)m * T.ldd + n) = make_float4(o[0], o[1], o[2], o[3]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int nn = n + e;
                  if (nn >= T.N) continue;
                  if (nn < T.split) {
                    float v = o[e];
                    if (T.epi == EPI_SOFTPLUS100) v *= 1.f - __expf(-100.f * __ldg(T.E + (long long)m * T.lde + nn));
                    else if (T.epi == EPI_RELU) v = __ldg(T.E + (long long)m * T.lde + nn) > 0.f ? v : 0.f;
                    T.D[(long long)m * T.ldd + nn] = v;
                  } else if (T.D2) {
                    T.D2[(long long)m * T.ldd2 + (nn - T.split)] = o[e];
                  }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_local(BAR(kBarAccEmpty + buf));
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

namespace {
int g3_launch(G3Params& prm, int total_tiles, cudaStream_t st) {
  static bool done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!done[dev & 15]) {
    cudaError_t e = cudaFuncSetAttribute(gemm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    done[dev & 15] = true;
  }
  void* sd = nullptr;
  int s0 = device_status_record(&sd);
  if (s0) return s0;
  if (((DevStatus*)sd)->code != 0) return RECMV_E_DEVICE;
  prm.status = (DevStatus*)sd;
  if (total_tiles <= 0) return RECMV_OK;
  gemm3_kernel<<<total_tiles, kThreads, kSmemBytes, st>>>(prm);
  return launch_status();
}

void g3_set_tiles(G3Task& t, int& next_tile) {
  t.tiles_m = (t.M + kBM - 1) / kBM;
  t.tiles_n = (t.N + kBN - 1) / kBN;
  t.tile_base = next_tile;
  next_tile += t.tiles_m * t.tiles_n;
}

// d PE / d x applied to a cotangent: dx_j = g[j] + sum_k f_k (w_sin_k cos(f_k x_j) g[3+6k+j] - w_cos_k sin(f_k x_j) g[3+6k+3+j])
// (model/Embedder.py:43-50 order; `bands` = 6 for points, 4 for view directions).  g2: optional second cotangent of the
// same encoding (the SDF network's skip connection), added before the contraction.
__global__ void __launch_bounds__(256) pe_backward_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                          long long ldg, const float* __restrict__ g2, long long ldg2,
                                                          PeWeights pw, int bands, float* __restrict__ dx, long long P,
                                                          int accumulate) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= 3 * P) return;
  const long long p = i / 3;
  const int j = (int)(i - 3 * p);
  const float* gp = g + p * ldg;
  const float* gq = g2 ? g2 + p * ldg2 : nullptr;
  auto G = [&](int e) { return gp[e] + (gq ? gq[e] : 0.f); };
  const float xj = x[i];
  float acc = G(j);
  float f = 1.f;
  for (int k = 0; k < bands; ++k) {
    float sn, cs;
    sincosf(xj * f, &sn, &cs);
    acc += f * (pw.w[2 * k] * cs * G(3 + 6 * k + j) - pw.w[2 * k + 1] * sn * G(3 + 6 * k + 3 + j));
    f *= 2.f;
  }
  dx[i] = accumulate ? dx[i] + acc : acc;
}
// positional encoding written as a saved layer input: out[p * ld + e], e < 3 + 6 * bands (model/Embedder.py:43-50);
// out2 (optional) receives the same row (the SDF network's skip input)
__global__ void __launch_bounds__(256) pe_forward_kernel(const float* __restrict__ x, PeWeights pw, int bands,
                                                         float* __restrict__ out, long long ld, float* __restrict__ out2,
                                                         long long ld2, long long P) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= P) return;
  float pe[39];
  positional_encode(x[3 * p], x[3 * p + 1], x[3 * p + 2], pw.w, pe);
  const int n = 3 + 6 * bands;
  for (int e = 0; e < n; ++e) {
    out[p * ld + e] = pe[e];
    if (out2) out2[p * ld2 + e] = pe[e];
  }
}
}  // namespace
}  // namespace recmv

using namespace recmv;

extern "C" int recmv_mlp_fwd_layer(const float* X, int64_t ldx, const float* W, const float* bias, int out_dim, int in_dim,
                                   int act, float pre_scale, int split, float* Y, int64_t ldy, float* Y2, int64_t ldy2,
                                   int64_t P, recmv_stream_t stream) {
  if (P < 0 || out_dim <= 0 || in_dim <= 0) return RECMV_E_SHAPE;
  if (P == 0) return RECMV_OK;
  if (!X || !W || !Y) return RECMV_E_NULL;
  if (act < 0 || act > 2) return RECMV_E_DTYPE;
  if (P > (int64_t)1 << 30) return RECMV_E_RANGE;
  G3Params prm = {};
  G3Task& t = prm.t[0];
  t.A = X; t.lda = ldx; t.a_transposed = 0;
  t.B = W; t.ldb = in_dim; t.b_transposed = 0;   // element (n = output row, k = input column) at W[n * in_dim + k]
  t.bias = bias;
  t.D = Y; t.ldd = ldy; t.D2 = Y2; t.ldd2 = ldy2;
  t.M = (int)P; t.N = out_dim; t.K = in_dim;
  t.split = (split > 0 && split < out_dim) ? split : out_dim;
  t.a_scale = kActScale; t.b_scale = kWgtScale; t.d_scale = pre_scale;
  t.epi = EPI_FWD_NONE + act;
  int tiles = 0;
  g3_set_tiles(t, tiles);
  prm.ntasks = 1; prm.chunk_kb = 1 << 20; prm.dyn_scale = nullptr;
  return g3_launch(prm, tiles, (cudaStream_t)stream);
}

extern "C" int recmv_pe_forward(const float* x, const float* pe_w, int bands, float* out, int64_t ld, float* out2,
                                int64_t ld2, int64_t P, recmv_stream_t stream) {
  if (P < 0 || bands < 0 || bands > 6) return RECMV_E_SHAPE;
  if (P == 0) return RECMV_OK;
  if (!x || !pe_w || !out) return RECMV_E_NULL;
  PeWeights pw;
  for (int i = 0; i < 12; ++i) pw.w[i] = i < 2 * bands ? pe_w[i] : 0.f;
  pe_forward_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, pw, bands, out, ld, out2, ld2, P);
  return launch_status();
}

// --------------------------------------------------------------------------------------------------------------------
// C ABI (include/recmv_b200.h): generic layer description, so the three networks share the two entry points
// --------------------------------------------------------------------------------------------------------------------
extern "C" int recmv_mlp_bwd_data_layer(const float* G, int64_t ldg, const float* W, int out_dim, int in_dim,
                                        const float* saved_input, int64_t lds, int act, int split, float* G_prev,
                                        int64_t ldgp, float* D2, int64_t ldd2, float out_scale, const float* dyn_scale,
                                        int64_t P, recmv_stream_t stream) {
  if (P < 0 || out_dim <= 0 || in_dim <= 0) return RECMV_E_SHAPE;
  if (P == 0) return RECMV_OK;
  if (!G || !W || !G_prev) return RECMV_E_NULL;
  if (act != EPI_STORE && act != EPI_SOFTPLUS100 && act != EPI_RELU) return RECMV_E_DTYPE;
  if (act != EPI_STORE && !saved_input) return RECMV_E_NULL;
  if (P > (int64_t)1 << 30) return RECMV_E_RANGE;
  G3Params prm = {};
  G3Task& t = prm.t[0];
  t.A = G; t.lda = ldg; t.a_transposed = 0;
  t.B = W; t.ldb = in_dim; t.b_transposed = 1; // element (n = input column, k = output row) at W[k * in_dim + n]
  t.D = G_prev; t.ldd = ldgp; t.D2 = D2; t.ldd2 = ldd2;
  t.E = saved_input; t.lde = lds;
  t.M = (int)P; t.N = in_dim; t.K = out_dim;
  t.split = (split > 0 && split < in_dim) ? split : in_dim;
  t.a_scale = kActScale; t.b_scale = kWgtScale; t.d_scale = out_scale;
  t.epi = act;
  int tiles = 0;
  g3_set_tiles(t, tiles);
  prm.ntasks = 1; prm.chunk_kb = 1 << 20; prm.dyn_scale = dyn_scale;
  return g3_launch(prm, tiles, (cudaStream_t)stream);
}

extern "C" int recmv_mlp_bwd_weight(int num_layers, const float* const* G, const int64_t* ldg, const float* const* X,
                                    const int64_t* ldx, const int* out_dim, const int* in_dim, float* const* dW,
                                    float* const* db, const float* out_scale, const float* dyn_scale, int64_t P,
                                    recmv_stream_t stream) {
  if (num_layers <= 0 || num_layers > kMaxTasks || P < 0) return RECMV_E_SHAPE;
  if (P == 0) return RECMV_OK;
  if (!G || !X || !dW || !out_dim || !in_dim || !ldg || !ldx) return RECMV_E_NULL;
  G3Params prm = {};
  int tiles = 0;
  for (int l = 0; l < num_layers; ++l) {
    if (!G[l] || !X[l] || !dW[l]) return RECMV_E_NULL;
    G3Task& t = prm.t[l];
    t.A = G[l]; t.lda = ldg[l]; t.a_transposed = 1;     // element (m = output row, k = sample) at G[k * ldg + m]
    t.B = X[l]; t.ldb = ldx[l]; t.b_transposed = 1;     // element (n = input column, k = sample) at X[k * ldx + n]
    t.D = dW[l]; t.ldd = in_dim[l]; t.D2 = nullptr; t.E = nullptr;
    t.colsum = db ? db[l] : nullptr;
    t.M = out_dim[l]; t.N = in_dim[l]; t.K = P; t.split = in_dim[l];
    t.a_scale = kActScale; t.b_scale = kActScale; t.d_scale = out_scale ? out_scale[l] : 1.f;
    t.epi = EPI_ACCUM;
    g3_set_tiles(t, tiles);
  }
  prm.ntasks = num_layers; prm.chunk_kb = 32; prm.dyn_scale = dyn_scale;   // 32 K blocks = 2048 samples per accumulator chunk
  return g3_launch(prm, tiles, (cudaStream_t)stream);
}

extern "C" int recmv_pe_backward(const float* x, const float* g, int64_t ldg, const float* g2, int64_t ldg2,
                                 const float* pe_w, int bands, float* dx, int accumulate, int64_t P,
                                 recmv_stream_t stream) {
  if (P < 0 || bands < 0 || bands > 6) return RECMV_E_SHAPE;
  if (P == 0) return RECMV_OK;
  if (!x || !g || !pe_w || !dx) return RECMV_E_NULL;
  PeWeights pw;
  for (int i = 0; i < 12; ++i) pw.w[i] = i < 2 * bands ? pe_w[i] : 0.f;
  const long long n = 3 * P;
  pe_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, g, ldg, g2, ldg2, pw, bands, dx, P,
                                                                                   accumulate);
  return launch_status();
}

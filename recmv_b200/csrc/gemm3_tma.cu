// Training-path layer GEMMs, second generation: forward layers and backward-data with BOTH operands already split into
// fp16 hi / lo planes in HBM and fed by TMA -- no conversion in the main loop.
//
// gemm3.cu converts fp32 operands inside its main loop (eight producer warps: load -> scale -> split -> st.shared); at
// ~6 us per 64-wide K block against 0.4 us of MMAs that chain bounds it (profiles/r02_notes.md).  Here every tensor a
// later GEMM consumes is WRITTEN in consumable form by the kernel that produces it:
//   * a layer's output  Y  -> fp32 [P][ld] (activation derivative, weight gradient) AND two fp16 planes [P][ldp]
//     hi = fp16(64 y), lo = fp16(64 y - hi)  (K-major rows: exactly what the next layer's A operand is);
//   * a cotangent       G  -> fp32 AND planes of 64 * dyn * g  (dyn = the call's power-of-two gradient scale, device scalar);
//   * weights           W  -> planes of 1024 w, once per step: [out][in] for the forward, [in][out] for backward-data
//     (recmv_split_planes, optionally transposing).
// Main loop = the inference engine's recipe on one CTA: a TMA thread streams [128 rows x 64 k] SW128 boxes of the four
// planes through a 3-stage ring (complete_tx on the stage's mbarrier; out-of-range rows / columns are zero-filled by TMA,
// so ragged M, N, K need no masking), one thread issues 12 tcgen05 MMAs per K block (lo*hi, hi*lo, hi*hi), tcgen05.commit
// frees the stage.  PERSISTENT: a CTA walks tiles blockIdx.x, + gridDim.x, ...; the two 128-column TMEM accumulators
// alternate per tile, so the four epilogue warps (bias / activation or activation derivative, fp32 store, hi / lo split and
// plane store) work on tile i while tile i + 1's loads and MMAs run.
#include "sdf_mlp.cuh"
#include "tc_common.cuh"

namespace recmv {
using namespace tc;

namespace {

constexpr int kBM = 128, kBN = 128, kBK = 64;
constexpr int kStages = 3;
constexpr uint32_t kPlane = 16384;                 // one [128 rows][64 fp16] tile
constexpr uint32_t kStageBytes = 4 * kPlane;       // A hi | A lo | B hi | B lo
constexpr int kEpiWarps = 16;                      // 4 per TMEM lane quadrant, one 32-column group of the 128-column tile each:
                                                   // with 4 warps the epilogue (activation, fp32 store, hi / lo split, plane
                                                   // store for 128 values per thread) took ~30 us per tile against ~8 us of
                                                   // main loop and bounded the kernel (r02_launches_train_131k_planes.csv)
constexpr int kThreads = 32 * (2 + kEpiWarps);     // warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator
constexpr uint32_t kOffBar = kStages * kStageBytes;
constexpr int kBarFull = 0, kBarEmpty = kStages, kBarAccFull = 2 * kStages, kBarAccEmpty = 2 * kStages + 2;
constexpr int kNumBars = 2 * kStages + 4;
constexpr uint32_t kOffMisc = kOffBar + kNumBars * 8;
constexpr uint32_t kSmemBytes = kOffMisc + 64 + 1024 /* alignment slack */;

enum { TEPI_BWD_NONE = 0, TEPI_BWD_SOFTPLUS100 = 1, TEPI_BWD_RELU = 2, TEPI_FWD_NONE = 4, TEPI_FWD_SOFTPLUS100 = 5,
       TEPI_FWD_RELU = 6 };

struct GtParams {
  float* D; long long ldd;                 // fp32 output, columns n < split
  float* D2; long long ldd2;               // fp32 output, columns n >= split (no activation derivative), may be NULL
  __half* DH; __half* DL; long long ldp;   // optional planes of the < split columns: split(plane_scale * value)
  const float* E; long long lde;           // saved layer input (backward-data activation derivative)
  const float* bias;                       // forward
  const float* dyn_scale;                  // backward-data: device scalar in the A planes' scale
  float unscale;                           // d_scale / (a_scale * b_scale), before dyn
  float plane_scale;                       // 64 (x dyn for cotangents)
  int M, N, K, split, tiles_m, tiles_n, epi, dyn_in_planes;
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
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// softplus(beta = 100, threshold 20) in the inference engine's form (sdf_mlp_tc.cu softplus100_scaled): two MUFU ops;
// absolute error <= ~2e-9 on activations of scale 0.05 (lg2.approx: 2^-22 absolute near 1)
__device__ __forceinline__ float softplus100_fast(float z) {
  constexpr float kUThr = 20.f * 1.4426950408889634f;
  const float u = z * kSoftplusLog2Scale;
  const float y = lg2_approx(1.f + ex2_approx(fminf(u, kUThr))) * 0.0069314718055994531f;
  return u > kUThr ? z : y;
}
// 256-bit global accesses (sm_100: LDG / STG .256).  A thread of the epilogue owns one output ROW, so a 128-bit store covers
// half a 32-byte sector: the L2 then merges two partial writes per sector and, with ECC, fills the sector from DRAM first
// (ncu: 33.5 M sector writes and 134 MB of DRAM reads beyond the operands per 131 k x 512 layer).  32 bytes per thread and
// instruction = whole sectors.
// .cs (evict-first): the fp32 copy of a layer's output is not read again before the backward pass -- it should not push the
// operand planes (re-read by the other column tiles, and by the next layer) out of L2
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void st_global_v8(__half* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]),
               "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
// 2-D tile load into this CTA's shared memory; completion bytes on a local mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
gemm3_tma_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, const GtParams prm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kOffBar;
  volatile int* abort_flag = reinterpret_cast<volatile int*>(gbase + kOffMisc + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kOffMisc);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = prm.tiles_m * prm.tiles_n;
  const int nkb = (prm.K + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < kStages; ++s) { mbar_init(BAR(kBarFull + s), 1); mbar_init(BAR(kBarEmpty + s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(BAR(kBarAccFull + b), 1); mbar_init(BAR(kBarAccEmpty + b), kEpiWarps); }
    fence_mbar_init();
    prefetch_tmap(&tm_a_hi); prefetch_tmap(&tm_a_lo); prefetch_tmap(&tm_b_hi); prefetch_tmap(&tm_b_lo);
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
    // =========================================== TMA producer =======================================================
    if (lane == 0) {
      long long kbg = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tm = tile / prm.tiles_n, tn = tile - tm * prm.tiles_n;
        for (int kb = 0; kb < nkb; ++kb, ++kbg) {
          const int s = (int)(kbg % kStages);
          const uint32_t par = (uint32_t)((kbg / kStages) & 1);
          mbar_wait(BAR(kBarEmpty + s), par ^ 1u, abort_flag, prm.status, 3000 + s);
          const uint32_t st = base + (uint32_t)s * kStageBytes;
          mbar_expect_tx_local(BAR(kBarFull + s), kStageBytes);
          tma_load_2d(st, &tm_a_hi, BAR(kBarFull + s), kb * kBK, tm * kBM);
          tma_load_2d(st + kPlane, &tm_a_lo, BAR(kBarFull + s), kb * kBK, tm * kBM);
          tma_load_2d(st + 2 * kPlane, &tm_b_hi, BAR(kBarFull + s), kb * kBK, tn * kBN);
          tma_load_2d(st + 3 * kPlane, &tm_b_lo, BAR(kBarFull + s), kb * kBK, tn * kBN);
        }
      }
      // collect the last stage releases (asynchronous tcgen05.commit arrivals) before the CTA may exit
      for (long long k2 = kbg; k2 < kbg + kStages; ++k2) {
        if (k2 - kStages < 0) continue;
        const int s = (int)(k2 % kStages);
        mbar_wait(BAR(kBarEmpty + s), (uint32_t)(((k2 / kStages) & 1) ^ 1), abort_flag, prm.status, 3050 + s);
      }
    }
  } else if (warp == 1) {
    // =========================================== MMA issuer ========================================================
    if (lane == 0) {
      const uint32_t idesc = idesc_f16(kBM, kBN);
      long long kbg = 0, item = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++item) {
        const int buf = (int)(item & 1);
        mbar_wait(BAR(kBarAccEmpty + buf), (uint32_t)(((item >> 1) & 1) ^ 1), abort_flag, prm.status, 3200 + buf);
        tc_fence_after();
        const uint32_t dcol = tmem_base + (uint32_t)(buf * kBN);
        for (int kb = 0; kb < nkb; ++kb, ++kbg) {
          const int s = (int)(kbg % kStages);
          mbar_wait(BAR(kBarFull + s), (uint32_t)((kbg / kStages) & 1), abort_flag, prm.status, 3100 + s);
          tc_fence_after();
          const uint32_t st = base + (uint32_t)s * kStageBytes;
          const uint64_t a_hi = smem_desc_sw128(st), a_lo = smem_desc_sw128(st + kPlane);
          const uint64_t b_hi = smem_desc_sw128(st + 2 * kPlane), b_lo = smem_desc_sw128(st + 3 * kPlane);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_1cta(dcol, a_lo + 2 * k, b_hi + 2 * k, idesc, (kb == 0 && k == 0) ? 0u : 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_1cta(dcol, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_1cta(dcol, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
          umma_commit_1cta(BAR(kBarEmpty + s));
        }
        umma_commit_1cta(BAR(kBarAccFull + buf));
      }
    }
  } else {
    // =========================================== epilogue ===========================================================
    const int q = warp & 3;                    // TMEM lane quadrant of this warp
    const int cg = (warp - 2) >> 2;            // its 32-column group of the tile
    const int row = q * 32 + lane;
    const float dyn = prm.dyn_scale ? __ldg(prm.dyn_scale) : 1.f;
    // truncating-accumulation compensation (acc_trunc_gain, tc_common.cuh)
    const float unscale = prm.unscale / dyn * acc_trunc_gain(nkb);
    const float pscale = prm.plane_scale * (prm.dyn_in_planes ? dyn : 1.f);
    const bool fwd = prm.epi >= TEPI_FWD_NONE;
    long long item = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++item) {
      const int tm = tile / prm.tiles_n, tn = tile - tm * prm.tiles_n;
      const int m = tm * kBM + row, n0 = tn * kBN;
      const int buf = (int)(item & 1);
      mbar_wait(BAR(kBarAccFull + buf), (uint32_t)((item >> 1) & 1), abort_flag, prm.status, 3400 + buf);
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kBN + cg * 32), r);
        const int nb = n0 + cg * 32;
        // Fast path (every full 32-column group of an aligned layer, i.e. almost all of the work): the group's bias /
        // saved-input values are fetched as eight float4 while the TMEM load is in flight, the 32 outputs are produced
        // without a loop-carried dependency, and softplus uses the engine's ex2 / lg2 form.  (ncu on the previous epilogue:
        // 84 thread instructions per output -- log1pf(expf()) and a bias load waited for per element -- tensor pipe 22 %.)
        // (warp-uniform: tcgen05.wait::ld is .sync.aligned, the row bound m < M only predicates the global accesses)
        const bool fast = nb + 32 <= prm.split && nb + 32 <= prm.N && (prm.ldd & 3) == 0 &&
                          (prm.DH == nullptr || (prm.ldp & 7) == 0) &&
                          (fwd ? true : (prm.epi == TEPI_BWD_NONE || (prm.lde & 3) == 0));
        // 256-bit accesses need 32-byte aligned rows (warp-uniform: pointers and strides only)
        const bool wide = (prm.ldd & 7) == 0 && (reinterpret_cast<uintptr_t>(prm.D) & 31) == 0 &&
                          (fwd ? (reinterpret_cast<uintptr_t>(prm.bias) & 31) == 0
                               : (prm.epi == TEPI_BWD_NONE || ((prm.lde & 7) == 0 && (reinterpret_cast<uintptr_t>(prm.E) & 31) == 0)));
        const bool wide_p = prm.DH != nullptr && (prm.ldp & 15) == 0 && (reinterpret_cast<uintptr_t>(prm.DH) & 31) == 0 &&
                            (reinterpret_cast<uintptr_t>(prm.DL) & 31) == 0;
        if (fast) {
          float ex[32];
          const bool has_ex = fwd ? prm.bias != nullptr : prm.epi != TEPI_BWD_NONE;
          const bool row_ok = m < prm.M;
          if (has_ex) {
            const float* src = fwd ? prm.bias + nb : prm.E + (long long)(row_ok ? m : 0) * prm.lde + nb;
            if (wide) {
#pragma unroll
              for (int j = 0; j < 4; ++j) ld_global_nc_v8(src + 8 * j, ex + 8 * j);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(src) + j);
                ex[4 * j] = t.x; ex[4 * j + 1] = t.y; ex[4 * j + 2] = t.z; ex[4 * j + 3] = t.w;
              }
            }
          }
          tmem_ld_wait();
          float o[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float v = __uint_as_float(r[e]) * unscale;
            if (fwd) {
              if (has_ex) v += ex[e];
              if (prm.epi == TEPI_FWD_SOFTPLUS100) v = softplus100_fast(v);
              else if (prm.epi == TEPI_FWD_RELU) v = fmaxf(v, 0.f);
            } else if (prm.epi == TEPI_BWD_SOFTPLUS100) {
              v *= 1.f - ex2_approx(-kSoftplusLog2Scale * ex[e]);
            } else if (prm.epi == TEPI_BWD_RELU) {
              v = ex[e] > 0.f ? v : 0.f;
            }
            o[e] = v;
          }
          if (row_ok) {
            float* d = prm.D + (long long)m * prm.ldd + nb;
            if (wide) {
#pragma unroll
              for (int j = 0; j < 4; ++j) st_global_v8(d + 8 * j, o + 8 * j);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(d)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
          }
          if (prm.DH && row_ok) {
            __half* dh = prm.DH + (long long)m * prm.ldp + nb;
            __half* dl = prm.DL + (long long)m * prm.ldp + nb;
#pragma unroll
            for (int j16 = 0; j16 < 2; ++j16) {
              uint4 hi[2], lo[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = o[16 * j16 + 8 * h + e] * pscale;
                range_check8(v, prm.status, 3500);
                split8(v, hi[h], lo[h]);
              }
              if (wide_p) {
                st_global_v8(dh + 16 * j16, hi[0], hi[1]);
                st_global_v8(dl + 16 * j16, lo[0], lo[1]);
              } else {
                reinterpret_cast<uint4*>(dh + 16 * j16)[0] = hi[0]; reinterpret_cast<uint4*>(dh + 16 * j16)[1] = hi[1];
                reinterpret_cast<uint4*>(dl + 16 * j16)[0] = lo[0]; reinterpret_cast<uint4*>(dl + 16 * j16)[1] = lo[1];
              }
            }
          }
        } else {
        tmem_ld_wait();
#pragma unroll 1
        for (int j8 = 0; j8 < 4 && m < prm.M; ++j8) {
          const int n = nb + 8 * j8;
          if (n >= prm.N) break;
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(r[8 * j8 + e]) * unscale;
          const bool full = n + 8 <= prm.split;                       // the 8 columns are all "activation" columns
          if (fwd) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (n + e >= prm.N) { o[e] = 0.f; continue; }
              float v = o[e] + (prm.bias ? __ldg(prm.bias + n + e) : 0.f);
              if (prm.epi == TEPI_FWD_SOFTPLUS100) v = softplus100(v);
              else if (prm.epi == TEPI_FWD_RELU) v = fmaxf(v, 0.f);
              o[e] = v;
            }
          } else if (prm.epi != TEPI_BWD_NONE) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (n + e >= prm.split) continue;
              const float a = __ldg(prm.E + (long long)m * prm.lde + n + e);
              o[e] = prm.epi == TEPI_BWD_SOFTPLUS100 ? o[e] * (1.f - __expf(-100.f * a)) : (a > 0.f ? o[e] : 0.f);
            }
          }
          if (full && (prm.ldd & 3) == 0) {
            float* d = prm.D + (long long)m * prm.ldd + n;
            *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(d + 4) = make_float4(o[4], o[5], o[6], o[7]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int nn = n + e;
              if (nn >= prm.N) continue;
              if (nn < prm.split) prm.D[(long long)m * prm.ldd + nn] = o[e];
              else if (prm.D2) prm.D2[(long long)m * prm.ldd2 + (nn - prm.split)] = o[e];
            }
          }
          if (prm.DH) {   // the same values as the next GEMM's A operand: fp16 hi / lo of plane_scale * value
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (n + e < prm.split) ? o[e] * pscale : 0.f;
            range_check8(v, prm.status, 3500);
            uint4 hi, lo;
            split8(v, hi, lo);
            if (full && (prm.ldp & 7) == 0) {
              *reinterpret_cast<uint4*>(prm.DH + (long long)m * prm.ldp + n) = hi;
              *reinterpret_cast<uint4*>(prm.DL + (long long)m * prm.ldp + n) = lo;
            } else {
              const __half* hh = reinterpret_cast<const __half*>(&hi);
              const __half* ll = reinterpret_cast<const __half*>(&lo);
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (n + e < prm.split) { prm.DH[(long long)m * prm.ldp + n + e] = hh[e]; prm.DL[(long long)m * prm.ldp + n + e] = ll[e]; }
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

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient on the SAME planes:  dW[m][n] = sum_p G[p][m] X[p][n]  (m = output row, n = input column, p = sample).
// Both operands are the row-major planes the layer GEMMs already wrote ([P][ld], the reduction index p is the ROW), i.e.
// they are MN-major for this product.  tcgen05 reads MN-major operands natively (instruction-descriptor bits 15 / 16;
// canonical SW128 layout ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) in fp16 elements: a [8 k-rows][64 mn] swizzle atom of
// 1024 B, atoms stacked along k every SBO = 1024 B and along mn every LBO), and that layout is exactly what TMA writes for
// a {64 columns, 64 rows} SW128 box: a 128-wide operand tile is two such boxes 8192 B apart (LBO), one MMA (K = 16
// samples) advances the start address by 16 rows = 2048 B.  So the weight gradient needs NO transposed copy of anything.
// Work split: the output is at most 4 x 4 tiles, so the sample range is split over CTAs (grid = tiles x splits, one wave);
// inside a CTA the accumulation runs in chunks of 32 K blocks alternating between two TMEM accumulators, and the epilogue
// warps add each finished chunk into fp32 registers (a 10^5-sample reduction in one truncating accumulator would be cut
// short, gemm3.cu); partial tiles go to a workspace and are summed in split order by a second launch (deterministic).
struct WgParams {
  float* ws;                 // [splits][tiles_m * 128][tiles_n * 128] raw partial sums (accumulator units, gain applied)
  float* ws_db;              // [splits][tiles_m * 128] raw partial row sums of A (the bias gradient), or NULL
  long long nkb;             // ceil(P / 64)
  int M, N, tiles_m, tiles_n, splits, chunk_kb;
  DevStatus* status;
};

// Bias gradient for free: db[m] = sum_p G[p][m] is one more product of the A tile, with a [16 x 64] tile of ones as a K-major B
// operand (2 KB of shared memory, written once per CTA): 8 extra N = 16 MMAs per K block in the CTAs of the first column of
// tiles, accumulated in 16 more TMEM columns per buffer and chunk-summed by the same epilogue warps.
constexpr uint32_t kOffOnes = (kOffMisc + 64 + 1023u) & ~1023u;
constexpr uint32_t kSmemBytesWg = kOffOnes + 2048 + 1024 /* alignment slack */;
constexpr uint32_t kDbCol = 2 * kBN;               // TMEM columns [256, 288): two 16-column accumulators

__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(kThreads, 1)
gemm3_wgrad_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, const WgParams prm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kOffBar;
  volatile int* abort_flag = reinterpret_cast<volatile int*>(gbase + kOffMisc + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kOffMisc);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / prm.splits, split = blockIdx.x - tile * prm.splits;
  const int tm = tile / prm.tiles_n, tn = tile - tm * prm.tiles_n;
  const long long kb0 = prm.nkb * split / prm.splits, kb1 = prm.nkb * (split + 1) / prm.splits;
  const long long nchunks = (kb1 - kb0 + prm.chunk_kb - 1) / prm.chunk_kb;

  // every CTA of a tile row does the bias-gradient product for the K blocks kb with kb % tiles_n == tn (spreading the extra
  // MMAs over the row's CTAs: with all of them on the tn == 0 CTAs the launch took 185 instead of 135 us)
  const bool with_db = prm.ws_db != nullptr;
  auto db_blocks = [&](long long c0, long long c1) {   // number of K blocks of [c0, c1) this CTA owns for the bias gradient
    const long long first = c0 + ((tn - c0 % prm.tiles_n) + prm.tiles_n) % prm.tiles_n;
    return first < c1 ? (c1 - first + prm.tiles_n - 1) / prm.tiles_n : 0LL;
  };
  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < kStages; ++s) { mbar_init(BAR(kBarFull + s), 1); mbar_init(BAR(kBarEmpty + s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(BAR(kBarAccFull + b), 1); mbar_init(BAR(kBarAccEmpty + b), kEpiWarps); }
    fence_mbar_init();
    prefetch_tmap(&tm_a_hi); prefetch_tmap(&tm_a_lo); prefetch_tmap(&tm_b_hi); prefetch_tmap(&tm_b_lo);
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 128) {    // the ones tile: 128 x 16 bytes of fp16 1.0 (layout-free: all equal)
    *reinterpret_cast<uint4*>(gbase + kOffOnes + 16 * (threadIdx.x - 64)) = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      long long kbg = 0;
      for (long long kb = kb0; kb < kb1; ++kb, ++kbg) {
        const int s = (int)(kbg % kStages);
        mbar_wait(BAR(kBarEmpty + s), (uint32_t)(((kbg / kStages) & 1) ^ 1), abort_flag, prm.status, 3800 + s);
        const uint32_t st = base + (uint32_t)s * kStageBytes;
        mbar_expect_tx_local(BAR(kBarFull + s), kStageBytes);
        const int row = (int)(kb * kBK);
#pragma unroll
        for (int h = 0; h < 2; ++h) {   // the two 64-wide halves of each 128-wide operand tile
          tma_load_2d(st + h * 8192u, &tm_a_hi, BAR(kBarFull + s), tm * kBM + 64 * h, row);
          tma_load_2d(st + kPlane + h * 8192u, &tm_a_lo, BAR(kBarFull + s), tm * kBM + 64 * h, row);
          tma_load_2d(st + 2 * kPlane + h * 8192u, &tm_b_hi, BAR(kBarFull + s), tn * kBN + 64 * h, row);
          tma_load_2d(st + 3 * kPlane + h * 8192u, &tm_b_lo, BAR(kBarFull + s), tn * kBN + 64 * h, row);
        }
      }
      for (long long k2 = kbg; k2 < kbg + kStages; ++k2) {
        if (k2 - kStages < 0) continue;
        const int s = (int)(k2 % kStages);
        mbar_wait(BAR(kBarEmpty + s), (uint32_t)(((k2 / kStages) & 1) ^ 1), abort_flag, prm.status, 3850 + s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_f16(kBM, kBN) | (1u << 15) | (1u << 16);   // A and B MN-major
      const uint32_t idesc_db = idesc_f16(kBM, 16) | (1u << 15);               // A MN-major, B (ones) K-major
      const uint64_t ones = smem_desc_sw128(base + kOffOnes);
      long long kbg = 0;
      for (long long ch = 0; ch < nchunks; ++ch) {
        const int buf = (int)(ch & 1);
        mbar_wait(BAR(kBarAccEmpty + buf), (uint32_t)(((ch >> 1) & 1) ^ 1), abort_flag, prm.status, 3900 + buf);
        tc_fence_after();
        const uint32_t dcol = tmem_base + (uint32_t)(buf * kBN);
        const long long c0 = kb0 + ch * prm.chunk_kb, c1 = c0 + prm.chunk_kb < kb1 ? c0 + prm.chunk_kb : kb1;
        bool db_started = false;
        for (long long kb = c0; kb < c1; ++kb, ++kbg) {
          const int s = (int)(kbg % kStages);
          mbar_wait(BAR(kBarFull + s), (uint32_t)((kbg / kStages) & 1), abort_flag, prm.status, 3950 + s);
          tc_fence_after();
          const uint32_t st = base + (uint32_t)s * kStageBytes;
          const uint64_t a_hi = smem_desc_sw128_mn(st), a_lo = smem_desc_sw128_mn(st + kPlane);
          const uint64_t b_hi = smem_desc_sw128_mn(st + 2 * kPlane), b_lo = smem_desc_sw128_mn(st + 3 * kPlane);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_1cta(dcol, a_lo + 128 * k, b_hi + 128 * k, idesc, (kb == c0 && k == 0) ? 0u : 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_1cta(dcol, a_hi + 128 * k, b_lo + 128 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_1cta(dcol, a_hi + 128 * k, b_hi + 128 * k, idesc, 1u);
          if (with_db && kb % prm.tiles_n == tn) {
            const uint32_t dbcol = tmem_base + kDbCol + (uint32_t)(buf * 16);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_1cta(dbcol, a_lo + 128 * k, ones + 2 * k, idesc_db, (!db_started && k == 0) ? 0u : 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_1cta(dbcol, a_hi + 128 * k, ones + 2 * k, idesc_db, 1u);
            db_started = true;
          }
          umma_commit_1cta(BAR(kBarEmpty + s));
        }
        umma_commit_1cta(BAR(kBarAccFull + buf));
      }
    }
  } else {
    const int q = warp & 3, cg = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    float dbacc = 0.f;
    const bool db_warp = with_db && cg == 0;          // warp-uniform
    for (long long ch = 0; ch < nchunks; ++ch) {
      const int buf = (int)(ch & 1);
      const long long c0 = kb0 + ch * prm.chunk_kb;
      const int len = (int)((c0 + prm.chunk_kb < kb1 ? c0 + prm.chunk_kb : kb1) - c0);
      const float gain = acc_trunc_gain(len);
      mbar_wait(BAR(kBarAccFull + buf), (uint32_t)((ch >> 1) & 1), abort_flag, prm.status, 3980 + buf);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kBN + cg * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = fmaf(__uint_as_float(r[j]), gain, acc[j]);
      const long long nb_db = db_warp ? db_blocks(c0, c0 + len) : 0;      // warp-uniform
      if (nb_db > 0) {
        uint32_t rb[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + kDbCol + (uint32_t)(buf * 16), rb);
        tmem_ld_wait();
        dbacc = fmaf(__uint_as_float(rb[0]), acc_trunc_gain((int)nb_db), dbacc);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_local(BAR(kBarAccEmpty + buf));
    }
    const long long ldw = (long long)prm.tiles_n * kBN;
    float* d = prm.ws + ((long long)split * prm.tiles_m * kBM + tm * kBM + row) * ldw + tn * kBN + cg * 32;
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4)
      *reinterpret_cast<float4*>(d + 4 * j4) = make_float4(acc[4 * j4], acc[4 * j4 + 1], acc[4 * j4 + 2], acc[4 * j4 + 3]);
    if (db_warp) prm.ws_db[((long long)split * prm.tiles_n + tn) * prm.tiles_m * kBM + tm * kBM + row] = dbacc;
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

namespace {

// dW[m][n] = unscale * sum over splits (in split order) of the partial tiles
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, int splits, long long split_stride,
                                                           long long ldw, int M, int N, float scale,
                                                           const float* __restrict__ dyn_scale, float* __restrict__ dW,
                                                           long long ldd) {
  const float s = scale / (dyn_scale ? __ldg(dyn_scale) : 1.f);
  const long long total = (long long)M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / N), n = (int)(i - (long long)m * N);
    float a = 0.f;
    for (int k = 0; k < splits; ++k) a += ws[k * split_stride + m * ldw + n];
    dW[m * ldd + n] = a * s;
  }
}

// db[m] = (1 / (64 dyn)) * sum over splits of the partial row sums
__global__ void __launch_bounds__(256) wgrad_reduce_db_kernel(const float* __restrict__ ws_db, int splits, long long split_stride,
                                                              int M, const float* __restrict__ dyn_scale, float* __restrict__ db) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float a = 0.f;
  for (int k = 0; k < splits; ++k) a += ws_db[k * split_stride + m];
  db[m] = a / (64.f * (dyn_scale ? __ldg(dyn_scale) : 1.f));
}

// column sums of an fp32 matrix (bias gradient), two deterministic stages: partial[rs][c] over row slab rs, then the slabs
constexpr int kCsSlabs = 128;
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ g, long long ld, long long R, int C,
                                                             float* __restrict__ partial) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
  const long long r0 = R * blockIdx.y / kCsSlabs, r1 = R * (blockIdx.y + 1) / kCsSlabs;
  float a = 0.f;
  if (c < C)
    for (long long r = r0 + ry; r < r1; r += 8) a += __ldg(g + r * ld + c);
  sh[ry][threadIdx.x & 31] = a;
  __syncthreads();
  if (ry == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x & 31];
    partial[(long long)blockIdx.y * C + c] = t;
  }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int k = 0; k < kCsSlabs; ++k) t += partial[(long long)k * C + c];
  out[c] = t;
}

}  // namespace

namespace {

// fp32 [R][C] (row stride ld) -> fp16 hi / lo planes of scale * value: out[r][c] (transpose == 0, row stride ldp) or
// out[c][r] (transpose != 0).  scale_dev: optional device scalar multiplied into scale.
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ in, long long ld, long long R, int C,
                                                           float scale, const float* __restrict__ scale_dev, int transpose,
                                                           __half* __restrict__ hi, __half* __restrict__ lo, long long ldp,
                                                           DevStatus* status) {
  const float s = scale * (scale_dev ? __ldg(scale_dev) : 1.f);
  const long long total = R * (long long)C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    float v = __ldg(in + r * ld + c) * s;
    if (!(fabsf(v) < 65504.f)) { report_range(status, 3600); v = fminf(fmaxf(v, -65504.f), 65504.f); }
    const __half h = __float2half_rn(v);
    const long long o = transpose ? (long long)c * ldp + r : r * ldp + c;
    hi[o] = h;
    lo[o] = __float2half_rn(v - __half2float(h));
  }
}

// Second-order helpers (recmv_b200/second_order.py): the element-wise step between two layer GEMMs of the tangent pass and of
// the downward pass, each ONE launch that also writes the next GEMM's operand planes.
//   softplus_tangent: tz = raw tangent of z_l, a = softplus_100(z_l) (saved), h = first-order cotangent at z_l
//       u   = s tz                 (tangent of a_{l+1};  s = softplus' = 1 - exp(-100 a))      -> fp32 + planes
//       inj = 100 (1 - s) h tz     (softplus'' u_{l+1} tz: what the second differentiation adds to the cotangent of z_l)
//   add_split: y += addend, planes of y
__global__ void __launch_bounds__(256) softplus_tangent_kernel(const float* __restrict__ tz, long long ldt, const float* __restrict__ a,
                                                               long long lda, const float* __restrict__ h, long long ldh,
                                                               long long R, int C, float plane_scale, float* __restrict__ u,
                                                               long long ldu, __half* __restrict__ uh, __half* __restrict__ ul,
                                                               long long ldp, float* __restrict__ inj, long long ldi,
                                                               DevStatus* status) {
  const long long total = R * (long long)C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float t = __ldg(tz + r * ldt + c);
    const float em = expm1f(-100.f * __ldg(a + r * lda + c));      // -s, accurate where s is small
    const float uv = -em * t;
    u[r * ldu + c] = uv;
    inj[r * ldi + c] = 100.f * (1.f + em) * __ldg(h + r * ldh + c) * t;
    float v = uv * plane_scale;
    if (!(fabsf(v) < 65504.f)) { report_range(status, 3700); v = fminf(fmaxf(v, -65504.f), 65504.f); }
    const __half hh = __float2half_rn(v);
    uh[r * ldp + c] = hh;
    ul[r * ldp + c] = __float2half_rn(v - __half2float(hh));
  }
}

__global__ void __launch_bounds__(256) add_split_kernel(float* __restrict__ y, long long ldy, const float* __restrict__ addend,
                                                        long long lda, long long R, int C, float scale,
                                                        const float* __restrict__ scale_dev, __half* __restrict__ yh,
                                                        __half* __restrict__ yl, long long ldp, DevStatus* status) {
  const float s = scale * (scale_dev ? __ldg(scale_dev) : 1.f);
  const long long total = R * (long long)C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float o = y[r * ldy + c] + __ldg(addend + r * lda + c);
    y[r * ldy + c] = o;
    float v = o * s;
    if (!(fabsf(v) < 65504.f)) { report_range(status, 3710); v = fminf(fmaxf(v, -65504.f), 65504.f); }
    const __half hh = __float2half_rn(v);
    yh[r * ldp + c] = hh;
    yl[r * ldp + c] = __float2half_rn(v - __half2float(hh));
  }
}

// positional encoding as a saved layer input AND as operand planes.  One thread per (point, column): consecutive threads write
// consecutive columns (the one-thread-per-point form wrote 39-element rows at a 2 KB stride from every lane: 139 us for 131 k
// points under ncu).  Same arithmetic as positional_encode (common.cuh): w * sin / cos (2^k x), accurate sincosf.
__global__ void __launch_bounds__(256) pe_forward_planes_kernel(const float* __restrict__ x, PeWeights pw, int bands,
                                                                float* __restrict__ out, long long ld, __half* __restrict__ oh,
                                                                __half* __restrict__ ol, long long ldp, long long P) {
  const int n = 3 + 6 * bands;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long p = i / 40;
  const int e = (int)(i - p * 40);
  if (p >= P || e >= n) return;
  float val;
  if (e < 3) {
    val = x[3 * p + e];
  } else {
    const int k = (e - 3) / 6, r = (e - 3) - 6 * k, c = r % 3;
    float s, co;
    sincosf(x[3 * p + c] * (float)(1 << k), &s, &co);
    val = r < 3 ? pw.w[2 * k] * s : pw.w[2 * k + 1] * co;
  }
  if (out) out[p * ld + e] = val;
  const float v = val * kActScale;
  const __half h = __float2half_rn(v);
  oh[p * ldp + e] = h;
  ol[p * ldp + e] = __float2half_rn(v - __half2float(h));
}

int make_plane_tmap(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t row_pitch_elems,
                    uint32_t box_rows = 128) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || !p) return e != cudaSuccess ? (int)e : (int)cudaErrorNotSupported;
    fn = (EncodeTiledFn)p;
  }
  if (((uintptr_t)base & 15) != 0 || (row_pitch_elems & 7) != 0) return RECMV_E_SHAPE;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_pitch_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

}  // namespace
}  // namespace recmv

using namespace recmv;

// fp32 [R][C] -> fp16 hi / lo planes of scale (* *scale_dev) * value, optionally transposed ([C][R]); ldp = row stride of the
// planes in elements (multiple of 8).  Weights: scale 1024; activations 64; cotangents 64 with scale_dev = the dyn scale.
extern "C" int recmv_split_planes(const float* in, int64_t ld, int64_t R, int C, float scale, const float* scale_dev,
                                  int transpose, void* hi, void* lo, int64_t ldp, recmv_stream_t stream) {
  if (R < 0 || C <= 0) return RECMV_E_SHAPE;
  if (R == 0) return RECMV_OK;
  if (!in || !hi || !lo) return RECMV_E_NULL;
  void* sd = nullptr;
  int s0 = device_status_record(&sd);
  if (s0) return s0;
  split_planes_kernel<<<stride_grid(R * (int64_t)C, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      in, ld, R, C, scale, scale_dev, transpose, (__half*)hi, (__half*)lo, ldp, (DevStatus*)sd);
  return launch_status();
}

extern "C" int recmv_softplus_tangent_planes(const float* tz, int64_t ldt, const float* a, int64_t lda, const float* h,
                                            int64_t ldh, int64_t rows, int cols, float plane_scale, float* u, int64_t ldu,
                                            void* u_hi, void* u_lo, int64_t ldp, float* inj, int64_t ldi,
                                            recmv_stream_t stream) {
  if (rows < 0 || cols <= 0) return RECMV_E_SHAPE;
  if (rows == 0) return RECMV_OK;
  if (!tz || !a || !h || !u || !u_hi || !u_lo || !inj) return RECMV_E_NULL;
  void* sd = nullptr;
  int s0 = device_status_record(&sd);
  if (s0) return s0;
  softplus_tangent_kernel<<<stride_grid(rows * (int64_t)cols, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      tz, ldt, a, lda, h, ldh, rows, cols, plane_scale, u, ldu, (__half*)u_hi, (__half*)u_lo, ldp, inj, ldi, (DevStatus*)sd);
  return launch_status();
}

extern "C" int recmv_add_split_planes(float* y, int64_t ldy, const float* addend, int64_t lda, int64_t rows, int cols,
                                      float scale, const float* scale_dev, void* y_hi, void* y_lo, int64_t ldp,
                                      recmv_stream_t stream) {
  if (rows < 0 || cols <= 0) return RECMV_E_SHAPE;
  if (rows == 0) return RECMV_OK;
  if (!y || !addend || !y_hi || !y_lo) return RECMV_E_NULL;
  void* sd = nullptr;
  int s0 = device_status_record(&sd);
  if (s0) return s0;
  add_split_kernel<<<stride_grid(rows * (int64_t)cols, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      y, ldy, addend, lda, rows, cols, scale, scale_dev, (__half*)y_hi, (__half*)y_lo, ldp, (DevStatus*)sd);
  return launch_status();
}

extern "C" int recmv_pe_forward_planes(const float* x, const float* pe_w, int bands, float* out, int64_t ld, void* out_hi,
                                       void* out_lo, int64_t ldp, int64_t P, recmv_stream_t stream) {
  if (P < 0 || bands < 0 || bands > 6) return RECMV_E_SHAPE;
  if (P == 0) return RECMV_OK;
  if (!x || !pe_w || !out_hi || !out_lo) return RECMV_E_NULL;
  PeWeights pw;
  for (int i = 0; i < 12; ++i) pw.w[i] = i < 2 * bands ? pe_w[i] : 0.f;
  pe_forward_planes_kernel<<<(unsigned)((P * 40 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, pw, bands, out, ld, (__half*)out_hi,
                                                                                          (__half*)out_lo, ldp, P);
  return launch_status();
}

// One layer GEMM with TMA-fed operand planes.
//   forward       (mode 4/5/6 = none / softplus100 / ReLU):  Y = act(pre_scale * (A . B^T) + bias)
//   backward-data (mode 0/1/2 = none / softplus100' / ReLU'): Y = (A . B^T) * out_scale * act'(saved_input), columns >= split
//                 to Y2 without the derivative
// A planes [M][lda_p] (K columns used, scaled 64 [x dyn when a_has_dyn]), B planes [N][ldb_p] (scaled 1024).
// Outputs: Y fp32 [M][ldy]; optional planes of the < split columns, scaled 64 (x dyn when planes_with_dyn).
extern "C" int recmv_mlp_layer_planes(const void* a_hi, const void* a_lo, int64_t lda_p, const void* b_hi, const void* b_lo,
                                      int64_t ldb_p, int64_t M, int N, int K, int mode, const float* bias,
                                      const float* saved_input, int64_t lds, float scale, const float* dyn_scale,
                                      int a_has_dyn, int split, float* Y, int64_t ldy, float* Y2, int64_t ldy2, void* y_hi,
                                      void* y_lo, int64_t ldyp, int planes_with_dyn, recmv_stream_t stream) {
  if (M < 0 || N <= 0 || K <= 0) return RECMV_E_SHAPE;
  if (M == 0) return RECMV_OK;
  if (!a_hi || !a_lo || !b_hi || !b_lo || !Y) return RECMV_E_NULL;
  if (!(mode == 0 || mode == 1 || mode == 2 || mode == 4 || mode == 5 || mode == 6)) return RECMV_E_DTYPE;
  if ((mode == 1 || mode == 2) && !saved_input) return RECMV_E_NULL;
  if ((y_hi == nullptr) != (y_lo == nullptr)) return RECMV_E_NULL;
  if (M > (int64_t)1 << 30) return RECMV_E_RANGE;
  CUtensorMap ta_h, ta_l, tb_h, tb_l;
  int s = make_plane_tmap(&ta_h, a_hi, (uint64_t)K, (uint64_t)M, (uint64_t)lda_p);
  if (!s) s = make_plane_tmap(&ta_l, a_lo, (uint64_t)K, (uint64_t)M, (uint64_t)lda_p);
  if (!s) s = make_plane_tmap(&tb_h, b_hi, (uint64_t)K, (uint64_t)N, (uint64_t)ldb_p);
  if (!s) s = make_plane_tmap(&tb_l, b_lo, (uint64_t)K, (uint64_t)N, (uint64_t)ldb_p);
  if (s) return s;
  static bool done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!done[dev & 15]) {
    cudaError_t e = cudaFuncSetAttribute(gemm3_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    done[dev & 15] = true;
  }
  void* sd = nullptr;
  int s0 = device_status_record(&sd);
  if (s0) return s0;
  if (((DevStatus*)sd)->code != 0) return RECMV_E_DEVICE;
  GtParams prm = {};
  prm.D = Y; prm.ldd = ldy; prm.D2 = Y2; prm.ldd2 = ldy2;
  prm.DH = (__half*)y_hi; prm.DL = (__half*)y_lo; prm.ldp = ldyp;
  prm.E = saved_input; prm.lde = lds; prm.bias = bias;
  prm.dyn_scale = a_has_dyn ? dyn_scale : nullptr;
  prm.unscale = scale / (kActScale * kWgtScale);
  prm.plane_scale = kActScale;
  prm.dyn_in_planes = 0;
  if (planes_with_dyn) {
    // cotangent planes carry the dyn scale; when the A planes do not (never the case in backward-data) it would have to be
    // multiplied in separately
    if (!a_has_dyn || !dyn_scale) return RECMV_E_UNSUPPORTED;
    prm.dyn_in_planes = 1;
  }
  prm.M = (int)M; prm.N = N; prm.K = K;
  prm.split = (split > 0 && split < N) ? split : N;
  prm.tiles_m = (int)((M + kBM - 1) / kBM); prm.tiles_n = (N + kBN - 1) / kBN;
  prm.epi = mode; prm.status = (DevStatus*)sd;
  const int total = prm.tiles_m * prm.tiles_n;
  const int grid = total < num_sms() ? total : num_sms();
  gemm3_tma_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(ta_h, ta_l, tb_h, tb_l, prm);
  return launch_status();
}

// dW [out_dim][in_dim] (row stride in_dim) = scale * G^T X from the planes the layer GEMMs wrote: g planes [P][ldg_p] hold
// 64 * dyn * g (columns < out_dim), x planes [P][ldx_p] hold 64 * x (columns < in_dim).  workspace: fp32, at least
// recmv_mlp_wgrad_workspace_floats() elements, reused across calls on one stream.
extern "C" size_t recmv_mlp_wgrad_workspace_floats(void) {
  int dev = 0;
  cudaGetDevice(&dev);
  return (size_t)(num_sms() + 16) * (kBM * kBN + kBM);   // tiles x splits <= SMs partial tiles + as many partial row sums
}

extern "C" int recmv_mlp_wgrad_planes(const void* g_hi, const void* g_lo, int64_t ldg_p, const void* x_hi, const void* x_lo,
                                      int64_t ldx_p, int64_t P, int out_dim, int in_dim, float scale, const float* dyn_scale,
                                      float* workspace, float* dW, float* db, recmv_stream_t stream) {
  if (P < 0 || out_dim <= 0 || in_dim <= 0) return RECMV_E_SHAPE;
  if (!g_hi || !g_lo || !x_hi || !x_lo || !workspace || !dW) return RECMV_E_NULL;
  if (P > (int64_t)1 << 30) return RECMV_E_RANGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (P == 0) {
    cudaError_t e0 = cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)out_dim * in_dim, st);
    if (e0 == cudaSuccess && db) e0 = cudaMemsetAsync(db, 0, sizeof(float) * (size_t)out_dim, st);
    return (int)e0;
  }
  CUtensorMap ta_h, ta_l, tb_h, tb_l;
  int s = make_plane_tmap(&ta_h, g_hi, (uint64_t)out_dim, (uint64_t)P, (uint64_t)ldg_p, 64);
  if (!s) s = make_plane_tmap(&ta_l, g_lo, (uint64_t)out_dim, (uint64_t)P, (uint64_t)ldg_p, 64);
  if (!s) s = make_plane_tmap(&tb_h, x_hi, (uint64_t)in_dim, (uint64_t)P, (uint64_t)ldx_p, 64);
  if (!s) s = make_plane_tmap(&tb_l, x_lo, (uint64_t)in_dim, (uint64_t)P, (uint64_t)ldx_p, 64);
  if (s) return s;
  void* sd = nullptr;
  s = device_status_record(&sd);
  if (s) return s;
  static bool attr_done[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 15]) {
    cudaError_t e = cudaFuncSetAttribute(gemm3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesWg);
    if (e != cudaSuccess) return (int)e;
    attr_done[dev & 15] = true;
  }
  WgParams prm = {};
  prm.ws = workspace;
  prm.nkb = (P + kBK - 1) / kBK;
  prm.M = out_dim; prm.N = in_dim;
  prm.tiles_m = (out_dim + kBM - 1) / kBM; prm.tiles_n = (in_dim + kBN - 1) / kBN;
  const int tiles = prm.tiles_m * prm.tiles_n;
  int splits = num_sms() / tiles;
  if (splits < 1) splits = 1;
  if ((long long)splits > prm.nkb) splits = (int)prm.nkb;
  prm.splits = splits;
  prm.chunk_kb = 32;
  prm.status = (DevStatus*)sd;
  const long long db_stride = (long long)prm.tiles_m * kBM;
  prm.ws_db = db ? workspace + (size_t)splits * prm.tiles_m * kBM * prm.tiles_n * kBN : nullptr;
  gemm3_wgrad_kernel<<<tiles * splits, kThreads, kSmemBytesWg, st>>>(ta_h, ta_l, tb_h, tb_l, prm);
  s = launch_status();
  if (s) return s;
  const long long ldw = (long long)prm.tiles_n * kBN, split_stride = (long long)prm.tiles_m * kBM * ldw;
  wgrad_reduce_kernel<<<stride_grid((int64_t)out_dim * in_dim, 256, 4), 256, 0, st>>>(
      workspace, splits, split_stride, ldw, out_dim, in_dim, scale / (64.f * 64.f), dyn_scale, dW, in_dim);
  s = launch_status();
  if (s || !db) return s;
  wgrad_reduce_db_kernel<<<(out_dim + 255) / 256, 256, 0, st>>>(prm.ws_db, splits * prm.tiles_n, db_stride, out_dim, dyn_scale, db);
  return launch_status();
}

// out[c] = sum_r g[r][c], c < cols (bias gradient); partial: fp32 scratch of 128 * cols elements
extern "C" int recmv_colsum(const float* g, int64_t ld, int64_t rows, int cols, float* partial, float* out,
                            recmv_stream_t stream) {
  if (rows < 0 || cols <= 0) return RECMV_E_SHAPE;
  if (!g || !partial || !out) return RECMV_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  colsum_partial_kernel<<<dim3((unsigned)((cols + 31) / 32), kCsSlabs), 256, 0, st>>>(g, ld, rows, cols, partial);
  int s = launch_status();
  if (s) return s;
  colsum_final_kernel<<<(cols + 255) / 256, 256, 0, st>>>(partial, cols, out);
  return launch_status();
}

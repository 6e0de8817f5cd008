// tcgen05 issue-rate microbenchmark (diagnostics only; not on any product path).
// Elected threads issue back-to-back kind::f16 MMAs of a given M x N x 16 shape over SW128 K-major operand
// tiles in shared memory and report cycles per MMA, with the issue-thread overheads of a real pipeline
// (tcgen05.commit, mbarrier try_wait) and background load (epilogue-like warps, bulk copies) switched on
// one by one.  Used to choose the issue structure of sdf_mlp_tc.cu ("MMA issue path").
// F is a compile-time bit set so the issue loop contains only what is being measured:
//   1 alternate two accumulators per K block      2 eight epilogue-like warps (TMEM reads + smem stores)
//   4 background bulk copies global -> B stages   8 commits are multicast to both CTAs
//  16 commit every K block (4 MMAs)             256 commit every 2 K blocks (8 MMAs)    512 every 3 (12 MMAs)
//  32 blocking try_wait (completed phase) per K block    64 the same, issued BEFORE the K block's MMAs and
//     consumed after them (software pipelined)  128 two issuing threads (warps 0 and 2), one accumulator each
#include "../../include/recmv_b200_diag.h"
#include "tc_common.cuh"

namespace recmv {
namespace {
using namespace tc;

template <int CG>
__device__ __forceinline__ void mma_any(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 2) umma_f16_pair(d, a, b, idesc, acc);
  else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG, bool kMcast>
__device__ __forceinline__ void commit_any(uint32_t bar) {
  if (CG == 2 && kMcast) umma_commit_pair(bar, 3);
  else if (CG == 2) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t try_wait_issue(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}

template <int CG, int F>
__global__ void __launch_bounds__(384, 1)
tc_microbench_kernel(int M, int N, int iters, const uint8_t* gsrc, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int rows_a = M / CG, rows_b = N / CG;
  const uint32_t a_kb = rows_a * 128u, b_kb = rows_b * 128u;   // bytes per 64-wide K block
  const uint32_t a_base = base;                                // 8 K blocks
  const uint32_t b_base = base + 8 * a_kb;                     // 2 stages
  const uint32_t bar = b_base + 2 * b_kb;                      // [2] final commit per issuing thread
  const uint32_t tslot = bar + 16;
  const uint32_t bar_dummy = bar + 32;                         // [2] periodic commits (nobody waits)
  const uint32_t bar_copy = bar + 48;                          // [2] bulk-copy completion per B stage
  const uint32_t bar_done = bar + 64;                          // a barrier whose phase 1 wait succeeds at once
  volatile int* done = reinterpret_cast<volatile int*>(smem_raw + (bar + 96 - smem_u32(smem_raw)));
  for (uint32_t i = threadIdx.x; i < (8 * a_kb + 2 * b_kb) / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4 * i), "r"(0x2C003C00u ^ ((i * 2654435761u) & 0x03FF03FFu)));
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(bar + 8 * i + (i >= 2 ? 16 : 0), 1);
    *done = 0;
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    if (CG == 2) asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tslot) : "memory");
    else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tslot) : "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tslot));

  constexpr bool kTwo = (F & 128) != 0;
  const int issuer = threadIdx.x == 0 ? 0 : ((kTwo && threadIdx.x == 64) ? 1 : -1);
  if (issuer >= 0 && rank == 0) {
    const uint32_t idesc = idesc_f16(M, N);
    const int my_iters = kTwo ? iters / 2 : iters;
    const uint32_t my_bar = bar + 8 * issuer, my_dummy = bar_dummy + 8 * issuer;
    const long long t0 = clock64();
    for (int it = 0; it < my_iters; ++it) {
#pragma unroll
      for (int kb = 0; kb < 8; ++kb) {
        const uint64_t da = smem_desc_sw128(a_base + kb * a_kb);
        const uint64_t db = smem_desc_sw128(b_base + (kb & 1) * b_kb);
        uint32_t tmem_d = tmem;
        if (kTwo) tmem_d += issuer * 256u;
        else if ((F & 1) && (kb & 1)) tmem_d += 256u;
        uint32_t early = 1;
        if (F & 64) early = try_wait_issue(bar_done, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_any<CG>(tmem_d, da + 2 * k, db + 2 * k, idesc, (it | (kb >> 1) | k) ? 1u : 0u);
        if ((F & 16) || ((F & 256) && (kb & 1)) || ((F & 512) && (kb % 3 == 2))) commit_any<CG, (F & 8) != 0>(my_dummy);
        if (F & 64) { int spins = 0; while (!early && ++spins < 4) early = try_wait_issue(bar_done, 1); }
        if (F & 32) { int spins = 0; while (!mbar_try_wait(bar_done, 1) && ++spins < 4) {} }
      }
    }
    commit_any<CG, false>(my_bar);
    const long long t_issue = clock64();
    int spins = 0;
    while (!mbar_try_wait(my_bar, 0) && ++spins < (1 << 26)) {}
    const long long t1 = clock64();
    if (blockIdx.x == 0) { out[4 * issuer + 0] = (unsigned long long)(t1 - t0); out[4 * issuer + 1] = (unsigned long long)(t_issue - t0); }
    if (kTwo) {   // both issuers finished -> stop the helpers (the second arrival releases them)
      if (atomicAdd((int*)done, 1) == 1) {
        if (CG == 2) asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(mapa(bar + 96, 1)), "r"(2) : "memory");
      }
    } else {
      *done = 2;
      if (CG == 2) asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(mapa(bar + 96, 1)), "r"(2) : "memory");
    }
  }
  // ---- background load: epilogue-like TMEM reads + smem stores (warps 4..11) ---------------------------------
  if ((F & 2) && threadIdx.x >= 128) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lane_addr = (uint32_t)((w & 3) * 32) << 16;
    uint32_t n = 0;
    while (*done < 2 && n < (1u << 22)) {
      uint32_t r[32];
      tmem_ld32(tmem + lane_addr + ((n * 32u) & 127u), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * j + e]) * 0.001f + 1.f;
        uint4 hi, lo;
        split8(v, hi, lo);
        const uint32_t off = ((n & 7u) * a_kb) + sw128_offset(lane + 32 * ((w >> 2) & 1), j + 4 * (n & 1));
        st_shared_v4(a_base + off, hi);
        st_shared_v4(a_base + ((off + 4 * a_kb) & (8 * a_kb - 1)), lo);
      }
      ++n;
    }
  }
  // ---- background load: bulk copies global -> B stages (one thread of warp 1) ---------------------------------
  if ((F & 4) && threadIdx.x == 32) {
    uint32_t n = 0, par[2] = {0, 0};
    const uint8_t* src = gsrc + (size_t)blockIdx.x * 65536;
    while (*done < 2 && n < (1u << 20)) {
      const int st = n & 1;
      if (n >= 2) { int spins = 0; while (!mbar_try_wait(bar_copy + 8 * st, par[st]) && ++spins < (1 << 22)) {} par[st] ^= 1u; }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_copy + 8 * st), "r"(b_kb) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(b_base + st * b_kb), "l"(src + ((n * b_kb) & 32767u)), "r"(b_kb), "r"(bar_copy + 8 * st) : "memory");
      ++n;
    }
    for (int st = 0; st < 2; ++st)   // drain outstanding copies before the CTA exits
      if (n > (uint32_t)st) { int spins = 0; while (!mbar_try_wait(bar_copy + 8 * st, par[st]) && ++spins < (1 << 22)) {} }
    if (blockIdx.x == 0) out[3] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (threadIdx.x < 32) {
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int CG, int F>
cudaError_t launch_mb(cudaLaunchConfig_t& cfg, int M, int N, int iters, const uint8_t* gsrc, unsigned long long* out) {
  cudaFuncSetAttribute(tc_microbench_kernel<CG, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
  return cudaLaunchKernelEx(&cfg, tc_microbench_kernel<CG, F>, M, N, iters, gsrc, out);
}

}  // namespace
}  // namespace recmv

// out (device, 8 x u64): per issuing thread {total cycles incl. drain, issue-loop cycles, -, copies}; iters * 32
// MMAs are issued in total.  `flags` must be one of the instantiated sets (RECMV_E_UNSUPPORTED otherwise).
extern "C" int recmv_tc_microbench(int cta_group, int M, int N, int iters, int num_ctas, int flags, const void* gsrc,
                                   unsigned long long* out, recmv_stream_t stream) {
  using namespace recmv;
  if (!out || ((flags & 4) && !gsrc)) return RECMV_E_NULL;
  if ((cta_group != 1 && cta_group != 2) || iters <= 0 || (iters & 1) || num_ctas <= 0 || num_ctas % cta_group) return RECMV_E_RANGE;
  if (N % 16 || N < 16 || N > 256) return RECMV_E_SHAPE;
  if (cta_group == 1 ? (M != 64 && M != 128) : (M != 128 && M != 256)) return RECMV_E_SHAPE;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_ctas);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = 1024 + 8 * (size_t)(M / cta_group) * 128 + 2 * (size_t)(N / cta_group) * 128 + 160;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = cta_group; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  const uint8_t* g = (const uint8_t*)gsrc;
  cudaError_t e = cudaErrorInvalidValue;
  if (cta_group == 1) {
    if (flags == 0) e = launch_mb<1, 0>(cfg, M, N, iters, g, out);
    else if (flags == 16) e = launch_mb<1, 16>(cfg, M, N, iters, g, out);
    else return RECMV_E_UNSUPPORTED;
  } else {
    switch (flags) {
#define MB_CASE(F) case F: e = launch_mb<2, F>(cfg, M, N, iters, g, out); break;
      MB_CASE(0) MB_CASE(1) MB_CASE(2) MB_CASE(4) MB_CASE(16) MB_CASE(24) MB_CASE(32) MB_CASE(64) MB_CASE(256) MB_CASE(512)
      MB_CASE(16 + 32) MB_CASE(16 + 64) MB_CASE(256 + 64) MB_CASE(512 + 64) MB_CASE(128) MB_CASE(128 + 16) MB_CASE(128 + 16 + 32)
      MB_CASE(128 + 16 + 64) MB_CASE(128 + 256 + 64)
      MB_CASE(2 + 16) MB_CASE(2 + 16 + 32) MB_CASE(2 + 16 + 64) MB_CASE(2 + 256 + 64) MB_CASE(2 + 512 + 64) MB_CASE(2 + 128)
      MB_CASE(2 + 128 + 16 + 32) MB_CASE(2 + 128 + 16 + 64) MB_CASE(2 + 128 + 256 + 64) MB_CASE(2 + 4 + 128 + 16 + 64 + 8)
      MB_CASE(2 + 4 + 512 + 64 + 8)
#undef MB_CASE
      default: return RECMV_E_UNSUPPORTED;
    }
  }
  return e == cudaSuccess ? RECMV_OK : RECMV_E_DEVICE;
}

// sm_100a primitives used by the tcgen05 kernels: mbarrier (with bounded waits), TMA tile loads for a
// CTA pair, tcgen05 alloc / mma / commit / ld, shared-memory and instruction descriptors.
// Encodings follow the PTX ISA as exposed by CUTLASS's cute/arch/mma_sm100_desc.hpp (field layout of
// the 64-bit smem descriptor and the 32-bit instruction descriptor) -- restated here, not included.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace recmv {
namespace tc {

// ------------------------------------------------------------------------------------------------
// error reporting from device code: first failing wait wins; later waits return immediately
// ------------------------------------------------------------------------------------------------
struct DevStatus {
  int code;      // 0 ok, else which wait timed out
  int detail;    // role / barrier tag
  int block;
  int pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA's layout) in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (TMA / tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Arrive on a barrier that lives in CTA `rank` of the cluster.  Default (release.cta) semantics, as
// CUTLASS's ClusterBarrier does: the data these barriers guard is consumed through the async proxy
// (tensor-core operand reads) after a fence.proxy.async by the writer.  The .release.cluster /
// .acquire.cluster forms compile to MEMBAR + ERRBAR on the arrive and a CCTL.IVALL (L1 invalidate-all,
// ~700 cycles measured) after EVERY try_wait -- that made the single MMA-issuing thread the bottleneck
// of the first version of this kernel (profiles/r01_tc_timeline_before.txt).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  uint32_t remote = mapa(bar, rank);
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar, uint32_t rank, uint32_t bytes) {
  uint32_t remote = mapa(bar, rank);
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(remote), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// Bounded wait.  `abort_flag` is a CTA-shared int: once any wait in the CTA has timed out all later
// waits fall through immediately, so a protocol bug ends the kernel in milliseconds with a status
// instead of hanging the GPU.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag,
                                          DevStatus* status, int tag) {
  if (mbar_try_wait(bar, parity)) return true;
  long long t0 = clock64();
  while (true) {
    if (mbar_try_wait(bar, parity)) return true;
    if (*abort_flag) return false;
    if (clock64() - t0 > 400000000LL) {  // ~0.2 s at 2 GHz
      *abort_flag = 1;
      // status lives in mapped pinned host memory: plain system-visible stores (first writer wins is not
      // needed -- any timed-out wait is a valid report)
      volatile DevStatus* vs = status;
      vs->detail = tag;
      vs->block = blockIdx.x;
      __threadfence_system();
      vs->code = 1;
      __threadfence_system();
      return false;
    }
  }
}

// ---- TMA ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 2-D tile load, multicast: the tile lands at the same smem offset of every CTA in `mask`; with cta_group::2
// the completion bytes of each destination are signalled on the barrier (same offset) of the leader of
// that destination's CTA pair.
__device__ __forceinline__ void tma_load_2d_pair_mcast(uint32_t smem_dst, const void* tmap, uint32_t bar_cluster_addr,
                                                       uint16_t mask, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(smem_dst), "l"(tmap), "r"(bar_cluster_addr), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_local(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// ---- tcgen05 ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16, issued by ONE thread of the leader CTA for the pair
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread complete -> one arrival on the barrier at the same smem
// offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

// 32 lanes x 32 columns of fp32: thread i of the warp gets TMEM lane (quadrant base + i), columns c..c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t r[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
template <int kCols> __device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* r) {
  static_assert(kCols == 16 || kCols == 32, "chunk width");
  if (kCols == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------------
// K-major, 128-byte swizzle: rows of 128 B (64 fp16 along K), 8-row groups 1024 B apart (SBO), atoms
// 1024-B aligned (base_offset 0), descriptor version 1 (Blackwell).  bits: [0,14) addr>>4,
// [16,30) LBO>>4 (ignored for swizzled K-major; 1), [32,46) SBO>>4, [46,48) version, [61,64) layout.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D fp32 (bits 4-5 = 1), A/B fp16 (formats 0), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28.  M is the pair's M (128 for two 64-row CTAs).
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a [rows x 128 B] SW128 tile
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// Gain that undoes the tensor core's truncating fp32 accumulation after `kblocks` 64-wide K blocks of 3-pass MMAs into
// one accumulator (the layer GEMMs of gemm3.cu / gemm3_tma.cu).  Measured on B200 as the signed projection
// <ours, fp64> / <fp64, fp64> - 1 of a GEMM's output (tools/dbg_bias.py; identical for positive and signed operands):
//     K blocks        1        5        8        32
//     bias x 2^24   -1.2    -12.8    -24.4    -102          ->  bias ~ -(3.23 n - 1.5) 2^-24
// (the first MMA overwrites the accumulator and loses nothing; later ones drop a fraction of an ulp of a growing sum).
// With the earlier 4 n 2^-24 every GEMM came out +4.5e-7 high, +4.8e-6 on the 18-GEMM input gradient of the SDF network.
__host__ __device__ __forceinline__ float acc_trunc_gain(int kblocks) {
  return 1.f + (3.23f * (float)kblocks - 1.5f) * 5.9604645e-8f;
}

// fp32 -> fp16 hi/lo split of 8 consecutive K values, packed for one 16-byte store each.
// cvt.rn.f16x2.f32 packs two conversions into one ALU-pipe instruction (F2FP.PACK_AB); the scalar
// __float2half_rn path compiles to F2F on the quarter-rate XU pipe, which made the epilogue XU-bound.
// .satfinite: an operand beyond fp16's range saturates to +-65504 instead of becoming inf (hi = inf would make
// lo = v - inf = -inf and the MMA NaN); the writers below additionally raise the overflow status (range_check8).
__device__ __forceinline__ uint32_t pack_f16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
// Operand range of the tcgen05 path: scaled activations (2^6 a) and weights (2^10 w) must stay below fp16's
// 65504, i.e. |a| < 1023.5, |w| < 63.97.  The reference's fp32 has no such limit, so a violation is REPORTED
// (status code 2 in the mapped record -> recmv_check_async_errors / RECMV_E_DEVICE), never silent.
constexpr int kStatusTimeout = 1, kStatusRange = 2;
__device__ __forceinline__ void report_range(DevStatus* status, int tag) {
  volatile DevStatus* vs = status;
  if (vs->code == 0) {
    vs->detail = tag;
    vs->block = blockIdx.x;
    __threadfence_system();
    vs->code = kStatusRange;
  }
}
__device__ __forceinline__ void range_check8(const float v[8], DevStatus* status, int tag) {
  const float m = fmaxf(fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))),
                        fmaxf(fmaxf(fabsf(v[4]), fabsf(v[5])), fmaxf(fabsf(v[6]), fabsf(v[7]))));
  if (!(m < 65504.f)) report_range(status, tag);   // also catches NaN
}
__device__ __forceinline__ void split8(const float v[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_f16x2(v[2 * i], v[2 * i + 1]);
    const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
    l[i] = pack_f16x2(v[2 * i] - back.x, v[2 * i + 1] - back.y);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- host: tensor map over the fp16 weight panels -------------------------------------------------------
// 2-D tensor [rows, 64] fp16 (128-byte rows), box [64 x box_rows], 128-byte swizzle.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int make_panel_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint32_t box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || !p) return e != cudaSuccess ? (int)e : (int)cudaErrorNotSupported;
    fn = (EncodeTiledFn)p;
  }
  cuuint64_t dims[2] = {64, rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

}  // namespace tc
}  // namespace recmv

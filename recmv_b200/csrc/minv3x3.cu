// Batched 3x3 inverse and its VJP.
// Replaces FastMinv/Matrix3x3InvKernels.cu:21-104 (cu3x3MInv, cu3x3MInv_backward).
//
// HBM-bound: 72 B in + 37 B out per matrix (fp32).  The reference uses one thread per matrix with
// nine strided scalar loads; here a CTA stages 256 matrices through shared memory with fully
// coalesced 128-bit loads/stores (scalar for the ragged tail), then each thread inverts one matrix out of shared memory
// (stride 9 words => conflict-free), on the caller's stream (the reference launches on the legacy
// default stream).
#include "common.cuh"

namespace recmv {

constexpr int kMinvThreads = 256;

// Global <-> shared staging of one CTA tile (256 matrices).  Full tiles of a 16-byte aligned array move as
// 128-bit words (a tile is 9216 / 18432 bytes, so every tile start stays aligned); the ragged last tile and
// unaligned views fall back to scalars.
template <typename T>
__device__ __forceinline__ void stage_in(const T* __restrict__ g, T* s, int64_t base, int64_t n_elems) {
  // n_elems = valid scalars for this CTA (<= 256*9)
  if (n_elems == kMinvThreads * 9 && ((uintptr_t)(g + base) & 15) == 0) {
    const uint4* gv = reinterpret_cast<const uint4*>(g + base);
    uint4* sv = reinterpret_cast<uint4*>(s);
    constexpr int kVec = kMinvThreads * 9 * (int)sizeof(T) / 16;
#pragma unroll
    for (int i = threadIdx.x; i < kVec; i += kMinvThreads) sv[i] = __ldg(gv + i);
    return;
  }
  for (int i = threadIdx.x; i < kMinvThreads * 9; i += kMinvThreads)
    s[i] = (i < n_elems) ? g[base + i] : (T)0;
}
template <typename T>
__device__ __forceinline__ void stage_out(T* __restrict__ g, const T* s, int64_t base, int64_t n_elems) {
  if (n_elems == kMinvThreads * 9 && ((uintptr_t)(g + base) & 15) == 0) {
    uint4* gv = reinterpret_cast<uint4*>(g + base);
    const uint4* sv = reinterpret_cast<const uint4*>(s);
    constexpr int kVec = kMinvThreads * 9 * (int)sizeof(T) / 16;
#pragma unroll
    for (int i = threadIdx.x; i < kVec; i += kMinvThreads) gv[i] = sv[i];
    return;
  }
  for (int i = threadIdx.x; i < n_elems; i += kMinvThreads) g[base + i] = s[i];
}

template <typename T>
__global__ void __launch_bounds__(kMinvThreads) minv3x3_fwd_kernel(const T* __restrict__ ms,
                                                                   T* __restrict__ invs,
                                                                   uint8_t* __restrict__ ok,
                                                                   int64_t n) {
  __shared__ __align__(16) T s[kMinvThreads * 9];
  for (int64_t blk = blockIdx.x; blk * kMinvThreads < n; blk += gridDim.x) {
    int64_t m0 = blk * kMinvThreads;
    int64_t cnt = min((int64_t)kMinvThreads, n - m0);
    stage_in(ms, s, m0 * 9, cnt * 9);
    __syncthreads();
    T m[9], inv[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = s[threadIdx.x * 9 + i];
    bool good = inv3x3<T>(m, inv);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 9; ++i) s[threadIdx.x * 9 + i] = inv[i];
    if (threadIdx.x < cnt) ok[m0 + threadIdx.x] = good ? 1 : 0;
    __syncthreads();
    stage_out(invs, s, m0 * 9, cnt * 9);
    __syncthreads();
  }
}

// out = -inv^T g inv^T   (Matrix3x3InvKernels.cu:91-102 expanded form)
template <typename T>
__global__ void __launch_bounds__(kMinvThreads) minv3x3_bwd_kernel(const T* __restrict__ grads,
                                                                   const T* __restrict__ invs,
                                                                   T* __restrict__ outs, int64_t n) {
  __shared__ __align__(16) T sg[kMinvThreads * 9];
  __shared__ __align__(16) T si[kMinvThreads * 9];
  for (int64_t blk = blockIdx.x; blk * kMinvThreads < n; blk += gridDim.x) {
    int64_t m0 = blk * kMinvThreads;
    int64_t cnt = min((int64_t)kMinvThreads, n - m0);
    stage_in(grads, sg, m0 * 9, cnt * 9);
    stage_in(invs, si, m0 * 9, cnt * 9);
    __syncthreads();
    T g[9], c[9], t[9], o[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { g[i] = sg[threadIdx.x * 9 + i]; c[i] = si[threadIdx.x * 9 + i]; }
    // t = g * inv^T  : t[r][k] = sum_q g[r][q] c[k][q]
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        t[3 * r + k] = g[3 * r + 0] * c[3 * k + 0] + g[3 * r + 1] * c[3 * k + 1] + g[3 * r + 2] * c[3 * k + 2];
    // o = -inv^T * t : o[i][k] = -sum_r c[r][i] t[r][k]
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        o[3 * i + k] = -(c[0 + i] * t[0 + k] + c[3 + i] * t[3 + k] + c[6 + i] * t[6 + k]);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 9; ++i) sg[threadIdx.x * 9 + i] = o[i];
    __syncthreads();
    stage_out(outs, sg, m0 * 9, cnt * 9);
    __syncthreads();
  }
}

}  // namespace recmv

using namespace recmv;

extern "C" int recmv_minv3x3_fwd(const void* ms, void* invs, uint8_t* ok, int64_t n, int dtype,
                                 recmv_stream_t stream) {
  if (n < 0) return RECMV_E_SHAPE;
  if (n == 0) return RECMV_OK;
  if (!ms || !invs || !ok) return RECMV_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = stride_grid(n, kMinvThreads, 8);
  if (dtype == RECMV_F32)
    minv3x3_fwd_kernel<float><<<grid, kMinvThreads, 0, st>>>((const float*)ms, (float*)invs, ok, n);
  else if (dtype == RECMV_F64)
    minv3x3_fwd_kernel<double><<<grid, kMinvThreads, 0, st>>>((const double*)ms, (double*)invs, ok, n);
  else
    return RECMV_E_DTYPE;
  return launch_status();
}

extern "C" int recmv_minv3x3_bwd(const void* grads, const void* invs, void* outs, int64_t n,
                                 int dtype, recmv_stream_t stream) {
  if (n < 0) return RECMV_E_SHAPE;
  if (n == 0) return RECMV_OK;
  if (!grads || !invs || !outs) return RECMV_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = stride_grid(n, kMinvThreads, 4);
  if (dtype == RECMV_F32)
    minv3x3_bwd_kernel<float><<<grid, kMinvThreads, 0, st>>>((const float*)grads, (const float*)invs, (float*)outs, n);
  else if (dtype == RECMV_F64)
    minv3x3_bwd_kernel<double><<<grid, kMinvThreads, 0, st>>>((const double*)grads, (const double*)invs, (double*)outs, n);
  else
    return RECMV_E_DTYPE;
  return launch_status();
}

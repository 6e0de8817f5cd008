// Marching cubes with shared vertices -- deterministic count / scan / emit.
// Replaces MCGpu/CudaKernels.cu:316-521 (d_mc_get_mesh_on_gpu, d_conver_ijkd_to_pindex, d_set_int,
// d_scale_vertices) and the MCGpu singleton (CudaKernels.cu:524-639).
//
// Reference cost per call: memset of 3*N ints (-1), case tables in global memory, two global atomics
// per triangle, output buffers sized to 5 % of the cells with no overflow check, blocking D2H copy.
// Here: HBM-bound sweep.  Algorithmic bytes = 4 B x N (one read of the grid) + 12 B x V + 24 B x F.
//   pass 1 (count): one read of the grid -> 1-byte case index per voxel + per-CTA (verts, tris) sums
//   pass 2 (scan) : exclusive scan of the per-CTA sums (one CTA), totals -> host
//   pass 3 (verts): case bytes -> CTA-local exclusive scan -> owned-edge vertices (v*step+origin) and a
//                   packed (first vertex id, rank of edges 0/3/8) word for ACTIVE cells only (no memset:
//                   only entries that were written are ever read)
//   pass 4 (faces): case bytes -> scan -> int64 faces, neighbours' vertex ids from the packed words
// Output order = the order a sequential sweep of the reference kernel would produce (cells by linear
// index, vertices by first appearance in the cell's triangle list), so results are reproducible and
// comparable index-for-index with the CPU restatement (oracle/mc_oracle.c).
#include "common.cuh"
#include "mc_tables.h"

namespace recmv {

constexpr int kMcThreads = 1024;

// Division of a 31-bit index by a runtime constant through a precomputed multiplier (Granlund-Montgomery):
// the per-cell (i,j,k) decomposition otherwise costs two ~20-instruction integer divisions per cell.
struct FastDiv {
  unsigned d, m, s;
  __host__ static FastDiv make(unsigned d) {
    FastDiv f;
    f.d = d;
    unsigned s = 0;
    while ((1ull << s) < d) ++s;
    f.s = s;
    f.m = (unsigned)(((1ull << 32) * ((1ull << s) - d)) / d + 1);
    return f;
  }
  __device__ __forceinline__ unsigned div(unsigned n) const {
    unsigned t = __umulhi(m, n);
    return (t + n) >> s;  // n < 2^31 and t <= n: no overflow
  }
};

struct McTables {
  signed char tri[256][16];
  unsigned char ntri[256];
  unsigned char vinfo[256];  // bits 0-1 rank(edge0), 2-3 rank(edge3), 4-5 rank(edge8) (3 = absent), 6-7 count
  signed char owner[12][4];  // owning voxel offset + axis of every cube edge
};
__constant__ McTables c_mc;
// copy of the triangle table in global memory: the face kernel indexes it with per-thread case numbers, and
// divergent __constant__ reads serialise 32-way (L1-cached global loads do not)
__device__ signed char g_mc_tri[256][16];

static void build_tables(McTables& t) {
  for (int c = 0; c < 256; ++c) {
    int n = 0;
    while (n < 5 && kMcTriTable[c][3 * n] >= 0) ++n;
    t.ntri[c] = (unsigned char)n;
    int rank[3] = {3, 3, 3}, cnt = 0;
    for (int q = 0; q < 3 * n; ++q) {
      int e = kMcTriTable[c][q];
      int slot = e == 0 ? 0 : (e == 3 ? 1 : (e == 8 ? 2 : -1));
      if (slot >= 0 && rank[slot] == 3) rank[slot] = cnt++;
    }
    t.vinfo[c] = (unsigned char)(rank[0] | (rank[1] << 2) | (rank[2] << 4) | (cnt << 6));
    for (int q = 0; q < 16; ++q) t.tri[c][q] = kMcTriTable[c][q];
  }
  for (int e = 0; e < 12; ++e)
    for (int q = 0; q < 4; ++q) t.owner[e][q] = kMcEdgeOwner[e][q];
}

static int ensure_tables() {
  static int dev_done[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return RECMV_E_RANGE;
  if (!dev_done[dev]) {
    static McTables host;
    build_tables(host);
    cudaError_t e = cudaMemcpyToSymbol(c_mc, &host, sizeof(McTables));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(g_mc_tri, host.tri, sizeof(host.tri));
    if (e != cudaSuccess) return (int)e;
    dev_done[dev] = 1;
  }
  return RECMV_OK;
}

// CTA-wide exclusive scan of an int2 (x = verts, y = tris); returns exclusive prefix, total in *tot.
__device__ __forceinline__ int2 block_excl_scan(int2 v, int2* tot) {
  __shared__ int2 warp_sums[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int2 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int ax = __shfl_up_sync(0xffffffffu, inc.x, o);
    int ay = __shfl_up_sync(0xffffffffu, inc.y, o);
    if (lane >= o) { inc.x += ax; inc.y += ay; }
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int2 w = warp_sums[lane];
    int2 wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int ax = __shfl_up_sync(0xffffffffu, wi.x, o);
      int ay = __shfl_up_sync(0xffffffffu, wi.y, o);
      if (lane >= o) { wi.x += ax; wi.y += ay; }
    }
    warp_sums[lane] = make_int2(wi.x - w.x, wi.y - w.y);  // exclusive
    if (lane == 31) *tot = wi;
  }
  __syncthreads();
  int2 base = warp_sums[wid];
  int2 r = make_int2(base.x + inc.x - v.x, base.y + inc.y - v.y);
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kMcThreads) mc_count_kernel(const float* __restrict__ sdf, int NX,
                                                              int NY, int NZ, float iso,
                                                              unsigned char* __restrict__ cube,
                                                              int2* __restrict__ block_sums, int nseg,
                                                              FastDiv dplane, FastDiv dnz) {
  __shared__ int s_cnt[2];
  const unsigned N = (unsigned)NX * NY * NZ, plane = (unsigned)NY * NZ;
 for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {  // one 1024-cell segment per iteration
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const unsigned idx = (unsigned)seg * kMcThreads + threadIdx.x;
  int nv = 0, nt = 0;
  if (idx < N) {
    const unsigned i = dplane.div(idx), rem = idx - i * plane;
    const unsigned j = dnz.div(rem), k = rem - j * NZ;
    int ci = 0;
    if (i < (unsigned)NX - 1 && j < (unsigned)NY - 1 && k < (unsigned)NZ - 1) {
      const float* p = sdf + idx;
      const size_t sj = NZ, si = plane;
      float v0 = __ldg(p), v1 = __ldg(p + si), v2 = __ldg(p + si + sj), v3 = __ldg(p + sj);
      float v4 = __ldg(p + 1), v5 = __ldg(p + si + 1), v6 = __ldg(p + si + sj + 1), v7 = __ldg(p + sj + 1);
      ci = (v0 < iso) | ((v1 < iso) << 1) | ((v2 < iso) << 2) | ((v3 < iso) << 3) | ((v4 < iso) << 4) |
           ((v5 < iso) << 5) | ((v6 < iso) << 6) | ((v7 < iso) << 7);
      if (ci != 0 && ci != 255) { nv = c_mc.vinfo[ci] >> 6; nt = c_mc.ntri[ci]; }
    }
    cube[idx] = (unsigned char)ci;
  }
  // only the per-CTA totals are needed here: warp reduction, one shared atomic per non-empty warp
  const int packed = __reduce_add_sync(0xffffffffu, nv | (nt << 16));
  if ((threadIdx.x & 31) == 0 && packed) {
    atomicAdd(&s_cnt[0], packed & 0xffff);
    atomicAdd(&s_cnt[1], packed >> 16);
  }
  __syncthreads();
  if (threadIdx.x == 0) block_sums[seg] = make_int2(s_cnt[0], s_cnt[1]);
  __syncthreads();
 }
}

__global__ void __launch_bounds__(kMcThreads) mc_scan_kernel(int2* __restrict__ block_sums, int nb,
                                                             int* __restrict__ totals) {
  __shared__ int2 tot;
  int2 carry = make_int2(0, 0);
  for (int base = 0; base < nb; base += kMcThreads) {
    int b = base + threadIdx.x;
    int2 v = b < nb ? block_sums[b] : make_int2(0, 0);
    int2 ex = block_excl_scan(v, &tot);
    if (b < nb) block_sums[b] = make_int2(ex.x + carry.x, ex.y + carry.y);
    carry.x += tot.x; carry.y += tot.y;
    __syncthreads();
  }
  if (threadIdx.x == 0) { totals[0] = carry.x; totals[1] = carry.y; }
}

// (iso - v1) / (v2 - v1) evaluated like d_fGetOffset (CudaKernels.cu:304-314): float differences,
// double division, rounded to float; 0.5 when the edge is flat.
__device__ __forceinline__ float edge_offset(float v1, float v2, float iso) {
  double d = (double)(v2 - v1);
  if (d == 0.0) return 0.5f;
  return (float)((double)(iso - v1) / d);
}

__global__ void __launch_bounds__(kMcThreads) mc_vertex_kernel(
    const float* __restrict__ sdf, int NX, int NY, int NZ, float iso,
    const unsigned char* __restrict__ cube, const int2* __restrict__ block_offs,
    int* __restrict__ cellinfo, float sx, float sy, float sz, float ox, float oy, float oz,
    float* __restrict__ verts, int nseg, FastDiv dplane, FastDiv dnz) {
  __shared__ int2 tot;
  const unsigned N = (unsigned)NX * NY * NZ, plane = (unsigned)NY * NZ;
 for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
  const unsigned idx = (unsigned)seg * kMcThreads + threadIdx.x;
  int ci = idx < N ? cube[idx] : 0;
  int info = (ci != 0 && ci != 255) ? c_mc.vinfo[ci] : (3 | (3 << 2) | (3 << 4));
  int nv = info >> 6;
  if (!__syncthreads_or(nv > 0)) continue;  // ~98 % of the segments hold no surface cell
  int2 ex = block_excl_scan(make_int2(nv, 0), &tot);
  if (nv == 0) continue;
  int vbase = block_offs[seg].x + ex.x;
  cellinfo[idx] = (vbase << 6) | (info & 63);
  const unsigned i = dplane.div(idx), rem = idx - i * plane;
  const unsigned j = dnz.div(rem), k = rem - j * NZ;
  const float* p = sdf + idx;
  float v0 = __ldg(p);
  float fX = (float)i, fY = (float)j, fZ = (float)k;
  int r0 = info & 3, r3 = (info >> 2) & 3, r8 = (info >> 4) & 3;
  if (r0 != 3) {  // edge 0: corner 0 -> 1, +x
    float off = edge_offset(v0, __ldg(p + (size_t)NY * NZ), iso);
    float* o = verts + (size_t)(vbase + r0) * 3;
    o[0] = fmaf(fX + off, sx, ox); o[1] = fmaf(fY, sy, oy); o[2] = fmaf(fZ, sz, oz);
  }
  if (r3 != 3) {  // edge 3: corner 3 (0,1,0) -> 0, -y
    float off = edge_offset(__ldg(p + NZ), v0, iso);
    float* o = verts + (size_t)(vbase + r3) * 3;
    o[0] = fmaf(fX, sx, ox); o[1] = fmaf(fY + (1.f - off), sy, oy); o[2] = fmaf(fZ, sz, oz);
  }
  if (r8 != 3) {  // edge 8: corner 0 -> 4, +z
    float off = edge_offset(v0, __ldg(p + 1), iso);
    float* o = verts + (size_t)(vbase + r8) * 3;
    o[0] = fmaf(fX, sx, ox); o[1] = fmaf(fY, sy, oy); o[2] = fmaf(fZ + off, sz, oz);
  }
 }
}

__global__ void __launch_bounds__(kMcThreads) mc_face_kernel(int NX, int NY, int NZ,
                                                             const unsigned char* __restrict__ cube,
                                                             const int2* __restrict__ block_offs,
                                                             const int* __restrict__ cellinfo,
                                                             long long* __restrict__ faces, int nseg,
                                                             FastDiv dplane, FastDiv dnz) {
  __shared__ int2 tot;
  const unsigned N = (unsigned)NX * NY * NZ, plane = (unsigned)NY * NZ;
 for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
  const unsigned idx = (unsigned)seg * kMcThreads + threadIdx.x;
  int ci = idx < N ? cube[idx] : 0;
  int nt = (ci != 0 && ci != 255) ? c_mc.ntri[ci] : 0;
  if (!__syncthreads_or(nt > 0)) continue;
  int2 ex = block_excl_scan(make_int2(0, nt), &tot);
  if (nt == 0) continue;
  int64_t fbase = (int64_t)block_offs[seg].y + ex.y;
  const int i = (int)dplane.div(idx), rem = (int)(idx - (unsigned)i * plane);
  const int j = (int)dnz.div((unsigned)rem), k = rem - j * NZ;
  for (int t = 0; t < nt; ++t) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int e = g_mc_tri[ci][3 * t + c];
      int oi = i + c_mc.owner[e][0], oj = j + c_mc.owner[e][1], ok = k + c_mc.owner[e][2], d = c_mc.owner[e][3];
      long long vid = -1;
      if (oi < NX - 1 && oj < NY - 1 && ok < NZ - 1) {
        int info = cellinfo[((int64_t)oi * NY + oj) * NZ + ok];
        vid = (long long)(info >> 6) + ((info >> (2 * d)) & 3);
      }
      faces[(fbase + t) * 3 + (2 - c)] = vid;  // reversed winding (CudaKernels.cu:502)
    }
  }
 }
}

struct McScratch {
  unsigned char* cube;
  int* cellinfo;
  int2* block_sums;
  int* totals;
  int nb;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static McScratch carve(void* scratch, int NX, int NY, int NZ) {
  int64_t N = (int64_t)NX * NY * NZ;
  McScratch s;
  s.nb = (int)((N + kMcThreads - 1) / kMcThreads);
  char* p = (char*)scratch;
  s.cube = (unsigned char*)p; p += align256((size_t)N);
  s.cellinfo = (int*)p; p += align256((size_t)N * 4);
  s.block_sums = (int2*)p; p += align256((size_t)s.nb * 8);
  s.totals = (int*)p;
  return s;
}

}  // namespace recmv

using namespace recmv;

extern "C" int recmv_mc_scratch_bytes(int NX, int NY, int NZ, size_t* bytes) {
  if (!bytes) return RECMV_E_NULL;
  if (NX <= 0 || NY <= 0 || NZ <= 0) return RECMV_E_SHAPE;
  int64_t N = (int64_t)NX * NY * NZ;
  if (N > 2000000000LL) return RECMV_E_RANGE;
  int nb = (int)((N + kMcThreads - 1) / kMcThreads);
  *bytes = align256((size_t)N) + align256((size_t)N * 4) + align256((size_t)nb * 8) + 256;
  return RECMV_OK;
}

extern "C" int recmv_mc_count(const float* sdf, int NX, int NY, int NZ, float iso, void* scratch,
                              int64_t* num_verts, int64_t* num_faces, recmv_stream_t stream) {
  if (!sdf || !scratch || !num_verts || !num_faces) return RECMV_E_NULL;
  if (NX <= 0 || NY <= 0 || NZ <= 0) return RECMV_E_SHAPE;
  if ((int64_t)NX * NY * NZ > 2000000000LL) return RECMV_E_RANGE;
  int s = ensure_tables();
  if (s) return s;
  cudaStream_t st = (cudaStream_t)stream;
  McScratch sc = carve(scratch, NX, NY, NZ);
  const int grid = sc.nb < 2 * num_sms() ? sc.nb : 2 * num_sms();
  const FastDiv dplane = FastDiv::make((unsigned)NY * NZ), dnz = FastDiv::make((unsigned)NZ);
  mc_count_kernel<<<grid, kMcThreads, 0, st>>>(sdf, NX, NY, NZ, iso, sc.cube, sc.block_sums, sc.nb, dplane, dnz);
  s = launch_status();
  if (s) return s;
  mc_scan_kernel<<<1, kMcThreads, 0, st>>>(sc.block_sums, sc.nb, sc.totals);
  s = launch_status();
  if (s) return s;
  int h[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(h, sc.totals, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return (int)e;
  *num_verts = h[0];
  *num_faces = h[1];
  if (h[0] >= (1 << 25)) return RECMV_E_RANGE;  // packed vertex ids are 26 bit signed-safe
  return RECMV_OK;
}

extern "C" int recmv_mc_emit(const float* sdf, int NX, int NY, int NZ, float iso, void* scratch,
                             const float step[3], const float origin[3], float* verts,
                             int64_t* faces, recmv_stream_t stream) {
  if (!sdf || !scratch || !step || !origin) return RECMV_E_NULL;
  if (NX <= 0 || NY <= 0 || NZ <= 0) return RECMV_E_SHAPE;
  int s = ensure_tables();
  if (s) return s;
  cudaStream_t st = (cudaStream_t)stream;
  McScratch sc = carve(scratch, NX, NY, NZ);
  const int grid = sc.nb < 2 * num_sms() ? sc.nb : 2 * num_sms();
  const FastDiv dplane = FastDiv::make((unsigned)NY * NZ), dnz = FastDiv::make((unsigned)NZ);
  if (verts) {
    mc_vertex_kernel<<<grid, kMcThreads, 0, st>>>(sdf, NX, NY, NZ, iso, sc.cube, sc.block_sums,
                                                  sc.cellinfo, step[0], step[1], step[2], origin[0],
                                                  origin[1], origin[2], verts, sc.nb, dplane, dnz);
    s = launch_status();
    if (s) return s;
  }
  if (faces) {
    if (!verts) return RECMV_E_NULL;  // faces need the packed words written by the vertex pass
    mc_face_kernel<<<grid, kMcThreads, 0, st>>>(NX, NY, NZ, sc.cube, sc.block_sums, sc.cellinfo,
                                                (long long*)faces, sc.nb, dplane, dnz);
    s = launch_status();
    if (s) return s;
  }
  return RECMV_OK;
}

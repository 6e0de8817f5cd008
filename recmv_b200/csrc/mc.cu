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
    while (n < 5 && mc_tri_entry(c, 3 * n) >= 0) ++n;
    t.ntri[c] = (unsigned char)n;
    int rank[3] = {3, 3, 3}, cnt = 0;
    for (int q = 0; q < 3 * n; ++q) {
      int e = mc_tri_entry(c, q);
      int slot = e == 0 ? 0 : (e == 3 ? 1 : (e == 8 ? 2 : -1));
      if (slot >= 0 && rank[slot] == 3) rank[slot] = cnt++;
    }
    t.vinfo[c] = (unsigned char)(rank[0] | (rank[1] << 2) | (rank[2] << 4) | (cnt << 6));
    for (int q = 0; q < 16; ++q) t.tri[c][q] = (signed char)mc_tri_entry(c, q);
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

constexpr int kMcWarpBlock = 256;  // 8 warps per CTA in the warp-per-segment passes

// Pass 1, one WARP per 1024-cell segment, 32 rows of 32 consecutive cells: 4 coalesced loads per cell (its
// four z-columns at k); the k+1 values come from the next lane by shuffle (lane 31 loads them).  No block
// barrier, no shared memory.  Writes the 1-byte case index of every voxel and the segment's (verts, tris).
__global__ void __launch_bounds__(kMcWarpBlock) mc_count_kernel(const float* __restrict__ sdf, int NX,
                                                                int NY, int NZ, float iso,
                                                                unsigned char* __restrict__ cube,
                                                                int2* __restrict__ block_sums, int nseg,
                                                                FastDiv dplane, FastDiv dnz) {
  const unsigned N = (unsigned)NX * NY * NZ, plane = (unsigned)NY * NZ;
  const int lane = threadIdx.x & 31;
  const int seg = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (seg >= nseg) return;
  int acc = 0;  // verts | tris << 16 of this lane
  const size_t sj = NZ, si = plane;
#pragma unroll 4
  for (int r = 0; r < 32; ++r) {
    const unsigned idx = (unsigned)seg * kMcThreads + r * 32 + lane;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // columns (i,j) (i+1,j) (i+1,j+1) (i,j+1) at k
    bool cell = false;
    unsigned i = 0, j = 0, k = 0;
    if (idx < N) {
      i = dplane.div(idx);
      const unsigned rem = idx - i * plane;
      j = dnz.div(rem);
      k = rem - j * NZ;
      const bool col = i < (unsigned)NX - 1 && j < (unsigned)NY - 1;
      if (col) {
        const float* p = sdf + idx;
        a0 = __ldg(p); a1 = __ldg(p + si); a2 = __ldg(p + si + sj); a3 = __ldg(p + sj);
      }
      cell = col && k < (unsigned)NZ - 1;
    }
    // k+1 values: lane+1 holds idx+1 (same column unless the row wraps, and then this cell is not a cell)
    float b0 = __shfl_down_sync(0xffffffffu, a0, 1), b1 = __shfl_down_sync(0xffffffffu, a1, 1);
    float b2 = __shfl_down_sync(0xffffffffu, a2, 1), b3 = __shfl_down_sync(0xffffffffu, a3, 1);
    if (lane == 31 && cell) {
      const float* p = sdf + idx + 1;
      b0 = __ldg(p); b1 = __ldg(p + si); b2 = __ldg(p + si + sj); b3 = __ldg(p + sj);
    }
    int ci = 0;
    if (cell) {
      ci = (a0 < iso) | ((a1 < iso) << 1) | ((a2 < iso) << 2) | ((a3 < iso) << 3) | ((b0 < iso) << 4) |
           ((b1 < iso) << 5) | ((b2 < iso) << 6) | ((b3 < iso) << 7);
      if (ci != 0 && ci != 255) acc += (c_mc.vinfo[ci] >> 6) | ((int)c_mc.ntri[ci] << 16);
    }
    if (idx < N) cube[idx] = (unsigned char)ci;
  }
  const int tot = __reduce_add_sync(0xffffffffu, acc);
  if (lane == 0) block_sums[seg] = make_int2(tot & 0xffff, tot >> 16);
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

// warp-wide exclusive scan
__device__ __forceinline__ int warp_excl_scan(int v, int lane, int* total) {
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  *total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

// Passes 3 and 4, one warp per 1024-cell segment, 8 rows of 128 cells: each lane loads the case bytes of 4
// consecutive cells as one 32-bit word; a row whose words are all 0x00000000 / 0xffffffff (no surface: ~98 %
// of them) costs one load and one ballot.
template <bool kFaces>
__global__ void __launch_bounds__(kMcWarpBlock) mc_emit_kernel(
    const float* __restrict__ sdf, int NX, int NY, int NZ, float iso,
    const unsigned char* __restrict__ cube, const int2* __restrict__ block_offs,
    int* __restrict__ cellinfo, float sx, float sy, float sz, float ox, float oy, float oz,
    float* __restrict__ verts, long long* __restrict__ faces, int nseg, FastDiv dplane, FastDiv dnz) {
  const unsigned N = (unsigned)NX * NY * NZ, plane = (unsigned)NY * NZ;
  const int lane = threadIdx.x & 31;
  const int seg = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (seg >= nseg) return;
  const int2 segoff = block_offs[seg];
  if (!kFaces) { if (seg + 1 < nseg && block_offs[seg + 1].x == segoff.x) return; }   // nothing to emit here
  else { if (seg + 1 < nseg && block_offs[seg + 1].y == segoff.y) return; }
  int running = kFaces ? segoff.y : segoff.x;
  for (int r = 0; r < 8; ++r) {
    const unsigned idx0 = (unsigned)seg * kMcThreads + r * 128 + lane * 4;
    unsigned word = 0;
    if (idx0 + 3 < N) word = *reinterpret_cast<const unsigned*>(cube + idx0);   // N*1 bytes; idx0 % 4 == 0
    else { for (int q = 0; q < 4; ++q) if (idx0 + q < N) word |= (unsigned)cube[idx0 + q] << (8 * q); }
    if (__ballot_sync(0xffffffffu, word != 0u && word != 0xffffffffu) == 0u) continue;
    int cnt[4], mine = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ci = (word >> (8 * q)) & 255;
      cnt[q] = (ci != 0 && ci != 255) ? (kFaces ? (int)c_mc.ntri[ci] : (int)(c_mc.vinfo[ci] >> 6)) : 0;
      mine += cnt[q];
    }
    int row_total;
    int off = running + warp_excl_scan(mine, lane, &row_total);
    running += row_total;
    if (mine == 0) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (cnt[q] == 0) continue;
      const unsigned idx = idx0 + q;
      const int ci = (word >> (8 * q)) & 255;
      const int i = (int)dplane.div(idx), rem = (int)(idx - (unsigned)i * plane);
      const int j = (int)dnz.div((unsigned)rem), k = rem - j * NZ;
      if (!kFaces) {
        const int info = c_mc.vinfo[ci];
        const int vbase = off;
        cellinfo[idx] = (vbase << 6) | (info & 63);
        const float* p = sdf + idx;
        const float v0 = __ldg(p);
        const float fX = (float)i, fY = (float)j, fZ = (float)k;
        const int r0 = info & 3, r3 = (info >> 2) & 3, r8 = (info >> 4) & 3;
        if (r0 != 3) {  // edge 0: corner 0 -> 1, +x
          const float o_ = edge_offset(v0, __ldg(p + plane), iso);
          float* o = verts + (size_t)(vbase + r0) * 3;
          o[0] = fmaf(fX + o_, sx, ox); o[1] = fmaf(fY, sy, oy); o[2] = fmaf(fZ, sz, oz);
        }
        if (r3 != 3) {  // edge 3: corner 3 (0,1,0) -> 0, -y
          const float o_ = edge_offset(__ldg(p + NZ), v0, iso);
          float* o = verts + (size_t)(vbase + r3) * 3;
          o[0] = fmaf(fX, sx, ox); o[1] = fmaf(fY + (1.f - o_), sy, oy); o[2] = fmaf(fZ, sz, oz);
        }
        if (r8 != 3) {  // edge 8: corner 0 -> 4, +z
          const float o_ = edge_offset(v0, __ldg(p + 1), iso);
          float* o = verts + (size_t)(vbase + r8) * 3;
          o[0] = fmaf(fX, sx, ox); o[1] = fmaf(fY, sy, oy); o[2] = fmaf(fZ + o_, sz, oz);
        }
      } else {
        const long long fbase = off;
        for (int t = 0; t < cnt[q]; ++t) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int e = g_mc_tri[ci][3 * t + c];
            const int oi = i + c_mc.owner[e][0], oj = j + c_mc.owner[e][1], ok = k + c_mc.owner[e][2];
            const int d = c_mc.owner[e][3];
            long long vid = -1;
            if (oi < NX - 1 && oj < NY - 1 && ok < NZ - 1) {
              const int info = cellinfo[((int64_t)oi * NY + oj) * NZ + ok];
              vid = (long long)(info >> 6) + ((info >> (2 * d)) & 3);
            }
            faces[(fbase + t) * 3 + (2 - c)] = vid;  // reversed winding (CudaKernels.cu:502)
          }
        }
      }
      off += cnt[q];
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
  const int grid = (sc.nb * 32 + kMcWarpBlock - 1) / kMcWarpBlock;  // one warp per 1024-cell segment
  const FastDiv dplane = FastDiv::make((unsigned)NY * NZ), dnz = FastDiv::make((unsigned)NZ);
  mc_count_kernel<<<grid, kMcWarpBlock, 0, st>>>(sdf, NX, NY, NZ, iso, sc.cube, sc.block_sums, sc.nb, dplane, dnz);
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
  const int grid = (sc.nb * 32 + kMcWarpBlock - 1) / kMcWarpBlock;
  const FastDiv dplane = FastDiv::make((unsigned)NY * NZ), dnz = FastDiv::make((unsigned)NZ);
  if (verts) {
    mc_emit_kernel<false><<<grid, kMcWarpBlock, 0, st>>>(sdf, NX, NY, NZ, iso, sc.cube, sc.block_sums, sc.cellinfo,
                                                         step[0], step[1], step[2], origin[0], origin[1], origin[2],
                                                         verts, nullptr, sc.nb, dplane, dnz);
    s = launch_status();
    if (s) return s;
  }
  if (faces) {
    if (!verts) return RECMV_E_NULL;  // faces need the packed words written by the vertex pass
    mc_emit_kernel<true><<<grid, kMcWarpBlock, 0, st>>>(sdf, NX, NY, NZ, iso, sc.cube, sc.block_sums, sc.cellinfo,
                                                        0.f, 0.f, 0.f, 0.f, 0.f, 0.f, nullptr, (long long*)faces,
                                                        sc.nb, dplane, dnz);
    s = launch_status();
    if (s) return s;
  }
  return RECMV_OK;
}

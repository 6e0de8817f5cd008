/* TEST INFRASTRUCTURE ONLY -- sequential CPU restatement of the reference's marching cubes
 * (MCGpu/CudaKernels.cu:316-521): d_mc_get_mesh_on_gpu run as a single thread over the cell index in
 * ascending order, then d_conver_ijkd_to_pindex and d_scale_vertices.  The reference's output order
 * depends on atomic arrival; a sequential sweep is one of its legal schedules, and it is the canonical
 * order the CUDA implementation reproduces.  Parity status: restatement checked (a) against the
 * reference's tables (oracle/check_mc_tables.py) and (b) on the GPU box against nothing else -- the
 * reference has no CPU marching cubes; see DESIGN.md "oracle pins".
 *
 *   int mc_oracle(const float* sdf, int NX, int NY, int NZ, float iso,
 *                 const float step[3], const float origin[3],
 *                 float* verts, long long* faces, long long cap_v, long long cap_f,
 *                 long long* nv, long long* nf)
 * returns 0, or 1 when a capacity is too small (nv/nf still hold the required sizes).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../recmv_b200/csrc/mc_tables.h"

static float get_offset(float v1, float v2, float want) { /* CudaKernels.cu:304-314 */
  double d = (double)(v2 - v1);
  if (d == 0.0) return 0.5f;
  return (float)((double)(want - v1) / d);
}

int mc_oracle(const float* sdf, int NX, int NY, int NZ, float iso, const float step[3],
              const float origin[3], float* verts, long long* faces, long long cap_v,
              long long cap_f, long long* nv_out, long long* nf_out) {
  long long N = (long long)NX * NY * NZ;
  int* edge_state = (int*)malloc(sizeof(int) * (size_t)N * 3); /* d_edge_point_state_, -1 filled */
  if (!edge_state) return 2;
  for (long long q = 0; q < N * 3; ++q) edge_state[q] = -1;
  long long nv = 0, nf = 0;
  int overflow = 0;
  /* pass 1: vertices + (i,j,k,dir) face records, cell index ascending */
  long long cap_rec = cap_f > 0 ? cap_f : 1;
  int* ijkd = (int*)malloc(sizeof(int) * 12 * (size_t)cap_rec);
  if (!ijkd) { free(edge_state); return 2; }
  for (long long idx = 0; idx < N; ++idx) {
    int i = (int)(idx / ((long long)NY * NZ));
    int j = (int)((idx - (long long)i * NY * NZ) / NZ);
    int k = (int)(idx - (long long)i * NY * NZ - (long long)j * NZ);
    if (!(i < NX - 1 && j < NY - 1 && k < NZ - 1)) continue;
    float val[8];
    int flag = 0;
    for (int c = 0; c < 8; ++c) {
      long long id = (long long)(i + kMcCornerOffset[c][0]) * NY * NZ +
                     (long long)(j + kMcCornerOffset[c][1]) * NZ + (k + kMcCornerOffset[c][2]);
      val[c] = sdf[id];
      if (val[c] < iso) flag |= 1 << c;
    }
    if (mc_tri_entry(flag, 0) < 0) continue;
    int is_new[12];
    for (int e = 0; e < 12; ++e) is_new[e] = 1;
    for (int t = 0; t < 5 && mc_tri_entry(flag, 3 * t) >= 0; ++t) {
      long long fid = nf++;
      for (int c = 0; c < 3; ++c) {
        int e = mc_tri_entry(flag, 3 * t + c);
        int bx = i + kMcEdgeOwner[e][0], by = j + kMcEdgeOwner[e][1], bz = k + kMcEdgeOwner[e][2];
        int dir = kMcEdgeOwner[e][3];
        if (is_new[e] && (e == 0 || e == 3 || e == 8)) {
          int c0 = kMcEdgeCorners[e][0], c1 = kMcEdgeCorners[e][1];
          float off = get_offset(val[c0], val[c1], iso);
          float pos[3];
          for (int a = 0; a < 3; ++a) {
            float d = (float)(kMcCornerOffset[c1][a] - kMcCornerOffset[c0][a]); /* edge direction */
            float base = (a == 0 ? (float)i : (a == 1 ? (float)j : (float)k));
            pos[a] = base + ((float)kMcCornerOffset[c0][a] + off * d);
          }
          long long vid = nv++;
          if (vid < cap_v) {
            for (int a = 0; a < 3; ++a) verts[vid * 3 + a] = fmaf(pos[a], step[a], origin[a]);
          } else {
            overflow = 1;
          }
          edge_state[((long long)bx * NY * NZ + (long long)by * NZ + bz) * 3 + dir] = (int)vid;
          is_new[e] = 0;
        }
        if (fid < cap_f) {
          ijkd[fid * 12 + c * 4 + 0] = bx; ijkd[fid * 12 + c * 4 + 1] = by;
          ijkd[fid * 12 + c * 4 + 2] = bz; ijkd[fid * 12 + c * 4 + 3] = dir;
        } else {
          overflow = 1;
        }
      }
    }
  }
  /* pass 2: (i,j,k,dir) -> vertex id, winding reversed (CudaKernels.cu:492-505) */
  long long lim = nf < cap_f ? nf : cap_f;
  for (long long f = 0; f < lim; ++f)
    for (int p = 0; p < 3; ++p) {
      int bx = ijkd[f * 12 + p * 4], by = ijkd[f * 12 + p * 4 + 1], bz = ijkd[f * 12 + p * 4 + 2];
      int d = ijkd[f * 12 + p * 4 + 3];
      faces[f * 3 + (2 - p)] =
          (long long)edge_state[((long long)bx * NY * NZ + (long long)by * NZ + bz) * 3 + d];
    }
  free(edge_state);
  free(ijkd);
  *nv_out = nv;
  *nf_out = nf;
  return overflow;
}

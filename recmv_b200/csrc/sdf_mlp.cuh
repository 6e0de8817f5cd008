// Internal interface shared by the SIMT and tcgen05 implementations of the fused SDF path.
#pragma once
#include "common.cuh"

namespace recmv {

// Where a tile's canonical points come from.
//   x != nullptr : canonical points [P,3] read from memory (ImplicitNetwork.forward)
//   x == nullptr : render mode -- point p = ray r * S + k, x_obs = cam + t_k dir_r, inverse LBS
struct PointSource {
  const float* x;
  // render mode
  const float* ray_dirs;
  const float* A;
  const float* trans;
  const int32_t* frame_of_ray;
  int64_t rays_per_frame;
  int num_frames;
  Voxel vox;
  float cam[3];
  float t_near, dt;  // t_k = t_near + (k + 0.5) * dt
  int S;
  float* out_xc;  // optional [P,3]
};

// 12 annealing weights, passed by value to kernels
struct PeWeights {
  float w[12];
};

// Canonical point of global sample index p; returns false when the inverse warp is singular.
__device__ __forceinline__ bool fetch_point(const PointSource& src, int64_t p, float& cx, float& cy,
                                            float& cz) {
  if (src.x) {
    cx = __ldg(src.x + 3 * p); cy = __ldg(src.x + 3 * p + 1); cz = __ldg(src.x + 3 * p + 2);
    return true;
  }
  int64_t r = p / src.S;
  int k = (int)(p - r * src.S);
  float t = src.t_near + ((float)k + 0.5f) * src.dt;
  float dx = __ldg(src.ray_dirs + 3 * r), dy = __ldg(src.ray_dirs + 3 * r + 1), dz = __ldg(src.ray_dirs + 3 * r + 2);
  float ox = src.cam[0] + t * dx, oy = src.cam[1] + t * dy, oz = src.cam[2] + t * dz;
  int64_t f = src.frame_of_ray ? (int64_t)__ldg(src.frame_of_ray + r)
                               : (src.rays_per_frame > 0 ? r / src.rays_per_frame : 0);
  if (f < 0) f = 0;
  if (f >= src.num_frames) f = src.num_frames - 1;
  bool ok = inverse_lbs_point(src.vox, src.A + (size_t)f * 384, src.trans + 3 * f, ox, oy, oz, cx, cy, cz);
  if (src.out_xc) { src.out_xc[3 * p] = cx; src.out_xc[3 * p + 1] = cy; src.out_xc[3 * p + 2] = cz; }
  return ok;
}

constexpr float kInvalidSdf = 1e10f;

// implemented in sdf_mlp_simt.cu / sdf_mlp_tc.cu
int simt_sdf_forward(const PointSource& src, const void* packed, const PeWeights& pw, float* out_sdf,
                     float* out_feat, int64_t P, cudaStream_t st);
int tc_sdf_forward(const PointSource& src, const void* packed, const PeWeights& pw, float* out_sdf,
                   float* out_feat, int64_t P, int passes, cudaStream_t st, const int* P_dev = nullptr);

// the per-device status record (mapped pinned host memory) the tcgen05 kernels report into: bounded-wait
// time-outs (code 1) and fp16 operand-range violations (code 2); see recmv_check_async_errors
int device_status_record(void** out);

// value + input gradient in one forward-mode launch (tcgen05 path only)
int tc_sdf_forward_grad(const float* x, const void* packed, const PeWeights& pw, float* out_sdf, float* out_feat,
                        float* out_grad, int64_t P, int passes, cudaStream_t st);

}  // namespace recmv

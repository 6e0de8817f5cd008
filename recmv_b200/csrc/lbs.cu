// Linear blend skinning on the diffused SMPL skinning voxel: forward warp and inverse warp.
// Replaces model/Deformer.py:406-445 (LBSkinner.forward after the bone matrices are built):
//   GridSamplerMine3dFunction (K4) -> matmul(ps_ws, A) -> matmul(T, [p;1]) -> + trans,
// including the per-frame python loop with a host sync per frame (Deformer.py:438-444).
//
// One thread per point: 8 corners x 24 channels gathered from the channels-last voxel as 48 128-bit
// loads (768 algorithmic bytes per point), 24x12 FMA bone blend with the bone matrices broadcast from
// L1/constant cache, 12 FMA apply; nothing but the 12-byte result goes back to HBM.
#include "common.cuh"

namespace recmv {

__device__ __forceinline__ int frame_of(const int64_t* __restrict__ batch_inds, int64_t p,
                                        int64_t ppf, int nf) {
  int64_t f = batch_inds ? batch_inds[p] : (ppf > 0 ? p / ppf : 0);
  if (f < 0) f = 0;
  if (f >= nf) f = nf - 1;
  return (int)f;
}

__global__ void __launch_bounds__(128) lbs_fwd_kernel(const float* __restrict__ ps,
                                                      const float* __restrict__ tps,
                                                      const float* __restrict__ A,
                                                      const float* __restrict__ trans,
                                                      const int64_t* __restrict__ batch_inds,
                                                      int64_t ppf, int nf, Voxel vox,
                                                      float* __restrict__ out,
                                                      float* __restrict__ weights_out, int64_t P) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P;
       p += (int64_t)gridDim.x * blockDim.x) {
    float x = ps[3 * p], y = ps[3 * p + 1], z = ps[3 * p + 2];
    float tx = x, ty = y, tz = z;
    if (tps) { tx = tps[3 * p]; ty = tps[3 * p + 1]; tz = tps[3 * p + 2]; }
    float w[24], T[12];
    sample_skin24(vox, tx, ty, tz, w);
    int f = frame_of(batch_inds, p, ppf, nf);
    blend_bones(A + (size_t)f * 24 * 16, w, T);
    const float* t = trans + 3 * f;
    // T [p;1] accumulated in the order of a 4-term dot product, then + trans (Deformer.py:430-431)
    out[3 * p + 0] = (T[0] * x + T[1] * y + T[2] * z + T[3]) + __ldg(t + 0);
    out[3 * p + 1] = (T[4] * x + T[5] * y + T[6] * z + T[7]) + __ldg(t + 1);
    out[3 * p + 2] = (T[8] * x + T[9] * y + T[10] * z + T[11]) + __ldg(t + 2);
    if (weights_out) {
#pragma unroll
      for (int q = 0; q < 6; ++q)
        reinterpret_cast<float4*>(weights_out + p * 24)[q] =
            make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
  }
}

__global__ void __launch_bounds__(128) lbs_inverse_kernel(const float* __restrict__ xo,
                                                          const float* __restrict__ A,
                                                          const float* __restrict__ trans,
                                                          const int64_t* __restrict__ batch_inds,
                                                          int64_t ppf, int nf, Voxel vox,
                                                          float* __restrict__ xc,
                                                          uint8_t* __restrict__ valid, int64_t P) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P;
       p += (int64_t)gridDim.x * blockDim.x) {
    int f = frame_of(batch_inds, p, ppf, nf);
    float cx, cy, cz;
    bool ok = inverse_lbs_point(vox, A + (size_t)f * 24 * 16, trans + 3 * f, xo[3 * p],
                                xo[3 * p + 1], xo[3 * p + 2], cx, cy, cz);
    xc[3 * p] = cx; xc[3 * p + 1] = cy; xc[3 * p + 2] = cz;
    if (valid) valid[p] = ok ? 1 : 0;
  }
}


// Bone matrices A = G . init_pose of the SMPL kinematic chain (model/Deformer.py:372-405): axis-angle -> rotation
// (the un-vendored smpl_pytorch.util.batch_rodrigues, standard HMR quaternion form: angle = |theta + 1e-8|), chain
// G_i = G_parent(i) . [R_i | J_i - J_parent(i)], then A_i = G_i . init_pose_i (or G_i - [0 | G_i J_i] without an init
// pose).  One thread per frame walks the 24 joints (parents precede children); replaces ~150 tiny torch launches
// per call of LBSkinner.forward / .inverse when no autograd graph is needed.
__global__ void __launch_bounds__(64) bone_matrices_kernel(const float* __restrict__ poses, const float* __restrict__ Js,
                                                           const int* __restrict__ parents, const float* __restrict__ init_pose,
                                                           float* __restrict__ G, float* __restrict__ A, int F) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  float* g = G + (size_t)f * 24 * 16;
  for (int i = 0; i < 24; ++i) {
    const float* th = poses + ((size_t)f * 24 + i) * 3;
    const float ax = th[0], ay = th[1], az = th[2];
    const float bx = ax + 1e-8f, by = ay + 1e-8f, bz = az + 1e-8f;
    const float angle = sqrtf(bx * bx + by * by + bz * bz);
    const float half = angle * 0.5f;
    float sn, cs;
    sincosf(half, &sn, &cs);
    float qw = cs, qx = sn * (ax / angle), qy = sn * (ay / angle), qz = sn * (az / angle);
    const float qn = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
    qw /= qn; qx /= qn; qy /= qn; qz /= qn;
    const float w2 = qw * qw, x2 = qx * qx, y2 = qy * qy, z2 = qz * qz;
    const float wx = qw * qx, wy = qw * qy, wz = qw * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz;
    float L[16];   // local transform [R | t; 0 0 0 1]
    L[0] = w2 + x2 - y2 - z2; L[1] = 2 * xy - 2 * wz;   L[2] = 2 * wy + 2 * xz;
    L[4] = 2 * wz + 2 * xy;   L[5] = w2 - x2 + y2 - z2; L[6] = 2 * yz - 2 * wx;
    L[8] = 2 * xz - 2 * wy;   L[9] = 2 * wx + 2 * yz;   L[10] = w2 - x2 - y2 + z2;
    const int par = i == 0 ? -1 : parents[i];
    L[3] = Js[3 * i] - (par >= 0 ? Js[3 * par] : 0.f);
    L[7] = Js[3 * i + 1] - (par >= 0 ? Js[3 * par + 1] : 0.f);
    L[11] = Js[3 * i + 2] - (par >= 0 ? Js[3 * par + 2] : 0.f);
    L[12] = 0.f; L[13] = 0.f; L[14] = 0.f; L[15] = 1.f;
    float* gi = g + 16 * i;
    if (par < 0) {
#pragma unroll
      for (int e = 0; e < 16; ++e) gi[e] = L[e];
    } else {
      const float* gp = g + 16 * par;
      float Pm[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) Pm[e] = gp[e];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          gi[4 * r + c] = Pm[4 * r] * L[c] + Pm[4 * r + 1] * L[4 + c] + Pm[4 * r + 2] * L[8 + c] + Pm[4 * r + 3] * L[12 + c];
    }
  }
  if (!A) return;
  for (int i = 0; i < 24; ++i) {
    const float* gi = g + 16 * i;
    float* a = A + ((size_t)f * 24 + i) * 16;
    if (init_pose) {
      const float* ip = init_pose + 16 * i;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          a[4 * r + c] = gi[4 * r] * ip[c] + gi[4 * r + 1] * ip[4 + c] + gi[4 * r + 2] * ip[8 + c] + gi[4 * r + 3] * ip[12 + c];
    } else {
      const float jx = Js[3 * i], jy = Js[3 * i + 1], jz = Js[3 * i + 2];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        a[4 * r] = gi[4 * r]; a[4 * r + 1] = gi[4 * r + 1]; a[4 * r + 2] = gi[4 * r + 2];
        a[4 * r + 3] = gi[4 * r + 3] - (gi[4 * r] * jx + gi[4 * r + 1] * jy + gi[4 * r + 2] * jz);
      }
    }
  }
}

static int lbs_check(const recmv_voxel_t* vox, int nf, int64_t P) {
  if (!vox || !vox->ws_cl) return RECMV_E_NULL;
  if (vox->D <= 0 || vox->H <= 0 || vox->W <= 0 || nf <= 0 || P < 0) return RECMV_E_SHAPE;
  if (((uintptr_t)vox->ws_cl & 15) != 0) return RECMV_E_SHAPE;  // float4 gathers
  return RECMV_OK;
}

}  // namespace recmv

using namespace recmv;

extern "C" int recmv_lbs_fwd(const float* ps, const float* tps, const float* A, const float* trans,
                             const int64_t* batch_inds, int64_t points_per_frame, int num_frames,
                             const recmv_voxel_t* vox, float* out, float* weights_out, int64_t P,
                             recmv_stream_t stream) {
  int s = lbs_check(vox, num_frames, P);
  if (s) return s;
  if (P == 0) return RECMV_OK;
  if (!ps || !A || !trans || !out) return RECMV_E_NULL;
  int g = stride_grid(P, 128, 8);
  lbs_fwd_kernel<<<g, 128, 0, (cudaStream_t)stream>>>(ps, tps, A, trans, batch_inds, points_per_frame,
                                                      num_frames, to_voxel(vox), out, weights_out, P);
  return launch_status();
}

extern "C" int recmv_lbs_inverse(const float* x_obs, const float* A, const float* trans,
                                 const int64_t* batch_inds, int64_t points_per_frame, int num_frames,
                                 const recmv_voxel_t* vox, float* x_can, uint8_t* valid, int64_t P,
                                 recmv_stream_t stream) {
  int s = lbs_check(vox, num_frames, P);
  if (s) return s;
  if (P == 0) return RECMV_OK;
  if (!x_obs || !A || !trans || !x_can) return RECMV_E_NULL;
  int g = stride_grid(P, 128, 8);
  lbs_inverse_kernel<<<g, 128, 0, (cudaStream_t)stream>>>(x_obs, A, trans, batch_inds, points_per_frame,
                                                          num_frames, to_voxel(vox), x_can, valid, P);
  return launch_status();
}

extern "C" int recmv_bone_matrices(const float* poses, const float* Js, const int* parents, const float* init_pose,
                                   float* G, float* A, int num_frames, recmv_stream_t stream) {
  if (num_frames < 0) return RECMV_E_SHAPE;
  if (num_frames == 0) return RECMV_OK;
  if (!poses || !Js || !parents || !G) return RECMV_E_NULL;
  bone_matrices_kernel<<<(num_frames + 63) / 64, 64, 0, (cudaStream_t)stream>>>(poses, Js, parents, init_pose, G, A,
                                                                                num_frames);
  return launch_status();
}

// Per-ray algebra of the implicit-surface gradient (engineer/networks/OptimNetwork.py:726-879,
// propagateTmpPsGrad; garment variant OptimGarmentNetwork.py:2159-2313): the surface point p* of a ray is defined
// implicitly by  f(p) = 0  and  (D(p) - c) x v = 0;  its sensitivity to the network parameters is the
// least-squares solution of  b dp = rhs  with the 4 x 3 matrix  b = [grad f ; [v]x J]  (J = dD/dp).  Given the
// loss gradient g = dL/dp*, the reference forms  r = g (b^T b)^-1 b^T  (1 x 4) with FastMinv's singularity rule and
// then back-propagates  -r[0]  through the SDF network and  r[1:4] (-[v]x)  through the deformer.  This kernel does
// everything between the two network evaluations in registers: 21 floats in, 8 floats + 1 flag out per ray
// (the reference: ~40 small kernels and a FastMinv call on the legacy stream).
#include "../../include/recmv_b200.h"
#include "common.cuh"

namespace recmv {
namespace {

__global__ void __launch_bounds__(256) surface_grad_kernel(const float* __restrict__ gl, const float* __restrict__ gf,
                                                           const float* __restrict__ J, const float* __restrict__ v,
                                                           const float* __restrict__ dc, float* __restrict__ sdf_coef,
                                                           float* __restrict__ def_vec, float* __restrict__ ray_grad,
                                                           unsigned char* __restrict__ ok, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float vx = __ldg(v + 3 * i), vy = __ldg(v + 3 * i + 1), vz = __ldg(v + 3 * i + 2);
    float Jm[9], a[9], b0[3], g[3];
#pragma unroll
    for (int e = 0; e < 9; ++e) Jm[e] = __ldg(J + 9 * i + e);
#pragma unroll
    for (int e = 0; e < 3; ++e) { b0[e] = __ldg(gf + 3 * i + e); g[e] = __ldg(gl + 3 * i + e); }
    // a = [v]x J : row 0 = -vz J1 + vy J2, row 1 = vz J0 - vx J2, row 2 = -vy J0 + vx J1
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      a[c] = -vz * Jm[3 + c] + vy * Jm[6 + c];
      a[3 + c] = vz * Jm[c] - vx * Jm[6 + c];
      a[6 + c] = -vy * Jm[c] + vx * Jm[3 + c];
    }
    // btb = b^T b = gf gf^T + a^T a
    float M[9], Mi[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        M[3 * r + c] = b0[r] * b0[c] + a[r] * a[c] + a[3 + r] * a[3 + c] + a[6 + r] * a[6 + c];
    const bool good = inv3x3<float>(M, Mi);   // |det| < 1e-4 -> zeros, flag false (Matrix3x3InvKernels.cu:29-60)
    // w = (btb_inv b^T) contracted with g from the left: r = g btb_inv b^T  ->  w = g btb_inv (1 x 3)
    float w[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) w[c] = g[0] * Mi[c] + g[1] * Mi[3 + c] + g[2] * Mi[6 + c];
    const float r0 = w[0] * b0[0] + w[1] * b0[1] + w[2] * b0[2];
    float r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = w[0] * a[3 * k] + w[1] * a[3 * k + 1] + w[2] * a[3 * k + 2];
    sdf_coef[i] = -r0;
    // temp = r (-[v]x) = -(r x ... ) : (r [v]x)_j = sum_i r_i [v]x_ij  = (r_1 vz - r_2 vy, -r_0 vz + r_2 vx, r_0 vy - r_1 vx)
    def_vec[3 * i + 0] = -(r[1] * vz - r[2] * vy);
    def_vec[3 * i + 1] = -(-r[0] * vz + r[2] * vx);
    def_vec[3 * i + 2] = -(r[0] * vy - r[1] * vx);
    if (ray_grad) {   // r [dc]x : the gradient w.r.t. the ray direction (OptimNetwork.py:862-873)
      const float dx = __ldg(dc + 3 * i), dy = __ldg(dc + 3 * i + 1), dz = __ldg(dc + 3 * i + 2);
      ray_grad[3 * i + 0] = r[1] * dz - r[2] * dy;
      ray_grad[3 * i + 1] = -r[0] * dz + r[2] * dx;
      ray_grad[3 * i + 2] = r[0] * dy - r[1] * dx;
    }
    ok[i] = good ? 1 : 0;
  }
}

}  // namespace
}  // namespace recmv

using namespace recmv;

extern "C" int recmv_surface_grad_coeffs(const float* grad_l_p, const float* grad_f_p, const float* jac,
                                         const float* rays, const float* d_minus_c, float* sdf_coef, float* def_vec,
                                         float* ray_grad, uint8_t* ok, int64_t n, recmv_stream_t stream) {
  if (n < 0) return RECMV_E_SHAPE;
  if (n == 0) return RECMV_OK;
  if (!grad_l_p || !grad_f_p || !jac || !rays || !sdf_coef || !def_vec || !ok) return RECMV_E_NULL;
  if (ray_grad && !d_minus_c) return RECMV_E_NULL;
  surface_grad_kernel<<<stride_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(grad_l_p, grad_f_p, jac, rays, d_minus_c,
                                                                               sdf_coef, def_vec, ray_grad, ok, n);
  return launch_status();
}

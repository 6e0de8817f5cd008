// Coarse-to-fine sweep helpers (Seg3dLossless, MCAcc/seg3d_lossless.py:233-428).
//  * recmv_interp2x_boundary3d_{fwd,bwd}: the 2x-1 upsampling + "mixed occupancy" flag of the reference's optional
//    extension (MCAcc/cuda/interp2x_boundary3d_kernel.cu:9-242; off by default there, use_cuda_impl=False) -- and,
//    selectable by `order`, the same operation with the rounding of the default path
//    (F.interpolate(trilinear, align_corners=True) on the values and on the 0/1 flags, seg3d_lossless.py:270-281),
//    so the default sweep can use one fused pass and stay bit-identical.
//  * recmv_c2f_todo_mask: 3x3x3 dilation of the flags (smooth_conv3x3 > 0) AND NOT already-evaluated.
// HBM-bound: 4 B read per coarse voxel (L2-resident re-reads), 5 B written per fine voxel.
#include "../../include/recmv_b200.h"
#include "common.cuh"

namespace recmv {
namespace {

// order 0: sequential sum / count as the reference kernel; order 1: ATen's nested x, y, z interpolation with
// lambda in {0, 1/2} (every product by 1/2 is exact, one rounding per addition)
template <int kOrder>
__global__ void __launch_bounds__(256) interp2x_fwd_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           unsigned char* __restrict__ flag, int NC, int D, int H,
                                                           int W, float balance) {
  const int d = 2 * D - 1, h = 2 * H - 1, w = 2 * W - 1;
  const long long total = (long long)NC * d * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h), z = (int)((i / ((long long)w * h)) % d);
    const long long nc = i / ((long long)w * h * d);
    const float* p = in + nc * ((long long)D * H * W);
    const int x0 = x >> 1, y0 = y >> 1, z0 = z >> 1;             // (x-1)/2 for odd x, x/2 for even x
    const int ox = x & 1, oy = y & 1, oz = z & 1;                 // odd: second neighbour at +1
    float v[8];
    bool any_in = false, any_out = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int dz = (k >> 2) & oz, dy = ((k >> 1) & 1) & oy, dx = (k & 1) & ox;
      v[k] = __ldg(p + ((long long)(z0 + dz) * H + (y0 + dy)) * W + (x0 + dx));
      const bool f = v[k] > balance;
      any_in |= f; any_out |= !f;
    }
    float r;
    if (kOrder == 0) {
      // the reference's summation orders (interp2x_boundary3d_kernel.cu:38-129): left to right over
      //   1 axis odd : the two neighbours;   x,y odd : (y-,x-)(y-,x+)(y+,x-)(y+,x+);
      //   y,z odd    : (z-,y-)(z+,y-)(z-,y+)(z+,y+);   x,z odd : (z-,x-)(z+,x-)(z-,x+)(z+,x+);   all odd : x, y, z nested
      const int n = (1 << ox) << (oy + oz);
      float s;
      if (n == 1) s = v[0];
      else if (n == 2) s = v[0] + (ox ? v[1] : (oy ? v[2] : v[4]));
      else if (n == 8) s = ((((((v[0] + v[1]) + v[2]) + v[3]) + v[4]) + v[5]) + v[6]) + v[7];
      else if (!oz) s = ((v[0] + v[1]) + v[2]) + v[3];
      else if (!ox) s = ((v[0] + v[4]) + v[2]) + v[6];
      else s = ((v[0] + v[4]) + v[1]) + v[5];
      r = n == 1 ? s : (float)((double)s / (double)n);
    } else {
      const float a00 = ox ? 0.5f * v[0] + 0.5f * v[1] : v[0], a01 = ox ? 0.5f * v[2] + 0.5f * v[3] : v[2];
      const float a10 = ox ? 0.5f * v[4] + 0.5f * v[5] : v[4], a11 = ox ? 0.5f * v[6] + 0.5f * v[7] : v[6];
      const float b0 = oy ? 0.5f * a00 + 0.5f * a01 : a00, b1 = oy ? 0.5f * a10 + 0.5f * a11 : a10;
      r = oz ? 0.5f * b0 + 0.5f * b1 : b0;
    }
    out[i] = r;
    flag[i] = (any_in && any_out) ? 1 : 0;
  }
}

// gradient of the (order 0 == order 1 up to rounding) linear map: gather of the 27 fine neighbours of 2*(z,y,x)
__global__ void __launch_bounds__(256) interp2x_bwd_kernel(const float* __restrict__ go, float* __restrict__ gi,
                                                           int NC, int D, int H, int W) {
  const int d = 2 * D - 1, h = 2 * H - 1, w = 2 * W - 1;
  const long long total = (long long)NC * D * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)((i / ((long long)W * H)) % D);
    const long long nc = i / ((long long)W * H * D);
    const float* g = go + nc * ((long long)d * h * w);
    // same accumulation order as the reference: centre, 6 axis neighbours (x-, x+, y-, y+, z-, z+) / 2, 12 face
    // diagonals (xy, xz, yz) / 4, 8 corners / 8
    auto at = [&](int dz, int dy, int dx) { return __ldg(g + ((long long)(2 * z + dz) * h + (2 * y + dy)) * w + (2 * x + dx)); };
    const bool xm = x > 0, xp = x < W - 1, ym = y > 0, yp = y < H - 1, zm = z > 0, zp = z < D - 1;
    float a = at(0, 0, 0);
    if (xm) a += at(0, 0, -1) / 2.f;
    if (xp) a += at(0, 0, 1) / 2.f;
    if (ym) a += at(0, -1, 0) / 2.f;
    if (yp) a += at(0, 1, 0) / 2.f;
    if (zm) a += at(-1, 0, 0) / 2.f;
    if (zp) a += at(1, 0, 0) / 2.f;
    if (xm && ym) a += at(0, -1, -1) / 4.f;
    if (xp && ym) a += at(0, -1, 1) / 4.f;
    if (xm && yp) a += at(0, 1, -1) / 4.f;
    if (xp && yp) a += at(0, 1, 1) / 4.f;
    if (xm && zm) a += at(-1, 0, -1) / 4.f;
    if (xp && zm) a += at(-1, 0, 1) / 4.f;
    if (xm && zp) a += at(1, 0, -1) / 4.f;
    if (xp && zp) a += at(1, 0, 1) / 4.f;
    if (ym && zm) a += at(-1, -1, 0) / 4.f;
    if (yp && zm) a += at(-1, 1, 0) / 4.f;
    if (ym && zp) a += at(1, -1, 0) / 4.f;
    if (yp && zp) a += at(1, 1, 0) / 4.f;
    if (xm && ym && zm) a += at(-1, -1, -1) / 8.f;
    if (xp && ym && zm) a += at(-1, -1, 1) / 8.f;
    if (xm && yp && zm) a += at(-1, 1, -1) / 8.f;
    if (xp && yp && zm) a += at(-1, 1, 1) / 8.f;
    if (xm && ym && zp) a += at(1, -1, -1) / 8.f;
    if (xp && ym && zp) a += at(1, -1, 1) / 8.f;
    if (xm && yp && zp) a += at(1, 1, -1) / 8.f;
    if (xp && yp && zp) a += at(1, 1, 1) / 8.f;
    gi[i] = a;
  }
}

// todo = dilate3x3x3(flag) & ~done   (zero padding at the volume border)
__global__ void __launch_bounds__(256) c2f_todo_kernel(const unsigned char* __restrict__ flag,
                                                       const unsigned char* __restrict__ done,
                                                       unsigned char* __restrict__ todo, int D, int H, int W) {
  const long long total = (long long)D * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    unsigned char r = 0;
    if (!done[i]) {
      const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)(i / ((long long)W * H));
      for (int dz = -1; dz <= 1 && !r; ++dz) {
        const int zz = z + dz;
        if (zz < 0 || zz >= D) continue;
        for (int dy = -1; dy <= 1 && !r; ++dy) {
          const int yy = y + dy;
          if (yy < 0 || yy >= H) continue;
          const unsigned char* row = flag + ((long long)zz * H + yy) * W;
          r = (x > 0 && row[x - 1]) || row[x] || (x < W - 1 && row[x + 1]);
        }
      }
    }
    todo[i] = r;
  }
}

}  // namespace
}  // namespace recmv

using namespace recmv;

extern "C" int recmv_interp2x_boundary3d_fwd(const float* input, float* output, uint8_t* is_boundary, int NC, int D,
                                             int H, int W, float balance_value, int order, recmv_stream_t stream) {
  if (NC < 0 || D <= 0 || H <= 0 || W <= 0) return RECMV_E_SHAPE;
  if (order != 0 && order != 1) return RECMV_E_RANGE;
  if (NC == 0) return RECMV_OK;
  if (!input || !output || !is_boundary) return RECMV_E_NULL;
  const int64_t total = (int64_t)NC * (2 * D - 1) * (2 * H - 1) * (2 * W - 1);
  const int g = stride_grid(total, 256, 16);
  if (order == 0) interp2x_fwd_kernel<0><<<g, 256, 0, (cudaStream_t)stream>>>(input, output, is_boundary, NC, D, H, W, balance_value);
  else interp2x_fwd_kernel<1><<<g, 256, 0, (cudaStream_t)stream>>>(input, output, is_boundary, NC, D, H, W, balance_value);
  return launch_status();
}

extern "C" int recmv_interp2x_boundary3d_bwd(const float* grad_output, float* grad_input, int NC, int D, int H, int W,
                                             recmv_stream_t stream) {
  if (NC < 0 || D <= 0 || H <= 0 || W <= 0) return RECMV_E_SHAPE;
  if (NC == 0) return RECMV_OK;
  if (!grad_output || !grad_input) return RECMV_E_NULL;
  interp2x_bwd_kernel<<<stride_grid((int64_t)NC * D * H * W, 256, 16), 256, 0, (cudaStream_t)stream>>>(grad_output, grad_input, NC, D, H, W);
  return launch_status();
}

extern "C" int recmv_c2f_todo_mask(const uint8_t* is_boundary, const uint8_t* done, uint8_t* todo, int D, int H, int W,
                                   recmv_stream_t stream) {
  if (D <= 0 || H <= 0 || W <= 0) return RECMV_E_SHAPE;
  if (!is_boundary || !done || !todo) return RECMV_E_NULL;
  c2f_todo_kernel<<<stride_grid((int64_t)D * H * W, 256, 16), 256, 0, (cudaStream_t)stream>>>(is_boundary, done, todo, D, H, W);
  return launch_status();
}

// Camera rays and sphere intersections.
//   rend_util.get_camera_params + lift   (/root/reference/code/lib/utils/rend_util.py:45-87)
//   rend_util.get_sphere_intersections   (rend_util.py:131-147)
//   LaplaceDensity elementwise           (lib/model/density.py:20-25)
#include "common.cuh"

namespace mp {

__global__ void camera_rays_kernel(const float* __restrict__ uv, const float* __restrict__ pose,
                                   const float* __restrict__ K, int R, float* __restrict__ dirs,
                                   float* __restrict__ cam) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  float fx = K[0], sk = K[1], cx = K[2], fy = K[5], cy = K[6];
  float x = uv[2 * i], y = uv[2 * i + 1];
  // x_lift = (x - cx + cy*sk/fy - sk*y/fy) / fx * z ;  z = 1      (rend_util.py:84)
  float t1 = __fsub_rn(x, cx);
  float t2 = __fdiv_rn(__fmul_rn(cy, sk), fy);
  float t3 = __fdiv_rn(__fmul_rn(sk, y), fy);
  float xl = __fmul_rn(__fdiv_rn(__fsub_rn(__fadd_rn(t1, t2), t3), fx), 1.0f);
  float yl = __fmul_rn(__fdiv_rn(__fsub_rn(y, cy), fy), 1.0f);
  float o[3], d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float* p = pose + 4 * r;
    // world = p @ [xl, yl, 1, 1]
    float w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p[0], xl), __fmul_rn(p[1], yl)), p[2]), p[3]);
    o[r] = p[3];
    d[r] = __fsub_rn(w, p[3]);
  }
  // F.normalize(dim=2): v / max(||v||, 1e-12)
  float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
  n = fmaxf(n, 1e-12f);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    dirs[3 * i + r] = __fdiv_rn(d[r], n);
    cam[3 * i + r] = o[r];
  }
}

__device__ __forceinline__ void sphere_isect(const float* o, const float* d, float r, float& tn, float& tf, bool& bad) {
  float dot = __fadd_rn(__fadd_rn(__fmul_rn(d[0], o[0]), __fmul_rn(d[1], o[1])), __fmul_rn(d[2], o[2]));
  float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(o[0], o[0]), __fmul_rn(o[1], o[1])), __fmul_rn(o[2], o[2])));
  float under = __fsub_rn(__fmul_rn(dot, dot), __fsub_rn(__fmul_rn(nrm, nrm), __fmul_rn(r, r)));
  bad = !(under > 0.f);
  float s = sqrtf(under);
  tn = fmaxf(__fsub_rn(__fmul_rn(s, -1.f), dot), 0.f);
  tf = fmaxf(__fsub_rn(__fmul_rn(s, 1.f), dot), 0.f);
}

__global__ void sphere_kernel(const float* __restrict__ cam, const float* __restrict__ dirs, int R, float r,
                              float* __restrict__ nf, int* __restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  float tn, tf;
  bool bad;
  sphere_isect(cam + 3 * i, dirs + 3 * i, r, tn, tf, bad);
  if (bad && flag) atomicOr(flag, 1);
  nf[2 * i] = tn;
  nf[2 * i + 1] = tf;
}

__global__ void density_kernel(const float* __restrict__ sdf, int N, float beta, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = laplace_density(sdf[i], beta);
}

}  // namespace mp

extern "C" {

int mp_camera_rays(const float* uv, const float* pose, const float* intrinsics, int R, float* ray_dirs,
                   float* cam_loc, void* stream) {
  if (R <= 0) return 0;
  mp::camera_rays_kernel<<<mp::div_up(R, 256), 256, 0, (cudaStream_t)stream>>>(uv, pose, intrinsics, R, ray_dirs,
                                                                               cam_loc);
  MP_LAUNCH_CHECK();
  return 0;
}

int mp_sphere_intersections(const float* cam_loc, const float* ray_dirs, int R, float r, float* near_far,
                            int* status_flag, void* stream) {
  if (R <= 0) return 0;
  mp::sphere_kernel<<<mp::div_up(R, 256), 256, 0, (cudaStream_t)stream>>>(cam_loc, ray_dirs, R, r, near_far,
                                                                          status_flag);
  MP_LAUNCH_CHECK();
  return 0;
}

int mp_laplace_density(const float* sdf, int N, float beta, float* sigma, void* stream) {
  if (N <= 0) return 0;
  mp::density_kernel<<<mp::div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(sdf, N, beta, sigma);
  MP_LAUNCH_CHECK();
  return 0;
}
}

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

// Ray / oriented-box culling (replaces the host-side trimesh RayMeshIntersector of multiply.py:208-214, :256-263).
// Slab test in the box frame, fp64 like the host code it replaces; one CTA compacts the hit ray ids IN ORDER
// (the reference sorts them, :258-259) with ballot + block scan, so no separate sort is needed.
__global__ void __launch_bounds__(1024) ray_box_hits_kernel(const float* __restrict__ cam, const float* __restrict__ dirs,
                                                            int R, double cx, double cy, double cz, double hx, double hy,
                                                            double hz, const double* __restrict__ rot,
                                                            int64_t* __restrict__ idx_out, int* __restrict__ count,
                                                            const double* __restrict__ box_dev, int finalize) {
  __shared__ int s_warp[32];
  __shared__ int s_base, s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  if (box_dev) {     // centre / half extents computed on the device (aabb_kernel)
    cx = box_dev[0]; cy = box_dev[1]; cz = box_dev[2];
    hx = box_dev[3]; hy = box_dev[4]; hz = box_dev[5];
  }
  double Rm[9];
  for (int k = 0; k < 9; ++k) Rm[k] = rot ? rot[k] : ((k % 4 == 0) ? 1.0 : 0.0);
  for (int r0 = 0; r0 < R; r0 += blockDim.x) {
    int r = r0 + tid;
    bool hit = false;
    if (r < R) {
      double o[3] = {(double)cam[3 * r] - cx, (double)cam[3 * r + 1] - cy, (double)cam[3 * r + 2] - cz};
      double d[3] = {(double)dirs[3 * r], (double)dirs[3 * r + 1], (double)dirs[3 * r + 2]};
      double h[3] = {hx, hy, hz};
      double tmin = -1e300, tmax = 1e300;
      for (int a = 0; a < 3; ++a) {
        double oa = Rm[3 * a] * o[0] + Rm[3 * a + 1] * o[1] + Rm[3 * a + 2] * o[2];
        double da = Rm[3 * a] * d[0] + Rm[3 * a + 1] * d[1] + Rm[3 * a + 2] * d[2];
        if (fabs(da) < 1e-12) da = 1e-12;
        double inv = 1.0 / da;
        double t1 = (-h[a] - oa) * inv, t2 = (h[a] - oa) * inv;
        tmin = fmax(tmin, fmin(t1, t2));
        tmax = fmin(tmax, fmax(t1, t2));
      }
      hit = tmax >= fmax(tmin, 0.0);
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
      int v = s_warp[lane];
      int incl = v;
      for (int o2 = 1; o2 < 32; o2 <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o2);
        if (lane >= o2) incl += t;
      }
      s_warp[lane] = incl - v;          // exclusive prefix of the warp counts
      if (lane == 31) s_total = incl;
    }
    __syncthreads();
    if (hit) idx_out[s_base + s_warp[warp] + __popc(m & ((1u << lane) - 1u))] = r;
    __syncthreads();
    if (tid == 0) s_base += s_total;
    __syncthreads();
  }
  if (tid == 0 && count) {
    if (finalize && s_base == 0) {     // multiply.py:262-263: an empty hit list becomes the single ray 0
      idx_out[0] = 0;
      s_base = 1;
    }
    *count = s_base;
  }
}

// centre and inflated half extents of the axis-aligned bounds of verts [V,3] (one CTA; fp64 like the host code)
__global__ void __launch_bounds__(1024) aabb_kernel(const float* __restrict__ verts, int V, double inflate,
                                                    double* __restrict__ box) {
  __shared__ float s_lo[3][32], s_hi[3][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int v = tid; v < V; v += blockDim.x)
    for (int k = 0; k < 3; ++k) {
      float x = verts[3 * v + k];
      lo[k] = fminf(lo[k], x);
      hi[k] = fmaxf(hi[k], x);
    }
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    if (lane == 0) {
      s_lo[k][warp] = lo[k];
      s_hi[k][warp] = hi[k];
    }
  }
  __syncthreads();
  if (tid < 3) {
    float l = 3.4e38f, h = -3.4e38f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      l = fminf(l, s_lo[tid][w]);
      h = fmaxf(h, s_hi[tid][w]);
    }
    box[tid] = ((double)l + (double)h) / 2.0;
    box[3 + tid] = ((double)h - (double)l) / 2.0 * inflate;
  }
}

__global__ void hit_list_finalize_kernel(int64_t* __restrict__ idx, int* __restrict__ count) {
  if (*count == 0) {
    idx[0] = 0;
    *count = 1;
  }
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

int mp_ray_box_hits(const float* cam_loc, const float* ray_dirs, int R, const double* center_host,
                    const double* half_extent_host, const double* rot_dev, int64_t* idx_out, int* count_dev,
                    void* stream) {
  MP_REQUIRE(cam_loc && ray_dirs && center_host && half_extent_host && idx_out && count_dev,
             "mp_ray_box_hits: null argument");
  if (R <= 0) return 0;
  mp::ray_box_hits_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(cam_loc, ray_dirs, R, center_host[0], center_host[1],
                                                               center_host[2], half_extent_host[0],
                                                               half_extent_host[1], half_extent_host[2], rot_dev,
                                                               idx_out, count_dev, nullptr, 0);
  MP_LAUNCH_CHECK();
  return 0;
}

int mp_hit_list_finalize(int64_t* idx, int* count_dev, void* stream) {
  MP_REQUIRE(idx && count_dev, "mp_hit_list_finalize: null argument");
  mp::hit_list_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(idx, count_dev);
  MP_LAUNCH_CHECK();
  return 0;
}

int mp_ray_aabb_hits(const float* cam_loc, const float* ray_dirs, int R, const float* verts, int V, double inflate,
                     int64_t* idx_out, int* count_dev, void* box_ws, void* stream) {
  MP_REQUIRE(cam_loc && ray_dirs && verts && idx_out && count_dev && box_ws, "mp_ray_aabb_hits: null argument");
  MP_REQUIRE(V > 0, "mp_ray_aabb_hits: V must be positive");
  if (R <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  mp::aabb_kernel<<<1, 1024, 0, st>>>(verts, V, inflate, (double*)box_ws);
  MP_LAUNCH_CHECK();
  // the hit list comes out finalised (multiply.py:262-263): empty -> ray 0
  mp::ray_box_hits_kernel<<<1, 1024, 0, st>>>(cam_loc, ray_dirs, R, 0, 0, 0, 0, 0, 0, nullptr, idx_out, count_dev,
                                              (const double*)box_ws, 1);
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

// NeRF++ inverted-sphere background.
//   reference: /root/reference/code/lib/model/multiply.py
//     background rendering block   :514-539
//     bg_volume_rendering          :682-696  (AbsDensity, lib/model/density.py:32-34)
//     depth2pts_outside            :698-726
//   and the inverse-sphere sample depths of ray_sampler.py:215-218 / multiply.py:482-484.
#include "common.cuh"

namespace mp {

int field_bg(const Field& f, const float* pts, const float* dirs, int N, float* sdf, float* rgb, void* ws,
             size_t ws_bytes, cudaStream_t st);   // render.cu (engine dispatch)
size_t field_bg_ws_bytes(int N);

__device__ __forceinline__ float bg_linspace32(int i) {
  float step = 1.0f / 31.0f;
  return (i < 16) ? fmaf(step, (float)i, 0.f) : fmaf(-step, (float)(31 - i), 1.0f);
}

// one thread per (ray, sample): depth = flip(linspace(0,1,32) / bound)[j] ; pts = depth2pts_outside(o, d, depth)
// inverse-sphere depth k of ray r: linspace(0,1,32)[k] / bound (ray_sampler.py:215-218); with t_rand (training mode:
// the UniformSampler of the inverse sphere sees model.training, ray_sampler.py:32-40) stratified between the midpoints
__device__ __forceinline__ float bg_depth(int r, int k, float inv_bound, const float* __restrict__ t_rand) {
  float zk = bg_linspace32(k);
  if (t_rand) {
    float lower = (k == 0) ? zk : .5f * (zk + bg_linspace32(k - 1));
    float upper = (k == 31) ? zk : .5f * (bg_linspace32(k + 1) + zk);
    zk = lower + (upper - lower) * t_rand[(size_t)r * 32 + k];
  }
  return zk * inv_bound;
}

__global__ void bg_points_kernel(const float* __restrict__ dirs, const float* __restrict__ cam, int R, float bound,
                                 float inv_bound, float* __restrict__ pts, float* __restrict__ dirs_out,
                                 const float* __restrict__ t_rand) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * 32) return;
  int r = idx >> 5, j = idx & 31;
  float depth = bg_depth(r, 31 - j, inv_bound, t_rand);     // torch.flip(z_vals_bg)   multiply.py:516
  const float* o = cam + 3 * r;
  const float* d = dirs + 3 * r;
  float odd = d[0] * o[0] + d[1] * o[1] + d[2] * o[2];
  float oo = o[0] * o[0] + o[1] * o[1] + o[2] * o[2];
  float under = odd * odd - (oo - bound * bound);
  float dsph = sqrtf(under) - odd;
  float ps[3], pm[3];
  for (int k = 0; k < 3; ++k) {
    ps[k] = o[k] + dsph * d[k];
    pm[k] = o[k] - odd * d[k];
  }
  float pmn = sqrtf(pm[0] * pm[0] + pm[1] * pm[1] + pm[2] * pm[2]);
  float ax[3] = {o[1] * ps[2] - o[2] * ps[1], o[2] * ps[0] - o[0] * ps[2], o[0] * ps[1] - o[1] * ps[0]};
  float an = sqrtf(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  for (int k = 0; k < 3; ++k) ax[k] = ax[k] / an;
  float phi = asinf(pmn / bound);
  float theta = asinf(pmn * depth);
  float ang = phi - theta;
  float ca = cosf(ang), sa = sinf(ang);
  float cr[3] = {ax[1] * ps[2] - ax[2] * ps[1], ax[2] * ps[0] - ax[0] * ps[2], ax[0] * ps[1] - ax[1] * ps[0]};
  float dt = ax[0] * ps[0] + ax[1] * ps[1] + ax[2] * ps[2];
  float pn[3];
  for (int k = 0; k < 3; ++k) pn[k] = ps[k] * ca + cr[k] * sa + ax[k] * dt * (1.f - ca);
  float nn = sqrtf(pn[0] * pn[0] + pn[1] * pn[1] + pn[2] * pn[2]);
  float* q = pts + 4 * (size_t)idx;
  q[0] = pn[0] / nn;
  q[1] = pn[1] / nn;
  q[2] = pn[2] / nn;
  q[3] = depth;
  dirs_out[3 * (size_t)idx] = d[0];
  dirs_out[3 * (size_t)idx + 1] = d[1];
  dirs_out[3 * (size_t)idx + 2] = d[2];
}

// one warp per ray, lane = sample     multiply.py:682-696, :539
__global__ void bg_composite_kernel(const float* __restrict__ sdf, const float* __restrict__ rgb, int R,
                                    float inv_bound, float* __restrict__ out, const float* __restrict__ t_rand) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= R) return;
  size_t i = (size_t)w * 32 + lane;
  float dens = fabsf(sdf[i]);
  float zc = bg_depth(w, 31 - lane, inv_bound, t_rand);
  float zn = (lane < 31) ? bg_depth(w, 30 - lane, inv_bound, t_rand) : 0.f;
  float dist = (lane < 31) ? (zc - zn) : 1e10f;
  float fe = dist * dens;
  float T = expf(-warp_scan_excl(fe, lane));
  float wgt = (1.f - expf(-fe)) * T;
  for (int k = 0; k < 3; ++k) {
    float v = warp_sum(wgt * rgb[3 * i + k]);
    if (lane == 0) out[3 * (size_t)w + k] = v;
  }
}

size_t bg_ws_bytes(int R) {
  size_t N = (size_t)(R > 0 ? R : 1) * 32;
  return align_up(N * 4 * 4, 256) + align_up(N * 3 * 4, 256) * 2 + align_up(N * 4, 256) + field_bg_ws_bytes((int)N) +
         4096;
}

int render_background(const Field& f, const float* dirs, const float* cam, int R, float bound, float* bg_rgb,
                      void* ws, size_t ws_bytes, cudaStream_t st, const float* t_rand) {
  if (R <= 0) return 0;
  Arena a(ws, ws_bytes);
  int N = R * 32;
  float* pts = a.take<float>((size_t)N * 4);
  float* dexp = a.take<float>((size_t)N * 3);
  float* rgb = a.take<float>((size_t)N * 3);
  float* sdf = a.take<float>(N);
  size_t mb = field_bg_ws_bytes(N);
  void* mws = a.take<char>(mb);
  MP_REQUIRE(a.ok, "background: workspace too small (%zu needed, %zu given)", a.off, ws_bytes);
  float inv_bound = (float)(1.0 / bound);
  bg_points_kernel<<<div_up(N, 256), 256, 0, st>>>(dirs, cam, R, bound, inv_bound, pts, dexp, t_rand);
  MP_LAUNCH_CHECK();
  MP_TRY(field_bg(f, pts, dexp, N, sdf, rgb, mws, mb, st));
  bg_composite_kernel<<<div_up(N, 256), 256, 0, st>>>(sdf, rgb, R, inv_bound, bg_rgb, t_rand);
  MP_LAUNCH_CHECK();
  return 0;
}

}  // namespace mp

extern "C" {
size_t mp_background_workspace_bytes(int R) { return mp::bg_ws_bytes(R); }

int mp_background(mp_net_t* bg_field, const float* ray_dirs, const float* cam_loc, int R, float bound_r, float* bg_rgb,
                  void* workspace, size_t workspace_bytes, void* stream) {
  MP_REQUIRE(bg_field && ray_dirs && cam_loc && bg_rgb, "mp_background: null argument");
  return mp::render_background(bg_field->f, ray_dirs, cam_loc, R, bound_r, bg_rgb, workspace, workspace_bytes,
                               (cudaStream_t)stream, nullptr);
}
}

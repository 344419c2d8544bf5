// Engine dispatch and the fused eval-mode forward.
//   reference: /root/reference/code/lib/model/multiply.py:174-598 (Multiply.forward, eval branch,
//   using_nerfacc=True): per person  sample -> deform -> SDF -> normals/colour, then the
//   multi-person composite, the background, and the final blend.
#include "common.cuh"

namespace mp {

int g_engine = 1;

// mlp_tc.cu
int tc_sdf_list(const Field& f, const float* xc_list, const int* slot_list, const int* count_dev, int cap,
                float* sdf_out, void* ws, size_t ws_bytes, cudaStream_t st);
int tc_shade_list(const Field& f, const float* xc_list, const int* slot_list, const int* count_dev, int cap,
                  const float* Jinv_list, float* sdf_out, float* rgb_out, float* normal_out, float* grad_out,
                  float* feat_out, void* ws, size_t ws_bytes, cudaStream_t st);
int tc_bg(const Field& f, const float* pts, const float* dirs, int N, float* sdf, float* rgb, void* ws,
          size_t ws_bytes, cudaStream_t st);
size_t tc_workspace_bytes(int N);
int tc_trace_read(unsigned long long* out, int n);
int prof_enable(int on);
int prof_read(double* ms, long long* launches, double* points, int reset);

// sampler.cu / composite.cu / background.cu
int sample_rays(const mp_sampler_cfg_t& c, const Body& body, const Field& field, const float* dirs,
                const float* cam, int R, float* z_final, float* z_bg, int* trips_out, void* ws, size_t ws_bytes,
                cudaStream_t st);
size_t sampler_ws_bytes(const mp_sampler_cfg_t& c, int R);
struct CompositePersons {
  int P;
  int n_rows[MP_MAX_PERSONS];
  const int* row_of_ray[MP_MAX_PERSONS];
  const float* z[MP_MAX_PERSONS];
  const float* sdf[MP_MAX_PERSONS];
  const float* rgb[MP_MAX_PERSONS];
  const float* nrm[MP_MAX_PERSONS];
};
int launch_composite(const CompositePersons& cp, int R, int n, float beta, float* fg_rgb, float* normal, float* acc,
                     float* acc_person, float* bg_T, cudaStream_t st);
int launch_row_of_ray(const int64_t* idx, int n_rows, int R, int* row_of_ray, cudaStream_t st);
int launch_final_compose(const float* fg, const float* bgT, const float* bg, int R, float* rgb, float* fg_out,
                         cudaStream_t st);
int render_background(const Field& f, const float* dirs, const float* cam, int R, float bound, float* bg_rgb,
                      void* ws, size_t ws_bytes, cudaStream_t st);
size_t bg_ws_bytes(int R);

static size_t engine_ws_bytes(int N) {
  size_t a = simt_workspace_bytes(N), b = tc_workspace_bytes(N);
  return a > b ? a : b;
}
size_t field_sdf_ws_bytes(int cap) { return engine_ws_bytes(cap); }
size_t field_bg_ws_bytes(int N) { return engine_ws_bytes(N); }

int field_sdf_list(const Field& f, const float* xc_list, const int* slot_list, const int* count_dev, int cap,
                   float* sdf_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (cap <= 0) return 0;
  if (g_engine == 1) return tc_sdf_list(f, xc_list, slot_list, count_dev, cap, sdf_out, ws, ws_bytes, st);
  return simt_sdf_list(f, xc_list, slot_list, count_dev, cap, sdf_out, ws, ws_bytes, st);
}
int field_shade_list(const Field& f, const float* xc_list, const int* slot_list, const int* count_dev, int cap,
                     const float* Jinv_list, float* sdf_out, float* rgb_out, float* normal_out, float* grad_out,
                     float* feat_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (cap <= 0) return 0;
  if (g_engine == 1)
    return tc_shade_list(f, xc_list, slot_list, count_dev, cap, Jinv_list, sdf_out, rgb_out, normal_out, grad_out,
                         feat_out, ws, ws_bytes, st);
  return simt_shade_list(f, xc_list, slot_list, count_dev, cap, Jinv_list, sdf_out, rgb_out, normal_out, grad_out,
                         feat_out, ws, ws_bytes, st);
}
int field_bg(const Field& f, const float* pts, const float* dirs, int N, float* sdf, float* rgb, void* ws,
             size_t ws_bytes, cudaStream_t st) {
  if (g_engine == 1) return tc_bg(f, pts, dirs, N, sdf, rgb, ws, ws_bytes, st);
  return simt_bg(f, pts, dirs, N, sdf, rgb, ws, ws_bytes, st);
}

__global__ void gather_rays_kernel(const float* __restrict__ dirs, const float* __restrict__ cam,
                                   const int64_t* __restrict__ idx, int n, float* __restrict__ d_out,
                                   float* __restrict__ c_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t r = idx[i];
  for (int k = 0; k < 3; ++k) {
    d_out[3 * i + k] = dirs[3 * r + k];
    c_out[3 * i + k] = cam[3 * r + k];
  }
}
__global__ void force_outlier_sdf_kernel(const uint8_t* __restrict__ outl, int n, float* __restrict__ sdf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && outl[i]) sdf[i] = 4.0f;     // multiply.py:142-143
}
__global__ void iota_kernel(int* p, int n, int* count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
  if (i == 0 && count) *count = n;
}

// Outlier samples may be skipped in the colour pass only if their compositing weight is exactly
// zero: sigma(4; beta) == 0 in fp32, i.e. expm1(-4/beta) == -1 (density.py:24) — true for
// beta < 0.23.  (The SDF clamp itself, multiply.py:142-143, is always exact.)
static bool prune_is_exact(float beta) { return (4.0f / beta) > 18.0f; }

// The persons of a scene and the background are independent until the compositor (multiply.py:266-410 is a Python
// loop over persons; :514-539 only needs the rays).  Each branch runs on its own stream so that the small
// latency-bound kernels of one (deformer, sampler trips) fill the SMs under the persistent MLP kernel of another,
// which issues on a third of the cycles.  mp_set_streams(0) restores the single-stream schedule.
struct BranchStreams {
  bool ready = false;
  cudaStream_t s[MP_MAX_PERSONS + 1];
  cudaEvent_t fork, join[MP_MAX_PERSONS + 1];
};
static BranchStreams g_bs;
static int g_streams_on = -1;

static int branch_streams_init() {
  if (g_bs.ready) return 0;
  for (int i = 0; i <= MP_MAX_PERSONS; ++i) {
    MP_CHECK_CUDA(cudaStreamCreateWithFlags(&g_bs.s[i], cudaStreamNonBlocking));
    MP_CHECK_CUDA(cudaEventCreateWithFlags(&g_bs.join[i], cudaEventDisableTiming));
  }
  MP_CHECK_CUDA(cudaEventCreateWithFlags(&g_bs.fork, cudaEventDisableTiming));
  g_bs.ready = true;
  return 0;
}

struct PersonBufs {
  float *dirs, *cam, *z, *sdf, *rgb, *nrm, *xc_list, *jinv;
  int *slot_list, *count, *row_of_ray;
  uint8_t* outl;
};

struct RenderWs {
  float *dirs, *cam, *fg, *nrm, *acc, *accp, *bgT, *bg;
  PersonBufs pb[MP_MAX_PERSONS];
  // scratch of the sampler / MLP engine: one per concurrently running branch (persons, background)
  void* sub[MP_MAX_PERSONS + 1];
  size_t sub_bytes;
};

static bool render_carve(Arena& a, const mp_scene_t& sc, int R, RenderWs& w) {
  const mp_sampler_cfg_t& c = sc.sampler;
  int n = c.N_samples + c.N_samples_extra + 1;
  w.dirs = a.take<float>((size_t)R * 3);
  w.cam = a.take<float>((size_t)R * 3);
  w.fg = a.take<float>((size_t)R * 3);
  w.nrm = a.take<float>((size_t)R * 3);
  w.acc = a.take<float>(R);
  w.accp = a.take<float>((size_t)R * sc.P);
  w.bgT = a.take<float>(R);
  w.bg = a.take<float>((size_t)R * 3);
  size_t sub = bg_ws_bytes(R);
  for (int p = 0; p < sc.P; ++p) {
    int Rp = sc.hit_count[p];
    PersonBufs& b = w.pb[p];
    b.dirs = a.take<float>((size_t)Rp * 3);
    b.cam = a.take<float>((size_t)Rp * 3);
    b.z = a.take<float>((size_t)Rp * (n + 1));
    b.sdf = a.take<float>((size_t)Rp * n);
    b.rgb = a.take<float>((size_t)Rp * n * 3);
    b.nrm = a.take<float>((size_t)Rp * n * 3);
    b.xc_list = a.take<float>((size_t)Rp * n * 3);
    b.jinv = a.take<float>((size_t)Rp * n * 12);
    b.slot_list = a.take<int>((size_t)Rp * n);
    b.count = a.take<int>(1);
    b.row_of_ray = a.take<int>(R);
    b.outl = a.take<uint8_t>((size_t)Rp * n);
    sub = max(sub, sampler_ws_bytes(c, Rp));
    sub = max(sub, engine_ws_bytes(Rp * n));
  }
  w.sub_bytes = sub;
  for (int i = 0; i <= sc.P; ++i) w.sub[i] = a.take<char>(sub);     // [P] = background branch
  return a.ok;
}

}  // namespace mp

extern "C" {

int mp_set_engine(int engine) {
  MP_REQUIRE(engine == 0 || engine == 1, "mp_set_engine: engine must be 0 (simt fp32) or 1 (tcgen05)");
  mp::g_engine = engine;
  return 0;
}
int mp_get_engine(void) { return mp::g_engine; }

int mp_profile_enable(int on) { return mp::prof_enable(on); }
int mp_set_streams(int on) {
  mp::g_streams_on = on ? 1 : 0;
  return 0;
}

int mp_tc_trace_read(unsigned long long* out, int n) {
  MP_REQUIRE(out && n > 0, "mp_tc_trace_read: null argument");
  return mp::tc_trace_read(out, n);
}

int mp_profile_read(double* ms_host, long long* launches_host, double* points_host, int reset) {
  MP_REQUIRE(ms_host && launches_host && points_host, "mp_profile_read: null argument");
  return mp::prof_read(ms_host, launches_host, points_host, reset);
}

size_t mp_mlp_workspace_bytes(int N) { return mp::engine_ws_bytes(N) + (size_t)N * (9 + 1) * sizeof(float) + 4096; }

int mp_implicit_forward(mp_net_t* f, const float* x, int N, float* sdf, float* feat, void* workspace,
                        size_t workspace_bytes, void* stream) {
  MP_REQUIRE(f && x, "mp_implicit_forward: null argument");
  if (N <= 0) return 0;      // networks.py:131
  cudaStream_t st = (cudaStream_t)stream;
  if (feat == nullptr) return mp::field_sdf_list(f->f, x, nullptr, nullptr, N, sdf, workspace, workspace_bytes, st);
  return mp::field_shade_list(f->f, x, nullptr, nullptr, N, nullptr, sdf, nullptr, nullptr, nullptr, feat, workspace,
                              workspace_bytes, st);
}

int mp_implicit_forward_grad(mp_net_t* f, const float* x, int N, float* sdf, float* feat, float* grad,
                             void* workspace, size_t workspace_bytes, void* stream) {
  MP_REQUIRE(f && x && grad, "mp_implicit_forward_grad: null argument");
  if (N <= 0) return 0;
  return mp::field_shade_list(f->f, x, nullptr, nullptr, N, nullptr, sdf, nullptr, nullptr, grad, feat, workspace,
                              workspace_bytes, (cudaStream_t)stream);
}

int mp_render_forward(mp_net_t* f, const float* points, const float* normals, const float* feat, int N, float* rgb,
                      void* workspace, size_t workspace_bytes, void* stream) {
  MP_REQUIRE(f && points && normals && feat && rgb, "mp_render_forward: null argument");
  if (N <= 0) return 0;
  // the standalone colour operator always runs on the fp32 SIMT kernels (the fused tcgen05 chain
  // consumes features straight from shared memory and has no (points, normals, feat) entry)
  return mp::simt_render(f->f, points, normals, feat, N, rgb, workspace, workspace_bytes, (cudaStream_t)stream);
}

int mp_sdf_with_deformer(mp_body_t* body, mp_net_t* field, const float* x, int N, float* sdf, float* x_c,
                         float* feat, void* workspace, size_t workspace_bytes, void* stream) {
  MP_REQUIRE(body && field && x && sdf && x_c, "mp_sdf_with_deformer: null argument");
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  mp::Arena a(workspace, workspace_bytes);
  uint8_t* outl = a.take<uint8_t>(N);
  size_t mb = mp::engine_ws_bytes(N);
  void* mws = a.take<char>(mb);
  MP_REQUIRE(a.ok, "mp_sdf_with_deformer: workspace too small (%zu needed)", a.off);
  MP_TRY(mp_deform_inverse(body, x, N, x_c, outl, 1, stream));
  if (feat)
    MP_TRY(mp::field_shade_list(field->f, x_c, nullptr, nullptr, N, nullptr, sdf, nullptr, nullptr, nullptr, feat, mws,
                                mb, st));
  else
    MP_TRY(mp::field_sdf_list(field->f, x_c, nullptr, nullptr, N, sdf, mws, mb, st));
  mp::force_outlier_sdf_kernel<<<mp::div_up(N, 256), 256, 0, st>>>(outl, N, sdf);
  MP_LAUNCH_CHECK();
  return 0;
}

size_t mp_render_workspace_bytes(const mp_scene_t* scene, int R) {
  if (!scene) return 0;
  mp::Arena a(nullptr, 0);
  mp::RenderWs w;
  mp::render_carve(a, *scene, R, w);
  return a.off + 8192;
}

int mp_render_rays(const mp_scene_t* scene, const float* uv, const float* pose, const float* intrinsics, int R,
                   const mp_render_out_t* out, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace mp;
  MP_REQUIRE(scene && uv && pose && intrinsics && out, "mp_render_rays: null argument");
  MP_REQUIRE(scene->P >= 1 && scene->P <= MP_MAX_PERSONS, "mp_render_rays: P out of range");
  MP_REQUIRE(R > 0, "mp_render_rays: R must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  const mp_sampler_cfg_t& c = scene->sampler;
  const int n = c.N_samples + c.N_samples_extra + 1;     // multiply.py:290-292
  const float beta = fabsf(c.beta_param) + c.beta_min;
  const int prune = prune_is_exact(beta) ? 1 : 0;
  Arena a(workspace, workspace_bytes);
  RenderWs w;
  MP_REQUIRE(render_carve(a, *scene, R, w), "mp_render_rays: workspace too small (%zu needed, %zu given)", a.off,
             workspace_bytes);
  MP_TRY(mp_camera_rays(uv, pose, intrinsics, R, w.dirs, w.cam, stream));     // multiply.py:223-227
  if (g_streams_on < 0) {
    const char* e = getenv("MP_RENDER_STREAMS");
    g_streams_on = e ? (atoi(e) != 0) : 1;
  }
  const bool fork = g_streams_on == 1;
  const cudaStream_t caller = st;
  if (fork) {
    MP_TRY(branch_streams_init());
    MP_CHECK_CUDA(cudaEventRecord(g_bs.fork, caller));
  }
  // background branch first: its MLP launch is the longest independent piece (multiply.py:514-541)
  const float* bg = nullptr;
  if (scene->bg_field) {
    cudaStream_t sb = fork ? g_bs.s[scene->P] : caller;
    if (fork) MP_CHECK_CUDA(cudaStreamWaitEvent(sb, g_bs.fork, 0));
    MP_TRY(render_background(scene->bg_field->f, w.dirs, w.cam, R, c.scene_bounding_sphere, w.bg, w.sub[scene->P],
                             w.sub_bytes, sb));
    if (fork) MP_CHECK_CUDA(cudaEventRecord(g_bs.join[scene->P], sb));
    bg = w.bg;
  }
  CompositePersons cp;
  cp.P = scene->P;
  for (int p = 0; p < scene->P; ++p) {
    st = fork ? g_bs.s[p] : caller;
    if (fork) MP_CHECK_CUDA(cudaStreamWaitEvent(st, g_bs.fork, 0));
    MP_REQUIRE(scene->body[p] && scene->field[p] && scene->hit_index[p] && scene->hit_count[p] >= 1,
               "mp_render_rays: person %d incomplete", p);
    const Body& body = scene->body[p]->b;
    const Field& field = scene->field[p]->f;
    MP_REQUIRE(body.tfs, "mp_render_rays: body %d has no pose", p);
    const int Rp = scene->hit_count[p];
    PersonBufs& b = w.pb[p];
    gather_rays_kernel<<<div_up(Rp, 256), 256, 0, st>>>(w.dirs, w.cam, scene->hit_index[p], Rp, b.dirs, b.cam);
    MP_LAUNCH_CHECK();
    // ray_sampler.get_z_vals (multiply.py:285-289)
    MP_TRY(sample_rays(c, body, field, b.dirs, b.cam, Rp, b.z, nullptr, out->trips ? out->trips + p : nullptr, w.sub[p],
                       w.sub_bytes, st));
    // main pass (multiply.py:295-308, 403-404): deform, SDF, normals, colour
    MP_CHECK_CUDA(cudaMemsetAsync(b.count, 0, sizeof(int), st));
    MP_CHECK_CUDA(cudaMemsetAsync(b.rgb, 0, (size_t)Rp * n * 3 * sizeof(float), st));
    MP_CHECK_CUDA(cudaMemsetAsync(b.nrm, 0, (size_t)Rp * n * 3 * sizeof(float), st));
    MP_TRY(launch_deform_rays(body, b.dirs, b.cam, b.z, n + 1, nullptr, 0, n, Rp, prune, b.sdf, n, b.xc_list,
                              b.slot_list, b.count, b.outl, nullptr, st));
    MP_TRY(launch_forward_jac(body, b.xc_list, Rp * n, b.count, nullptr, b.jinv, 12, st));
    MP_TRY(field_shade_list(field, b.xc_list, b.slot_list, b.count, Rp * n, b.jinv, b.sdf, b.rgb, b.nrm, nullptr,
                            nullptr, w.sub[p], w.sub_bytes, st));
    if (!prune) {
      force_outlier_sdf_kernel<<<div_up(Rp * n, 256), 256, 0, st>>>(b.outl, Rp * n, b.sdf);
      MP_LAUNCH_CHECK();
    }
    MP_TRY(launch_row_of_ray(scene->hit_index[p], Rp, R, b.row_of_ray, st));
    cp.n_rows[p] = Rp;
    cp.row_of_ray[p] = b.row_of_ray;
    cp.z[p] = b.z;
    cp.sdf[p] = b.sdf;
    cp.rgb[p] = b.rgb;
    cp.nrm[p] = b.nrm;
    auto tap = [&](float* dst, const float* src, size_t cnt) -> int {
      if (dst) MP_CHECK_CUDA(cudaMemcpyAsync(dst, src, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
      return 0;
    };
    MP_TRY(tap(out->z_vals[p], b.z, (size_t)Rp * (n + 1)));
    MP_TRY(tap(out->sdf[p], b.sdf, (size_t)Rp * n));
    MP_TRY(tap(out->rgb[p], b.rgb, (size_t)Rp * n * 3));
    MP_TRY(tap(out->normals[p], b.nrm, (size_t)Rp * n * 3));
    if (fork) MP_CHECK_CUDA(cudaEventRecord(g_bs.join[p], st));
  }
  st = caller;
  if (fork) {
    for (int p = 0; p < scene->P; ++p) MP_CHECK_CUDA(cudaStreamWaitEvent(caller, g_bs.join[p], 0));
    if (scene->bg_field) MP_CHECK_CUDA(cudaStreamWaitEvent(caller, g_bs.join[scene->P], 0));
  }
  float* normal = out->normal_values ? out->normal_values : w.nrm;
  float* acc = out->acc_map ? out->acc_map : w.acc;
  float* accp = out->acc_person_list ? out->acc_person_list : w.accp;
  float* bgT = out->bg_T ? out->bg_T : w.bgT;
  MP_TRY(launch_composite(cp, R, n, beta, w.fg, normal, acc, accp, bgT, st));     // multiply.py:427-480
  MP_REQUIRE(out->rgb_values, "mp_render_rays: rgb_values output is required");
  MP_TRY(launch_final_compose(w.fg, bgT, bg, R, out->rgb_values, out->fg_rgb_values, st));   // :544-545, :590
  return 0;
}
}

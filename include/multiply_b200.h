/*
 * multiply_b200 — C ABI of the B200-native (sm_100a) MultiPly volume-rendering hot path.
 *
 * The reference (eth-ait/MultiPly) has no FFI / plugin registry: its boundary for this path
 * is the Python operator surface of code/lib/model.  Each entry point below
 * names the reference function it replaces (file:line relative to /root/reference/code).
 * INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - every `const float*` / `float*` / `int*` is a CUDA DEVICE pointer owned by the caller
 *     unless the parameter name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and never call
 *     cudaDeviceSynchronize.
 *   - return 0 on success, negative on error; mp_last_error() gives the (thread-local) text.
 *   - fp32 row-major contiguous tensors; B = 1 (the reference indexes [0] everywhere,
 *     multiply.py:208, deformer.py:22-24).
 *   - scratch memory comes from caller-provided workspaces sized by the *_workspace_bytes calls.
 */
#ifndef MULTIPLY_B200_H
#define MULTIPLY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP_MAX_LAYERS 12
#define MP_MAX_PERSONS 8
#define MP_NUM_JOINTS 24

/* ------------------------------------------------------------------------------------------
 * misc
 * ---------------------------------------------------------------------------------------- */
int mp_version(void);
const char* mp_last_error(void);
/* number of SMs of the current device (grid sizing) */
int mp_device_sm_count(void);
/* host helper: torch.linspace(start, end, n) in fp32, bit-exact with the ATen CPU kernel
 * (fma(step, i, start) for i < n/2, fma(-step, n-1-i, end) otherwise).  ray_sampler.py:29,167,204,215 */
int mp_linspace_host(float start, float end, int n, float* out_host);
/* total number of this library's kernels launched since the last reset (bench.py "gpu_launches") */
long long mp_launch_count(int reset);

/* ------------------------------------------------------------------------------------------
 * networks: ImplicitNet / RenderingNet  (lib/model/networks.py:7-208, 223-312)
 * ---------------------------------------------------------------------------------------- */
typedef struct mp_net mp_net_t;

/* Raw parameters of a stack of nn.Linear layers exactly as they sit in the state dict
 * (`lin{l}.weight_v`, `lin{l}.weight_g`, `lin{l}.bias`; `lin{l}.weight` when weight_g == NULL). */
typedef struct {
  int n_layers;
  const float* weight_v[MP_MAX_LAYERS]; /* [out,in] */
  const float* weight_g[MP_MAX_LAYERS]; /* [out,1] or NULL (no weight norm) */
  const float* bias[MP_MAX_LAYERS];     /* [out] */
  int in_dim[MP_MAX_LAYERS];
  int out_dim[MP_MAX_LAYERS];
} mp_linear_stack_t;

/* ImplicitNet description (networks.py:7-116). */
typedef struct {
  mp_linear_stack_t lin;   /* 9 layers for the shipped configs */
  int d_in;                /* 3 (fg) or 4 (bg) */
  int multires;            /* 6 (fg) or 10 (bg) ; embedding dim = d_in*(1+2*multires) */
  int cond_dim;            /* 69 ('smpl') or 32 ('frame'); cond is concatenated at layer 0 */
  int skip_layer;          /* 4 */
} mp_implicit_desc_t;

/* RenderingNet description (networks.py:223-262). mode 0 = 'pose_no_view', 1 = 'nerf_frame_encoding'. */
typedef struct {
  mp_linear_stack_t lin;   /* lin0..lin{n-1} */
  int mode;
  int multires_view;       /* -1 or 4 */
  const float* lin_pose_weight; /* [8,69] (mode 0) or NULL */
  const float* lin_pose_bias;   /* [8] */
} mp_render_desc_t;

/* A foreground field = ImplicitNet + RenderingNet of one person; background field = bg pair.
 * Packing folds weight-norm (networks.py:82-83), the 1/sqrt(2) of the skip layer (:166-167) and
 * lays the weights out for the kernels (fp32 transposed for the SIMT engine, fp16 hi/lo
 * swizzled K-major tiles for the tcgen05 engine).  The handle is immutable afterwards. */
size_t mp_field_pack_bytes(void);
int mp_field_pack(const mp_implicit_desc_t* imp, const mp_render_desc_t* ren, int is_background,
                  void* storage, size_t storage_bytes, mp_net_t** out, void* stream);
void mp_field_free(mp_net_t* f);
/* Per-call conditioning: folds cond (pose[3:]/pi, multiply.py:270, or the frame code, :407-410) into
 * the layer-0 bias and lin_pose(body_pose) (networks.py:277-281) / frame code into the colour layer-0 bias. */
int mp_field_set_cond(mp_net_t* f, const float* cond /*[cond_dim]*/, void* stream);

/* engine selection: 0 = fp32 SIMT (validation engine), 1 = tcgen05 split-fp16 tensor-core engine */
int mp_set_engine(int engine);
int mp_get_engine(void);
/* Precision mode of the tcgen05 engine: which split-precision product terms each MLP layer issues (fp16 hi/lo operand
 * pairs, fp32 accumulation).  0 = parity (default): A_hi.W_hi + A_lo.W_hi + A_hi.W_lo everywhere (RGB / SDF within 1e-4 of
 * the fp32 reference); 1 = the colour layers issue A_hi.W_hi only (SDF / normals unchanged, RGB ~2e-5); 2 = throughput:
 * every layer single-term, i.e. plain fp16 operands — outside the 1e-4 gate, reported separately. */
int mp_set_precision(int mode);
int mp_get_precision(void);
/* mp_render_rays schedule: 1 (default) = persons and background on their own streams, joined before the compositor;
 * 0 = everything on the caller's stream (used for per-kernel timing).  Environment override: MP_RENDER_STREAMS. */
int mp_set_streams(int on);

/* Per-launch timing of the tcgen05 MLP kernel (CUDA events on the launching stream), by program kind:
 * [0] sdf-only, [1] forward (sdf + features), [2] full shade, [3] background.  mp_profile_read synchronises
 * on the recorded events and returns summed milliseconds, launch counts and processed points (host arrays of 4). */
int mp_profile_enable(int on);
int mp_profile_read(double* ms_host, long long* launches_host, double* points_host, int reset);
/* Diagnostics: cycle stamps of one tile of CTA 0 of the last tcgen05 launch (recorded when MP_TC_KNOBS has bit 1 set):
 * out[s*8 + 0..6] epilogue warp (step start, accumulator ready, chunk 0..3 done, step end), out[2048 + s*8 + 0..4] MMA
 * issuer (operand K-block 0..3 ready, commit).  scripts/gpu_trace.py prints them. */
int mp_tc_trace_read(unsigned long long* out, int n);

/* ImplicitNet.forward (networks.py:126-208): x [N,d_in] -> out [N,257] (sdf | feature).
 * sdf / feat may be NULL.  Replaces `self.foreground_implicit_network_list[p](x_c, cond)`. */
int mp_implicit_forward(mp_net_t* f, const float* x, int N, float* sdf /*[N]*/, float* feat /*[N,256]*/,
                        void* workspace, size_t workspace_bytes, void* stream);
/* forward + d sdf / d x (replaces the autograd.grad at multiply.py:653-659) */
int mp_implicit_forward_grad(mp_net_t* f, const float* x, int N, float* sdf, float* feat, float* grad /*[N,3]*/,
                             void* workspace, size_t workspace_bytes, void* stream);
/* RenderingNet.forward 'pose_no_view' (networks.py:263-312): -> rgb [N,3] */
int mp_render_forward(mp_net_t* f, const float* points, const float* normals, const float* feat, int N,
                      float* rgb, void* workspace, size_t workspace_bytes, void* stream);
size_t mp_mlp_workspace_bytes(int N);
/* Background pair at given points (multiply.py:523-526): bg_implicit_network(pts [N,4], {'frame': code}) ->
 * sdf [N] (may be NULL) and bg_rendering_network(None, None, view_dirs [N,3], None, feature, code) -> rgb [N,3].
 * The frame code is the field's cond (mp_field_set_cond).  Workspace: mp_mlp_workspace_bytes(N). */
int mp_bg_nets_forward(mp_net_t* bg_field, const float* pts, const float* view_dirs, int N, float* sdf, float* rgb,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Canonical SDF on the dense lattice of lib/utils/mesh.py:generate_mesh (:78-105; the values MISE's octree queries
 * through Multiply.query_oc, multiply.py:169-172, batch by batch): values[(ix*(res+1)+iy)*(res+1)+iz] =
 * ImplicitNet(p)[0] with p = ((idx/res - 0.5) * pad) * extent + centre (fp32, rounded step by step as numpy does;
 * pad = 1.1, extent = the longest side of the SMPL bounds).  The pose conditioning is the field's cond
 * (mp_field_set_cond).  The lattice is generated on the device and streamed through the sdf-only MLP program in
 * 2^20-point slabs. */
size_t mp_sdf_grid_workspace_bytes(int res);
int mp_sdf_grid(mp_net_t* field, const float* center_host /*[3]*/, float extent, float pad, int res,
                float* values /*[(res+1)^3]*/, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * deformer: SMPLDeformer (lib/model/deformer.py:6-89)
 * ---------------------------------------------------------------------------------------- */
typedef struct mp_body mp_body_t;
size_t mp_body_bytes(int V);
/* verts_cano [V,3], weights [V,24]: SMPLDeformer.smpl_verts / smpl_weights (deformer.py:16-17).
 * Builds the canonical-space vertex grid once. */
int mp_body_create(const float* verts_cano, const float* weights, int V, float cano_cell,
                   void* storage, size_t storage_bytes, mp_body_t** out, void* stream);
void mp_body_free(mp_body_t* b);
/* per frame: posed verts [V,3] (smpl_output['smpl_verts']) and bone transforms [24,4,4] (smpl_tfs);
 * rebuilds the posed-space vertex grid (cell >= 0.1 = the outlier radius of deformer.py:49). */
int mp_body_set_pose(mp_body_t* b, const float* verts_posed, const float* tfs, void* stream);

/* SMPLDeformer.forward(x, smpl_tfs, return_weights=False, inverse=True, smpl_verts) (deformer.py:19-30):
 * x [N,3] -> x_c [N,3], outlier [N] (uint8).  exact_far != 0 also resolves the exact nearest
 * vertex of points farther than the grid radius (needed only when outliers are not pruned). */
int mp_deform_inverse(mp_body_t* b, const float* x, int N, float* x_c, uint8_t* outlier, int exact_far,
                      void* stream);
/* SMPLDeformer.forward_skinning (deformer.py:31-35) + the Jacobian the reference obtains by three
 * autograd VJPs (multiply.py:625-640): x_c [N,3] -> x_d [N,3] (may be NULL), Jinv [N,9] = inverse of
 * the upper-left 3x3 of sum_j w_j tfs_j with w from the nearest CANONICAL vertex. */
int mp_deform_forward_jac(mp_body_t* b, const float* x_c, int N, float* x_d, float* Jinv, void* stream);

/* Optional root finder (BASELINE.json north_star: "Broyden-root-finds canonical points").  The
 * reference has NO such step (its deformer is KNN + closed-form inverse LBS, deformer.py:19-50), so
 * this is non-default and checked against its own CPU restatement (oracle/port.py:deform_broyden), not against MultiPly.
 * Solves forward_skinning(x_c) = x (deformer.py:31-35: weights of the nearest CANONICAL vertex) by Broyden's method
 * started from the closed-form inverse (weights of the nearest POSED vertex), J^-1 initialised with the inverse
 * blended 3x3 at the start point, at most max_steps rank-one updates, lowest-residual iterate returned.
 *   x [N,3] -> x_c [N,3]; residual [N] = |forward_skinning(x_c) - x| (NULL ok); converged [N] = residual <
 *   cvg_threshold (NULL ok); outlier [N] as mp_deform_inverse (NULL ok); steps [N] iterations taken (NULL ok). */
int mp_deform_broyden(mp_body_t* b, const float* x, int N, int max_steps, float cvg_threshold, float* x_c,
                      float* residual, uint8_t* converged, uint8_t* outlier, int* steps, void* stream);
/* max_steps > 0 makes every inverse-deformer call on this body (mp_deform_inverse, mp_sdf_with_deformer, the sampler
 * and the main pass of mp_render_rays) refine its non-outlier points this way; 0 (the default) = reference behaviour. */
int mp_body_set_root_finder(mp_body_t* b, int max_steps, float cvg_threshold);

/* ------------------------------------------------------------------------------------------
 * density: LaplaceDensity (lib/model/density.py:11-29)
 * ---------------------------------------------------------------------------------------- */
int mp_laplace_density(const float* sdf, int N, float beta, float* sigma, void* stream);

/* ------------------------------------------------------------------------------------------
 * rays: rend_util.get_camera_params / get_sphere_intersections (lib/utils/rend_util.py:45-87,131-147)
 * ---------------------------------------------------------------------------------------- */
int mp_camera_rays(const float* uv /*[R,2]*/, const float* pose /*[4,4]*/, const float* intrinsics /*[4,4]*/,
                   int R, float* ray_dirs /*[R,3]*/, float* cam_loc /*[R,3]*/, void* stream);
/* status_flag (device int) is set to 1 if any ray misses the sphere (the reference calls exit(), :140-142) */
int mp_sphere_intersections(const float* cam_loc, const float* ray_dirs, int R, float r,
                            float* near_far /*[R,2]*/, int* status_flag, void* stream);

/* Ray / box culling (replaces the host-side trimesh ray/triangle test on the x1.2 oriented box,
 * multiply.py:208-214, :256-263): box = centre + half extents (host doubles) + optional 3x3 rotation (device
 * doubles, rows = box axes, NULL = axis aligned).  idx_out [R] receives the hit ray ids in ascending order,
 * count_dev their number.  An empty result is the caller's to replace by ray 0 (multiply.py:262-263). */
int mp_ray_box_hits(const float* cam_loc, const float* ray_dirs, int R, const double* center_host,
                    const double* half_extent_host, const double* rot_dev, int64_t* idx_out, int* count_dev,
                    void* stream);
/* multiply.py:262-263 on the device: an empty hit list becomes the single ray 0 (idx[0] = 0, *count_dev = 1). */
int mp_hit_list_finalize(int64_t* idx, int* count_dev, void* stream);
/* Same test with the box taken from the posed vertices on the device (no host read of the vertices): centre and half
 * extents of the axis-aligned bounds of verts [V,3], inflated by `inflate` (1.2, multiply.py:212).  box_ws: >= 64 bytes
 * of device scratch. */
int mp_ray_aabb_hits(const float* cam_loc, const float* ray_dirs, int R, const float* verts, int V, double inflate,
                     int64_t* idx_out, int* count_dev, void* box_ws, void* stream);

/* ------------------------------------------------------------------------------------------
 * SMPL server: SMPLServer.forward (lib/model/smpl.py:50-95) -> SMPL.forward (lib/smpl/body_models.py:278-364)
 * -> lbs (lib/smpl/lbs.py:136-229).  Model arrays are the ones the SMPL pkl provides (device pointers, kept by
 * reference); parents is a host array of 24 ints.  The canonical pose of smpl.py:35-47 is evaluated at creation.
 * ---------------------------------------------------------------------------------------- */
typedef struct mp_smpl mp_smpl_t;
size_t mp_smpl_bytes(int V);
int mp_smpl_create(const float* v_template /*[V,3]*/, const float* shapedirs /*[V,3,10]*/,
                   const float* posedirs /*[207,V*3]*/, const float* J_regressor /*[24,V]*/,
                   const int* parents_host /*[24]*/, const float* lbs_weights /*[V,24]*/, int V,
                   const float* betas_canonical /*[10] device or NULL*/, void* storage, size_t storage_bytes,
                   mp_smpl_t** out, void* stream);
void mp_smpl_free(mp_smpl_t* s);
/* SMPLServer.verts_c [V,3] and tfs_c_inv [24,4,4] (either may be NULL) */
int mp_smpl_canonical(mp_smpl_t* s, float* verts_c, float* tfs_c_inv, void* stream);
/* scale [1], transl [3], thetas [72], betas [10] (device) -> smpl_verts [V,3], smpl_tfs [24,4,4] */
int mp_smpl_forward(mp_smpl_t* s, const float* scale, const float* transl, const float* thetas, const float* betas,
                    int absolute, float* smpl_verts, float* smpl_tfs, void* stream);

/* ------------------------------------------------------------------------------------------
 * sampler: ErrorBoundSampler (lib/model/ray_sampler.py:45-230), eval mode
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  float scene_bounding_sphere; /* 3.0 (multiply.py:85) */
  float near;                  /* 0.0 */
  int N_samples;               /* S */
  int N_samples_eval;          /* E */
  int N_samples_extra;         /* X */
  float eps;                   /* 0.1 */
  int beta_iters;              /* 10 */
  int max_total_iters;         /* 5 */
  float add_tiny;              /* 1e-6 */
  float beta_param;            /* density.beta parameter; beta = |beta_param| + beta_min */
  float beta_min;              /* 1e-4 */
} mp_sampler_cfg_t;

size_t mp_sampler_workspace_bytes(const mp_sampler_cfg_t* cfg, int R);
/* ErrorBoundSampler.get_z_vals(ray_dirs, cam_loc, model, cond, smpl_tfs, eval_mode=True, smpl_verts, person_id)
 * (ray_sampler.py:66-220).  The SDF callback of the reference (`model.sdf_func_with_smpl_deformer`)
 * is the (body, field) pair.  Outputs z_vals [R, S+X+2], z_bg [R,32] (may be NULL); trips_out (device
 * int, may be NULL) receives the number of Algorithm-1 iterations executed.  No host sync. */
int mp_sample_rays(const mp_sampler_cfg_t* cfg, mp_body_t* body, mp_net_t* field,
                   const float* ray_dirs, const float* cam_loc, int R,
                   float* z_vals, float* z_bg, int* trips_out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Training-mode get_z_vals (model.training: ray_sampler.py:32-40 stratified start samples, :171 random abscissae of the
 * final set, :202 randperm extras, :212-213 the eikonal pick, :216 jittered inverse-sphere depths; the SDF callback does
 * not clamp outliers, multiply.py:142).  Every random draw of the reference is an INPUT so that a caller can replay the
 * reference's RNG stream.  The draws made after the Algorithm-1 loop depend on the number of trips T+1 the loop took
 * (randperm((T+1) E), and the generator state after it), which only the device knows: they are passed for EVERY possible
 * T and the kernels pick the row of the trip the loop ended on.  T_max = cfg->max_total_iters. */
typedef struct {
  const float* t_rand;     /* [R, E]                      torch.rand of UniformSampler.get_z_vals */
  const float* u_final;    /* [R, S]                      torch.rand at the final inverse-CDF step */
  const int* extra_perm;   /* [T_max, T_max*E] int32      row T: torch.randperm((T+1)*E), first X entries are used */
  const int* eik_idx;      /* [T_max, R] int32            row T: torch.randint(S+X+2, (R,)) */
  const float* t_rand_bg;  /* [T_max, R, 32] or NULL      row T: torch.rand of the inverse-sphere UniformSampler */
} mp_sampler_rng_t;
/* outputs as mp_sample_rays plus z_eik [R] (z_samples_eik, may be NULL) */
int mp_sample_rays_train(const mp_sampler_cfg_t* cfg, mp_body_t* body, mp_net_t* field,
                         const float* ray_dirs, const float* cam_loc, int R, const mp_sampler_rng_t* rng,
                         float* z_vals, float* z_bg, float* z_eik, int* trips_out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Multiply.sdf_func_with_smpl_deformer (multiply.py:137-151, eval): x [N,3] -> sdf [N] (4.0 on outliers),
 * x_c [N,3], feat [N,256] (may be NULL). */
int mp_sdf_with_deformer(mp_body_t* body, mp_net_t* field, const float* x, int N,
                         float* sdf, float* x_c, float* feat,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * compositing (multiply.py:427-480 with nerfacc; 682-696 background)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int n_rows;               /* rays in this person's hit list (R_p) */
  const int64_t* ray_index; /* [R_p] sorted ray ids (index_ray_box, multiply.py:256-263) */
  const float* z_vals;      /* [R_p, n+1] (last column = z_max) */
  const float* sdf;         /* [R_p, n] */
  const float* rgb;         /* [R_p, n, 3] */
  const float* normal;      /* [R_p, n, 3] */
} mp_person_samples_t;

size_t mp_composite_workspace_bytes(int R, int P);
/* outputs: fg_rgb [R,3], normal [R,3], acc [R], acc_person [R,P], bg_T [R] */
int mp_composite(const mp_person_samples_t* persons_host, int P, int R, int n, float beta,
                 float* fg_rgb, float* normal, float* acc, float* acc_person, float* bg_T,
                 void* workspace, size_t workspace_bytes, void* stream);

/* rgb_values = fg_rgb + bg_T * bg_rgb (bg_rgb NULL -> white, multiply.py:540-545), fg_rgb_values = fg_rgb + bg_T * 1
 * (multiply.py:590); fg_rgb_values may be NULL.  The last stage of Multiply.forward, exported for callers that composite
 * ray blocks themselves (person-sharded rendering, multiply_b200/parallel.py). */
int mp_final_compose(const float* fg_rgb /*[R,3]*/, const float* bg_T /*[R]*/, const float* bg_rgb /*[R,3] or NULL*/, int R,
                     float* rgb_values /*[R,3]*/, float* fg_rgb_values /*[R,3] or NULL*/, void* stream);

/* background: inverse-sphere samples -> depth2pts_outside (multiply.py:698-726) -> bg nets -> bg_volume_rendering */
size_t mp_background_workspace_bytes(int R);
int mp_background(mp_net_t* bg_field, const float* ray_dirs, const float* cam_loc, int R, float bound_r,
                  float* bg_rgb /*[R,3]*/, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * the fused entry used by Multiply.forward (multiply.py:174-598, eval branch)
 * ---------------------------------------------------------------------------------------- */
/* Training-mode forward VALUES (multiply.py:174-598 with self.training, shipped loss weights, current_epoch >= 250):
 * stochastic sampling per person (mp_sampler_rng_t), no outlier clamp in the SDF callback or the main pass
 * (multiply.py:142 is eval-only), jittered inverse-sphere depths of the background pass (the second
 * inverse_sphere_sampler.get_z_vals call, :482).  Gradients are NOT produced. */
typedef struct {
  const mp_sampler_rng_t* rng[MP_MAX_PERSONS]; /* host structs of device pointers, one per person */
  float* z_eik[MP_MAX_PERSONS];                /* [R_p] z_samples_eik out, or NULL */
  const float* t_rand_bg;                      /* [R,32] torch.rand of the background's UniformSampler, or NULL */
} mp_train_t;

typedef struct {
  mp_sampler_cfg_t sampler;
  int P;
  mp_body_t* body[MP_MAX_PERSONS];
  mp_net_t* field[MP_MAX_PERSONS];
  mp_net_t* bg_field;              /* NULL -> white background (multiply.py:540-541) */
  const int64_t* hit_index[MP_MAX_PERSONS]; /* device, sorted ray ids per person */
  int hit_count[MP_MAX_PERSONS];            /* >=1 (the reference substitutes ray 0 for an empty list) */
  /* Optional device-side row counts (GPU culling without a host round trip, mp_ray_box_hits + mp_hit_list_finalize):
   * when hit_count_dev[p] != NULL, hit_count[p] is the CAPACITY of hit_index[p] (normally R) and the number of valid
   * rows is read on the device by every kernel of person p's branch.  NULL -> hit_count[p] is exact. */
  const int* hit_count_dev[MP_MAX_PERSONS];
  const mp_train_t* train;                  /* NULL: eval mode */
} mp_scene_t;

typedef struct {
  float* rgb_values;      /* [R,3] */
  float* fg_rgb_values;   /* [R,3] */
  float* normal_values;   /* [R,3] */
  float* acc_map;         /* [R] */
  float* acc_person_list; /* [R,P] */
  /* optional debug taps (NULL to skip): per person p, [R_p, n+1], [R_p, n], [R_p,n,3], [R_p,n,3] */
  float* z_vals[MP_MAX_PERSONS];
  float* sdf[MP_MAX_PERSONS];
  float* rgb[MP_MAX_PERSONS];
  float* normals[MP_MAX_PERSONS];
  int* trips;             /* device [P] or NULL */
  float* bg_T;            /* [R] or NULL */
  int* status;            /* device int or NULL: bit 0 = some ray misses the bounding sphere (the reference prints
                             'BOUNDING SPHERE PROBLEM' and exits, rend_util.py:140-142; its pixels are undefined here) */
} mp_render_out_t;

size_t mp_render_workspace_bytes(const mp_scene_t* scene, int R);
/* uv [R,2], pose [4,4], intrinsics [4,4] (device).  Replaces Multiply.forward(input) in eval mode. */
int mp_render_rays(const mp_scene_t* scene, const float* uv, const float* pose, const float* intrinsics,
                   int R, const mp_render_out_t* out, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MULTIPLY_B200_H */

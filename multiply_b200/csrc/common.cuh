// Shared declarations of the multiply_b200 CUDA library (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "../../include/multiply_b200.h"

namespace mp {

// ---- error handling (no exceptions across the ABI) ----------------------------------------
void set_error(const char* fmt, ...);
extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

#define MP_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      mp::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,                \
                    cudaGetErrorString(_e));                                             \
      return -2;                                                                         \
    }                                                                                    \
  } while (0)

#define MP_REQUIRE(cond, ...)                                                            \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      mp::set_error(__VA_ARGS__);                                                        \
      return -1;                                                                         \
    }                                                                                    \
  } while (0)

#define MP_LAUNCH_CHECK()                                                                \
  do {                                                                                   \
    mp::g_launches++;                                                                    \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      mp::set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__,                \
                    cudaGetErrorString(_e));                                             \
      return -3;                                                                         \
    }                                                                                    \
  } while (0)

#define MP_TRY(expr)                                                                     \
  do {                                                                                   \
    int _r = (expr);                                                                     \
    if (_r != 0) return _r;                                                              \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over a caller-provided workspace
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  bool ok;
  Arena(void* p, size_t n) : base((char*)p), cap(n), off(0), ok(true) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    size_t bytes = n * sizeof(T);
    if (base == nullptr || off + bytes > cap) {
      ok = false;
      off += bytes;
      return nullptr;
    }
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
};

int sm_count();

// ---- device structures --------------------------------------------------------------------

// uniform vertex grid for exact nearest-vertex queries (deform.cu)
struct GridHeader {
  float lo[3];
  float inv_h;
  float h;
  int dim[3];
  int ncell;
  int R0;        // search radius in cells that covers the query radius the grid was built for (h * R0 >= radius)
};
constexpr int kMaxCells = 32768;

struct Body {
  int V;
  const float* weights;      // [V,24]
  const float* verts_cano;   // [V,3]
  const float* verts_posed;  // [V,3] (caller memory, valid for the frame)
  const float* tfs;          // [24,4,4]
  float cano_cell;
  // grids (device, inside the body's storage)
  GridHeader* cano_hdr;
  int* cano_cell_start;      // [kMaxCells+1]
  float4* cano_sorted;       // [V]
  GridHeader* posed_hdr;
  int* posed_cell_start;
  float4* posed_sorted;
  int* scratch;              // [kMaxCells + 8]
  float4* vert_tf;           // [V][3] per-vertex inverse blended transform: rows (I_r0, I_r1, I_r2, c_r), x_c = I (x - c)
  // optional root finder (mp_body_set_root_finder): 0 steps = the reference's closed-form inverse only
  int root_steps;
  float root_thr;
};

// packed network (mlp_pack.cu)
constexpr int kHidden = 256;
struct Field {
  int is_bg;
  int d_in, multires, emb_dim, cond_dim, skip_layer, n_imp;   // implicit
  int imp_in[MP_MAX_LAYERS], imp_out[MP_MAX_LAYERS];
  int n_ren, ren_mode, multires_view;
  int ren_in[MP_MAX_LAYERS], ren_out[MP_MAX_LAYERS];
  int ren_extra;          // leading inputs of colour layer 0 handled outside the 256-wide feature block
  // fp32 SIMT layout: Wt[l] is [in][out] (transposed), folded weight norm / skip scale
  float* imp_W[MP_MAX_LAYERS];      // natural [out][in] (backward pass, tcgen05 packing)
  float* ren_W[MP_MAX_LAYERS];
  float* imp_Wt[MP_MAX_LAYERS];
  float* imp_b[MP_MAX_LAYERS];     // layer 0: raw bias; imp_b0_eff has the cond folded in
  float* imp_W0cond;               // [cond_dim][out0] transposed cond columns of layer 0
  float* imp_b0_eff;               // [out0]
  float* ren_Wt[MP_MAX_LAYERS];
  float* ren_b[MP_MAX_LAYERS];
  float* ren_b0_eff;               // [out0] (lin_pose(cond) / frame code folded)
  float* ren_W0cond;               // [cdim][out0] : (W0[:, 6:14] @ lin_pose.W) for mode 0 ([69][out0]); W0[:,27:59]^T for mode 1 ([32][out0])
  float* ren_b0_base;              // [out0] : b0 + W0[:,6:14] @ lin_pose.b  (mode 0) ; b0 (mode 1)
  int ren_cond_dim;
  float* ren_cb;                   // [out0] Wc0[:, feat] . b8[1:]  (colour layer 0 folded onto the feature layer)
  float* ren_b0_fold;              // [out0] ren_b0_eff + ren_cb : bias of the folded layer (tcgen05 chains)
  // tcgen05 engine blobs (mlp_tc.cu); null until packed
  void* tc;
  char* storage;
  size_t storage_bytes;
};

// launch helpers
static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline int clamp_wpc(size_t v) { return v < 1 ? 1 : (v > 8 ? 8 : (int)v); }

// engines (mlp_simt.cu, mlp_tc.cu)
size_t simt_workspace_bytes(int N);
int simt_sdf_list(const Field& f, const float* xc_list, const int* slot_list, const int* count_dev, int cap,
                  float* sdf_out, void* ws, size_t ws_bytes, cudaStream_t st);
int simt_shade_list(const Field& f, const float* xc_list, const int* slot_list, const int* count_dev, int cap,
                    const float* Jinv_list, float* sdf_out, float* rgb_out, float* normal_out, float* grad_out,
                    float* feat_out, void* ws, size_t ws_bytes, cudaStream_t st);
int simt_bg(const Field& f, const float* pts, const float* dirs, int N, float* sdf, float* rgb, void* ws,
            size_t ws_bytes, cudaStream_t st);
int simt_render(const Field& f, const float* pts, const float* nrm, const float* feat, int N, float* rgb, void* ws,
                size_t ws_bytes, cudaStream_t st);
extern std::atomic<int> g_engine;

// cross-file launchers
int launch_deform_rays(const Body& b, const float* dirs, const float* cam, const float* z, int z_stride,
                       const int* zpos, int zpos_stride, int n_per_ray, int R, int prune, float* sdf_out,
                       int sdf_stride, float* xc_list, int* slot_list, int* count, uint8_t* outlier_out,
                       const int* active, cudaStream_t st, const int* R_dev = nullptr);
int launch_forward_jac(const Body& b, const float* x_c, int N, const int* n_dev, float* x_d, float* Jinv,
                       int jstride, cudaStream_t st);

}  // namespace mp

struct mp_body {
  mp::Body b;
};
struct mp_net {
  mp::Field f;
};

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__
namespace mp {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// inclusive warp scan
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive warp scan (no "inclusive minus own" cancellation when one lane holds a huge term)
__device__ __forceinline__ float warp_scan_excl(float v, int lane) {
  float incl = warp_scan_incl(v, lane);
  float up = __shfl_up_sync(0xffffffffu, incl, 1);
  return lane == 0 ? 0.f : up;
}

// LaplaceDensity.density_func (lib/model/density.py:20-25):
//   alpha * (0.5 + 0.5 * sign(sdf) * expm1(-|sdf| / beta)),  alpha = 1 / beta
__device__ __forceinline__ float laplace_density(float sdf, float beta) {
  float alpha = __fdiv_rn(1.0f, beta);
  float sg = (sdf > 0.f) ? 1.f : ((sdf < 0.f) ? -1.f : 0.f);
  float e = expm1f(__fdiv_rn(-fabsf(sdf), beta));
  float t = __fmul_rn(__fmul_rn(0.5f, sg), e);
  return __fmul_rn(alpha, __fadd_rn(0.5f, t));
}

// torch.nn.Softplus(beta=100, threshold=20) (networks.py:85)
__device__ __forceinline__ float softplus100(float x) {
  float t = 100.f * x;
  if (t > 20.f) return x;
  return log1pf(expf(t)) / 100.f;
}
// d softplus100 / dx = sigmoid(100 x)  (1 above the threshold, as autograd does)
__device__ __forceinline__ float softplus100_grad(float x) {
  float t = 100.f * x;
  if (t > 20.f) return 1.f;
  float z = expf(t);
  return z / (z + 1.f);
}

}  // namespace mp
#endif

// SMPL body model server: shape/pose blend shapes, forward kinematics, linear blend skinning.
//   reference: /root/reference/code/lib/model/smpl.py:50-95 (SMPLServer.forward)
//              -> lib/smpl/body_models.py:278-364 (SMPL.forward) -> lib/smpl/lbs.py:136-229 (lbs),
//              :276-307 (batch_rodrigues), :323-378 (batch_rigid_transform)
// The reference issues ~100 tiny ATen kernels per person per forward (launch-bound); here it is two
// launches: one CTA does the 10-coefficient shape blend, the joint regression, Rodrigues and the 24-joint
// kinematic chain (plus SMPLServer's scale / translation / canonical-inverse), then a grid-wide kernel does the
// 207-term pose blend and the skinning per vertex with coalesced reads of `posedirs`.
#include "common.cuh"

namespace mp {

struct Smpl {
  int V;
  const float* v_template;   // [V,3]
  const float* shapedirs;    // [V,3,10]
  const float* posedirs;     // [207, V*3]
  const float* J_regressor;  // [24,V]
  const float* lbs_weights;  // [V,24]
  int parents[MP_NUM_JOINTS];
  // storage
  float* v_shaped;           // [V,3]
  float* pose_feature;       // [207]
  float* A_abs;              // [24,16] scaled/translated bone transforms w.r.t. theta = 0
  float* tfs_c_inv;          // [24,16]
  float* verts_c;            // [V,3]
  float* tmp_tfs;            // [24,16]
  float* J_t;                // [24,3]     J_regressor @ v_template           (precomputed at creation)
  float* J_s;                // [24,3,10]  J_regressor @ shapedirs: J(betas) = J_t + J_s . betas
};

__device__ __forceinline__ void mat4_mul(const float* a, const float* b, float* c) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) s = fmaf(a[4 * i + k], b[4 * k + j], s);
      c[4 * i + j] = s;
    }
}

// Joint regression is linear in the shape coefficients: J = Jr @ (v_template + shapedirs . betas) (lbs.py:184-188) =
// Jr @ v_template + (Jr @ shapedirs) . betas.  The two regressed tensors are computed once per model (one warp per joint),
// which takes the 24 x V regression -- 58 of the SMPL server's 90 us per call -- out of the per-frame path.
__global__ void __launch_bounds__(1024) smpl_jreg_kernel(Smpl m) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= MP_NUM_JOINTS) return;
  const float* jr = m.J_regressor + (size_t)warp * m.V;
  float acc[33];
#pragma unroll
  for (int k = 0; k < 33; ++k) acc[k] = 0.f;
  for (int v = lane; v < m.V; v += 32) {
    const float w = jr[v];
    if (w == 0.f) continue;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      acc[a] = fmaf(w, m.v_template[3 * v + a], acc[a]);
      const float* sd = m.shapedirs + ((size_t)3 * v + a) * 10;
#pragma unroll
      for (int l = 0; l < 10; ++l) acc[3 + a * 10 + l] = fmaf(w, sd[l], acc[3 + a * 10 + l]);
    }
  }
#pragma unroll
  for (int k = 0; k < 33; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0) {
    for (int a = 0; a < 3; ++a) {
      m.J_t[warp * 3 + a] = acc[a];
      for (int l = 0; l < 10; ++l) m.J_s[(warp * 3 + a) * 10 + l] = acc[3 + a * 10 + l];
    }
  }
}

// one CTA of 1024 threads
__global__ void __launch_bounds__(1024) smpl_pose_kernel(Smpl m, const float* __restrict__ scale_p,
                                                         const float* __restrict__ transl, const float* __restrict__ thetas,
                                                         const float* __restrict__ betas, int absolute,
                                                         float* __restrict__ tfs_out) {
  __shared__ float sJ[MP_NUM_JOINTS][3];
  __shared__ float sR[MP_NUM_JOINTS][9];
  __shared__ float sG[MP_NUM_JOINTS][16];
  __shared__ float sb[10];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 10) sb[tid] = betas[tid];
  __syncthreads();
  // J = J_regressor @ v_shaped (lbs.py:188, :232-249) through the precomputed regressions (smpl_jreg_kernel)
  if (tid < MP_NUM_JOINTS * 3) {
    float j = m.J_t[tid];
#pragma unroll
    for (int l = 0; l < 10; ++l) j = fmaf(sb[l], m.J_s[tid * 10 + l], j);
    sJ[tid / 3][tid % 3] = j;
  }
  // Rodrigues      lbs.py:276-307
  if (tid < MP_NUM_JOINTS) {
    float rx = thetas[3 * tid], ry = thetas[3 * tid + 1], rz = thetas[3 * tid + 2];
    float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
    float angle = sqrtf(ax * ax + ay * ay + az * az);
    float dx = rx / angle, dy = ry / angle, dz = rz / angle;
    float c = cosf(angle), s = sinf(angle);
    float K[9] = {0.f, -dz, dy, dz, 0.f, -dx, -dy, dx, 0.f};
    float KK[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        float t = 0.f;
        for (int k = 0; k < 3; ++k) t = fmaf(K[3 * i + k], K[3 * k + j], t);
        KK[3 * i + j] = t;
      }
    for (int i = 0; i < 9; ++i) {
      float ident = (i == 0 || i == 4 || i == 8) ? 1.f : 0.f;
      float r = ident + s * K[i] + (1.f - c) * KK[i];
      sR[tid][i] = r;
      // pose_feature = (rot_mats[1:] - I)      lbs.py:199
      if (tid >= 1) m.pose_feature[(tid - 1) * 9 + i] = r - ident;
    }
  }
  __syncthreads();
  // kinematic chain      lbs.py:323-378 ; then SMPLServer's scale / translation / canonical inverse  smpl.py:86-91
  if (tid == 0) {
    for (int i = 0; i < MP_NUM_JOINTS; ++i) {
      float T[16];
      int p = m.parents[i];
      for (int r = 0; r < 3; ++r) {
        for (int c2 = 0; c2 < 3; ++c2) T[4 * r + c2] = sR[i][3 * r + c2];
        T[4 * r + 3] = (i == 0) ? sJ[0][r] : (sJ[i][r] - sJ[p][r]);
      }
      T[12] = T[13] = T[14] = 0.f;
      T[15] = 1.f;
      if (i == 0) {
        for (int k = 0; k < 16; ++k) sG[0][k] = T[k];
      } else {
        mat4_mul(sG[p], T, sG[i]);
      }
    }
  }
  __syncthreads();
  if (tid < MP_NUM_JOINTS) {
    // rel_transforms = G - pad(G @ [J;0])      lbs.py:371-376
    float A[16];
    for (int k = 0; k < 16; ++k) A[k] = sG[tid][k];
    for (int r = 0; r < 4; ++r) {
      float t = sG[tid][4 * r] * sJ[tid][0] + sG[tid][4 * r + 1] * sJ[tid][1] + sG[tid][4 * r + 2] * sJ[tid][2];
      A[4 * r + 3] -= t;
    }
    const float sc = scale_p[0];
    // tf_mats[:, :, :3, :] *= scale ; tf_mats[:, :, :3, 3] += transl * scale      smpl.py:86-88
    for (int r = 0; r < 3; ++r) {
      for (int c2 = 0; c2 < 4; ++c2) A[4 * r + c2] *= sc;
      A[4 * r + 3] += transl[r] * sc;
    }
    for (int k = 0; k < 16; ++k) m.A_abs[tid * 16 + k] = A[k];
    float O[16];
    if (absolute) {
      for (int k = 0; k < 16; ++k) O[k] = A[k];
    } else {
      mat4_mul(A, m.tfs_c_inv + tid * 16, O);      // einsum('bnij,njk->bnik')  smpl.py:91
    }
    for (int k = 0; k < 16; ++k) tfs_out[tid * 16 + k] = O[k];
  }
}

// per vertex: pose blend shapes + skinning + SMPLServer's scale/translation
//   v_posed = v_shaped + pose_feature @ posedirs ; T = W @ A ; verts = T v_posed      lbs.py:201-227, smpl.py:78
// A_abs already carries the scale and translation: (s*A_rot) v + (s*A_t + t*s) = s*(A v) + t*s.
__global__ void smpl_skin_kernel(Smpl m, const float* __restrict__ betas, float* __restrict__ verts_out) {
  __shared__ float spf[207];
  __shared__ float sA[MP_NUM_JOINTS * 16];
  __shared__ float sb[10];
  if (threadIdx.x < 10) sb[threadIdx.x] = betas[threadIdx.x];
  for (int i = threadIdx.x; i < 207; i += blockDim.x) spf[i] = m.pose_feature[i];
  for (int i = threadIdx.x; i < MP_NUM_JOINTS * 16; i += blockDim.x) sA[i] = m.A_abs[i];
  __syncthreads();
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= m.V) return;
  const size_t ld = (size_t)m.V * 3;
  float p0 = 0.f, p1 = 0.f, p2 = 0.f;
  for (int k = 0; k < 207; ++k) {
    const float* pd = m.posedirs + (size_t)k * ld + 3 * (size_t)v;
    float f = spf[k];
    p0 = fmaf(f, pd[0], p0);
    p1 = fmaf(f, pd[1], p1);
    p2 = fmaf(f, pd[2], p2);
  }
  // v_shaped = v_template + blend_shapes(betas, shapedirs)      lbs.py:184, :252-273
  float vs[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float* sd = m.shapedirs + ((size_t)3 * v + a) * 10;
    float sacc = 0.f;
#pragma unroll
    for (int l = 0; l < 10; ++l) sacc = fmaf(sb[l], sd[l], sacc);
    vs[a] = m.v_template[3 * v + a] + sacc;
  }
  float x = vs[0] + p0, y = vs[1] + p1, z = vs[2] + p2;
  float T[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) T[k] = 0.f;
  const float* w = m.lbs_weights + (size_t)v * MP_NUM_JOINTS;
  for (int j = 0; j < MP_NUM_JOINTS; ++j) {
    float wj = w[j];
    if (wj == 0.f) continue;
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = fmaf(wj, sA[16 * j + k], T[k]);
  }
  verts_out[3 * v] = T[0] * x + T[1] * y + T[2] * z + T[3];
  verts_out[3 * v + 1] = T[4] * x + T[5] * y + T[6] * z + T[7];
  verts_out[3 * v + 2] = T[8] * x + T[9] * y + T[10] * z + T[11];
}

// tfs_c_inv = inverse of 24 affine 4x4 matrices (bottom row 0 0 0 1)      smpl.py:47
__global__ void affine_inverse_kernel(const float* __restrict__ T, float* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* A = T + 16 * i;
  float a = A[0], b = A[1], c = A[2], d = A[4], e = A[5], f = A[6], g = A[8], h = A[9], k = A[10];
  float c00 = e * k - f * h, c01 = -(d * k - f * g), c02 = d * h - e * g;
  float det = a * c00 + b * c01 + c * c02;
  float r = 1.0f / det;
  float I[9] = {c00 * r, -(b * k - c * h) * r, (b * f - c * e) * r, c01 * r, (a * k - c * g) * r, -(a * f - c * d) * r,
                c02 * r, -(a * h - b * g) * r, (a * e - b * d) * r};
  float* o = out + 16 * i;
  for (int rr = 0; rr < 3; ++rr) {
    for (int cc = 0; cc < 3; ++cc) o[4 * rr + cc] = I[3 * rr + cc];
    o[4 * rr + 3] = -(I[3 * rr] * A[3] + I[3 * rr + 1] * A[7] + I[3 * rr + 2] * A[11]);
  }
  o[12] = o[13] = o[14] = 0.f;
  o[15] = 1.f;
}

static int smpl_run(const Smpl& m, const float* scale, const float* transl, const float* thetas, const float* betas,
                    int absolute, float* verts, float* tfs, cudaStream_t st) {
  smpl_pose_kernel<<<1, 1024, 0, st>>>(m, scale, transl, thetas, betas, absolute, tfs);
  MP_LAUNCH_CHECK();
  smpl_skin_kernel<<<div_up(m.V, 128), 128, 0, st>>>(m, betas, verts);
  MP_LAUNCH_CHECK();
  return 0;
}

}  // namespace mp

struct mp_smpl {
  mp::Smpl m;
};

extern "C" {

size_t mp_smpl_bytes(int V) {
  return mp::align_up((size_t)V * 3 * 4, 256) * 2 + 256 * 8 + mp::align_up(207 * 4, 256) + 3 * mp::align_up(24 * 16 * 4, 256) +
         mp::align_up(86 * 4, 256) + mp::align_up(24 * 3 * 4, 256) + mp::align_up(24 * 3 * 10 * 4, 256) + 4096;
}

int mp_smpl_create(const float* v_template, const float* shapedirs, const float* posedirs, const float* J_regressor,
                   const int* parents_host, const float* lbs_weights, int V, const float* betas_canonical, void* storage,
                   size_t storage_bytes, mp_smpl_t** out, void* stream) {
  using namespace mp;
  MP_REQUIRE(v_template && shapedirs && posedirs && J_regressor && parents_host && lbs_weights && storage && out,
             "mp_smpl_create: null argument");
  MP_REQUIRE(storage_bytes >= mp_smpl_bytes(V), "mp_smpl_create: storage too small");
  cudaStream_t st = (cudaStream_t)stream;
  mp_smpl* h = new mp_smpl();
  Smpl& m = h->m;
  m.V = V;
  m.v_template = v_template;
  m.shapedirs = shapedirs;
  m.posedirs = posedirs;
  m.J_regressor = J_regressor;
  m.lbs_weights = lbs_weights;
  for (int i = 0; i < MP_NUM_JOINTS; ++i) m.parents[i] = parents_host[i];
  Arena a(storage, storage_bytes);
  m.v_shaped = a.take<float>((size_t)V * 3);
  m.verts_c = a.take<float>((size_t)V * 3);
  m.pose_feature = a.take<float>(207);
  m.A_abs = a.take<float>(24 * 16);
  m.tfs_c_inv = a.take<float>(24 * 16);
  m.tmp_tfs = a.take<float>(24 * 16);
  m.J_t = a.take<float>(24 * 3);
  m.J_s = a.take<float>(24 * 3 * 10);
  float* canon = a.take<float>(86);
  if (!a.ok) {
    delete h;
    set_error("mp_smpl_create: arena overflow");
    return -1;
  }
  // canonical pose (smpl.py:35-47): scale 1, no translation, hips +-pi/6 about z, the person's betas; absolute
  // transforms, inverted once
  float host[86];
  memset(host, 0, sizeof(host));
  host[0] = 1.f;
  host[4 + 5] = (float)(M_PI / 6.0);
  host[4 + 8] = (float)(-M_PI / 6.0);
  cudaError_t e = cudaMemcpyAsync(canon, host, sizeof(host), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // host[] is a stack buffer
  if (e != cudaSuccess) {
    delete h;
    set_error("mp_smpl_create: %s", cudaGetErrorString(e));
    return -2;
  }
  const float* betas = betas_canonical ? betas_canonical : canon + 76;
  smpl_jreg_kernel<<<1, 1024, 0, st>>>(m);
  g_launches++;
  int rc = smpl_run(m, canon, canon + 1, canon + 4, betas, 1, m.verts_c, m.tmp_tfs, st);
  if (rc == 0) {
    affine_inverse_kernel<<<1, 32, 0, st>>>(m.tmp_tfs, m.tfs_c_inv, 24);
    g_launches++;
  }
  if (rc) {
    delete h;
    return rc;
  }
  *out = h;
  return 0;
}

void mp_smpl_free(mp_smpl_t* h) { delete h; }

int mp_smpl_canonical(mp_smpl_t* h, float* verts_c, float* tfs_c_inv, void* stream) {
  MP_REQUIRE(h, "mp_smpl_canonical: null handle");
  cudaStream_t st = (cudaStream_t)stream;
  if (verts_c)
    MP_CHECK_CUDA(cudaMemcpyAsync(verts_c, h->m.verts_c, (size_t)h->m.V * 3 * 4, cudaMemcpyDeviceToDevice, st));
  if (tfs_c_inv)
    MP_CHECK_CUDA(cudaMemcpyAsync(tfs_c_inv, h->m.tfs_c_inv, 24 * 16 * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int mp_smpl_forward(mp_smpl_t* h, const float* scale, const float* transl, const float* thetas, const float* betas,
                    int absolute, float* smpl_verts, float* smpl_tfs, void* stream) {
  MP_REQUIRE(h && scale && transl && thetas && betas && smpl_verts && smpl_tfs, "mp_smpl_forward: null argument");
  return mp::smpl_run(h->m, scale, transl, thetas, betas, absolute, smpl_verts, smpl_tfs, (cudaStream_t)stream);
}
}

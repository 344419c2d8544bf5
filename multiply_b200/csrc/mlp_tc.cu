// tcgen05 engine: the fused ImplicitNet (+ backward for normals) + RenderingNet chain on the
// 5th-generation tensor cores of sm_100a.
//   reference: /root/reference/code/lib/model/networks.py:126-208 (ImplicitNet.forward),
//              :263-312 (RenderingNet.forward), lib/model/multiply.py:620-661 (forward_gradient)
//
// Design:
//   * one persistent CTA per SM; a tile is 128 sample points (MMA M = 128, cta_group::1).
//   * every layer is D[128 x 256] (fp32, 256 TMEM columns) = A[128 x K] . W^T, K in chunks of 64.
//     A (the activations) lives in shared memory as fp16 hi + fp16 lo (K-major, 128B swizzle);
//     the weights are streamed as pre-swizzled fp16 hi/lo tiles ("slots", 256 x 64, 32 KB) by
//     the TMA engine (cp.async.bulk + mbarrier complete_tx) through a 3-slot ring.
//   * split precision: D = A_hi.W_hi + A_lo.W_hi + A_hi.W_lo (three kind::f16 MMAs per K-step,
//     fp32 accumulate) — 22 significand bits per operand, which is what keeps RGB/SDF within the
//     1e-4 gate that a single bf16/fp16 pass misses by two orders of magnitude.
//   * warp roles: warp 0 = weight loader, warp 1 = MMA issuer (one elected lane),
//     warps 2..17 = epilogue (TMEM -> registers -> activation -> fp16 hi/lo -> shared memory):
//     TMEM lane quadrant x column part; a part owns 16 columns of each 64-wide K-block of the next
//     operand and hands them over K-block by K-block, so the next layer's MMAs start after a quarter
//     of the epilogue, into the other of two accumulators (512 TMEM columns in all).
//   * the whole per-sample chain runs inside the tile: embed, L0..L7 (+ sigma' stash), SDF dot,
//     the reverse sweep B7..B0 (d sdf / d x_c), normals, colour layers (the feature layer L8 folded
//     into colour layer 0, its extra inputs as a fifth K-block), RGB.
#include "common.cuh"
#include <vector>
#include <mutex>
#include <stdlib.h>

namespace mp {

// ---------------------------------------------------------------------------------------------
// program description
// ---------------------------------------------------------------------------------------------
enum { EPI_SOFTPLUS = 0, EPI_FEAT = 1, EPI_BWD = 2, EPI_RELU = 3 };
enum {
  F_SAVE_SIG = 1,      // store d softplus / dz to the sigma' scratch (forward, grad mode)
  F_INJECT_EMB = 2,    // columns >= inj_col receive the input embedding (skip connection, networks.py:166)
  F_SDF_DOT = 4,       // sdf = h7 . W8[0,:] + b8[0]
  F_SEED_BWD = 8,      // after this step: A = W8[0,:] * sigma'_7 (start of the reverse sweep)
  F_SKIP_GRAD = 16,    // reverse sweep: columns >= inj_col are d/d embed of the skip; park them, zero A there
  F_FINAL_GRAD = 32,   // reverse sweep end: d sdf / d x -> normal ; then reload the features into A
  F_EXTRA_IN = 64,     // colour layer 0: a fifth K-block carries the extra inputs ([x_c, n] or the view embedding): they are
                       // staged into A's K-block 0 once its MMAs have drained, and accumulate into the same tile
  F_RGB_OUT = 128,     // last colour layer: rgb = sigmoid(h . Wrgb^T + b)
  F_FEAT_OUT = 256,    // write the fp32 features to feat_out (operator API)
  F_STASH_FEAT = 512   // park the feature chunks in scratch (they return as the colour net's input after the reverse sweep)
};
constexpr int kMaxSteps = 24;
constexpr int kSlotBytes = 32768;          // 256 rows x 64 fp16
constexpr int kRing = 3;

struct TcStep {
  int nk;               // 64-wide K chunks of A consumed by this layer (5 with F_EXTRA_IN: the last one re-uses K-block 0)
  int epi;
  int flags;
  int sig;              // sigma' scratch layer (save: forward, load: reverse) or -1
  const float* bias;    // [256] or nullptr
  int ncols;            // output columns that carry data (the rest are padding)
  int slot_off;         // first weight slot of this step in the blob
  int sc;               // index of this layer's 2^-s in inv_scale[]
  int terms;            // split-precision terms of this step's products: 3 = A_hi.W_hi + A_lo.W_hi + A_hi.W_lo (default),
                        // 1 = A_hi.W_hi only (the lo weight slots are then neither loaded nor issued)
};

struct TcProgram {
  int nsteps;
  int slots_per_tile;
  TcStep step[kMaxSteps];
  const uint4* blob;     // weight slots in consumption order
  const float* inv_scale;   // [nsteps] 2^-s of each step's weights
  // network constants
  int d_in, multires, E, inj_col, n_extra, col_n;   // col_n: width of the colour hidden layers
  const float* w8row;    // W8[0,:]  [256]
  const float* b8;       // b8[0]
  const float* Wrgb;     // [3][256]
  const float* brgb;     // [3]
};

struct TcIO {
  const float* x;        // [cap, d_in] canonical points (compact list)
  const int* slot;       // [cap] output slot of each point or nullptr (identity)
  const int* count;      // device count or nullptr (= cap)
  int cap;
  const float* jinv;     // [cap, 12] (3x3 row-major, padded to three float4) or nullptr
  const float* extra;    // [cap, n_extra] extra colour inputs (background view embedding) or nullptr
  float* sdf_out;        // scattered by slot
  float* rgb_out;        // [slots,3]
  float* nrm_out;        // [slots,3]
  float* grad_out;       // [cap,3] dense or nullptr
  float* feat_out;       // [cap,256] dense or nullptr
  int knobs;             // diagnostics (MP_TC_KNOBS bit mask): bit 1 = record the cycle stamps of mp_tc_trace_read,
                         // bit 2 = keep the dead scratch lines (no discard.global.L2)
  float rz;              // relative truncation loss of ONE tensor-core accumulation (see kRzPerMma)
  char* scratch;         // per-CTA scratch
  size_t scratch_per_cta;
};

// per-CTA scratch layout (bytes)
constexpr size_t kSigBytes = (size_t)8 * 64 * 128 * 16;      // sigma' [8][64][128] float4
constexpr size_t kFeatBytes = (size_t)2 * 32 * 128 * 16;     // features hi/lo chunks
constexpr size_t kGeBytes = (size_t)96 * 128 * 4;            // skip gradient [E<=96][128]
constexpr size_t kMiscBytes = (size_t)128 * 32 * 4;          // partial sums / normals [128][32]
constexpr size_t kEmbBytes = (size_t)96 * 128 * 4;           // input embedding of the tile [E<=96][128]
constexpr size_t kScratchPerCta = kSigBytes + kFeatBytes + kGeBytes + kMiscBytes + kEmbBytes;

// shared memory carve-up
constexpr int kABytes = 2 * 4 * 128 * 128;                   // hi + lo, 4 K-blocks of [128 x 128B]
constexpr int kXchBytes = 2 * 128 * 4;                        // d sdf / d x_1, d x_2 of the tile's rows (final-gradient step)
constexpr int kSmemBytes = kABytes + kRing * kSlotBytes + 256 + kXchBytes + 1024;
static_assert(kSmemBytes <= 232448, "shared memory budget of one CTA (227 KB)");

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major, 1), [32,46) SBO>>4 = 1024B between
//   8-row groups, [46,48) version = 1 (sm_100), [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// kind::f16 instruction descriptor: D = F32, A = B = F16, both K-major, N = 256, M = 128
__device__ __forceinline__ uint32_t make_idesc() { return (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24); }

// ---------------------------------------------------------------------------------------------
// epilogue helpers
// ---------------------------------------------------------------------------------------------
// byte offset of (row, 8-column chunk `ch` of K-block `kb`) in the swizzled A image
__device__ __forceinline__ uint32_t a_off(int row, int kb, int ch) {
  return (uint32_t)(kb * 16384 + row * 128 + ((ch ^ (row & 7)) << 4));
}

__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    float2 f = __half22float2(h[i]);
    l[i] = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
  }
  hi = make_uint4(*(uint32_t*)&h[0], *(uint32_t*)&h[1], *(uint32_t*)&h[2], *(uint32_t*)&h[3]);
  lo = make_uint4(*(uint32_t*)&l[0], *(uint32_t*)&l[1], *(uint32_t*)&l[2], *(uint32_t*)&l[3]);
}

// 16-byte store into the operand image through a 32-bit shared-window address.  (A pointer derived from the aligned
// dynamic-shared base loses its address space: the compiler emitted generic ST.E.128 with 64-bit address arithmetic.)
// (volatile, no memory clobber: ordered against the volatile proxy fence / mbarrier arrive that publish the image; no
// C++ access reads these bytes back.)
__device__ __forceinline__ void sts128(uint32_t A32, uint32_t off, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(A32 + off), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ void store_a8(uint32_t A32, int row, int col, const float* v) {
  uint4 hi, lo;
  split8(v, hi, lo);
  uint32_t o = a_off(row, col >> 6, (col >> 3) & 7);
  sts128(A32, o, hi);
  sts128(A32, 65536 + o, lo);
}

// element k of the positional embedding of x (embedders.py:8-34)
__device__ __forceinline__ float embed_elem(const float* x, int d, int k) {
  if (k < d) return x[k];
  int f = (k - d) / (2 * d), r = (k - d) - f * 2 * d;
  float t = __fmul_rn(x[r < d ? r : r - d], (float)(1 << f));
  return r < d ? sinf(t) : cosf(t);
}

// The per-CTA scratch (sigma', stashed features) streams: it is written once and read once per tile.  L2-only
// accesses keep it out of the small L1 that is left beside 224 KB of shared memory, so that the read-only vectors
// every chunk needs (bias, W8 row, extra-input and rgb weights) stay L1-resident.
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// The scratch lines of a tile are dead once read (sigma' and the stashed features are written once and read once):
// `discard.global.L2` drops the line without write-back, so the dead lines neither travel to DRAM when they are
// evicted nor compete for L2 capacity with the live part of the stash.  `dep` is a value computed FROM the loaded
// data: the consumer instruction waits for the load of every lane of the warp, so the discard cannot overtake it.
__device__ __forceinline__ void discard_line(const void* p, float dep) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p), "f"(dep) : "memory");
}
__device__ __forceinline__ void discard_line(const void* p, uint32_t dep) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p), "r"(dep) : "memory");
}

template <int N>
__device__ __forceinline__ void ep_bar() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

// softplus(beta=100, threshold=20) and its derivative (networks.py:85)
// Branch-free so that the elements of a chunk pipeline through the MUFU unit (a per-element branch serialises
// them: measured 5x slower).  Raw MUFU approximations (ex2 / lg2 / rcp .approx.ftz) without the
// denormal fix-ups of __expf/__logf: 1+u >= 1, the overflow side is replaced by the linear branch, results only need
// ~1e-7 absolute accuracy (softplus = log1p(exp(100 z))/100).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float softplus_fast(float z) {
  float t = z * 144.26950408889634f;                 // 100 z log2(e)
  float u = ex2_approx(t);                           // inf above ~128: discarded by the select below
  float y = lg2_approx(1.f + u) * 0.0069314718055994531f;   // ln2 / 100
  return t > 28.853900817779268f ? z : y;
}
__device__ __forceinline__ void softplus_fast_grad(float z, float& y, float& d) {
  float t = z * 144.26950408889634f;
  float u = ex2_approx(t);                           // inf above ~128: u * rcp(inf) = NaN, discarded below
  float w = 1.f + u;
  float r = rcp_approx(w);
  float ys = lg2_approx(w) * 0.0069314718055994531f;
  bool big = t > 28.853900817779268f;
  y = big ? z : ys;
  d = big ? 1.f : u * r;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// tcgen05.ld is asynchronous: `tmem_issue` starts the load of CW columns of this thread's row, `tmem_wait`
// (tcgen05.wait::ld) makes the registers valid.  The wait lists the destination registers as in/out
// operands so the compiler cannot touch them between the two.
template <int CW>
__device__ __forceinline__ void tmem_issue(uint32_t taddr, float* v);
template <>
__device__ __forceinline__ void tmem_issue<32>(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_issue<16>(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
template <int CW>
__device__ __forceinline__ void tmem_wait(float* v);
template <>
__device__ __forceinline__ void tmem_wait<16>(float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_wait<32>(float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                 "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                 "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// Step kinds: the chunk loop of a layer step is compiled once per kind (compile-time tag), so that every kind gets its own
// instruction schedule and register allocation instead of one loop that tests the step descriptor in its body.  (The
// single generic loop of round 1 was at the mercy of the optimiser: unrelated edits elsewhere in the kernel moved the
// forward steps between 13.5 k and 18 k cycles.)
enum { K_SP_PLAIN = 0, K_SP_SAVE = 1, K_SP_SEED = 2, K_FEAT = 3, K_BWD = 4, K_RELU = 5 };
template <int K>
struct KTag {
  static constexpr int value = K;
};

// cycle stamps of CTA 0 (MP_TC_KNOBS bit 1): [0..] epilogue warp 2, [2048..] MMA issuer; see mp_tc_trace_read
#ifndef MP_TC_TRACE
#define MP_TC_TRACE 0
#endif
__device__ unsigned long long g_trace[4096];

// NW epilogue warps (8 or 16): warp w owns TMEM lane quadrant w % 4 (rows) and column part (w-2)/4.
//
// PIPE (NW = 16 only): K-block-granular hand-over between the epilogue and the MMA issuer.
//   * a column part no longer owns one 64-column K-block of the next layer's operand; it owns 16 columns of
//     each of the four, visited in K-block order, and arrives on a per-K-block barrier after each chunk.  The
//     next layer's MMAs over K-block kb therefore start after a quarter of the epilogue instead of after all
//     of it;
//   * those MMAs write a second accumulator (TMEM columns 256..511, alternating per step) because the
//     epilogue is still draining the first.
//   The operand A stays single-buffered: all MMAs of a step have completed before its epilogue starts.
template <int NW, bool PIPE>
__global__ void __launch_bounds__(64 + 32 * NW, 1) tc_chain_kernel(const __grid_constant__ TcProgram P,
                                                                   const __grid_constant__ TcIO io) {
  static_assert(!PIPE || NW == 16, "PIPE needs four column parts of 16-column chunks");
  constexpr int NBAR = PIPE ? 4 : 1;       // operand hand-over barriers (one per K-block when pipelined)
  constexpr int TCOLS = PIPE ? 512 : 256;  // TMEM columns
  constexpr int NPART = NW / 4;            // column parts
  constexpr int PCOLS = 256 / NPART;       // columns per part
  constexpr int CW = (NW == 16) ? 16 : 32; // columns per TMEM load (register budget)
  constexpr int G4 = CW / 4;
  constexpr int NEPI = 32 * NW;
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1024-byte aligned carve-up (SWIZZLE_128B atoms)
  char* base = (char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  char* A = base;                                   // [hi | lo] x 4 K-blocks
  const uint32_t A32 = smem_u32(A);
  char* ring = base + kABytes;                      // kRing weight slots
  uint64_t* bars = (uint64_t*)(ring + kRing * kSlotBytes);
  uint64_t* full = bars;                            // [kRing]
  uint64_t* empty = bars + kRing;                   // [kRing]
  uint64_t* d_full = bars + 2 * kRing;
  uint64_t* a_ready = bars + 2 * kRing + 1;
  uint64_t* x_free = bars + 2 * kRing + 1 + NBAR;      // K-block 0 of A drained by the MMAs (extra-input steps)
  uint64_t* x_ready = x_free + 1;                      // extra inputs staged there
  uint32_t* tmem_slot = (uint32_t*)(x_ready + 1);
  float* xs = (float*)((char*)bars + 256);             // [2][128] exchange of the final-gradient step

  const int count = io.count ? min(io.cap, *io.count) : io.cap;
  const int ntiles = (count + 127) >> 7;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(d_full, 1);
    for (int i = 0; i < NBAR; ++i) mbar_init(&a_ready[i], NEPI);
    mbar_init(x_free, 1);
    mbar_init(x_ready, NEPI);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem0 = *tmem_slot;

  if (warp == 0) {
    // ===================== weight loader =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int s = 0; s < P.nsteps; ++s) {
          const char* src = (const char*)P.blob + (size_t)P.step[s].slot_off * kSlotBytes;
          const int nslot = 2 * P.step[s].nk;
          const int jstep = P.step[s].terms == 1 ? 2 : 1;
          for (int j = 0; j < nslot; j += jstep, ++it) {
            int r = it % kRing;
            uint32_t ph = (it / kRing) & 1;
            mbar_wait(&empty[r], ph ^ 1);
            mbar_expect_tx(&full[r], kSlotBytes);
            bulk_g2s(ring + (size_t)r * kSlotBytes, src + (size_t)j * kSlotBytes, kSlotBytes, &full[r]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc();
      const uint32_t a_hi = smem_u32(A), a_lo = smem_u32(A) + 65536;
      uint32_t it = 0, ar_ph = 0, buf = 0, xr_ph = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int s = 0; s < P.nsteps; ++s) {
          if (!PIPE) {
            mbar_wait(a_ready, ar_ph);
            tc_fence_after();
          }
          const int nk = P.step[s].nk;
          const bool one_term = P.step[s].terms == 1;
          const uint32_t tmem = tmem0 + (PIPE ? buf * 256u : 0u);
          uint32_t acc = 0;
          for (int kc = 0; kc < nk; ++kc) {
            if (kc == 4) {
              // extra-input K-block: lives where K-block 0 was
              mbar_wait(x_ready, xr_ph);
              xr_ph ^= 1;
              tc_fence_after();
            } else if (PIPE) {
              mbar_wait(&a_ready[kc], ar_ph);
              tc_fence_after();
#if MP_TC_TRACE
              if ((io.knobs & 2) && blockIdx.x == 0 && tile == (int)(blockIdx.x + gridDim.x)) g_trace[2048 + (P.nsteps > 12 ? 0 : 1024) + s * 8 + kc] = clock64();
#endif
            }
            // hi slot: A_hi.W_hi + A_lo.W_hi
            int r = it % kRing;
            mbar_wait(&full[r], (it / kRing) & 1);
            tc_fence_after();
            uint32_t wb = smem_u32(ring + (size_t)r * kSlotBytes);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              uint64_t bd = make_desc(wb + ks * 32);
              umma_f16(tmem, make_desc(a_hi + (kc & 3) * 16384 + ks * 32), bd, idesc, acc);
              acc = 1;
              if (!one_term) umma_f16(tmem, make_desc(a_lo + (kc & 3) * 16384 + ks * 32), bd, idesc, 1);
            }
            umma_commit(&empty[r]);
            ++it;
            if (!one_term) {
            // lo slot: A_hi.W_lo
            r = it % kRing;
            mbar_wait(&full[r], (it / kRing) & 1);
            tc_fence_after();
            wb = smem_u32(ring + (size_t)r * kSlotBytes);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16(tmem, make_desc(a_hi + (kc & 3) * 16384 + ks * 32), make_desc(wb + ks * 32), idesc, 1);
            umma_commit(&empty[r]);
            ++it;
            }
            if (kc == 0 && nk == 5) umma_commit(x_free);   // K-block 0 may be overwritten once these have completed
          }
          umma_commit(d_full);
#if MP_TC_TRACE
          if ((io.knobs & 2) && blockIdx.x == 0 && tile == (int)(blockIdx.x + gridDim.x)) g_trace[2048 + (P.nsteps > 12 ? 0 : 1024) + s * 8 + 4] = clock64();
#endif
          if (PIPE) {
            // keep the phases of the unused K-block barriers in step
            for (int kc = nk < 4 ? nk : 4; kc < 4; ++kc) mbar_wait(&a_ready[kc], ar_ph);
            buf ^= 1;
          }
          ar_ph ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int part = (warp - 2) >> 2;       // column part
    const int row = q * 32 + lane;
    const uint32_t t_row0 = tmem0 + ((uint32_t)(q * 32) << 16);
    uint32_t ebuf = 0;                       // accumulator the next step drains (PIPE)
    char* scr = io.scratch + (size_t)blockIdx.x * io.scratch_per_cta;
    float4* sig = (float4*)scr;                                  // [8][64][128]
    uint4* fsc = (uint4*)(scr + kSigBytes);                      // [2][32][128]
    float* ge = (float*)(scr + kSigBytes + kFeatBytes);          // [96][128]
    float* misc = (float*)(scr + kSigBytes + kFeatBytes + kGeBytes);   // [128][32]
    float* emb = (float*)(scr + kSigBytes + kFeatBytes + kGeBytes + kMiscBytes);   // [96][128]
    uint32_t df_ph = 0, xf_ph = 0;
    const int d = P.d_in, E = P.E;
    const int cbeg = part * PCOLS;
    constexpr int NCH = PCOLS / CW;          // chunks per thread and step
    // first column of this thread's i-th chunk: contiguous, or 16 columns of every K-block in K order (PIPE)
    auto col_of = [&](int i) { return PIPE ? i * 64 + part * CW : cbeg + i * CW; };
#ifndef MP_AOFF
#define MP_AOFF 1
#endif
#if MP_AOFF
    static_assert(PIPE && CW == 16, "precomputed operand offsets assume 16-column chunks in K order");
    const uint32_t aoff0 = a_off(row, 0, (part * 2) & 7), aoff1 = a_off(row, 0, (part * 2 + 1) & 7);
#endif
    auto arrive_all = [&]() {
      fence_async_smem();
      tc_fence_before();
#pragma unroll
      for (int i = 0; i < NBAR; ++i) mbar_arrive(&a_ready[i]);
    };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int pt = tile * 128 + row;
      const bool valid = pt < count;
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      if (valid)
        for (int a = 0; a < d; ++a) x[a] = io.x[(size_t)pt * d + a];
      const int slot = valid ? (io.slot ? io.slot[pt] : pt) : 0;
      // ---- tile prologue: embedding -> A (K-blocks 0 .. nk0-1), zero padded ----
      {
        // positional embedding (embedders.py:8-34): the (frequency, axis) pairs of a row are split over its
        // column-part threads, one sincosf each, parked in scratch so that the skip connection of layer 4 and
        // the chain rule at the end of the reverse sweep re-read instead of recomputing them
        const int npair = d * P.multires;
        for (int pi = part; pi < npair; pi += NPART) {
          int f = pi / d, a = pi - f * d;
          float sn, cs;
          sincosf(__fmul_rn(x[a], (float)(1 << f)), &sn, &cs);
          emb[(size_t)(d + 2 * f * d + a) * 128 + row] = sn;
          emb[(size_t)(d + (2 * f + 1) * d + a) * 128 + row] = cs;
        }
        if (part == 0)
          for (int a = 0; a < d; ++a) emb[(size_t)a * 128 + row] = x[a];
        if (io.extra) {
          // background: embedding of the view direction (n_extra = 3 + 6 * frequencies values), same split
          float dv[3] = {0.f, 0.f, 0.f};
          if (valid)
            for (int a = 0; a < 3; ++a) dv[a] = io.extra[(size_t)pt * 3 + a];
          const int nvp = (P.n_extra - 3) / 2;
          for (int pi = part; pi < nvp; pi += NPART) {
            int f = pi / 3, a = pi - f * 3;
            float sn, cs;
            sincosf(__fmul_rn(dv[a], (float)(1 << f)), &sn, &cs);
            ge[(size_t)(3 + 6 * f + a) * 128 + row] = sn;
            ge[(size_t)(3 + 6 * f + 3 + a) * 128 + row] = cs;
          }
          if (part == 0)
            for (int a = 0; a < 3; ++a) ge[(size_t)a * 128 + row] = dv[a];
        }
        __threadfence_block();
        ep_bar<NEPI>();
        const int ncol = P.step[0].nk * 64;
        const int c0 = part * (ncol / NPART), c1 = c0 + ncol / NPART;
        for (int c = c0; c < c1; c += 8) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (c + j < E) ? emb[(size_t)(c + j) * 128 + row] : 0.f;
          store_a8(A32, row, c, v);
        }
        arrive_all();
      }
      // colour-net extra inputs: foreground [x_c, n] (networks.py:281) live in registers (n arrives at the end of the
      // reverse sweep); the background view-dir embedding (:275) was parked in `ge` by the prologue
      float nrm[3] = {0.f, 0.f, 0.f};
      for (int s = 0; s < P.nsteps; ++s) {
        const TcStep st = P.step[s];
        // 2^-s of the weight scaling, times the compensation of the accumulator's round-toward-zero (kRzPerMma)
        const float isc = P.inv_scale[st.sc] * fmaf(io.rz, (float)(4 * st.nk * (st.terms == 1 ? 1 : 3)), 1.f);
        // reverse-sweep steps: start fetching sigma' of the first chunk before blocking on the accumulator
        float4 s4[G4];
        const bool need_sig = (st.epi == EPI_BWD) && st.sig >= 0;
#pragma unroll
        for (int g4 = 0; g4 < G4; ++g4) s4[g4] = make_float4(0.f, 0.f, 0.f, 0.f);      // (always initialised: see b4 below)
        if (need_sig) {
#pragma unroll
          for (int g4 = 0; g4 < G4; ++g4) s4[g4] = ld_stream(&sig[((size_t)st.sig * 64 + ((col_of(0) >> 2) + g4)) * 128 + row]);
        }
        if (st.nk == 5) {
          // colour layer 0: once the MMAs over K-block 0 have drained it, this row's extra inputs (16 columns per
          // column part, zero padded to 64) take its place as the fifth K-block of the same accumulation
          mbar_wait(x_free, xf_ph);
          xf_ph ^= 1;
          constexpr int XC = 64 / NPART;       // columns of the extra K-block staged by each column part
#pragma unroll
          for (int c8 = 0; c8 < XC; c8 += 8) {
            const int e0 = part * XC + c8;
            float xv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) xv[j] = 0.f;
            if (io.extra) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (e0 + j < P.n_extra) xv[j] = ge[(size_t)(e0 + j) * 128 + row];
            } else if (e0 == 0) {
              xv[0] = x[0]; xv[1] = x[1]; xv[2] = x[2];
              xv[3] = nrm[0]; xv[4] = nrm[1]; xv[5] = nrm[2];
            }
            store_a8(A32, row, e0, xv);
          }
          fence_async_smem();
          mbar_arrive(x_ready);
        }
        // cycle stamps of one tile (scripts/gpu_trace.py): compiled in only with -DMP_TC_TRACE=1 (scripts/build_variant.sh);
        // even predicated off they cost ~1.5 % of the epilogue's issue slots
#if MP_TC_TRACE
        const bool tr = (io.knobs & 2) && blockIdx.x == 0 && warp == 2 && lane == 0 && tile == (int)(blockIdx.x + gridDim.x);
#else
        constexpr bool tr = false;
#endif
        unsigned long long* trp = g_trace + (P.nsteps > 12 ? 0 : 1024) + s * 8;
        if (tr) trp[0] = clock64();
        mbar_wait(d_full, df_ph);
        df_ph ^= 1;
        tc_fence_after();
        if (tr) trp[1] = clock64();
        auto reload_features = [&]() {
#pragma unroll
          for (int c8 = 0; c8 < PCOLS / 8; c8 += 4) {
            uint4 fh[4], fl[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int chunk = (cbeg >> 3) + c8 + u;
              fh[u] = ld_stream(&fsc[(size_t)chunk * 128 + row]);
              fl[u] = ld_stream(&fsc[(size_t)(32 + chunk) * 128 + row]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int chunk = (cbeg >> 3) + c8 + u;
              const uint32_t o = a_off(row, chunk >> 3, chunk & 7);
              sts128(A32, o, fh[u]);
              sts128(A32, 65536 + o, fl[u]);
              if ((lane & 7) == 0 && !(io.knobs & 4)) {
                discard_line(&fsc[(size_t)chunk * 128 + row], fh[u].x);
                discard_line(&fsc[(size_t)(32 + chunk) * 128 + row], fl[u].x);
              }
            }
          }
          arrive_all();
        };
        if (PIPE && (st.flags & F_FINAL_GRAD) && s + 1 < P.nsteps) {
          // all MMAs of the reverse sweep are done, A is free: the features return for the colour net right away and
          // its first layer's MMAs run under the rest of this step (the accumulators alternate)
          reload_features();
          if (tr) g_trace[512] = clock64();
        }
        const uint32_t t_row = t_row0 + (PIPE ? ebuf * 256u : 0u);
        ebuf ^= 1;
        // steps whose operand for the next step is complete chunk by chunk (no tail rewrites A)
        const bool chunk_handover = PIPE && s + 1 < P.nsteps && !(st.flags & (F_RGB_OUT | F_FINAL_GRAD));
        float dot0 = 0.f, dot1 = 0.f, dot2 = 0.f;        // sdf / rgb partial dots
        float va[CW];
        // the accumulator registers are dead once a chunk has been split to fp16: the next chunk's TMEM load is
        // started there, so its latency hides behind this chunk's stores and hand-over
        auto issue_next = [&](const int ci) {
          if (ci + 1 < NCH) tmem_issue<CW>(t_row + (uint32_t)col_of(ci + 1), va);
        };
        auto process_chunk = [&](auto ktag, float* v, const int ci) {
          constexpr int KIND = decltype(ktag)::value;
          const int c = col_of(ci);
          // the bias of this chunk is fetched under the TMEM load
          // (loaded unconditionally where the kind uses it and not declared live otherwise: a conditionally initialised
          // array makes the compiler keep it in local memory -- 8 local loads / stores per chunk, measured 13.5 k -> 20 k
          // cycles per forward step)
          float4 b4[G4];
          if constexpr (KIND != K_BWD) {
#pragma unroll
            for (int g4 = 0; g4 < G4; ++g4) b4[g4] = __ldg((const float4*)(st.bias + c + 4 * g4));
          }
          tmem_wait<CW>(v);
          if constexpr (KIND == K_SP_PLAIN || KIND == K_SP_SAVE || KIND == K_SP_SEED) {
            if constexpr (KIND == K_SP_SEED) {
              // last SDF layer of the fused chain: sigma'_7 is consumed right here -- the reverse sweep starts from
              // A = W8[0,:] * sigma'_7 (d sdf / d z7), h7 only feeds the sdf dot and the feature stash
              float seed[CW];
#pragma unroll
              for (int g4 = 0; g4 < G4; ++g4) {
                float4 w4 = __ldg((const float4*)(P.w8row + c + 4 * g4));
                float dd[4];
                softplus_fast_grad(fmaf(v[4 * g4 + 0], isc, b4[g4].x), v[4 * g4 + 0], dd[0]);
                softplus_fast_grad(fmaf(v[4 * g4 + 1], isc, b4[g4].y), v[4 * g4 + 1], dd[1]);
                softplus_fast_grad(fmaf(v[4 * g4 + 2], isc, b4[g4].z), v[4 * g4 + 2], dd[2]);
                softplus_fast_grad(fmaf(v[4 * g4 + 3], isc, b4[g4].w), v[4 * g4 + 3], dd[3]);
                seed[4 * g4 + 0] = dd[0] * w4.x;
                seed[4 * g4 + 1] = dd[1] * w4.y;
                seed[4 * g4 + 2] = dd[2] * w4.z;
                seed[4 * g4 + 3] = dd[3] * w4.w;
                dot0 = fmaf(v[4 * g4 + 0], w4.x, dot0);
                dot0 = fmaf(v[4 * g4 + 1], w4.y, dot0);
                dot0 = fmaf(v[4 * g4 + 2], w4.z, dot0);
                dot0 = fmaf(v[4 * g4 + 3], w4.w, dot0);
              }
              uint4 fh[CW / 8], fl[CW / 8];
#pragma unroll
              for (int j = 0; j < CW; j += 8) split8(v + j, fh[j >> 3], fl[j >> 3]);
              issue_next(ci);
#pragma unroll
              for (int j = 0; j < CW; j += 8) {
                if (st.flags & F_STASH_FEAT) {
                  int chunk = (c + j) >> 3;
                  st_stream(&fsc[(size_t)chunk * 128 + row], fh[j >> 3]);
                  st_stream(&fsc[(size_t)(32 + chunk) * 128 + row], fl[j >> 3]);
                }
                store_a8(A32, row, c + j, seed + j);
              }
              return;
            }
            if constexpr (KIND == K_SP_SAVE) {
#pragma unroll
              for (int g4 = 0; g4 < G4; ++g4) {
                float dd[4];
                softplus_fast_grad(fmaf(v[4 * g4 + 0], isc, b4[g4].x), v[4 * g4 + 0], dd[0]);
                softplus_fast_grad(fmaf(v[4 * g4 + 1], isc, b4[g4].y), v[4 * g4 + 1], dd[1]);
                softplus_fast_grad(fmaf(v[4 * g4 + 2], isc, b4[g4].z), v[4 * g4 + 2], dd[2]);
                softplus_fast_grad(fmaf(v[4 * g4 + 3], isc, b4[g4].w), v[4 * g4 + 3], dd[3]);
                st_stream(&sig[((size_t)st.sig * 64 + ((c >> 2) + g4)) * 128 + row], make_float4(dd[0], dd[1], dd[2], dd[3]));
              }
            } else {
#pragma unroll
              for (int g4 = 0; g4 < G4; ++g4) {
                v[4 * g4 + 0] = softplus_fast(fmaf(v[4 * g4 + 0], isc, b4[g4].x));
                v[4 * g4 + 1] = softplus_fast(fmaf(v[4 * g4 + 1], isc, b4[g4].y));
                v[4 * g4 + 2] = softplus_fast(fmaf(v[4 * g4 + 2], isc, b4[g4].z));
                v[4 * g4 + 3] = softplus_fast(fmaf(v[4 * g4 + 3], isc, b4[g4].w));
              }
            }
            if ((st.flags & F_INJECT_EMB) && c + CW > P.inj_col) {
#pragma unroll
              for (int j = 0; j < CW; ++j)
                if (c + j >= P.inj_col) v[j] = emb[(size_t)(c + j - P.inj_col) * 128 + row];
            }
            if (st.flags & F_SDF_DOT) {
#pragma unroll
              for (int g4 = 0; g4 < G4; ++g4) {
                float4 w4 = __ldg((const float4*)(P.w8row + c + 4 * g4));
                dot0 = fmaf(v[4 * g4 + 0], w4.x, dot0);
                dot0 = fmaf(v[4 * g4 + 1], w4.y, dot0);
                dot0 = fmaf(v[4 * g4 + 2], w4.z, dot0);
                dot0 = fmaf(v[4 * g4 + 3], w4.w, dot0);
              }
            }
          } else if constexpr (KIND == K_FEAT) {
#pragma unroll
            for (int g4 = 0; g4 < G4; ++g4) {
              const float4 b = b4[g4];
              v[4 * g4 + 0] = fmaf(v[4 * g4 + 0], isc, b.x);
              v[4 * g4 + 1] = fmaf(v[4 * g4 + 1], isc, b.y);
              v[4 * g4 + 2] = fmaf(v[4 * g4 + 2], isc, b.z);
              v[4 * g4 + 3] = fmaf(v[4 * g4 + 3], isc, b.w);
            }
            if ((st.flags & F_FEAT_OUT) && io.feat_out && valid) {
#pragma unroll
              for (int j = 0; j < CW; j += 4)
                *(float4*)(io.feat_out + (size_t)pt * 256 + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else if constexpr (KIND == K_BWD) {
            if ((st.flags & F_SKIP_GRAD) && c + CW > P.inj_col) {
              // columns >= inj_col are d/d embed through the skip connection: park them, zero them in A
              const float* s4f = reinterpret_cast<const float*>(s4);
#pragma unroll
              for (int j = 0; j < CW; ++j) {
                float gval = v[j] * isc;
                if (c + j >= P.inj_col) {
                  ge[(size_t)(c + j - P.inj_col) * 128 + row] = gval;
                  v[j] = 0.f;
                } else {
                  v[j] = gval * s4f[j];
                }
              }
            } else {
#pragma unroll
              for (int g4 = 0; g4 < G4; ++g4) {
                v[4 * g4 + 0] *= isc * s4[g4].x;
                v[4 * g4 + 1] *= isc * s4[g4].y;
                v[4 * g4 + 2] *= isc * s4[g4].z;
                v[4 * g4 + 3] *= isc * s4[g4].w;
              }
            }
            if (need_sig && (lane & 7) == 0 && !(io.knobs & 4)) {   // this chunk's sigma' lines are dead: drop them from L2
#pragma unroll
              for (int g4 = 0; g4 < G4; ++g4)
                discard_line(&sig[((size_t)st.sig * 64 + ((c >> 2) + g4)) * 128 + row], v[4 * g4]);
            }
            if (need_sig && ci + 1 < NCH) {   // next chunk's sigma' streams in behind the stores below
#pragma unroll
              for (int g4 = 0; g4 < G4; ++g4)
                s4[g4] = ld_stream(&sig[((size_t)st.sig * 64 + ((col_of(ci + 1) >> 2) + g4)) * 128 + row]);
            }
          } else {   // EPI_RELU
#pragma unroll
            for (int g4 = 0; g4 < G4; ++g4) {
              const float4 b = b4[g4];
              v[4 * g4 + 0] = fmaf(v[4 * g4 + 0], isc, b.x);
              v[4 * g4 + 1] = fmaf(v[4 * g4 + 1], isc, b.y);
              v[4 * g4 + 2] = fmaf(v[4 * g4 + 2], isc, b.z);
              v[4 * g4 + 3] = fmaf(v[4 * g4 + 3], isc, b.w);
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = fmaxf(v[j], 0.f);
            if (st.flags & F_RGB_OUT) {
#pragma unroll
              for (int g4 = 0; g4 < G4; ++g4) {
                float4 w0 = __ldg((const float4*)(P.Wrgb + c + 4 * g4));
                float4 w1 = __ldg((const float4*)(P.Wrgb + 256 + c + 4 * g4));
                float4 w2 = __ldg((const float4*)(P.Wrgb + 512 + c + 4 * g4));
                dot0 = fmaf(v[4 * g4 + 0], w0.x, fmaf(v[4 * g4 + 1], w0.y, fmaf(v[4 * g4 + 2], w0.z, fmaf(v[4 * g4 + 3], w0.w, dot0))));
                dot1 = fmaf(v[4 * g4 + 0], w1.x, fmaf(v[4 * g4 + 1], w1.y, fmaf(v[4 * g4 + 2], w1.z, fmaf(v[4 * g4 + 3], w1.w, dot1))));
                dot2 = fmaf(v[4 * g4 + 0], w2.x, fmaf(v[4 * g4 + 1], w2.y, fmaf(v[4 * g4 + 2], w2.z, fmaf(v[4 * g4 + 3], w2.w, dot2))));
              }
            }
          }
          // activations of this chunk -> A (fp16 hi/lo, swizzled) unless this is the last layer
          // (F_FINAL_GRAD steps never get here; F_RGB_OUT exists only on ReLU steps; F_STASH_FEAT only on the seed step,
          // which returned above -- testing them here cost seven predicated-off instructions per chunk)
          bool to_a = true;
          if constexpr (KIND == K_RELU) to_a = !(st.flags & F_RGB_OUT);
          if (to_a) {
            uint4 hi[CW / 8], lo[CW / 8];
#pragma unroll
            for (int j = 0; j < CW; j += 8) split8(v + j, hi[j >> 3], lo[j >> 3]);
            issue_next(ci);
#pragma unroll
            for (int j = 0; j < CW; j += 8) {
#if MP_AOFF
              // (PIPE) chunk ci is K-block ci; the 16-byte slot inside the row depends only on the column part
              const uint32_t o = (uint32_t)ci * 16384u + (j ? aoff1 : aoff0);
#else
              const uint32_t o = a_off(row, (c + j) >> 6, ((c + j) >> 3) & 7);
#endif
              sts128(A32, o, hi[j >> 3]);
              sts128(A32, 65536 + o, lo[j >> 3]);
            }
          } else {
            issue_next(ci);
          }
        };
        if (st.flags & F_FINAL_GRAD) {
          // ---- last step of the reverse sweep: its own straight-line code (everything it needs lives only here, so the
          // hot chunk loops of the other steps do not carry its registers) ----
          // B0's output columns are permuted at pack time BY AXIS: column part a (a < d_in) holds, in its 16 columns of
          // K-block 0, every embedding index that depends on x_a -- [x_a, sin(2^0 x_a), cos(2^0 x_a), sin(2^1 x_a), ...]
          // (embedders.py:8-34) -- so the chain rule
          //   d sdf / d x_a = sum_k (g_k + skip_k) * d embed_k / d x_a
          // of one axis runs inside one thread's registers.  The skip gradient (parked at the F_SKIP_GRAD step) and the
          // partner sin / cos of each index (parked by the tile prologue) come from scratch: independent loads issued
          // together under the TMEM load.
          float gax = 0.f;                                 // d sdf / d x_part (column parts 0..d-1)
          if (part < d) {
            tmem_issue<CW>(t_row + (uint32_t)col_of(0), va);
            float pg[14], pe[14];
#pragma unroll
            for (int j = 0; j < 14; ++j) {
              // j = 0: x_a itself; j = 1 + 2 f: sin(2^f x_a); j = 2 + 2 f: cos(2^f x_a)
              const int fq = (j - 1) >> 1;
              const int k = (j == 0) ? part : d + 2 * fq * d + ((j - 1) & 1) * d + part;
              const bool ok = j < 1 + 2 * P.multires;
              pg[j] = ok ? ge[(size_t)k * 128 + row] : 0.f;
              // partner: cos for a sin entry (+d), sin for a cos entry (-d)
              pe[j] = (ok && j > 0) ? emb[(size_t)(((j - 1) & 1) ? k - d : k + d) * 128 + row] : 0.f;
            }
            tmem_wait<CW>(va);
#pragma unroll
            for (int j = 0; j < 14; ++j) {
              const float tot = fmaf(va[j], isc, pg[j]);           // through layer 0 + through the skip connection
              const int fq = (j - 1) >> 1;
              // d x / d x = 1 ; d sin(2^f x) = 2^f cos(2^f x) ; d cos(2^f x) = -2^f sin(2^f x)
              const float w = (j == 0) ? 1.f : (float)(1 << fq) * (((j - 1) & 1) ? -pe[j] : pe[j]);
              gax = fmaf(w, tot, gax);
            }
          }
          tc_fence_before();
          if (tr) g_trace[513] = clock64();
          // axes 1..d-1 travel to column part 0 through shared memory (one float per row); it derives
          // normal = normalize(g . J^-1) (multiply.py:661), normalised again with eps 1e-6 (:606), writes the outputs and
          // is the part that stages the colour net's extra inputs [x_c, n].  Parts 1.. go on without waiting.
          // (Round 1 exchanged every term through global scratch behind two 512-thread barriers: 30 k cycles per tile.)
          float n0 = 0.f, n1 = 0.f, n2 = 0.f;
          if (part != 0) {
            if (part < d) {
              xs[(part - 1) * 128 + row] = gax;
              asm volatile("bar.arrive 4, %0;" ::"r"(128 * d) : "memory");
            }
          } else {
            // the inverse Jacobian of this row's point (cold: written by the deformer kernel) travels under the barrier
            float4 ja = make_float4(0.f, 0.f, 0.f, 0.f), jb = ja, jc = ja;
            if (io.jinv && valid) {
              const float4* J4 = (const float4*)(io.jinv + 12 * (size_t)pt);
              ja = __ldg(J4);      // J[0..3]
              jb = __ldg(J4 + 1);  // J[4..7]
              jc = __ldg(J4 + 2);  // J[8..11]
            }
            asm volatile("bar.sync 4, %0;" ::"r"(128 * d) : "memory");
            const float gx0 = gax, gx1 = d > 1 ? xs[row] : 0.f, gx2 = d > 2 ? xs[128 + row] : 0.f;
            if (io.grad_out && valid) {
              io.grad_out[3 * (size_t)pt] = gx0;
              io.grad_out[3 * (size_t)pt + 1] = gx1;
              io.grad_out[3 * (size_t)pt + 2] = gx2;
            }
            if (io.jinv && valid) {
              float v0 = gx0 * ja.x + gx1 * ja.w + gx2 * jb.z;
              float v1 = gx0 * ja.y + gx1 * jb.x + gx2 * jb.w;
              float v2 = gx0 * ja.z + gx1 * jb.y + gx2 * jc.x;
              float nr = fmaxf(sqrtf(v0 * v0 + v1 * v1 + v2 * v2), 1e-12f);     // multiply.py:661
              const float inr = 1.f / nr;
              v0 *= inr; v1 *= inr; v2 *= inr;
              float n2r = fmaxf(sqrtf(v0 * v0 + v1 * v1 + v2 * v2), 1e-6f);     // multiply.py:606
              const float in2 = 1.f / n2r;
              n0 = v0 * in2; n1 = v1 * in2; n2 = v2 * in2;
              if (io.nrm_out) {
                io.nrm_out[3 * (size_t)slot] = n0;
                io.nrm_out[3 * (size_t)slot + 1] = n1;
                io.nrm_out[3 * (size_t)slot + 2] = n2;
              }
            }
          }
          nrm[0] = n0;
          nrm[1] = n1;
          nrm[2] = n2;
          if (tr) g_trace[517] = clock64();
          // hand-over: the features already went back into A (reload_features above); nothing else to arrive on
          if (tr) trp[6] = clock64();
          continue;
        }
        // (a second register buffer for the TMEM reads was measured slower: +20 % kernel time from spills / code size)
        auto run_chunks = [&](auto ktag) {
          tmem_issue<CW>(t_row + (uint32_t)col_of(0), va);
#pragma unroll 1
          for (int ci = 0; ci < NCH; ++ci) {
            process_chunk(ktag, va, ci);
            if (tr) trp[2 + ci] = clock64();
            if (chunk_handover) {
              // K-block ci of the next layer's operand is complete in this thread
              fence_async_smem();
              tc_fence_before();
              mbar_arrive(&a_ready[ci]);
            }
          }
        };
        if (st.epi == EPI_SOFTPLUS) {
          if ((st.flags & (F_SAVE_SIG | F_SEED_BWD)) == (F_SAVE_SIG | F_SEED_BWD))
            run_chunks(KTag<K_SP_SEED>{});
          else if (st.flags & F_SAVE_SIG)
            run_chunks(KTag<K_SP_SAVE>{});
          else
            run_chunks(KTag<K_SP_PLAIN>{});
        } else if (st.epi == EPI_FEAT) {
          run_chunks(KTag<K_FEAT>{});
        } else if (st.epi == EPI_BWD) {
          run_chunks(KTag<K_BWD>{});
        } else {
          run_chunks(KTag<K_RELU>{});
        }
        tc_fence_before();
        // ---- step-specific tails ----
        if (st.flags & F_SDF_DOT) {
          misc[row * 32 + part] = dot0;
          __threadfence_block();
          ep_bar<NEPI>();
          if (part == 0 && valid && io.sdf_out) {
            float sacc = __ldg(P.b8);
#pragma unroll
            for (int pp = 0; pp < NPART; ++pp) sacc += misc[row * 32 + pp];
            io.sdf_out[slot] = sacc;
          }
          // (no second barrier: these scratch slots are next written a whole tile -- many barriers -- later)
        }
        if (st.flags & F_SKIP_GRAD) __threadfence_block();
        if (st.flags & F_RGB_OUT) {
          // last step of the tile: the MMAs are done with A, its K-block 3 serves as the exchange buffer for the
          // partial dots (the next write there is a whole layer step -- and several barriers -- away)
          float* xch = reinterpret_cast<float*>(A + 3 * 16384);
          xch[(part * 3 + 0) * 128 + row] = dot0;
          xch[(part * 3 + 1) * 128 + row] = dot1;
          xch[(part * 3 + 2) * 128 + row] = dot2;
          ep_bar<NEPI>();
          if (part == 0 && valid && io.rgb_out) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              float z = __ldg(P.brgb + k);
#pragma unroll
              for (int pp = 0; pp < NPART; ++pp) z += xch[(pp * 3 + k) * 128 + row];
              io.rgb_out[3 * (size_t)slot + k] = 1.f / (1.f + __expf(-z));
            }
          }
        }
        // hand A (and the drained accumulator) to the MMA warp for the next step of this tile;
        // the last step's hand-over is the next tile's prologue arrival
        if (s + 1 < P.nsteps && !chunk_handover) {
          if (!(st.flags & F_FINAL_GRAD))
            arrive_all();
          else if (!PIPE)
            reload_features();     // single accumulator: only after this step has drained it
        }
        if (tr) trp[6] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "n"(TCOLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// packing
// ---------------------------------------------------------------------------------------------
struct TcBlob {
  TcProgram sdf_prog;     // L0..L7 + sdf dot
  TcProgram full_prog;    // forward + reverse + colour
  TcProgram fwd_prog;     // L0..L8 (sdf + features), operator API
  bool has_full;
};

__global__ void absmax_kernel(const float* __restrict__ W, int n, float* __restrict__ out) {
  __shared__ float s[256];
  float m = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(W[i]));
  s[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] = fmaxf(s[threadIdx.x], s[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // scale = 2^s with max|W| * 2^s in (2^13, 2^14]  (fp16 hi/lo both stay in the normal range)
    float mx = s[0];
    int ex = 0;
    if (mx > 0.f) frexpf(mx, &ex);       // mx = f * 2^ex, f in [0.5,1)
    float sc = ldexpf(1.f, 14 - ex);
    out[0] = sc;
    out[1] = 1.f / sc;
  }
}

// One weight slot: B[n][k] for n < 256, k < 64 at K offset kc*64; value from W (natural [out][in], ld):
//   transposed == 0 : B[n][k] = W[(n_off + n) * ld + k_off + kc*64 + k]   (n < n_valid, kk < k_valid)
//   transposed == 1 : B[n][k] = W[(k_off + kc*64 + k) * ld + n_off + n]
__global__ void pack_slot_kernel(const float* __restrict__ W, int ld, int transposed, int n_off, int k_off,
                                 int n_valid, int k_valid, int kc, const float* __restrict__ scale,
                                 uint8_t* __restrict__ dst_hi, uint8_t* __restrict__ dst_lo, int perm16) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 256 * 64) return;
  int n = idx >> 6, k = idx & 63;
  int kk = kc * 64 + k;
  // perm16 = d_in > 0 (final step of the reverse sweep): output column a * 16 + j (a < d_in) carries the embedding index
  // of axis a in the order [x_a, sin(2^0 x_a), cos(2^0 x_a), sin(2^1 x_a), ...] (embedders.py:8-34), i.e.
  // j = 0 -> a ; j = 1 + 2 f + t -> d + 2 f d + t d + a ; every other column is padding.  Column part a of the epilogue
  // owns columns [16 a, 16 a + 16) of K-block 0, so it holds every term of d sdf / d x_a.
  int ns = n;
  if (perm16) {
    const int dd = perm16, a = n >> 4, j = n & 15;
    ns = n_valid;
    if (n < 64 && a < dd) {
      const int k = (j == 0) ? a : dd + ((j - 1) >> 1) * 2 * dd + ((j - 1) & 1) * dd + a;
      if (k < n_valid) ns = k;
    }
  }
  float w = 0.f;
  if (ns < n_valid && kk < k_valid)
    w = transposed ? W[(size_t)(k_off + kk) * ld + n_off + ns] : W[(size_t)(n_off + ns) * ld + k_off + kk];
  w *= scale[0];
  __half h = __float2half_rn(w);
  __half l = __float2half_rn(w - __half2float(h));
  uint32_t off = (uint32_t)(n * 128 + ((((k >> 3) ^ (n & 7))) << 4) + (k & 7) * 2);
  *(__half*)(dst_hi + off) = h;
  *(__half*)(dst_lo + off) = l;
}

__global__ void copy_strided_kernel(const float* __restrict__ src, int stride, int n, float* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[(size_t)i * stride];
}

// dst[r][c] (ld 256, zero padded) = src[r * lds + c] for c < ncols
__global__ void pad_rows_kernel(const float* __restrict__ src, int lds, int nrows, int ncols, float* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows * 256) return;
  int r = i >> 8, c = i & 255;
  dst[i] = (c < ncols) ? src[(size_t)r * lds + c] : 0.f;
}

// M[n][k] = sum_j Wc[n][coff + j] * W8f[j][k]   (n < n_rows, k < 256; W8f = W8[1:], row stride ld8; M row stride ldm)
// cb[n]   = sum_j Wc[n][coff + j] * b8f[j]
__global__ void fold_mm_kernel(const float* __restrict__ Wc, int ldc, int coff, const float* __restrict__ W8f, int ld8,
                               const float* __restrict__ b8f, int n_rows, float* __restrict__ M, int ldm,
                               float* __restrict__ cb) {
  int k = blockIdx.x * 16 + threadIdx.x, n = blockIdx.y * 16 + threadIdx.y;
  if (n >= n_rows || k >= 256) return;
  const float* w = Wc + (size_t)n * ldc + coff;
  double acc = 0.0;
  for (int j = 0; j < 256; ++j) acc += (double)w[j] * (double)W8f[(size_t)j * ld8 + k];
  M[(size_t)n * ldm + k] = (float)acc;
  if (k == 0) {
    double a2 = 0.0;
    for (int j = 0; j < 256; ++j) a2 += (double)w[j] * (double)b8f[j];
    cb[n] = (float)a2;
  }
}

// columns 256 .. 319 of the combined colour layer 0: M[n][256 + e] = Wc0[n][e] for the extra inputs e < n_extra
// (Wt = Wc0 transposed, [in][n_out]), zero elsewhere
__global__ void fill_extra_cols_kernel(const float* __restrict__ Wt, int n_out, int n_extra, float* __restrict__ M, int ldm) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 256 * 64) return;
  int n = i >> 6, e = i & 63;
  M[(size_t)n * ldm + 256 + e] = (n < n_out && e < n_extra) ? Wt[(size_t)e * n_out + n] : 0.f;
}

size_t tc_pack_bytes() {
  // sdf (58) + fwd (66) + full (162) slots ... share: full program contains fwd which contains sdf
  return (size_t)170 * kSlotBytes + (1 << 20);
}

struct PackCtx {
  Arena* a;
  cudaStream_t st;
  uint8_t* blob;
  int nslots;
  int nlayers;        // packed layers so far (index into scales / inv_scale)
  float* scales;      // [kMaxLayers][2] (scale, inv)
  float* inv_scale;   // [kMaxSteps]
  int rc;
};

static int pack_layer(PackCtx& c, TcStep& stp, const float* W, int ld, int transposed, int n_off, int k_off,
                      int n_valid, int k_valid, int nk, int total_elems, int perm16 = 0) {
  const int first = c.nslots;
  const int step = c.nlayers++;
  stp.slot_off = first;
  stp.sc = step;
  stp.terms = 3;
  if (c.rc) return first;
  absmax_kernel<<<1, 256, 0, c.st>>>(W, total_elems, c.scales + 2 * step);
  g_launches++;
  for (int kc = 0; kc < nk; ++kc) {
    uint8_t* hi = c.blob + (size_t)c.nslots * kSlotBytes;
    uint8_t* lo = hi + kSlotBytes;
    pack_slot_kernel<<<64, 256, 0, c.st>>>(W, ld, transposed, n_off, k_off, n_valid, k_valid, kc,
                                           c.scales + 2 * step, hi, lo, perm16);
    g_launches++;
    c.nslots += 2;
  }
  copy_strided_kernel<<<1, 1, 0, c.st>>>(c.scales + 2 * step + 1, 1, 1, c.inv_scale + step);
  g_launches++;
  return first;
}

void tc_free(Field& f) {
  delete (TcBlob*)f.tc;
  f.tc = nullptr;
}

int tc_pack(Field& f, Arena& a, cudaStream_t st) {
  TcBlob* tb = new TcBlob();
  memset(tb, 0, sizeof(*tb));
  f.tc = tb;
  const int E = f.emb_dim;
  PackCtx c;
  c.a = &a;
  c.st = st;
  c.rc = 0;
  c.nslots = 0;
  c.nlayers = 0;
  c.blob = (uint8_t*)a.take<uint4>((size_t)170 * kSlotBytes / 16);
  c.scales = a.take<float>(2 * 32);
  c.inv_scale = a.take<float>(32);
  float* w8row = a.take<float>(256);
  float* Wrgb = a.take<float>(3 * 256);
  float* b8feat = a.take<float>(256);      // b8[1:], 16-byte aligned copy (the epilogue loads float4)
  MP_REQUIRE(a.ok, "tc_pack: storage too small");
  MP_CHECK_CUDA(cudaMemsetAsync(c.blob, 0, (size_t)170 * kSlotBytes, st));
  TcProgram P;
  memset(&P, 0, sizeof(P));
  P.blob = (const uint4*)c.blob;
  P.inv_scale = c.inv_scale;
  P.d_in = f.d_in;
  P.multires = f.multires;
  P.E = E;
  P.inj_col = kHidden - E;
  P.w8row = w8row;
  P.b8 = f.imp_b[8];
  P.Wrgb = Wrgb;
  P.n_extra = f.ren_extra;
  copy_strided_kernel<<<1, 256, 0, st>>>(f.imp_W[8], 1, 256, w8row);     // W8[0,:]
  g_launches++;
  copy_strided_kernel<<<1, 256, 0, st>>>(f.imp_b[8] + 1, 1, 256, b8feat);
  g_launches++;
  int s = 0;
  const int nk0 = (E + 63) / 64;
  // ---- forward L0..L7 ----
  for (int l = 0; l < 8; ++l) {
    int in = f.imp_in[l], out = f.imp_out[l];
    int nk = (l == 0) ? nk0 : 4;
    pack_layer(c, P.step[s], f.imp_W[l], in, 0, 0, 0, out, (l == 0) ? E : in, nk, out * in);
    P.step[s].nk = nk;
    P.step[s].epi = EPI_SOFTPLUS;
    P.step[s].flags = F_SAVE_SIG | ((l == f.skip_layer - 1) ? F_INJECT_EMB : 0) | ((l == 7) ? F_SDF_DOT : 0);
    P.step[s].sig = l;
    P.step[s].bias = (l == 0) ? f.imp_b0_eff : f.imp_b[l];
    P.step[s].ncols = out;
    ++s;
  }
  // sdf-only program: the first 8 steps, no sigma' stash
  tb->sdf_prog = P;
  tb->sdf_prog.nsteps = 8;
  tb->sdf_prog.slots_per_tile = c.nslots;
  for (int i = 0; i < 8; ++i) tb->sdf_prog.step[i].flags &= ~F_SAVE_SIG;
  // ---- L8 features (operator API: sdf + features) ----
  {
    TcProgram F = P;
    pack_layer(c, F.step[8], f.imp_W[8], 256, 0, 1, 0, 256, 256, 4, 257 * 256);
    F.step[8].nk = 4;
    F.step[8].epi = EPI_FEAT;
    F.step[8].flags = F_FEAT_OUT;
    F.step[8].sig = -1;
    F.step[8].bias = b8feat;
    F.step[8].ncols = 256;
    F.nsteps = 9;
    F.slots_per_tile = c.nslots;
    for (int i = 0; i < 8; ++i) F.step[i].flags &= ~F_SAVE_SIG;
    tb->fwd_prog = F;
  }
  const bool fg_chain = (f.ren_mode == 0) && (f.n_ren == 5) && f.ren_out[0] == 256;
  const bool bg_chain = (f.ren_mode == 1) && (f.n_ren == 2) && f.ren_out[0] <= 256 && f.ren_extra <= 27;
  tb->has_full = fg_chain || bg_chain;
  // The feature layer L8 and the colour layer 0 have no non-linearity in between (networks.py:199-207 -> :281,:305):
  //   C0_pre = Wc0[:, feat] (W8[1:] h7 + b8[1:]) + Wc0[:, extra] extra + b0  =  M h7 + (Wc0f b8f + b0) + ...
  // so the fused chains run ONE layer with M = Wc0[:, feat] . W8[1:, :] instead of two.
  // the colour layer 0 the chains run: [ M | extra-input columns | 0 ]  (256 x 320, K-blocks 0..3 | 4)
  constexpr int kLdM = 320;
  float* Mfold = a.take<float>(256 * kLdM);
  f.ren_cb = a.take<float>(264);
  f.ren_b0_fold = a.take<float>(264);
  MP_REQUIRE(a.ok, "tc_pack: storage too small");
  if (tb->has_full) {
    const int in0 = f.ren_in[0] == 0 ? 0 : (f.ren_mode == 0 ? 6 + 8 + 256 : f.ren_extra + 32 + 256);
    const int coff = f.ren_mode == 0 ? 14 : f.ren_extra + 32;
    const int o0 = f.ren_out[0];
    fold_mm_kernel<<<dim3(256 / 16, div_up(o0, 16)), dim3(16, 16), 0, st>>>(f.ren_W[0], in0, coff, f.imp_W[8] + 256, 256,
                                                                          f.imp_b[8] + 1, o0, Mfold, kLdM, f.ren_cb);
    g_launches++;
    MP_REQUIRE(f.ren_extra <= 64, "tc_pack: more than 64 extra colour inputs");
    fill_extra_cols_kernel<<<64, 256, 0, st>>>(f.ren_Wt[0], o0, f.ren_extra, Mfold, kLdM);
    g_launches++;
  }
  if (bg_chain) {
    // background: folded colour layer 0 (view embedding + h7 -> 128, ReLU) and the rgb head (multiply.py:531)
    const int o0 = f.ren_out[0];
    pack_layer(c, P.step[s], Mfold, kLdM, 0, 0, 0, o0, kLdM, 5, 256 * kLdM);
    P.step[s].nk = 5;
    P.step[s].epi = EPI_RELU;
    P.step[s].flags = F_EXTRA_IN | F_RGB_OUT;
    P.step[s].sig = -1;
    P.step[s].bias = f.ren_b0_fold;
    P.step[s].ncols = o0;
    ++s;
    pad_rows_kernel<<<div_up(3 * 256, 256), 256, 0, st>>>(f.ren_W[1], o0, 3, o0, Wrgb);
    g_launches++;
    P.brgb = f.ren_b[1];
    P.nsteps = s;
    P.slots_per_tile = c.nslots;
    tb->full_prog = P;
    for (int i = 0; i < 8; ++i) tb->full_prog.step[i].flags &= ~F_SAVE_SIG;   // no reverse sweep in the background
  }
  if (fg_chain) {
    // h7 is parked (it returns as the folded colour layer's input) and the reverse sweep starts right after L7
    P.step[s - 1].flags |= F_SEED_BWD | F_STASH_FEAT;
    // ---- reverse sweep B7..B1: g_{l-1} = (g_l * sigma'_l) . W_l ----
    for (int l = 7; l >= 1; --l) {
      int in = f.imp_in[l], out = f.imp_out[l];
      // B[n][k] = W_l[k][n] : n over in (valid in), k over out (valid out)
      pack_layer(c, P.step[s], f.imp_W[l], in, 1, 0, 0, in, out, 4, out * in);
      P.step[s].nk = 4;
      P.step[s].epi = EPI_BWD;
      P.step[s].flags = (l == f.skip_layer) ? F_SKIP_GRAD : 0;
      P.step[s].sig = l - 1;
      P.step[s].bias = nullptr;
      P.step[s].ncols = in;
      ++s;
    }
    // ---- B0: d/d embed = (g_0 * sigma'_0) . W0[:, :E] ----
    MP_REQUIRE(f.d_in <= 4 && 1 + 2 * f.multires <= 14,
               "tc_pack: the final-gradient step keeps 1 + 2 * multires <= 14 embedding columns per axis");
    pack_layer(c, P.step[s], f.imp_W[0], f.imp_in[0], 1, 0, 0, E, 256, 4, 256 * f.imp_in[0], /*perm16=*/f.d_in);
    P.step[s].nk = 4;
    P.step[s].epi = EPI_BWD;
    P.step[s].flags = F_FINAL_GRAD;
    P.step[s].sig = -1;
    P.step[s].bias = nullptr;
    P.step[s].ncols = E;
    ++s;
    // ---- colour net: folded layer 0, then layers 1..3 ----
    for (int l = 0; l < 4; ++l) {
      if (l == 0)
        pack_layer(c, P.step[s], Mfold, kLdM, 0, 0, 0, 256, kLdM, 5, 256 * kLdM);
      else
        pack_layer(c, P.step[s], f.ren_W[l], 256, 0, 0, 0, 256, 256, 4, 256 * 256);
      P.step[s].nk = (l == 0) ? 5 : 4;
      P.step[s].epi = EPI_RELU;
      P.step[s].flags = ((l == 0) ? F_EXTRA_IN : 0) | ((l == 3) ? F_RGB_OUT : 0);
      P.step[s].sig = -1;
      P.step[s].bias = (l == 0) ? f.ren_b0_fold : f.ren_b[l];
      P.step[s].ncols = 256;
      ++s;
    }
    // rgb head [3][256]
    MP_CHECK_CUDA(cudaMemcpyAsync(Wrgb, f.ren_W[4], (size_t)3 * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    P.brgb = f.ren_b[4];
    P.nsteps = s;
    P.slots_per_tile = c.nslots;
    tb->full_prog = P;
  }
  MP_REQUIRE(c.nslots <= 170, "tc_pack: slot budget exceeded (%d)", c.nslots);
  return c.rc;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int tc_trace_read(unsigned long long* out, int n) {
  if (n > 4096) n = 4096;
  MP_CHECK_CUDA(cudaDeviceSynchronize());
  MP_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_trace, (size_t)n * sizeof(unsigned long long)));
  return 0;
}

size_t tc_workspace_bytes(int N) { return (size_t)sm_count() * kScratchPerCta + 4096; }

// optional per-launch timing of the tcgen05 kernel (bench.py roofline): CUDA events on the
// launching stream + an async copy of the device-side point count into pinned memory
struct ProfEntry {
  cudaEvent_t e0, e1;
  int kind;        // 0 sdf-only, 1 forward (sdf+features), 2 full shade, 3 background
  int cap;
  int* host_count; // pinned
};
// The profile is a process-wide diagnostic (bench.py): its state is guarded by g_prof_mu so that concurrent
// launches from several host threads stay well defined.
static std::mutex g_prof_mu;
static std::atomic<bool> g_prof_on{false};
static std::vector<ProfEntry>* g_prof = nullptr;
static int* g_prof_pinned = nullptr;
static int g_prof_used = 0;
constexpr int kProfMax = 1 << 16;

int prof_enable(int on) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  if (on && !g_prof) {
    g_prof = new std::vector<ProfEntry>();
    MP_CHECK_CUDA(cudaMallocHost(&g_prof_pinned, kProfMax * sizeof(int)));
  }
  g_prof_on = on != 0;
  return 0;
}
int prof_read(double* ms, long long* launches, double* points, int reset) {
  for (int k = 0; k < 4; ++k) {
    ms[k] = 0;
    launches[k] = 0;
    points[k] = 0;
  }
  std::lock_guard<std::mutex> g(g_prof_mu);
  if (!g_prof) return 0;
  for (auto& e : *g_prof) {
    MP_CHECK_CUDA(cudaEventSynchronize(e.e1));
    float t = 0.f;
    MP_CHECK_CUDA(cudaEventElapsedTime(&t, e.e0, e.e1));
    ms[e.kind] += t;
    launches[e.kind] += 1;
    int n = e.host_count ? *e.host_count : e.cap;
    points[e.kind] += (double)(n < e.cap ? n : e.cap);
  }
  if (reset) {
    for (auto& e : *g_prof) {
      cudaEventDestroy(e.e0);
      cudaEventDestroy(e.e1);
    }
    g_prof->clear();
    g_prof_used = 0;
  }
  return 0;
}

// Precision mode of the tcgen05 engine (mp_set_precision): which of the three split-precision product terms each layer
// step issues.  The weight blob always holds the hi and lo slots; a single-term step just skips the lo slots.
//   0  parity (default): every step A_hi.W_hi + A_lo.W_hi + A_hi.W_lo  -- RGB / SDF within 1e-4 of the fp32 reference
//   1  colour layers single-term (A_hi.W_hi): SDF / normals unchanged, RGB error ~2e-5 (still inside the gate)
//   2  throughput: every step single-term, i.e. plain fp16 operands with fp32 accumulation -- misses the 1e-4 gate
//      (SDF ~3e-4, normals ~1e-3; measured values in DESIGN.md) and is reported separately by bench.py
std::atomic<int> g_precision{0};

// The tensor core adds each MMA's products into the fp32 accumulator with round-toward-zero: every one of the
// n = 4 nk terms accumulations of a layer step (K = 16 per tcgen05.mma) drops on average half an ulp of the running sum,
// always toward zero.  Unlike round-to-nearest noise this loss is coherent -- every pre-activation shrinks by the same
// relative amount, layer after layer -- and it is what kept the engine's SDF at 4e-6 and its gradients at 6e-6 from
// the fp64 value of the same weights while the fp32 SIMT engine sits at 3e-7.  First-order model: a running sum that
// grows linearly to its final value z loses  sum_i ulp(z i/n)/2 ~= (n/2) * E[ulp(z)/|z|]/2 * |z|, with
// E[ulp/|z|] = 2^-23 / (2 ln 2) for a log-uniform mantissa: 2.15e-8 |z| per accumulation, 1.03e-6 |z| for the 48
// accumulations of a 256-wide three-term layer.  The epilogue multiplies the accumulator by (1 + n * kRzPerMma), which
// is free (it is folded into the 2^-s rescale).  The constant is the model's 2.15e-8 calibrated by one factor measured on
// the device (scripts/gpu_normal_diag.py, 4096 points around the surface and 4096 near the canonical origin, against
// the fp64 evaluation of the same weights; MP_TC_RZ_SCALE sweeps it):
//   factor   SDF L-inf   d sdf/dx L-inf (mean)    rendered normals, 48-ray sample of the benchmark batch
//   0        3.7e-6      6.0e-6 (4.0e-6)          4.8e-4   (one sample at |grad| ~ 1e-4 off by 0.07)
//   0.8      5.8e-7      1.3e-6 (4.3e-7)          3.0e-6   <- kRzPerMma
//   1.0      1.2e-6      2.1e-6 (1.1e-6)          3.7e-6
//   1.2      1.6e-6      2.8e-6 (1.7e-6)          6.9e-6
// (fp32 SIMT engine: 3.1e-7 / 7.1e-7 (1.9e-7) / 2.9e-6.)
constexpr float kRzPerMma = 1.72e-8f;

static int tc_launch(const TcProgram& P0, TcIO io, void* ws, size_t ws_bytes, cudaStream_t st, int kind) {
  TcProgram P = P0;
  {
    const int mode = g_precision.load();
    for (int s = 0; s < P.nsteps; ++s)
      P.step[s].terms = (mode == 2 || (mode == 1 && P.step[s].epi == EPI_RELU)) ? 1 : 3;
  }
  int grid = sm_count();
  {
    static int grid_override = -1;
    if (grid_override < 0) {
      const char* eg = getenv("MP_TC_GRID");     // experiment knob: number of persistent CTAs
      grid_override = eg ? atoi(eg) : 0;
    }
    if (grid_override > 0 && grid_override < grid) grid = grid_override;
  }
  int maxtiles = (io.cap + 127) / 128;
  if (grid > maxtiles) grid = maxtiles;
  if (grid < 1) return 0;
  MP_REQUIRE(ws && ws_bytes >= (size_t)grid * kScratchPerCta, "tcgen05 engine: workspace too small (%zu < %zu)",
             ws_bytes, (size_t)grid * kScratchPerCta);
  io.scratch = (char*)ws;
  io.scratch_per_cta = kScratchPerCta;
  {
    static int knobs = -1;
    static float rz_scale = 1.f;
    if (knobs < 0) {
      const char* er = getenv("MP_TC_RZ_SCALE");      // experiment knob: multiplies kRzPerMma (0 switches it off)
      rz_scale = er ? (float)atof(er) : 1.f;
      const char* ek = getenv("MP_TC_KNOBS");
      knobs = ek ? atoi(ek) : 0;
    }
    io.knobs = knobs;
    io.rz = kRzPerMma * rz_scale;
  }
  {
    // the > 48 KB dynamic shared memory opt-in is per device
    static std::mutex attr_mu;
    static unsigned long long attr_done = 0;      // bit d: cudaFuncSetAttribute done on device d
    int dev = 0;
    MP_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> g(attr_mu);
    if (dev < 0 || dev >= 64 || !((attr_done >> dev) & 1ull)) {
      MP_CHECK_CUDA(cudaFuncSetAttribute(tc_chain_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      if (dev >= 0 && dev < 64) attr_done |= 1ull << dev;
    }
  }
  ProfEntry pe;
  std::unique_lock<std::mutex> prof_lock(g_prof_mu, std::defer_lock);
  if (g_prof_on.load()) prof_lock.lock();
  const bool prof = prof_lock.owns_lock() && g_prof_on.load() && g_prof && g_prof_used < kProfMax;
  if (prof) {
    MP_CHECK_CUDA(cudaEventCreate(&pe.e0));
    MP_CHECK_CUDA(cudaEventCreate(&pe.e1));
    pe.kind = kind;
    pe.cap = io.cap;
    pe.host_count = nullptr;
    if (io.count) {
      pe.host_count = g_prof_pinned + g_prof_used++;
      MP_CHECK_CUDA(cudaMemcpyAsync(pe.host_count, io.count, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    MP_CHECK_CUDA(cudaEventRecord(pe.e0, st));
  }
  tc_chain_kernel<16, true><<<grid, 64 + 32 * 16, kSmemBytes, st>>>(P, io);
  MP_LAUNCH_CHECK();
  if (prof) {
    MP_CHECK_CUDA(cudaEventRecord(pe.e1, st));
    g_prof->push_back(pe);
  }
  return 0;
}

int tc_sdf_list(const Field& f, const float* xc_list, const int* slot_list, const int* count_dev, int cap,
                float* sdf_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  MP_REQUIRE(f.tc, "tcgen05 engine: field not packed");
  TcBlob* tb = (TcBlob*)f.tc;
  TcIO io;
  memset(&io, 0, sizeof(io));
  io.x = xc_list;
  io.slot = slot_list;
  io.count = count_dev;
  io.cap = cap;
  io.sdf_out = sdf_out;
  return tc_launch(tb->sdf_prog, io, ws, ws_bytes, st, 0);
}

int tc_shade_list(const Field& f, const float* xc_list, const int* slot_list, const int* count_dev, int cap,
                  const float* Jinv_list, float* sdf_out, float* rgb_out, float* normal_out, float* grad_out,
                  float* feat_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  MP_REQUIRE(f.tc, "tcgen05 engine: field not packed");
  TcBlob* tb = (TcBlob*)f.tc;
  TcIO io;
  memset(&io, 0, sizeof(io));
  io.x = xc_list;
  io.slot = slot_list;
  io.count = count_dev;
  io.cap = cap;
  io.jinv = Jinv_list;
  io.sdf_out = sdf_out;
  io.rgb_out = rgb_out;
  io.nrm_out = normal_out;
  io.grad_out = grad_out;
  io.feat_out = feat_out;
  if (!Jinv_list && !grad_out) return tc_launch(tb->fwd_prog, io, ws, ws_bytes, st, 1);
  MP_REQUIRE(tb->has_full, "tcgen05 engine: this field has no fused shading program");
  return tc_launch(tb->full_prog, io, ws, ws_bytes, st, 2);
}

int tc_bg(const Field& f, const float* pts, const float* dirs, int N, float* sdf, float* rgb, void* ws,
          size_t ws_bytes, cudaStream_t st) {
  MP_REQUIRE(f.tc, "tcgen05 engine: field not packed");
  TcBlob* tb = (TcBlob*)f.tc;
  MP_REQUIRE(tb->has_full && f.ren_mode == 1, "tcgen05 engine: not a background field");
  TcIO io;
  memset(&io, 0, sizeof(io));
  io.x = pts;
  io.cap = N;
  io.extra = dirs;
  io.sdf_out = sdf;
  io.rgb_out = rgb;
  return tc_launch(tb->full_prog, io, ws, ws_bytes, st, 3);
}

}  // namespace mp

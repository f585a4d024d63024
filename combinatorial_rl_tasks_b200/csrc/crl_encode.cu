// crl_encode.cu -- the consumer side of the observation: ZoneEnvModel's per-zone network and
// mean-pool (main/src/env_model.py:56-78), fused into one sm_100a kernel.
//
//   zone_emb[e] = 1/N * sum_z  L3( relu( L2( relu( L1( [obs[e], zone_obs[e][z]] ))))),   L_i = nn.Linear
//               = L3( pooled[e] ),   pooled[e] = 1/N * sum_z relu( L2( relu( L1( . ))))    (L3 is affine)
//
// The reference materialises three (B*N, h) activations (h = 185 by default,
// main/scripts/train_ppo.py:66).  Here the kernel produces `pooled` (B, h) and nothing else; the
// mean is taken BEFORE the third Linear, which turns L3 from a (B N, h) x (h, h) GEMM into a
// (B, h) x (h, h) one (a third of the tensor work gone): L3 and combine_net_ (env_model.py:79) stay
// plain fp32 library GEMMs on the caller's side.
//
// The two wide GEMMs run TRANSPOSED on the 5th-gen tensor cores: H^T = W X^T, i.e. the weights are
// the A operand (M = hidden units, in blocks of 128 = the TMEM lanes) and a tile of 128 (env, zone) rows
// -- 8 envs of 16 zone slots, or 16 envs of 8 slots when N <= 8; slots beyond N are padding -- is the N
// dimension (TMEM columns).
// tcgen05.mma.cta_group::1.kind::f16: bf16 operands, fp32 accumulators in TMEM, M = 128, N = 128,
// K = 16 per instruction.  Why transposed: an accumulator row lives in ONE thread's registers after
// tcgen05.ld, so with the 16 zone slots of an env on consecutive COLUMNS the mean over zones is 15
// register adds per env -- no shuffles -- and lane j stores hidden unit j, i.e. coalesced rows of the
// output.  (The first version had rows = (env, zone) on the lanes: half of its instructions were a
// shuffle butterfly.)  W1 / W2 stay resident in shared memory for the whole persistent kernel.
//
// Epilogues are the bottleneck of such a narrow MLP (K <= 192 gives the tensor pipe ~1,700 cycles per
// tile while 2 x 96 KB of accumulator come back through registers), so they are minimal:
//  * the BIASES ARE FOLDED INTO THE GEMMS -- the K dimension carries constant-1 entries (layer 1:
//    input columns in_dim, in_dim + 1; layer 2: hidden rows h, h + 1, produced by two "generator" rows
//    of W1) that multiply the bias split into a bf16 high and low part;
//  * epilogue 1 = one cvt.rn.relu.bf16x2.f32 per pair of values + 16-byte stores: 8 consecutive rows m of
//    one hidden unit j are exactly one 16-byte unit of the layer-2 B operand in MN-major layout;
//  * epilogue 2 = one max per value + the register adds + two coalesced stores per 32 values.
// A padding row is all zeros including its ones, hence stays exactly zero through both layers and
// drops out of the mean by itself.
//
// A CTA = two independent groups of 256 threads, each walking its own sequence of tiles with its own
// accumulators, operand buffers and mbarrier: while one group is in an epilogue (CUDA cores) the other
// group's MMAs occupy the tensor pipe.  Warp w of a group owns TMEM lane quadrant w % 4 of M-block
// w / 4.  The next tile's inputs are prefetched into registers.  Every wait is bounded.
//
// A separate translation unit on purpose: nothing here can move the register allocation of the
// step kernels (crl_kernels.cu; see profiles/r01_notes.md on how easily that happens).
//
// Shared-memory operand images (no swizzle; canonical layouts of cute::UMMA, mma_traits_sm100.hpp), all
// made of 8 x 8 core matrices of 128 contiguous bytes:
//  K-major (W1, W2 as A; the input rows X as B of layer 1): element (r, k) of an R x K matrix at
//      (r % 8) * 16 + (r / 8) * (16 K) + (k / 8) * 128 + (k % 8) * 2      LBO = 128 (next core matrix along K),
//                                                                         SBO = 16 K (next 8 rows)
//  MN-major (the layer-1 activations H1 as B of layer 2, N = row m, K = hidden unit j): element (m, j) at
//      (j % 8) * 16 + (j / 8) * 2048 + (m / 8) * 128 + (m % 8) * 2        LBO = 2048 (next 8 j), SBO = 128 (next 8 m)
// One tcgen05.mma consumes K = 16 = two core matrices along K.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/crl_b200.h"
#include "crl_core.cuh"

namespace crl_enc {

constexpr int kRows = 128;          // (env, zone) rows of a tile = the N of every MMA = accumulator columns
constexpr int kGroupThreads = 256;  // 8 warps: one per (M-block, TMEM lane quadrant)
constexpr int kGroups = 2;          // independent groups per CTA
// an env owns S = 8 (N <= 8) or 16 consecutive row slots of a tile; the slots beyond N are padding rows
__host__ __device__ inline int slots_per_env(int n_zones) { return n_zones <= 8 ? 8 : 16; }
// padded input width of layer 1: the per-env features (obs, and goal / one-hot skill for the reference's
// ZoneEnvGoalModel / ZoneEnvSkillModel, which the caller concatenates to obs), the zone row and at
// least one ones column; one or two K = 16 steps
__host__ __device__ inline int padded_k1(int in_dim) { return in_dim + 1 <= 16 ? 16 : 32; }
constexpr uint32_t kTmemCols = 512; // two M-blocks x 128 columns per group; the CTA owns the SM
constexpr uint32_t kSpinLimit = 1u << 24;

// K of layer 2: the h hidden units plus the two ones rows, multiple of 16
__host__ __device__ inline int padded_k(int h) { return (h + 2 + 15) & ~15; }
// M of both layers: blocks of 128 lanes covering padded_k (the ones rows are outputs of layer 1)
__host__ __device__ inline int padded_m(int h) { return (padded_k(h) + 127) & ~127; }

// byte offset of element (r, k) in the K-major image of a matrix with K columns
__host__ __device__ inline uint32_t canon(int r, int k, int K) {
  return (uint32_t)((r & 7) * 16 + (r >> 3) * (16 * K) + (k >> 3) * 128 + (k & 7) * 2);
}

struct Offsets {   // byte offsets: packed weight buffer == start of shared memory; then per-group buffers
  uint32_t w2, w1, packed_end, group0, h1, xbuf, bar, group_bytes, tmem_slot, smem_end;
};
__host__ __device__ inline Offsets offsets(int h, int in_dim) {
  const int KP = padded_k(h), MP = padded_m(h), kK1 = padded_k1(in_dim);
  Offsets o;
  o.w2 = 0;
  o.w1 = o.w2 + (uint32_t)MP * KP * 2;
  o.packed_end = o.w1 + (uint32_t)MP * kK1 * 2;
  o.group0 = (o.packed_end + 127u) & ~127u;
  o.h1 = 0;                                              // relative to the group's base
  o.xbuf = o.h1 + (uint32_t)KP * kRows * 2;
  o.bar = o.xbuf + (uint32_t)kRows * kK1 * 2;
  o.group_bytes = (o.bar + 8 + 127u) & ~127u;
  o.tmem_slot = o.group0 + kGroups * o.group_bytes;
  o.smem_end = o.tmem_slot + 16;
  return o;
}

// ---- packing: torch-layout fp32 weights -> the shared-memory image -------------------------
struct PackArgs {
  const float *w1, *b1, *w2, *b2;
  uint8_t* out;
  int in_dim, h;
};

__global__ void pack_kernel(const PackArgs a) {
  const Offsets o = offsets(a.h, a.in_dim);
  const int KP = padded_k(a.h), MP = padded_m(a.h), kK1 = padded_k1(a.in_dim);
  const int n_w = MP * KP, n_w1 = MP * kK1;
  const int total = n_w + n_w1;
  // bias = hi + lo with hi = bf16(bias), lo = bf16(bias - hi): entry `ones` of the K dimension carries 1
  // and multiplies hi, entry `ones + 1` multiplies lo (layer 1 has the second one only if in_dim < 15)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const bool second = i < n_w;
    const int K = second ? KP : kK1, ones = second ? a.h : a.in_dim;
    const int j = second ? i : i - n_w, n = j / K, k = j % K;
    const float* w = second ? a.w2 : a.w1;
    const float* b = second ? a.b2 : a.b1;
    float v = 0.f;
    if (n < a.h) {
      if (k < ones) v = w[(size_t)n * ones + k];
      else if (k == ones) v = __bfloat162float(__float2bfloat16_rn(b[n]));
      else if (k == ones + 1) v = b[n] - __bfloat162float(__float2bfloat16_rn(b[n]));
    } else if (!second && (n == a.h || n == a.h + 1) && k == ones) {
      v = 1.f;                                   // generator rows of W1: layer 2's ones entries h, h + 1
    }
    *reinterpret_cast<__nv_bfloat16*>(a.out + (second ? o.w2 : o.w1) + canon(n, k, K)) = __float2bfloat16_rn(v);
  }
}

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: start address, LBO, SBO, version 1 (sm_100), no swizzle
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by one thread
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// bounded wait on an mbarrier phase: returns false if it never completed (the caller reports it)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; it < kSpinLimit; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
// this warp's 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread.  Asynchronous:
// tmem_ld_wait(v) before v is read (several loads may be in flight).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
// the registers are in/out operands so that no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
      : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
        "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
        "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
        "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
      :: "memory");
}
// barrier over the 256 threads of one group (barrier 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int group) { asm volatile("bar.sync %0, 256;" :: "r"(group + 1) : "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);   // .x (low half) = lo
  return *reinterpret_cast<const uint32_t*>(&t);
}
// {bf16(max(lo, 0)), bf16(max(hi, 0))} in one instruction (the first source lands in the upper half)
__device__ __forceinline__ uint32_t relu_pack_bf16(uint32_t lo_bits, uint32_t hi_bits) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
  return d;
}

struct EncArgs {
  const float* obs;        // [B][obs_dim]
  const float* zone_obs;   // [B][N][Z]
  const uint8_t* packed;
  float* out;              // [B][h]: pooled hidden activation
  int* status;             // device int: set to 1 if a tensor-core wait expired
  int B, N, Z, obs_dim, h, n_tiles, S;   // S = slots_per_env(N)
  // crl_zone_encode_state: the zone part of the input rows is built from the STATE planes instead of being read
  // from a materialised zone_obs (aux != nullptr selects it; zone_obs is then not touched)
  const float4* aux;           // float4[B]: .w = steps | visited mask / colour codes << 16
  const float2* zone_xy;       // float2[N][B]
  const uint32_t* zone_tmax;   // uint32[ceil(N/2)][B] (TimedTSP)
  const uint2* cooldown;       // uint2[B] (ColourMatch)
  int task, num_steps;
  crl::DivConst div_steps, div_cd;
};

// Feature f of zone `slot` of env e from the state planes: EXACTLY what step_kernel's zone_row writes to
// zone_obs[e][slot][f] (crl_kernels.cu; tests/test_gpu_encode.py compares the two paths bit for bit):
// x/3, y/3, r, g, b, 0.25 [, time left | cooldown / max_cooldown]
__device__ __forceinline__ void zone_features_from_state(const EncArgs& a, int e, int slot, float (&z)[7]) {
  const float2 c = __ldg(a.zone_xy + (size_t)slot * a.B + e);
  const uint32_t bits = (uint32_t)__float_as_int(__ldg(reinterpret_cast<const float*>(a.aux + e) + 3));
  const int steps = (int)(bits & 0xffffu);
  const uint32_t hi = bits >> 16;
  const float third = 1.0f / 3.0f;
  z[0] = c.x * third; z[1] = c.y * third; z[5] = 0.25f; z[6] = 0.f;
  if (a.task == CRL_TASK_CM) {
    const uint32_t col = (hi >> (2 * slot)) & 3u;               // 0 B, 1 G, 2 R
    z[2] = col == 2u ? 1.f : 0.f; z[3] = col == 1u ? 1.f : 0.f; z[4] = col == 0u ? 1.f : 0.f;
    const uint2 cd = __ldg(a.cooldown + e);
    const uint32_t w = slot < 4 ? cd.x : cd.y;
    z[6] = crl::div_const((float)((w >> (8 * (slot & 3))) & 0xffu), a.div_cd);
  } else {
    const bool v = (hi >> slot) & 1u;                          // Yellow (1,1,0) visited / Cyan (0,1,1)
    z[2] = v ? 1.f : 0.f; z[3] = 1.f; z[4] = v ? 0.f : 1.f;
    if (a.task == CRL_TASK_TTSP) {
      const uint32_t w = __ldg(a.zone_tmax + (size_t)(slot >> 1) * a.B + e);
      const int tm = (int)((w >> (16 * (slot & 1))) & 0xffffu);
      z[6] = v ? 1.0f : crl::div_const((float)(tm - steps), a.div_steps);
    }
  }
}

// eight consecutive values (k = 8 kc .. 8 kc + 7) of row (e, slot) of the layer-1 input
// [obs[e], zone_obs[e][slot], 1, 1, 0...]; all zeros -- the ones included -- for a padding row.
// STATE: the zone part comes from the state planes (obs_dim == 8, so chunk 1 is exactly the zone features + ones);
// a template parameter so that the materialised path compiles exactly as it did without the feature.
template <bool STATE>
__device__ __forceinline__ void load_half_row(const EncArgs& a, int tile, int m, int kc, float (&x)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = 0.f;
  const int e = tile * (kRows / a.S) + m / a.S, slot = m % a.S;
  if (tile < a.n_tiles && e < a.B && slot < a.N) {
    const float* ob = a.obs + (size_t)e * a.obs_dim;
    if (STATE) {
      if (kc == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __ldg(ob + j);
      } else {
        float z[7];
        zone_features_from_state(a, e, slot, z);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = j < 7 && j < a.Z ? z[j < 7 ? j : 0] : (j <= a.Z + 1 ? 1.f : 0.f);
      }
    } else {
      const float* zo = a.zone_obs + ((size_t)e * a.N + slot) * a.Z;
      const int in_dim = a.obs_dim + a.Z;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = 8 * kc + j;
        if (k < a.obs_dim) x[j] = __ldg(ob + k);
        else if (k < in_dim) x[j] = __ldg(zo + (k - a.obs_dim));
        else if (k <= in_dim + 1) x[j] = 1.f;
      }
    }
  }
}

// relu of 32 consecutive rows m of this thread's hidden unit -> bf16 -> four 16-byte units of H1
__device__ __forceinline__ void relu_to_h1(const uint32_t (&v)[32], uint8_t* h1_row, int c) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    *reinterpret_cast<uint4*>(h1_row + (uint32_t)((c * 4 + q) * 128)) =
        make_uint4(relu_pack_bf16(v[q * 8 + 0], v[q * 8 + 1]), relu_pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                   relu_pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), relu_pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
  }
}

// 32 consecutive rows m = the zone slots of envs e0 .. e0 + 32 / S - 1: relu, sum in registers, store column j
__device__ __forceinline__ void relu_pool_store(const uint32_t (&v)[32], const EncArgs& a, int e0, int j, float inv_n) {
  float q[4];                                                 // sums of 8 consecutive rows
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      s[i] = fmaxf(__uint_as_float(v[8 * g + 2 * i]), 0.f) + fmaxf(__uint_as_float(v[8 * g + 2 * i + 1]), 0.f);
    q[g] = (s[0] + s[1]) + (s[2] + s[3]);
  }
  if (j < a.h) {
    if (a.S == 16) {
      if (e0 < a.B) a.out[(size_t)e0 * a.h + j] = (q[0] + q[1]) * inv_n;
      if (e0 + 1 < a.B) a.out[(size_t)(e0 + 1) * a.h + j] = (q[2] + q[3]) * inv_n;
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (e0 + g < a.B) a.out[(size_t)(e0 + g) * a.h + j] = q[g] * inv_n;
    }
  }
}

template <bool STATE>
__global__ void __launch_bounds__(kGroupThreads * kGroups, 1) zone_encode_kernel(const EncArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const Offsets o = offsets(a.h, a.obs_dim + a.Z);
  const int KP = padded_k(a.h), MP = padded_m(a.h), kK1 = padded_k1(a.obs_dim + a.Z);
  const int n_mblocks = MP / 128;
  const int group = threadIdx.x / kGroupThreads;              // 0 / 1
  const int t = threadIdx.x % kGroupThreads;
  const int m = t & (kRows - 1), half = t >> 7;               // input row / half of it this thread stages
  const int warp = t >> 5, lane = t & 31;
  const int mblock = warp >> 2, quad = warp & 3;              // accumulator rows this warp drains
  const int j = mblock * 128 + quad * 32 + lane;              // hidden unit (accumulator row) of this thread
  uint8_t* gbase = smem + o.group0 + (uint32_t)group * o.group_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(gbase + o.bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + o.tmem_slot);
  const uint32_t bar_addr = smem_u32(bar);

  // ---- one-time setup: barriers, TMEM, resident weights ------------------------------------
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_addr) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.packed);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = threadIdx.x; i < o.packed_end / 16; i += kGroupThreads * kGroups) dst[i] = __ldg(src + i);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc = tmem_base + (uint32_t)group * 256u;    // this group's accumulators: M-block b at + 128 b
  const uint32_t my_acc = acc + (uint32_t)(mblock * 128) + ((uint32_t)(quad * 32) << 16);
  // instruction descriptors: D fp32, A and B bf16, M = 128, N = 128; layer 1: both K-major; layer 2: B MN-major
  const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kRows >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc2 = idesc1 | (1u << 16);
  uint8_t* h1buf = gbase + o.h1;
  uint8_t* xbuf = gbase + o.xbuf;
  const uint32_t h1_addr = smem_u32(h1buf), x_addr = smem_u32(xbuf);
  const uint32_t w1_addr = smem_u32(smem + o.w1), w2_addr = smem_u32(smem + o.w2);
  const uint32_t x_off = (uint32_t)((m & 7) * 16 + (m >> 3) * (16 * kK1) + half * 128);
  const uint32_t h1_off = (uint32_t)((j & 7) * 16 + (j >> 3) * 2048);
  const bool drains1 = mblock < n_mblocks && mblock * 128 + quad * 32 < KP;     // layer 2 reads hidden rows < KP
  const bool drains2 = mblock < n_mblocks && mblock * 128 + quad * 32 < a.h;    // the output has h columns
  const float inv_n = 1.0f / (float)a.N;
  uint32_t parity = 0u;
  bool healthy = true;

  const int tile_stride = kGroups * gridDim.x;
  int tile = kGroups * blockIdx.x + group;
  float x[8], x2[8];                                         // 16-byte units `half` and, for 32-wide inputs, `half + 2`
  load_half_row<STATE>(a, tile, m, half, x);
  if (!STATE && kK1 == 32) load_half_row<STATE>(a, tile, m, half + 2, x2);
  for (; tile < a.n_tiles; tile += tile_stride) {
    // ---- layer-1 B operand (the tile's 128 input rows) from the prefetched registers ----------
    *reinterpret_cast<uint4*>(xbuf + x_off) =
        make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
    if (kK1 == 32)
      *reinterpret_cast<uint4*>(xbuf + x_off + 256) =
          make_uint4(pack_bf16(x2[0], x2[1]), pack_bf16(x2[2], x2[3]), pack_bf16(x2[4], x2[5]), pack_bf16(x2[6], x2[7]));
    fence_async_smem();
    tc_fence_before();
    group_sync(group);
    // ---- layer 1: H1^T[128 b ..][m] = W1[128 b ..][16] X^T -------------------------------------
    if (t == 0) {
      tc_fence_after();
      for (int b = 0; b < n_mblocks; ++b)
        for (int s = 0; s < kK1 / 16; ++s)
          mma_bf16(acc + (uint32_t)(b * 128), smem_desc(w1_addr + (uint32_t)(b * 16 * 16 * kK1) + 256u * s, 128u, 16 * kK1),
                   smem_desc(x_addr + 256u * s, 128u, 16 * kK1), idesc1, s > 0);
      mma_commit(bar_addr);
    }
    load_half_row<STATE>(a, tile + tile_stride, m, half, x);  // prefetch: in flight for the rest of the tile
    if (!STATE && kK1 == 32) load_half_row<STATE>(a, tile + tile_stride, m, half + 2, x2);
    healthy = mbar_wait(bar_addr, parity) && healthy;
    parity ^= 1u;
    tc_fence_after();
    if (drains1) {
      // this thread: hidden unit j, rows m = 32 c .. 32 c + 31 -> relu -> bf16 -> four 16-byte units of H1
#ifdef CRL_ENC_DIAG_HALF_DRAIN                                // timing diagnostic (wrong results): half of epilogue 1's TMEM reads
      const int c_stop = kRows / 64;
#else
      const int c_stop = kRows / 32;
#endif
#pragma unroll 1
      for (int c = 0; c < c_stop; c += 2) {                   // two TMEM loads in flight
        uint32_t v0[32], v1[32];
        tmem_ld32(my_acc + (uint32_t)(c * 32), v0);
        tmem_ld32(my_acc + (uint32_t)(c * 32 + 32), v1);
        tmem_ld_wait(v0);
        relu_to_h1(v0, h1buf + h1_off, c);
        tmem_ld_wait(v1);
        relu_to_h1(v1, h1buf + h1_off, c + 1);
      }
    }
    fence_async_smem();
    tc_fence_before();
    group_sync(group);
    // ---- layer 2: H2^T[128 b ..][m] = W2[128 b ..][KP] H1^T (layer 1's values have been read) ---
    if (t == 0) {
      tc_fence_after();
#ifdef CRL_ENC_DIAG_HALF_MMA                                  // timing diagnostic (wrong results): half of layer 2's K steps
      const int s_stop = KP / 32;
#else
      const int s_stop = KP / 16;
#endif
      for (int b = 0; b < n_mblocks; ++b)
        for (int s = 0; s < s_stop; ++s)
          mma_bf16(acc + (uint32_t)(b * 128), smem_desc(w2_addr + (uint32_t)(b * 16 * 16 * KP) + 256u * s, 128u, 16 * KP),
                   smem_desc(h1_addr + 4096u * s, 2048u, 128u), idesc2, s > 0);
      mma_commit(bar_addr);
    }
    healthy = mbar_wait(bar_addr, parity) && healthy;
    parity ^= 1u;
    tc_fence_after();
    // ---- ReLU, mean over each env's 16 zone slots (register adds), coalesced stores -----------
    if (drains2) {
#pragma unroll 1
      for (int c = 0; c < kRows / 32; c += 2) {               // two TMEM loads in flight
        uint32_t v0[32], v1[32];
        tmem_ld32(my_acc + (uint32_t)(c * 32), v0);
        tmem_ld32(my_acc + (uint32_t)(c * 32 + 32), v1);
        tmem_ld_wait(v0);
        const int per_chunk = 32 / a.S, e0 = tile * (kRows / a.S) + c * per_chunk;
        relu_pool_store(v0, a, e0, j, inv_n);
        tmem_ld_wait(v1);
        relu_pool_store(v1, a, e0 + per_chunk, j, inv_n);
      }
    }
    // the next tile's layer-1 MMA overwrites the accumulators: ordered after these loads by the fence
    // and the group barrier that precede it
  }
  if (!healthy && a.status) atomicExch(a.status, 1);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---- the head: forward() = combine_net_([obs, L3(pooled)]) (env_model.py:79) as ONE more tcgen05 GEMM ----------
// Both Linears are affine, so  out[e] = Wc [obs[e], W3 pooled[e] + b3] + bc = W' [obs[e], pooled[e], 1, 1]  with
// W' = [Wc_obs | Wc_emb W3 | bias hi | bias lo] folded once on the host side (encoder.py).  The same transposed
// formulation as above: out^T[j][e] = W'[j][:] X^T, the folded weights resident in shared memory as the A operand
// (M = output units in blocks of 128 = TMEM lanes), a tile of 128 ENVS the N dimension, K = obs_dim + h + 2 padded to
// 16.  After tcgen05.ld lane j holds unit j of 32 consecutive envs: 32 coalesced 128-byte stores per load.  4.7 GFLOP
// at 65,536 envs and h = 185: the kernel is bound by reading (obs, pooled) and writing out once (~100 MB), not by the
// tensor pipe -- the library fp32 GEMM it replaces took longer than the whole fused zone kernel (profiles/r01_notes.md).
__host__ __device__ inline int head_k(int obs_dim, int h) { return (obs_dim + h + 2 + 15) & ~15; }

struct HeadOffsets { uint32_t w, packed_end, group0, xbuf, bar, group_bytes, tmem_slot, smem_end; };
__host__ __device__ inline HeadOffsets head_offsets(int obs_dim, int h) {
  const int KH = head_k(obs_dim, h), MP = padded_m(h);
  HeadOffsets o;
  o.w = 0;
  o.packed_end = (uint32_t)MP * KH * 2;
  o.group0 = (o.packed_end + 127u) & ~127u;
  o.xbuf = 0;
  o.bar = (uint32_t)kRows * KH * 2;
  o.group_bytes = (o.bar + 8 + 127u) & ~127u;
  o.tmem_slot = o.group0 + kGroups * o.group_bytes;
  o.smem_end = o.tmem_slot + 16;
  return o;
}

struct HeadPackArgs { const float *w, *b; uint8_t* out; int obs_dim, h; };

// W' rows n < h: [w[n][0 .. obs_dim + h), bf16(b[n]), b[n] - bf16(b[n]), 0...]; rows >= h: zeros
__global__ void head_pack_kernel(const HeadPackArgs a) {
  const int KH = head_k(a.obs_dim, a.h), MP = padded_m(a.h), in = a.obs_dim + a.h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < MP * KH; i += gridDim.x * blockDim.x) {
    const int n = i / KH, k = i % KH;
    float v = 0.f;
    if (n < a.h) {
      if (k < in) v = a.w[(size_t)n * in + k];
      else if (k == in) v = __bfloat162float(__float2bfloat16_rn(a.b[n]));
      else if (k == in + 1) v = a.b[n] - __bfloat162float(__float2bfloat16_rn(a.b[n]));
    }
    *reinterpret_cast<__nv_bfloat16*>(a.out + canon(n, k, KH)) = __float2bfloat16_rn(v);
  }
}

struct HeadArgs {
  const float* obs;      // [B][obs_dim]
  const float* pooled;   // [B][h]
  const uint8_t* packed;
  float* out;            // [B][h]
  int* status;
  int B, obs_dim, h, n_tiles;
};

__global__ void __launch_bounds__(kGroupThreads * kGroups, 1) encoder_head_kernel(const HeadArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const HeadOffsets o = head_offsets(a.obs_dim, a.h);
  const int KH = head_k(a.obs_dim, a.h), MP = padded_m(a.h), in = a.obs_dim + a.h;
  const int n_mblocks = MP / 128, chunks = KH / 8;
  const int group = threadIdx.x / kGroupThreads, t = threadIdx.x % kGroupThreads;
  const int warp = t >> 5, lane = t & 31, mblock = warp >> 2, quad = warp & 3;
  const int j = mblock * 128 + quad * 32 + lane;              // output unit (accumulator row) of this thread
  uint8_t* gbase = smem + o.group0 + (uint32_t)group * o.group_bytes;
  uint8_t* xbuf = gbase + o.xbuf;
  uint64_t* bar = reinterpret_cast<uint64_t*>(gbase + o.bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + o.tmem_slot);
  const uint32_t bar_addr = smem_u32(bar);
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_addr) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.packed);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = threadIdx.x; i < o.packed_end / 16; i += kGroupThreads * kGroups) dst[i] = __ldg(src + i);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc = tmem_base + (uint32_t)group * 256u;
  const uint32_t my_acc = acc + (uint32_t)(mblock * 128) + ((uint32_t)(quad * 32) << 16);
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kRows >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t x_addr = smem_u32(xbuf), w_addr = smem_u32(smem + o.w);
  const bool drains = mblock < n_mblocks && mblock * 128 + quad * 32 < a.h;
  uint32_t parity = 0u;
  bool healthy = true;
  for (int tile = kGroups * blockIdx.x + group; tile < a.n_tiles; tile += kGroups * gridDim.x) {
    // ---- B operand: the tile's 128 rows [obs, pooled, 1, 1, 0..] as bf16, K-major; consecutive threads take
    // consecutive 8-value chunks of one env (coalesced reads); a row beyond the batch is all zeros.  Four chunks
    // per iteration with UNCONDITIONAL loads from clamped addresses (the selects come after), so that 32 loads
    // are in flight per thread: this kernel is bound by how fast it reads (obs, pooled), not by the tensor pipe.
    for (int u0 = t; u0 < kRows * chunks; u0 += 4 * kGroupThreads) {
      float x[4][8];
      int mm[4], cc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int u = u0 + i * kGroupThreads;
        const int m = u / chunks, c = u - m * chunks, e = tile * kRows + m;
        mm[i] = m; cc[i] = u < kRows * chunks ? c : -1;
        const size_t e_ok = (size_t)(e < a.B ? e : 0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = 8 * c + q;
          const float* src = k < a.obs_dim ? a.obs + e_ok * a.obs_dim + k
                                           : a.pooled + e_ok * a.h + (k < in ? k - a.obs_dim : 0);
          x[i][q] = __ldg(src);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (cc[i] < 0) continue;
        const int e = tile * kRows + mm[i];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = 8 * cc[i] + q;
          x[i][q] = e < a.B ? (k < in ? x[i][q] : (k <= in + 1 ? 1.f : 0.f)) : 0.f;
        }
        *reinterpret_cast<uint4*>(xbuf + (uint32_t)((mm[i] & 7) * 16 + (mm[i] >> 3) * (16 * KH) + cc[i] * 128)) =
            make_uint4(pack_bf16(x[i][0], x[i][1]), pack_bf16(x[i][2], x[i][3]), pack_bf16(x[i][4], x[i][5]),
                       pack_bf16(x[i][6], x[i][7]));
      }
    }
    fence_async_smem();
    tc_fence_before();
    group_sync(group);
    if (t == 0) {
      tc_fence_after();
      for (int b = 0; b < n_mblocks; ++b)
        for (int s = 0; s < KH / 16; ++s)
          mma_bf16(acc + (uint32_t)(b * 128), smem_desc(w_addr + (uint32_t)(b * 16 * 16 * KH) + 256u * s, 128u, 16 * KH),
                   smem_desc(x_addr + 256u * s, 128u, 16 * KH), idesc, s > 0);
      mma_commit(bar_addr);
    }
    healthy = mbar_wait(bar_addr, parity) && healthy;
    parity ^= 1u;
    tc_fence_after();
    if (drains) {
#pragma unroll 1
      for (int c = 0; c < kRows / 32; c += 2) {               // two TMEM loads in flight
        uint32_t v0[32], v1[32];
        tmem_ld32(my_acc + (uint32_t)(c * 32), v0);
        tmem_ld32(my_acc + (uint32_t)(c * 32 + 32), v1);
        tmem_ld_wait(v0);
        if (j < a.h) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int e = tile * kRows + c * 32 + i;
            if (e < a.B) a.out[(size_t)e * a.h + j] = __uint_as_float(v0[i]);
          }
        }
        tmem_ld_wait(v1);
        if (j < a.h) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int e = tile * kRows + (c + 1) * 32 + i;
            if (e < a.B) a.out[(size_t)e * a.h + j] = __uint_as_float(v1[i]);
          }
        }
      }
    }
    // the next tile's MMAs overwrite the accumulators and the operand buffer: ordered after these loads by the
    // fence and the group barrier that precede them
    tc_fence_before();
  }
  if (!healthy && a.status) atomicExch(a.status, 1);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

static int check_shape(const CrlEncoderShape* s) {
  if (!s) return CRL_ERR_NULL;
  if (s->obs_dim <= 0 || s->zone_dim <= 0 || s->obs_dim + s->zone_dim + 1 > 32) return CRL_ERR_CONFIG;   // + a ones column
  if (s->num_zones <= 0 || s->num_zones > 16) return CRL_ERR_CONFIG;
  if (s->hidden <= 0) return CRL_ERR_CONFIG;
  if (padded_k(s->hidden) > 192) return CRL_ERR_UNSUPPORTED;   // resident W2 + two groups' operand buffers must fit an SM
  return CRL_OK;
}

}  // namespace crl_enc

using namespace crl_enc;

extern "C" {

int crl_encoder_packed_bytes(const CrlEncoderShape* s, int64_t* bytes) {
  const int rc = check_shape(s);
  if (rc) return rc;
  if (!bytes) return CRL_ERR_NULL;
  *bytes = (int64_t)offsets(s->hidden, s->obs_dim + s->zone_dim).packed_end;
  return CRL_OK;
}

int crl_encoder_pack(const CrlEncoderShape* s, const float* w1, const float* b1, const float* w2, const float* b2,
                     void* packed, void* stream) {
  const int rc = check_shape(s);
  if (rc) return rc;
  if (!w1 || !b1 || !w2 || !b2 || !packed) return CRL_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(packed) & 15u) return CRL_ERR_ALIGN;
  PackArgs a{w1, b1, w2, b2, static_cast<uint8_t*>(packed), s->obs_dim + s->zone_dim, s->hidden};
  pack_kernel<<<148, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? CRL_OK : CRL_ERR_LAUNCH;
}

static int zone_encode_launch(const CrlEncoderShape* s, int32_t num_envs, const float* obs, const float* zone_obs,
                              const CrlConfig* cfg, const CrlState* st, const void* packed, float* pooled, int32_t* status,
                              void* stream) {
  const int rc = check_shape(s);
  if (rc) return rc;
  if (!obs || (!zone_obs && !st) || !packed || !pooled) return CRL_ERR_NULL;
  if (num_envs <= 0) return CRL_ERR_CONFIG;
  if (reinterpret_cast<uintptr_t>(packed) & 15u) return CRL_ERR_ALIGN;
  const Offsets o = offsets(s->hidden, s->obs_dim + s->zone_dim);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return CRL_ERR_DEVICE;
  static bool attr_set[64] = {false};                       // per device: opt in to > 48 KB of dynamic shared memory
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    if (cudaFuncSetAttribute(zone_encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(zone_encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return CRL_ERR_DEVICE;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  EncArgs a{};
  a.obs = obs; a.zone_obs = zone_obs; a.packed = static_cast<const uint8_t*>(packed); a.out = pooled; a.status = status;
  a.B = num_envs; a.N = s->num_zones; a.Z = s->zone_dim; a.obs_dim = s->obs_dim; a.h = s->hidden;
  a.S = slots_per_env(s->num_zones);
  a.n_tiles = (num_envs + kRows / a.S - 1) / (kRows / a.S);
  if (st) {
    // the zone features come from the state planes: the shapes must be the env's own
    if (!cfg || !st->aux || !st->zone_xy) return CRL_ERR_NULL;
    if (s->obs_dim != 8 || cfg->num_envs != num_envs || cfg->num_zones != s->num_zones || cfg->task < 0 || cfg->task > 2 ||
        s->zone_dim != (cfg->task == CRL_TASK_TSP ? 6 : 7) || cfg->num_steps <= 0 || cfg->num_steps > 65534 ||
        cfg->max_cooldown < 0 || cfg->max_cooldown > 255)
      return CRL_ERR_CONFIG;
    if (cfg->task == CRL_TASK_TTSP && !st->zone_tmax) return CRL_ERR_NULL;
    if (cfg->task == CRL_TASK_CM && !st->cooldown) return CRL_ERR_NULL;
    a.aux = reinterpret_cast<const float4*>(st->aux);
    a.zone_xy = reinterpret_cast<const float2*>(st->zone_xy);
    a.zone_tmax = st->zone_tmax;
    a.cooldown = reinterpret_cast<const uint2*>(st->cooldown);
    a.task = cfg->task; a.num_steps = cfg->num_steps;
    // the same exact-division constants the step kernel uses (proven on the host once per divisor)
    static struct { int d, lo, hi; crl::DivConst k; } cache[8];
    static int used = 0;
    auto div_for = [&](int d, int lo, int hi) {
      for (int i = 0; i < used; ++i)
        if (cache[i].d == d && cache[i].lo == lo && cache[i].hi == hi) return cache[i].k;
      const crl::DivConst k = crl::make_div_const(d, lo, hi);
      if (used < 8) { cache[used].d = d; cache[used].lo = lo; cache[used].hi = hi; cache[used].k = k; ++used; }
      return k;
    };
    a.div_steps = div_for(cfg->num_steps, -65535, 65535);
    a.div_cd = div_for(cfg->max_cooldown > 0 ? cfg->max_cooldown : 1, 0, 255);
    if (!a.div_steps.exact || !a.div_cd.exact) return CRL_ERR_CONFIG;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int want = (a.n_tiles + kGroups - 1) / kGroups;
  const int grid = want < sms ? want : sms;                   // persistent: one CTA per SM, weights loaded once
  if (st) zone_encode_kernel<true><<<grid, kGroupThreads * kGroups, o.smem_end, static_cast<cudaStream_t>(stream)>>>(a);
  else zone_encode_kernel<false><<<grid, kGroupThreads * kGroups, o.smem_end, static_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? CRL_OK : CRL_ERR_LAUNCH;
}

int crl_zone_encode(const CrlEncoderShape* s, int32_t num_envs, const float* obs, const float* zone_obs,
                    const void* packed, float* pooled, int32_t* status, void* stream) {
  if (!zone_obs) return CRL_ERR_NULL;
  return zone_encode_launch(s, num_envs, obs, zone_obs, nullptr, nullptr, packed, pooled, status, stream);
}

int crl_zone_encode_state(const CrlEncoderShape* s, const CrlConfig* cfg, const CrlState* st, const float* obs,
                          const void* packed, float* pooled, int32_t* status, void* stream) {
  if (!cfg || !st) return CRL_ERR_NULL;
  return zone_encode_launch(s, cfg->num_envs, obs, nullptr, cfg, st, packed, pooled, status, stream);
}

int crl_encoder_head_packed_bytes(const CrlEncoderShape* s, int64_t* bytes) {
  const int rc = check_shape(s);
  if (rc) return rc;
  if (!bytes) return CRL_ERR_NULL;
  const HeadOffsets o = head_offsets(s->obs_dim, s->hidden);
  if (o.smem_end > 227u * 1024u) return CRL_ERR_UNSUPPORTED;
  *bytes = (int64_t)o.packed_end;
  return CRL_OK;
}

int crl_encoder_pack_head(const CrlEncoderShape* s, const float* w, const float* b, void* packed, void* stream) {
  const int rc = check_shape(s);
  if (rc) return rc;
  if (!w || !b || !packed) return CRL_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(packed) & 15u) return CRL_ERR_ALIGN;
  if (head_offsets(s->obs_dim, s->hidden).smem_end > 227u * 1024u) return CRL_ERR_UNSUPPORTED;
  HeadPackArgs a{w, b, static_cast<uint8_t*>(packed), s->obs_dim, s->hidden};
  head_pack_kernel<<<148, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? CRL_OK : CRL_ERR_LAUNCH;
}

int crl_encoder_head(const CrlEncoderShape* s, int32_t num_envs, const float* obs, const float* pooled,
                     const void* packed_head, float* out, int32_t* status, void* stream) {
  const int rc = check_shape(s);
  if (rc) return rc;
  if (!obs || !pooled || !packed_head || !out) return CRL_ERR_NULL;
  if (num_envs <= 0) return CRL_ERR_CONFIG;
  if (reinterpret_cast<uintptr_t>(packed_head) & 15u) return CRL_ERR_ALIGN;
  const HeadOffsets o = head_offsets(s->obs_dim, s->hidden);
  if (o.smem_end > 227u * 1024u) return CRL_ERR_UNSUPPORTED;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return CRL_ERR_DEVICE;
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    if (cudaFuncSetAttribute(encoder_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return CRL_ERR_DEVICE;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  HeadArgs a{obs, pooled, static_cast<const uint8_t*>(packed_head), out, status, num_envs, s->obs_dim, s->hidden, 0};
  a.n_tiles = (num_envs + kRows - 1) / kRows;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int want = (a.n_tiles + kGroups - 1) / kGroups;
  const int grid = want < sms ? want : sms;
  encoder_head_kernel<<<grid, kGroupThreads * kGroups, o.smem_end, static_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? CRL_OK : CRL_ERR_LAUNCH;
}

}  // extern "C"

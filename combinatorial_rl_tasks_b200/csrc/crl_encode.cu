// crl_encode.cu -- the consumer side of the observation: ZoneEnvModel's per-zone network and
// mean-pool (main/src/env_model.py:56-78), fused into one sm_100a kernel.
//
//   zone_emb[e] = 1/N * sum_z  L3( relu( L2( relu( L1( [obs[e], zone_obs[e][z]] ))))),   L_i = nn.Linear
//               = L3( pooled[e] ),   pooled[e] = 1/N * sum_z relu( L2( relu( L1( . ))))    (L3 is affine)
//
// The reference materialises three (B*N, h) activations (h = 185 by default,
// main/scripts/train_ppo.py:66).  Here the kernel produces `pooled` (B, h) and nothing else: the two
// wide GEMMs run on the 5th-gen tensor cores (tcgen05.mma, kind::f16 with bf16 operands, fp32
// accumulators in TMEM, M = 128, N = HP = h padded to a multiple of 32, K = 16 per instruction),
// W1 / W2 stay resident in shared memory for the whole (persistent) kernel, the epilogue of layer 1
// (TMEM -> registers -> bias, ReLU, bf16 -> shared memory in the canonical K-major operand layout)
// feeds layer 2, and the epilogue of layer 2 pools each env's zones with a butterfly of warp
// shuffles.  The mean is taken BEFORE the third Linear, which turns L3 from a (B N, h) x (h, h) GEMM
// into a (B, h) x (h, h) one: a third of the tensor work disappears, and L3 / combine_net_
// (env_model.py:79) stay plain fp32 library GEMMs on the caller's side.
//
// A CTA = two independent groups of 256 threads, each walking its own sequence of tiles (8 envs =
// 128 (env, zone) rows; zone slot 15 of an env is padding) with its own accumulator, operand buffers
// and mbarrier: while one group is in an epilogue (CUDA cores) the other group's MMAs occupy the
// tensor pipe.  Within a group, warps w and w + 4 share TMEM lane quadrant w % 4 (rows 32 (w % 4) ..)
// and split the accumulator's columns.  The next tile's inputs are prefetched into registers.
//
// The epilogues are the bottleneck of such a narrow MLP (K = 192 gives the tensor pipe 13 x 96 cycles
// per tile, while 2 x 96 KB of accumulator have to come back through registers), so they are kept
// minimal: the BIASES ARE FOLDED INTO THE GEMMS -- the operand rows carry constant-1 columns (layer 1:
// columns in_dim, in_dim + 1 of the 16; layer 2: columns h, h + 1 of HP, produced by two "generator"
// rows of W1) that multiply the bias split into a bf16 high and low part -- so epilogue 1 is one
// cvt.rn.relu.bf16x2 per pair of values and epilogue 2 one max per value plus the pooling butterfly.
// A padding row is all zeros including its ones, hence stays exactly zero through both layers and
// drops out of the mean by itself.
//
// A separate translation unit on purpose: nothing here can move the register allocation of the
// step kernels (crl_kernels.cu; see profiles/r01_notes.md on how easily that happens).
//
// Operand layout in shared memory (no swizzle, "interleaved" K-major canonical layout of
// cute::UMMA, mma_traits_sm100.hpp: ((8,n),2):((1,SBO),LBO) in 16-byte units): element (row r,
// column k) of an R x K bf16 matrix lives at byte
//     (r % 8) * 16 + (r / 8) * (16 K) + (k / 8) * 128 + (k % 8) * 2
// i.e. 8x8 core matrices of 128 contiguous bytes, consecutive along K (LBO = 128 B), 8-row groups
// 16 K bytes apart (SBO).  One tcgen05.mma consumes K = 16 = two core matrices; k-step s starts
// 256 s bytes into the image.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/crl_b200.h"

namespace crl_enc {

constexpr int kRows = 128;          // rows of a tile = TMEM lanes
constexpr int kGroupThreads = 256;  // 8 warps: two per TMEM lane quadrant
constexpr int kGroups = 2;          // independent groups per CTA
constexpr int kEnvsPerTile = 8;     // 16 row slots per env (N <= 16)
constexpr int kK1 = 16;             // padded input width of layer 1 (obs_dim + zone_dim + a ones column <= 16)
constexpr uint32_t kTmemCols = 512; // one accumulator of up to 256 columns per group; the CTA owns the SM
constexpr uint32_t kSpinLimit = 1u << 24;

// accumulator width: h plus at least one spare column (the ones column of layer 2), multiple of 32
__host__ __device__ inline int padded_hidden(int h) { return (h + 1 + 31) & ~31; }

// byte offset of element (r, k) in the canonical image of a matrix with K columns
__host__ __device__ inline uint32_t canon(int r, int k, int K) {
  return (uint32_t)((r & 7) * 16 + (r >> 3) * (16 * K) + (k >> 3) * 128 + (k & 7) * 2);
}

struct Offsets {   // byte offsets: packed weight buffer == start of shared memory; then per-group buffers
  uint32_t w2, w1, packed_end, group0, abuf, a1buf, bar, group_bytes, tmem_slot, smem_end;
};
__host__ __device__ inline Offsets offsets(int HP) {
  Offsets o;
  o.w2 = 0;
  o.w1 = o.w2 + (uint32_t)HP * HP * 2;
  o.packed_end = o.w1 + (uint32_t)HP * kK1 * 2;
  o.group0 = (o.packed_end + 127u) & ~127u;
  o.abuf = 0;                                            // relative to the group's base
  o.a1buf = o.abuf + (uint32_t)kRows * HP * 2;
  o.bar = o.a1buf + (uint32_t)kRows * kK1 * 2;
  o.group_bytes = (o.bar + 8 + 127u) & ~127u;
  o.tmem_slot = o.group0 + kGroups * o.group_bytes;
  o.smem_end = o.tmem_slot + 16;
  return o;
}

// ---- packing: torch-layout fp32 weights -> the shared-memory image -------------------------
struct PackArgs {
  const float *w1, *b1, *w2, *b2;
  uint8_t* out;
  int in_dim, h, HP;
};

__global__ void pack_kernel(const PackArgs a) {
  const Offsets o = offsets(a.HP);
  const int HP = a.HP;
  const int n_w = HP * HP, n_w1 = HP * kK1;
  const int total = n_w + n_w1;
  // bias = hi + lo with hi = bf16(bias), lo = bf16(bias - hi): column `ones` of the operand carries 1
  // and multiplies hi, column `ones + 1` (when the matrix has one) multiplies lo
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const bool second = i < n_w;
    const int K = second ? HP : kK1, ones = second ? a.h : a.in_dim;
    const int j = second ? i : i - n_w, n = j / K, k = j % K;
    const float* w = second ? a.w2 : a.w1;
    const float* b = second ? a.b2 : a.b1;
    float v = 0.f;
    if (n < a.h) {
      if (k < ones) v = w[(size_t)n * ones + k];
      else if (k == ones) v = __bfloat162float(__float2bfloat16_rn(b[n]));
      else if (k == ones + 1) v = b[n] - __bfloat162float(__float2bfloat16_rn(b[n]));
    } else if (!second && (n == a.h || n == a.h + 1) && k == ones) {
      v = 1.f;                                   // generator rows of W1: layer 2's ones columns h, h + 1
    }
    *reinterpret_cast<__nv_bfloat16*>(a.out + (second ? o.w2 : o.w1) + canon(n, k, K)) = __float2bfloat16_rn(v);
  }
}

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: start address, LBO = 128 B, SBO, version 1 (sm_100), no swizzle
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) |
         (1ull << 46);
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
  int B, N, Z, obs_dim, h, HP, n_tiles;
};

// eight consecutive values (k = 8 half .. 8 half + 7) of row (e, slot) of the layer-1 operand
// [obs[e], zone_obs[e][slot], 1, 1, 0...]; all zeros -- the ones included -- for a padding row
__device__ __forceinline__ void load_half_row(const EncArgs& a, int tile, int m, int half, float (&x)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = 0.f;
  const int e = tile * kEnvsPerTile + (m >> 4), slot = m & 15;
  if (tile < a.n_tiles && e < a.B && slot < a.N) {
    const float* ob = a.obs + (size_t)e * a.obs_dim;
    const float* zo = a.zone_obs + ((size_t)e * a.N + slot) * a.Z;
    const int in_dim = a.obs_dim + a.Z;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = 8 * half + j;
      if (k < a.obs_dim) x[j] = __ldg(ob + k);
      else if (k < in_dim) x[j] = __ldg(zo + (k - a.obs_dim));
      else if (k <= in_dim + 1) x[j] = 1.f;
    }
  }
}

// relu(acc) of 32 columns of this thread's row -> bf16 -> the next layer's A operand
__device__ __forceinline__ void relu_to_smem(const uint32_t (&v)[32], uint8_t* abuf, uint32_t row_off, int c) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    *reinterpret_cast<uint4*>(abuf + row_off + (uint32_t)((c * 4 + q) * 128)) =
        make_uint4(relu_pack_bf16(v[q * 8 + 0], v[q * 8 + 1]), relu_pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                   relu_pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), relu_pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
  }
}

// relu(acc) of 32 columns, summed over the 16 rows (lanes) of an env.  A butterfly in which every
// step halves what a lane carries: with partner lane ^ off, the lane whose bit `off` is clear keeps
// the lower half of its values and receives the partner's lower half, the other lane the upper
// halves: 16 + 8 + 4 + 2 = 30 shuffles instead of 32 x 4.  Lane L of the half-warp ends up with the
// sums of columns 2L and 2L + 1 of the chunk, which it stores scaled by 1/N.
__device__ __forceinline__ void relu_pool_store(const uint32_t (&v)[32], int lane, float inv_n, float* dst_row, int c, int h) {
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = fmaxf(__uint_as_float(v[j]), 0.f);
#define CRL_BUTTERFLY(HALF, OFF)                                        \
  {                                                                     \
    const bool upper = (lane & OFF) != 0;                               \
    _Pragma("unroll") for (int i = 0; i < HALF; ++i) {                  \
      const float send = upper ? x[i] : x[i + HALF];                    \
      const float keep = upper ? x[i + HALF] : x[i];                    \
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);            \
    }                                                                   \
  }
  CRL_BUTTERFLY(16, 8)
  CRL_BUTTERFLY(8, 4)
  CRL_BUTTERFLY(4, 2)
  CRL_BUTTERFLY(2, 1)
#undef CRL_BUTTERFLY
  if (dst_row) {
    const int col = c * 32 + 2 * (lane & 15);
    if (col < h) dst_row[col] = x[0] * inv_n;
    if (col + 1 < h) dst_row[col + 1] = x[1] * inv_n;
  }
}

__global__ void __launch_bounds__(kGroupThreads * kGroups, 1) zone_encode_kernel(const EncArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const Offsets o = offsets(a.HP);
  const int group = threadIdx.x / kGroupThreads;              // 0 / 1
  const int t = threadIdx.x % kGroupThreads;
  const int m = t & (kRows - 1);                              // row of the tile = TMEM lane
  const int half = t >> 7;                                    // which half of the columns / of the input row
  const int quad = (t >> 5) & 3, lane = t & 31;               // TMEM lane quadrant of this warp
  const int HP = a.HP;
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
  const uint32_t acc = tmem_base + (uint32_t)group * 256u;    // this group's accumulator columns
  const uint32_t my_acc = acc + ((uint32_t)(quad * 32) << 16);
  // instruction descriptor: D fp32, A and B bf16, both K-major, N = HP, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HP >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
  uint8_t* abuf = gbase + o.abuf;
  uint8_t* a1buf = gbase + o.a1buf;
  const uint32_t abuf_addr = smem_u32(abuf), a1_addr = smem_u32(a1buf);
  const uint32_t w1_addr = smem_u32(smem + o.w1), w2_addr = smem_u32(smem + o.w2);
  const uint32_t a1_off = (uint32_t)((m & 7) * 16 + (m >> 3) * (16 * kK1) + half * 128);
  const uint32_t a_off = (uint32_t)((m & 7) * 16 + (m >> 3) * (16 * HP));
  const int n_chunks = HP / 32;
  const int c_split = (n_chunks + 1) / 2;
  const int c_begin = half ? c_split : 0, c_end = half ? n_chunks : c_split;   // this warp's column chunks
  const float inv_n = 1.0f / (float)a.N;
  uint32_t parity = 0u;
  bool healthy = true;

  const int tile_stride = kGroups * gridDim.x;
  int tile = kGroups * blockIdx.x + group;
  float x[8];
  load_half_row(a, tile, m, half, x);
  for (; tile < a.n_tiles; tile += tile_stride) {
    const int e = tile * kEnvsPerTile + (m >> 4);
    // ---- layer-1 operand from the prefetched registers ---------------------------------------
    *reinterpret_cast<uint4*>(a1buf + a1_off) =
        make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
    fence_async_smem();
    tc_fence_before();
    group_sync(group);
    // ---- layer 1: [128 x 16] x [16 x HP] -> acc ---------------------------------------------
    if (t == 0) {
      tc_fence_after();
      mma_bf16(acc, smem_desc(a1_addr, 16 * kK1), smem_desc(w1_addr, 16 * kK1), idesc, 0u);
      mma_commit(bar_addr);
    }
    load_half_row(a, tile + tile_stride, m, half, x);         // prefetch: in flight for the rest of the tile
    healthy = mbar_wait(bar_addr, parity) && healthy;
    parity ^= 1u;
    tc_fence_after();
    for (int c = c_begin; c < c_end; ++c) {
      uint32_t v[32];
      tmem_ld32(my_acc + (uint32_t)(c * 32), v);
      tmem_ld_wait(v);
      relu_to_smem(v, abuf, a_off, c);
    }
    fence_async_smem();
    tc_fence_before();
    group_sync(group);
    // ---- layer 2: [128 x HP] x [HP x HP] -> acc (layer 1's values have been read) -------------
    if (t == 0) {
      tc_fence_after();
      for (int s = 0; s < HP / 16; ++s)
        mma_bf16(acc, smem_desc(abuf_addr + 256u * s, 16 * HP), smem_desc(w2_addr + 256u * s, 16 * HP), idesc, s > 0);
      mma_commit(bar_addr);
    }
    healthy = mbar_wait(bar_addr, parity) && healthy;
    parity ^= 1u;
    tc_fence_after();
    // ---- ReLU, mean over the env's zones, store ----------------------------------------------
    float* dst_row = e < a.B ? a.out + (size_t)e * a.h : nullptr;
    for (int c = c_begin; c < c_end; ++c) {
      uint32_t v[32];
      tmem_ld32(my_acc + (uint32_t)(c * 32), v);
      tmem_ld_wait(v);
      relu_pool_store(v, lane, inv_n, dst_row, c, a.h);
    }
    // the next tile's layer-1 MMA overwrites acc: ordered after these loads by the fence and the
    // group barrier that precede it
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
  if (s->obs_dim <= 0 || s->zone_dim <= 0 || s->obs_dim + s->zone_dim >= kK1) return CRL_ERR_CONFIG;   // + a ones column
  if (s->num_zones <= 0 || s->num_zones > 16) return CRL_ERR_CONFIG;
  if (s->hidden <= 0) return CRL_ERR_CONFIG;
  if (padded_hidden(s->hidden) > 192) return CRL_ERR_UNSUPPORTED;   // resident W2 + two operand buffers must fit an SM
  return CRL_OK;
}

}  // namespace crl_enc

using namespace crl_enc;

extern "C" {

int crl_encoder_packed_bytes(const CrlEncoderShape* s, int64_t* bytes) {
  const int rc = check_shape(s);
  if (rc) return rc;
  if (!bytes) return CRL_ERR_NULL;
  *bytes = (int64_t)offsets(padded_hidden(s->hidden)).packed_end;
  return CRL_OK;
}

int crl_encoder_pack(const CrlEncoderShape* s, const float* w1, const float* b1, const float* w2, const float* b2,
                     void* packed, void* stream) {
  const int rc = check_shape(s);
  if (rc) return rc;
  if (!w1 || !b1 || !w2 || !b2 || !packed) return CRL_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(packed) & 15u) return CRL_ERR_ALIGN;
  PackArgs a{w1, b1, w2, b2, static_cast<uint8_t*>(packed), s->obs_dim + s->zone_dim, s->hidden, padded_hidden(s->hidden)};
  pack_kernel<<<148, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? CRL_OK : CRL_ERR_LAUNCH;
}

int crl_zone_encode(const CrlEncoderShape* s, int32_t num_envs, const float* obs, const float* zone_obs,
                    const void* packed, float* pooled, int32_t* status, void* stream) {
  const int rc = check_shape(s);
  if (rc) return rc;
  if (!obs || !zone_obs || !packed || !pooled) return CRL_ERR_NULL;
  if (num_envs <= 0) return CRL_ERR_CONFIG;
  if (reinterpret_cast<uintptr_t>(packed) & 15u) return CRL_ERR_ALIGN;
  const int HP = padded_hidden(s->hidden);
  const Offsets o = offsets(HP);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(zone_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return CRL_ERR_DEVICE;
    attr_set = true;
  }
  EncArgs a{obs, zone_obs, static_cast<const uint8_t*>(packed), pooled, status, num_envs, s->num_zones, s->zone_dim,
            s->obs_dim, s->hidden, HP, (num_envs + kEnvsPerTile - 1) / kEnvsPerTile};
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int want = (a.n_tiles + kGroups - 1) / kGroups;
  const int grid = want < sms ? want : sms;                   // persistent: one CTA per SM, weights loaded once
  zone_encode_kernel<<<grid, kGroupThreads * kGroups, o.smem_end, static_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? CRL_OK : CRL_ERR_LAUNCH;
}

}  // extern "C"

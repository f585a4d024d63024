// crl_encode.cu -- the consumer side of the observation: ZoneEnvModel's per-zone network and
// mean-pool (main/src/env_model.py:56-78), fused into one sm_100a kernel.
//
//   zone_emb[e] = 1/N * sum_z  L3( relu( L2( relu( L1( [obs[e], zone_obs[e][z]] ))))),   L_i = nn.Linear
//
// The reference materialises three (B*N, h) activations (h = 185 by default,
// main/scripts/train_ppo.py:66); here they never leave the SM: a CTA owns a tile of 8 envs =
// 128 (env, zone) rows (zone 15 of each env is padding), the three GEMMs run on the 5th-gen tensor
// cores (tcgen05.mma, kind::f16 with bf16 operands, fp32 accumulators in TMEM, M = 128,
// N = HP = h padded to a multiple of 32, K = 16 per instruction), the weights stay resident in
// shared memory for the whole (persistent) kernel, and each epilogue (TMEM -> registers -> bias,
// ReLU, bf16 -> shared memory in the canonical K-major operand layout) feeds the next GEMM.
// Only obs / zone_obs are read and (B, h) written.  combine_net_ (one plain Linear on
// [obs, zone_emb], env_model.py:79) stays a library GEMM on the caller's side.
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

constexpr int kRows = 128;          // rows of a tile = TMEM lanes = threads of the CTA
constexpr int kEnvsPerTile = 8;     // 16 row slots per env (N <= 16)
constexpr int kK1 = 16;             // padded input width of layer 1 (obs_dim + zone_dim <= 16)
constexpr uint32_t kTmemCols = 512; // two accumulators of up to 192 columns; the CTA owns the SM
constexpr uint32_t kSpinLimit = 1u << 24;

__host__ __device__ inline int padded_hidden(int h) { return (h + 31) & ~31; }

// byte offset of element (r, k) in the canonical image of a matrix with K columns
__host__ __device__ inline uint32_t canon(int r, int k, int K) {
  return (uint32_t)((r & 7) * 16 + (r >> 3) * (16 * K) + (k >> 3) * 128 + (k & 7) * 2);
}

struct Offsets {   // byte offsets inside the packed weight buffer == inside shared memory
  uint32_t w2, w3, w1, b1, b2, b3, packed_end, abuf, a1buf, bar, tmem_slot, smem_end;
};
__host__ __device__ inline Offsets offsets(int HP) {
  Offsets o;
  o.w2 = 0;
  o.w3 = o.w2 + (uint32_t)HP * HP * 2;
  o.w1 = o.w3 + (uint32_t)HP * HP * 2;
  o.b1 = o.w1 + (uint32_t)HP * kK1 * 2;
  o.b2 = o.b1 + (uint32_t)HP * 4;
  o.b3 = o.b2 + (uint32_t)HP * 4;
  o.packed_end = o.b3 + (uint32_t)HP * 4;
  o.abuf = (o.packed_end + 127u) & ~127u;
  o.a1buf = o.abuf + (uint32_t)kRows * HP * 2;
  o.bar = o.a1buf + (uint32_t)kRows * kK1 * 2;
  o.tmem_slot = o.bar + 8;
  o.smem_end = o.tmem_slot + 8;
  return o;
}

// ---- packing: torch-layout fp32 weights -> the shared-memory image -------------------------
struct PackArgs {
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  uint8_t* out;
  int in_dim, h, HP;
};

__global__ void pack_kernel(const PackArgs a) {
  const Offsets o = offsets(a.HP);
  const int HP = a.HP;
  const int n_w = HP * HP, n_w1 = HP * kK1;
  const int total = 2 * n_w + n_w1 + 3 * HP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < 2 * n_w) {
      const int which = i / n_w, j = i % n_w, n = j / HP, k = j % HP;
      const float* w = which ? a.w3 : a.w2;
      const float v = (n < a.h && k < a.h) ? w[(size_t)n * a.h + k] : 0.f;
      *reinterpret_cast<__nv_bfloat16*>(a.out + (which ? o.w3 : o.w2) + canon(n, k, HP)) = __float2bfloat16_rn(v);
    } else if (i < 2 * n_w + n_w1) {
      const int j = i - 2 * n_w, n = j / kK1, k = j % kK1;
      const float v = (n < a.h && k < a.in_dim) ? a.w1[(size_t)n * a.in_dim + k] : 0.f;
      *reinterpret_cast<__nv_bfloat16*>(a.out + o.w1 + canon(n, k, kK1)) = __float2bfloat16_rn(v);
    } else {
      const int j = i - 2 * n_w - n_w1, which = j / HP, n = j % HP;
      const float* b = which == 0 ? a.b1 : (which == 1 ? a.b2 : a.b3);
      reinterpret_cast<float*>(a.out + o.b1)[j] = n < a.h ? b[n] : 0.f;
    }
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
// this warp's 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);   // .x (low half) = lo
  return *reinterpret_cast<const uint32_t*>(&t);
}

struct EncArgs {
  const float* obs;        // [B][obs_dim]
  const float* zone_obs;   // [B][N][Z]
  const uint8_t* packed;
  float* out;              // [B][h]
  int* status;             // device int: set to 1 if a tensor-core wait expired
  int B, N, Z, obs_dim, h, HP, n_tiles;
};

// bias + ReLU + bf16 of one accumulator (this thread's row), written as the next layer's A operand
__device__ __forceinline__ void epilogue_to_smem(uint32_t tmem_acc, const float* bias, uint8_t* abuf, int m, int HP) {
  const uint32_t row_off = (uint32_t)((m & 7) * 16 + (m >> 3) * (16 * HP));
  for (int c = 0; c < HP / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem_acc + (uint32_t)(c * 32), v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = c * 32 + q * 8 + 2 * j;
        const float x0 = fmaxf(__uint_as_float(v[q * 8 + 2 * j]) + bias[col], 0.f);
        const float x1 = fmaxf(__uint_as_float(v[q * 8 + 2 * j + 1]) + bias[col + 1], 0.f);
        w[j] = pack_bf16(x0, x1);
      }
      *reinterpret_cast<uint4*>(abuf + row_off + (uint32_t)((c * 4 + q) * 128)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

__global__ void __launch_bounds__(kRows, 1) zone_encode_kernel(const EncArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const Offsets o = offsets(a.HP);
  const int m = threadIdx.x, warp = m >> 5, lane = m & 31;
  const int HP = a.HP;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + o.bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + o.tmem_slot);
  const uint32_t bar_addr = smem_u32(bar);

  // ---- one-time setup: barrier, TMEM, resident weights -----------------------------------
  if (m == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_addr) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.packed);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = m; i < o.packed_end / 16; i += kRows) dst[i] = __ldg(src + i);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc0 = tmem_base, acc1 = tmem_base + 256u;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;      // this warp's TMEM lanes
  // instruction descriptor: D fp32, A and B bf16, both K-major, N = HP, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HP >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
  const float* b1 = reinterpret_cast<const float*>(smem + o.b1);
  const float* b2 = reinterpret_cast<const float*>(smem + o.b2);
  const float* b3 = reinterpret_cast<const float*>(smem + o.b3);
  uint8_t* abuf = smem + o.abuf;
  uint8_t* a1buf = smem + o.a1buf;
  const uint32_t abuf_addr = smem_u32(abuf), a1_addr = smem_u32(a1buf);
  const uint32_t w1_addr = smem_u32(smem + o.w1), w2_addr = smem_u32(smem + o.w2), w3_addr = smem_u32(smem + o.w3);
  uint32_t parity = 0u;
  bool healthy = true;
  const int slot = m & 15;                                    // zone slot of this row
  const float inv_n = 1.0f / (float)a.N;

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int e = tile * kEnvsPerTile + (m >> 4);
    const bool live = e < a.B && slot < a.N;
    // ---- layer-1 operand: [obs[e], zone_obs[e][slot], 0...] as 16 bf16 ----------------------
    {
      float x[kK1];
#pragma unroll
      for (int k = 0; k < kK1; ++k) x[k] = 0.f;
      if (live) {
        const float* ob = a.obs + (size_t)e * a.obs_dim;
        const float* zo = a.zone_obs + ((size_t)e * a.N + slot) * a.Z;
#pragma unroll
        for (int k = 0; k < kK1; ++k) {
          if (k < a.obs_dim) x[k] = __ldg(ob + k);
          else if (k - a.obs_dim < a.Z) x[k] = __ldg(zo + (k - a.obs_dim));
        }
      }
      const uint32_t off = (uint32_t)((m & 7) * 16 + (m >> 3) * (16 * kK1));
      *reinterpret_cast<uint4*>(a1buf + off) =
          make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
      *reinterpret_cast<uint4*>(a1buf + off + 128) =
          make_uint4(pack_bf16(x[8], x[9]), pack_bf16(x[10], x[11]), pack_bf16(x[12], x[13]), pack_bf16(x[14], x[15]));
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- layer 1: [128 x 16] x [16 x HP] -> acc0 --------------------------------------------
    if (m == 0) {
      tc_fence_after();
      mma_bf16(acc0, smem_desc(a1_addr, 16 * kK1), smem_desc(w1_addr, 16 * kK1), idesc, 0u);
      mma_commit(bar_addr);
    }
    healthy = mbar_wait(bar_addr, parity) && healthy;
    parity ^= 1u;
    tc_fence_after();
    epilogue_to_smem(acc0 + lane_sel, b1, abuf, m, HP);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- layer 2: [128 x HP] x [HP x HP] -> acc1 --------------------------------------------
    if (m == 0) {
      tc_fence_after();
      for (int s = 0; s < HP / 16; ++s)
        mma_bf16(acc1, smem_desc(abuf_addr + 256u * s, 16 * HP), smem_desc(w2_addr + 256u * s, 16 * HP), idesc, s > 0);
      mma_commit(bar_addr);
    }
    healthy = mbar_wait(bar_addr, parity) && healthy;
    parity ^= 1u;
    tc_fence_after();
    epilogue_to_smem(acc1 + lane_sel, b2, abuf, m, HP);       // layer 2 has finished reading abuf
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- layer 3: [128 x HP] x [HP x HP] -> acc0 --------------------------------------------
    if (m == 0) {
      tc_fence_after();
      for (int s = 0; s < HP / 16; ++s)
        mma_bf16(acc0, smem_desc(abuf_addr + 256u * s, 16 * HP), smem_desc(w3_addr + 256u * s, 16 * HP), idesc, s > 0);
      mma_commit(bar_addr);
    }
    healthy = mbar_wait(bar_addr, parity) && healthy;
    parity ^= 1u;
    tc_fence_after();
    // ---- bias, mean over the env's zones (16 consecutive lanes), store ------------------------
    for (int c = 0; c < HP / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(acc0 + lane_sel + (uint32_t)(c * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = live ? __uint_as_float(v[j]) + b3[c * 32 + j] : 0.f;
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 4);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        v[j] = __float_as_uint(x * inv_n);
      }
      // lane j of each half-warp writes column j, j + 16 of its env: two coalesced 64-byte rows
      if (e < a.B) {
        float* dst = a.out + (size_t)e * a.h + c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if ((lane & 15) == (j & 15) && c * 32 + j < a.h) dst[j] = __uint_as_float(v[j]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                          // acc0 / a1buf are free for the next tile
  }
  if (!healthy && a.status) atomicExch(a.status, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

static int check_shape(const CrlEncoderShape* s) {
  if (!s) return CRL_ERR_NULL;
  if (s->obs_dim <= 0 || s->zone_dim <= 0 || s->obs_dim + s->zone_dim > kK1) return CRL_ERR_CONFIG;
  if (s->num_zones <= 0 || s->num_zones > 16) return CRL_ERR_CONFIG;
  if (s->hidden <= 0) return CRL_ERR_CONFIG;
  if (padded_hidden(s->hidden) > 192) return CRL_ERR_UNSUPPORTED;   // two resident HP x HP bf16 weights must fit an SM
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
                     const float* w3, const float* b3, void* packed, void* stream) {
  const int rc = check_shape(s);
  if (rc) return rc;
  if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !packed) return CRL_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(packed) & 15u) return CRL_ERR_ALIGN;
  PackArgs a{w1, b1, w2, b2, w3, b3, static_cast<uint8_t*>(packed), s->obs_dim + s->zone_dim, s->hidden,
             padded_hidden(s->hidden)};
  pack_kernel<<<148, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? CRL_OK : CRL_ERR_LAUNCH;
}

int crl_zone_encode(const CrlEncoderShape* s, int32_t num_envs, const float* obs, const float* zone_obs,
                    const void* packed, float* zone_emb, int32_t* status, void* stream) {
  const int rc = check_shape(s);
  if (rc) return rc;
  if (!obs || !zone_obs || !packed || !zone_emb) return CRL_ERR_NULL;
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
  EncArgs a{obs, zone_obs, static_cast<const uint8_t*>(packed), zone_emb, status, num_envs, s->num_zones, s->zone_dim,
            s->obs_dim, s->hidden, HP, (num_envs + kEnvsPerTile - 1) / kEnvsPerTile};
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.n_tiles < sms ? a.n_tiles : sms;         // persistent: one CTA per SM, weights loaded once
  zone_encode_kernel<<<grid, kRows, o.smem_end, static_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? CRL_OK : CRL_ERR_LAUNCH;
}

}  // extern "C"

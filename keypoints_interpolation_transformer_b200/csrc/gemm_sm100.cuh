// tcgen05 / TMEM / TMA GEMM for sm_100a.  One 128 x BN fp32 accumulator tile per CTA lives in
// tensor memory; operands are staged by TMA into 128B-swizzled shared memory through a
// STAGES-deep mbarrier ring; one elected thread issues tcgen05.mma; four epilogue warps read the
// accumulator back with tcgen05.ld and apply the fused epilogue.
//
//   MODE 0 (TN)    C[M,N] = A[M,K] * B[N,K]^T      both operands K-major      (y = x W^T, dx = dy W)
//   MODE 1 (wgrad) C[M,N] = A[K,M]^T * B[K,N]      both operands MN-major     (dW = dy^T x)
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp % 4).
#pragma once
#include "common.cuh"

namespace kit {

enum { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_ATOMIC = 2 };
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_GELU_BWD = 2 };

struct GemmParams {
  int M, N, K;  // C is [M,N]; K is the reduction extent
  void* C;
  int64_t ldc;
  const float* bias;  // [N] or null
  const bf16* addend; // [M, ld_addend] or null
  int64_t ld_addend;
  bf16* aux;  // ACT_GELU: pre-activation out; ACT_GELU_BWD: pre-activation in
  int64_t ld_aux;
  int out_kind, act;
  int kb_per_split;  // 64-wide k-blocks handled by one blockIdx.z
};

struct GemmPlan {
  CUtensorMap tmA, tmB;
  GemmParams p;
  int mode;
  dim3 grid;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

template <int BN, int STAGES>
constexpr int gemm_smem_bytes() {
  return STAGES * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + 1024 /*align slack*/ + 256 /*barriers*/;
}

template <int BN, int MODE, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                      const __grid_constant__ CUtensorMap tmB,
                                                                      const GemmParams p) {
  constexpr int BM = GEMM_BM, BK = GEMM_BK;
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN must be a power-of-two TMEM allocation");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
  const int num_kb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        uint8_t* sA = smem + s * STAGE_BYTES;
        uint8_t* sB = sA + A_BYTES;
        const int kc = (kb_begin + i) * BK;
        if (MODE == 0) {
          tma_load_2d(sA, &tmA, &full[s], kc, m0);
          tma_load_2d(sB, &tmB, &full[s], kc, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + j * (BK * 128), &tmA, &full[s], m0 + 64 * j, kc);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * (BK * 128), &tmB, &full[s], n0 + 64 * j, kc);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, MODE == 1, MODE == 1);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          uint64_t adesc, bdesc;
          if (MODE == 0) {  // K-major, one 128B swizzle atom along K: +32 B per UMMA_K
            adesc = make_smem_desc_sw128(a_base + k * 32, 0, 1024);
            bdesc = make_smem_desc_sw128(b_base + k * 32, 0, 1024);
          } else {  // MN-major: 64-element MN atoms LBO apart, 8-row k groups SBO apart
            adesc = make_smem_desc_sw128(a_base + k * 16 * 128, BK * 128, 1024);
            bdesc = make_smem_desc_sw128(b_base + k * 16 * 128, BK * 128, 1024);
          }
          umma_bf16(tmem_base, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const bool vec_ok = ((p.ldc & 7) == 0) && ((p.N & 7) == 0) && (p.addend == nullptr || (p.ld_addend & 7) == 0) &&
                        (p.aux == nullptr || (p.ld_aux & 7) == 0);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c * 32), r);
      tmem_ld_wait();
      const int col0 = n0 + c * 32;
      if (!row_ok || col0 >= p.N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const bool full_chunk = (col0 + 32 <= p.N) && vec_ok;
      if (full_chunk) {
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(p.bias + col0 + j);
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
        }
        if (p.addend != nullptr) {
          const bf16* ap = p.addend + (int64_t)row * p.ld_addend + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float t[8];
            load8(ap + j, t);
#pragma unroll
            for (int u = 0; u < 8; ++u) v[j + u] += t[u];
          }
        }
        if (p.act == ACT_GELU) {
          bf16* xp = p.aux + (int64_t)row * p.ld_aux + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            store8(xp + j, v + j);
#pragma unroll
            for (int u = 0; u < 8; ++u) v[j + u] = gelu_erf(v[j + u]);
          }
        } else if (p.act == ACT_GELU_BWD) {
          const bf16* xp = p.aux + (int64_t)row * p.ld_aux + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float t[8];
            load8(xp + j, t);
#pragma unroll
            for (int u = 0; u < 8; ++u) v[j + u] *= gelu_erf_grad(t[u]);
          }
        }
        if (p.out_kind == OUT_BF16) {
          bf16* cp = reinterpret_cast<bf16*>(p.C) + (int64_t)row * p.ldc + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) store8(cp + j, v + j);
        } else if (p.out_kind == OUT_F32) {
          float* cp = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          float* cp = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + j), "f"(v[j]), "f"(v[j + 1]),
                         "f"(v[j + 2]), "f"(v[j + 3])
                         : "memory");
          }
        }
      } else {  // ragged N edge or unaligned leading dimension: scalar path
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          if (col >= p.N) continue;
          float x = v[j];
          if (p.bias != nullptr) x += p.bias[col];
          if (p.addend != nullptr) x += __bfloat162float(p.addend[(int64_t)row * p.ld_addend + col]);
          if (p.act == ACT_GELU) {
            p.aux[(int64_t)row * p.ld_aux + col] = __float2bfloat16(x);
            x = gelu_erf(x);
          } else if (p.act == ACT_GELU_BWD) {
            x *= gelu_erf_grad(__bfloat162float(p.aux[(int64_t)row * p.ld_aux + col]));
          }
          if (p.out_kind == OUT_BF16) {
            reinterpret_cast<bf16*>(p.C)[(int64_t)row * p.ldc + col] = __float2bfloat16(x);
          } else if (p.out_kind == OUT_F32) {
            reinterpret_cast<float*>(p.C)[(int64_t)row * p.ldc + col] = x;
          } else {
            atomicAdd(reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col, x);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

// ---------------------------------------------------------------- host side
int make_tensor_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_pitch_bytes,
                       uint32_t box_inner, uint32_t box_outer);
// Builds a plan: tensor maps + grid.  split_k <= 0 lets the planner choose (wgrad only).
int gemm_plan(GemmPlan* plan, int mode, const bf16* A, int64_t lda, const bf16* B, int64_t ldb, void* C, int64_t ldc,
              int M, int N, int K, const float* bias, const bf16* addend, int64_t ld_addend, int out_kind, int act,
              bf16* aux, int64_t ld_aux, int split_k);
int gemm_launch(const GemmPlan* plan, cudaStream_t stream);
int gemm_init_attributes();

}  // namespace kit

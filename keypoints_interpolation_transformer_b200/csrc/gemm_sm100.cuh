// tcgen05 / TMEM / TMA GEMM for sm_100a -- persistent, warp-specialised.
//
// Each CTA (one per SM) loops over 128 x BN output tiles.  Operands are staged by TMA into
// 128B-swizzled shared memory through a STAGES-deep mbarrier ring that runs ahead across tile
// boundaries; one elected thread issues tcgen05.mma into one of TWO fp32 accumulators in tensor
// memory, so the epilogue of tile i (tcgen05.ld -> fused epilogue -> global stores, 8 warps)
// overlaps the TMA + MMA main loop of tile i+1.
//
//   MODE 0 (TN)    C[M,N] = A[M,K] * B[N,K]^T      both operands K-major      (y = x W^T, dx = dy W)
//   MODE 1 (wgrad) C[M,N] = A[K,M]^T * B[K,N]      both operands MN-major     (dW = dy^T x), split-K
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2.. = epilogue
// (BN/16 of them: TMEM lane quadrant = warp % 4, 64-column group = (warp - 2) / 4).  The epilogue is
// ALU / latency bound (GELU, casts), hence the many warps.
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
  int kb_per_split;  // 64-wide k-blocks per work item
  int tiles_m, tiles_n, splits;
  int tma_store;  // 1: outputs leave through shared memory + TMA store / reduce-add (clipped at the edges)
  int dbg;        // experiments only (KIT_GEMM_DBG): 1 = no TMA stores, 2 = no wait on staging reuse, 4 = no tmem loads
  int tma_in;     // 1: the addend / GELU' pre-activation tile arrives through TMA (tmAux) instead of per-row loads
};

struct GemmPlan {
  CUtensorMap tmA, tmB, tmC, tmAux;
  GemmParams p;
  int mode, bn;
  int grid;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_MAX_EPI_WARPS = 16;
constexpr int GEMM_EPI_BUF = 4096;  // per epilogue warp: 2 KB out tile (bf16, 64B-swizzled) + 2 KB second tile
                                    // (GELU pre-activation out / TMA-staged input), or one 4 KB fp32 out tile

template <int BN>
constexpr int gemm_epi_warps() { return BN / 16; }
template <int BN>
constexpr int gemm_threads() { return 64 + 32 * gemm_epi_warps<BN>(); }

template <int BN, int STAGES>
constexpr int gemm_smem_bytes() {
  return STAGES * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + gemm_epi_warps<BN>() * GEMM_EPI_BUF + 1024 /*align slack*/ +
         256 /*barriers*/;
}

// bias / residual / activation on 32 accumulator columns of one row, in registers (packed fp32 pairs).  Columns >= N
// of a ragged last chunk are computed on garbage and clipped by the TMA store.
__device__ __forceinline__ void add_pair(float& a, float& b, float x, float y) { up2(add2(pk2(a, b), pk2(x, y)), a, b); }
__device__ __forceinline__ void gemm_epilogue_math(const GemmParams& p, int row, int col0, bool row_ok, const float* in_vals,
                                                   float (&v)[32]) {
  const bool full = col0 + 32 <= p.N;
  if (p.bias != nullptr) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        add_pair(v[j], v[j + 1], b4.x, b4.y);
        add_pair(v[j + 2], v[j + 3], b4.z, b4.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
    }
  }
  if (in_vals != nullptr) {   // tile staged by TMA (zero-filled outside the tensor)
    if (p.act == ACT_GELU_BWD) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) gelu_grad_mul_pair(in_vals[j], in_vals[j + 1], v[j], v[j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 2) add_pair(v[j], v[j + 1], in_vals[j], in_vals[j + 1]);
    }
    return;
  }
  if (p.addend != nullptr && row_ok) {
    const bf16* ap = p.addend + (int64_t)row * p.ld_addend + col0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float t[8];
        load8(ap + j, t);
#pragma unroll
        for (int u = 0; u < 8; u += 2) add_pair(v[j + u], v[j + u + 1], t[u], t[u + 1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) v[j] += __bfloat162float(ap[j]);
    }
  }
  if (p.act == ACT_GELU_BWD && row_ok) {
    const bf16* xp = p.aux + (int64_t)row * p.ld_aux + col0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float t[8];
        load8(xp + j, t);
#pragma unroll
        for (int u = 0; u < 8; u += 2) gelu_grad_mul_pair(t[u], t[u + 1], v[j + u], v[j + u + 1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) v[j] *= gelu_act_grad(__bfloat162float(xp[j]));
    }
  }
}
// 32 bf16 values of this lane's row into a [32 rows x 64 B] tile, 64B-swizzled (conflict-free, TMA SWIZZLE_64B)
__device__ __forceinline__ void stage_bf16(uint8_t* buf, int lane, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<uint4*>(buf + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = pack8(&v[8 * i]);
}
__device__ __forceinline__ void unstage_bf16(const uint8_t* buf, int lane, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 u = *reinterpret_cast<const uint4*>(buf + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4));
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    v[8 * i] = a.x; v[8 * i + 1] = a.y; v[8 * i + 2] = b.x; v[8 * i + 3] = b.y;
    v[8 * i + 4] = c.x; v[8 * i + 5] = c.y; v[8 * i + 6] = d.x; v[8 * i + 7] = d.y;
  }
}
// 32 fp32 values into a [32 rows x 128 B] tile, 128B-swizzled
__device__ __forceinline__ void stage_f32(uint8_t* ebuf, int lane, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(ebuf + lane * 128 + ((i ^ (lane & 7)) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// One row x 32 consecutive columns of the accumulator through the fused epilogue.
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, int row, int col0, float (&v)[32], bool vec_ok) {
  const bool full_chunk = (col0 + 32 <= p.N) && vec_ok;
  if (full_chunk) {
    if (p.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
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
        for (int u = 0; u < 8; u += 2) gelu_pair(v[j + u], v[j + u + 1]);
      }
    } else if (p.act == ACT_GELU_BWD) {
      const bf16* xp = p.aux + (int64_t)row * p.ld_aux + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float t[8];
        load8(xp + j, t);
#pragma unroll
        for (int u = 0; u < 8; u += 2) gelu_grad_mul_pair(t[u], t[u + 1], v[j + u], v[j + u + 1]);
      }
    }
    if (p.out_kind == OUT_BF16) {
      bf16* cp = reinterpret_cast<bf16*>(p.C) + (int64_t)row * p.ldc + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) store8(cp + j, v + j);
    } else if (p.out_kind == OUT_F32) {
      float* cp = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      float* cp = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]),
                     "f"(v[j + 3])
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
        x = gelu_act(x);
      } else if (p.act == ACT_GELU_BWD) {
        x *= gelu_act_grad(__bfloat162float(p.aux[(int64_t)row * p.ld_aux + col]));
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

// CL = 2: CTA pairs (thread-block cluster) work on two vertically adjacent M tiles of the same N tile; each CTA
// fetches half of the shared B tile and TMA-multicasts it to both, cutting the L2 -> SM operand traffic by a third.
template <int BN, int MODE, int STAGES, int CL>
__global__ void __launch_bounds__(gemm_threads<BN>(), 1) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                         const __grid_constant__ CUtensorMap tmB,
                                                                         const __grid_constant__ CUtensorMap tmC,
                                                                         const __grid_constant__ CUtensorMap tmAux,
                                                                         const GemmParams p) {
  constexpr int BM = GEMM_BM, BK = GEMM_BK;
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = 2 * BN;  // two accumulators
  static_assert(BN == 128 || BN == 256, "BN must give a power-of-two TMEM allocation <= 512");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
  constexpr int EPI_WARPS = gemm_epi_warps<BN>();
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_smem + EPI_WARPS * GEMM_EPI_BUF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint64_t* in_bars = tmem_empty + 2;     // [EPI_WARPS] epilogue input tiles (addend / GELU' aux)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_bars + GEMM_MAX_EPI_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_total = (p.K + BK - 1) / BK;
  // work items are (split, n tile, group of CL m tiles); CTA `rank` of a cluster takes m tile CL*group + rank
  const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;
  const int first_item = blockIdx.x / CL, item_stride = gridDim.x / CL;
  const int groups_m = (p.tiles_m + CL - 1) / CL;
  const int tiles_mn = groups_m * p.tiles_n;
  const int n_items = tiles_mn * p.splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CL);   // every CTA of the cluster must have consumed the slot (multicast B)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], EPI_WARPS);
    }
    for (int s = 0; s < EPI_WARPS; ++s) mbar_init(&in_bars[s], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();   // everything above overlapped the previous kernel; operands / outputs are touched only below

  if (warp == 0) {
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int item = first_item; item < n_items; item += item_stride) {
        const int split = item / tiles_mn, rem = item - split * tiles_mn;
        const int n0 = (rem / groups_m) * BN, m0 = ((rem % groups_m) * CL + rank) * BM;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb, ++cnt) {
          const int s = cnt % STAGES;
          const uint32_t ph = (cnt / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
          uint8_t* sA = smem + s * STAGE_BYTES;
          uint8_t* sB = sA + A_BYTES;
          const int kc = kb * BK;
          if (MODE == 0) {
            tma_load_2d(sA, &tmA, &full[s], kc, m0);
            if (CL == 1) {
              tma_load_2d(sB, &tmB, &full[s], kc, n0);
            } else {   // my half of the B rows, delivered to both CTAs
              tma_load_2d_mc(sB + rank * (BN / CL) * 128, &tmB, &full[s], kc, n0 + rank * (BN / CL), (uint16_t)((1 << CL) - 1));
            }
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + j * (BK * 128), &tmA, &full[s], m0 + 64 * j, kc);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {
              if (CL == 1) tma_load_2d(sB + j * (BK * 128), &tmB, &full[s], n0 + 64 * j, kc);
              else if ((j % CL) == rank) tma_load_2d_mc(sB + j * (BK * 128), &tmB, &full[s], n0 + 64 * j, kc, (uint16_t)((1 << CL) - 1));
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, MODE == 1, MODE == 1);
      uint32_t cnt = 0, it = 0;
      for (int item = first_item; item < n_items; item += item_stride, ++it) {
        const int split = item / tiles_mn;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aph ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb_begin; kb < kb_end; ++kb, ++cnt) {
          const int s = cnt % STAGES;
          const uint32_t ph = (cnt / STAGES) & 1;
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
            umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          if (CL == 1) umma_commit(&empty[s]); else umma_commit_mc(&empty[s], (uint16_t)((1 << CL) - 1));
        }
        umma_commit(&tmem_full[as]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;   // 64-column group of the tile owned by this warp
    const bool vec_ok = ((p.ldc & 7) == 0) && ((p.N & 7) == 0) && (p.addend == nullptr || (p.ld_addend & 7) == 0) &&
                        (p.aux == nullptr || (p.ld_aux & 7) == 0);
    uint8_t* ebuf = epi_smem + (warp - 2) * GEMM_EPI_BUF;
    uint8_t* ebuf2 = ebuf + 2048;
    uint64_t* in_bar = &in_bars[warp - 2];
    uint32_t in_ph = 0;
    const bool tma_in = p.tma_in != 0 && p.out_kind == OUT_BF16;
    const bool gelu = p.act == ACT_GELU;
    uint32_t it = 0;
    for (int item = first_item; item < n_items; item += item_stride, ++it) {
      const int split = item / tiles_mn, rem = item - split * tiles_mn;
      const int n0 = (rem / groups_m) * BN, m0 = ((rem % groups_m) * CL + rank) * BM;
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      const int row0 = m0 + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < p.M;
      const int colg = n0 + cg * 64;
      const uint32_t tmem_row = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN + cg * 64);
      auto issue_in = [&](int col) {   // TMA: [32 rows x 32 cols] bf16 input tile -> second staging tile
        if (lane == 0) {
          mbar_arrive_expect_tx(in_bar, 2048);
          tma_load_2d(ebuf2, &tmAux, in_bar, col, row0);
        }
      };
      auto release_tmem = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[as]);
      };
      auto staging_free = [&]() {   // the previous TMA stores have finished READING the staging tiles
        if (lane == 0 && !(p.dbg & 2)) tma_store_wait_read();
        __syncwarp();
      };
      if (tma_in && colg < p.N) issue_in(colg);   // overlaps the wait for the MMAs
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        const int col0 = colg + sub * 32;
        uint32_t r[32];
        __syncwarp();
        if (!(p.dbg & 4)) tmem_ld32(tmem_row + uint32_t(sub * 32), r);
        float v[32];
        if (col0 >= p.N) {   // nothing to write (uniform): only keep the TMEM protocol alive
          tmem_ld_wait();
          if (sub == 1) release_tmem();
          continue;
        }
        if (p.tma_store && p.out_kind == OUT_BF16) {
          float in_vals[32];
          if (tma_in) {
            mbar_wait(in_bar, in_ph);
            in_ph ^= 1;
            unstage_bf16(ebuf2, lane, in_vals);
            __syncwarp();
            if (sub == 0 && col0 + 32 < p.N) issue_in(col0 + 32);   // prefetch the second half's input tile
          }
          tmem_ld_wait();
          if (sub == 1) release_tmem();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          gemm_epilogue_math(p, row, col0, row_ok, tma_in ? in_vals : nullptr, v);
          staging_free();
          if (gelu) {   // pre-activation to the second tile, activation to the first
            stage_bf16(ebuf2, lane, v);
#pragma unroll
            for (int j = 0; j < 32; j += 2) gelu_pair(v[j], v[j + 1]);
          }
          stage_bf16(ebuf, lane, v);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && !(p.dbg & 1)) {
            if (gelu) tma_store_2d(&tmAux, ebuf2, col0, row0);
            tma_store_2d(&tmC, ebuf, col0, row0);
            tma_store_commit();
          }
        } else if (p.tma_store) {   // fp32 tile: plain store or L2 reduce-add (split-K weight gradients)
          tmem_ld_wait();
          if (sub == 1) release_tmem();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          gemm_epilogue_math(p, row, col0, row_ok, nullptr, v);
          staging_free();
          stage_f32(ebuf, lane, v);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (p.out_kind == OUT_F32_ATOMIC) tma_reduce_add_2d(&tmC, ebuf, col0, row0); else tma_store_2d(&tmC, ebuf, col0, row0);
            tma_store_commit();
          }
        } else {   // layouts a tensor map cannot describe (unaligned pitch): direct global accesses
          tmem_ld_wait();
          if (sub == 1) release_tmem();
          if (!row_ok) continue;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          gemm_epilogue_chunk(p, row, col0, v, vec_ok);
        }
      }
    }
    if (p.tma_store && lane == 0) tma_store_wait_all();   // global writes complete before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // no CTA exits while its peer may still multicast into it / signal its barriers
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
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

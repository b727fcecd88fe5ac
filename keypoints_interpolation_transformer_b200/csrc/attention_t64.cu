// Attention for sequences of at most 64 frames on the 5th-generation tensor cores, forward and backward: the shape every
// attention call of the default model has (A1_train.py:120-124 -> model.py:141-145 -> torch/nn/functional.py:6682 with
// T = 64, 8 heads of d = 32), i.e. B * NH independent 64 x 64 x 32 problems.  One tcgen05 tile is 128 lanes, so a work unit
// is (a PAIR of sequences, a PACK of two heads = one 128B-swizzled 64-column tile of Q / K / V):
//
//   * TMA loads the [2 sequences x 64 frames] x 64 columns tiles through a 3-D tensor map ([B, T, ld]; frames >= T and
//     sequences >= B arrive as zeros), double-buffered across units;
//   * one thread issues tcgen05.mma with M = 64 per sequence: sequence s of the pair writes the accumulator lanes
//     (i % 16) + 32 (i / 16) + 16 s  (row i), so the two sequences interleave in the 128 lanes of the same tensor-memory
//     columns and no score outside the two 64 x 64 diagonal blocks is ever computed; a head's 32 columns of a packed tile are
//     selected by a +64 byte start address (K-major operands) or as an N = 32 slice (MN-major operands);
//   * two softmax groups of four warps, one per head of the pack, one thread per (sequence, query) row: tcgen05.ld of the
//     row's 64 scores, mask terms folded into two floats per key ({bias * log2 e, cut index}: -inf iff (j > i and m[j] = 1),
//     + m[j] -- model.py:193-202, torch/nn/functional.py:6620), base-2 softmax, P as bf16 into the K-major swizzled layout
//     the tensor core reads back for O = P V (V = MN-major B operand, N = 32).
//
// Backward (recomputes P from the saved log-sum-exp):  S = Q K^T, dP = dO V^T  ->  P = 2^(S c + bias - lse),
// delta = rowsum(P * dP) (= rowsum(dO * O)), dS = P (dP - delta) scale  ->  dV = P^T dO, dK = dS^T Q (P / dS tiles as MN-major
// A operands), dQ = dS K; all five products on tcgen05, accumulators in tensor memory (448 of the 512 columns).
// SASS: UTCHMMA / LDTM / UTMALDG; no HMMA.  Verified layouts: tools/umma_probe.py.
#include "attention.cuh"
#include "gemm_sm100.cuh"

namespace kit {

constexpr int A64_TILE = 16384;   // [128 rows x 64 columns] bf16
constexpr float A64_LOG2E = 1.4426950408889634f, A64_LN2 = 0.6931471805599453f;

struct A64Params {
  int B, NH, Sq, Sk, packs, units;
  float scale, scale2;
  const float* frame_mask;
  int64_t frame_mask_stride;
  int flags;
  bf16* out;        // forward
  int64_t ldo;
  float* lse;
  const float* lse_in;   // backward
  bf16* dq;
  int64_t ld_dq;
  bf16* dk;
  int64_t ld_dk;
  bf16* dv;
  int64_t ld_dv;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ float a64_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// Folded mask terms of the 2 x 64 keys of a sequence pair, warp-private: kb = additive term in base 2 (-inf beyond Sk),
// kc = the key's index when it is cut for every earlier query (repeat-inc with m[j] = 1, or triangle), else -1.  The frame-mask
// values of the NEXT unit are requested (a64_load_fm) while the current one is processed and folded at its start.
__device__ __forceinline__ void a64_load_fm(const A64Params& p, int pair, int lane, float (&fmv)[4]) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int idx = lane * 4 + u, s = idx >> 6, j = idx & 63, b = pair * 2 + s;
    fmv[u] = (p.frame_mask != nullptr && b < p.B && j < p.Sk) ? __ldg(p.frame_mask + (int64_t)b * p.frame_mask_stride + j) : 0.f;
  }
}
__device__ __forceinline__ bool a64_fold_terms(const A64Params& p, const float (&fmv)[4], float* kb, float* kc, int lane) {
  bool any_cut = false;
  float b4[4], c4[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = (lane * 4 + u) & 63;
    const float fm = fmv[u];
    const bool cut = ((p.flags & KIT_MASK_REPEAT_INC) && fm == 1.f) || (p.flags & KIT_MASK_TRIANGLE);
    b4[u] = (j >= p.Sk) ? -INFINITY : ((p.flags & KIT_MASK_KEYPAD_ADD) ? fm * A64_LOG2E : 0.f);
    c4[u] = cut ? (float)j : -1.f;
    any_cut |= cut;
  }
  *reinterpret_cast<float4*>(kb + lane * 4) = make_float4(b4[0], b4[1], b4[2], b4[3]);
  *reinterpret_cast<float4*>(kc + lane * 4) = make_float4(c4[0], c4[1], c4[2], c4[3]);
  return __any_sync(0xffffffffu, any_cut);
}

// ------------------------------------------------------------------------------------------------ forward
struct A64FwdSmem {
  uint8_t q[2][A64_TILE];
  uint8_t k[2][A64_TILE];
  uint8_t v[2][A64_TILE];
  uint8_t p[4][A64_TILE];      // [slot * 2 + head of the pack]: [2 sequences x 64 queries] x 64 keys
  float kterm[8][2][128];      // per softmax warp: kb / kc
  uint64_t bars[20];
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(64 + 256, 1) attn64_fwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                 const __grid_constant__ CUtensorMap tmK,
                                                                 const __grid_constant__ CUtensorMap tmV, const A64Params p) {
  extern __shared__ uint8_t smem_raw[];
  A64FwdSmem& s = *reinterpret_cast<A64FwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* kv_full = &s.bars[0];    // [2]
  uint64_t* kv_empty = &s.bars[2];   // [2] every MMA that reads the slot's tiles has completed
  uint64_t* s_full = &s.bars[4];     // [4] scores written (tcgen05.commit)
  uint64_t* s_free = &s.bars[8];     // [4] scores read by the group (4 warps)
  uint64_t* p_full = &s.bars[12];    // [4] P in shared memory (4 warps)
  uint64_t* pv_done = &s.bars[16];   // [4] O accumulated (tcgen05.commit)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    pdl_launch_dependents();
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&kv_full[i], 1);
        mbar_init(&kv_empty[i], 1);
      }
      for (int i = 0; i < 4; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&s_free[i], 4);
        mbar_init(&p_full[i], 4);
        mbar_init(&pv_done[i], 1);
      }
      fence_barrier_init();
      fence_proxy_async();
    }
  }
  if (warp == 1) tmem_alloc<512>(&s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_slot;
  // tensor-memory columns: S[idx] at idx * 64, O[idx] at 256 + idx * 32   (idx = slot * 2 + head of the pack)
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int iu = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++iu) {
        const int sl = iu & 1, pair = u / p.packs, hp = u % p.packs;
        mbar_wait(&kv_empty[sl], ((iu >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[sl], 3 * A64_TILE);
        tma_load_3d(s.q[sl], &tmQ, &kv_full[sl], hp * 64, 0, pair * 2);
        tma_load_3d(s.k[sl], &tmK, &kv_full[sl], hp * 64, 0, pair * 2);
        tma_load_3d(s.v[sl], &tmV, &kv_full[sl], hp * 64, 0, pair * 2);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(64, 64, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(64, 32, false, true);
      const int n_units = (p.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int total = n_units * 2;
      int qk = 0, pv = 0;
      uint32_t idle = 0;
      // two in-order queues (as attention_tc.cu): S(n) needs its score columns drained and the unit's tiles; O(m) needs P(m)
      while (pv < total) {
        if (++idle > (1u << 24)) {   // a protocol bug becomes a trap instead of a hang
          printf("kit: attn64_fwd MMA queue stalled (block %d, qk %d, pv %d)\n", blockIdx.x, qk, pv);
          __trap();
        }
        if (qk < total) {
          const int iu = qk >> 1, g = qk & 1, sl = iu & 1, idx = sl * 2 + g;
          bool ready = mbar_try_wait(&s_free[idx], ((iu >> 1) & 1) ^ 1);
          if (ready && g == 0) ready = mbar_try_wait(&kv_full[sl], (iu >> 1) & 1);
          if (ready) {
            tc_fence_after();
            const uint32_t q_base = smem_u32(s.q[sl]) + g * 64, k_base = smem_u32(s.k[sl]) + g * 64;
#pragma unroll
            for (int sq = 0; sq < 2; ++sq)
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint64_t adesc = make_smem_desc_sw128(q_base + sq * 8192 + kk * 32, 0, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(k_base + sq * 8192 + kk * 32, 0, 1024);
                umma_bf16(tmem_base + (uint32_t(16 * sq) << 16) + uint32_t(idx * 64), adesc, bdesc, idesc_qk, kk > 0 ? 1u : 0u);
              }
            umma_commit(&s_full[idx]);
            ++qk;
            idle = 0;
          }
        }
        if (pv < qk) {
          const int iu = pv >> 1, g = pv & 1, sl = iu & 1, idx = sl * 2 + g;
          if (mbar_try_wait(&p_full[idx], (iu >> 1) & 1)) {
            tc_fence_after();
            const uint32_t p_base = smem_u32(s.p[idx]), v_base = smem_u32(s.v[sl]) + g * 64;
#pragma unroll
            for (int sq = 0; sq < 2; ++sq)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {   // 16 keys per MMA
                const uint64_t adesc = make_smem_desc_sw128(p_base + sq * 8192 + kk * 32, 0, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(v_base + sq * 8192 + kk * 2048, 8192, 1024);
                umma_bf16(tmem_base + (uint32_t(16 * sq) << 16) + uint32_t(256 + idx * 32), adesc, bdesc, idesc_pv, kk > 0 ? 1u : 0u);
              }
            umma_commit(&pv_done[idx]);
            if (g == 1) umma_commit(&kv_empty[sl]);
            ++pv;
            idle = 0;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax groups: group = head of the pack
    const int grp = (warp - 2) >> 2, qd = warp & 3;
    const int sq = lane >> 4;                    // sequence of the pair
    const int qi = (lane & 15) + 16 * qd;        // query row within the sequence (M = 64 accumulator lane layout)
    const int prow = sq * 64 + qi;               // row of the [128 x 64] P tile
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    float* kb = s.kterm[warp - 2][0];
    float* kc = s.kterm[warp - 2][1];
    const uint32_t kb_s = smem_u32(kb) + sq * 256, kc_s = smem_u32(kc) + sq * 256;
    const float qif = (float)qi;
    const uint64_t sc2 = pk2(p.scale2, p.scale2);
    // The output of unit n (O / l, log-sum-exp) is drained after the softmax of unit n + 1: the P V products complete while
    // this group works on the next score tile instead of being waited for.
    struct Pending { int idx, b, h; uint32_t ph; float l, m; } pend = {0, 0, 0, 0, 1.f, 0.f};
    bool have_pend = false;
    auto drain = [&]() {
      mbar_wait(&pv_done[pend.idx], pend.ph);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(tmem_base + lane_base + uint32_t(256 + pend.idx * 32), o);
      tmem_ld_wait();
      tc_fence_before();
      if (pend.b < p.B && qi < p.Sq) {
        const float inv = 1.f / pend.l;
        bf16* dst = p.out + ((int64_t)pend.b * p.Sq + qi) * p.ldo + pend.h * 32;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float f[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) f[t] = __uint_as_float(o[8 * e + t]) * inv;
          store8(dst + 8 * e, f);
        }
        if (p.lse != nullptr) p.lse[((int64_t)pend.b * p.NH + pend.h) * p.Sq + qi] = (pend.m + log2f(pend.l)) * A64_LN2;
      }
    };
    float fmv[4];
    if ((int)blockIdx.x < p.units) a64_load_fm(p, (int)blockIdx.x / p.packs, lane, fmv);
    int iu = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++iu) {
      const int sl = iu & 1, idx = sl * 2 + grp;
      const uint32_t ph = (iu >> 1) & 1;
      const int pair = u / p.packs, hp = u % p.packs;
      const int b = pair * 2 + sq, h = hp * 2 + grp;
      __syncwarp();
      const bool need_cut = a64_fold_terms(p, fmv, kb, kc, lane);
      if (u + (int)gridDim.x < p.units) a64_load_fm(p, (u + (int)gridDim.x) / p.packs, lane, fmv);
      __syncwarp();
      mbar_wait(&s_full[idx], ph);
      tc_fence_after();
      float x[64];
      uint32_t* xr = reinterpret_cast<uint32_t*>(x);
      const uint32_t s_addr = tmem_base + lane_base + uint32_t(idx * 64);
      tmem_ld32(s_addr, xr);
      tmem_ld32(s_addr + 32, xr + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[idx]);
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const uint4 kbw = lds128(kb_s + 16 * c4);
        const uint32_t kbv[4] = {kbw.x, kbw.y, kbw.z, kbw.w};
        uint4 kcw = make_uint4(0, 0, 0, 0);
        if (need_cut) kcw = lds128(kc_s + 16 * c4);
        const uint32_t kcv[4] = {kcw.x, kcw.y, kcw.z, kcw.w};
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
          const int c = 4 * c4 + e;
          float v0, v1;
          up2(fma2(pk2(x[c], x[c + 1]), sc2, pk2(__uint_as_float(kbv[e]), __uint_as_float(kbv[e + 1]))), v0, v1);
          if (need_cut) {
            v0 = (__uint_as_float(kcv[e]) > qif) ? -INFINITY : v0;
            v1 = (__uint_as_float(kcv[e + 1]) > qif) ? -INFINITY : v1;
          }
          x[c] = v0;
          x[c + 1] = v1;
          mx4[e] = fmaxf(mx4[e], v0);
          mx4[e + 1] = fmaxf(mx4[e + 1], v1);
        }
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_ref = (mx == -INFINITY) ? 0.f : mx;
      uint64_t rs2[4] = {pk2(0.f, 0.f), pk2(0.f, 0.f), pk2(0.f, 0.f), pk2(0.f, 0.f)};
      const uint64_t nm2 = pk2(-m_ref, -m_ref);
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float a0, a1;
        up2(add2(pk2(x[2 * c], x[2 * c + 1]), nm2), a0, a1);
        const float p0 = a64_ex2(a0), p1 = a64_ex2(a1);
        rs2[c & 3] = add2(rs2[c & 3], pk2(p0, p1));
        pk[c] = pack_bf16(p0, p1);
      }
      float r0, r1, r2, r3, r4, r5, r6, r7;
      up2(rs2[0], r0, r1); up2(rs2[1], r2, r3); up2(rs2[2], r4, r5); up2(rs2[3], r6, r7);
      const float l = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
      // P buffer idx was last read by the P V product of two units ago, whose completion this group waited for in drain()
      const uint32_t p_row = smem_u32(s.p[idx]) + prow * 128;
      const uint32_t sw = prow & 7;
#pragma unroll
      for (int c = 0; c < 8; ++c) sts128(p_row + ((uint32_t(c) ^ sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[idx]);
      if (have_pend) drain();
      pend = Pending{idx, b, h, ph, l, m_ref};
      have_pend = true;
    }
    if (have_pend) drain();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ backward
struct A64BwdSmem {
  uint8_t q[2][A64_TILE];
  uint8_t k[2][A64_TILE];
  uint8_t v[2][A64_TILE];
  uint8_t d_o[2][A64_TILE];
  uint8_t p[2][A64_TILE];      // [head of the pack]: P  [2 sequences x 64 queries] x 64 keys
  uint8_t ds[2][A64_TILE];     //                     dS
  float kterm[8][2][128];
  uint64_t bars[16];
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(64 + 256, 1) attn64_bwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                 const __grid_constant__ CUtensorMap tmK,
                                                                 const __grid_constant__ CUtensorMap tmV,
                                                                 const __grid_constant__ CUtensorMap tmDO, const A64Params p) {
  extern __shared__ uint8_t smem_raw[];
  A64BwdSmem& s = *reinterpret_cast<A64BwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* t_full = &s.bars[0];      // [2] tiles of a unit have landed
  uint64_t* t_empty = &s.bars[2];     // [2] every MMA that reads them has completed
  uint64_t* sdp_full = &s.bars[4];    // [2] per head: S and dP written
  uint64_t* s_free = &s.bars[6];      // [2] per head: S and dP read by the group (4 warps)
  uint64_t* p_full = &s.bars[8];      // [2] per head: P and dS in shared memory (4 warps)
  uint64_t* grad_full = &s.bars[10];  // [2] per head: dQ, dK, dV accumulated
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    pdl_launch_dependents();
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&t_full[i], 1);
        mbar_init(&t_empty[i], 1);
        mbar_init(&sdp_full[i], 1);
        mbar_init(&s_free[i], 4);
        mbar_init(&p_full[i], 4);
        mbar_init(&grad_full[i], 1);
      }
      fence_barrier_init();
      fence_proxy_async();
    }
  }
  if (warp == 1) tmem_alloc<512>(&s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_slot;
  // tensor-memory columns (g = head of the pack): S at g * 64, dP at 128 + g * 64, dQ at 256 + g * 32, dK at 320 + g * 32,
  // dV at 384 + g * 32
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int iu = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++iu) {
        const int sl = iu & 1, pair = u / p.packs, hp = u % p.packs;
        mbar_wait(&t_empty[sl], ((iu >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&t_full[sl], 4 * A64_TILE);
        tma_load_3d(s.q[sl], &tmQ, &t_full[sl], hp * 64, 0, pair * 2);
        tma_load_3d(s.k[sl], &tmK, &t_full[sl], hp * 64, 0, pair * 2);
        tma_load_3d(s.v[sl], &tmV, &t_full[sl], hp * 64, 0, pair * 2);
        tma_load_3d(s.d_o[sl], &tmDO, &t_full[sl], hp * 64, 0, pair * 2);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_kk = make_idesc_bf16(64, 64, false, false);   // S, dP
      constexpr uint32_t idesc_mm = make_idesc_bf16(64, 32, true, true);     // dV, dK
      constexpr uint32_t idesc_km = make_idesc_bf16(64, 32, false, true);    // dQ
      const int n_units = (p.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int total = n_units * 2;
      int sd = 0, gr = 0;
      uint32_t idle = 0;
      while (gr < total) {
        if (++idle > (1u << 24)) {
          printf("kit: attn64_bwd MMA queue stalled (block %d, sd %d, gr %d)\n", blockIdx.x, sd, gr);
          __trap();
        }
        if (sd < total) {
          const int iu = sd >> 1, g = sd & 1, sl = iu & 1;
          bool ready = mbar_try_wait(&s_free[g], (iu & 1) ^ 1);
          if (ready && g == 0) ready = mbar_try_wait(&t_full[sl], (iu >> 1) & 1);
          if (ready) {
            tc_fence_after();
            const uint32_t q_base = smem_u32(s.q[sl]) + g * 64, k_base = smem_u32(s.k[sl]) + g * 64;
            const uint32_t v_base = smem_u32(s.v[sl]) + g * 64, do_base = smem_u32(s.d_o[sl]) + g * 64;
#pragma unroll
            for (int sq = 0; sq < 2; ++sq) {
              const uint32_t lanes = uint32_t(16 * sq) << 16;
#pragma unroll
              for (int kk = 0; kk < 2; ++kk)
                umma_bf16(tmem_base + lanes + uint32_t(g * 64), make_smem_desc_sw128(q_base + sq * 8192 + kk * 32, 0, 1024),
                          make_smem_desc_sw128(k_base + sq * 8192 + kk * 32, 0, 1024), idesc_kk, kk > 0 ? 1u : 0u);
#pragma unroll
              for (int kk = 0; kk < 2; ++kk)
                umma_bf16(tmem_base + lanes + uint32_t(128 + g * 64), make_smem_desc_sw128(do_base + sq * 8192 + kk * 32, 0, 1024),
                          make_smem_desc_sw128(v_base + sq * 8192 + kk * 32, 0, 1024), idesc_kk, kk > 0 ? 1u : 0u);
            }
            umma_commit(&sdp_full[g]);
            ++sd;
            idle = 0;
          }
        }
        if (gr < sd) {
          const int iu = gr >> 1, g = gr & 1, sl = iu & 1;
          if (mbar_try_wait(&p_full[g], iu & 1)) {
            tc_fence_after();
            const uint32_t p_base = smem_u32(s.p[g]), ds_base = smem_u32(s.ds[g]);
            const uint32_t q_base = smem_u32(s.q[sl]) + g * 64, k_base = smem_u32(s.k[sl]) + g * 64;
            const uint32_t do_base = smem_u32(s.d_o[sl]) + g * 64;
#pragma unroll
            for (int sq = 0; sq < 2; ++sq) {
              const uint32_t lanes = uint32_t(16 * sq) << 16;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {   // 16 queries (dV, dK) / 16 keys (dQ) per MMA
                const uint32_t roff = sq * 8192 + kk * 2048;
                umma_bf16(tmem_base + lanes + uint32_t(384 + g * 32), make_smem_desc_sw128(p_base + roff, 8192, 1024),
                          make_smem_desc_sw128(do_base + roff, 8192, 1024), idesc_mm, kk > 0 ? 1u : 0u);
                umma_bf16(tmem_base + lanes + uint32_t(320 + g * 32), make_smem_desc_sw128(ds_base + roff, 8192, 1024),
                          make_smem_desc_sw128(q_base + roff, 8192, 1024), idesc_mm, kk > 0 ? 1u : 0u);
                umma_bf16(tmem_base + lanes + uint32_t(256 + g * 32), make_smem_desc_sw128(ds_base + sq * 8192 + kk * 32, 0, 1024),
                          make_smem_desc_sw128(k_base + roff, 8192, 1024), idesc_km, kk > 0 ? 1u : 0u);
              }
            }
            umma_commit(&grad_full[g]);
            if (g == 1) umma_commit(&t_empty[sl]);
            ++gr;
            idle = 0;
          }
        }
      }
    }
    __syncwarp();
  } else {
    const int grp = (warp - 2) >> 2, qd = warp & 3;
    const int sq = lane >> 4;
    const int qi = (lane & 15) + 16 * qd;   // query row (S, dP, dQ) = key row (dK, dV) of this lane
    const int prow = sq * 64 + qi;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    float* kb = s.kterm[warp - 2][0];
    float* kc = s.kterm[warp - 2][1];
    const uint32_t kb_s = smem_u32(kb) + sq * 256, kc_s = smem_u32(kc) + sq * 256;
    const float qif = (float)qi;
    const uint64_t sc2 = pk2(p.scale2, p.scale2);
    // The gradients of unit n (dQ / dK / dV accumulators) are drained in the middle of unit n + 1 -- after its P / dS values are
    // computed, before they are written to the shared tiles the products of unit n still read -- so the 24 MMAs of a unit
    // run under the next unit's exponentials instead of being waited for.
    struct Pending { int b, h; uint32_t ph; } pend = {0, 0, 0};
    bool have_pend = false;
    auto drain_one = [&](uint32_t col, bf16* base, int64_t ld, int n_rows) {
      uint32_t o[32];
      tmem_ld32(tmem_base + lane_base + col, o);
      tmem_ld_wait();
      if (pend.b < p.B && qi < n_rows) {
        bf16* dst = base + ((int64_t)pend.b * n_rows + qi) * ld + pend.h * 32;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float f[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) f[t] = __uint_as_float(o[8 * e + t]);
          store8(dst + 8 * e, f);
        }
      }
    };
    auto drain = [&]() {   // dQ (row = query), dK / dV (row = key) -> global
      mbar_wait(&grad_full[grp], pend.ph);
      tc_fence_after();
      drain_one(uint32_t(256 + grp * 32), p.dq, p.ld_dq, p.Sq);
      drain_one(uint32_t(320 + grp * 32), p.dk, p.ld_dk, p.Sk);
      drain_one(uint32_t(384 + grp * 32), p.dv, p.ld_dv, p.Sk);
      tc_fence_before();
    };
    auto load_lse = [&](int u) {
      const int bb = (u / p.packs) * 2 + sq, hh = (u % p.packs) * 2 + grp;
      return (bb < p.B && qi < p.Sq) ? __ldg(p.lse_in + ((int64_t)bb * p.NH + hh) * p.Sq + qi) * A64_LOG2E : INFINITY;
    };
    float fmv[4], lse_next = INFINITY;
    if ((int)blockIdx.x < p.units) {
      a64_load_fm(p, (int)blockIdx.x / p.packs, lane, fmv);
      lse_next = load_lse((int)blockIdx.x);
    }
    int iu = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++iu) {
      const uint32_t ph = iu & 1;
      const int pair = u / p.packs, hp = u % p.packs;
      const int b = pair * 2 + sq, h = hp * 2 + grp;
      const float lse2 = lse_next;
      __syncwarp();
      const bool need_cut = a64_fold_terms(p, fmv, kb, kc, lane);
      if (u + (int)gridDim.x < p.units) {
        a64_load_fm(p, (u + (int)gridDim.x) / p.packs, lane, fmv);
        lse_next = load_lse(u + (int)gridDim.x);
      }
      __syncwarp();
      mbar_wait(&sdp_full[grp], ph);
      tc_fence_after();
      float x[64], dp[64];
      const uint32_t s_addr = tmem_base + lane_base + uint32_t(grp * 64);
      tmem_ld32(s_addr, reinterpret_cast<uint32_t*>(x));
      tmem_ld32(s_addr + 32, reinterpret_cast<uint32_t*>(x) + 32);
      tmem_ld32(s_addr + 128, reinterpret_cast<uint32_t*>(dp));
      tmem_ld32(s_addr + 160, reinterpret_cast<uint32_t*>(dp) + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[grp]);
      // P = 2^(S c + bias - lse2) (0 where masked); delta = sum_j P dP
      const uint64_t nl2 = pk2(-lse2, -lse2);
      uint64_t dl2[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const uint4 kbw = lds128(kb_s + 16 * c4);
        const uint32_t kbv[4] = {kbw.x, kbw.y, kbw.z, kbw.w};
        uint4 kcw = make_uint4(0, 0, 0, 0);
        if (need_cut) kcw = lds128(kc_s + 16 * c4);
        const uint32_t kcv[4] = {kcw.x, kcw.y, kcw.z, kcw.w};
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
          const int c = 4 * c4 + e;
          float v0, v1;
          up2(add2(fma2(pk2(x[c], x[c + 1]), sc2, pk2(__uint_as_float(kbv[e]), __uint_as_float(kbv[e + 1]))), nl2), v0, v1);
          if (need_cut) {
            v0 = (__uint_as_float(kcv[e]) > qif) ? -INFINITY : v0;
            v1 = (__uint_as_float(kcv[e + 1]) > qif) ? -INFINITY : v1;
          }
          const float p0 = a64_ex2(v0), p1 = a64_ex2(v1);
          x[c] = p0;
          x[c + 1] = p1;
          dl2[(c >> 1) & 1] = fma2(pk2(p0, p1), pk2(dp[c], dp[c + 1]), dl2[(c >> 1) & 1]);
        }
      }
      float d0, d1, d2, d3;
      up2(dl2[0], d0, d1); up2(dl2[1], d2, d3);
      const float delta = (d0 + d1) + (d2 + d3);
      // dS = P (dP - delta) scale; both tiles as bf16 rows of the [128 x 64] swizzled tiles
      const uint64_t nd2 = pk2(-delta, -delta), scl2 = pk2(p.scale, p.scale);
      uint32_t pp[32], dd[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float a0, a1;
        const uint64_t pr = pk2(x[2 * c], x[2 * c + 1]);
        up2(mul2(mul2(pr, add2(pk2(dp[2 * c], dp[2 * c + 1]), nd2)), scl2), a0, a1);
        pp[c] = pack_bf16(x[2 * c], x[2 * c + 1]);
        dd[c] = pack_bf16(a0, a1);
      }
      if (have_pend) drain();   // the previous unit's products have read the P / dS tiles; its accumulators are free after this
      const uint32_t p_row = smem_u32(s.p[grp]) + prow * 128, ds_row = smem_u32(s.ds[grp]) + prow * 128;
      const uint32_t sw = prow & 7;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        sts128(p_row + ((uint32_t(c) ^ sw) << 4), pp[4 * c], pp[4 * c + 1], pp[4 * c + 2], pp[4 * c + 3]);
        sts128(ds_row + ((uint32_t(c) ^ sw) << 4), dd[4 * c], dd[4 * c + 1], dd[4 * c + 2], dd[4 * c + 3]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[grp]);
      pend = Pending{b, h, ph};
      have_pend = true;
    }
    if (have_pend) drain();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host
bool attention_t64_supported(int NH, int Sq, int Sk, int d, const KitAttnMask* mask, const void* const* ptrs, const int64_t* lds, int n) {
  const char* e = getenv("KIT_ATTN_T64");   // KIT_ATTN_T64=0: the mma.sync tile kernels (A/B measurements, tests)
  if (e != nullptr && e[0] == '0') return false;
  if (d != 32 || (NH & 1) != 0 || Sq > 64 || Sk > 64 || Sq != Sk) return false;
  if (mask != nullptr && mask->bias != nullptr) return false;
  for (int i = 0; i < n; ++i)
    if ((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) != 0 || (lds[i] * 2) % 16 != 0) return false;
  return true;
}

static int a64_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}
static int a64_map(CUtensorMap* m, const bf16* ptr, int64_t ld, int B, int S, int NH) {
  return make_tensor_map_3d(m, ptr, (uint64_t)NH * 32, (uint64_t)S, (uint64_t)B, (uint64_t)ld * 2, (uint64_t)S * ld * 2, 64, 64, 2);
}
static void a64_params(A64Params& p, int B, int NH, int Sq, int Sk, const KitAttnMask* mask) {
  p = A64Params{};
  p.B = B; p.NH = NH; p.Sq = Sq; p.Sk = Sk;
  p.packs = NH / 2;
  p.units = ((B + 1) / 2) * p.packs;
  p.scale = rsqrtf(32.f);
  p.scale2 = p.scale * A64_LOG2E;
  p.frame_mask = mask != nullptr ? mask->frame_mask : nullptr;
  p.frame_mask_stride = mask != nullptr ? mask->frame_mask_stride : 0;
  p.flags = (mask != nullptr && mask->frame_mask != nullptr) ? mask->flags : (mask != nullptr ? (mask->flags & KIT_MASK_TRIANGLE) : 0);
}
template <typename... KArgs, typename... Args>
static int a64_launch(void (*kernel)(KArgs...), int units, int smem, cudaStream_t st, Args... args) {
  const int sms = a64_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(units < sms ? units : sms);
  cfg.blockDim = dim3(64 + 256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  KIT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  return KIT_OK;
}

int attention_t64_fwd(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, bf16* out, int64_t ldo,
                      float* lse, int B, int NH, int Sq, int Sk, const KitAttnMask* mask, cudaStream_t st) {
  constexpr int smem = (int)sizeof(A64FwdSmem) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    KIT_CHECK_CUDA(cudaFuncSetAttribute(attn64_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = a64_map(&tmQ, q, ldq, B, Sq, NH))) return rc;
  if ((rc = a64_map(&tmK, k, ldk, B, Sk, NH))) return rc;
  if ((rc = a64_map(&tmV, v, ldv, B, Sk, NH))) return rc;
  A64Params p;
  a64_params(p, B, NH, Sq, Sk, mask);
  p.out = out; p.ldo = ldo; p.lse = lse;
  return a64_launch(attn64_fwd_kernel, p.units, smem, st, tmQ, tmK, tmV, p);
}

int attention_t64_bwd(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, const bf16* dout,
                      int64_t ld_do, const float* lse, bf16* dq, int64_t ld_dq, bf16* dk, int64_t ld_dk, bf16* dv, int64_t ld_dv,
                      int B, int NH, int Sq, int Sk, const KitAttnMask* mask, cudaStream_t st) {
  constexpr int smem = (int)sizeof(A64BwdSmem) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    KIT_CHECK_CUDA(cudaFuncSetAttribute(attn64_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  CUtensorMap tmQ, tmK, tmV, tmDO;
  int rc;
  if ((rc = a64_map(&tmQ, q, ldq, B, Sq, NH))) return rc;
  if ((rc = a64_map(&tmK, k, ldk, B, Sk, NH))) return rc;
  if ((rc = a64_map(&tmV, v, ldv, B, Sk, NH))) return rc;
  if ((rc = a64_map(&tmDO, dout, ld_do, B, Sq, NH))) return rc;
  A64Params p;
  a64_params(p, B, NH, Sq, Sk, mask);
  p.lse_in = lse;
  p.dq = dq; p.ld_dq = ld_dq; p.dk = dk; p.ld_dk = ld_dk; p.dv = dv; p.ld_dv = ld_dv;
  return a64_launch(attn64_bwd_kernel, p.units, smem, st, tmQ, tmK, tmV, tmDO, p);
}

}  // namespace kit

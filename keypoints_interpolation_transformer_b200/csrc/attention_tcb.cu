// Attention BACKWARD on the 5th-generation tensor cores for longer sequences (more than one 64-key tile): the gradient of
// softmax(Q K^T / sqrt(d) + mask) V that autograd derives for model.py:141-145 (A1_train.py:134), with the masks of
// A1_train.py:117-124 synthesised from the [B, T] frame mask as in the forward kernels.  All five products run as tcgen05.mma with
// fp32 accumulators in tensor memory, operands staged by TMA (128B swizzle):
//
//   unit  = (batch, 64 feature columns of the packed head dimension = one head of 64 or two heads of 32, 128-key tile j);
//   step  = (128-query tile i, head hh of the pack), P recomputed from the saved log-sum-exp:
//
//     warp 1 (one elected thread):  S  = Q_i K_j^T          d / 16 MMAs 128 x 128 x 16   -> TMEM columns   0 .. 127
//                                   dP = dO_i V_j^T         d / 16 MMAs                  -> TMEM columns 128 .. 255
//     two softmax groups (4 warps each, one thread per query row, group g = key columns 64 g .. 64 g + 63 of the tile):
//                                   P = 2^(S c + bias - lse),  dS = P (dP - delta) / sqrt(d)   (delta = rowsum(dO * O), once per call:
//                                   attn_delta_kernel) -> bf16 tiles [128 queries x 128 keys] in shared memory
//     warp 1:                       dV_j += P^T dO_i        8 MMAs 128 x d x 16 (P / dS as MN-major A operands)  -> columns 384 ..
//                                   dK_j += dS^T Q_i        8 MMAs                                                -> columns 320 ..
//                                   dQ_i  = dS K_j          8 MMAs (K tile as MN-major B operand)                 -> columns 256 ..
//
// dK / dV accumulate in tensor memory over the unit's query tiles and leave once as bf16; the dQ partial of every step is added
// to the fp32 workspace with 16-byte vector reductions (red.global.add.v4.f32) by the softmax warps in the middle of the NEXT step
// -- after its exponentials, before its P / dS tiles overwrite the ones the products still read -- so the 24 MMAs of a step run
// under the next step's softmax.  S / dP of step n + 1 are issued as soon as step n's scores are in registers.  Persistent over
// units (K / V double-buffered across units, Q / dO across query tiles).  Rows beyond the sequence get lse = +inf (P = 0), keys
// beyond it bias = -inf.  Explicit additive bias tensors and other head sizes go through the mma.sync kernels of attention.cu.
// SASS: UTCHMMA / LDTM / UTMALDG, no HMMA.
#include "attention.cuh"
#include "gemm_sm100.cuh"

namespace kit {

constexpr int ATB_TILE = 128 * 64 * 2;   // [128 rows x 64 columns] bf16 = 16 KB
constexpr float ATB_LOG2E = 1.4426950408889634f;

struct AtbParams {
  int B, NH, Sq, Sk, q_tiles, k_tiles, units;
  float scale, scale2;
  const float* frame_mask;
  int64_t frame_mask_stride;
  int flags;
  const float* lse;      // [B, NH, Sq]
  const float* delta;    // [B, NH, Sq]
  float* dq_acc;         // [B * Sq, NH * D] fp32, zero on entry
  int64_t ld_acc;
  bf16* dk;
  int64_t ld_dk;
  bf16* dv;
  int64_t ld_dv;
};

__device__ __forceinline__ bool atb_test(uint64_t* bar, uint32_t parity) {   // non-blocking probe (see attention_t64.cu mbar_test)
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool atb_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ bool atb_uniform(bool v) { return __shfl_sync(0xffffffffu, (int)v, 0) != 0; }
__device__ __forceinline__ float atb_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct AtbSmem {
  uint8_t k[2][ATB_TILE];      // per unit, double-buffered across units
  uint8_t v[2][ATB_TILE];
  uint8_t q[2][ATB_TILE];      // per query tile
  uint8_t d_o[2][ATB_TILE];
  uint8_t p[2 * ATB_TILE];     // [128 queries x 128 keys] bf16: two [128 x 64] blocks (keys 0..63 | 64..127)
  uint8_t ds[2 * ATB_TILE];
  float kterm[8][2][128];      // per softmax warp: kb / kc of the unit's 128 keys
  uint64_t bars[16];
  uint32_t tmem_slot;
};

template <int D>
__global__ void __launch_bounds__(64 + 256, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                  const __grid_constant__ CUtensorMap tmK,
                                                                  const __grid_constant__ CUtensorMap tmV,
                                                                  const __grid_constant__ CUtensorMap tmDO, const AtbParams p) {
  constexpr int HP = 64 / D, KS = D / 16;   // heads per 64-column pack, K = 16 steps of a head
  extern __shared__ uint8_t smem_raw[];
  AtbSmem& s = *reinterpret_cast<AtbSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* kv_full = &s.bars[0];     // [2]
  uint64_t* kv_empty = &s.bars[2];    // [2] every MMA of the unit has completed (tcgen05.commit)
  uint64_t* qd_full = &s.bars[4];     // [2]
  uint64_t* qd_empty = &s.bars[6];    // [2] every MMA that reads this Q / dO tile has completed
  uint64_t* sdp_full = &s.bars[8];    // S and dP written (tcgen05.commit)
  uint64_t* s_free = &s.bars[9];      // S and dP read by the 8 softmax warps
  uint64_t* p_full = &s.bars[10];     // P and dS in shared memory (8 warps)
  uint64_t* grad_full = &s.bars[11];  // the step's dV / dK / dQ products have completed: P / dS tiles free, dQ ready
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NS = p.q_tiles * HP;      // steps per unit
  const int packs = p.NH / HP;
  auto decode = [&](int u, int& b, int& hg, int& j) {
    j = u % p.k_tiles;
    const int rest = u / p.k_tiles;
    hg = rest % packs;
    b = rest / packs;
  };

  if (warp == 0) {
    pdl_launch_dependents();
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&kv_full[i], 1);
        mbar_init(&kv_empty[i], 1);
        mbar_init(&qd_full[i], 1);
        mbar_init(&qd_empty[i], 1);
      }
      mbar_init(sdp_full, 1);
      mbar_init(s_free, 8);
      mbar_init(p_full, 8);
      mbar_init(grad_full, 1);
      fence_barrier_init();
      fence_proxy_async();
    }
  }
  if (warp == 1) tmem_alloc<512>(&s.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int iu = 0, it = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++iu) {
        int b, hg, j;
        decode(u, b, hg, j);
        const int sl = iu & 1;
        mbar_wait(&kv_empty[sl], ((iu >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[sl], 2 * ATB_TILE);
        tma_load_2d(s.k[sl], &tmK, &kv_full[sl], hg * 64, b * p.Sk + j * 128);
        tma_load_2d(s.v[sl], &tmV, &kv_full[sl], hg * 64, b * p.Sk + j * 128);
        for (int i = 0; i < p.q_tiles; ++i, ++it) {
          const int st = it & 1;
          mbar_wait(&qd_empty[st], ((it >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&qd_full[st], 2 * ATB_TILE);
          tma_load_2d(s.q[st], &tmQ, &qd_full[st], hg * 64, b * p.Sq + i * 128);
          tma_load_2d(s.d_o[st], &tmDO, &qd_full[st], hg * 64, b * p.Sq + i * 128);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: warp-uniform loop, instructions under elect.sync
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);   // S, dP
    constexpr uint32_t idesc_mm = make_idesc_bf16(128, D, true, true);      // dV, dK: P^T / dS^T (MN-major A), dO / Q (MN-major B)
    constexpr uint32_t idesc_km = make_idesc_bf16(128, D, false, true);     // dQ: dS (K-major A), K tile (MN-major B)
    const bool leader = atb_elect();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int n_units = (p.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = n_units * NS;
    const uint32_t p_base = smem_u32(s.p), ds_base = smem_u32(s.ds);
    int sd = 0, gr = 0;
    uint32_t idle = 0;
    while (gr < total) {
      if (++idle > (1u << 27)) {   // a protocol bug becomes a trap instead of a hang
        if (leader) printf("kit: attn_bwd_tc MMA queue stalled (block %d, sd %d, gr %d of %d)\n", blockIdx.x, sd, gr, total);
        __trap();
      }
      if (sd < total) {
        const int iu = sd / NS, ns = sd - iu * NS, i = ns / HP, hh = ns - i * HP, it = iu * p.q_tiles + i;
        bool ready = atb_test(s_free, (sd & 1) ^ 1);   // the previous step's S / dP are in registers
        if (ready && hh == 0) ready = atb_test(&qd_full[it & 1], (it >> 1) & 1);
        if (ready && ns == 0) ready = atb_test(&kv_full[iu & 1], (iu >> 1) & 1);
        if (atb_uniform(ready)) {
          tc_fence_after();
          const uint32_t q_b = smem_u32(s.q[it & 1]) + hh * D * 2, k_b = smem_u32(s.k[iu & 1]) + hh * D * 2;
          const uint32_t do_b = smem_u32(s.d_o[it & 1]) + hh * D * 2, v_b = smem_u32(s.v[iu & 1]) + hh * D * 2;
#pragma unroll
          for (int kk = 0; kk < KS; ++kk)
            if (leader)
              umma_bf16(tb, make_smem_desc_sw128(q_b + kk * 32, 0, 1024), make_smem_desc_sw128(k_b + kk * 32, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < KS; ++kk)
            if (leader)
              umma_bf16(tb + 128, make_smem_desc_sw128(do_b + kk * 32, 0, 1024), make_smem_desc_sw128(v_b + kk * 32, 0, 1024), idesc_s,
                        kk > 0 ? 1u : 0u);
          if (leader) umma_commit(sdp_full);
          ++sd;
          idle = 0;
        }
      }
      if (gr < sd) {
        const int iu = gr / NS, ns = gr - iu * NS, i = ns / HP, hh = ns - i * HP, it = iu * p.q_tiles + i;
        if (atb_uniform(atb_test(p_full, gr & 1))) {
          tc_fence_after();
          const uint32_t q_b = smem_u32(s.q[it & 1]) + hh * D * 2, k_b = smem_u32(s.k[iu & 1]) + hh * D * 2;
          const uint32_t do_b = smem_u32(s.d_o[it & 1]) + hh * D * 2;
          const uint32_t acc = (i > 0) ? 1u : 0u;   // dK / dV accumulate over the unit's query tiles
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {   // 16 queries (dV, dK) / 16 keys (dQ) per MMA
            const uint32_t roff = kk * 2048;
            if (leader) {
              // MN-major operands: 64-element atoms LBO apart (P / dS: the second 64-key block; B operands: one atom), 8-row groups 1024 B apart
              umma_bf16(tb + uint32_t(384 + hh * D), make_smem_desc_sw128(p_base + roff, ATB_TILE, 1024),
                        make_smem_desc_sw128(do_b + roff, 8192, 1024), idesc_mm, (acc | (kk > 0 ? 1u : 0u)));
              umma_bf16(tb + uint32_t(320 + hh * D), make_smem_desc_sw128(ds_base + roff, ATB_TILE, 1024),
                        make_smem_desc_sw128(q_b + roff, 8192, 1024), idesc_mm, (acc | (kk > 0 ? 1u : 0u)));
              umma_bf16(tb + uint32_t(256 + hh * D), make_smem_desc_sw128(ds_base + (kk >> 2) * ATB_TILE + (kk & 3) * 32, 0, 1024),
                        make_smem_desc_sw128(k_b + roff, 8192, 1024), idesc_km, kk > 0 ? 1u : 0u);
            }
          }
          if (leader) {
            umma_commit(grad_full);
            if (hh == HP - 1) umma_commit(&qd_empty[it & 1]);
            if (ns == NS - 1) umma_commit(&kv_empty[iu & 1]);
          }
          ++gr;
          idle = 0;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax groups: group g = key columns [64 g, 64 g + 64) of the tile
    const int grp = (warp - 2) >> 2, qd = warp & 3;
    const int row = qd * 32 + lane;             // query row of the S / dP / dQ tiles = key row of dK / dV
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    float* kb = s.kterm[warp - 2][0];
    float* kc = s.kterm[warp - 2][1];
    const uint32_t kb_s = smem_u32(kb) + grp * 256, kc_s = smem_u32(kc) + grp * 256;
    const uint64_t sc2 = pk2(p.scale2, p.scale2), scl2 = pk2(p.scale, p.scale);
    const uint32_t p_row = smem_u32(s.p) + grp * ATB_TILE + row * 128, ds_row = smem_u32(s.ds) + grp * ATB_TILE + row * 128;
    const uint32_t sw = row & 7;
    // the dQ partial of step n is drained in the middle of step n + 1 (or at the end of the unit)
    bool have_pend = false, pend_mine = false;
    uint32_t pend_ph = 0, pend_col = 0;
    float* pend_acc = nullptr;
    auto drain_dq = [&]() {
      mbar_wait(grad_full, pend_ph);
      tc_fence_after();
      if (pend_mine) {
        uint32_t o[32];
        tmem_ld32(tmem_base + lane_base + pend_col, o);
        tmem_ld_wait();
        if (pend_acc != nullptr) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(pend_acc + 4 * e), "f"(__uint_as_float(o[4 * e])),
                         "f"(__uint_as_float(o[4 * e + 1])), "f"(__uint_as_float(o[4 * e + 2])), "f"(__uint_as_float(o[4 * e + 3]))
                         : "memory");
        }
      }
      tc_fence_before();
      have_pend = false;
    };
    auto drain_bf16 = [&](uint32_t col, bf16* dst) {
      uint32_t o[32];
      tmem_ld32(tmem_base + lane_base + col, o);
      tmem_ld_wait();
      if (dst != nullptr) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          *reinterpret_cast<uint4*>(dst + 8 * e) =
              make_uint4(pack_bf16(__uint_as_float(o[8 * e]), __uint_as_float(o[8 * e + 1])), pack_bf16(__uint_as_float(o[8 * e + 2]), __uint_as_float(o[8 * e + 3])),
                         pack_bf16(__uint_as_float(o[8 * e + 4]), __uint_as_float(o[8 * e + 5])), pack_bf16(__uint_as_float(o[8 * e + 6]), __uint_as_float(o[8 * e + 7])));
      }
    };
    int n = 0;   // steps of this CTA so far
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      int b, hg, j;
      decode(u, b, hg, j);
      // folded mask terms of the unit's 128 keys, warp-private (4 keys per lane): kb = additive term in base 2 (-inf beyond Sk),
      // kc = the key's index when it is cut for every earlier query (repeat-inc with m[j] = 1, or triangle), else -1
      __syncwarp();
      bool any_cut = false;
      {
        float b4[4], c4[4];
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const int kj = j * 128 + lane * 4 + uu;
          const float fm = (p.frame_mask != nullptr && kj < p.Sk) ? __ldg(p.frame_mask + (int64_t)b * p.frame_mask_stride + kj) : 0.f;
          const bool cut = ((p.flags & KIT_MASK_REPEAT_INC) && fm == 1.f) || (p.flags & KIT_MASK_TRIANGLE);
          b4[uu] = (kj >= p.Sk) ? -INFINITY : ((p.flags & KIT_MASK_KEYPAD_ADD) ? fm * ATB_LOG2E : 0.f);
          c4[uu] = cut ? (float)kj : -1.f;
          any_cut |= cut;
        }
        *reinterpret_cast<float4*>(kb + lane * 4) = make_float4(b4[0], b4[1], b4[2], b4[3]);
        *reinterpret_cast<float4*>(kc + lane * 4) = make_float4(c4[0], c4[1], c4[2], c4[3]);
      }
      const bool need_cut = __any_sync(0xffffffffu, any_cut);
      __syncwarp();
      for (int i = 0; i < p.q_tiles; ++i) {
#pragma unroll 1
        for (int hh = 0; hh < HP; ++hh, ++n) {
          const int h = hg * HP + hh, qi = i * 128 + row;
          const bool valid = qi < p.Sq;
          const int64_t ridx = ((int64_t)b * p.NH + h) * p.Sq + qi;
          const float lse2 = valid ? __ldg(p.lse + ridx) * ATB_LOG2E : INFINITY;
          const float dl = valid ? __ldg(p.delta + ridx) : 0.f;
          mbar_wait(sdp_full, n & 1);
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
          if (lane == 0) mbar_arrive(s_free);
          // P = 2^(S c + bias - lse2) (0 where masked);  dS = P (dP - delta) scale
          const uint64_t nl2 = pk2(-lse2, -lse2), nd2 = pk2(-dl, -dl);
          const float qif = (float)qi;
          uint32_t pp[32], dd[32];
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
              float v0, v1, a0, a1;
              up2(add2(fma2(pk2(x[c], x[c + 1]), sc2, pk2(__uint_as_float(kbv[e]), __uint_as_float(kbv[e + 1]))), nl2), v0, v1);
              if (need_cut) {
                v0 = (__uint_as_float(kcv[e]) > qif) ? -INFINITY : v0;
                v1 = (__uint_as_float(kcv[e + 1]) > qif) ? -INFINITY : v1;
              }
              const float p0 = atb_ex2(v0), p1 = atb_ex2(v1);
              up2(mul2(mul2(pk2(p0, p1), add2(pk2(dp[c], dp[c + 1]), nd2)), scl2), a0, a1);
              pp[c >> 1] = pack_bf16(p0, p1);
              dd[c >> 1] = pack_bf16(a0, a1);
            }
          }
          if (have_pend) drain_dq();   // the previous step's products have read the P / dS tiles; its dQ leaves for the workspace
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            sts128(p_row + ((uint32_t(c) ^ sw) << 4), pp[4 * c], pp[4 * c + 1], pp[4 * c + 2], pp[4 * c + 3]);
            sts128(ds_row + ((uint32_t(c) ^ sw) << 4), dd[4 * c], dd[4 * c + 1], dd[4 * c + 2], dd[4 * c + 3]);
          }
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full);
          have_pend = true;
          pend_ph = n & 1;
          pend_mine = (HP == 1) || (grp == hh);              // d = 64: each group drains 32 of the 64 columns; d = 32: group hh drains head hh
          pend_col = uint32_t(256 + hh * D + (HP == 1 ? grp * 32 : 0));
          pend_acc = valid ? p.dq_acc + ((int64_t)b * p.Sq + qi) * p.ld_acc + h * D + (HP == 1 ? grp * 32 : 0) : nullptr;
        }
      }
      // end of the unit: the last step's dQ, then dK / dV of the 128 keys (group g: columns [32 g, 32 g + 32) of the 64-column pack)
      drain_dq();
      const int kj = j * 128 + row;
      const bool okk = kj < p.Sk;
      drain_bf16(uint32_t(320 + grp * 32), okk ? p.dk + ((int64_t)b * p.Sk + kj) * p.ld_dk + hg * 64 + grp * 32 : nullptr);
      drain_bf16(uint32_t(384 + grp * 32), okk ? p.dv + ((int64_t)b * p.Sk + kj) * p.ld_dv + hg * 64 + grp * 32 : nullptr);
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host
bool attention_bwd_tc_supported(int NH, int Sq, int Sk, int d, const KitAttnMask* mask, const void* const* ptrs, const int64_t* lds, int n,
                                const float* dq_acc) {
  const char* e = getenv("KIT_ATTN_TC");   // KIT_ATTN_TC=0: the mma.sync streaming kernels (A/B measurements, tests)
  if (e != nullptr && e[0] == '0') return false;
  if ((d != 32 && d != 64) || NH % (64 / d) != 0 || Sk <= 64 || dq_acc == nullptr) return false;
  (void)Sq;
  if (mask != nullptr && mask->bias != nullptr) return false;
  for (int i = 0; i < n; ++i)
    if ((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) != 0 || (lds[i] * 2) % 16 != 0) return false;
  return true;
}

template <int D>
static int launch_tcb(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, const bf16* dout, int64_t ld_do,
                      const float* lse, const float* delta, float* dq_acc, bf16* dk, int64_t ld_dk, bf16* dv, int64_t ld_dv, int B, int NH,
                      int Sq, int Sk, const KitAttnMask* mask, cudaStream_t st) {
  constexpr int HP = 64 / D;
  constexpr int smem = (int)sizeof(AtbSmem) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    KIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  CUtensorMap tmQ, tmK, tmV, tmDO;
  int rc;
  // [B * S rows, NH * D columns]: box = 64 columns (one head pack) x 128 rows, 128B swizzle; rows past the end read as zeros
  if ((rc = make_tensor_map_2d(&tmQ, q, (uint64_t)NH * D, (uint64_t)B * Sq, (uint64_t)ldq * 2, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&tmK, k, (uint64_t)NH * D, (uint64_t)B * Sk, (uint64_t)ldk * 2, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&tmV, v, (uint64_t)NH * D, (uint64_t)B * Sk, (uint64_t)ldv * 2, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&tmDO, dout, (uint64_t)NH * D, (uint64_t)B * Sq, (uint64_t)ld_do * 2, 64, 128))) return rc;
  AtbParams p;
  p.B = B; p.NH = NH; p.Sq = Sq; p.Sk = Sk;
  p.q_tiles = (Sq + 127) / 128;
  p.k_tiles = (Sk + 127) / 128;
  const int64_t units64 = (int64_t)B * (NH / HP) * p.k_tiles;
  KIT_REQUIRE(units64 < (1ll << 31) / (p.q_tiles * HP), "attention backward: too many (batch, head pack, key tile) units");
  p.units = (int)units64;
  p.scale = rsqrtf((float)D);
  p.scale2 = p.scale * ATB_LOG2E;
  p.frame_mask = mask != nullptr ? mask->frame_mask : nullptr;
  p.frame_mask_stride = mask != nullptr ? mask->frame_mask_stride : 0;
  p.flags = (mask != nullptr && mask->frame_mask != nullptr) ? mask->flags : (mask != nullptr ? (mask->flags & KIT_MASK_TRIANGLE) : 0);
  p.lse = lse; p.delta = delta;
  p.dq_acc = dq_acc; p.ld_acc = (int64_t)NH * D;
  p.dk = dk; p.ld_dk = ld_dk; p.dv = dv; p.ld_dv = ld_dv;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    KIT_CHECK_CUDA(cudaGetDevice(&dev));
    KIT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.units < sms - sm_reserve() ? p.units : sms - sm_reserve());
  cfg.blockDim = dim3(64 + 256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  KIT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_bwd_tc_kernel<D>, tmQ, tmK, tmV, tmDO, p));
  return KIT_OK;
}

int attention_bwd_tc(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, const bf16* dout, int64_t ld_do,
                     const float* lse, const float* delta, float* dq_acc, bf16* dk, int64_t ld_dk, bf16* dv, int64_t ld_dv, int B, int NH,
                     int Sq, int Sk, int d, const KitAttnMask* mask, cudaStream_t st) {
  if (d == 32)
    return launch_tcb<32>(q, ldq, k, ldk, v, ldv, dout, ld_do, lse, delta, dq_acc, dk, ld_dk, dv, ld_dv, B, NH, Sq, Sk, mask, st);
  return launch_tcb<64>(q, ldq, k, ldk, v, ldv, dout, ld_do, lse, delta, dq_acc, dk, ld_dk, dv, ld_dv, B, NH, Sq, Sk, mask, st);
}

}  // namespace kit

// Attention forward on the 5th-generation tensor cores for longer sequences (more than one 64-key tile): softmax(Q K^T / sqrt(d)
// + mask) V with 128-query x 128-key score tiles, tcgen05.mma issued by one thread, accumulators in tensor memory, operands
// staged by TMA.  model.py:141-145 (nn.Transformer's scaled-dot-product attention) with the masks of A1_train.py:117-124
// synthesised in the kernel from the [B, T] frame mask, folded into two floats per key.
//
// A CTA works on units of (batch, 64 feature columns of the packed head dimension, query block) with TWO softmax groups:
// at d = 32 the two heads that share the 128B-swizzled [rows x 64] Q / K / V tiles (a head's 32-column half is selected by the
// start address of its K-major descriptors; 128 queries per unit), at d = 64 two 128-query tiles of one head (256 queries per
// unit).  Steps n = (key tile j, group g):
//
//     warp 1 (one thread):  S[n & 1] (TMEM, 128 columns) = Q_g K_j,g^T        2 or 4 MMAs 128 x 128 x 16
//     softmax group g:      one thread per query row: tcgen05.ld of the whole 128-key row, scale + mask, online max / sum in
//                           base 2, rescale of the running O_g (tcgen05.ld / st), P (bf16) -> shared memory in the K-major
//                           128B-swizzled layout the tensor core reads
//     warp 1:               O_g (TMEM, 64 columns) += P V_j                  8 MMAs 128 x 64 x 16, V tile as MN-major B operand
//                           (with two heads the full 64-column V tile is multiplied and each head keeps its own half)
//
// issued as QK(n + 1) before PV(n), so the tensor pipe computes the next score tile while a softmax group works on this one;
// with two heads the two groups (4 warps each) alternate.  Bound by the exponentials (one MUFU.EX2 per score), not by the
// tensor pipe.  Explicit additive bias tensors go through the mma.sync kernels of attention.cu.
#include "attention.cuh"
#include "gemm_sm100.cuh"

namespace kit {

constexpr int ATC_Q = 128, ATC_K = 128;
constexpr int ATC_TILE = 128 * 64 * 2;   // [128 rows x 64 columns] bf16 = 16 KB

struct AtcParams {
  int NH, Sq, Sk, n_ktiles, q_tiles, units;
  float scale2;              // scale * log2(e)
  const float* frame_mask;   // [B, T] or null
  int64_t frame_mask_stride;
  int flags;
  bf16* out;
  int64_t ldo;
  float* lse;                // [B, NH, Sq] or null
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <int D>
struct AtcSmem {
  static constexpr int QT = D == 64 ? 2 : 1;   // query tiles per unit
  uint8_t q[2][QT][ATC_TILE];
  uint8_t k[2][ATC_TILE];
  uint8_t v[2][ATC_TILE];
  uint8_t p[2][2 * ATC_TILE];   // [128 queries x 128 keys] bf16: two [128 x 64] k-blocks
  float kbias[8][2][ATC_K];     // per softmax warp: folded key terms of the current tile (bias * log2e, cut flag)
  uint64_t bars[20];
  uint32_t tmem_slot;
};

template <int D>
__global__ void __launch_bounds__(64 + 256, 1) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                             const __grid_constant__ CUtensorMap tmK,
                                                                             const __grid_constant__ CUtensorMap tmV,
                                                                             const AtcParams p) {
  constexpr int G = 2, HP = 64 / D, QT = D == 64 ? 2 : 1, KS = D / 16;   // groups, heads per pack, query tiles per unit
  extern __shared__ uint8_t smem_raw[];
  AtcSmem<D>& s = *reinterpret_cast<AtcSmem<D>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = &s.bars[0];     // [2]
  uint64_t* q_empty = &s.bars[2];    // [2] every Q K^T of the unit has been issued (tcgen05.commit)
  uint64_t* kv_full = &s.bars[4];    // [2]
  uint64_t* kv_empty = &s.bars[6];   // [2]
  uint64_t* s_full = &s.bars[8];     // [2] score tile written (tcgen05.commit)
  uint64_t* s_free = &s.bars[10];    // [2] score tile read by its softmax group (4 warps)
  uint64_t* p_full = &s.bars[12];    // [2] P in shared memory, O rescaled (4 warps)
  uint64_t* pv_done = &s.bars[14];   // [2] P V accumulated (tcgen05.commit): P buffer free, O stable

  // Persistent over (batch, head pack, 128-query tile) units: TMEM, barriers and the K / V ring carry on from unit to unit, the
  // next unit's Q tile is prefetched into the other buffer.  Counters: iu = unit number of this CTA, jt = key tiles so far,
  // n = steps so far (slot = n & 1).
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NS = p.n_ktiles * G;   // steps per unit
  auto decode = [&](int u, int& b, int& hg, int& q0) {
    const int qt = u % p.q_tiles, rest = u / p.q_tiles;
    q0 = qt * ATC_Q * QT;
    hg = rest % (p.NH / HP);
    b = rest / (p.NH / HP);
  };

  if (warp == 0) {
    pdl_launch_dependents();
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&q_full[i], 1);
        mbar_init(&q_empty[i], 1);
        mbar_init(&kv_full[i], 1);
        mbar_init(&kv_empty[i], 1);
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
  // TMEM columns: S[slot] at slot * 128, O[g] at 256 + g * 64
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int iu = 0, jt = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++iu) {
        int b, hg, q0;
        decode(u, b, hg, q0);
        mbar_wait(&q_empty[iu & 1], ((iu >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[iu & 1], QT * ATC_TILE);
#pragma unroll
        for (int t = 0; t < QT; ++t) tma_load_2d(s.q[iu & 1][t], &tmQ, &q_full[iu & 1], hg * 64, b * p.Sq + q0 + t * ATC_Q);
        for (int j = 0; j < p.n_ktiles; ++j, ++jt) {
          const int st = jt & 1;
          mbar_wait(&kv_empty[st], ((jt >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[st], 2 * ATC_TILE);
          tma_load_2d(s.k[st], &tmK, &kv_full[st], hg * 64, b * p.Sk + j * ATC_K);
          tma_load_2d(s.v[st], &tmV, &kv_full[st], hg * 64, b * p.Sk + j * ATC_K);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // The whole warp runs the loop on warp-uniform values and issues every tcgen05 instruction under an elect.sync predicate
    // (operands stay in uniform registers, consecutive UTCHMMA are back to back -- attention_t64.cu), and the barriers are PROBED
    // (mbarrier.test_wait): try_wait may suspend the thread for a system-dependent time, which stalled the ready queue behind
    // the other one.
    {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, false, true);   // B = V tile [keys][64 columns]: MN-major
      uint32_t el;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
      const bool leader = el != 0;
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      auto probe = [&](uint64_t* bar, uint32_t parity) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        return ok != 0;
      };
      auto uniform = [&](bool v) { return __shfl_sync(0xffffffffu, (int)v, 0) != 0; };
      auto issue_pv = [&](int m, int j, int jt, int g) {   // O_g += P[m & 1] V_j   (m = global step)
        const int sl = m & 1;
        tc_fence_after();
        const uint32_t p_base = smem_u32(s.p[sl]), v_base = smem_u32(s.v[jt & 1]);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {   // 16 keys per MMA
          const uint64_t adesc = make_smem_desc_sw128(p_base + (kk >> 2) * ATC_TILE + (kk & 3) * 32, 0, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(v_base + kk * 16 * 128, 64 * 128, 1024);
          if (leader) umma_bf16(tb + 256 + g * 64, adesc, bdesc, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        }
        if (leader) {
          umma_commit(&pv_done[sl]);
          if (g == G - 1) umma_commit(&kv_empty[jt & 1]);   // every MMA that reads this K / V stage has been issued
        }
      };
      // Two in-order queues, issued as their inputs become ready: Q K^T of step n needs its score slot drained (s_free, early in
      // step n - 2 of the same group) and the K / V stage; P V of step m needs P (p_full, the end of softmax step m).  The score
      // product of step n + 2 is issued as soon as step n's scores are in registers, so it is complete long before its group asks.
      const int n_units = (p.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int total = n_units * NS;
      int qk = 0, pv = 0;
      uint32_t idle = 0;
      while (pv < total) {
        if (++idle > (1u << 27)) {   // a protocol bug becomes a trap instead of a hang
          if (leader) printf("kit: attn_fwd_tc MMA queue stalled (block %d, qk %d, pv %d of %d)\n", blockIdx.x, qk, pv, total);
          __trap();
        }
        if (qk < total) {
          const int iuq = qk / NS, ns = qk % NS, j = ns / G, g = ns % G, sl = qk & 1, jt = iuq * p.n_ktiles + j;
          bool ready = probe(&s_free[sl], ((qk >> 1) & 1) ^ 1);
          if (ready && ns == 0) ready = probe(&q_full[iuq & 1], (iuq >> 1) & 1);
          if (ready && g == 0) ready = probe(&kv_full[jt & 1], (jt >> 1) & 1);
          if (uniform(ready)) {
            tc_fence_after();
            const uint32_t q_buf = smem_u32(s.q[iuq & 1][0]), k_base = smem_u32(s.k[jt & 1]);
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
              // d = 32: group = head half of the shared tiles; d = 64: group = query tile, the whole K tile
              const uint32_t a_off = (D == 64) ? g * ATC_TILE : g * D * 2, b_off = (D == 64) ? 0 : g * D * 2;
              const uint64_t adesc = make_smem_desc_sw128(q_buf + a_off + kk * 32, 0, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(k_base + b_off + kk * 32, 0, 1024);
              if (leader) umma_bf16(tb + sl * 128, adesc, bdesc, idesc_qk, kk > 0 ? 1u : 0u);
            }
            if (leader) {
              umma_commit(&s_full[sl]);
              if (ns == NS - 1) umma_commit(&q_empty[iuq & 1]);   // the producer may refill this Q buffer
            }
            ++qk;
            idle = 0;
          }
        }
        if (pv < qk) {
          const int iup = pv / NS, ns = pv % NS, j = ns / G, g = ns % G, jt = iup * p.n_ktiles + j;
          if (uniform(probe(&p_full[pv & 1], (pv >> 1) & 1))) {
            issue_pv(pv, j, jt, g);
            ++pv;
            idle = 0;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax groups: group g = head g of the 64-column pack
    const int grp = (warp - 2) >> 2;          // 0 / 1
    const int qd = warp & 3;                  // TMEM lane quadrant of this warp
    const int row = qd * 32 + lane;           // query row of the tile
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    float* kb = s.kbias[warp - 2][0];
    float* kc = s.kbias[warp - 2][1];
    int n0 = 0;   // global steps before this unit
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, n0 += NS) {
    int b, hg, q0;
    decode(u, b, hg, q0);
    const int qi = q0 + (D == 64 ? grp * ATC_Q : 0) + row;
    const int o_col = 256 + grp * 64 + (D == 64 ? 0 : grp * D);   // this group's D accumulator columns
    float m_run = -INFINITY, l_run = 0.f;
    float fm_next[4];
    auto load_fm = [&](int jn) {
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) {
        const int kj = jn * ATC_K + lane * 4 + uu;
        fm_next[uu] = (p.frame_mask != nullptr && kj < p.Sk) ? __ldg(p.frame_mask + (int64_t)b * p.frame_mask_stride + kj) : 0.f;
      }
    };
    load_fm(0);
    for (int j = 0; j < p.n_ktiles; ++j) {
      const int n = n0 + j * G + grp, sl = n & 1;
      // folded mask terms of the 128 keys of this tile, warp-private (4 keys per lane)
      __syncwarp();
      // kb: additive term in base 2 (-inf for keys beyond the sequence); kc: the key's own index when it is cut for every earlier
      // query (repeat-inc / triangle), else -1 -- a score is masked iff kc > query index
      // A tile entirely behind this warp's last query has its cut keys masked for every row: folded into kb (-inf), no compare;
      // a tile entirely at or before the warp's first query cuts nothing.  Only tiles that cross the diagonal compare per score.
      const bool tile_after = j * ATC_K > qi - lane + 31, tile_before = j * ATC_K + ATC_K - 1 <= qi - lane;
      bool any_cut = false;
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) {
        const int kl = lane * 4 + uu, kj = j * ATC_K + kl;
        const float fm = fm_next[uu];
        const bool cut = ((p.flags & KIT_MASK_REPEAT_INC) && fm == 1.f) || (p.flags & KIT_MASK_TRIANGLE);
        kb[kl] = (kj >= p.Sk || (cut && tile_after)) ? -INFINITY : ((p.flags & KIT_MASK_KEYPAD_ADD) ? fm * 1.4426950408889634f : 0.f);
        kc[kl] = cut ? (float)kj : -1.f;
        any_cut |= cut;
      }
      const bool need_cut = __any_sync(0xffffffffu, any_cut) && !tile_after && !tile_before;
      if (j + 1 < p.n_ktiles) load_fm(j + 1);   // the next tile's frame mask travels during this step
      __syncwarp();
      mbar_wait(&s_full[sl], (n >> 1) & 1);
      tc_fence_after();
      float x[128];
      uint32_t* xr = reinterpret_cast<uint32_t*>(x);
      const uint32_t s_addr = tmem_base + lane_base + uint32_t(sl * 128);
      tmem_ld32(s_addr, xr);
      tmem_ld32(s_addr + 32, xr + 32);
      tmem_ld_wait();
      tmem_ld32(s_addr + 64, xr + 64);    // the second half of the row travels while the first is scaled and masked
      tmem_ld32(s_addr + 96, xr + 96);
      // scale + mask (base 2), row maximum
      float mx4[8] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY};   // independent chains
      const uint32_t kb_s = smem_u32(kb), kc_s = smem_u32(kc);
      const float qif = (float)qi;
      const uint64_t sc2 = pk2(p.scale2, p.scale2);
      auto half = [&](int h0) {
        if (need_cut) {
#pragma unroll
          for (int c4 = h0; c4 < h0 + 16; ++c4) {
            const uint4 kbw = lds128(kb_s + 16 * c4), kcw = lds128(kc_s + 16 * c4);
            const uint32_t kbv[4] = {kbw.x, kbw.y, kbw.z, kbw.w}, kcv[4] = {kcw.x, kcw.y, kcw.z, kcw.w};
#pragma unroll
            for (int u = 0; u < 4; u += 2) {   // packed fp32 pairs for the scale + bias
              const int c = 4 * c4 + u;
              float v0, v1;
              up2(fma2(pk2(x[c], x[c + 1]), sc2, pk2(__uint_as_float(kbv[u]), __uint_as_float(kbv[u + 1]))), v0, v1);
              v0 = (__uint_as_float(kcv[u]) > qif) ? -INFINITY : v0;
              v1 = (__uint_as_float(kcv[u + 1]) > qif) ? -INFINITY : v1;
              x[c] = v0;
              x[c + 1] = v1;
              mx4[u + 4 * (c4 & 1)] = fmaxf(mx4[u + 4 * (c4 & 1)], v0);
              mx4[u + 1 + 4 * (c4 & 1)] = fmaxf(mx4[u + 1 + 4 * (c4 & 1)], v1);
            }
          }
        } else {
#pragma unroll
          for (int c4 = h0; c4 < h0 + 16; ++c4) {
            const uint4 kbw = lds128(kb_s + 16 * c4);
            const uint32_t kbv[4] = {kbw.x, kbw.y, kbw.z, kbw.w};
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
              const int c = 4 * c4 + u;
              float v0, v1;
              up2(fma2(pk2(x[c], x[c + 1]), sc2, pk2(__uint_as_float(kbv[u]), __uint_as_float(kbv[u + 1]))), v0, v1);
              x[c] = v0;
              x[c + 1] = v1;
              mx4[u + 4 * (c4 & 1)] = fmaxf(mx4[u + 4 * (c4 & 1)], v0);
              mx4[u + 1 + 4 * (c4 & 1)] = fmaxf(mx4[u + 1 + 4 * (c4 & 1)], v1);
            }
          }
        }
      };
      half(0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[sl]);
      half(16);
      const float mx = fmaxf(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])), fmaxf(fmaxf(mx4[4], mx4[5]), fmaxf(mx4[6], mx4[7])));
      // Lazy rescale: a stale reference is as good as the true row maximum while x - m_ref stays small (P <= 2^8 is exact enough
      // in bf16, l and O are fp32), so the running sum and O_g (a tcgen05.ld / st round trip) are rescaled only when some row of
      // the warp has a new maximum more than 2^8 above its reference -- after the first tiles that is rare.
      const bool any_grow = __any_sync(0xffffffffu, mx > m_run + 8.f);
      const float m_new = any_grow ? fmaxf(m_run, mx) : m_run;
      const float m_ref = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = ex2f(m_run - m_ref);
      uint64_t rs2[4] = {pk2(0.f, 0.f), pk2(0.f, 0.f), pk2(0.f, 0.f), pk2(0.f, 0.f)};
      const uint64_t nm2 = pk2(-m_ref, -m_ref);
      uint32_t* pk = reinterpret_cast<uint32_t*>(x);   // the packed probabilities overwrite the scores they came from
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float a0, a1;
        up2(add2(pk2(x[2 * c], x[2 * c + 1]), nm2), a0, a1);
        const float p0 = ex2f(a0), p1 = ex2f(a1);
        rs2[c & 3] = add2(rs2[c & 3], pk2(p0, p1));
        pk[c] = pack_bf16(p0, p1);
      }
      float r0, r1, r2, r3, r4, r5, r6, r7;
      up2(rs2[0], r0, r1); up2(rs2[1], r2, r3); up2(rs2[2], r4, r5); up2(rs2[3], r6, r7);
      const float rs = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
      l_run = l_run * corr + rs;
      m_run = m_new;
      // the previous step of this slot has been accumulated (P buffer free); the previous tile of this head too (O stable)
      if (n >= 2) mbar_wait(&pv_done[sl], ((n - 2) >> 1) & 1);
      tc_fence_after();
      if (j > 0 && any_grow) {   // O_g *= corr (this head's D columns)
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          uint32_t o[32];
          const uint32_t addr = tmem_base + lane_base + uint32_t(o_col + c * 32);
          tmem_ld32(addr, o);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 32; ++u) o[u] = __float_as_uint(__uint_as_float(o[u]) * corr);
          tmem_st32(addr, o);
        }
        tmem_st_wait();
      }
      const uint32_t p_row = smem_u32(s.p[sl]) + row * 128;
      const uint32_t sw = row & 7;
#pragma unroll
      for (int c = 0; c < 16; ++c)   // 16-byte chunk c of the 256-byte row: k-block c >> 3, chunk (c & 7) ^ (row & 7)
        sts128(p_row + (c >> 3) * ATC_TILE + ((uint32_t(c & 7) ^ sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sl]);
    }
    // ---- O_g / l -> out, log-sum-exp
    const int n_last = n0 + (p.n_ktiles - 1) * G + grp;
    mbar_wait(&pv_done[n_last & 1], (n_last >> 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const int h = hg * HP + (D == 64 ? 0 : grp);
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t o[32];
      tmem_ld32(tmem_base + lane_base + uint32_t(o_col + c * 32), o);
      tmem_ld_wait();
      if (qi < p.Sq) {
        bf16* dst = p.out + ((int64_t)b * p.Sq + qi) * p.ldo + h * D + c * 32;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[8 * u + e]) * inv;
          store8(dst + 8 * u, f);
        }
      }
    }
    if (p.lse != nullptr && qi < p.Sq) p.lse[((int64_t)b * p.NH + h) * p.Sq + qi] = (m_run + log2f(l_run)) * 0.6931471805599453f;
    tc_fence_before();
    }   // units
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

bool attention_fwd_tc_supported(int64_t ldq, int64_t ldk, int64_t ldv, int NH, int Sq, int Sk, int d, const KitAttnMask* mask,
                                const void* q, const void* k, const void* v) {
  const char* e = getenv("KIT_ATTN_TC");   // KIT_ATTN_TC=0: the mma.sync streaming kernel (A/B measurements, tests)
  if (e != nullptr && e[0] == '0') return false;
  if (d != 32 && d != 64) return false;
  if (NH % (64 / d) != 0) return false;
  (void)Sq;
  if (Sk <= 64 || (mask != nullptr && mask->bias != nullptr)) return false;
  auto ok = [](const void* ptr, int64_t ld) { return (reinterpret_cast<uintptr_t>(ptr) & 127) == 0 && (ld * 2) % 16 == 0; };
  return ok(q, ldq) && ok(k, ldk) && ok(v, ldv);
}

template <int D>
static int launch_tc(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, bf16* out, int64_t ldo,
                     float* lse, int B, int NH, int Sq, int Sk, const KitAttnMask* mask, cudaStream_t st) {
  constexpr int HP = 64 / D, QT = D == 64 ? 2 : 1;
  constexpr int smem = (int)sizeof(AtcSmem<D>) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    KIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  // [B * S rows, NH * D columns]: box = 64 columns (one head pack) x 128 rows, 128B swizzle
  if ((rc = make_tensor_map_2d(&tmQ, q, (uint64_t)NH * D, (uint64_t)B * Sq, (uint64_t)ldq * 2, 64, ATC_Q))) return rc;
  if ((rc = make_tensor_map_2d(&tmK, k, (uint64_t)NH * D, (uint64_t)B * Sk, (uint64_t)ldk * 2, 64, ATC_K))) return rc;
  if ((rc = make_tensor_map_2d(&tmV, v, (uint64_t)NH * D, (uint64_t)B * Sk, (uint64_t)ldv * 2, 64, ATC_K))) return rc;
  AtcParams p;
  p.NH = NH; p.Sq = Sq; p.Sk = Sk; p.n_ktiles = (Sk + ATC_K - 1) / ATC_K;
  p.q_tiles = (Sq + ATC_Q * QT - 1) / (ATC_Q * QT);   // query blocks of 128 (d = 32) or 256 (d = 64) rows
  const int64_t units64 = (int64_t)B * (NH / HP) * p.q_tiles;
  KIT_REQUIRE(units64 < (1ll << 31), "attention forward: too many (batch, head pack, query tile) units");
  p.units = (int)units64;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    KIT_CHECK_CUDA(cudaGetDevice(&dev));
    KIT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  p.scale2 = rsqrtf((float)D) * 1.4426950408889634f;
  p.frame_mask = mask != nullptr ? mask->frame_mask : nullptr;
  p.frame_mask_stride = mask != nullptr ? mask->frame_mask_stride : 0;
  p.flags = (mask != nullptr && mask->frame_mask != nullptr) ? mask->flags : (mask != nullptr ? (mask->flags & KIT_MASK_TRIANGLE) : 0);
  p.out = out; p.ldo = ldo; p.lse = lse;
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
  KIT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_fwd_tc_kernel<D>, tmQ, tmK, tmV, p));
  return KIT_OK;
}

int attention_fwd_tc(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, bf16* out, int64_t ldo,
                     float* lse, int B, int NH, int Sq, int Sk, int d, const KitAttnMask* mask, cudaStream_t st) {
  if (d == 32) return launch_tc<32>(q, ldq, k, ldk, v, ldv, out, ldo, lse, B, NH, Sq, Sk, mask, st);
  return launch_tc<64>(q, ldq, k, ldk, v, ldv, out, ldo, lse, B, NH, Sq, Sk, mask, st);
}

}  // namespace kit

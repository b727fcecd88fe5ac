// Fused feed-forward block of a post-norm transformer layer (torch/nn/modules/transformer.py:956-959, 1147-1153):
//
//     s = x + linear2(gelu(linear1(x)));   y = LayerNorm(s)
//
// in ONE kernel for the model width H = 256: the [rows, FF] hidden activation never makes the HBM round trip between the two
// GEMMs.  A CTA pair owns 256 token rows (128 per CTA, tcgen05.mma.cta_group::2).  The x tile [128 x 256] stays in shared
// memory for the whole item: it is the A operand of GEMM1 and the residual of the final epilogue.  The hidden dimension is
// walked in chunks of 128 columns:
//
//     GEMM1(c):  acc1[c & 1] (TMEM, 128 columns, double-buffered) = x . W1[c]^T                 K = 256
//     EPI1(c):   16 epilogue warps: tcgen05.ld -> + b1 -> GELU -> bf16 -> shared memory in the 128B-swizzled K-major layout
//                the tensor core reads (the chunk is the A operand of GEMM2); when the step trains, the same tiles (and the
//                pre-activation z) leave by TMA store for the backward pass
//     GEMM2(c):  acc2 (TMEM, 256 columns) += h[c] . W2[:, c]^T                                   K = 128
//
// issued as GEMM1(c+1) before GEMM2(c) so that EPI1(c) overlaps GEMM1(c+1) on the tensor pipe.  Both weight streams come
// through one ring of 16 KB slots (each CTA stages HALF of every B tile).  The final epilogue is the EPI_ADD_LN epilogue of
// gemm_sm100.cuh (bias + residual, bf16 sum out, two-pass row statistics across the four column-group warps, LayerNorm out).
// Persistent over 256-row items.
//
// BWD = true is the input-gradient pass of the same block with the same skeleton:
//
//     dz = (g . W2) * gelu'(z)   (GEMM1 against W2^T chunks, z chunk prefetched by TMA, dz stored for the weight gradients)
//     dx = dz . W1 + g           (GEMM2 against W1^T chunks; the residual g is the resident A tile)
//
// i.e. x := g (the gradient w.r.t. the pre-norm sum), W1 := W2^T [FF, H], W2 := W1^T [H, FF], no biases, no LayerNorm.
#pragma once
#include "gemm_sm100.cuh"

namespace kit {

constexpr int FFN_H = 256;
constexpr int FFN_FC = 128;
// FFN_XT (an experiment kept behind KIT_FFN_XT, default off -- it measured no faster, see gemm.cu ffn_xt_mask): the x tile is
// copied into tensor memory once per item (tcgen05.st by the epilogue warps, 128 columns of bf16 pairs) and GEMM1 takes its A
// operand from there.  tools/umma_rate.py: an M = 128, N = 128 tcgen05.mma costs 112 cycles with both
// operands in shared memory (48 of them the A read, which does not overlap the math) and 79 with A in tensor memory.  The 512
// columns then hold acc1 (128, single-buffered: GEMM1(c + 1) waits for EPI1(c)'s tcgen05.ld, which GEMM2(c - 1) covers on the
// pipe), x (128) and acc2 (256).
// A template parameter (KIT_FFN_XT=0 at plan time selects the shared-memory A operand: A/B measurements).
template <bool BWD> constexpr int ffn_ring() { return BWD ? 3 : 5; }    // 16 KB weight slots
template <bool BWD> constexpr int ffn_zbufs() { return BWD ? 2 : 1; }   // z chunk buffers: TMA-in (double-buffered) / staging out
constexpr int FFN_SLOT = 16384;
constexpr int FFN_X_BYTES = 65536;   // [128 x 256] bf16: four [128 x 64] k-blocks
constexpr int FFN_HC_BYTES = 32768;  // [128 x 128] bf16: two [128 x 64] k-blocks
constexpr int FFN_STATS = 4096;
constexpr int FFN_EPI_WARPS = 16;
constexpr int FFN_THREADS = 64 + 32 * FFN_EPI_WARPS;
template <bool BWD> constexpr int ffn_smem() {
  return FFN_X_BYTES + ffn_ring<BWD>() * FFN_SLOT + (1 + ffn_zbufs<BWD>()) * FFN_HC_BYTES + FFN_STATS + 1024 + 1024;
}

struct FfnParams {
  int M, FF, n_items;
  const float* b1;
  const float* b2;
  const float* ln_gamma;
  const float* ln_beta;
  float* ln_mean;
  float* ln_rstd;
  float ln_eps;
  int store_zh;   // forward, training: the pre-activation z and the activation h leave for the backward pass
  LnBwdArgs lnb;  // BWD: the LayerNorm whose backward follows dx (s == null: none)
  long long* trace;   // experiments only (KIT_FFN_TRACE): clock64 marks of CTA 0, first item; see kit_ffn_trace_read
};
struct FfnPlan {
  CUtensorMap tmX, tmW1, tmW2, tmZ, tmHh, tmS, tmY;
  FfnParams p;
  int grid;
  int bwd;
  int xt;   // x tile as a tensor-memory A operand (FFN_XT)
};

#ifdef KIT_FFN_IMPL   // the kernel is compiled into gemm.cu only
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// acquire at cluster scope: the arrivals come from the peer CTA after it wrote ITS shared memory
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("kit: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void umma_bf16_ts_cg2(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
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
__device__ __forceinline__ void mbar_arrive_cluster_release(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

template <bool BWD, bool FFN_XT>
__global__ void __launch_bounds__(FFN_THREADS, 1) ffn_kernel(const __grid_constant__ CUtensorMap tmX,
                                                             const __grid_constant__ CUtensorMap tmW1,
                                                             const __grid_constant__ CUtensorMap tmW2,
                                                             const __grid_constant__ CUtensorMap tmZ,
                                                             const __grid_constant__ CUtensorMap tmHh,
                                                             const __grid_constant__ CUtensorMap tmS,
                                                             const __grid_constant__ CUtensorMap tmY, const FfnParams p) {
  constexpr int FFN_RING = ffn_ring<BWD>();
  constexpr int ZBUFS = ffn_zbufs<BWD>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* x_s = smem;
  uint8_t* ring = x_s + FFN_X_BYTES;
  uint8_t* h_s = ring + FFN_RING * FFN_SLOT;
  uint8_t* z_s = h_s + FFN_HC_BYTES;
  float2* ln_stats = reinterpret_cast<float2*>(z_s + ZBUFS * FFN_HC_BYTES);
  uint64_t* ring_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ln_stats) + FFN_STATS);
  uint64_t* ring_empty = ring_full + FFN_RING;
  uint64_t* x_full = ring_empty + FFN_RING;
  uint64_t* x_empty = x_full + 1;
  uint64_t* acc1_full = x_empty + 1;    // [2]
  uint64_t* acc1_empty = acc1_full + 2; // [2]
  uint64_t* h_full = acc1_empty + 2;    // [2]
  uint64_t* h_empty = h_full + 2;       // [2]
  uint64_t* acc2_full = h_empty + 2;
  uint64_t* acc2_empty = acc2_full + 1;
  uint64_t* z_full = acc2_empty + 1;    // [2]  BWD: the z chunk has landed (local)
  uint64_t* z_empty = z_full + 2;       // [2]  BWD: the epilogue warps have read it (local, 16 arrivals)
  uint64_t* s_bars = z_empty + 2;       // [16][2] BWD + LayerNorm backward: the saved-sum tiles (local)
  uint64_t* xt_full = s_bars + 2 * FFN_EPI_WARPS;   // FFN_XT: both CTAs' epilogue warps have copied x into tensor memory (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xt_full + 1);
  constexpr int N_BARS = 2 * FFN_RING + 16 + 2 * FFN_EPI_WARPS + 1;
  const bool store_h = BWD || p.store_zh, store_z = !BWD && p.store_zh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int first_item = blockIdx.x >> 1, item_stride = gridDim.x >> 1;
  const int NC = p.FF / FFN_FC;
  auto mark = [&](int slot) {
    if (p.trace != nullptr && blockIdx.x < 2 && lane == 0 && slot < 128) p.trace[blockIdx.x * 128 + slot] = clock64();
  };
  if (warp == 0) mark(0);

  if (warp == 0) {
    pdl_launch_dependents();
    if (lane == 0) {
      tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
      tma_prefetch_desc(&tmS);
      if (!BWD || p.lnb.s != nullptr) tma_prefetch_desc(&tmY);
      if (store_h) tma_prefetch_desc(&tmHh);
      if (BWD || store_z) tma_prefetch_desc(&tmZ);
    }
    for (int i = lane; i < N_BARS; i += 32) {
      uint64_t* b = &ring_full[i];
      uint32_t count = 1;
      if (b == x_empty || b == z_empty || b == z_empty + 1) count = FFN_EPI_WARPS;
      else if (b == acc1_empty || b == acc1_empty + 1 || b == h_full || b == h_full + 1 || b == acc2_empty || b == xt_full) count = 2 * FFN_EPI_WARPS;
      mbar_init(b, count);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc_cg2<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_setup();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: the x tile, then the two weight streams
    if (lane == 0) {
      uint32_t cnt = 0, it = 0, zc = 0;
      auto slot_acquire = [&](uint32_t& s) {
        s = cnt % FFN_RING;
        mbar_wait(&ring_empty[s], ((cnt / FFN_RING) & 1) ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&ring_full[s], 2 * FFN_SLOT);
        ++cnt;
      };
      auto load_w1 = [&](int c) {   // W1 rows [c*128 + rank*64, +64), two 64-wide k-blocks per slot
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t s;
          slot_acquire(s);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            tma_load_2d_cg2(ring + s * FFN_SLOT + kk * 8192, &tmW1, &ring_full[s], (2 * j + kk) * 64, c * FFN_FC + rank * 64);
        }
      };
      auto load_w2 = [&](int c) {   // W2 rows [rank*128, +128), k columns [c*128 + j*64, +64)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t s;
          slot_acquire(s);
          tma_load_2d_cg2(ring + s * FFN_SLOT, &tmW2, &ring_full[s], c * FFN_FC + j * 64, rank * 128);
        }
      };
      for (int item = first_item; item < p.n_items; item += item_stride, ++it) {
        const int m0 = (item * 2 + rank) * 128;
        mbar_wait(x_empty, (it & 1) ^ 1);   // the final epilogue of the previous item has read the residual
        if (it == 0) pdl_wait();
        if (FFN_XT) {   // each CTA's epilogue warps wait for THEIR tile (local barrier) before copying it into tensor memory
          mbar_arrive_expect_tx(x_full, FFN_X_BYTES);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) tma_load_2d(x_s + kb * 16384, &tmX, x_full, kb * 64, m0);
        } else {
          if (rank == 0) mbar_arrive_expect_tx(x_full, 2 * FFN_X_BYTES);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) tma_load_2d_cg2(x_s + kb * 16384, &tmX, x_full, kb * 64, m0);
        }
        auto load_z = [&](int c) {   // BWD: z[m0 .. +128, c*128 .. +128] -> z_s[zc & 1], two [128 x 64] k-block tiles
          const uint32_t zb = zc & 1;
          mbar_wait(&z_empty[zb], ((zc >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&z_full[zb], FFN_HC_BYTES);
          tma_load_2d(z_s + zb * FFN_HC_BYTES, &tmZ, &z_full[zb], c * FFN_FC, m0);
          tma_load_2d(z_s + zb * FFN_HC_BYTES + 16384, &tmZ, &z_full[zb], c * FFN_FC + 64, m0);
          ++zc;
        };
        if (BWD) {
          load_z(0);
          if (NC > 1) load_z(1);
        }
        load_w1(0);
        for (int c = 0; c < NC; ++c) {
          if (c + 1 < NC) load_w1(c + 1);
          load_w2(c);
          if (BWD && c + 2 < NC) load_z(c + 2);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    // The whole warp runs the loop on warp-uniform values and every tcgen05 instruction is issued under an elect.sync
    // predicate: descriptors and addresses then stay in uniform registers and consecutive UTCHMMA are back to back.  Inside
    // `if (lane == 0)` each operand went through R2UR and an ELECT / BRA.U.ANY loop (~120 cycles per MMA, tools/umma_rate.py)
    // -- more than an N = 128 instruction takes on the tensor pipe.
    if (__shfl_sync(0xffffffffu, rank, 0) == 0) {
      constexpr uint32_t idesc1 = make_idesc_bf16(256, FFN_FC, false, false);
      constexpr uint32_t idesc2 = make_idesc_bf16(256, FFN_H, false, false);
      uint32_t el;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
      const bool leader = el != 0;
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t x_base = smem_u32(x_s), h_base = smem_u32(h_s), ring_base = smem_u32(ring);
      uint32_t cnt = 0, gc = 0, hc = 0, it = 0;
      auto slot_wait = [&](uint32_t& s) {
        s = cnt % FFN_RING;
        mbar_wait(&ring_full[s], (cnt / FFN_RING) & 1);
        tc_fence_after();
        ++cnt;
      };
      auto gemm2 = [&](int c) {
        if (c == 0) {
          mbar_wait(acc2_empty, (it & 1) ^ 1);   // the final epilogue of the previous item has drained acc2
          tc_fence_after();
        }
        mbar_wait_cluster(h_full, hc & 1);   // both CTAs' epilogue warps have written (and fenced) their half of h[c]
        tc_fence_after();
        if (it == 0) mark(17 + c);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t s;
          slot_wait(s);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(h_base + j * 16384 + k * 32, 0, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(ring_base + s * FFN_SLOT + k * 32, 0, 1024);
            if (leader) umma_bf16_cg2(tb + 256, adesc, bdesc, idesc2, (c > 0 || j > 0 || k > 0) ? 1u : 0u);
          }
          if (leader) umma_commit_cg2(&ring_empty[s], (uint16_t)3);
        }
        if (leader) umma_commit_cg2(h_empty, (uint16_t)3);
        ++hc;
      };
      for (int item = first_item; item < p.n_items; item += item_stride, ++it) {
        if (FFN_XT) mbar_wait_cluster(xt_full, it & 1); else mbar_wait(x_full, it & 1);
        tc_fence_after();
        for (int c = 0; c < NC; ++c, ++gc) {
          const uint32_t b = FFN_XT ? 0u : (gc & 1);
          mbar_wait(&acc1_empty[b], (FFN_XT ? (gc & 1) : ((gc >> 1) & 1)) ^ 1);   // EPI1 two chunks ago (FFN_XT: of the previous chunk) has drained this accumulator
          tc_fence_after();
          if (it == 0) mark(1 + c);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t s;
            slot_wait(s);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int kb = 2 * j + kk;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t adesc = make_smem_desc_sw128(x_base + kb * 16384 + k * 32, 0, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(ring_base + s * FFN_SLOT + kk * 8192 + k * 32, 0, 1024);
                if (leader) {
                  if (FFN_XT) umma_bf16_ts_cg2(tb, tb + FFN_FC + kb * 32 + k * 8, bdesc, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
                  else umma_bf16_cg2(tb + b * FFN_FC, adesc, bdesc, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
                }
              }
            }
            if (leader) umma_commit_cg2(&ring_empty[s], (uint16_t)3);
          }
          if (leader) umma_commit_cg2(&acc1_full[b], (uint16_t)3);
          if (c >= 1) gemm2(c - 1);
        }
        gemm2(NC - 1);
        if (leader) umma_commit_cg2(acc2_full, (uint16_t)3);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;            // TMEM lane quadrant: rows [32q, 32q + 32) of this CTA's tile
    const int cg = (warp - 2) >> 2;    // EPI1: 32-column group of the chunk; final epilogue: 64-column group of the row
    const int row = q * 32 + lane;
    const uint32_t sw = row & 7;
    const int pair_bar = 1 + q * 2 + (cg >> 1);     // the two warps that fill one [32 x 64] tile of h / z
    const bool issuer = (cg & 1) == 0;
    const uint32_t hz_off = (cg >> 1) * 16384 + row * 128;   // this lane's row in the chunk's k-block tile
    const uint32_t tile_off = (cg >> 1) * 16384 + q * 4096;  // the [32 x 64] tile this warp pair fills
    const uint32_t h_u32 = smem_u32(h_s), z_u32 = smem_u32(z_s), x_u32 = smem_u32(x_s);
    uint32_t gc = 0, it = 0;
    for (int item = first_item; item < p.n_items; item += item_stride, ++it) {
      const int m0 = (item * 2 + rank) * 128;
      const int row0 = m0 + q * 32;
      if (it == 0) pdl_wait();
      if (FFN_XT) {   // x[row, 64 cg .. 64 cg + 64) = k-block cg of the tile: 32 words of bf16 pairs -> tensor-memory columns 128 + 32 cg ..
        mbar_wait(x_full, it & 1);
        uint32_t xr[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 v = lds128(x_u32 + cg * 16384 + row * 128 + ((uint32_t(j) ^ sw) << 4));
          xr[4 * j] = v.x; xr[4 * j + 1] = v.y; xr[4 * j + 2] = v.z; xr[4 * j + 3] = v.w;
        }
        tmem_st32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(FFN_FC + cg * 32), xr);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(xt_full); else mbar_arrive_cluster_release(xt_full, 0);
        }
      }
      for (int c = 0; c < NC; ++c, ++gc) {
        const uint32_t zq = gc & 1;                          // z chunk buffer (BWD)
        const uint32_t b = FFN_XT ? 0u : (gc & 1);           // accumulator
        mbar_wait(&acc1_full[b], FFN_XT ? (gc & 1) : ((gc >> 1) & 1));
        tc_fence_after();
        if (it == 0 && warp == 2) mark(33 + c);
        uint32_t r[32];
        tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(b * FFN_FC + cg * 32), r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(&acc1_empty[b]); else mbar_arrive_cluster_relaxed(&acc1_empty[b], 0);
        }
        const int col0 = c * FFN_FC + cg * 32;   // hidden unit of r[0]
        uint32_t zp[16], hp[16];
        if (BWD) {   // dz = dh * gelu'(z): this lane's 64 bytes of the z chunk the producer prefetched
          mbar_wait(&z_full[zq], (gc >> 1) & 1);
          uint4 zin[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) zin[i] = lds128(z_u32 + zq * FFN_HC_BYTES + hz_off + ((uint32_t((cg & 1) * 4 + i) ^ sw) << 4));
          __syncwarp();
          if (lane == 0) mbar_arrive(&z_empty[zq]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[8 * i + u]);
            const float2 za = unpack_bf16(zin[i].x), zb = unpack_bf16(zin[i].y), zc = unpack_bf16(zin[i].z), zd = unpack_bf16(zin[i].w);
            gelu_grad_mul_pair(za.x, za.y, v[0], v[1]); gelu_grad_mul_pair(zb.x, zb.y, v[2], v[3]);
            gelu_grad_mul_pair(zc.x, zc.y, v[4], v[5]); gelu_grad_mul_pair(zd.x, zd.y, v[6], v[7]);
#pragma unroll
            for (int u = 0; u < 4; ++u) hp[4 * i + u] = pack_bf16(v[2 * u], v[2 * u + 1]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[8 * i + u]);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.b1 + col0 + 8 * i));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.b1 + col0 + 8 * i + 4));
            add_pair(v[0], v[1], b0.x, b0.y); add_pair(v[2], v[3], b0.z, b0.w);
            add_pair(v[4], v[5], b1.x, b1.y); add_pair(v[6], v[7], b1.z, b1.w);
#pragma unroll
            for (int u = 0; u < 4; ++u) zp[4 * i + u] = pack_bf16(v[2 * u], v[2 * u + 1]);
            gelu_pair(v[0], v[1]); gelu_pair(v[2], v[3]); gelu_pair(v[4], v[5]); gelu_pair(v[6], v[7]);
#pragma unroll
            for (int u = 0; u < 4; ++u) hp[4 * i + u] = pack_bf16(v[2 * u], v[2 * u + 1]);
          }
        }
        if (it == 0 && warp == 2) mark(49 + c);
        mbar_wait(h_empty, (gc & 1) ^ 1);   // GEMM2 of the previous chunk has read h
        if (it == 0 && warp == 2) mark(65 + c);
        if (store_h) {   // ... and so have the TMA stores of the previous chunk's h / z tiles
          if (issuer && lane == 0) tma_store_wait_read_n<0>();
          named_bar_sync(pair_bar, 64);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t off = hz_off + ((uint32_t((cg & 1) * 4 + i) ^ sw) << 4);
          sts128(h_u32 + off, hp[4 * i], hp[4 * i + 1], hp[4 * i + 2], hp[4 * i + 3]);
          if (!BWD && store_z) sts128(z_u32 + off, zp[4 * i], zp[4 * i + 1], zp[4 * i + 2], zp[4 * i + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(h_full); else mbar_arrive_cluster_relaxed(h_full, 0);
        }
        if (it == 0 && warp == 2) mark(81 + c);
        if (it == 0 && warp == 17) mark(101 + c);
        if (store_h) {
          named_bar_sync(pair_bar, 64);
          if (issuer && lane == 0) {
            const int gcol = c * FFN_FC + (cg >> 1) * 64;
            tma_store_2d_a(&tmHh, h_u32 + tile_off, gcol, row0);
            if (!BWD && store_z) tma_store_2d_a(&tmZ, z_u32 + tile_off, gcol, row0);
            tma_store_commit();
          }
        }
      }
      // ---- final epilogue: s = acc2 + b2 + x ; y = LN(s)   (cg = 64-column group of the 256-wide row)
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
      if (it == 0 && warp == 2) mark(97);
      if (store_h && issuer && lane == 0) tma_store_wait_read_n<0>();
      named_bar_sync(13, 32 * FFN_EPI_WARPS);   // every h / z tile has been read: the region becomes the staging tiles
      const uint32_t tmem_row = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(256 + cg * 64);
      const int colg = cg * 64;
      if (BWD) {   // dx = acc2 + g, optionally followed by the backward of the LayerNorm that fed this block (lnbwd_rows)
        const uint32_t t_sub[2] = {h_u32 + uint32_t(warp - 2) * 4096, h_u32 + uint32_t(warp - 2) * 4096 + 2048};
        const bool fuse_ln = p.lnb.s != nullptr;
        // saved-sum tiles: every z chunk and weight slot of this item has been consumed (acc2_full), and the producer does not
        // start the next item before x_empty: warps 0..7 stage in the second z buffer, warps 8..15 in the weight ring
        const int ew = warp - 2;
        const uint32_t s_base = ew < 8 ? z_u32 + FFN_HC_BYTES + uint32_t(ew) * 4096 : smem_u32(ring) + uint32_t(ew - 8) * 4096;
        const uint32_t s_sub[2] = {s_base, s_base + 2048};
        uint64_t* s_bar = &s_bars[ew * 2];
        float ln_mean = 0.f, ln_rstd = 0.f;
        if (fuse_ln) {
          if (lane == 0) {
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              mbar_arrive_expect_tx(&s_bar[sub], 2048);
              tma_load_2d_a(s_sub[sub], &tmY, &s_bar[sub], colg + sub * 32, row0);
            }
          }
          __syncwarp();
          if (row0 + lane < p.M) {
            ln_mean = __ldg(p.lnb.mean + row0 + lane);
            ln_rstd = __ldg(p.lnb.rstd + row0 + lane);
          }
        }
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t r[32];
          tmem_ld32(tmem_row + uint32_t(sub * 32), r);
          tmem_ld_wait();
          if (sub == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (rank == 0) mbar_arrive(acc2_empty); else mbar_arrive_cluster_relaxed(acc2_empty, 0);
            }
          }
          const uint32_t row64 = t_sub[sub] + lane * 64, sw64 = (lane >> 1) & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[8 * i + u]);
            const uint4 in = lds128(x_u32 + cg * 16384 + row * 128 + ((uint32_t(sub * 4 + i) ^ sw) << 4));
            const float2 a = unpack_bf16(in.x), bb = unpack_bf16(in.y), cc = unpack_bf16(in.z), d = unpack_bf16(in.w);
            add_pair(v[0], v[1], a.x, a.y); add_pair(v[2], v[3], bb.x, bb.y);
            add_pair(v[4], v[5], cc.x, cc.y); add_pair(v[6], v[7], d.x, d.y);
            sts128(row64 + ((i ^ sw64) << 4), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          }
        }
        auto store = [&](int sub) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d_a(&tmS, t_sub[sub], colg + sub * 32, row0);
            tma_store_commit();
          }
        };
        if (fuse_ln) {
          mbar_wait(&s_bar[0], it & 1);
          mbar_wait(&s_bar[1], it & 1);
          lnbwd_rows(p.lnb, t_sub, s_sub, ln_mean, ln_rstd, colg, q, cg, lane, ln_stats, 9 + q, 14, 32 * FFN_EPI_WARPS, store);
        } else {
          store(0);
          store(1);
        }
        // the staging tiles cover h_s and the first z buffer: the producer (next item's x, then z) starts only now
        if (lane == 0) tma_store_wait_read_n<0>();
        __syncwarp();
        if (lane == 0) mbar_arrive(x_empty);
        if (item + item_stride < p.n_items) named_bar_sync(13, 32 * FFN_EPI_WARPS);
        continue;
      }
      const uint32_t t_sub[2] = {h_u32 + uint32_t(warp - 2) * 4096, h_u32 + uint32_t(warp - 2) * 4096 + 2048};
      uint32_t sreg[2][16];
      float rsum = 0.f;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        uint32_t r[32];
        tmem_ld32(tmem_row + uint32_t(sub * 32), r);
        tmem_ld_wait();
        if (sub == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(acc2_empty); else mbar_arrive_cluster_relaxed(acc2_empty, 0);
          }
        }
        const int col0 = colg + sub * 32;
        const uint32_t row64 = t_sub[sub] + lane * 64, sw64 = (lane >> 1) & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[8 * i + u]);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.b2 + col0 + 8 * i));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.b2 + col0 + 8 * i + 4));
          add_pair(v[0], v[1], b0.x, b0.y); add_pair(v[2], v[3], b0.z, b0.w);
          add_pair(v[4], v[5], b1.x, b1.y); add_pair(v[6], v[7], b1.z, b1.w);
          // residual: x[row, col0 + 8i ..] sits in k-block cg of the x tile, 16-byte chunk sub*4 + i of the 128-byte row
          const uint4 in = lds128(x_u32 + cg * 16384 + row * 128 + ((uint32_t(sub * 4 + i) ^ sw) << 4));
          const float2 a = unpack_bf16(in.x), bb = unpack_bf16(in.y), cc = unpack_bf16(in.z), d = unpack_bf16(in.w);
          add_pair(v[0], v[1], a.x, a.y); add_pair(v[2], v[3], bb.x, bb.y);
          add_pair(v[4], v[5], cc.x, cc.y); add_pair(v[6], v[7], d.x, d.y);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t pk = pack_bf16(v[2 * u], v[2 * u + 1]);
            sreg[sub][4 * i + u] = pk;
            const float2 f = unpack_bf16(pk);
            rsum += f.x + f.y;
          }
          sts128(row64 + ((i ^ sw64) << 4), sreg[sub][4 * i], sreg[sub][4 * i + 1], sreg[sub][4 * i + 2], sreg[sub][4 * i + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d_a(&tmS, t_sub[sub], col0, row0);
          tma_store_commit();
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(x_empty);   // the residual has been read: the producer may load the next item's x
      const int rl = q * 32 + lane;
      ln_stats[cg * 128 + rl].x = rsum;
      named_bar_sync(9 + q, 128);
      const float mean = (ln_stats[rl].x + ln_stats[128 + rl].x + ln_stats[256 + rl].x + ln_stats[384 + rl].x) * (1.f / 256.f);
      float rsq = 0.f;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub)
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float2 f = unpack_bf16(sreg[sub][u]);
          const float dx = f.x - mean, dy = f.y - mean;
          rsq = fmaf(dx, dx, fmaf(dy, dy, rsq));
        }
      ln_stats[cg * 128 + rl].y = rsq;
      named_bar_sync(9 + q, 128);
      const float var = (ln_stats[rl].y + ln_stats[128 + rl].y + ln_stats[256 + rl].y + ln_stats[384 + rl].y) * (1.f / 256.f);
      const float rstd = rsqrtf(var + p.ln_eps);
      named_bar_sync(9 + q, 128);
      if (cg == 0 && row0 + lane < p.M) {
        p.ln_mean[row0 + lane] = mean;
        p.ln_rstd[row0 + lane] = rstd;
      }
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int col0 = colg + sub * 32;
        const uint32_t row64 = t_sub[sub] + lane * 64, sw64 = (lane >> 1) & 3;
        if (lane == 0) tma_store_wait_read_n<1>();   // the store of s from this tile has read it
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 f = unpack_bf16(sreg[sub][4 * i + u]);
            v[2 * u] = (f.x - mean) * rstd;
            v[2 * u + 1] = (f.y - mean) * rstd;
          }
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + col0 + 8 * i));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + col0 + 8 * i + 4));
          const float4 e0 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + col0 + 8 * i));
          const float4 e1 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + col0 + 8 * i + 4));
          v[0] = fmaf(v[0], g0.x, e0.x); v[1] = fmaf(v[1], g0.y, e0.y); v[2] = fmaf(v[2], g0.z, e0.z); v[3] = fmaf(v[3], g0.w, e0.w);
          v[4] = fmaf(v[4], g1.x, e1.x); v[5] = fmaf(v[5], g1.y, e1.y); v[6] = fmaf(v[6], g1.z, e1.z); v[7] = fmaf(v[7], g1.w, e1.w);
          sts128(row64 + ((i ^ sw64) << 4), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d_a(&tmY, t_sub[sub], col0, row0);
          tma_store_commit();
        }
      }
      // the staging tiles go back to being h / z: their stores must have read them before the next item's EPI1 writes
      if (item + item_stride < p.n_items) {
        if (lane == 0) tma_store_wait_read_n<0>();
        named_bar_sync(13, 32 * FFN_EPI_WARPS);
      }
    }
    if (lane == 0) tma_store_wait_read();   // the staging tiles have been read; the writes complete with the kernel
    if (warp == 2) mark(98);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_exit();
  if (warp == 1) tmem_dealloc_cg2<512>(tmem_base);
  if (warp == 0) mark(99);
}

#endif  // KIT_FFN_IMPL

int ffn_fwd_plan(FfnPlan* plan, const bf16* x, int64_t ldx, const bf16* w1, int64_t ldw1, const bf16* w2, int64_t ldw2,
                 const float* b1, const float* b2, bf16* z, bf16* hh, int64_t ldzh, bf16* s, int64_t lds, bf16* y, int64_t ldy,
                 const float* gamma, const float* beta, float* mean, float* rstd, float eps, int M, int H, int FF, int store_zh);
// dz = (g . W2) * gelu'(z) -> dz_out [M, FF];  dx = dz . W1 + g -> dx [M, H].  w2t = W2^T [FF, H], w1t = W1^T [H, FF] (bf16).
int ffn_bwd_plan(FfnPlan* plan, const bf16* g, int64_t ldg, const bf16* w2t, int64_t ldw2t, const bf16* w1t, int64_t ldw1t,
                 const bf16* z, bf16* dz_out, int64_t ldzh, bf16* dx, int64_t lddx, int M, int H, int FF,
                 const LnBwdArgs* lnb = nullptr);
bool ffn_fwd_supported(int H, int FF);
int ffn_launch(const FfnPlan* plan, cudaStream_t stream);

}  // namespace kit

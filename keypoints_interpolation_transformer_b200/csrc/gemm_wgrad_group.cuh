// Grouped stream-K weight-gradient GEMM: every dW = dy^T x (+ db = column sums of dy) of one transformer layer in ONE launch.
//
// A layer's backward has 4 (encoder) or 7 (decoder) weight gradients, all with the same reduction extent K = B*T and
// most of them tiny (256 x 256 outputs): launched one by one, each pays launch + pipeline fill + a split-K reduce epilogue
// for ~1 us of tensor work.  Here the k-blocks of all output tiles of all problems form one sequence of work units
// (problem-major, tile-major, k-minor) that is cut into equal contiguous ranges, one per CTA pair ("stream-K"): a pair
// streams through its range -- at most two partial tiles plus whole tiles -- and adds each (partial) tile to dW with the
// TMA reduce-add the split-K epilogue uses anyway, so no fix-up pass exists.  Same warp roles, cta_group::2 MMA, operand
// ring and all-ones bias-gradient MMA as gemm_tcgen05_kernel<256, 1, 2, EPI_F32, true> (gemm_sm100.cuh).
#pragma once
#include "gemm_sm100.cuh"

namespace kit {

constexpr int WG_MAX_PROBLEMS = 8;

struct WgradProblem {
  int M, N;               // dW is [M, N]: M = features of dy, N = features of x
  int groups_m, tiles_n;  // 256-row groups (CTA pairs) and 256-column tiles
  int unit_begin;         // first work unit (k-block of a pair tile) of this problem
  float* bias_grad;       // db [M] or null
};
struct WgradGroupParams {
  int n_problems, kb_total, total_units, units_per_cluster;
  WgradProblem prob[WG_MAX_PROBLEMS];
};
struct WgradGroupMaps {
  CUtensorMap a[WG_MAX_PROBLEMS], b[WG_MAX_PROBLEMS], c[WG_MAX_PROBLEMS];
};
struct WgradGroupPlan {
  WgradGroupMaps maps;
  WgradGroupParams p;
  int grid;
};

// One contiguous piece of a cluster's unit range: k-blocks [kb_begin, kb_end) of one pair tile of one problem.
struct WgradSegment {
  int prob, m_group, n_tile, kb_begin, kb_end;
};
struct WgradSegmentIter {
  const WgradGroupParams& p;
  int u, u_end;
  __device__ WgradSegmentIter(const WgradGroupParams& params, int cluster) : p(params) {
    u = cluster * p.units_per_cluster;
    u_end = min(u + p.units_per_cluster, p.total_units);
  }
  __device__ bool next(WgradSegment& s) {
    if (u >= u_end) return false;
    int pr = 0;
    while (pr + 1 < p.n_problems && p.prob[pr + 1].unit_begin <= u) ++pr;
    const int local = u - p.prob[pr].unit_begin;
    const int tau = local / p.kb_total;
    s.prob = pr;
    s.kb_begin = local - tau * p.kb_total;
    s.kb_end = min(p.kb_total, s.kb_begin + (u_end - u));
    s.n_tile = tau / p.prob[pr].groups_m;
    s.m_group = tau - s.n_tile * p.prob[pr].groups_m;
    u += s.kb_end - s.kb_begin;
    return true;
  }
};

constexpr int WG_BN = 256, WG_CL = 2;
#ifdef KIT_WGRAD_GROUP_IMPL   // the kernel itself is compiled into gemm.cu only
constexpr int WG_STAGES = gemm_stages<WG_BN, WG_CL, EPI_F32>();
constexpr int WG_SMEM = gemm_smem_bytes<WG_BN, WG_CL, EPI_F32>();

__global__ void __launch_bounds__(gemm_threads<WG_BN>(), 1) gemm_wgrad_group_kernel(const __grid_constant__ WgradGroupMaps maps,
                                                                                   const WgradGroupParams p) {
  constexpr int BM = GEMM_BM, BK = GEMM_BK, BN = WG_BN, CL = WG_CL, STAGES = WG_STAGES;
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = (BN / CL) * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = 2 * BN;   // accumulator in columns [0, BN), all-ones (bias gradient) product at column BN
  constexpr int EPI_WARPS = gemm_epi_warps<BN>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_smem + EPI_WARPS * 2 * 2048);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [1]
  uint64_t* tmem_empty = tmem_full + 1;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  uint8_t* ones_tile = reinterpret_cast<uint8_t*>(full) + 1024;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cluster = blockIdx.x / CL;

  if (warp == 0) {
    pdl_launch_dependents();
    constexpr int N_BARS = 2 * STAGES + 2;
    for (int i = lane; i < N_BARS; i += 32) mbar_init(&full[i], i == 2 * STAGES + 1 ? EPI_WARPS * CL : 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc_cg2<TMEM_COLS>(tmem_slot);
  if (warp >= 2 && threadIdx.x < 128) {
    *reinterpret_cast<uint4*>(ones_tile + (threadIdx.x - 64) * 16) = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_setup();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      WgradSegmentIter it(p, cluster);
      WgradSegment s;
      uint32_t cnt = 0;
      bool waited = false;
      while (it.next(s)) {
        const CUtensorMap* tmA = &maps.a[s.prob];
        const CUtensorMap* tmB = &maps.b[s.prob];
        const int m0 = (s.m_group * CL + rank) * BM;
        const int nb = s.n_tile * BN + rank * (BN / CL);
        for (int kb = s.kb_begin; kb < s.kb_end; ++kb, ++cnt) {
          const int st = cnt % STAGES;
          const uint32_t ph = (cnt / STAGES) & 1;
          mbar_wait(&empty[st], ph ^ 1);
          if (!waited) {
            pdl_wait();
            waited = true;
          }
          if (rank == 0) mbar_arrive_expect_tx(&full[st], STAGE_BYTES * CL);
          uint8_t* sA = smem + st * STAGE_BYTES;
          uint8_t* sB = sA + A_BYTES;
          const int kc = kb * BK;
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d_cg2(sA + j * (BK * 128), tmA, &full[st], m0 + 64 * j, kc);
#pragma unroll
          for (int j = 0; j < BN / CL / 64; ++j) tma_load_2d_cg2(sB + j * (BK * 128), tmB, &full[st], nb + 64 * j, kc);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (__shfl_sync(0xffffffffu, rank, 0) == 0) {
      uint32_t el_;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el_));
      const bool leader = el_ != 0;   // warp-uniform loop, tcgen05 instructions under elect.sync: operands stay in uniform registers
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc = make_idesc_bf16(BM * CL, BN, true, true);
      constexpr uint32_t idesc_ones = make_idesc_bf16(BM * CL, 16, true, true);
      const uint64_t odesc = make_smem_desc_sw128(smem_u32(ones_tile), 0, 0);
      WgradSegmentIter it(p, cluster);
      WgradSegment s;
      uint32_t cnt = 0, sidx = 0;
      while (it.next(s)) {
        mbar_wait(&tmem_empty[0], (sidx & 1) ^ 1);
        tc_fence_after();
        const bool with_bias = s.n_tile == 0 && p.prob[s.prob].bias_grad != nullptr;
        for (int kb = s.kb_begin; kb < s.kb_end; ++kb, ++cnt) {
          const int st = cnt % STAGES;
          const uint32_t ph = (cnt / STAGES) & 1;
          mbar_wait(&full[st], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + st * STAGE_BYTES);
          const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(a_base + k * 16 * 128, BK * 128, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(b_base + k * 16 * 128, BK * 128, 1024);
            const uint32_t acc = (kb > s.kb_begin || k > 0) ? 1u : 0u;
            if (leader) {
              umma_bf16_cg2(tb, adesc, bdesc, idesc, acc);
              if (with_bias) umma_bf16_cg2(tb + BN, adesc, odesc, idesc_ones, acc);
            }
          }
          if (leader) umma_commit_cg2(&empty[st], (uint16_t)3);
        }
        if (leader) umma_commit_cg2(&tmem_full[0], (uint16_t)3);
        ++sidx;
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const uint32_t tile = smem_u32(epi_smem) + (warp - 2) * (2 * 2048);   // one [32 x 128 B] fp32 staging tile
    WgradSegmentIter it(p, cluster);
    WgradSegment s;
    uint32_t sidx = 0;
    bool waited = false;
    while (it.next(s)) {
      const WgradProblem& pr = p.prob[s.prob];
      const CUtensorMap* tmC = &maps.c[s.prob];
      const int m0 = (s.m_group * CL + rank) * BM, n0 = s.n_tile * BN;
      const int row0 = m0 + q * 32, colg = n0 + cg * 64;
      const uint32_t tmem_row = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(cg * 64);
      auto release_tmem = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(&tmem_empty[0]); else mbar_arrive_cluster_relaxed(&tmem_empty[0], 0);
        }
      };
      if (!waited) {
        pdl_wait();
        waited = true;
      }
      mbar_wait(&tmem_full[0], sidx & 1);
      ++sidx;
      tc_fence_after();
      const bool have0 = colg < pr.N, have1 = colg + 32 < pr.N;
      if (!have0) {
        release_tmem();
        continue;
      }
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        const int col0 = colg + sub * 32;
        if (sub == 1 && !have1) {
          release_tmem();
          break;
        }
        uint32_t r[32];
        if (sub == 0 && cg == 0 && s.n_tile == 0 && pr.bias_grad != nullptr) {   // column 0 of the all-ones product
          tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(BN), r);
          tmem_ld_wait();
          if (row0 + lane < pr.M) atomicAdd(pr.bias_grad + row0 + lane, __uint_as_float(r[0]));
        }
        tmem_ld32(tmem_row + uint32_t(sub * 32), r);
        tmem_ld_wait();
        if (sub == 1) release_tmem();
        if (lane == 0) tma_store_wait_read_n<0>();
        __syncwarp();
        const uint32_t row128 = tile + lane * 128, sw128 = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) sts128(row128 + ((i ^ sw128) << 4), r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d_a(tmC, tile, col0, row0);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_read();   // the staging tiles have been read; the reduce-adds complete with the kernel
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_exit();
  if (warp == 1) tmem_dealloc_cg2<TMEM_COLS>(tmem_base);
}

#endif  // KIT_WGRAD_GROUP_IMPL

// problems: dW_i [M_i, N_i] (fp32, row pitch ldc_i) += A_i^T B_i with A_i = dy_i [K, M_i] (pitch lda_i), B_i = x_i [K, N_i];
// bias_grad_i [M_i] += column sums of dy_i (or null).  Every pitch / pointer must satisfy the TMA alignment rules.
struct WgradProblemDesc {
  const bf16* A;
  int64_t lda;
  const bf16* B;
  int64_t ldb;
  float* C;
  int64_t ldc;
  int M, N;
  float* bias_grad;
};
bool wgrad_group_supported(const WgradProblemDesc& d);
int wgrad_group_plan(WgradGroupPlan* plan, const WgradProblemDesc* probs, int n, int K);
int wgrad_group_launch(const WgradGroupPlan* plan, cudaStream_t stream);

}  // namespace kit

// tcgen05 / TMEM / TMA GEMM for sm_100a -- persistent, warp-specialised.
//
// Each CTA (one per SM) loops over 128 x BN output tiles.  Operands are staged by TMA into
// 128B-swizzled shared memory through a multi-stage mbarrier ring that runs ahead across tile
// boundaries; one elected thread issues tcgen05.mma into one of TWO fp32 accumulators in tensor
// memory, so the epilogue of tile i (tcgen05.ld -> fused epilogue -> TMA stores) overlaps the
// TMA + MMA main loop of tile i+1.
//
//   MODE 0 (TN)    C[M,N] = A[M,K] * B[N,K]^T      both operands K-major      (y = x W^T, dx = dy W)
//   MODE 1 (wgrad) C[M,N] = A[K,M]^T * B[K,N]      both operands MN-major     (dW = dy^T x), split-K
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2.. = epilogue
// (BN/16 of them: TMEM lane quadrant = warp % 4, 64-column group = (warp - 2) / 4).
//
// CL = 2: a CTA pair (thread-block cluster) computes a 256 x BN tile with tcgen05.mma.cta_group::2: each CTA stages its own
// 128 rows of A and only HALF of the B tile (32 KB per k-block instead of 48 KB), the leader CTA's MMA thread issues for
// both, and each CTA's epilogue drains its own 128 accumulator rows.
//
// The epilogue is a compile-time kind (EPI_*): what bounds the K = 256 GEMMs of this model is not the tensor pipe but the
// CUDA-core work per output element and the latency chain tcgen05.ld -> math -> st.shared -> fence -> TMA store of each
// epilogue warp, so every kind is a straight-line loop over 16-byte chunks with shared-space accesses, inputs (residual /
// GELU' pre-activation) prefetched by TMA before the accumulator is ready, and 2 or 3 staging tiles per warp in rotation so
// that a warp rarely waits for its own previous store.
#pragma once
#include "common.cuh"

namespace kit {

enum { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_ATOMIC = 2 };
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_GELU_BWD = 2 };
enum {
  EPI_STORE = 0,     // bf16 C = acc (+ bias)
  EPI_ADD = 1,       // bf16 C = acc (+ bias) + addend            (addend tile by TMA, result written in place)
  EPI_GELU = 2,      // bf16 aux = acc + bias ; C = gelu(aux)
  EPI_GELU_BWD = 3,  // bf16 C = acc * gelu'(aux)                  (aux tile by TMA, result written in place)
  EPI_F32 = 4,       // fp32 C = acc (+ bias): TMA store or TMA reduce-add (split-K weight gradients)
  EPI_GENERIC = 5,   // any combination through direct global accesses (pitches / extents a tensor map cannot describe)
  EPI_ADD_LN = 6,    // EPI_ADD + LayerNorm of the sum over the full row (N = BN = 256): s = acc + bias + addend -> C (bf16),
                     // y = LN(s) * gamma + beta -> D (bf16), mean / rstd of the bf16-rounded s (what the backward reads)
  EPI_ADD_LNBWD = 7, // EPI_ADD followed by the BACKWARD of the LayerNorm whose output gradient the sum is (N = BN = 256):
                     // dy = acc + addend ; dx = rstd (dy g - mean(dy g) - xhat mean(dy g xhat)) -> C (bf16); dgamma / dbeta partials
  EPI_KINDS = 8
};

// LayerNorm backward fused behind the GEMM (or the fused FFN kernel) that produces the gradient w.r.t. the LayerNorm output.
struct LnBwdArgs {
  const bf16* s;        // the saved pre-norm sum [M, 256] (what the forward normalised), null = not fused
  int64_t ld_s;
  const float* gamma;   // LayerNorm weight [256]
  const float* mean;    // per-row statistics of s [M]
  const float* rstd;
  float* dgamma;        // [256] += sum_rows dy * xhat
  float* dbeta;         // [256] += sum_rows dy
};

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
  const float* ln_gamma;   // EPI_ADD_LN: LayerNorm weight / bias [N], per-row statistics out [M]
  const float* ln_beta;
  float* ln_mean;
  float* ln_rstd;
  float ln_eps;
  LnBwdArgs lnb;      // EPI_ADD_LNBWD
  float* bias_grad;   // MODE 1, BG kernels: bias_grad[m] += sum_k A[k, m] (column sums of dy), else null
  int manual_out;     // EPI_F32 over fp32 rows a tensor map cannot describe (pitch a multiple of 8 but not of 16 bytes: the 142-wide
                      // prediction and embedding gradients; TMA needs 16-byte aligned row starts): the staged [32 x 32] tile leaves
                      // by coalesced 8-byte stores / vector reductions, two rows of 128 bytes per warp instruction
  long long* trace;   // experiments only (KIT_GEMM_TRACE): clock64 marks of CTA 0, see kit_gemm_trace_read
};

struct GemmPlan {
  CUtensorMap tmA, tmB, tmC, tmAux, tmD;
  GemmParams p;
  int mode, bn, epi;
  int grid;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_MAX_EPI_WARPS = 16;
constexpr int GEMM_SMEM_LIMIT = 232448;   // 227 KB per CTA
constexpr int GEMM_SMEM_TAIL = 3072;      // barriers (1 KB) + all-ones operand tile (1 KB, BG kernels) + 1024-byte alignment slack
constexpr int GEMM_LN_STATS = 4096;       // EPI_ADD_LN: [4 column groups][128 rows] float2 partial row statistics

template <int BN>
constexpr int gemm_epi_warps() { return BN / 16; }
template <int BN>
constexpr int gemm_threads() { return 64 + 32 * gemm_epi_warps<BN>(); }
// 2 KB staging tiles ([32 rows x 64 B] bf16, or half of a [32 x 128 B] fp32 tile) per epilogue warp
template <int EPI>
constexpr int gemm_epi_tiles() { return EPI == EPI_GENERIC ? 0 : (EPI == EPI_GELU || EPI == EPI_GELU_BWD) ? 3 : 2; }
template <int EPI>
constexpr int gemm_tail_bytes() { return GEMM_SMEM_TAIL + ((EPI == EPI_ADD_LN || EPI == EPI_ADD_LNBWD) ? GEMM_LN_STATS : 0); }
template <int BN, int CL>
constexpr int gemm_stage_bytes() { return GEMM_BM * GEMM_BK * 2 + (BN / CL) * GEMM_BK * 2; }
template <int BN, int CL, int EPI>
constexpr int gemm_stages() {
  const int avail = GEMM_SMEM_LIMIT - gemm_tail_bytes<EPI>() - gemm_epi_warps<BN>() * gemm_epi_tiles<EPI>() * 2048;
  const int s = avail / gemm_stage_bytes<BN, CL>();
  return s > 6 ? 6 : s;
}
template <int BN, int CL, int EPI>
constexpr int gemm_smem_bytes() {
  return gemm_stages<BN, CL, EPI>() * gemm_stage_bytes<BN, CL>() + gemm_epi_warps<BN>() * gemm_epi_tiles<EPI>() * 2048 + gemm_tail_bytes<EPI>();
}

// ---------------------------------------------------------------- shared-space accesses / TMA by 32-bit shared address
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tma_load_2d_a(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_a(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d_a(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read_n() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void add_pair(float& a, float& b, float x, float y) { up2(add2(pk2(a, b), pk2(x, y)), a, b); }

// One row x 32 consecutive columns of the accumulator through the fused epilogue with direct global accesses (EPI_GENERIC).
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

// 32 accumulator columns of this lane's row -> the staging tile(s), 8 columns (one 16-byte bf16 chunk) at a time.
//   t_out / t_aux: shared addresses of [32 rows x 64 B] tiles in the TMA SWIZZLE_64B layout (16-byte chunk i of row r sits at
//   chunk i ^ ((r >> 1) & 3)); EPI_F32: t_out is a [32 x 128 B] SWIZZLE_128B tile (chunk i ^ (r & 7)).
//   EPI_ADD / EPI_GELU_BWD read their input from t_out and overwrite it in place.
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_sub(const GemmParams& p, int col0, int lane, const uint32_t (&r)[32], uint32_t t_out,
                                                  uint32_t t_aux) {
  const uint32_t row64 = t_out + lane * 64, sw64 = (lane >> 1) & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[8 * i + u]);
    if (p.bias != nullptr && col0 + 8 * i < p.N) {   // TMA kinds with a bias have N % 8 == 0 (planner); manual_out may end mid-chunk
      if (EPI != EPI_F32 || col0 + 8 * i + 8 <= p.N) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 8 * i));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 8 * i + 4));
        add_pair(v[0], v[1], b0.x, b0.y); add_pair(v[2], v[3], b0.z, b0.w);
        add_pair(v[4], v[5], b1.x, b1.y); add_pair(v[6], v[7], b1.z, b1.w);
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (col0 + 8 * i + u < p.N) v[u] += __ldg(p.bias + col0 + 8 * i + u);
      }
    }
    if (EPI == EPI_F32) {
      const uint32_t row128 = t_out + lane * 128, sw128 = lane & 7;
      sts128(row128 + (((2 * i) ^ sw128) << 4), __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
      sts128(row128 + (((2 * i + 1) ^ sw128) << 4), __float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
      continue;
    }
    const uint32_t off = (i ^ sw64) << 4;
    if (EPI == EPI_ADD || EPI == EPI_GELU_BWD) {
      const uint4 in = lds128(row64 + off);
      const float2 a = unpack_bf16(in.x), b = unpack_bf16(in.y), c = unpack_bf16(in.z), d = unpack_bf16(in.w);
      if (EPI == EPI_ADD) {
        add_pair(v[0], v[1], a.x, a.y); add_pair(v[2], v[3], b.x, b.y);
        add_pair(v[4], v[5], c.x, c.y); add_pair(v[6], v[7], d.x, d.y);
      } else {
        gelu_grad_mul_pair(a.x, a.y, v[0], v[1]); gelu_grad_mul_pair(b.x, b.y, v[2], v[3]);
        gelu_grad_mul_pair(c.x, c.y, v[4], v[5]); gelu_grad_mul_pair(d.x, d.y, v[6], v[7]);
      }
    }
    if (EPI == EPI_GELU) {
      sts128(t_aux + lane * 64 + off, pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      gelu_pair(v[0], v[1]); gelu_pair(v[2], v[3]); gelu_pair(v[4], v[5]); gelu_pair(v[6], v[7]);
    }
    sts128(row64 + off, pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

// ---------------------------------------------------------------- LayerNorm backward on accumulator rows
// A row's 256 columns sit in the four epilogue warps of one TMEM lane quadrant: warp column group cg holds columns
// [64 cg, 64 cg + 64) of row (32 q + lane) as 2 x 16 packed bf16 pairs.
// sum over the 32 lanes of a[c] for every c: on return lane l holds column l's sum (31 shuffles: halve the array each step)
__device__ __forceinline__ float warp_col_reduce32(float (&a)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float keep = up ? a[i + h] : a[i];
      const float send = up ? a[i] : a[i + h];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return a[0];
}
// The gradient w.r.t. the LayerNorm output (dy, bf16) sits in the warp's two [32 rows x 32 columns] staging tiles (SWIZZLE_64B:
// 16-byte chunk i of row r at chunk i ^ ((r >> 1) & 3)); s_sub: the saved pre-norm sum in two more tiles of the same layout.  On return the tiles hold dx and
// store(sub) has been called for each (fence + TMA store are the caller's).  ln_stats: [4][128] float2 scratch shared by the
// quadrant's four warps; bar_id: their named barrier (128 threads).  Register budget (96 per thread next to 16 epilogue
// warps): dy is re-read from the tiles in every pass, only one 32-column array is alive at a time.
template <typename StoreFn>
__device__ __forceinline__ void lnbwd_rows(const LnBwdArgs& a, const uint32_t (&t_sub)[2], const uint32_t (&s_sub)[2], float mean,
                                           float rstd, int colg, int q, int cg, int lane, float2* ln_stats, int bar_id,
                                           int bar_all, int n_epi_threads, StoreFn&& store) {
  const uint32_t sw64 = (lane >> 1) & 3;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int sub = 0; sub < 2; ++sub)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 in = lds128(t_sub[sub] + lane * 64 + ((i ^ sw64) << 4));
      const uint32_t vw[4] = {in.x, in.y, in.z, in.w};
      const uint4 sc = lds128(s_sub[sub] + lane * 64 + ((i ^ sw64) << 4));
      const uint32_t sw_[4] = {sc.x, sc.y, sc.z, sc.w};
      const float4 ga = __ldg(reinterpret_cast<const float4*>(a.gamma + colg + sub * 32 + 8 * i));
      const float4 gb = __ldg(reinterpret_cast<const float4*>(a.gamma + colg + sub * 32 + 8 * i + 4));
      const float gm[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 v = unpack_bf16(vw[u]), x = unpack_bf16(sw_[u]);
        const float g0 = v.x * gm[2 * u], g1 = v.y * gm[2 * u + 1];
        s1 += g0 + g1;
        s2 = fmaf(g0, (x.x - mean) * rstd, s2);
        s2 = fmaf(g1, (x.y - mean) * rstd, s2);
      }
    }
  const int rl = q * 32 + lane;
  ln_stats[cg * 128 + rl] = make_float2(s1, s2);
  asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
  const float2 p0 = ln_stats[rl], p1 = ln_stats[128 + rl], p2 = ln_stats[256 + rl], p3 = ln_stats[384 + rl];
  const float c1 = ((p0.x + p1.x) + (p2.x + p3.x)) * (1.f / 256.f);
  const float c2 = ((p0.y + p1.y) + (p2.y + p3.y)) * (1.f / 256.f);
  asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // everyone has read the partials: the slots may be reused
  float sb[2], sg[2];   // this warp's 32-row partial of dbeta / dgamma for column colg + 32 sub + lane
#pragma unroll
  for (int sub = 0; sub < 2; ++sub) {
    const uint32_t row64 = t_sub[sub] + lane * 64;
    float col[32];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 in = lds128(row64 + ((i ^ sw64) << 4));
      const float2 v0 = unpack_bf16(in.x), v1 = unpack_bf16(in.y), v2 = unpack_bf16(in.z), v3 = unpack_bf16(in.w);
      col[8 * i] = v0.x; col[8 * i + 1] = v0.y; col[8 * i + 2] = v1.x; col[8 * i + 3] = v1.y;
      col[8 * i + 4] = v2.x; col[8 * i + 5] = v2.y; col[8 * i + 6] = v3.x; col[8 * i + 7] = v3.y;
    }
    sb[sub] = warp_col_reduce32(col, lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 in = lds128(row64 + ((i ^ sw64) << 4));
      const uint32_t vw[4] = {in.x, in.y, in.z, in.w};
      const uint4 sc = lds128(s_sub[sub] + lane * 64 + ((i ^ sw64) << 4));
      const uint32_t sw_[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 v = unpack_bf16(vw[u]), x = unpack_bf16(sw_[u]);
        col[8 * i + 2 * u] = v.x * ((x.x - mean) * rstd);
        col[8 * i + 2 * u + 1] = v.y * ((x.y - mean) * rstd);
      }
    }
    sg[sub] = warp_col_reduce32(col, lane);
  }
#pragma unroll
  for (int sub = 0; sub < 2; ++sub) {   // dx = rstd (dy gamma - c1 - xhat c2), in place, and out
    const uint32_t row64 = t_sub[sub] + lane * 64;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t addr = row64 + ((i ^ sw64) << 4);
      const uint4 in = lds128(addr);
      const uint32_t vw[4] = {in.x, in.y, in.z, in.w};
      const uint4 sc = lds128(s_sub[sub] + lane * 64 + ((i ^ sw64) << 4));
      const uint32_t sw_[4] = {sc.x, sc.y, sc.z, sc.w};
      const float4 ga = __ldg(reinterpret_cast<const float4*>(a.gamma + colg + sub * 32 + 8 * i));
      const float4 gb = __ldg(reinterpret_cast<const float4*>(a.gamma + colg + sub * 32 + 8 * i + 4));
      const float gm[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
      uint32_t o[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 v = unpack_bf16(vw[u]), x = unpack_bf16(sw_[u]);
        const float d0 = rstd * (v.x * gm[2 * u] - c1 - (x.x - mean) * rstd * c2);
        const float d1 = rstd * (v.y * gm[2 * u + 1] - c1 - (x.y - mean) * rstd * c2);
        o[u] = pack_bf16(d0, d1);
      }
      sts128(addr, o[0], o[1], o[2], o[3]);
    }
    store(sub);
  }
  // dbeta / dgamma: the four quadrants' partials meet in shared memory (the statistics scratch, now [4][256] floats), so each
  // column costs ONE atomic per CTA -- 512-way contention per address (one per warp) was slower than the kernel it replaces
  float* scratch = reinterpret_cast<float*>(ln_stats);
  const int c0 = colg + lane;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_all), "r"(n_epi_threads) : "memory");   // the scratch is free (statistics / previous round read)
    scratch[q * 256 + c0] = which == 0 ? sb[0] : sg[0];
    scratch[q * 256 + c0 + 32] = which == 0 ? sb[1] : sg[1];
    asm volatile("bar.sync %0, %1;" ::"r"(bar_all), "r"(n_epi_threads) : "memory");
    if (q == 0) {
      float* dst = which == 0 ? a.dbeta : a.dgamma;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int c = c0 + 32 * sub;
        atomicAdd(dst + c, (scratch[c] + scratch[256 + c]) + (scratch[512 + c] + scratch[768 + c]));
      }
    }
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar_all), "r"(n_epi_threads) : "memory");   // before the scratch holds statistics again
}

// BG (MODE 1 only): the bias gradient of the same Linear -- the column sums of dy, i.e. the row sums of the A^T operand --
// comes out of the tensor pipe too: every k-step issues one more MMA of the A tile against an all-ones [16 x 16] B tile
// into 16 spare accumulator columns (one accumulator stage instead of two: weight-gradient work items are one per CTA
// pair), and the epilogue of the first n tile adds column 0 of it to bias_grad.  This replaces a separate pass over dy.
template <int BN, int MODE, int CL, int EPI, bool BG = false>
__global__ void __launch_bounds__(gemm_threads<BN>(), 1) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                         const __grid_constant__ CUtensorMap tmB,
                                                                         const __grid_constant__ CUtensorMap tmC,
                                                                         const __grid_constant__ CUtensorMap tmAux,
                                                                         const __grid_constant__ CUtensorMap tmD,
                                                                         const GemmParams p) {
  constexpr int BM = GEMM_BM, BK = GEMM_BK;
  constexpr int STAGES = gemm_stages<BN, CL, EPI>();
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = (BN / CL) * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = 2 * BN;  // two accumulators
  constexpr int EPI_WARPS = gemm_epi_warps<BN>();
  constexpr int EPI_TILES = gemm_epi_tiles<EPI>();
  static_assert(BN == 128 || BN == 256, "BN must give a power-of-two TMEM allocation <= 512");
  static_assert(STAGES >= 3, "operand ring too shallow");
  static_assert(!BG || (MODE == 1 && EPI == EPI_F32), "the fused bias gradient exists for weight-gradient GEMMs only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_smem + EPI_WARPS * EPI_TILES * 2048);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint64_t* in_bars = tmem_empty + 2;     // [EPI_WARPS][2] epilogue input tiles (addend / GELU' pre-activation)
  uint64_t* s_bars = in_bars + 2 * GEMM_MAX_EPI_WARPS;   // [EPI_WARPS][2] EPI_ADD_LNBWD: the saved-sum tiles
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_bars + 2 * GEMM_MAX_EPI_WARPS);
  uint8_t* ones_tile = reinterpret_cast<uint8_t*>(full) + 1024;   // BG: 1 KB of bf16 1.0 (any layout of all-ones is all-ones)
  float2* ln_stats = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(full) + 2048);   // EPI_ADD_LN: [4][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto mark = [&](int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && lane == 0) p.trace[slot] = clock64();
  };
  if (warp == 0) mark(0);
  const int kb_total = (p.K + BK - 1) / BK;
  // work items are (split, n tile, group of CL m tiles); CTA `rank` of a cluster takes m tile CL*group + rank
  const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;
  const int first_item = blockIdx.x / CL, item_stride = gridDim.x / CL;
  const int groups_m = (p.tiles_m + CL - 1) / CL;
  const int tiles_mn = groups_m * p.tiles_n;
  const int n_items = tiles_mn * p.splits;

  if (warp == 0) {   // one barrier per lane: a single thread initialising ~45 barriers costs ~1.5k cycles per launch
    pdl_launch_dependents();   // the next kernel's CTAs may be scheduled as soon as SMs free up
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      if (EPI != EPI_GENERIC) tma_prefetch_desc(&tmC);
      if (EPI == EPI_ADD || EPI == EPI_GELU || EPI == EPI_GELU_BWD || EPI == EPI_ADD_LN || EPI == EPI_ADD_LNBWD) tma_prefetch_desc(&tmAux);
      if (EPI == EPI_ADD_LN || EPI == EPI_ADD_LNBWD) tma_prefetch_desc(&tmD);
    }
    // full[s]: the producer's arrive.expect_tx (CL = 2: only the leader's is used; it collects the bytes of both CTAs);
    // empty[s], tmem_full[s]: one tcgen05.commit arrival (multicast to both CTAs when CL = 2);
    // tmem_empty[s]: every epilogue warp (CL = 2: the peer's warps arrive on the leader's barrier); in_bars: one TMA load
    constexpr int N_BARS = 2 * STAGES + 4 + 2 * EPI_WARPS;
    for (int i = lane; i < N_BARS; i += 32) {
      const bool is_tmem_empty = i >= 2 * STAGES + 2 && i < 2 * STAGES + 4;
      mbar_init(&full[i], is_tmem_empty ? EPI_WARPS * CL : 1);
    }
    if (EPI == EPI_ADD_LNBWD) mbar_init(&s_bars[lane], 1);
    fence_barrier_init();
    fence_proxy_async();
    mark(12);
  }
  if (warp == 1) {
    if (CL == 1) tmem_alloc<TMEM_COLS>(tmem_slot); else tmem_alloc_cg2<TMEM_COLS>(tmem_slot);
    mark(13);
  }
  // The producer is one thread: the index arithmetic of its first work item (three integer divisions, ~160 dependent
  // instructions, ~0.9 k cycles between the set-up barrier and the first TMA load) is done here by an idle epilogue warp.
  int* first_idx = reinterpret_cast<int*>(tmem_slot + 2);   // n0, m0, kb_begin, kb_end
  if (warp == 3 && lane == 0 && first_item < n_items) {
    const int split = first_item / tiles_mn, rem = first_item - split * tiles_mn;
    first_idx[0] = (rem / groups_m) * BN;
    first_idx[1] = ((rem % groups_m) * CL + rank) * BM;
    first_idx[2] = split * p.kb_per_split;
    first_idx[3] = min(kb_total, split * p.kb_per_split + p.kb_per_split);
  }
  if (BG && warp >= 2 && threadIdx.x < 64 + 64) {   // 64 threads x 16 bytes
    *reinterpret_cast<uint4*>(ones_tile + (threadIdx.x - 64) * 16) = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async();   // read by the tensor core (async proxy)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) mark(14);
  if (CL > 1) cluster_sync_setup();   // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) mark(1);
  // Everything above overlapped the previous kernel.  Only the threads that touch global memory wait for it to complete
  // (pdl_wait below: the producer before its first TMA load, the epilogue warps before their first global access), after
  // the index arithmetic of their first work item.

  if (warp == 0) {
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int item = first_item; item < n_items; item += item_stride) {
        int n0, m0, kb_begin, kb_end;
        if (item == first_item) {
          n0 = first_idx[0]; m0 = first_idx[1]; kb_begin = first_idx[2]; kb_end = first_idx[3];
        } else {
          const int split = item / tiles_mn, rem = item - split * tiles_mn;
          n0 = (rem / groups_m) * BN;
          m0 = ((rem % groups_m) * CL + rank) * BM;
          kb_begin = split * p.kb_per_split;
          kb_end = min(kb_total, kb_begin + p.kb_per_split);
        }
        const int nb = n0 + rank * (BN / CL);   // CL = 2: this CTA stages its half of the B tile
        for (int kb = kb_begin; kb < kb_end; ++kb, ++cnt) {
          const int s = cnt % STAGES;
          const uint32_t ph = (cnt / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          if (cnt == 0) {
            pdl_wait();
            mark(3);
          }
          if (rank == 0) mbar_arrive_expect_tx(&full[s], STAGE_BYTES * CL);
          uint8_t* sA = smem + s * STAGE_BYTES;
          uint8_t* sB = sA + A_BYTES;
          const int kc = kb * BK;
          if (MODE == 0) {
            if (CL == 1) {
              tma_load_2d(sA, &tmA, &full[s], kc, m0);
              tma_load_2d(sB, &tmB, &full[s], kc, nb);
            } else {
              tma_load_2d_cg2(sA, &tmA, &full[s], kc, m0);
              tma_load_2d_cg2(sB, &tmB, &full[s], kc, nb);
            }
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) {
              if (CL == 1) tma_load_2d(sA + j * (BK * 128), &tmA, &full[s], m0 + 64 * j, kc);
              else tma_load_2d_cg2(sA + j * (BK * 128), &tmA, &full[s], m0 + 64 * j, kc);
            }
#pragma unroll
            for (int j = 0; j < BN / CL / 64; ++j) {
              if (CL == 1) tma_load_2d(sB + j * (BK * 128), &tmB, &full[s], nb + 64 * j, kc);
              else tma_load_2d_cg2(sB + j * (BK * 128), &tmB, &full[s], nb + 64 * j, kc);
            }
          }
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
      constexpr uint32_t idesc = make_idesc_bf16(BM * CL, BN, MODE == 1, MODE == 1);
      uint32_t cnt = 0, it = 0;
      for (int item = first_item; item < n_items; item += item_stride, ++it) {
        const int split = item / tiles_mn;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
        const uint32_t as = BG ? 0 : (it & 1), aph = BG ? (it & 1) : ((it >> 1) & 1);
        mbar_wait(&tmem_empty[as], aph ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tb + as * BN;
        const bool first_n = BG && ((item - split * tiles_mn) / groups_m) == 0 && p.bias_grad != nullptr;
        for (int kb = kb_begin; kb < kb_end; ++kb, ++cnt) {
          const int s = cnt % STAGES;
          const uint32_t ph = (cnt / STAGES) & 1;
          mbar_wait(&full[s], ph);
          if (cnt == 0) mark(4);
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
            if (leader) {
              if (CL == 1) umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
              else umma_bf16_cg2(tmem_d, adesc, bdesc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            }
            if (BG && first_n) {   // row sums of A^T: same A tile against all-ones, 16 columns at tmem column BN
              constexpr uint32_t idesc_ones = make_idesc_bf16(BM * CL, 16, true, true);
              const uint64_t odesc = make_smem_desc_sw128(smem_u32(ones_tile), 0, 0);
              if (leader) {
                if (CL == 1) umma_bf16(tb + BN, adesc, odesc, idesc_ones, (kb > kb_begin || k > 0) ? 1u : 0u);
                else umma_bf16_cg2(tb + BN, adesc, odesc, idesc_ones, (kb > kb_begin || k > 0) ? 1u : 0u);
              }
            }
          }
          if (leader) {
            if (CL == 1) umma_commit(&empty[s]); else umma_commit_cg2(&empty[s], (uint16_t)3);   // slot free in both CTAs
          }
        }
        if (leader) {
          if (CL == 1) umma_commit(&tmem_full[as]); else umma_commit_cg2(&tmem_full[as], (uint16_t)3);
        }
        if (it == 0) mark(5);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;   // 64-column group of the tile owned by this warp
    const uint32_t tiles = smem_u32(epi_smem) + (warp - 2) * (EPI_TILES * 2048);
    uint64_t* in_bar = &in_bars[(warp - 2) * 2];
    uint32_t in_ph[2] = {0, 0};
    uint32_t rr = 0;   // rotation of the staging tiles (3-tile kinds)
    uint32_t it = 0;
    for (int item = first_item; item < n_items; item += item_stride, ++it) {
      const int split = item / tiles_mn, rem = item - split * tiles_mn;
      const int n0 = (rem / groups_m) * BN, m0 = ((rem % groups_m) * CL + rank) * BM;
      const uint32_t as = BG ? 0 : (it & 1), aph = BG ? (it & 1) : ((it >> 1) & 1);
      const int row0 = m0 + q * 32;
      const int colg = n0 + cg * 64;
      const uint32_t tmem_row = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN + cg * 64);
      auto release_tmem = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CL == 1 || rank == 0) mbar_arrive(&tmem_empty[as]); else mbar_arrive_cluster_relaxed(&tmem_empty[as], 0);   // the accumulator reads are complete (tcgen05.wait::ld): nothing for a release fence to order
        }
      };
      const bool have0 = colg < p.N, have1 = colg + 32 < p.N;   // warp-uniform
      if (it == 0) pdl_wait();
      // staging tiles of the two 32-column halves
      uint32_t t_sub[2] = {tiles, tiles}, t_aux[2] = {tiles, tiles};
      if (EPI == EPI_GELU) {            // stores in order: aux0 -> rr, out0 -> rr+1, aux1 -> rr+2, out1 -> rr+3 (= rr)
        t_aux[0] = tiles + (rr % 3) * 2048;       t_sub[0] = tiles + ((rr + 1) % 3) * 2048;
        t_aux[1] = tiles + ((rr + 2) % 3) * 2048; t_sub[1] = tiles + (rr % 3) * 2048;
        rr += 4;
      } else if (EPI == EPI_GELU_BWD) {  // stores go round the three tiles: a tile's previous store is 3 stores old
        t_sub[0] = tiles + (rr % 3) * 2048;
        t_sub[1] = tiles + ((rr + 1) % 3) * 2048;
        rr += 2;
      } else if (EPI == EPI_STORE || EPI == EPI_ADD || EPI == EPI_ADD_LN || EPI == EPI_ADD_LNBWD) {
        t_sub[1] = tiles + 2048;
      }
      if (EPI != EPI_GENERIC && have0 && p.bias != nullptr && lane < 2 && colg + 32 * lane < p.N) prefetch_l1(p.bias + colg + 32 * lane);
      if (EPI == EPI_ADD || EPI == EPI_GELU_BWD || EPI == EPI_ADD_LN || EPI == EPI_ADD_LNBWD) {   // input tiles land while the MMAs of this tile are still running
        if (lane == 0 && have0) {
          tma_store_wait_read_n<1>();   // the tile(s) below were last read by stores that are at least 2 groups old
          mbar_arrive_expect_tx(&in_bar[0], 2048);
          tma_load_2d_a(t_sub[0], &tmAux, &in_bar[0], colg, row0);
          if (have1) {
            if (EPI == EPI_ADD || EPI == EPI_ADD_LN || EPI == EPI_ADD_LNBWD) tma_store_wait_read_n<0>();   // two tiles only: tile 1 carried the most recent store
            mbar_arrive_expect_tx(&in_bar[1], 2048);
            tma_load_2d_a(t_sub[1], &tmAux, &in_bar[1], colg + 32, row0);
          }
        }
        __syncwarp();
      }
      float lnb_mean = 0.f, lnb_rstd = 0.f;
      if (EPI == EPI_ADD_LNBWD && have0 && row0 + lane < p.M) {
        lnb_mean = __ldg(p.lnb.mean + row0 + lane);
        lnb_rstd = __ldg(p.lnb.rstd + row0 + lane);
      }
      mbar_wait(&tmem_full[as], aph);
      if (it == 0 && warp == 2) mark(6);
      tc_fence_after();
      if (!have0) {   // nothing to write: only keep the TMEM protocol alive
        release_tmem();
        continue;
      }
      if (EPI == EPI_ADD_LNBWD) {   // dy = acc (+ bias) + addend (written over the addend tile), then the LayerNorm backward
        // The saved-sum tiles land in the operand ring: this kind runs one work item per CTA (planner), so once the
        // accumulator is complete nothing else touches the ring.
        const uint32_t s_sub[2] = {smem_u32(smem) + uint32_t(warp - 2) * 4096, smem_u32(smem) + uint32_t(warp - 2) * 4096 + 2048};
        uint64_t* s_bar = &s_bars[(warp - 2) * 2];
        if (lane == 0) {
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            mbar_arrive_expect_tx(&s_bar[sub], 2048);
            tma_load_2d_a(s_sub[sub], &tmD, &s_bar[sub], colg + sub * 32, row0);
          }
        }
        __syncwarp();
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t r[32];
          tmem_ld32(tmem_row + uint32_t(sub * 32), r);
          mbar_wait(&in_bar[sub], in_ph[sub]);
          in_ph[sub] ^= 1;
          tmem_ld_wait();
          if (sub == 1) release_tmem();
          const int col0 = colg + sub * 32;
          const uint32_t row64 = t_sub[sub] + lane * 64, sw64 = (lane >> 1) & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[8 * i + u]);
            if (p.bias != nullptr) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 8 * i));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 8 * i + 4));
              add_pair(v[0], v[1], b0.x, b0.y); add_pair(v[2], v[3], b0.z, b0.w);
              add_pair(v[4], v[5], b1.x, b1.y); add_pair(v[6], v[7], b1.z, b1.w);
            }
            const uint4 in = lds128(row64 + ((i ^ sw64) << 4));
            const float2 a = unpack_bf16(in.x), b = unpack_bf16(in.y), c = unpack_bf16(in.z), d = unpack_bf16(in.w);
            add_pair(v[0], v[1], a.x, a.y); add_pair(v[2], v[3], b.x, b.y);
            add_pair(v[4], v[5], c.x, c.y); add_pair(v[6], v[7], d.x, d.y);
            sts128(row64 + ((i ^ sw64) << 4), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          }
        }
        mbar_wait(&s_bar[0], 0);
        mbar_wait(&s_bar[1], 0);
        lnbwd_rows(p.lnb, t_sub, s_sub, lnb_mean, lnb_rstd, colg, q, cg, lane, ln_stats, 1 + q, 5, 32 * EPI_WARPS, [&](int sub) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d_a(&tmC, t_sub[sub], colg + sub * 32, row0);
            tma_store_commit();
          }
        });
        continue;
      }
      if (EPI == EPI_ADD_LN) {   // N = BN = 256: the four warps with this warp's TMEM quadrant hold complete rows between them
        uint32_t sreg[2][16];   // the bf16-rounded sum, two columns per register
        float rsum = 0.f;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t r[32];
          tmem_ld32(tmem_row + uint32_t(sub * 32), r);
          mbar_wait(&in_bar[sub], in_ph[sub]);
          in_ph[sub] ^= 1;
          tmem_ld_wait();
          if (sub == 1) release_tmem();
          const int col0 = colg + sub * 32;
          const uint32_t row64 = t_sub[sub] + lane * 64, sw64 = (lane >> 1) & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(r[8 * i + u]);
            if (p.bias != nullptr) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 8 * i));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 8 * i + 4));
              add_pair(v[0], v[1], b0.x, b0.y); add_pair(v[2], v[3], b0.z, b0.w);
              add_pair(v[4], v[5], b1.x, b1.y); add_pair(v[6], v[7], b1.z, b1.w);
            }
            const uint32_t off = (i ^ sw64) << 4;
            const uint4 in = lds128(row64 + off);
            const float2 a = unpack_bf16(in.x), b = unpack_bf16(in.y), c = unpack_bf16(in.z), d = unpack_bf16(in.w);
            add_pair(v[0], v[1], a.x, a.y); add_pair(v[2], v[3], b.x, b.y);
            add_pair(v[4], v[5], c.x, c.y); add_pair(v[6], v[7], d.x, d.y);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t pk = pack_bf16(v[2 * u], v[2 * u + 1]);
              sreg[sub][4 * i + u] = pk;
              const float2 f = unpack_bf16(pk);   // statistics of the values the backward will read
              rsum += f.x + f.y;
            }
            sts128(row64 + off, sreg[sub][4 * i], sreg[sub][4 * i + 1], sreg[sub][4 * i + 2], sreg[sub][4 * i + 3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d_a(&tmC, t_sub[sub], col0, row0);
            tma_store_commit();
          }
        }
        // row statistics across the four column groups (two-pass, like add_ln_fwd_kernel): named barrier 1 + q, 128 threads
        const int rl = q * 32 + lane;
        ln_stats[cg * 128 + rl].x = rsum;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
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
        asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
        const float var = (ln_stats[rl].y + ln_stats[128 + rl].y + ln_stats[256 + rl].y + ln_stats[384 + rl].y) * (1.f / 256.f);
        const float rstd = rsqrtf(var + p.ln_eps);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");   // everyone has read the partials: the slots may be reused
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
            tma_store_2d_a(&tmD, t_sub[sub], col0, row0);
            tma_store_commit();
          }
        }
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
        if (BG && sub == 0 && cg == 0 && n0 == 0 && p.bias_grad != nullptr) {   // column 0 of the all-ones product
          tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(BN), r);
          tmem_ld_wait();
          if (row0 + lane < p.M) atomicAdd(p.bias_grad + row0 + lane, __uint_as_float(r[0]));
        }
        tmem_ld32(tmem_row + uint32_t(sub * 32), r);
        if (EPI == EPI_ADD || EPI == EPI_GELU_BWD) {
          mbar_wait(&in_bar[sub], in_ph[sub]);
          in_ph[sub] ^= 1;
        }
        tmem_ld_wait();
        if (it == 0 && warp == 2 && sub == 0) mark(7);
        if (sub == 1) release_tmem();
        if (EPI == EPI_GENERIC) {
          const int row = row0 + lane;
          if (row < p.M) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            const bool vec_ok = ((p.ldc & 7) == 0) && ((p.N & 7) == 0) && (p.addend == nullptr || (p.ld_addend & 7) == 0) &&
                                (p.aux == nullptr || (p.ld_aux & 7) == 0);
            gemm_epilogue_chunk(p, row, col0, v, vec_ok);
          }
          continue;
        }
        // the staging tile(s) of this half are free once the bulk stores that last read them have done so
        if (EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_F32) {
          if (lane == 0) {
            if (EPI == EPI_F32) tma_store_wait_read_n<0>(); else tma_store_wait_read_n<1>();
          }
          __syncwarp();
        }
        gemm_epilogue_sub<EPI>(p, col0, lane, r, t_sub[sub], t_aux[sub]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (EPI == EPI_GELU) {   // one bulk group per store: the rotation waits on "all but the latest"
            tma_store_2d_a(&tmAux, t_aux[sub], col0, row0);
            tma_store_commit();
          }
          if (EPI == EPI_F32 && p.manual_out) {
            // (copied out below by the whole warp)
          } else {
            if (EPI == EPI_F32 && p.out_kind == OUT_F32_ATOMIC) tma_reduce_add_2d_a(&tmC, t_sub[sub], col0, row0);
            else tma_store_2d_a(&tmC, t_sub[sub], col0, row0);
            tma_store_commit();
          }
        }
        if (EPI == EPI_F32 && p.manual_out) {
          // lanes 0..15: row r, lanes 16..31: row r + 1; lane l: columns col0 + 2 (l & 15) .. + 1 of the staged tile (SWIZZLE_128B
          // layout: 16-byte chunk c of row r at chunk c ^ (r & 7))
          const int half = lane >> 4, l2 = (lane & 15) * 2, col = col0 + l2;
          float* cbase = reinterpret_cast<float*>(p.C);
#pragma unroll 4
          for (int r = 0; r < 32; r += 2) {
            const int rr = r + half, row = row0 + rr;
            const uint32_t addr = t_sub[sub] + rr * 128 + ((uint32_t(l2 >> 2) ^ uint32_t(rr & 7)) << 4) + (l2 & 3) * 4;
            float2 val;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(val.x), "=f"(val.y) : "r"(addr));
            if (row < p.M && col < p.N) {
              float* dst = cbase + (int64_t)row * p.ldc + col;
              if (col + 1 < p.N) {
                if (p.out_kind == OUT_F32_ATOMIC) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(val.x), "f"(val.y) : "memory");
                else *reinterpret_cast<float2*>(dst) = val;
              } else {
                if (p.out_kind == OUT_F32_ATOMIC) atomicAdd(dst, val.x);
                else *dst = val.x;
              }
            }
          }
          __syncwarp();   // the tile is free for the next half
        }
        if (it == 0 && warp == 2 && sub == 0) mark(8);
      }
    }
    if (warp == 2) mark(9);
    if (EPI != EPI_GENERIC && lane == 0) tma_store_wait_read();   // the staging tiles have been read: shared memory may go; the writes complete with the kernel
    if (warp == 2) mark(10);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_exit();   // no CTA exits while its peer may still read its shared memory / signal its barriers
  if (warp == 0) mark(11);
  if (warp == 1) {
    if (CL == 1) tmem_dealloc<TMEM_COLS>(tmem_base); else tmem_dealloc_cg2<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------- host side
int make_tensor_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_pitch_bytes,
                       uint32_t box_inner, uint32_t box_outer);
int make_tensor_map_3d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                       uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2);
// Builds a plan: tensor maps + grid.  split_k <= 0 lets the planner choose (wgrad only).
// Optional LayerNorm behind a +residual GEMM (EPI_ADD_LN): y = LN(C) * gamma + beta, statistics of the bf16-rounded C.
struct GemmLN {
  const float* gamma;
  const float* beta;
  float* mean;
  float* rstd;
  bf16* y;
  int64_t ldy;
  float eps;
};
int gemm_plan(GemmPlan* plan, int mode, const bf16* A, int64_t lda, const bf16* B, int64_t ldb, void* C, int64_t ldc,
              int M, int N, int K, const float* bias, const bf16* addend, int64_t ld_addend, int out_kind, int act,
              bf16* aux, int64_t ld_aux, int split_k, float* bias_grad = nullptr, const GemmLN* ln = nullptr,
              const LnBwdArgs* lnb = nullptr);
int gemm_launch(const GemmPlan* plan, cudaStream_t stream);
bool gemm_lnbwd_supported(int M);
int gemm_init_attributes();

}  // namespace kit

// softmax(Q K^T / sqrt(d) + mask) V, forward and backward, with the attention masks of the
// reference synthesised in-kernel from the [B,T] frame mask (model.py:193-202 "repeat-inc",
// A1_train.py:121 float key-padding mask that PyTorch ADDS to the logits) so that no [B*NH,S,S]
// tensor ever exists.  Flash-style: 64 queries x 64 keys per step, online softmax in fp32, bf16
// operands on the tensor cores (warp-level mma.sync m16n8k16 -- the per-head problems are
// 64x64x32, far below one tcgen05 tile; the kernels are bound by operand traffic and softmax ALU).
// The backward recomputes P from the saved log-sum-exp and works on the TRANSPOSED score tile
// (keys as MMA rows) so that dK/dV are warp-private and only dS crosses shared memory.
#include "attention.cuh"

namespace kit {

constexpr int AT = 64;  // tile edge (queries and keys)
constexpr int AT_THREADS = 128;

struct MaskDev {
  const float* frame_mask;
  int64_t frame_mask_stride;
  int flags;
  const float* bias;
  int64_t bias_sb, bias_sh;
};

__device__ __forceinline__ float mask_bias(const MaskDev& m, int b, int h, int i, int j, int Sk, float fm_j) {
  float r = 0.f;
  if ((m.flags & KIT_MASK_REPEAT_INC) && j > i && fm_j == 1.f) r = -INFINITY;
  if ((m.flags & KIT_MASK_TRIANGLE) && j > i) r = -INFINITY;
  if (m.flags & KIT_MASK_KEYPAD_ADD) r += fm_j;
  if (m.bias != nullptr) r += m.bias[(int64_t)b * m.bias_sb + (int64_t)h * m.bias_sh + (int64_t)i * Sk + j];
  return r;
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

// rows [r0, r0+64) x D columns of a bf16 matrix -> shared [64][D+8], zero rows beyond nrows
template <int D>
__device__ __forceinline__ void load_tile(bf16 (*dst)[D + 8], const bf16* base, int64_t ld, int r0, int nrows) {
  constexpr int VPR = D / 8;
  for (int idx = threadIdx.x; idx < AT * VPR; idx += AT_THREADS) {
    const int r = idx / VPR, v = idx % VPR;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r0 + r < nrows) val = *reinterpret_cast<const uint4*>(base + (int64_t)(r0 + r) * ld + v * 8);
    *reinterpret_cast<uint4*>(&dst[r][v * 8]) = val;
  }
}
template <int D>
__device__ __forceinline__ uint32_t lds_pair(const bf16 (*t)[D + 8], int r, int c) {
  return *reinterpret_cast<const uint32_t*>(&t[r][c]);
}

// ---------------------------------------------------------------------------------------- forward
template <int D>
struct FwdSmem {
  bf16 K[AT][D + 8];
  bf16 V[AT][D + 8];
  float fm[AT];
};

template <int D>
__global__ void __launch_bounds__(AT_THREADS) attn_fwd_kernel(const bf16* __restrict__ q, int64_t ldq,
                                                              const bf16* __restrict__ k, int64_t ldk,
                                                              const bf16* __restrict__ v, int64_t ldv,
                                                              bf16* __restrict__ out, int64_t ldo,
                                                              float* __restrict__ lse, int NH, int Sq, int Sk,
                                                              float scale, MaskDev mask) {
  pdl_grid_sync();
  __shared__ __align__(16) FwdSmem<D> s;
  constexpr int KS = D / 16, NT = D / 8;
  const int b = blockIdx.y / NH, h = blockIdx.y % NH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * AT + warp * 16;
  const bf16* qb = q + (int64_t)b * Sq * ldq + h * D;
  const bf16* kb = k + (int64_t)b * Sk * ldk + h * D;
  const bf16* vb = v + (int64_t)b * Sk * ldv + h * D;

  uint32_t aq[KS][4];
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    const int r0 = q0 + g, r1 = q0 + g + 8, c = kk * 16 + 2 * t;
    aq[kk][0] = r0 < Sq ? *reinterpret_cast<const uint32_t*>(qb + (int64_t)r0 * ldq + c) : 0u;
    aq[kk][1] = r1 < Sq ? *reinterpret_cast<const uint32_t*>(qb + (int64_t)r1 * ldq + c) : 0u;
    aq[kk][2] = r0 < Sq ? *reinterpret_cast<const uint32_t*>(qb + (int64_t)r0 * ldq + c + 8) : 0u;
    aq[kk][3] = r1 < Sq ? *reinterpret_cast<const uint32_t*>(qb + (int64_t)r1 * ldq + c + 8) : 0u;
  }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float o[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;

  for (int k0 = 0; k0 < Sk; k0 += AT) {
    __syncthreads();
    load_tile<D>(s.K, kb, ldk, k0, Sk);
    load_tile<D>(s.V, vb, ldv, k0, Sk);
    if (threadIdx.x < AT) {
      const int j = k0 + threadIdx.x;
      s.fm[threadIdx.x] = (mask.frame_mask != nullptr && j < Sk) ? mask.frame_mask[(int64_t)b * mask.frame_mask_stride + j] : 0.f;
    }
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        mma16816(acc[j], aq[kk], lds_pair<D>(s.K, j * 8 + g, kk * 16 + 2 * t), lds_pair<D>(s.K, j * 8 + g, kk * 16 + 2 * t + 8));
    // scale + mask, online softmax for rows (q0+g) and (q0+g+8)
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qi = q0 + g + 8 * r;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int kl = j * 8 + 2 * t + c, kj = k0 + kl;
          float x = -INFINITY;
          if (kj < Sk && qi < Sq) x = acc[j][2 * r + c] * scale + mask_bias(mask, b, h, qi, kj, Sk, s.fm[kl]);
          acc[j][2 * r + c] = x;
          mx = fmaxf(mx, x);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run[r], mx);
      const float m_ref = (m_new == -INFINITY) ? 0.f : m_new;
      corr[r] = __expf(m_run[r] - m_ref);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float p = __expf(acc[j][2 * r + c] - m_ref);
          acc[j][2 * r + c] = p;
          rs += p;
        }
      l_run[r] = l_run[r] * corr[r] + rs;   // per-thread partial; quad-reduced at the end
      m_run[r] = m_new;
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      o[j][0] *= corr[0]; o[j][1] *= corr[0];
      o[j][2] *= corr[1]; o[j][3] *= corr[1];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {   // 16 keys per step
      uint32_t ap[4];
      ap[0] = pack_bf16(acc[2 * kk][0], acc[2 * kk][1]);
      ap[1] = pack_bf16(acc[2 * kk][2], acc[2 * kk][3]);
      ap[2] = pack_bf16(acc[2 * kk + 1][0], acc[2 * kk + 1][1]);
      ap[3] = pack_bf16(acc[2 * kk + 1][2], acc[2 * kk + 1][3]);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, &s.V[kk * 16 + (lane & 15)][j * 8]);
        mma16816(o[j], ap, b0, b1);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = l_run[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const int qi = q0 + g + 8 * r;
    if (qi >= Sq) continue;
    const float inv = 1.f / l;
    bf16* op = out + ((int64_t)b * Sq + qi) * ldo + h * D;
#pragma unroll
    for (int j = 0; j < NT; ++j)
      *reinterpret_cast<uint32_t*>(op + j * 8 + 2 * t) = pack_bf16(o[j][2 * r] * inv, o[j][2 * r + 1] * inv);
    if (t == 0 && lse != nullptr) lse[((int64_t)b * NH + h) * Sq + qi] = m_run[r] + __logf(l);
  }
}

// ---------------------------------------------------------------------------------------- backward
template <int D>
struct BwdSmem {
  bf16 K[AT][D + 8];
  bf16 V[AT][D + 8];
  bf16 Q[AT][D + 8];
  bf16 dO[AT][D + 8];
  bf16 dS[AT][AT + 8];  // [key][query]
  float fm[AT];         // frame mask of the keys
  float lse[AT];        // per query
  float delta[AT];      // per query: sum_c dO*O
};

// One CTA per (key tile, b*h): warp w owns keys 16w..16w+15 of the tile.  dQ is written directly when
// there is a single key tile, else accumulated with fp32 atomics into dq_acc.
template <int D>
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_kernel(
    const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk, const bf16* __restrict__ v,
    int64_t ldv, const bf16* __restrict__ o, int64_t ldo, const bf16* __restrict__ dout, int64_t ld_do,
    const float* __restrict__ lse, bf16* __restrict__ dq, int64_t ld_dq, bf16* __restrict__ dk, int64_t ld_dk,
    bf16* __restrict__ dv, int64_t ld_dv, float* __restrict__ dq_acc, int NH, int Sq, int Sk, float scale, MaskDev mask) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BwdSmem<D>& s = *reinterpret_cast<BwdSmem<D>*>(smem_raw);
  constexpr int KS = D / 16, NT = D / 8;
  const int b = blockIdx.y / NH, h = blockIdx.y % NH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int k0 = blockIdx.x * AT;
  const bool single_tile = gridDim.x == 1;
  const bf16* qb = q + (int64_t)b * Sq * ldq + h * D;
  const bf16* kb = k + (int64_t)b * Sk * ldk + h * D;
  const bf16* vb = v + (int64_t)b * Sk * ldv + h * D;
  const bf16* ob = o + (int64_t)b * Sq * ldo + h * D;
  const bf16* dob = dout + (int64_t)b * Sq * ld_do + h * D;
  const float* lse_g = lse + ((int64_t)b * NH + h) * Sq;

  load_tile<D>(s.K, kb, ldk, k0, Sk);
  load_tile<D>(s.V, vb, ldv, k0, Sk);
  if (threadIdx.x < AT) {
    const int j = k0 + threadIdx.x;
    s.fm[threadIdx.x] = (mask.frame_mask != nullptr && j < Sk) ? mask.frame_mask[(int64_t)b * mask.frame_mask_stride + j] : 0.f;
  }
  __syncthreads();
  // A fragments of this warp's 16 keys: K (for S^T = K Q^T) and V (for dP^T = V dO^T)
  uint32_t ak[KS][4], av[KS][4];
  const int kr = warp * 16 + g;
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    const int c = kk * 16 + 2 * t;
    ak[kk][0] = lds_pair<D>(s.K, kr, c);     ak[kk][1] = lds_pair<D>(s.K, kr + 8, c);
    ak[kk][2] = lds_pair<D>(s.K, kr, c + 8); ak[kk][3] = lds_pair<D>(s.K, kr + 8, c + 8);
    av[kk][0] = lds_pair<D>(s.V, kr, c);     av[kk][1] = lds_pair<D>(s.V, kr + 8, c);
    av[kk][2] = lds_pair<D>(s.V, kr, c + 8); av[kk][3] = lds_pair<D>(s.V, kr + 8, c + 8);
  }
  float dk_acc[NT][4], dv_acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    dk_acc[j][0] = dk_acc[j][1] = dk_acc[j][2] = dk_acc[j][3] = 0.f;
    dv_acc[j][0] = dv_acc[j][1] = dv_acc[j][2] = dv_acc[j][3] = 0.f;
  }

  for (int q0 = 0; q0 < Sq; q0 += AT) {
    __syncthreads();   // previous iteration's readers of Q / dO / dS are done
    load_tile<D>(s.Q, qb, ldq, q0, Sq);
    load_tile<D>(s.dO, dob, ld_do, q0, Sq);
    {  // delta_i = sum_c dO[i,c] O[i,c]; two threads per query row
      const int r = threadIdx.x >> 1, hf = threadIdx.x & 1;
      float dsum = 0.f;
      if (q0 + r < Sq) {
        for (int c = hf * (D / 2); c < (hf + 1) * (D / 2); c += 8) {
          float a[8], gg[8];
          load8(ob + (int64_t)(q0 + r) * ldo + c, a);
          load8(dob + (int64_t)(q0 + r) * ld_do + c, gg);
#pragma unroll
          for (int u = 0; u < 8; ++u) dsum = fmaf(a[u], gg[u], dsum);
        }
      }
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
      if (hf == 0) {
        s.delta[r] = dsum;
        s.lse[r] = (q0 + r < Sq) ? lse_g[q0 + r] : 0.f;
      }
    }
    __syncthreads();
    // S^T and dP^T for (16 keys of this warp) x (64 queries)
    float st[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < KS; ++kk)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = kk * 16 + 2 * t;
        mma16816(st[j], ak[kk], lds_pair<D>(s.Q, j * 8 + g, c), lds_pair<D>(s.Q, j * 8 + g, c + 8));
        mma16816(dp[j], av[kk], lds_pair<D>(s.dO, j * 8 + g, c), lds_pair<D>(s.dO, j * 8 + g, c + 8));
      }
    // P^T = exp(S^T*scale + bias - lse[query]);  dS^T = P^T * (dP^T - delta[query]) * scale
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kl = warp * 16 + g + 8 * (e >> 1), kj = k0 + kl;   // key
        const int ql = j * 8 + 2 * t + (e & 1), qi = q0 + ql;        // query
        float p = 0.f;
        if (kj < Sk && qi < Sq) {
          const float x = st[j][e] * scale + mask_bias(mask, b, h, qi, kj, Sk, s.fm[kl]);
          p = __expf(x - s.lse[ql]);
        }
        st[j][e] = p;
        dp[j][e] = p * (dp[j][e] - s.delta[ql]) * scale;
      }
    // dV += P^T dO ; dK += dS^T Q      (reduction over the 64 queries, 16 per k-step)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ap[4], ad[4];
      ap[0] = pack_bf16(st[2 * kk][0], st[2 * kk][1]);         ap[1] = pack_bf16(st[2 * kk][2], st[2 * kk][3]);
      ap[2] = pack_bf16(st[2 * kk + 1][0], st[2 * kk + 1][1]); ap[3] = pack_bf16(st[2 * kk + 1][2], st[2 * kk + 1][3]);
      ad[0] = pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);         ad[1] = pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
      ad[2] = pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]); ad[3] = pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, &s.dO[kk * 16 + (lane & 15)][j * 8]);
        mma16816(dv_acc[j], ap, b0, b1);
        ldsm_x2_trans(b0, b1, &s.Q[kk * 16 + (lane & 15)][j * 8]);
        mma16816(dk_acc[j], ad, b0, b1);
      }
    }
    // dS^T (bf16) -> shared [key][query]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<uint32_t*>(&s.dS[warp * 16 + g][j * 8 + 2 * t]) = pack_bf16(dp[j][0], dp[j][1]);
      *reinterpret_cast<uint32_t*>(&s.dS[warp * 16 + g + 8][j * 8 + 2 * t]) = pack_bf16(dp[j][2], dp[j][3]);
    }
    __syncthreads();
    // dQ (16 queries of this warp) = dS (16 x 64 keys) K (64 x D)
    float dq_acc_r[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) dq_acc_r[j][0] = dq_acc_r[j][1] = dq_acc_r[j][2] = dq_acc_r[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {   // 16 keys per step
      uint32_t a[4];
      // 8x8 blocks (transposed on load): lanes 0-7 keys 0-7/queries 0-7, 8-15 queries 8-15, 16-23 keys 8-15, 24-31 both
      ldsm_x4_trans(a, &s.dS[kk * 16 + (lane & 7) + ((lane >> 4) << 3)][warp * 16 + (((lane >> 3) & 1) << 3)]);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, &s.K[kk * 16 + (lane & 15)][j * 8]);
        mma16816(dq_acc_r[j], a, b0, b1);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qi = q0 + warp * 16 + g + 8 * r;
      if (qi >= Sq) continue;
      if (single_tile) {
        bf16* dqp = dq + ((int64_t)b * Sq + qi) * ld_dq + h * D;
#pragma unroll
        for (int j = 0; j < NT; ++j)
          *reinterpret_cast<uint32_t*>(dqp + j * 8 + 2 * t) = pack_bf16(dq_acc_r[j][2 * r], dq_acc_r[j][2 * r + 1]);
      } else {
        float* ap = dq_acc + ((int64_t)b * Sq + qi) * (int64_t)(NH * D) + h * D;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          atomicAdd(ap + j * 8 + 2 * t, dq_acc_r[j][2 * r]);
          atomicAdd(ap + j * 8 + 2 * t + 1, dq_acc_r[j][2 * r + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int kj = k0 + warp * 16 + g + 8 * r;
    if (kj >= Sk) continue;
    bf16* dkp = dk + ((int64_t)b * Sk + kj) * ld_dk + h * D;
    bf16* dvp = dv + ((int64_t)b * Sk + kj) * ld_dv + h * D;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      *reinterpret_cast<uint32_t*>(dkp + j * 8 + 2 * t) = pack_bf16(dk_acc[j][2 * r], dk_acc[j][2 * r + 1]);
      *reinterpret_cast<uint32_t*>(dvp + j * 8 + 2 * t) = pack_bf16(dv_acc[j][2 * r], dv_acc[j][2 * r + 1]);
    }
  }
}

// fp32 dQ accumulator [rows, width] -> bf16 dq (row pitch ld_dq), multi-key-tile case only
__global__ void dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, int64_t ld_dq, int64_t rows, int width) {
  pdl_grid_sync();
  const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (idx >= rows * width) return;
  const int64_t r = idx / width;
  const int c = (int)(idx % width);
  float f[8];
  load8(acc + idx, f);
  store8(dq + r * ld_dq + c, f);
}


// ------------------------------------------------------------------------ long-sequence backward (more than one key tile)
constexpr float LOG2E_F = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx_f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// delta[b, h, i] = sum_c dO[i, c] O[i, c], once per call (the one-CTA-per-key-tile kernel above recomputes it in every key tile)
template <int D>
__global__ void attn_delta_kernel(const bf16* __restrict__ o, int64_t ldo, const bf16* __restrict__ dout, int64_t ld_do,
                                  float* __restrict__ delta, int NH, int Sq, int64_t n_rows) {
  pdl_grid_sync();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (b * Sq + i) * NH + h
  if (idx >= n_rows * NH) return;
  const int64_t row = idx / NH;
  const int h = (int)(idx - row * NH);
  const bf16* op = o + row * ldo + h * D;
  const bf16* gp = dout + row * ld_do + h * D;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < D; c += 8) {
    float a[8], g[8];
    load8(op + c, a);
    load8(gp + c, g);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc = fmaf(a[u], g[u], acc);
  }
  const int64_t b = row / Sq, i = row - b * Sq;
  delta[(b * NH + h) * Sq + i] = acc;
}

template <int D>
struct BwdStreamSmem {
  bf16 K[AT][D + 8];
  bf16 V[AT][D + 8];
  bf16 Q[2][AT][D + 8];
  bf16 dO[2][AT][D + 8];
  bf16 dS[AT][AT + 8];  // [key][query]
  float fm[AT];
  float kbias[AT];   // per key: additive bias * log2(e)            (key padding mask)
  float kcut[AT];    // per key: 1 = masked for every query before it (repeat-inc / triangle), else 0
  float lse[2][AT];
  float delta[2][AT];
};

// Same mathematics and warp layout as attn_bwd_kernel (one CTA per (key tile, batch * head), warp w owns keys 16w .. 16w + 15,
// the transposed score tile), but the query tiles STREAM: while tile i is being processed, cp.async is filling the other
// buffer with Q / dO / lse / delta of tile i + 1; delta comes from attn_delta_kernel; dQ partials leave as 8-byte vector
// reductions.  configs[4] (T = 512, d = 64): 1.2 ms -> see profiles/r01c_summary.md.
template <int D>
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_stream_kernel(
    const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk, const bf16* __restrict__ v,
    int64_t ldv, const bf16* __restrict__ dout, int64_t ld_do, const float* __restrict__ lse,
    const float* __restrict__ delta, bf16* __restrict__ dk, int64_t ld_dk, bf16* __restrict__ dv, int64_t ld_dv,
    float* __restrict__ dq_acc, int NH, int Sq, int Sk, float scale, MaskDev mask) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BwdStreamSmem<D>& s = *reinterpret_cast<BwdStreamSmem<D>*>(smem_raw);
  constexpr int KS = D / 16, NT = D / 8, VPR = D / 8;
  const int b = blockIdx.y / NH, h = blockIdx.y % NH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int k0 = blockIdx.x * AT;
  const bf16* qb = q + (int64_t)b * Sq * ldq + h * D;
  const bf16* kb = k + (int64_t)b * Sk * ldk + h * D;
  const bf16* vb = v + (int64_t)b * Sk * ldv + h * D;
  const bf16* dob = dout + (int64_t)b * Sq * ld_do + h * D;
  const float* lse_g = lse + ((int64_t)b * NH + h) * Sq;
  const float* delta_g = delta + ((int64_t)b * NH + h) * Sq;

  auto prefetch = [&](int q0, int buf) {   // Q / dO rows [q0, q0 + 64) and their lse / delta -> buffer buf (zero beyond Sq)
    for (int idx = threadIdx.x; idx < AT * VPR; idx += AT_THREADS) {
      const int r = idx / VPR, c = (idx % VPR) * 8;
      const bool ok = q0 + r < Sq;
      const int64_t row = ok ? q0 + r : 0;
      cp_async16(&s.Q[buf][r][c], qb + row * ldq + c, ok);
      cp_async16(&s.dO[buf][r][c], dob + row * ld_do + c, ok);
    }
    if (threadIdx.x < AT) {
      const bool ok = q0 + (int)threadIdx.x < Sq;
      cp_async4(&s.lse[buf][threadIdx.x], lse_g + (ok ? q0 + threadIdx.x : 0), ok);
    } else {
      const int r = threadIdx.x - AT;
      const bool ok = q0 + r < Sq;
      cp_async4(&s.delta[buf][r], delta_g + (ok ? q0 + r : 0), ok);
    }
    cp_async_commit();
  };
  prefetch(0, 0);
  load_tile<D>(s.K, kb, ldk, k0, Sk);
  load_tile<D>(s.V, vb, ldv, k0, Sk);
  if (threadIdx.x < AT) {
    const int j = k0 + threadIdx.x;
    const float fmj = (mask.frame_mask != nullptr && j < Sk) ? mask.frame_mask[(int64_t)b * mask.frame_mask_stride + j] : 0.f;
    s.fm[threadIdx.x] = fmj;
    s.kbias[threadIdx.x] = (mask.flags & KIT_MASK_KEYPAD_ADD) ? fmj * LOG2E_F : 0.f;
    s.kcut[threadIdx.x] = (((mask.flags & KIT_MASK_REPEAT_INC) && fmj == 1.f) || (mask.flags & KIT_MASK_TRIANGLE)) ? 1.f : 0.f;
  }
  __syncthreads();
  uint32_t ak[KS][4], av[KS][4];
  const int kr = warp * 16 + g;
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    const int c = kk * 16 + 2 * t;
    ak[kk][0] = lds_pair<D>(s.K, kr, c);     ak[kk][1] = lds_pair<D>(s.K, kr + 8, c);
    ak[kk][2] = lds_pair<D>(s.K, kr, c + 8); ak[kk][3] = lds_pair<D>(s.K, kr + 8, c + 8);
    av[kk][0] = lds_pair<D>(s.V, kr, c);     av[kk][1] = lds_pair<D>(s.V, kr + 8, c);
    av[kk][2] = lds_pair<D>(s.V, kr, c + 8); av[kk][3] = lds_pair<D>(s.V, kr + 8, c + 8);
  }
  float dk_acc[NT][4], dv_acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    dk_acc[j][0] = dk_acc[j][1] = dk_acc[j][2] = dk_acc[j][3] = 0.f;
    dv_acc[j][0] = dv_acc[j][1] = dv_acc[j][2] = dv_acc[j][3] = 0.f;
  }
  // this thread's two keys (rows g and g + 8 of the warp's 16): folded mask terms, base-2 softmax scale
  const bool folded = mask.bias == nullptr;
  const float scale2 = scale * LOG2E_F;
  float kb2[2], kc2[2];
  int kjj[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int kl = warp * 16 + g + 8 * r;
    kb2[r] = s.kbias[kl];
    kc2[r] = s.kcut[kl];
    kjj[r] = k0 + kl;
  }
  const int nq = (Sq + AT - 1) / AT;
  for (int it = 0; it < nq; ++it) {
    const int q0 = it * AT, buf = it & 1;
    if (it + 1 < nq) {
      prefetch(q0 + AT, buf ^ 1);   // its previous readers finished before the barrier that ended iteration it - 1
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();   // tile `it` has landed for every thread
    const bf16 (*sQ)[D + 8] = s.Q[buf];
    const bf16 (*sdO)[D + 8] = s.dO[buf];
    float st[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < KS; ++kk)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = kk * 16 + 2 * t;
        mma16816(st[j], ak[kk], lds_pair<D>(sQ, j * 8 + g, c), lds_pair<D>(sQ, j * 8 + g, c + 8));
        mma16816(dp[j], av[kk], lds_pair<D>(sdO, j * 8 + g, c), lds_pair<D>(sdO, j * 8 + g, c + 8));
      }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kl = warp * 16 + g + 8 * (e >> 1), kj = k0 + kl;   // key
        const int ql = j * 8 + 2 * t + (e & 1), qi = q0 + ql;        // query
        float p = 0.f;
        if (kj < Sk && qi < Sq) {
          if (folded) {   // exp(x - lse) = 2^(s * scale * log2e + bias * log2e - lse * log2e); a cut key is masked for earlier queries
            const int r = e >> 1;
            const bool cut = kc2[r] != 0.f && kjj[r] > qi;
            p = cut ? 0.f : ex2_approx_f(fmaf(st[j][e], scale2, kb2[r]) - s.lse[buf][ql] * LOG2E_F);
          } else {
            const float x = st[j][e] * scale + mask_bias(mask, b, h, qi, kj, Sk, s.fm[kl]);
            p = __expf(x - s.lse[buf][ql]);
          }
        }
        st[j][e] = p;
        dp[j][e] = p * (dp[j][e] - s.delta[buf][ql]) * scale;
      }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ap[4], ad[4];
      ap[0] = pack_bf16(st[2 * kk][0], st[2 * kk][1]);         ap[1] = pack_bf16(st[2 * kk][2], st[2 * kk][3]);
      ap[2] = pack_bf16(st[2 * kk + 1][0], st[2 * kk + 1][1]); ap[3] = pack_bf16(st[2 * kk + 1][2], st[2 * kk + 1][3]);
      ad[0] = pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);         ad[1] = pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
      ad[2] = pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]); ad[3] = pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, &sdO[kk * 16 + (lane & 15)][j * 8]);
        mma16816(dv_acc[j], ap, b0, b1);
        ldsm_x2_trans(b0, b1, &sQ[kk * 16 + (lane & 15)][j * 8]);
        mma16816(dk_acc[j], ad, b0, b1);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<uint32_t*>(&s.dS[warp * 16 + g][j * 8 + 2 * t]) = pack_bf16(dp[j][0], dp[j][1]);
      *reinterpret_cast<uint32_t*>(&s.dS[warp * 16 + g + 8][j * 8 + 2 * t]) = pack_bf16(dp[j][2], dp[j][3]);
    }
    __syncthreads();
    float dq_r[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) dq_r[j][0] = dq_r[j][1] = dq_r[j][2] = dq_r[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      ldsm_x4_trans(a, &s.dS[kk * 16 + (lane & 7) + ((lane >> 4) << 3)][warp * 16 + (((lane >> 3) & 1) << 3)]);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, &s.K[kk * 16 + (lane & 15)][j * 8]);
        mma16816(dq_r[j], a, b0, b1);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qi = q0 + warp * 16 + g + 8 * r;
      if (qi >= Sq) continue;
      float* ap = dq_acc + ((int64_t)b * Sq + qi) * (int64_t)(NH * D) + h * D;
#pragma unroll
      for (int j = 0; j < NT; ++j)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(ap + j * 8 + 2 * t), "f"(dq_r[j][2 * r]), "f"(dq_r[j][2 * r + 1]) : "memory");
    }
    __syncthreads();   // every warp is done with this tile's Q / dO / dS before they are overwritten
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int kj = k0 + warp * 16 + g + 8 * r;
    if (kj >= Sk) continue;
    bf16* dkp = dk + ((int64_t)b * Sk + kj) * ld_dk + h * D;
    bf16* dvp = dv + ((int64_t)b * Sk + kj) * ld_dv + h * D;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      *reinterpret_cast<uint32_t*>(dkp + j * 8 + 2 * t) = pack_bf16(dk_acc[j][2 * r], dk_acc[j][2 * r + 1]);
      *reinterpret_cast<uint32_t*>(dvp + j * 8 + 2 * t) = pack_bf16(dv_acc[j][2 * r], dv_acc[j][2 * r + 1]);
    }
  }
}

// ------------------------------------------------------------------------ single-tile kernels (Sq, Sk <= 64)
// The benchmark shape (T = 64, d = 32) is one 64 x 64 score tile per (batch, head): 4 KB per operand, far too little work
// to hide a load -> compute -> store chain per CTA.  These kernels are PERSISTENT over the (batch, head) units: while a CTA
// computes unit i, cp.async is already filling the other shared-memory buffer with unit i+1, and results leave through
// shared memory as 16-byte coalesced stores.  The mask is folded into two floats per key ({additive bias * log2(e),
// "masked when key > query"}), computed once per unit, so the inner loops carry no per-element flag logic.
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// rows [0, 64) x D columns of head h -> shared [64][D+8] through cp.async, rows >= nrows zero-filled
template <int D>
__device__ __forceinline__ void async_tile(bf16 (*dst)[D + 8], const bf16* base, int64_t ld, int nrows) {
  constexpr int VPR = D / 8;
  for (int idx = threadIdx.x; idx < AT * VPR; idx += AT_THREADS) {
    const int r = idx / VPR, v = idx % VPR;
    const bool ok = r < nrows;
    cp_async16(&dst[r][v * 8], ok ? base + (int64_t)r * ld + v * 8 : base, ok);
  }
}
// shared [64][D+8] rows [0, nrows) -> global rows, 16 bytes per thread
template <int D>
__device__ __forceinline__ void store_tile(bf16* base, int64_t ld, const bf16 (*src)[D + 8], int nrows) {
  constexpr int VPR = D / 8;
  for (int idx = threadIdx.x; idx < AT * VPR; idx += AT_THREADS) {
    const int r = idx / VPR, v = idx % VPR;
    if (r < nrows) *reinterpret_cast<uint4*>(base + (int64_t)r * ld + v * 8) = *reinterpret_cast<const uint4*>(&src[r][v * 8]);
  }
}
// {bias * log2e (or -inf beyond Sk), 1 when the key is masked for every query before it}
__device__ __forceinline__ float2 key_info(const MaskDev& m, int b, int j, int Sk) {
  const float fm = (m.frame_mask != nullptr && j < Sk) ? m.frame_mask[(int64_t)b * m.frame_mask_stride + j] : 0.f;
  float add = (m.flags & KIT_MASK_KEYPAD_ADD) ? fm * LOG2E : 0.f;
  if (j >= Sk) add = -INFINITY;
  const bool inc = ((m.flags & KIT_MASK_REPEAT_INC) && fm == 1.f) || (m.flags & KIT_MASK_TRIANGLE);
  return make_float2(add, inc ? 1.f : 0.f);
}

template <int D>
struct FwdTile {
  bf16 Q[AT][D + 8];   // re-used as the staging tile of O
  bf16 K[AT][D + 8];
  bf16 V[AT][D + 8];
  float2 kinfo[AT];
};

template <int D>
__global__ void __launch_bounds__(AT_THREADS) attn_fwd_tile_kernel(const bf16* __restrict__ q, int64_t ldq,
                                                                   const bf16* __restrict__ k, int64_t ldk,
                                                                   const bf16* __restrict__ v, int64_t ldv,
                                                                   bf16* __restrict__ out, int64_t ldo,
                                                                   float* __restrict__ lse, int NH, int Sq, int Sk,
                                                                   int units, float scale, MaskDev mask) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FwdTile<D>* tiles = reinterpret_cast<FwdTile<D>*>(smem_raw);
  constexpr int KS = D / 16, NT = D / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = warp * 16;
  const float sl2 = scale * LOG2E;
  auto issue = [&](int unit, FwdTile<D>& s) {
    const int b = unit / NH, h = unit % NH;
    async_tile<D>(s.Q, q + (int64_t)b * Sq * ldq + h * D, ldq, Sq);
    async_tile<D>(s.K, k + (int64_t)b * Sk * ldk + h * D, ldk, Sk);
    async_tile<D>(s.V, v + (int64_t)b * Sk * ldv + h * D, ldv, Sk);
    if (threadIdx.x < AT) s.kinfo[threadIdx.x] = key_info(mask, b, threadIdx.x, Sk);
  };
  int unit = blockIdx.x, buf = 0;
  if (unit < units) issue(unit, tiles[0]);
  cp_async_commit();
  for (; unit < units; unit += gridDim.x, buf ^= 1) {
    FwdTile<D>& s = tiles[buf];
    if (unit + (int)gridDim.x < units) issue(unit + gridDim.x, tiles[buf ^ 1]);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int b = unit / NH, h = unit % NH;
    uint32_t aq[KS][4];
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      const int c = kk * 16 + 2 * t;
      aq[kk][0] = lds_pair<D>(s.Q, q0 + g, c);     aq[kk][1] = lds_pair<D>(s.Q, q0 + g + 8, c);
      aq[kk][2] = lds_pair<D>(s.Q, q0 + g, c + 8); aq[kk][3] = lds_pair<D>(s.Q, q0 + g + 8, c + 8);
    }
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        mma16816(acc[j], aq[kk], lds_pair<D>(s.K, j * 8 + g, kk * 16 + 2 * t), lds_pair<D>(s.K, j * 8 + g, kk * 16 + 2 * t + 8));
    // logits (base 2) + mask: key (j*8 + 2t + c) is masked for query (q0 + g + 8r) iff it lies after it and kinfo.y != 0
    const int dq = 2 * t - g - q0;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 ki = *reinterpret_cast<const float4*>(&s.kinfo[j * 8 + 2 * t]);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float x0 = fmaf(acc[j][2 * r], sl2, ki.x), x1 = fmaf(acc[j][2 * r + 1], sl2, ki.z);
        if (ki.y != 0.f && dq > 8 * r - 8 * j) x0 = -INFINITY;
        if (ki.w != 0.f && dq + 1 > 8 * r - 8 * j) x1 = -INFINITY;
        acc[j][2 * r] = x0;
        acc[j][2 * r + 1] = x1;
        mx[r] = fmaxf(mx[r], fmaxf(x0, x1));
      }
    }
    float l[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));   // finite: key 0 is never masked and 0 < Sk
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p0 = ex2_approx(acc[j][2 * r] - mx[r]), p1 = ex2_approx(acc[j][2 * r + 1] - mx[r]);
        acc[j][2 * r] = p0;
        acc[j][2 * r + 1] = p1;
        l[r] += p0 + p1;
      }
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    }
    float o[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {   // 16 keys per step
      uint32_t ap[4];
      ap[0] = pack_bf16(acc[2 * kk][0], acc[2 * kk][1]);
      ap[1] = pack_bf16(acc[2 * kk][2], acc[2 * kk][3]);
      ap[2] = pack_bf16(acc[2 * kk + 1][0], acc[2 * kk + 1][1]);
      ap[3] = pack_bf16(acc[2 * kk + 1][2], acc[2 * kk + 1][3]);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, &s.V[kk * 16 + (lane & 15)][j * 8]);
        mma16816(o[j], ap, b0, b1);
      }
    }
    __syncwarp();   // every lane of this warp has its Q fragments: the warp's 16 Q rows become the O staging rows
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float inv = 1.f / l[r];
      const int ql = q0 + g + 8 * r;
#pragma unroll
      for (int j = 0; j < NT; ++j)
        *reinterpret_cast<uint32_t*>(&s.Q[ql][j * 8 + 2 * t]) = pack_bf16(o[j][2 * r] * inv, o[j][2 * r + 1] * inv);
      if (t == 0 && lse != nullptr && ql < Sq) lse[((int64_t)b * NH + h) * Sq + ql] = (mx[r] + log2f(l[r])) * LN2;
    }
    __syncthreads();
    store_tile<D>(out + (int64_t)b * Sq * ldo + h * D, ldo, s.Q, Sq);
    __syncthreads();   // this buffer is refilled by the next iteration's prefetch
  }
  cp_async_wait<0>();
}

// Longer sequences: the same persistent double-buffered scheme over (batch, head, 64-query tile) units, each streaming its
// 64-key tiles through an online softmax (flash-style); one "step" = one (unit, key tile), and step s+1 is in flight while
// step s computes, across unit boundaries too.
template <int D>
__global__ void __launch_bounds__(AT_THREADS) attn_fwd_stream_kernel(const bf16* __restrict__ q, int64_t ldq,
                                                                     const bf16* __restrict__ k, int64_t ldk,
                                                                     const bf16* __restrict__ v, int64_t ldv,
                                                                     bf16* __restrict__ out, int64_t ldo,
                                                                     float* __restrict__ lse, int NH, int Sq, int Sk,
                                                                     int q_tiles, int k_tiles, int units, float scale,
                                                                     MaskDev mask) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FwdTile<D>* tiles = reinterpret_cast<FwdTile<D>*>(smem_raw);
  constexpr int KS = D / 16, NT = D / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = warp * 16;
  const float sl2 = scale * LOG2E;
  // step -> (unit, key tile); unit -> (batch*head, query tile)
  auto issue = [&](int unit, int kt, FwdTile<D>& s) {
    const int bh = unit / q_tiles, qt = unit - bh * q_tiles;
    const int b = bh / NH, h = bh % NH;
    const int qrow = qt * AT, krow = kt * AT;
    async_tile<D>(s.Q, q + ((int64_t)b * Sq + qrow) * ldq + h * D, ldq, Sq - qrow);
    async_tile<D>(s.K, k + ((int64_t)b * Sk + krow) * ldk + h * D, ldk, Sk - krow);
    async_tile<D>(s.V, v + ((int64_t)b * Sk + krow) * ldv + h * D, ldv, Sk - krow);
    if (threadIdx.x < AT) s.kinfo[threadIdx.x] = key_info(mask, b, krow + threadIdx.x, Sk);
  };
  int unit = blockIdx.x, kt = 0, buf = 0;
  if (unit < units) issue(unit, 0, tiles[0]);
  cp_async_commit();
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float o[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  while (unit < units) {
    FwdTile<D>& s = tiles[buf];
    // next step
    int nunit = unit, nkt = kt + 1;
    if (nkt == k_tiles) {
      nkt = 0;
      nunit = unit + gridDim.x;
    }
    if (nunit < units) issue(nunit, nkt, tiles[buf ^ 1]);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int bh = unit / q_tiles, qt = unit - bh * q_tiles;
    const int b = bh / NH, h = bh % NH;
    uint32_t aq[KS][4];
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      const int c = kk * 16 + 2 * t;
      aq[kk][0] = lds_pair<D>(s.Q, q0 + g, c);     aq[kk][1] = lds_pair<D>(s.Q, q0 + g + 8, c);
      aq[kk][2] = lds_pair<D>(s.Q, q0 + g, c + 8); aq[kk][3] = lds_pair<D>(s.Q, q0 + g + 8, c + 8);
    }
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        mma16816(acc[j], aq[kk], lds_pair<D>(s.K, j * 8 + g, kk * 16 + 2 * t), lds_pair<D>(s.K, j * 8 + g, kk * 16 + 2 * t + 8));
    // key (kt*64 + j*8 + 2t + c) is masked for query (qt*64 + q0 + g + 8r) iff it lies after it and kinfo.y != 0
    const int dq = 2 * t - g - q0 + AT * (kt - qt);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 ki = *reinterpret_cast<const float4*>(&s.kinfo[j * 8 + 2 * t]);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float x0 = fmaf(acc[j][2 * r], sl2, ki.x), x1 = fmaf(acc[j][2 * r + 1], sl2, ki.z);
        if (ki.y != 0.f && dq > 8 * r - 8 * j) x0 = -INFINITY;
        if (ki.w != 0.f && dq + 1 > 8 * r - 8 * j) x1 = -INFINITY;
        acc[j][2 * r] = x0;
        acc[j][2 * r + 1] = x1;
        mx[r] = fmaxf(mx[r], fmaxf(x0, x1));
      }
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      const float m_ref = (m_new == -INFINITY) ? 0.f : m_new;
      corr[r] = ex2_approx(m_run[r] - m_ref);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p0 = ex2_approx(acc[j][2 * r] - m_ref), p1 = ex2_approx(acc[j][2 * r + 1] - m_ref);
        acc[j][2 * r] = p0;
        acc[j][2 * r + 1] = p1;
        rs += p0 + p1;
      }
      l_run[r] = l_run[r] * corr[r] + rs;   // per-thread partial; quad-reduced at the end of the unit
      m_run[r] = m_new;
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      o[j][0] *= corr[0]; o[j][1] *= corr[0];
      o[j][2] *= corr[1]; o[j][3] *= corr[1];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {   // 16 keys per step
      uint32_t ap[4];
      ap[0] = pack_bf16(acc[2 * kk][0], acc[2 * kk][1]);
      ap[1] = pack_bf16(acc[2 * kk][2], acc[2 * kk][3]);
      ap[2] = pack_bf16(acc[2 * kk + 1][0], acc[2 * kk + 1][1]);
      ap[3] = pack_bf16(acc[2 * kk + 1][2], acc[2 * kk + 1][3]);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, &s.V[kk * 16 + (lane & 15)][j * 8]);
        mma16816(o[j], ap, b0, b1);
      }
    }
    if (kt + 1 == k_tiles) {   // unit complete: normalise, stage through this buffer's Q rows, coalesced store
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float l = l_run[r];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const float inv = 1.f / l;
        const int ql = q0 + g + 8 * r, qi = qt * AT + ql;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          *reinterpret_cast<uint32_t*>(&s.Q[ql][j * 8 + 2 * t]) = pack_bf16(o[j][2 * r] * inv, o[j][2 * r + 1] * inv);
          o[j][2 * r] = o[j][2 * r + 1] = 0.f;
        }
        if (t == 0 && lse != nullptr && qi < Sq) lse[((int64_t)b * NH + h) * Sq + qi] = (m_run[r] + log2f(l)) * LN2;
        m_run[r] = -INFINITY;
        l_run[r] = 0.f;
      }
      __syncthreads();
      store_tile<D>(out + ((int64_t)b * Sq + qt * AT) * ldo + h * D, ldo, s.Q, Sq - qt * AT);
    }
    __syncthreads();   // this buffer is refilled by the next iteration's prefetch
    unit = nunit;
    kt = nkt;
    buf ^= 1;
  }
  cp_async_wait<0>();
}

template <int D>
struct BwdTile {
  bf16 Q[AT][D + 8];    // later: dK staging
  bf16 K[AT][D + 8];
  bf16 V[AT][D + 8];    // later: dV staging
  bf16 dO[AT][D + 8];   // later: dQ staging
  bf16 O[AT][D + 8];
  float lse[AT];
  float2 kinfo[AT];
};
template <int D>
struct BwdTileSmem {
  BwdTile<D> t[2];
  bf16 dS[AT][AT + 8];   // [key][query]
  float lse2[AT];        // lse * log2e, +inf for queries >= Sq
  float delta[AT];
};

// warp w owns keys 16w..16w+15 for S^T / dP^T / dK / dV and queries 16w..16w+15 for dQ (as attn_bwd_kernel)
template <int D>
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_tile_kernel(
    const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk, const bf16* __restrict__ v,
    int64_t ldv, const bf16* __restrict__ o, int64_t ldo, const bf16* __restrict__ dout, int64_t ld_do,
    const float* __restrict__ lse, bf16* __restrict__ dq, int64_t ld_dq, bf16* __restrict__ dk, int64_t ld_dk,
    bf16* __restrict__ dv, int64_t ld_dv, int NH, int Sq, int Sk, int units, float scale, MaskDev mask) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BwdTileSmem<D>& sm = *reinterpret_cast<BwdTileSmem<D>*>(smem_raw);
  constexpr int KS = D / 16, NT = D / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const float sl2 = scale * LOG2E;
  auto issue = [&](int unit, BwdTile<D>& s) {
    const int b = unit / NH, h = unit % NH;
    async_tile<D>(s.Q, q + (int64_t)b * Sq * ldq + h * D, ldq, Sq);
    async_tile<D>(s.K, k + (int64_t)b * Sk * ldk + h * D, ldk, Sk);
    async_tile<D>(s.V, v + (int64_t)b * Sk * ldv + h * D, ldv, Sk);
    async_tile<D>(s.dO, dout + (int64_t)b * Sq * ld_do + h * D, ld_do, Sq);
    async_tile<D>(s.O, o + (int64_t)b * Sq * ldo + h * D, ldo, Sq);
    if (threadIdx.x < AT) {
      const float* lg = lse + ((int64_t)b * NH + h) * Sq;
      const bool ok = threadIdx.x < Sq;
      cp_async4(&s.lse[threadIdx.x], ok ? lg + threadIdx.x : lg, ok);
      s.kinfo[threadIdx.x] = key_info(mask, b, threadIdx.x, Sk);
    }
  };
  int unit = blockIdx.x, buf = 0;
  if (unit < units) issue(unit, sm.t[0]);
  cp_async_commit();
  for (; unit < units; unit += gridDim.x, buf ^= 1) {
    BwdTile<D>& s = sm.t[buf];
    if (unit + (int)gridDim.x < units) issue(unit + gridDim.x, sm.t[buf ^ 1]);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int b = unit / NH, h = unit % NH;
    {  // delta_i = sum_c dO[i,c] O[i,c]; two threads per query row
      const int r = threadIdx.x >> 1, hf = threadIdx.x & 1;
      float dsum = 0.f;
#pragma unroll
      for (int c = hf * (D / 2); c < (hf + 1) * (D / 2); c += 8) {
        float a[8], gg[8];
        load8(&s.O[r][c], a);
        load8(&s.dO[r][c], gg);
#pragma unroll
        for (int u = 0; u < 8; ++u) dsum = fmaf(a[u], gg[u], dsum);
      }
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
      if (hf == 0) {
        sm.delta[r] = dsum;
        sm.lse2[r] = (r < Sq) ? s.lse[r] * LOG2E : INFINITY;
      }
    }
    uint32_t ak[KS][4], av[KS][4];
    const int kr = warp * 16 + g;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      const int c = kk * 16 + 2 * t;
      ak[kk][0] = lds_pair<D>(s.K, kr, c);     ak[kk][1] = lds_pair<D>(s.K, kr + 8, c);
      ak[kk][2] = lds_pair<D>(s.K, kr, c + 8); ak[kk][3] = lds_pair<D>(s.K, kr + 8, c + 8);
      av[kk][0] = lds_pair<D>(s.V, kr, c);     av[kk][1] = lds_pair<D>(s.V, kr + 8, c);
      av[kk][2] = lds_pair<D>(s.V, kr, c + 8); av[kk][3] = lds_pair<D>(s.V, kr + 8, c + 8);
    }
    __syncthreads();   // delta / lse2 visible
    float st[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < KS; ++kk)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = kk * 16 + 2 * t;
        mma16816(st[j], ak[kk], lds_pair<D>(s.Q, j * 8 + g, c), lds_pair<D>(s.Q, j * 8 + g, c + 8));
        mma16816(dp[j], av[kk], lds_pair<D>(s.dO, j * 8 + g, c), lds_pair<D>(s.dO, j * 8 + g, c + 8));
      }
    // P^T = 2^(S^T*scale*log2e + bias2 - lse2[query]);  dS^T = P^T * (dP^T - delta[query]) * scale
    const float2 ki0 = s.kinfo[kr], ki1 = s.kinfo[kr + 8];
    const int dkq = kr - 2 * t;   // key kr + 8*rr is after query j*8 + 2t + cc  iff  dkq + 8*rr - cc > 8*j
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 l2 = *reinterpret_cast<const float2*>(&sm.lse2[j * 8 + 2 * t]);
      const float2 de = *reinterpret_cast<const float2*>(&sm.delta[j * 8 + 2 * t]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int rr = e >> 1, cc = e & 1;
        const float2 ki = rr ? ki1 : ki0;
        float x = fmaf(st[j][e], sl2, ki.x) - (cc ? l2.y : l2.x);
        if (ki.y != 0.f && dkq + 8 * rr - cc > 8 * j) x = -INFINITY;
        const float p = ex2_approx(x);
        st[j][e] = p;
        dp[j][e] = p * (dp[j][e] - (cc ? de.y : de.x)) * scale;
      }
    }
    float dk_acc[NT][4], dv_acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      dk_acc[j][0] = dk_acc[j][1] = dk_acc[j][2] = dk_acc[j][3] = 0.f;
      dv_acc[j][0] = dv_acc[j][1] = dv_acc[j][2] = dv_acc[j][3] = 0.f;
    }
    // dV = P^T dO ; dK = dS^T Q      (reduction over the 64 queries, 16 per k-step)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ap[4], ad[4];
      ap[0] = pack_bf16(st[2 * kk][0], st[2 * kk][1]);         ap[1] = pack_bf16(st[2 * kk][2], st[2 * kk][3]);
      ap[2] = pack_bf16(st[2 * kk + 1][0], st[2 * kk + 1][1]); ap[3] = pack_bf16(st[2 * kk + 1][2], st[2 * kk + 1][3]);
      ad[0] = pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);         ad[1] = pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
      ad[2] = pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]); ad[3] = pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, &s.dO[kk * 16 + (lane & 15)][j * 8]);
        mma16816(dv_acc[j], ap, b0, b1);
        ldsm_x2_trans(b0, b1, &s.Q[kk * 16 + (lane & 15)][j * 8]);
        mma16816(dk_acc[j], ad, b0, b1);
      }
    }
    // dS^T (bf16) -> shared [key][query]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<uint32_t*>(&sm.dS[warp * 16 + g][j * 8 + 2 * t]) = pack_bf16(dp[j][0], dp[j][1]);
      *reinterpret_cast<uint32_t*>(&sm.dS[warp * 16 + g + 8][j * 8 + 2 * t]) = pack_bf16(dp[j][2], dp[j][3]);
    }
    __syncthreads();   // dS complete; nobody reads Q / dO / V any more
    // dK, dV (this warp's 16 keys) -> the dead Q / V tiles
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        *reinterpret_cast<uint32_t*>(&s.Q[kr + 8 * r][j * 8 + 2 * t]) = pack_bf16(dk_acc[j][2 * r], dk_acc[j][2 * r + 1]);
        *reinterpret_cast<uint32_t*>(&s.V[kr + 8 * r][j * 8 + 2 * t]) = pack_bf16(dv_acc[j][2 * r], dv_acc[j][2 * r + 1]);
      }
    }
    // dQ (16 queries of this warp) = dS (16 x 64 keys) K (64 x D)
    float dq_r[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) dq_r[j][0] = dq_r[j][1] = dq_r[j][2] = dq_r[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      ldsm_x4_trans(a, &sm.dS[kk * 16 + (lane & 7) + ((lane >> 4) << 3)][warp * 16 + (((lane >> 3) & 1) << 3)]);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, &s.K[kk * 16 + (lane & 15)][j * 8]);
        mma16816(dq_r[j], a, b0, b1);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int j = 0; j < NT; ++j)
        *reinterpret_cast<uint32_t*>(&s.dO[warp * 16 + g + 8 * r][j * 8 + 2 * t]) = pack_bf16(dq_r[j][2 * r], dq_r[j][2 * r + 1]);
    }
    __syncthreads();
    store_tile<D>(dq + (int64_t)b * Sq * ld_dq + h * D, ld_dq, s.dO, Sq);
    store_tile<D>(dk + (int64_t)b * Sk * ld_dk + h * D, ld_dk, s.Q, Sk);
    store_tile<D>(dv + (int64_t)b * Sk * ld_dv + h * D, ld_dv, s.V, Sk);
    __syncthreads();   // this buffer (and dS / delta) are rewritten by the next iterations
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------- host
static MaskDev to_dev(const KitAttnMask* m) {
  MaskDev d;
  d.frame_mask = m ? m->frame_mask : nullptr;
  d.frame_mask_stride = m ? m->frame_mask_stride : 0;
  d.flags = (m && m->frame_mask) ? m->flags : (m ? (m->flags & KIT_MASK_TRIANGLE) : 0);
  d.bias = m ? m->bias : nullptr;
  d.bias_sb = m ? m->bias_stride_b : 0;
  d.bias_sh = m ? m->bias_stride_h : 0;
  return d;
}

template <int D>
static int fwd_launch(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, bf16* out,
                      int64_t ldo, float* lse, int B, int NH, int Sq, int Sk, const MaskDev& md, cudaStream_t st) {
  if (Sq <= AT && Sk <= AT && md.bias == nullptr) {   // one score tile per (batch, head): persistent double-buffered kernel
    static int ctas_per_sm = 0, sms = 0;
    constexpr int smem = 2 * (int)sizeof(FwdTile<D>);
    if (ctas_per_sm == 0) {
      KIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tile_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      KIT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, attn_fwd_tile_kernel<D>, AT_THREADS, smem));
      int dev = 0;
      KIT_CHECK_CUDA(cudaGetDevice(&dev));
      KIT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      KIT_REQUIRE(ctas_per_sm > 0, "attention forward tile kernel does not fit on an SM");
    }
    const int units = B * NH;
    const int grid1 = units < (sms - sm_reserve()) * ctas_per_sm ? units : (sms - sm_reserve()) * ctas_per_sm;
    launch_kernel(attn_fwd_tile_kernel<D>, dim3(grid1), dim3(AT_THREADS), smem, st, q, ldq, k, ldk, v, ldv, out, ldo, lse, NH, Sq, Sk,
                  units, rsqrtf((float)D), md);
    KIT_LAUNCH_CHECK();
    return KIT_OK;
  }
  if (md.bias == nullptr) {   // longer sequences without an explicit additive mask: persistent streaming kernel
    static int ctas_per_sm = 0, sms = 0;
    constexpr int smem = 2 * (int)sizeof(FwdTile<D>);
    if (ctas_per_sm == 0) {
      KIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_stream_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      KIT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, attn_fwd_stream_kernel<D>, AT_THREADS, smem));
      int dev = 0;
      KIT_CHECK_CUDA(cudaGetDevice(&dev));
      KIT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      KIT_REQUIRE(ctas_per_sm > 0, "attention forward streaming kernel does not fit on an SM");
    }
    const int q_tiles = (Sq + AT - 1) / AT, k_tiles = (Sk + AT - 1) / AT;
    const int64_t units64 = (int64_t)B * NH * q_tiles;
    KIT_REQUIRE(units64 < (1ll << 31), "attention forward: too many (batch, head, query tile) units");
    const int units = (int)units64;
    const int grid1 = units < (sms - sm_reserve()) * ctas_per_sm ? units : (sms - sm_reserve()) * ctas_per_sm;
    launch_kernel(attn_fwd_stream_kernel<D>, dim3(grid1), dim3(AT_THREADS), smem, st, q, ldq, k, ldk, v, ldv, out, ldo, lse, NH, Sq, Sk,
                  q_tiles, k_tiles, units, rsqrtf((float)D), md);
    KIT_LAUNCH_CHECK();
    return KIT_OK;
  }
  dim3 grid((Sq + AT - 1) / AT, B * NH);
  launch_kernel(attn_fwd_kernel<D>, dim3(grid), dim3(AT_THREADS), 0, st, q, ldq, k, ldk, v, ldv, out, ldo, lse, NH, Sq, Sk, rsqrtf((float)D), md);
  KIT_LAUNCH_CHECK();
  return KIT_OK;
}
template <int D>
static int bwd_launch(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, const bf16* o,
                      int64_t ldo, const bf16* dout, int64_t ld_do, const float* lse, bf16* dq, int64_t ld_dq, bf16* dk,
                      int64_t ld_dk, bf16* dv, int64_t ld_dv, float* dq_acc, int B, int NH, int Sq, int Sk,
                      const MaskDev& md, cudaStream_t st) {
  if (Sq <= AT && Sk <= AT && md.bias == nullptr) {
    static int ctas_per_sm = 0, sms = 0;
    constexpr int smem = (int)sizeof(BwdTileSmem<D>);
    if (ctas_per_sm == 0) {
      KIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tile_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      KIT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, attn_bwd_tile_kernel<D>, AT_THREADS, smem));
      int dev = 0;
      KIT_CHECK_CUDA(cudaGetDevice(&dev));
      KIT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      KIT_REQUIRE(ctas_per_sm > 0, "attention backward tile kernel does not fit on an SM");
    }
    const int units = B * NH;
    const int grid1 = units < (sms - sm_reserve()) * ctas_per_sm ? units : (sms - sm_reserve()) * ctas_per_sm;
    launch_kernel(attn_bwd_tile_kernel<D>, dim3(grid1), dim3(AT_THREADS), smem, st, q, ldq, k, ldk, v, ldv, o, ldo, dout, ld_do, lse, dq,
                  ld_dq, dk, ld_dk, dv, ld_dv, NH, Sq, Sk, units, rsqrtf((float)D), md);
    KIT_LAUNCH_CHECK();
    return KIT_OK;
  }
  static bool attr_done = false;
  if (!attr_done) {
    KIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem<D>)));
    attr_done = true;
  }
  const int ktiles = (Sk + AT - 1) / AT;
  if (ktiles > 1) {   // long sequences: the streaming kernel (delta once per call, query tiles prefetched with cp.async)
    KIT_REQUIRE(dq_acc != nullptr, "attention backward with more than 64 keys needs the fp32 dq accumulator workspace");
    KIT_CHECK_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)B * Sq * NH * D * sizeof(float), st));
    static bool stream_attr_done = false;
    if (!stream_attr_done) {
      KIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_stream_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdStreamSmem<D>)));
      stream_attr_done = true;
    }
    float* delta = dq_acc + (int64_t)B * Sq * NH * D;   // workspace tail: [B, NH, Sq]
    const int64_t n_rows = (int64_t)B * Sq;
    launch_kernel(attn_delta_kernel<D>, dim3((unsigned)ceil_div(n_rows * NH, 256)), dim3(256), 0, st, o, ldo, dout, ld_do, delta, NH, Sq, n_rows);
    KIT_LAUNCH_CHECK();
    launch_kernel(attn_bwd_stream_kernel<D>, dim3(ktiles, B * NH), dim3(AT_THREADS), sizeof(BwdStreamSmem<D>), st, q, ldq, k, ldk, v, ldv,
                  dout, ld_do, lse, (const float*)delta, dk, ld_dk, dv, ld_dv, dq_acc, NH, Sq, Sk, rsqrtf((float)D), md);
    KIT_LAUNCH_CHECK();
    const int64_t n = (int64_t)B * Sq * NH * D;
    launch_kernel(dq_convert_kernel, dim3((unsigned)ceil_div(n / 8, 256)), dim3(256), 0, st, dq_acc, dq, ld_dq, (int64_t)B * Sq, NH * D);
    KIT_LAUNCH_CHECK();
    return KIT_OK;
  }
  // one key tile with an explicit additive mask: the plain kernel (dQ written directly)
  dim3 grid(ktiles, B * NH);
  launch_kernel(attn_bwd_kernel<D>, dim3(grid), dim3(AT_THREADS), sizeof(BwdSmem<D>), st, q, ldq, k, ldk, v, ldv, o, ldo, dout, ld_do, lse, dq, ld_dq,
                                                                    dk, ld_dk, dv, ld_dv, dq_acc, NH, Sq, Sk, rsqrtf((float)D), md);
  KIT_LAUNCH_CHECK();
  return KIT_OK;
}

static int check_attn(int d, int64_t l0, int64_t l1, int64_t l2, int64_t l3) {
  KIT_REQUIRE(d == 16 || d == 32 || d == 64, "attention head dim %d unsupported (16, 32, 64)", d);
  KIT_REQUIRE(l0 % 8 == 0 && l1 % 8 == 0 && l2 % 8 == 0 && l3 % 8 == 0, "attention leading dims must be multiples of 8");
  return KIT_OK;
}

int attention_fwd(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, bf16* out,
                  int64_t ldo, float* lse, int B, int NH, int Sq, int Sk, int d, const KitAttnMask* mask,
                  cudaStream_t st) {
  int rc = check_attn(d, ldq, ldk, ldv, ldo);
  if (rc) return rc;
  {
    const void* ptrs[4] = {q, k, v, out};
    const int64_t lds[4] = {ldq, ldk, ldv, ldo};
    if (attention_t64_supported(NH, Sq, Sk, d, mask, ptrs, lds, 4))
      return attention_t64_fwd(q, ldq, k, ldk, v, ldv, out, ldo, lse, B, NH, Sq, Sk, mask, st);
  }
  if (attention_fwd_tc_supported(ldq, ldk, ldv, NH, Sq, Sk, d, mask, q, k, v))
    return attention_fwd_tc(q, ldq, k, ldk, v, ldv, out, ldo, lse, B, NH, Sq, Sk, d, mask, st);
  const MaskDev md = to_dev(mask);
  switch (d) {
    case 16: return fwd_launch<16>(q, ldq, k, ldk, v, ldv, out, ldo, lse, B, NH, Sq, Sk, md, st);
    case 32: return fwd_launch<32>(q, ldq, k, ldk, v, ldv, out, ldo, lse, B, NH, Sq, Sk, md, st);
    default: return fwd_launch<64>(q, ldq, k, ldk, v, ldv, out, ldo, lse, B, NH, Sq, Sk, md, st);
  }
}
int attention_bwd(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, const bf16* o,
                  int64_t ldo, const bf16* dout, int64_t ld_do, const float* lse, bf16* dq, int64_t ld_dq, bf16* dk,
                  int64_t ld_dk, bf16* dv, int64_t ld_dv, float* dq_acc, int B, int NH, int Sq, int Sk, int d,
                  const KitAttnMask* mask, cudaStream_t st) {
  int rc = check_attn(d, ldq, ldk, ldv, ldo);
  if (rc) return rc;
  rc = check_attn(d, ld_do, ld_dq, ld_dk, ld_dv);
  if (rc) return rc;
  {
    const void* ptrs[7] = {q, k, v, dout, dq, dk, dv};
    const int64_t lds[7] = {ldq, ldk, ldv, ld_do, ld_dq, ld_dk, ld_dv};
    if (attention_t64_supported(NH, Sq, Sk, d, mask, ptrs, lds, 7))
      return attention_t64_bwd(q, ldq, k, ldk, v, ldv, dout, ld_do, lse, dq, ld_dq, dk, ld_dk, dv, ld_dv, B, NH, Sq, Sk, mask, st);
  }
  {
    const void* ptrs[6] = {q, k, v, dout, dk, dv};
    const int64_t lds[6] = {ldq, ldk, ldv, ld_do, ld_dk, ld_dv};
    if (attention_bwd_tc_supported(NH, Sq, Sk, d, mask, ptrs, lds, 6, dq_acc)) {   // long sequences on tcgen05 (attention_tcb.cu)
      const int64_t n_rows = (int64_t)B * Sq, n = n_rows * NH * d;
      KIT_CHECK_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)n * sizeof(float), st));
      float* delta = dq_acc + n;   // workspace tail: [B, NH, Sq]
      if (d == 32)
        launch_kernel(attn_delta_kernel<32>, dim3((unsigned)ceil_div(n_rows * NH, 256)), dim3(256), 0, st, o, ldo, dout, ld_do, delta, NH, Sq, n_rows);
      else
        launch_kernel(attn_delta_kernel<64>, dim3((unsigned)ceil_div(n_rows * NH, 256)), dim3(256), 0, st, o, ldo, dout, ld_do, delta, NH, Sq, n_rows);
      KIT_LAUNCH_CHECK();
      rc = attention_bwd_tc(q, ldq, k, ldk, v, ldv, dout, ld_do, lse, delta, dq_acc, dk, ld_dk, dv, ld_dv, B, NH, Sq, Sk, d, mask, st);
      if (rc) return rc;
      launch_kernel(dq_convert_kernel, dim3((unsigned)ceil_div(n / 8, 256)), dim3(256), 0, st, (const float*)dq_acc, dq, ld_dq, n_rows, NH * d);
      KIT_LAUNCH_CHECK();
      return KIT_OK;
    }
  }
  const MaskDev md = to_dev(mask);
  switch (d) {
    case 16: return bwd_launch<16>(q, ldq, k, ldk, v, ldv, o, ldo, dout, ld_do, lse, dq, ld_dq, dk, ld_dk, dv, ld_dv, dq_acc, B, NH, Sq, Sk, md, st);
    case 32: return bwd_launch<32>(q, ldq, k, ldk, v, ldv, o, ldo, dout, ld_do, lse, dq, ld_dq, dk, ld_dk, dv, ld_dv, dq_acc, B, NH, Sq, Sk, md, st);
    default: return bwd_launch<64>(q, ldq, k, ldk, v, ldv, o, ldo, dout, ld_do, lse, dq, ld_dq, dk, ld_dk, dv, ld_dv, dq_acc, B, NH, Sq, Sk, md, st);
  }
}

}  // namespace kit

using namespace kit;
extern "C" int kit_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                 void* out, int64_t ldo, float* lse, int32_t B, int32_t NH, int32_t Sq, int32_t Sk,
                                 int32_t d, const KitAttnMask* mask, void* stream) {
  return attention_fwd((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (bf16*)out, ldo, lse, B, NH, Sq,
                       Sk, d, mask, (cudaStream_t)stream);
}
extern "C" int kit_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                 const void* out, int64_t ldo, const void* dout, int64_t ld_do, const float* lse, void* dq,
                                 int64_t ld_dq, void* dk, int64_t ld_dk, void* dv, int64_t ld_dv, float* dq_accum,
                                 int32_t B, int32_t NH, int32_t Sq, int32_t Sk, int32_t d, const KitAttnMask* mask,
                                 void* stream) {
  return attention_bwd((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (const bf16*)out, ldo,
                       (const bf16*)dout, ld_do, lse, (bf16*)dq, ld_dq, (bf16*)dk, ld_dk, (bf16*)dv, ld_dv, dq_accum, B,
                       NH, Sq, Sk, d, mask, (cudaStream_t)stream);
}

// softmax(Q K^T / sqrt(d) + mask) V, forward and backward, with the attention masks of the
// reference synthesised in-kernel from the [B,T] frame mask (model.py:193-202 "repeat-inc",
// A1_train.py:121 float key-padding mask that PyTorch ADDS to the logits) so that no [B*NH,S,S]
// tensor ever exists.  Flash-style tiling: 64 queries x 64 keys per step, online softmax, fp32
// math on bf16 operands; the backward recomputes P from the saved log-sum-exp.
//
// v1: CUDA-core tiles (the attention matmuls are 3 % of the step's FLOPs at T=64).
#include "attention.cuh"

namespace kit {

constexpr int AT = 64;         // tile edge (queries and keys)
constexpr int AT_THREADS = 128;
constexpr int ATP = AT + 4;    // padded row length of [*][64] smem tiles (keeps float4 alignment)

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

// rows [r0, r0+64) of a bf16 [*, ld] matrix (columns c0..c0+D) -> fp32 smem, transposed [D][ATP]
template <int D>
__device__ __forceinline__ void load_tile_T(float (*dst)[ATP], const bf16* base, int64_t ld, int r0, int nrows) {
  constexpr int VPR = D / 8;  // 16-byte vectors per row
  for (int idx = threadIdx.x; idx < AT * VPR; idx += AT_THREADS) {
    const int r = idx / VPR, v = idx % VPR;
    float f[8];
    if (r0 + r < nrows) {
      load8(base + (int64_t)(r0 + r) * ld + v * 8, f);
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) dst[v * 8 + u][r] = f[u];
  }
}
// same rows, row-major [64][D+4]
template <int D>
__device__ __forceinline__ void load_tile(float (*dst)[D + 4], const bf16* base, int64_t ld, int r0, int nrows) {
  constexpr int VPR = D / 8;
  for (int idx = threadIdx.x; idx < AT * VPR; idx += AT_THREADS) {
    const int r = idx / VPR, v = idx % VPR;
    float f[8];
    if (r0 + r < nrows) {
      load8(base + (int64_t)(r0 + r) * ld + v * 8, f);
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) dst[r][v * 8 + u] = f[u];
  }
}

// acc[4][8] = sum_c At[c][4*ty + i] * Bt[c][8*tx + j]
template <int D>
__device__ __forceinline__ void outer_tile(const float (*At)[ATP], const float (*Bt)[ATP], int ty, int tx,
                                           float acc[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 8
  for (int c = 0; c < D; ++c) {
    const float4 a = *reinterpret_cast<const float4*>(&At[c][4 * ty]);
    const float4 b0 = *reinterpret_cast<const float4*>(&Bt[c][8 * tx]);
    const float4 b1 = *reinterpret_cast<const float4*>(&Bt[c][8 * tx + 4]);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// out[4][CPT] += sum_r Pt[r][4*ty + i] * V[r][CPT*tx + j]      (reduction over the 64 rows r)
template <int D>
__device__ __forceinline__ void reduce_tile(const float (*Pt)[ATP], const float (*V)[D + 4], int ty, int tx,
                                            float acc[4][D / 8]) {
  constexpr int CPT = D / 8;
#pragma unroll 4
  for (int r = 0; r < AT; ++r) {
    const float4 p = *reinterpret_cast<const float4*>(&Pt[r][4 * ty]);
    const float pv[4] = {p.x, p.y, p.z, p.w};
    float vv[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) vv[j] = V[r][CPT * tx + j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(pv[i], vv[j], acc[i][j]);
  }
}

template <int D>
struct FwdSmem {
  float Qt[D][ATP];
  float Kt[D][ATP];
  float V[AT][D + 4];
  float Pt[AT][ATP];  // [key][query]
  float fm[AT];
};

template <int D>
__global__ void __launch_bounds__(AT_THREADS) attn_fwd_kernel(const bf16* __restrict__ q, int64_t ldq,
                                                              const bf16* __restrict__ k, int64_t ldk,
                                                              const bf16* __restrict__ v, int64_t ldv,
                                                              bf16* __restrict__ out, int64_t ldo,
                                                              float* __restrict__ lse, int NH, int Sq, int Sk,
                                                              float scale, MaskDev mask) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FwdSmem<D>& s = *reinterpret_cast<FwdSmem<D>*>(smem_raw);
  constexpr int CPT = D / 8;
  const int b = blockIdx.y / NH, h = blockIdx.y % NH;
  const int q0 = blockIdx.x * AT;
  const int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
  const bf16* qb = q + (int64_t)b * Sq * ldq + h * D;
  const bf16* kb = k + (int64_t)b * Sk * ldk + h * D;
  const bf16* vb = v + (int64_t)b * Sk * ldv + h * D;

  load_tile_T<D>(s.Qt, qb, ldq, q0, Sq);
  float m_run[4], l_run[4], o[4][CPT];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < CPT; ++j) o[i][j] = 0.f;
  }
  for (int k0 = 0; k0 < Sk; k0 += AT) {
    __syncthreads();  // previous tile fully consumed (also covers the Qt load on the first pass)
    load_tile_T<D>(s.Kt, kb, ldk, k0, Sk);
    load_tile<D>(s.V, vb, ldv, k0, Sk);
    if (threadIdx.x < AT) {
      const int j = k0 + threadIdx.x;
      s.fm[threadIdx.x] = (mask.frame_mask != nullptr && j < Sk) ? mask.frame_mask[(int64_t)b * mask.frame_mask_stride + j] : 0.f;
    }
    __syncthreads();
    float acc[4][8];
    outer_tile<D>(s.Qt, s.Kt, ty, tx, acc);
    float p_scale[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + 4 * ty + i;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kj = k0 + 8 * tx + j;
        float x = -INFINITY;
        if (kj < Sk && qi < Sq) x = acc[i][j] * scale + mask_bias(mask, b, h, qi, kj, Sk, s.fm[8 * tx + j]);
        acc[i][j] = x;
        mx = fmaxf(mx, x);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      const float m_new = fmaxf(m_run[i], mx);
      const float m_ref = (m_new == -INFINITY) ? 0.f : m_new;
      p_scale[i] = __expf(m_run[i] - m_ref);  // exp(-inf) = 0 on the first tile
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p = __expf(acc[i][j] - m_ref);
        acc[i][j] = p;
        rs += p;
      }
      rs += __shfl_xor_sync(0xffffffffu, rs, 1);
      rs += __shfl_xor_sync(0xffffffffu, rs, 2);
      rs += __shfl_xor_sync(0xffffffffu, rs, 4);
      l_run[i] = l_run[i] * p_scale[i] + rs;
      m_run[i] = m_new;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(&s.Pt[8 * tx + j][4 * ty]) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < CPT; ++j) o[i][j] *= p_scale[i];
    __syncthreads();
    reduce_tile<D>(s.Pt, s.V, ty, tx, o);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + 4 * ty + i;
    if (qi >= Sq) continue;
    const float inv = 1.f / l_run[i];
    bf16* op = out + ((int64_t)b * Sq + qi) * ldo + h * D + CPT * tx;
#pragma unroll
    for (int j = 0; j < CPT; ++j) op[j] = __float2bfloat16(o[i][j] * inv);
    if (tx == 0 && lse != nullptr) lse[((int64_t)b * NH + h) * Sq + qi] = m_run[i] + __logf(l_run[i]);
  }
}

// ---------------------------------------------------------------- backward
template <int D>
struct BwdSmem {
  float At[D][ATP];     // Q^T  (dKV kernel: per q tile)     | Q^T (dQ kernel: fixed)
  float Bt[D][ATP];     // K^T
  float Ct[D][ATP];     // dO^T
  float Dt[D][ATP];     // V^T
  float R0[AT][D + 4];  // dKV: Q row-major   | dQ: K row-major
  float R1[AT][D + 4];  // dKV: dO row-major
  float P[AT][ATP];     // dKV: P[query][key] | dQ: dS^T [key][query]
  float dS[AT][ATP];    // dKV: dS[query][key]
  float fm[AT];
  float lse[AT];
  float delta[AT];
};

// delta_i = sum_c dO[i,c] * O[i,c] for the 64 rows starting at r0 (one thread pair per row)
template <int D>
__device__ __forceinline__ void load_delta_lse(float* delta, float* lse_s, const bf16* ob, int64_t ldo, const bf16* dob,
                                               int64_t ld_do, const float* lse_g, int r0, int nrows) {
  for (int r = threadIdx.x; r < AT; r += AT_THREADS) {
    float d = 0.f, l = 0.f;
    if (r0 + r < nrows) {
      for (int c = 0; c < D; c += 8) {
        float a[8], g[8];
        load8(ob + (int64_t)(r0 + r) * ldo + c, a);
        load8(dob + (int64_t)(r0 + r) * ld_do + c, g);
#pragma unroll
        for (int u = 0; u < 8; ++u) d = fmaf(a[u], g[u], d);
      }
      l = lse_g[r0 + r];
    }
    delta[r] = d;
    lse_s[r] = l;
  }
}

// grid (key tiles, B*NH): dK, dV for one key tile, looping over query tiles
template <int D>
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_dkv_kernel(
    const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk, const bf16* __restrict__ v,
    int64_t ldv, const bf16* __restrict__ o, int64_t ldo, const bf16* __restrict__ dout, int64_t ld_do,
    const float* __restrict__ lse, bf16* __restrict__ dk, int64_t ld_dk, bf16* __restrict__ dv, int64_t ld_dv, int NH,
    int Sq, int Sk, float scale, MaskDev mask) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BwdSmem<D>& s = *reinterpret_cast<BwdSmem<D>*>(smem_raw);
  constexpr int CPT = D / 8;
  const int b = blockIdx.y / NH, h = blockIdx.y % NH;
  const int k0 = blockIdx.x * AT;
  const int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
  const bf16* qb = q + (int64_t)b * Sq * ldq + h * D;
  const bf16* kb = k + (int64_t)b * Sk * ldk + h * D;
  const bf16* vb = v + (int64_t)b * Sk * ldv + h * D;
  const bf16* ob = o + (int64_t)b * Sq * ldo + h * D;
  const bf16* dob = dout + (int64_t)b * Sq * ld_do + h * D;
  const float* lse_g = lse + ((int64_t)b * NH + h) * Sq;

  load_tile_T<D>(s.Bt, kb, ldk, k0, Sk);
  load_tile_T<D>(s.Dt, vb, ldv, k0, Sk);
  if (threadIdx.x < AT) {
    const int j = k0 + threadIdx.x;
    s.fm[threadIdx.x] = (mask.frame_mask != nullptr && j < Sk) ? mask.frame_mask[(int64_t)b * mask.frame_mask_stride + j] : 0.f;
  }
  float dk_acc[4][CPT], dv_acc[4][CPT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CPT; ++j) dk_acc[i][j] = dv_acc[i][j] = 0.f;

  for (int q0 = 0; q0 < Sq; q0 += AT) {
    __syncthreads();
    load_tile_T<D>(s.At, qb, ldq, q0, Sq);
    load_tile_T<D>(s.Ct, dob, ld_do, q0, Sq);
    load_tile<D>(s.R0, qb, ldq, q0, Sq);
    load_tile<D>(s.R1, dob, ld_do, q0, Sq);
    load_delta_lse<D>(s.delta, s.lse, ob, ldo, dob, ld_do, lse_g, q0, Sq);
    __syncthreads();
    float sc[4][8], dp[4][8];
    outer_tile<D>(s.At, s.Bt, ty, tx, sc);  // Q K^T
    outer_tile<D>(s.Ct, s.Dt, ty, tx, dp);  // dO V^T
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + 4 * ty + i;
      const float l = s.lse[4 * ty + i], dl = s.delta[4 * ty + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kj = k0 + 8 * tx + j;
        float p = 0.f;
        if (kj < Sk && qi < Sq) {
          const float x = sc[i][j] * scale + mask_bias(mask, b, h, qi, kj, Sk, s.fm[8 * tx + j]);
          p = __expf(x - l);
        }
        s.P[4 * ty + i][8 * tx + j] = p;
        s.dS[4 * ty + i][8 * tx + j] = p * (dp[i][j] - dl) * scale;
      }
    }
    __syncthreads();
    // dV[key][c] += sum_i P[i][key] dO[i][c];  dK[key][c] += sum_i dS[i][key] Q[i][c]
    reduce_tile<D>(s.P, s.R1, ty, tx, dv_acc);
    reduce_tile<D>(s.dS, s.R0, ty, tx, dk_acc);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kj = k0 + 4 * ty + i;
    if (kj >= Sk) continue;
    bf16* dkp = dk + ((int64_t)b * Sk + kj) * ld_dk + h * D + CPT * tx;
    bf16* dvp = dv + ((int64_t)b * Sk + kj) * ld_dv + h * D + CPT * tx;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      dkp[j] = __float2bfloat16(dk_acc[i][j]);
      dvp[j] = __float2bfloat16(dv_acc[i][j]);
    }
  }
}

// grid (query tiles, B*NH): dQ for one query tile, looping over key tiles
template <int D>
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_dq_kernel(
    const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk, const bf16* __restrict__ v,
    int64_t ldv, const bf16* __restrict__ o, int64_t ldo, const bf16* __restrict__ dout, int64_t ld_do,
    const float* __restrict__ lse, bf16* __restrict__ dq, int64_t ld_dq, int NH, int Sq, int Sk, float scale,
    MaskDev mask) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BwdSmem<D>& s = *reinterpret_cast<BwdSmem<D>*>(smem_raw);
  constexpr int CPT = D / 8;
  const int b = blockIdx.y / NH, h = blockIdx.y % NH;
  const int q0 = blockIdx.x * AT;
  const int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
  const bf16* qb = q + (int64_t)b * Sq * ldq + h * D;
  const bf16* kb = k + (int64_t)b * Sk * ldk + h * D;
  const bf16* vb = v + (int64_t)b * Sk * ldv + h * D;
  const bf16* ob = o + (int64_t)b * Sq * ldo + h * D;
  const bf16* dob = dout + (int64_t)b * Sq * ld_do + h * D;
  const float* lse_g = lse + ((int64_t)b * NH + h) * Sq;

  load_tile_T<D>(s.At, qb, ldq, q0, Sq);
  load_tile_T<D>(s.Ct, dob, ld_do, q0, Sq);
  load_delta_lse<D>(s.delta, s.lse, ob, ldo, dob, ld_do, lse_g, q0, Sq);
  float dq_acc[4][CPT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CPT; ++j) dq_acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Sk; k0 += AT) {
    __syncthreads();
    load_tile_T<D>(s.Bt, kb, ldk, k0, Sk);
    load_tile_T<D>(s.Dt, vb, ldv, k0, Sk);
    load_tile<D>(s.R0, kb, ldk, k0, Sk);
    if (threadIdx.x < AT) {
      const int j = k0 + threadIdx.x;
      s.fm[threadIdx.x] = (mask.frame_mask != nullptr && j < Sk) ? mask.frame_mask[(int64_t)b * mask.frame_mask_stride + j] : 0.f;
    }
    __syncthreads();
    float sc[4][8], dp[4][8];
    outer_tile<D>(s.At, s.Bt, ty, tx, sc);
    outer_tile<D>(s.Ct, s.Dt, ty, tx, dp);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float ds[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int qi = q0 + 4 * ty + i, kj = k0 + 8 * tx + j;
        float p = 0.f;
        if (kj < Sk && qi < Sq) {
          const float x = sc[i][j] * scale + mask_bias(mask, b, h, qi, kj, Sk, s.fm[8 * tx + j]);
          p = __expf(x - s.lse[4 * ty + i]);
        }
        ds[i] = p * (dp[i][j] - s.delta[4 * ty + i]) * scale;
      }
      *reinterpret_cast<float4*>(&s.P[8 * tx + j][4 * ty]) = make_float4(ds[0], ds[1], ds[2], ds[3]);  // dS^T
    }
    __syncthreads();
    reduce_tile<D>(s.P, s.R0, ty, tx, dq_acc);  // dQ[i][c] += sum_j dS[i][j] K[j][c]
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + 4 * ty + i;
    if (qi >= Sq) continue;
    bf16* dqp = dq + ((int64_t)b * Sq + qi) * ld_dq + h * D + CPT * tx;
#pragma unroll
    for (int j = 0; j < CPT; ++j) dqp[j] = __float2bfloat16(dq_acc[i][j]);
  }
}

// ---------------------------------------------------------------- host
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
  static bool attr_done = false;
  if (!attr_done) {
    KIT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(FwdSmem<D>)));
    attr_done = true;
  }
  dim3 grid((Sq + AT - 1) / AT, B * NH);
  attn_fwd_kernel<D><<<grid, AT_THREADS, sizeof(FwdSmem<D>), st>>>(q, ldq, k, ldk, v, ldv, out, ldo, lse, NH, Sq, Sk,
                                                                   rsqrtf((float)D), md);
  KIT_LAUNCH_CHECK();
  return KIT_OK;
}
template <int D>
static int bwd_launch(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, const bf16* o,
                      int64_t ldo, const bf16* dout, int64_t ld_do, const float* lse, bf16* dq, int64_t ld_dq, bf16* dk,
                      int64_t ld_dk, bf16* dv, int64_t ld_dv, int B, int NH, int Sq, int Sk, const MaskDev& md,
                      cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    KIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(BwdSmem<D>)));
    KIT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(BwdSmem<D>)));
    attr_done = true;
  }
  const float scale = rsqrtf((float)D);
  dim3 gkv((Sk + AT - 1) / AT, B * NH);
  attn_bwd_dkv_kernel<D><<<gkv, AT_THREADS, sizeof(BwdSmem<D>), st>>>(q, ldq, k, ldk, v, ldv, o, ldo, dout, ld_do, lse,
                                                                       dk, ld_dk, dv, ld_dv, NH, Sq, Sk, scale, md);
  KIT_LAUNCH_CHECK();
  dim3 gq((Sq + AT - 1) / AT, B * NH);
  attn_bwd_dq_kernel<D><<<gq, AT_THREADS, sizeof(BwdSmem<D>), st>>>(q, ldq, k, ldk, v, ldv, o, ldo, dout, ld_do, lse, dq,
                                                                     ld_dq, NH, Sq, Sk, scale, md);
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
  const MaskDev md = to_dev(mask);
  switch (d) {
    case 16: return fwd_launch<16>(q, ldq, k, ldk, v, ldv, out, ldo, lse, B, NH, Sq, Sk, md, st);
    case 32: return fwd_launch<32>(q, ldq, k, ldk, v, ldv, out, ldo, lse, B, NH, Sq, Sk, md, st);
    default: return fwd_launch<64>(q, ldq, k, ldk, v, ldv, out, ldo, lse, B, NH, Sq, Sk, md, st);
  }
}
int attention_bwd(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, const bf16* o,
                  int64_t ldo, const bf16* dout, int64_t ld_do, const float* lse, bf16* dq, int64_t ld_dq, bf16* dk,
                  int64_t ld_dk, bf16* dv, int64_t ld_dv, int B, int NH, int Sq, int Sk, int d, const KitAttnMask* mask,
                  cudaStream_t st) {
  int rc = check_attn(d, ldq, ldk, ldv, ldo);
  if (rc) return rc;
  rc = check_attn(d, ld_do, ld_dq, ld_dk, ld_dv);
  if (rc) return rc;
  const MaskDev md = to_dev(mask);
  switch (d) {
    case 16: return bwd_launch<16>(q, ldq, k, ldk, v, ldv, o, ldo, dout, ld_do, lse, dq, ld_dq, dk, ld_dk, dv, ld_dv, B, NH, Sq, Sk, md, st);
    case 32: return bwd_launch<32>(q, ldq, k, ldk, v, ldv, o, ldo, dout, ld_do, lse, dq, ld_dq, dk, ld_dk, dv, ld_dv, B, NH, Sq, Sk, md, st);
    default: return bwd_launch<64>(q, ldq, k, ldk, v, ldv, o, ldo, dout, ld_do, lse, dq, ld_dq, dk, ld_dk, dv, ld_dv, B, NH, Sq, Sk, md, st);
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
                                 int64_t ld_dq, void* dk, int64_t ld_dk, void* dv, int64_t ld_dv, int32_t B, int32_t NH,
                                 int32_t Sq, int32_t Sk, int32_t d, const KitAttnMask* mask, void* stream) {
  return attention_bwd((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, (const bf16*)out, ldo,
                       (const bf16*)dout, ld_do, lse, (bf16*)dq, ld_dq, (bf16*)dk, ld_dk, (bf16*)dv, ld_dv, B, NH, Sq, Sk,
                       d, mask, (cudaStream_t)stream);
}

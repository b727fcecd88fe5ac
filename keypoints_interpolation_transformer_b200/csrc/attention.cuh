#pragma once
#include "common.cuh"

namespace kit {
int attention_fwd(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, bf16* out,
                  int64_t ldo, float* lse, int B, int NH, int Sq, int Sk, int d, const KitAttnMask* mask,
                  cudaStream_t st);
int attention_bwd(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, const bf16* o,
                  int64_t ldo, const bf16* dout, int64_t ld_do, const float* lse, bf16* dq, int64_t ld_dq, bf16* dk,
                  int64_t ld_dk, bf16* dv, int64_t ld_dv, float* dq_acc, int B, int NH, int Sq, int Sk, int d,
                  const KitAttnMask* mask, cudaStream_t st);
// attention_tc.cu: forward on tcgen05 tensor cores for more than one key tile (no explicit bias tensor, d = 32 / 64)
bool attention_fwd_tc_supported(int64_t ldq, int64_t ldk, int64_t ldv, int NH, int Sq, int Sk, int d, const KitAttnMask* mask,
                                const void* q, const void* k, const void* v);
int attention_fwd_tc(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, bf16* out, int64_t ldo,
                     float* lse, int B, int NH, int Sq, int Sk, int d, const KitAttnMask* mask, cudaStream_t st);
// attention_tcb.cu: backward on tcgen05 for more than one key tile (no explicit bias tensor, d = 32 / 64); delta = rowsum(dO * O)
// [B, NH, Sq] and the zeroed fp32 dQ accumulator [B * Sq, NH * d] come from the caller (attention.cu)
bool attention_bwd_tc_supported(int NH, int Sq, int Sk, int d, const KitAttnMask* mask, const void* const* ptrs, const int64_t* lds, int n,
                                const float* dq_acc);
int attention_bwd_tc(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, const bf16* dout, int64_t ld_do,
                     const float* lse, const float* delta, float* dq_acc, bf16* dk, int64_t ld_dk, bf16* dv, int64_t ld_dv, int B, int NH,
                     int Sq, int Sk, int d, const KitAttnMask* mask, cudaStream_t st);
// attention_t64.cu: forward and backward on tcgen05 for Sq = Sk <= 64, d = 32, an even number of heads, no explicit bias tensor
bool attention_t64_supported(int NH, int Sq, int Sk, int d, const KitAttnMask* mask, const void* const* ptrs, const int64_t* lds, int n);
int attention_t64_fwd(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, bf16* out, int64_t ldo,
                      float* lse, int B, int NH, int Sq, int Sk, const KitAttnMask* mask, cudaStream_t st);
int attention_t64_bwd(const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv, const bf16* dout,
                      int64_t ld_do, const float* lse, bf16* dq, int64_t ld_dq, bf16* dk, int64_t ld_dk, bf16* dv, int64_t ld_dv,
                      int B, int NH, int Sq, int Sk, const KitAttnMask* mask, cudaStream_t st);
}  // namespace kit

// Host launchers of the bandwidth-bound row kernels (rowops.cu).
#pragma once
#include "common.cuh"

namespace kit {

// One 2-D weight of the fp32 arena and where its bf16 copy (dst) and transposed bf16 copy (dstT)
// live in the bf16 weight arena (-1 = not needed).  cols_pad/rows_pad: zero-filled extents.
constexpr int WR_TILE = 64;   // weight_refresh_kernel: tile edge (the tile table built at bind time uses it too)
struct WeightDesc {
  int64_t src_off;
  int rows, cols;
  int64_t dst_off;
  int dst_ld, rows_pad, cols_pad;
  int64_t dstT_off;
  int dstT_ld;
};

// norm_scale: weight of the token-norm term in the position sum (1; 2 in KeypointCompleterCycle, model.py:283-284)
int embed_post_fwd(const bf16* raw, const float* pe, const float* learned, bf16* out, int64_t M, int H, int T,
                   float norm_scale, cudaStream_t st);
int embed_post_bwd(const bf16* dout, const bf16* raw, const bf16* addend, bf16* draw, float* dlearned, int64_t M, int H,
                   float norm_scale, cudaStream_t st);
int add_ln_fwd(const bf16* a, const bf16* b, const float* gamma, const float* beta, bf16* sum_out, bf16* y, float* mean,
               float* rstd, int64_t M, int H, cudaStream_t st);
int ln_bwd(const bf16* dy, const bf16* s_saved, const float* mean, const float* rstd, const float* gamma,
           const bf16* addend, bf16* dx, float* dgamma, float* dbeta, float* dxsum, int64_t M, int H, cudaStream_t st);
int swiglu_gate_fwd(const bf16* x12, bf16* g, int64_t M, int H, cudaStream_t st);
int swiglu_gate_bwd(const bf16* dg, const bf16* x12, bf16* dx12, int64_t M, int H, cudaStream_t st);
int final_norm_silu_fwd(const bf16* dec, const bf16* femb, bf16* z_out, bf16* out, int64_t M, int H, cudaStream_t st);
int final_norm_silu_bwd(const bf16* dout, const bf16* z, bf16* dz, int64_t M, int H, cudaStream_t st);
int colsum(const bf16* x, int64_t ld, float* out, int64_t M, int N, cudaStream_t st);
int cast_pad(const float* src, int64_t rows, int64_t cols, int64_t src_ld, bf16* dst, int64_t dst_ld, cudaStream_t st);
int pack_frames(const float* src, int64_t batch_stride, int B, int T, int cols, const float* zero_mask,
                int64_t zero_mask_stride, bf16* dst, int dst_ld, cudaStream_t st);
int weight_refresh(const float* params, bf16* wb, const WeightDesc* descs_dev, const int* tile_prefix_dev, int n_desc,
                   int total_tiles, cudaStream_t st);

}  // namespace kit

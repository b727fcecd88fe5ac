/* kit.h -- C ABI of libkit_b200.so: the B200 (sm_100a) implementation of the train / infer step
 * of JoeNatan30/keypoints_interpolation_transformer's KeypointCompleter.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; the "interface each entry point
 * replaces" is therefore the Python call site it stands in for, cited as file:line relative to the
 * reference checkout.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller (PyTorch) owns all
 *     memory; the library never allocates persistent device memory and never frees caller memory.
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), returns KIT_OK or
 *     a negative code and leaves a message for kit_last_error() (thread local).  Unsupported
 *     shapes are errors: there is no CPU fallback.
 *   - tensors are dense row-major fp32 unless stated; "bf16" = __nv_bfloat16.
 *
 * Numerical deviations from the reference (each bounded by a test)
 *   - GEMM / attention operands are bf16 with fp32 accumulation (north_star: 2e-2 relative on outputs, loss, gradients).
 *   - GELU: nn.Transformer(activation="gelu") (model.py:87) is the erf form; the kernels evaluate
 *     0.5 x (1 + tanh(x (c0 + c1 x^2 + c2 x^4))) with coefficients fitted to the erf form: |gelu - gelu_erf| <= 2.6e-5
 *     absolute, |gelu' - gelu_erf'| <= 1.1e-4 (csrc/common.cuh; tests/test_kernels_gpu.py::test_gelu_matches_erf_form).
 */
#ifndef KIT_B200_H_
#define KIT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KIT_OK 0
#define KIT_ERR_INVALID (-1)
#define KIT_ERR_CUDA (-2)
#define KIT_ERR_UNSUPPORTED (-3)

#define KIT_ABI_VERSION 2

/* attention-mask synthesis flags (model.py:172-209 + torch/nn/functional.py:6620) */
#define KIT_MASK_NONE 0
#define KIT_MASK_REPEAT_INC 1 /* -inf iff (j > i && frame_mask[j] == 1)   model.py:193-202         */
#define KIT_MASK_KEYPAD_ADD 2 /* + frame_mask[j] (float key-padding mask is ADDED, A1_train.py:121) */
#define KIT_MASK_TRIANGLE 4   /* -inf iff j > i                           model.py:174-180         */

/* loss kinds */
#define KIT_LOSS_EUCLID 0 /* euclidean_loss.py:8-17  mean_points sum_xy (o-t)^2                    */
#define KIT_LOSS_MSE 1    /* A1_train.py:254 torch.nn.MSELoss (= EUCLID / 2)                       */
#define KIT_LOSS_DISTANCE 2 /* euclidean_loss.py:19-37 EuclideanDistanceLoss: sum_points ||o-t||_2 (A4 validation) */

/* get_mask matrix types (model.py:172) */
#define KIT_MATRIX_TRIANGLE 0
#define KIT_MATRIX_REPEAT 1
#define KIT_MATRIX_REPEAT_INC 2
#define KIT_MATRIX_ALL 3

/* augmentation kinds (dataloader.py:649-663 dispatch) */
#define KIT_AUG_NONE 0
#define KIT_AUG_ROTATE 1      /* augmentation.py:121-142 */
#define KIT_AUG_SHEAR 2       /* augmentation.py:144-203 (squeeze and perspective: a 3x3 matrix) */
#define KIT_AUG_ARM_ROTATE 3  /* augmentation.py:206-233 */

const char* kit_last_error(void);
int kit_version(void);
/* Leave n SMs (even, <= 64) free in every persistent kernel's grid: data-parallel training runs NCCL's all-reduce kernels beside
 * backward (parallel.BucketReducer sets it together with NCCL_MAX_CTAS).  Call before the first engine is created. */
int kit_set_sm_reserve(int32_t n);

/* ------------------------------------------------------------------------------------------
 * Model description and parameter arena.  model.py:61-98 (constructor arguments) -- ff and max_len
 * are the implicit nn.Transformer / PositionalEncoding constants (2048, 2048).
 * ---------------------------------------------------------------------------------------- */
typedef struct KitModelConfig {
  int32_t input_size; /* 2*K */
  int32_t hidden;     /* H, multiple of 64, <= 1024 */
  int32_t layers;     /* encoder layers == decoder layers */
  int32_t heads;      /* H / heads in {16, 32, 64, 128} */
  int32_t ff;         /* dim_feedforward (2048) */
  int32_t max_len;    /* rows of the trig positional table (2048; 512 in KeypointCompleterCycle, model.py:226-227) */
  int32_t variant;    /* KIT_MODEL_* */
  int32_t reserved;
} KitModelConfig;
#define KIT_MODEL_COMPLETER 0 /* model.py:60-170  KeypointCompleter */
#define KIT_MODEL_CYCLE 1     /* model.py:212-321 KeypointCompleterCycle: the token-norm output enters the position sum twice
                               * (:279-284: PositionalEncoding already returns norm + pe), and tgt_pad_mask reaches
                               * nn.Transformer (:294) -- pass it as dec_mask.frame_mask with KIT_MASK_KEYPAD_ADD */

/* All parameters live in ONE fp32 arena (and their gradients in a second arena of the same layout)
 * so that the optimiser is one kernel and the data-parallel all-reduce is a few large contiguous
 * buckets.  Entries are the reference's state_dict tensors; entry `index` is found at
 * [offset, offset+numel).  Buffers (the two trig tables) come after all trainable entries. */
int32_t kit_layout_num_entries(const KitModelConfig* cfg);
int kit_layout_entry(const KitModelConfig* cfg, int32_t index, char* name, int32_t name_cap, int64_t* offset,
                     int64_t* numel, int64_t* rows, int64_t* cols, int32_t* is_buffer);
int64_t kit_layout_trainable_floats(const KitModelConfig* cfg); /* Adam / all-reduce extent */
int64_t kit_layout_total_floats(const KitModelConfig* cfg);     /* arena size incl. buffers */
/* Backward finishes gradient ranges in this order (bucket 0 first): used to overlap the NCCL
 * all-reduce with backward.  Returns the number of buckets; ranges are [begin,end) in floats. */
int32_t kit_layout_num_buckets(const KitModelConfig* cfg);
int kit_layout_bucket(const KitModelConfig* cfg, int32_t bucket, int64_t* begin, int64_t* end);

/* ------------------------------------------------------------------------------------------
 * Engine: the whole KeypointCompleter forward (model.py:100-170) and its backward for a fixed
 * (batch, seq_len).  Replaces `model(x, x_no_sota, src_pad_mask=..., src_mask=..., tgt_mask=...)`
 * at A1_train.py:120-124 / :175-179 and `loss.backward()` at A1_train.py:134.
 * ---------------------------------------------------------------------------------------- */
typedef struct KitEngine KitEngine;

typedef struct KitAttnMask {
  const float* frame_mask;   /* [B, T] 0/1 floats, row stride frame_mask_stride; NULL = none */
  int64_t frame_mask_stride; /* elements between consecutive batch rows */
  int32_t flags;             /* KIT_MASK_* applied to frame_mask */
  int32_t reserved;
  const float* bias;         /* optional explicit additive mask, rows of length T; NULL = none */
  int64_t bias_stride_b;     /* elements between batches (0 = shared)  */
  int64_t bias_stride_h;     /* elements between heads   (0 = shared)  */
} KitAttnMask;

/* training != 0 keeps every layer's activations for kit_engine_backward; training == 0 lets all layers
 * share one set of buffers (inference: 3_test_IA_interpolation / A1_train.py:139-218 eval). */
int kit_engine_create(const KitModelConfig* cfg, int32_t batch, int32_t seq_len, int32_t training, KitEngine** out);
int kit_engine_destroy(KitEngine* e);
int64_t kit_engine_workspace_bytes(const KitEngine* e);
/* params/grads: fp32 arenas of kit_layout_total_floats / kit_layout_trainable_floats floats. */
int kit_engine_bind(KitEngine* e, float* params, float* grads, void* workspace, int64_t workspace_bytes);
/* Re-derive the bf16 (and transposed bf16) GEMM operands from the fp32 arena.  Call after every
 * optimiser step / load_state_dict (A1_train.py:135, A4_train_with_pretrained.py:227). */
int kit_engine_refresh_weights(KitEngine* e, void* stream);
/* x_enc / x_dec: frame f of batch b at ptr + b*batch_stride + f*input_size (so the A1 slices
 * inputs[:-1] and inputs[1:] of one [B,T+1,K,2] tensor are two pointers, A1_train.py:93-94).
 * zero_masked_enc != 0 applies A4_train_with_pretrained.py:107-108 using enc_mask.frame_mask.
 * pred: [B,T,input_size] fp32.  save_for_backward = 0 skips nothing but allows buffer reuse. */
int kit_engine_forward(KitEngine* e, const float* x_enc, int64_t x_enc_batch_stride, const float* x_dec,
                       int64_t x_dec_batch_stride, const KitAttnMask* enc_mask, const KitAttnMask* dec_mask,
                       int32_t zero_masked_enc, float* pred, void* stream);
/* dpred: [B,T,input_size] fp32 = dLoss/dpred.  Gradients are ACCUMULATED (+=) into the bound grad
 * arena; `bucket_done`, if non-NULL, is called on the host right after the kernels that finish
 * bucket i have been enqueued (the caller records an event and launches its all-reduce). */
typedef void (*KitBucketCallback)(int32_t bucket, void* user);
int kit_engine_backward(KitEngine* e, const float* dpred, KitBucketCallback bucket_done, void* user,
                        void* stream);
/* The engine's own bf16 operand buffers [B*T, k2p] (k2p = input_size rounded up to 8): kit_prepass can write its x_enc_bf16 /
 * x_dec_bf16 outputs straight into them, after which kit_engine_forward is called with x_enc = x_dec = NULL ("operands are in
 * place": no packing pass; zero_masked_enc must then have been applied by the pre-pass). */
int kit_engine_operands(KitEngine* e, void** x_enc_bf16, void** x_dec_bf16, int32_t* k2p);
/* Debug/inspection: copy out a named internal activation as fp32 (tests only). */
int kit_engine_debug_read(KitEngine* e, const char* name, float* out, int64_t out_floats, void* stream);
/* Number of kernels the last forward / backward call launched (bench.py's gpu_launches). */
int64_t kit_engine_last_launches(const KitEngine* e);
/* Per-category CUDA-event timing of the engine's own launches (bench.py's roofline leg).  Counters are
 * reset by kit_engine_forward and accumulate over the following backward; profile_read synchronises
 * on the recorded events.  flops = algorithmic FLOPs of that category's launches (2MNK per GEMM). */
#define KIT_PROF_GEMM_TN 0
#define KIT_PROF_GEMM_WGRAD 1
#define KIT_PROF_ATTN_FWD 2
#define KIT_PROF_ATTN_BWD 3
#define KIT_PROF_FFN 4 /* the fused feed-forward kernels (forward and input-gradient pass) */
#define KIT_PROF_CATEGORIES 5
int kit_engine_set_profiling(KitEngine* e, int32_t on);
int kit_engine_profile_read(KitEngine* e, int32_t category, float* ms, int64_t* launches, double* flops);

/* ------------------------------------------------------------------------------------------
 * Fused per-frame passes (HBM-bound)
 * ---------------------------------------------------------------------------------------- */
/* Per-sequence preprocessing parameters, host-drawn exactly as the reference draws them. */
typedef struct KitSeqAug {
  int32_t kind;       /* KIT_AUG_* */
  int32_t reserved;
  float cos_t, sin_t; /* ROTATE: cos/sin of the angle (augmentation.py:76-77, python doubles) */
  double mtx[9];      /* SHEAR: cv2.getPerspectiveTransform 3x3 (float64, augmentation.py:173,187) */
  float zero_x, zero_y; /* SHEAR: float32 image of (0,0) (augmentation.py:198) */
  float arm_cos[8], arm_sin[8]; /* ARM: per (chain c, joint j) at [c*4+j]; cos=2 marks "coin failed" */
} KitSeqAug;

typedef struct KitPrepassConfig {
  int32_t B, T, K;
  int32_t normalize;                               /* dataloader.py:71-140 on/off */
  int32_t left_shoulder, right_shoulder, right_eye; /* body_dict indices (dataloader.py:81,120) */
  int32_t n_body, n_hand;                          /* lengths of body_ids / hand_ids */
  int32_t arm_chain[8];                            /* 2 chains x [chest, shoulder, elbow, wrist] */
  int32_t zero_masked_enc;                         /* A4_train_with_pretrained.py:107-108 */
  int32_t k2p;                                     /* bf16 row pitch (>= 2K, multiple of 8) or 0 */
} KitPrepassConfig;

/* raw [B,T,K,2] -> y (normalised + augmented ground truth `sota`, dataloader.py:632-663),
 * inputs [B,T+1,K,2] (hold-filled + SOS, dataloader.py:421-434,482-493) and mask [B,T+1];
 * optionally the two bf16 GEMM operands x_enc/x_dec [B*T, k2p] (A1_train.py:93-94 slices).
 * src_index [B,T] int32: frame t of the hold-filled video is frame src_index[t] of y (-1 = zeros);
 * frame_missing [B,T] float 0/1.  body_ids/hand_ids: int32 device arrays (may be NULL when no
 * augmentation uses them).  aug: [B] KitSeqAug on the device (may be NULL = no augmentation). */
int kit_prepass(const KitPrepassConfig* cfg, const float* raw, const int32_t* src_index,
                const float* frame_missing, const KitSeqAug* aug, const int32_t* body_ids,
                const int32_t* hand_ids, float* y, float* inputs, float* mask, void* x_enc_bf16,
                void* x_dec_bf16, void* stream);

/* Loss + gradient in one pass (euclidean_loss.py:8-17, A1_train.py:128,184-186).
 * pred,target [n_frames, K, 2]; frame_weight [n_frames] or NULL: when given, the eval blend
 * pred*m + y*(1-m) is applied first (only frames with m=1 contribute; denominator unchanged).
 * loss_out: one float, OVERWRITTEN (partials are reduced in a fixed order: deterministic).
 * dpred (optional) = d loss / d pred * grad_scale.  partials: workspace of kit_loss_partials()
 * floats. */
int64_t kit_loss_partials(int64_t n_frames, int32_t K);
int kit_loss_fwd_bwd(const float* pred, const float* target, const float* frame_weight, int64_t n_frames,
                     int32_t K, int32_t loss_kind, float grad_scale, float* loss_out, float* dpred,
                     float* partials, void* stream);

/* put_missing_frames' block policy (dataloader.py:364-434, the non-random mode every trainer uses) drawn ON THE DEVICE for a
 * whole batch: per sequence the quartiles of `samples` normals of each statistic, the block count, lengths and offsets, and
 * the hold-fill index map -- Philox stream (seed, offset), i.e. the reference's distribution, not its draws.  Outputs feed
 * kit_prepass directly.  blocks [B, 64, 2] / n_blocks [B] (optional, both or neither): the (start, end) pairs drawn. */
typedef struct KitMissingStats {
  float mean_consecutive_missing, std_consecutive_missing;   /* dataset_config.json */
  float mean_number_missing_blocks, std_number_missing_blocks;
  int32_t samples;
} KitMissingStats;
int kit_draw_missing(const KitMissingStats* stats, int32_t B, int32_t T, uint64_t seed, uint64_t offset,
                     int32_t* src_index, float* frame_missing, int32_t* blocks, int32_t* n_blocks, void* stream);

/* The whole per-batch random policy of LSP_Dataset.__getitem__ (dataloader.py:649-675) on the device: which augmentation with
 * which parameters (p = prob, uniform choice of rotate / perspective / squeeze / arm-joint rotate; augmentation.py:132,166-185,
 * 221-224) and the missing blocks of kit_draw_missing, on Philox streams whose offset is a DEVICE counter (8 bytes, advanced by
 * the call) -- so a CUDA graph holding the step draws fresh values at every replay.  aug_out [B] KitSeqAug (with aug_policy) feeds
 * kit_prepass; aug_draws [B,12] doubles (optional, tests): {selected kind or -1, scalar draws 0 / 1, coin, 8 arm angles (NaN =
 * coin failed)}. */
typedef struct KitAugPolicy {
  float prob;      /* augmentations_prob (dataloader.py:649) */
  float angle_deg; /* 15: rotate and arm-joint angles are U(-angle, angle) degrees (dataloader.py:654,663) */
  float squeeze;   /* 0.15: shear ratios U(-s, s) (dataloader.py:657,660) */
  float arm_prob;  /* 0.5 (dataloader.py:663) */
  int32_t has_arms; /* 0: the skeleton has no arm chains -- selection 3 leaves the sequence unchanged */
} KitAugPolicy;
int kit_draw_policy(const KitMissingStats* stats, const KitAugPolicy* aug_policy, int32_t B, int32_t T, uint64_t seed,
                    uint64_t* counter_dev, int32_t* src_index, float* frame_missing, KitSeqAug* aug_out, double* aug_draws,
                    void* stream);

/* The cubic-spline baseline of the evaluation (3_test_cubic_interpolation.py:32-58): frames with mask == 1 and every exact 0
 * are missing; each (keypoint, coordinate) series is filled by the not-a-knot cubic spline through its remaining samples
 * (pandas interpolate(method="cubicspline", limit_direction="both") = scipy CubicSpline, extrapolating at both ends),
 * all-missing series become 0 (np.nan_to_num).  data / out [B, T1, K, 2] fp32, mask [B, T1] fp32; T1 <= 1040. */
int kit_cubic_interpolate(const float* data, const float* mask, float* out, int32_t B, int32_t T1, int32_t K, void* stream);

/* model.get_mask (model.py:172-209) on the device: frame_mask [size] -> out [size,size]. */
int kit_get_mask(const float* frame_mask, int32_t size, int32_t matrix_type, float* out, void* stream);

/* Adam over the flat arena (A1_train.py:135,256: torch.optim.Adam defaults, no weight decay,
 * no amsgrad).  step is 1-based.  grad_scale multiplies the gradient first (1/world for DP). */
int kit_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, int32_t step, float grad_scale, void* stream);
/* Same update with the step count and the learning rate on the device, so that the step can be captured in a CUDA graph.
 * state: 16 bytes, 16-byte aligned: {int32 completed_steps; float lr; float step_size; float inv_sqrt_bc2} -- the host
 * writes completed_steps (checkpoint restore) and lr (A1_train.py:42-54); the call advances completed_steps by one. */
int kit_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, void* state,
                      float beta1, float beta2, float eps, float grad_scale, void* stream);
/* The same update over a sub-range of the arena (pointers already offset; 16-byte aligned, n a multiple of 4): data-parallel
 * training steps each gradient bucket as soon as its all-reduce has completed.  advance_step != 0 on the first range of an
 * optimiser step only (it advances completed_steps and recomputes the bias corrections). */
int kit_adam_step_dev_range(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, void* state,
                            float beta1, float beta2, float eps, float grad_scale, int32_t advance_step, void* stream);

/* ------------------------------------------------------------------------------------------
 * Building blocks exposed for unit tests (tests/ call these through the same ABI)
 * ---------------------------------------------------------------------------------------- */
/* tcgen05 GEMM.  C[M,N] = op(A) * op(B) (+bias[N]) (+addend[M,N] bf16):
 *   mode 0 (TN): A [M,K] bf16 row-major (ld lda), B [N,K] bf16 row-major (ld ldb)     y = x W^T
 *   mode 1 (wgrad): A [K,M] bf16 row-major, B [K,N] bf16 row-major, C fp32 += A^T B   dW = dy^T x
 * out_kind: 0 = bf16, 1 = fp32 overwrite, 2 = fp32 atomic accumulate (split-K allowed).
 * act: 0 none, 1 gelu (aux_out receives the pre-activation as bf16), 2 multiply by gelu'(aux_in). */
int kit_gemm_bf16(int32_t mode, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                  int32_t M, int32_t N, int32_t K, const float* bias, const void* addend, int64_t ld_addend,
                  int32_t out_kind, int32_t act, void* aux, int64_t ld_aux, int32_t split_k, void* stream);

/* The feed-forward block of a post-norm layer in one kernel (torch/nn/modules/transformer.py:956-959):
 * s = x + linear2(gelu(linear1(x))), y = LayerNorm(s) * gamma + beta, per-row mean / rstd of s.  H must be 256, FF a
 * multiple of 128.  x, s, y: bf16 [M,H]; w1 [FF,H], w2 [H,FF] bf16 row-major; store_zh != 0 also writes the
 * pre-activation z and the activation hh (bf16 [M,FF]) that the backward pass reads. */
int kit_ffn_fwd(const void* x, const void* w1, const void* w2, const float* b1, const float* b2, const float* gamma,
                const float* beta, void* z, void* hh, void* s, void* y, float* mean, float* rstd, int32_t M, int32_t H,
                int32_t FF, int32_t store_zh, void* stream);

/* Input gradients of that block in one kernel: dz = (g w2) * gelu'(z) -> dz [M,FF] (what the weight gradients read),
 * dx = dz w1 + g -> dx [M,H] (g = gradient w.r.t. the pre-norm sum s).  w2t = w2^T [FF,H], w1t = w1^T [H,FF], bf16 row-major. */
int kit_ffn_bwd(const void* g, const void* w2t, const void* w1t, const void* z, void* dz, void* dx, int32_t M, int32_t H,
                int32_t FF, void* stream);

/* LayerNorm backward in the epilogue of the GEMM that produces the gradient w.r.t. the LayerNorm output (tests):
 * dy = A B^T + addend (A [M,K], B [256,K], addend [M,256], bf16); dx = rstd (dy gamma - mean(dy gamma) - xhat mean(dy gamma xhat))
 * with xhat = (s - mean) rstd from the saved pre-norm sum s [M,256]; dgamma += sum_rows dy xhat, dbeta += sum_rows dy. */
int kit_gemm_lnbwd(const void* A, const void* B, const void* addend, const void* s, const float* gamma, const float* mean,
                   const float* rstd, void* dx, float* dgamma, float* dbeta, int32_t M, int32_t K, void* stream);

/* softmax(Q K^T / sqrt(d) + mask) V for B*NH heads.  q/k/v: bf16, element (b, t, h, c) at
 * ptr + (b*S + t)*ld + h*d + c.  out: bf16 [B*Sq, NH*d] (ld_o).  lse: fp32 [B, NH, Sq]. */
int kit_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                      int64_t ldo, float* lse, int32_t B, int32_t NH, int32_t Sq, int32_t Sk, int32_t d,
                      const KitAttnMask* mask, void* stream);
/* dq_accum: fp32 workspace of B*Sq*NH*d + B*NH*Sq floats, required only when Sk > 64 (dQ summed over key tiles, followed by
 * the per-row sums of dO * O). */
int kit_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                      const void* out, int64_t ldo, const void* dout, int64_t ld_do, const float* lse, void* dq,
                      int64_t ld_dq, void* dk, int64_t ld_dk, void* dv, int64_t ld_dv, float* dq_accum, int32_t B,
                      int32_t NH, int32_t Sq, int32_t Sk, int32_t d, const KitAttnMask* mask, void* stream);

/* Test-only probe of the tensor-core operand layouts: ONE sequence of `steps` tcgen05.mma (M x N x 16, kind::f16, bf16 in,
 * fp32 accumulate) from two operand tiles the caller lays out byte by byte (a_bytes / b_bytes are copied verbatim to 1024-byte
 * aligned shared memory; descriptors are built from args), accumulator read back as out[128 lanes][n_cols] fp32 (lanes the
 * instruction does not write hold -12345).  args_host: 17 int32 on the HOST = {idesc, steps, a_off, a_step, a_lbo, a_sbo,
 * a_layout, b_off, b_step, b_lbo, b_sbo, b_layout, a_from_tmem, a_tmem_cols, a_tmem_step, d_lane, n_cols}. */
int kit_umma_probe(const void* a_bytes, int32_t a_len, const void* b_bytes, int32_t b_len, const int32_t* args_host,
                   float* out, void* stream);

/* Row kernels over [M,H] bf16 (H multiple of 8, <= 1024).  See csrc/rowops.cu. */
int kit_add_layernorm_fwd(const void* a, const void* b, const float* gamma, const float* beta, void* sum_out,
                          void* y, float* mean, float* rstd, int64_t M, int32_t H, void* stream);
int kit_layernorm_bwd(const void* dy, const void* sum_saved, const float* mean, const float* rstd, const float* gamma,
                      const void* addend, void* dx, float* dgamma, float* dbeta, int64_t M, int32_t H, void* stream);
int kit_cast_fp32_to_bf16_padded(const float* src, int64_t rows, int64_t cols, int64_t src_ld, void* dst,
                                 int64_t dst_ld, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KIT_B200_H_ */

// Host side of the tcgen05 GEMM: TMA descriptor construction, planning and launch.
#include "gemm_sm100.cuh"
#define KIT_WGRAD_GROUP_IMPL
#include "gemm_wgrad_group.cuh"
#define KIT_FFN_IMPL
#include "ffn_fused.cuh"

#include <mutex>
#include <cstdlib>

namespace kit {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

static int make_tensor_map_2d_typed(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, uint64_t inner,
                                    uint64_t outer, uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                                    CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  KIT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KIT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return KIT_OK;
}

int make_tensor_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_pitch_bytes,
                       uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  KIT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  KIT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base address must be 16-byte aligned");
  KIT_REQUIRE((row_pitch_bytes & 15) == 0, "TMA row pitch must be a multiple of 16 bytes (got %llu)",
              (unsigned long long)row_pitch_bytes);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KIT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return KIT_OK;
}

// [d2][d1][d0] bf16 tensor (d0 contiguous; strides of d1 / d2 in bytes), box (box0, box1, box2), 128B swizzle, out-of-range
// elements read as zeros: the T <= 64 attention kernels load [2 sequences][64 frames][64 columns] tiles of [B, T, ld] tensors.
int make_tensor_map_3d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                       uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2) {
  EncodeTiledFn fn = get_encode_fn();
  KIT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  KIT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (stride1_bytes & 15) == 0 && (stride2_bytes & 15) == 0,
              "TMA base address and strides must be multiples of 16 bytes");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KIT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
  return KIT_OK;
}

static int g_num_sms = 0;
static long long* g_trace_buf = nullptr;

// every (tile width, mode, epilogue kind) the planner can select; CTA pairs (cta_group::2) for the 256-wide tiles
#define KIT_GEMM_KERNEL(BN, MODE, EPI) gemm_tcgen05_kernel<BN, MODE, ((BN) == 256 ? 2 : 1), EPI>
#define KIT_GEMM_SMEM(BN, EPI) gemm_smem_bytes<BN, ((BN) == 256 ? 2 : 1), EPI>()
#define KIT_GEMM_FOR_ALL(X)                                                                                     \
  X(128, 0, EPI_STORE) X(128, 0, EPI_ADD) X(128, 0, EPI_GELU) X(128, 0, EPI_GELU_BWD) X(128, 0, EPI_F32) X(128, 0, EPI_GENERIC) \
  X(256, 0, EPI_STORE) X(256, 0, EPI_ADD) X(256, 0, EPI_GELU) X(256, 0, EPI_GELU_BWD) X(256, 0, EPI_F32) X(256, 0, EPI_GENERIC) \
  X(256, 0, EPI_ADD_LN) X(256, 0, EPI_ADD_LNBWD)                                                                 \
  X(128, 1, EPI_F32) X(128, 1, EPI_GENERIC) X(256, 1, EPI_F32) X(256, 1, EPI_GENERIC)
#define KIT_GEMM_KERNEL_BG(BN) gemm_tcgen05_kernel<BN, 1, ((BN) == 256 ? 2 : 1), EPI_F32, true>

int gemm_init_attributes() {
  static int status = 1;
  static std::once_flag once;
  std::call_once(once, []() {
    status = 0;
#define KIT_SET_ATTR(BN, MODE, EPI)                                                                                         \
  {                                                                                                                         \
    cudaError_t e = cudaFuncSetAttribute(KIT_GEMM_KERNEL(BN, MODE, EPI), cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         KIT_GEMM_SMEM(BN, EPI));                                                           \
    if (e != cudaSuccess) {                                                                                                 \
      status = -1;                                                                                                          \
      set_error("cudaFuncSetAttribute(max dynamic smem, BN=%d mode=%d epi=%d) failed: %s", BN, MODE, EPI, cudaGetErrorString(e)); \
    }                                                                                                                       \
  }
    KIT_GEMM_FOR_ALL(KIT_SET_ATTR)
#undef KIT_SET_ATTR
    if (cudaFuncSetAttribute(KIT_GEMM_KERNEL_BG(128), cudaFuncAttributeMaxDynamicSharedMemorySize, KIT_GEMM_SMEM(128, EPI_F32)) != cudaSuccess ||
        cudaFuncSetAttribute(KIT_GEMM_KERNEL_BG(256), cudaFuncAttributeMaxDynamicSharedMemorySize, KIT_GEMM_SMEM(256, EPI_F32)) != cudaSuccess) {
      status = -1;
      set_error("cudaFuncSetAttribute(max dynamic smem) failed for the bias-gradient weight-gradient kernels");
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  });
  return status == 0 ? KIT_OK : KIT_ERR_CUDA;
}

// EPI_ADD_LNBWD stages the saved-sum tiles through the operand ring once the accumulator is complete: one 256-row item per CTA pair
bool gemm_lnbwd_supported(int M) {
  if (gemm_init_attributes()) return false;
  return (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM) <= (g_num_sms - sm_reserve()) / 2;
}

static bool aligned16(const void* ptr, int64_t ld_elems, size_t esize) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ((size_t)ld_elems * esize) % 16 == 0;
}

int gemm_plan(GemmPlan* plan, int mode, const bf16* A, int64_t lda, const bf16* B, int64_t ldb, void* C, int64_t ldc,
              int M, int N, int K, const float* bias, const bf16* addend, int64_t ld_addend, int out_kind, int act,
              bf16* aux, int64_t ld_aux, int split_k, float* bias_grad, const GemmLN* ln, const LnBwdArgs* lnb) {
  KIT_REQUIRE(mode == 0 || mode == 1, "gemm mode must be 0 (TN) or 1 (wgrad)");
  KIT_REQUIRE(M > 0 && N > 0 && K > 0, "gemm dims must be positive (M=%d N=%d K=%d)", M, N, K);
  KIT_REQUIRE(act == ACT_NONE || aux != nullptr, "gelu epilogues need the aux tensor");
  int rc = gemm_init_attributes();
  if (rc) return rc;
  GemmParams& p = plan->p;
  p.M = M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc;
  p.bias = bias;
  p.addend = addend; p.ld_addend = ld_addend;
  p.aux = aux; p.ld_aux = ld_aux;
  p.out_kind = out_kind; p.act = act;
  p.trace = nullptr;
  if (getenv("KIT_GEMM_TRACE") != nullptr) {
    static long long* trace_buf = nullptr;
    if (trace_buf == nullptr && cudaMalloc(&trace_buf, 16 * sizeof(long long)) != cudaSuccess) trace_buf = nullptr;
    p.trace = trace_buf;
    g_trace_buf = trace_buf;
  }
  plan->mode = mode;
  const int bn = (N > 128) ? 256 : 128;
  const int cl = (bn == 256) ? 2 : 1;
  plan->bn = bn;
  p.tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  p.tiles_n = (N + bn - 1) / bn;
  const int kb_total = (K + GEMM_BK - 1) / GEMM_BK;
  const int groups = ((p.tiles_m + cl - 1) / cl) * p.tiles_n;
  int splits = 1;
  if (mode == 0) {
    // A [M,K] K-major: box 64(k) x 128(m);  B [N,K] K-major: box 64(k) x BN/cl (n) -- each CTA of a pair stages half of B
    if ((rc = make_tensor_map_2d(&plan->tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, GEMM_BK, GEMM_BM))) return rc;
    if ((rc = make_tensor_map_2d(&plan->tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, GEMM_BK, bn / cl))) return rc;
    splits = split_k > 1 ? split_k : 1;
  } else {
    // A [K,M] row-major (MN-major operand): box 64(m) x 64(k);  B [K,N]: box 64(n) x 64(k)
    if ((rc = make_tensor_map_2d(&plan->tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, GEMM_BK))) return rc;
    if ((rc = make_tensor_map_2d(&plan->tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, GEMM_BK))) return rc;
    splits = split_k;
    if (splits <= 0) splits = ((g_num_sms - sm_reserve()) / cl) / groups;   // about one work item per CTA (pair)
  }
  // epilogue kind: outputs (and the residual / GELU' input) through shared memory + TMA whenever tensor maps can describe them
  const size_t esize = (out_kind == OUT_BF16) ? 2 : 4;
  int epi = EPI_GENERIC;
  const bool c_ok = aligned16(C, ldc, esize) && (bias == nullptr || (N % 8) == 0);
  // fp32 rows of an even number of floats that TMA cannot address (row starts only 8-byte aligned, e.g. N = ldc = 142): the staged
  // EPI_F32 epilogue with a coalesced copy-out by the warp instead of the per-lane scalar path
  const bool manual_ok = !c_ok && out_kind != OUT_BF16 && act == ACT_NONE && addend == nullptr && (ldc % 2) == 0 &&
                         (reinterpret_cast<uintptr_t>(C) & 7) == 0 && (bias == nullptr || aligned16(bias, 0, 4)) && bias_grad == nullptr &&
                         getenv("KIT_GEMM_NO_MANUAL_OUT") == nullptr;
  p.manual_out = 0;
  if (manual_ok) {
    epi = EPI_F32;
    p.manual_out = 1;
  }
  if (c_ok) {
    if (out_kind != OUT_BF16) {
      if (act == ACT_NONE && addend == nullptr) epi = EPI_F32;
    } else if (mode == 0) {
      if (act == ACT_NONE && addend == nullptr) epi = EPI_STORE;
      else if (act == ACT_NONE && aligned16(addend, ld_addend, 2)) {
        epi = EPI_ADD;
        // LayerNorm of the full row in the same epilogue: one 256-wide tile per row, every pointer TMA / float4 friendly
        if (ln != nullptr && N == 256 && bn == 256 && aligned16(ln->y, ln->ldy, 2) && aligned16(ln->gamma, 0, 4) &&
            aligned16(ln->beta, 0, 4) && (bias == nullptr || aligned16(bias, 0, 4)))
          epi = EPI_ADD_LN;
        else if (lnb != nullptr && lnb->s != nullptr && N == 256 && bn == 256 && aligned16(lnb->s, lnb->ld_s, 2) &&
                 aligned16(lnb->gamma, 0, 4) && (bias == nullptr || aligned16(bias, 0, 4)))
          epi = EPI_ADD_LNBWD;
      }
      else if (act == ACT_GELU && addend == nullptr && aligned16(aux, ld_aux, 2)) epi = EPI_GELU;
      else if (act == ACT_GELU_BWD && addend == nullptr && aligned16(aux, ld_aux, 2)) epi = EPI_GELU_BWD;
    }
  }
  if (epi == EPI_ADD_LNBWD && groups > (g_num_sms - sm_reserve()) / cl) epi = EPI_ADD;   // the fused kind stages through the operand ring: one item per CTA
  plan->epi = epi;
  p.bias_grad = (mode == 1 && epi == EPI_F32) ? bias_grad : nullptr;   // else the caller sums the columns of dy itself
  p.lnb = LnBwdArgs{nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (epi == EPI_ADD_LNBWD) p.lnb = *lnb;
  p.ln_gamma = p.ln_beta = nullptr;
  p.ln_mean = p.ln_rstd = nullptr;
  p.ln_eps = 0.f;
  plan->tmD = plan->tmA;
  if (epi == EPI_ADD_LN) {
    p.ln_gamma = ln->gamma; p.ln_beta = ln->beta; p.ln_mean = ln->mean; p.ln_rstd = ln->rstd; p.ln_eps = ln->eps;
  }
  plan->tmC = plan->tmA;
  plan->tmAux = plan->tmA;
  if (epi == EPI_F32 && p.manual_out) {
    // no tensor map for C: the warp copies the staged tile out itself
  } else if (epi == EPI_F32) {
    if ((rc = make_tensor_map_2d_typed(&plan->tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc * 4, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  } else if (epi != EPI_GENERIC) {
    if ((rc = make_tensor_map_2d_typed(&plan->tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    const bool add_kind = epi == EPI_ADD || epi == EPI_ADD_LN || epi == EPI_ADD_LNBWD;
    const bf16* side = add_kind ? addend : aux;
    const int64_t side_ld = add_kind ? ld_addend : ld_aux;
    if (epi == EPI_ADD_LNBWD) {
      if ((rc = make_tensor_map_2d_typed(&plan->tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, lnb->s, (uint64_t)N, (uint64_t)M, (uint64_t)lnb->ld_s * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    }
    if (epi == EPI_ADD_LN) {
      if ((rc = make_tensor_map_2d_typed(&plan->tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, ln->y, (uint64_t)N, (uint64_t)M, (uint64_t)ln->ldy * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    }
    if (epi != EPI_STORE) {
      if ((rc = make_tensor_map_2d_typed(&plan->tmAux, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, side, (uint64_t)N, (uint64_t)M, (uint64_t)side_ld * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    }
  }
  if (splits > kb_total) splits = kb_total;
  if (splits < 1) splits = 1;
  KIT_REQUIRE(splits == 1 || out_kind == OUT_F32_ATOMIC, "split-K needs the atomic fp32 epilogue");
  p.kb_per_split = (kb_total + splits - 1) / splits;
  p.splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
  const int items = groups * p.splits;   // work items per CTA (pair)
  const int max_clusters = (g_num_sms - sm_reserve()) / cl;
  plan->grid = (items < max_clusters ? items : max_clusters) * cl;
  return KIT_OK;
}

template <typename KernelT>
static int launch_one(KernelT kernel, int grid, int threads, int smem, int cluster, cudaStream_t stream, const GemmPlan* plan) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  KIT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, plan->tmA, plan->tmB, plan->tmC, plan->tmAux, plan->tmD, plan->p));
  return KIT_OK;
}

int gemm_launch(const GemmPlan* plan, cudaStream_t stream) {
  if (plan->p.bias_grad != nullptr) {
    if (plan->bn == 128)
      return launch_one(KIT_GEMM_KERNEL_BG(128), plan->grid, gemm_threads<128>(), KIT_GEMM_SMEM(128, EPI_F32), 1, stream, plan);
    return launch_one(KIT_GEMM_KERNEL_BG(256), plan->grid, gemm_threads<256>(), KIT_GEMM_SMEM(256, EPI_F32), 2, stream, plan);
  }
#define KIT_DISPATCH(BN, MODE, EPI)                                                                                          \
  if (plan->bn == BN && plan->mode == MODE && plan->epi == EPI)                                                              \
    return launch_one(KIT_GEMM_KERNEL(BN, MODE, EPI), plan->grid, gemm_threads<BN>(), KIT_GEMM_SMEM(BN, EPI), (BN) == 256 ? 2 : 1, \
                      stream, plan);
  KIT_GEMM_FOR_ALL(KIT_DISPATCH)
#undef KIT_DISPATCH
  set_error("gemm_launch: no kernel for bn=%d mode=%d epi=%d", plan->bn, plan->mode, plan->epi);
  return KIT_ERR_INVALID;
}

// ---------------------------------------------------------------- grouped stream-K weight gradients
bool wgrad_group_supported(const WgradProblemDesc& d) {
  return aligned16(d.A, d.lda, 2) && aligned16(d.B, d.ldb, 2) && aligned16(d.C, d.ldc, 4) && d.M > 0 && d.N > 0;
}

int wgrad_group_plan(WgradGroupPlan* plan, const WgradProblemDesc* probs, int n, int K) {
  KIT_REQUIRE(n >= 1 && n <= WG_MAX_PROBLEMS, "wgrad group: 1..%d problems (got %d)", WG_MAX_PROBLEMS, n);
  KIT_REQUIRE(K > 0, "wgrad group: K must be positive");
  int rc = gemm_init_attributes();
  if (rc) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(gemm_wgrad_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
  });
  KIT_REQUIRE(attr_err == cudaSuccess, "wgrad group: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  WgradGroupParams& p = plan->p;
  p.n_problems = n;
  p.kb_total = (K + GEMM_BK - 1) / GEMM_BK;
  int units = 0;
  for (int i = 0; i < n; ++i) {
    const WgradProblemDesc& d = probs[i];
    KIT_REQUIRE(wgrad_group_supported(d), "wgrad group: problem %d violates the TMA alignment rules", i);
    WgradProblem& q = p.prob[i];
    q.M = d.M;
    q.N = d.N;
    q.groups_m = (d.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
    q.tiles_n = (d.N + WG_BN - 1) / WG_BN;
    q.unit_begin = units;
    q.bias_grad = d.bias_grad;
    units += q.groups_m * q.tiles_n * p.kb_total;
    if ((rc = make_tensor_map_2d(&plan->maps.a[i], d.A, (uint64_t)d.M, (uint64_t)K, (uint64_t)d.lda * 2, 64, GEMM_BK))) return rc;
    if ((rc = make_tensor_map_2d(&plan->maps.b[i], d.B, (uint64_t)d.N, (uint64_t)K, (uint64_t)d.ldb * 2, 64, GEMM_BK))) return rc;
    if ((rc = make_tensor_map_2d_typed(&plan->maps.c[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d.C, (uint64_t)d.N, (uint64_t)d.M,
                                       (uint64_t)d.ldc * 4, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  for (int i = n; i < WG_MAX_PROBLEMS; ++i) {
    p.prob[i] = p.prob[0];
    p.prob[i].unit_begin = units;
    plan->maps.a[i] = plan->maps.a[0];
    plan->maps.b[i] = plan->maps.b[0];
    plan->maps.c[i] = plan->maps.c[0];
  }
  p.total_units = units;
  const int max_clusters = (g_num_sms - sm_reserve()) / WG_CL;
  const int clusters = units < max_clusters ? units : max_clusters;
  p.units_per_cluster = (units + clusters - 1) / clusters;
  plan->grid = ((units + p.units_per_cluster - 1) / p.units_per_cluster) * WG_CL;
  return KIT_OK;
}

int wgrad_group_launch(const WgradGroupPlan* plan, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan->grid);
  cfg.blockDim = dim3(gemm_threads<WG_BN>());
  cfg.dynamicSmemBytes = WG_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = WG_CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  KIT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_wgrad_group_kernel, plan->maps, plan->p));
  return KIT_OK;
}

// ---------------------------------------------------------------- fused feed-forward block (ffn_fused.cuh)
static long long* g_ffn_trace = nullptr;
bool ffn_fwd_supported(int H, int FF) { return H == FFN_H && FF >= FFN_FC && FF % FFN_FC == 0; }

// KIT_FFN_XT: bit 0 = forward, bit 1 = backward take the resident tile from tensor memory.  Default OFF: measured no faster
// (profiles/r02_summary.md: the chunk loop is paced by the 16 epilogue warps -- tcgen05.ld, GELU, pack, store -- not by the
// tensor pipe, and the single-buffered acc1 the 512 columns then allow costs the training variants 2-5 %).
static int ffn_xt_mask() {
  const char* v = getenv("KIT_FFN_XT");
  return v != nullptr ? atoi(v) : 0;
}
int ffn_fwd_plan(FfnPlan* plan, const bf16* x, int64_t ldx, const bf16* w1, int64_t ldw1, const bf16* w2, int64_t ldw2,
                 const float* b1, const float* b2, bf16* z, bf16* hh, int64_t ldzh, bf16* s, int64_t lds, bf16* y, int64_t ldy,
                 const float* gamma, const float* beta, float* mean, float* rstd, float eps, int M, int H, int FF, int store_zh) {
  KIT_REQUIRE(ffn_fwd_supported(H, FF), "fused FFN: H must be %d and FF a multiple of %d (got %d, %d)", FFN_H, FFN_FC, H, FF);
  KIT_REQUIRE(M > 0 && x && w1 && w2 && b1 && b2 && s && y && gamma && beta && mean && rstd, "fused FFN: null argument");
  KIT_REQUIRE(!store_zh || (z != nullptr && hh != nullptr), "fused FFN: training needs the z and h output tensors");
  KIT_REQUIRE(aligned16(b1, 0, 4) && aligned16(b2, 0, 4) && aligned16(gamma, 0, 4) && aligned16(beta, 0, 4),
              "fused FFN: bias / LayerNorm vectors must be 16-byte aligned");
  int rc = gemm_init_attributes();
  if (rc) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(ffn_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ffn_smem<false>());
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(ffn_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ffn_smem<false>());
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(ffn_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ffn_smem<true>());
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(ffn_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ffn_smem<true>());
  });
  KIT_REQUIRE(attr_err == cudaSuccess, "fused FFN: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  if ((rc = make_tensor_map_2d(&plan->tmX, x, (uint64_t)H, (uint64_t)M, (uint64_t)ldx * 2, 64, 128))) return rc;
  if ((rc = make_tensor_map_2d(&plan->tmW1, w1, (uint64_t)H, (uint64_t)FF, (uint64_t)ldw1 * 2, 64, 64))) return rc;
  if ((rc = make_tensor_map_2d(&plan->tmW2, w2, (uint64_t)FF, (uint64_t)H, (uint64_t)ldw2 * 2, 64, 128))) return rc;
  plan->tmZ = plan->tmX;
  plan->tmHh = plan->tmX;
  if (store_zh) {
    if ((rc = make_tensor_map_2d(&plan->tmZ, z, (uint64_t)FF, (uint64_t)M, (uint64_t)ldzh * 2, 64, 32))) return rc;
    if ((rc = make_tensor_map_2d(&plan->tmHh, hh, (uint64_t)FF, (uint64_t)M, (uint64_t)ldzh * 2, 64, 32))) return rc;
  }
  if ((rc = make_tensor_map_2d_typed(&plan->tmS, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, s, (uint64_t)H, (uint64_t)M, (uint64_t)lds * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tensor_map_2d_typed(&plan->tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, y, (uint64_t)H, (uint64_t)M, (uint64_t)ldy * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  FfnParams& p = plan->p;
  p.M = M; p.FF = FF; p.n_items = (M + 255) / 256;
  p.b1 = b1; p.b2 = b2; p.ln_gamma = gamma; p.ln_beta = beta; p.ln_mean = mean; p.ln_rstd = rstd; p.ln_eps = eps;
  p.store_zh = store_zh;
  p.lnb = LnBwdArgs{nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr};
  plan->bwd = 0;
  plan->xt = ffn_xt_mask() & 1;
  p.trace = nullptr;
  if (getenv("KIT_FFN_TRACE") != nullptr) {
    if (g_ffn_trace == nullptr && cudaMalloc(&g_ffn_trace, 256 * sizeof(long long)) != cudaSuccess) g_ffn_trace = nullptr;
    if (g_ffn_trace != nullptr) cudaMemset(g_ffn_trace, 0, 256 * sizeof(long long));
    p.trace = g_ffn_trace;
  }
  const int max_clusters = (g_num_sms - sm_reserve()) / 2;
  plan->grid = (p.n_items < max_clusters ? p.n_items : max_clusters) * 2;
  return KIT_OK;
}

int ffn_bwd_plan(FfnPlan* plan, const bf16* g, int64_t ldg, const bf16* w2t, int64_t ldw2t, const bf16* w1t, int64_t ldw1t,
                 const bf16* z, bf16* dz_out, int64_t ldzh, bf16* dx, int64_t lddx, int M, int H, int FF, const LnBwdArgs* lnb) {
  // the forward plan with x := g, W1 := W2^T, W2 := W1^T; the z map becomes a LOAD map ([128 x 64] boxes)
  static const float dummy[4] = {0.f, 0.f, 0.f, 0.f};
  KIT_REQUIRE(z != nullptr && dz_out != nullptr && dx != nullptr, "fused FFN backward: null tensor");
  float* unused = const_cast<float*>(dummy);
  int rc = ffn_fwd_plan(plan, g, ldg, w2t, ldw2t, w1t, ldw1t, dummy, dummy, const_cast<bf16*>(z), dz_out, ldzh, dx, lddx, dx, lddx, dummy,
                        dummy, unused, unused, 0.f, M, H, FF, 1);
  if (rc) return rc;
  if ((rc = make_tensor_map_2d(&plan->tmZ, z, (uint64_t)FF, (uint64_t)M, (uint64_t)ldzh * 2, 64, 128))) return rc;
  plan->p.b1 = plan->p.b2 = plan->p.ln_gamma = plan->p.ln_beta = nullptr;
  plan->p.ln_mean = plan->p.ln_rstd = nullptr;
  plan->p.store_zh = 0;
  plan->bwd = 1;
  plan->xt = (ffn_xt_mask() >> 1) & 1;
  if (lnb != nullptr && lnb->s != nullptr) {
    KIT_REQUIRE(aligned16(lnb->s, lnb->ld_s, 2) && aligned16(lnb->gamma, 0, 4), "fused FFN backward: LayerNorm tensors must be 16-byte aligned");
    plan->p.lnb = *lnb;   // tmY (unused by the backward otherwise) becomes the load map of the saved sum: [32 x 32] boxes
    if ((rc = make_tensor_map_2d_typed(&plan->tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, lnb->s, (uint64_t)H, (uint64_t)M,
                                       (uint64_t)lnb->ld_s * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  }
  return KIT_OK;
}

int ffn_launch(const FfnPlan* plan, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan->grid);
  cfg.blockDim = dim3(FFN_THREADS);
  cfg.dynamicSmemBytes = plan->bwd ? ffn_smem<true>() : ffn_smem<false>();
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
#define KIT_FFN_LAUNCH(B, X)                                                                                              \
  KIT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ffn_kernel<B, X>, plan->tmX, plan->tmW1, plan->tmW2, plan->tmZ, plan->tmHh, plan->tmS, \
                                    plan->tmY, plan->p))
  if (plan->bwd) {
    if (plan->xt) KIT_FFN_LAUNCH(true, true); else KIT_FFN_LAUNCH(true, false);
  } else {
    if (plan->xt) KIT_FFN_LAUNCH(false, true); else KIT_FFN_LAUNCH(false, false);
  }
#undef KIT_FFN_LAUNCH
  return KIT_OK;
}

}  // namespace kit

// experiments only: the 128 clock64 marks of CTA 0 of the last traced fused FFN launch (KIT_FFN_TRACE=1)
extern "C" int kit_ffn_trace_read(long long* out128) {
  if (kit::g_ffn_trace == nullptr) return KIT_ERR_INVALID;
  return cudaMemcpy(out128, kit::g_ffn_trace, 256 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? KIT_OK : KIT_ERR_CUDA;
}

// experiments only: the 16 clock64 marks of CTA 0 of the last traced GEMM (KIT_GEMM_TRACE=1)
extern "C" int kit_gemm_trace_read(long long* out16) {
  if (kit::g_trace_buf == nullptr) return KIT_ERR_INVALID;
  return cudaMemcpy(out16, kit::g_trace_buf, 16 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? KIT_OK : KIT_ERR_CUDA;
}

// Host side of the tcgen05 GEMM: TMA descriptor construction, planning and launch.
#include "gemm_sm100.cuh"

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

constexpr int STAGES_BN128 = 4, STAGES_BN256 = 3;
constexpr int CLUSTER_BN256 = 2;   // CTA pairs share the B tile through TMA multicast

static int g_num_sms = 0;

int gemm_init_attributes() {
  static int status = 1;
  static std::once_flag once;
  std::call_once(once, []() {
    cudaError_t e[4];
    e[0] = cudaFuncSetAttribute(gemm_tcgen05_kernel<128, 0, STAGES_BN128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                gemm_smem_bytes<128, STAGES_BN128>());
    e[1] = cudaFuncSetAttribute(gemm_tcgen05_kernel<256, 0, STAGES_BN256, CLUSTER_BN256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                gemm_smem_bytes<256, STAGES_BN256>());
    e[2] = cudaFuncSetAttribute(gemm_tcgen05_kernel<128, 1, STAGES_BN128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                gemm_smem_bytes<128, STAGES_BN128>());
    e[3] = cudaFuncSetAttribute(gemm_tcgen05_kernel<256, 1, STAGES_BN256, CLUSTER_BN256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                gemm_smem_bytes<256, STAGES_BN256>());
    status = 0;
    for (int i = 0; i < 4; ++i) {
      if (e[i] != cudaSuccess) {
        status = -1;
        set_error("cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(e[i]));
      }
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  });
  return status == 0 ? KIT_OK : KIT_ERR_CUDA;
}

int gemm_plan(GemmPlan* plan, int mode, const bf16* A, int64_t lda, const bf16* B, int64_t ldb, void* C, int64_t ldc,
              int M, int N, int K, const float* bias, const bf16* addend, int64_t ld_addend, int out_kind, int act,
              bf16* aux, int64_t ld_aux, int split_k) {
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
  { const char* d = getenv("KIT_GEMM_DBG"); p.dbg = d ? atoi(d) : 0; }
  plan->mode = mode;
  const int bn = (N > 128) ? 256 : 128;
  plan->bn = bn;
  p.tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  p.tiles_n = (N + bn - 1) / bn;
  const int tiles = p.tiles_m * p.tiles_n;
  const int kb_total = (K + GEMM_BK - 1) / GEMM_BK;
  int splits = 1;
  if (mode == 0) {
    // A [M,K] K-major: box 64(k) x 128(m);  B [N,K] K-major: box 64(k) x BN(n)
    if ((rc = make_tensor_map_2d(&plan->tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, GEMM_BK, GEMM_BM))) return rc;
    const int cl = (bn == 256) ? CLUSTER_BN256 : 1;
    if ((rc = make_tensor_map_2d(&plan->tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, GEMM_BK, bn / cl))) return rc;
    splits = split_k > 1 ? split_k : 1;
  } else {
    // A [K,M] row-major (MN-major operand): box 64(m) x 64(k);  B [K,N]: box 64(n) x 64(k)
    if ((rc = make_tensor_map_2d(&plan->tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, GEMM_BK))) return rc;
    if ((rc = make_tensor_map_2d(&plan->tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, GEMM_BK))) return rc;
    splits = split_k;
    if (splits <= 0) {   // about one work item per cluster
      const int clm = (bn == 256) ? CLUSTER_BN256 : 1;
      const int groups = ((p.tiles_m + clm - 1) / clm) * p.tiles_n;
      splits = (g_num_sms / clm) / groups;
    }
  }
  // outputs through shared memory + TMA (store / fp32 reduce-add) whenever the layout allows a tensor map
  const size_t esize = (out_kind == OUT_BF16) ? 2 : 4;
  p.tma_store = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (((size_t)ldc * esize) % 16 == 0);
  if (act == ACT_GELU) p.tma_store = p.tma_store && ((reinterpret_cast<uintptr_t>(aux) & 15) == 0) && (((size_t)ld_aux * 2) % 16 == 0);
  if (p.tma_store) {
    if (out_kind == OUT_BF16) {
      if ((rc = make_tensor_map_2d_typed(&plan->tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    } else {
      if ((rc = make_tensor_map_2d_typed(&plan->tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc * 4, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    }
    plan->tmAux = plan->tmC;
    if (act == ACT_GELU) {
      if ((rc = make_tensor_map_2d_typed(&plan->tmAux, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, aux, (uint64_t)N, (uint64_t)M, (uint64_t)ld_aux * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    }
  } else {
    plan->tmC = plan->tmA;
    plan->tmAux = plan->tmA;
  }
  // epilogue input tile (residual addend or GELU' pre-activation) through TMA as well
  p.tma_in = 0;
  if (p.tma_store && out_kind == OUT_BF16 && act != ACT_GELU) {
    const bf16* in_ptr = (act == ACT_GELU_BWD) ? aux : addend;
    const int64_t in_ld = (act == ACT_GELU_BWD) ? ld_aux : ld_addend;
    const bool both = (act == ACT_GELU_BWD) && addend != nullptr;
    if (in_ptr != nullptr && !both && (reinterpret_cast<uintptr_t>(in_ptr) & 15) == 0 && ((size_t)in_ld * 2) % 16 == 0) {
      if ((rc = make_tensor_map_2d_typed(&plan->tmAux, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, in_ptr, (uint64_t)N, (uint64_t)M, (uint64_t)in_ld * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
      p.tma_in = 1;
    }
  }
  if (splits > kb_total) splits = kb_total;
  if (splits < 1) splits = 1;
  KIT_REQUIRE(splits == 1 || out_kind == OUT_F32_ATOMIC, "split-K needs the atomic fp32 epilogue");
  p.kb_per_split = (kb_total + splits - 1) / splits;
  p.splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
  const int clw = (bn == 256) ? CLUSTER_BN256 : 1;
  const int items = ((p.tiles_m + clw - 1) / clw) * p.tiles_n * p.splits;   // work items per cluster
  const int max_clusters = g_num_sms / clw;
  plan->grid = (items < max_clusters ? items : max_clusters) * clw;
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
  KIT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, plan->tmA, plan->tmB, plan->tmC, plan->tmAux, plan->p));
  return KIT_OK;
}

int gemm_launch(const GemmPlan* plan, cudaStream_t stream) {
  if (plan->mode == 0 && plan->bn == 128)
    return launch_one(gemm_tcgen05_kernel<128, 0, STAGES_BN128, 1>, plan->grid, gemm_threads<128>(), gemm_smem_bytes<128, STAGES_BN128>(), 1, stream, plan);
  if (plan->mode == 0)
    return launch_one(gemm_tcgen05_kernel<256, 0, STAGES_BN256, CLUSTER_BN256>, plan->grid, gemm_threads<256>(), gemm_smem_bytes<256, STAGES_BN256>(), CLUSTER_BN256, stream, plan);
  if (plan->bn == 128)
    return launch_one(gemm_tcgen05_kernel<128, 1, STAGES_BN128, 1>, plan->grid, gemm_threads<128>(), gemm_smem_bytes<128, STAGES_BN128>(), 1, stream, plan);
  return launch_one(gemm_tcgen05_kernel<256, 1, STAGES_BN256, CLUSTER_BN256>, plan->grid, gemm_threads<256>(), gemm_smem_bytes<256, STAGES_BN256>(), CLUSTER_BN256, stream, plan);
}

}  // namespace kit

// Test-only: ONE tcgen05.mma sequence issued from operand tiles the caller lays out byte by byte, accumulator read back
// lane by lane.  tests/test_umma_probe_gpu.py uses it to pin the operand-layout facts the attention kernels rely on
// (K-major / MN-major 128B-swizzled tiles, sub-atom start offsets, the M = 64 accumulator lanes, A operand from tensor
// memory) against a numpy matrix product, independent of any kernel that uses them.
#include "attention.cuh"
#include "gemm_sm100.cuh"

namespace kit {

constexpr int PROBE_TILE_BYTES = 65536;

__device__ __forceinline__ uint64_t make_smem_desc_any(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7u) << 61;
  return d;
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void probe_tmem_st32(uint32_t taddr, const uint32_t* r) {
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

struct ProbeArgs {
  uint32_t idesc;
  int32_t steps;                 // tcgen05.mma instructions (K = 16 each), accumulating
  int32_t a_off, a_step, a_lbo, a_sbo, a_layout;   // bytes; start = tile + a_off + step * a_step
  int32_t b_off, b_step, b_lbo, b_sbo, b_layout;
  int32_t a_from_tmem;           // A operand = tensor memory [128 lanes x a_tmem_cols] (32-bit words = bf16 pairs) at column 256
  int32_t a_tmem_cols, a_tmem_step;   // columns; step advance in columns
  int32_t d_lane;                // lane field of the accumulator address
  int32_t n_cols;                // accumulator columns to read back (multiple of 32)
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const uint8_t* __restrict__ a_bytes, int a_len,
                                                            const uint8_t* __restrict__ b_bytes, int b_len, const ProbeArgs pa,
                                                            float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + PROBE_TILE_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + PROBE_TILE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!pa.a_from_tmem)
    for (int i = threadIdx.x * 16; i < a_len; i += 128 * 16) *reinterpret_cast<uint4*>(a_s + i) = *reinterpret_cast<const uint4*>(a_bytes + i);
  for (int i = threadIdx.x * 16; i < b_len; i += 128 * 16) *reinterpret_cast<uint4*>(b_s + i) = *reinterpret_cast<const uint4*>(b_bytes + i);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
  // sentinel in the accumulator region: lanes / columns the MMA does not write stay -12345
  {
    uint32_t s[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] = __float_as_uint(-12345.f);
    for (int c = 0; c < pa.n_cols; c += 32) probe_tmem_st32(tmem_base + lane_base + c, s);
    if (pa.a_from_tmem) {   // row r = this thread: a_tmem_cols 32-bit words at a_bytes + r * a_tmem_cols * 4
      const uint32_t* src = reinterpret_cast<const uint32_t*>(a_bytes) + (size_t)(warp * 32 + lane) * pa.a_tmem_cols;
      for (int c = 0; c < pa.a_tmem_cols; c += 32) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = src[c + i];
        probe_tmem_st32(tmem_base + lane_base + 256 + c, v);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    for (int s = 0; s < pa.steps; ++s) {
      const uint64_t bdesc = make_smem_desc_any(smem_u32(b_s) + pa.b_off + s * pa.b_step, pa.b_lbo, pa.b_sbo, pa.b_layout);
      const uint32_t d = tmem_base + (uint32_t(pa.d_lane) << 16);
      if (pa.a_from_tmem) {
        umma_bf16_ts(d, tmem_base + 256 + s * pa.a_tmem_step, bdesc, pa.idesc, s > 0 ? 1u : 0u);
      } else {
        const uint64_t adesc = make_smem_desc_any(smem_u32(a_s) + pa.a_off + s * pa.a_step, pa.a_lbo, pa.a_sbo, pa.a_layout);
        umma_bf16(d, adesc, bdesc, pa.idesc, s > 0 ? 1u : 0u);
      }
    }
    umma_commit(bar);
  }
  __syncwarp();
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c = 0; c < pa.n_cols; c += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_base + lane_base + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) out[(size_t)(warp * 32 + lane) * pa.n_cols + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// Issue rate of tcgen05.mma for a given shape: `reps` back-to-back instructions on zero-filled operand tiles by one thread,
// clock64 from before the first issue to the completion of the commit.  out[0] = cycles until the last issue returned,
// out[1] = cycles until all had completed.
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(uint32_t idesc, int reps, int a_mn, int b_mn, int a_tmem, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * PROBE_TILE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x * 16; i < 2 * PROBE_TILE_BYTES; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if ((threadIdx.x >> 5) == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x < 32) {   // warp-uniform loop, each instruction under elect.sync: operands stay in uniform registers
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    const bool leader = el != 0;
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + PROBE_TILE_BYTES);
    const uint64_t hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
    const uint64_t ad0 = hi | (uint64_t)(((a_base & 0x3FFFFu) >> 4) | (a_mn ? ((8192u >> 4) << 16) : 0u));
    const uint64_t bd0 = hi | (uint64_t)(((b_base & 0x3FFFFu) >> 4) | (b_mn ? ((8192u >> 4) << 16) : 0u));
    const uint32_t da = a_mn ? 128u : 2u, db = b_mn ? 128u : 2u;
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (leader) {
          if (a_tmem) umma_bf16_ts(tb, tb + 256 + k * 8, bd0 + k * db, idesc, (r + k) > 0 ? 1u : 0u);
          else umma_bf16(tb, ad0 + k * da, bd0 + k * db, idesc, (r + k) > 0 ? 1u : 0u);
        }
      }
    }
    const long long t1 = clock64();
    if (leader) umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (leader) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace kit

using namespace kit;

extern "C" int kit_umma_rate(uint32_t idesc, int32_t reps, int32_t a_mn, int32_t b_mn, int32_t a_tmem, long long* out_dev, void* stream) {
  const int smem = 2 * PROBE_TILE_BYTES + 1024 + 64;
  KIT_CHECK_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_rate_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(idesc, reps, a_mn, b_mn, a_tmem, out_dev);
  KIT_LAUNCH_CHECK();
  return KIT_OK;
}

// args: the 17 int32 fields of ProbeArgs in declaration order (idesc first, as its bit pattern).
extern "C" int kit_umma_probe(const void* a_bytes, int32_t a_len, const void* b_bytes, int32_t b_len, const int32_t* args_host,
                              float* out, void* stream) {
  KIT_REQUIRE(a_bytes && b_bytes && args_host && out, "kit_umma_probe: null argument");
  KIT_REQUIRE(a_len >= 0 && a_len <= PROBE_TILE_BYTES && b_len > 0 && b_len <= PROBE_TILE_BYTES && a_len % 16 == 0 && b_len % 16 == 0,
              "kit_umma_probe: tiles are at most 64 KB, multiples of 16 bytes");
  ProbeArgs pa;
  static_assert(sizeof(ProbeArgs) == 17 * 4, "ProbeArgs is 17 int32");
  memcpy(&pa, args_host, sizeof(pa));
  KIT_REQUIRE(pa.n_cols > 0 && pa.n_cols <= 256 && pa.n_cols % 32 == 0 && pa.steps > 0 && pa.steps <= 64, "kit_umma_probe: bad n_cols / steps");
  KIT_REQUIRE(!pa.a_from_tmem || (pa.a_tmem_cols > 0 && pa.a_tmem_cols <= 256 && pa.a_tmem_cols % 32 == 0), "kit_umma_probe: bad a_tmem_cols");
  const int smem = 2 * PROBE_TILE_BYTES + 1024 + 64;
  KIT_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint8_t*)a_bytes, a_len, (const uint8_t*)b_bytes, b_len, pa, out);
  KIT_LAUNCH_CHECK();
  return KIT_OK;
}

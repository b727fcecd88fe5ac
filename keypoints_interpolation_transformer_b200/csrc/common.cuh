// Shared device/host helpers for the kit_b200 library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kit.h"

namespace kit {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
#define KIT_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      kit::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,            \
                     cudaGetErrorString(_e));                                         \
      return KIT_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)
#define KIT_REQUIRE(cond, ...)                                                        \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      kit::set_error(__VA_ARGS__);                                                    \
      return KIT_ERR_INVALID;                                                         \
    }                                                                                 \
  } while (0)
#define KIT_LAUNCH_CHECK() KIT_CHECK_CUDA(cudaGetLastError())

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of the library is launched with the programmatic-stream-serialization attribute and starts with
// pdl_grid_sync(): it waits until the kernels before it in the stream have completed (their writes are visible), then lets
// the NEXT kernel's CTAs be scheduled, so that kernel's launch latency and prologue (barrier init, tensor-memory allocation,
// descriptor prefetch) overlap this kernel's execution instead of following it.  Nothing before pdl_grid_sync() may touch
// global memory another kernel writes.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_grid_sync() {
  pdl_wait();
  pdl_launch_dependents();
}
bool pdl_enabled();   // errors.cu: KIT_PDL=0 turns the attribute off (A/B measurements)
int sm_reserve();     // errors.cu: SMs the persistent kernels leave free for concurrent collectives (kit_set_sm_reserve)
void set_sm_reserve_active(bool on);   // engine.cu: off while no collective can be in flight (forward, backward before its first bucket)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
// Packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): two lanes of a 64-bit register pair per instruction.
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void up2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float tanh_approx(float x) {
#ifdef KIT_EXP_NO_MUFU
  return fmaf(x, 0.25f, 0.1f);   // timing experiment only
#else
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}
// GELU (activation="gelu" at model.py:87, the erf form 0.5 x (1 + erf(x / sqrt 2))) evaluated as 0.5 x (1 + tanh(u(x))) with the odd
// quintic u = x (c0 + c1 x^2 + c2 x^4) fitted to the erf form in the max norm: |difference| <= 2.6e-5 absolute (the textbook
// cubic "tanh approximation" is 4.8e-4 off), derivative within 1.1e-4 -- both far below the bf16 rounding (2^-9 relative) of the
// activation that is stored -- plus 2^-11 relative from MUFU.TANH.  x^2 is clamped at 49: beyond |x| = 7 the tanh is saturated
// and the negative c2 term must not turn u around.  One MUFU and 3 packed FMA-pipe instructions per element instead of
// rcp + ex2 + ~12 scalar ones for an erf polynomial: the epilogues that apply it are ALU-bound, not tensor-bound.  The derivative
// differentiates the same expression, so forward and backward stay consistent.  kit.h states the deviation.
constexpr float GELU_C0 = 0.797507884f, GELU_C1 = 0.0370056460f, GELU_C2 = -3.51516789e-4f, GELU_SQ_MAX = 49.f;
__device__ __forceinline__ uint64_t gelu_sq_clamped(uint64_t x) {
  float a, b;
  up2(mul2(x, x), a, b);
  return pk2(fminf(a, GELU_SQ_MAX), fminf(b, GELU_SQ_MAX));
}
__device__ __forceinline__ void gelu_pair(float& a, float& b) {
  const uint64_t x = pk2(a, b);
  const uint64_t sq = gelu_sq_clamped(x);
  const uint64_t u = mul2(x, fma2(sq, fma2(sq, pk2(GELU_C2, GELU_C2), pk2(GELU_C1, GELU_C1)), pk2(GELU_C0, GELU_C0)));
  float ua, ub;
  up2(u, ua, ub);
  const uint64_t t = pk2(tanh_approx(ua), tanh_approx(ub));
  const uint64_t hx = mul2(x, pk2(0.5f, 0.5f));
  up2(fma2(hx, t, hx), a, b);
}
// (ga, gb) *= gelu'(za), gelu'(zb)
__device__ __forceinline__ void gelu_grad_mul_pair(float za, float zb, float& ga, float& gb) {
  const uint64_t x = pk2(za, zb);
  const uint64_t sq = gelu_sq_clamped(x);
  const uint64_t u = mul2(x, fma2(sq, fma2(sq, pk2(GELU_C2, GELU_C2), pk2(GELU_C1, GELU_C1)), pk2(GELU_C0, GELU_C0)));
  float ua, ub;
  up2(u, ua, ub);
  const uint64_t t = pk2(tanh_approx(ua), tanh_approx(ub));
  const uint64_t du = fma2(sq, fma2(sq, pk2(5.f * GELU_C2, 5.f * GELU_C2), pk2(3.f * GELU_C1, 3.f * GELU_C1)), pk2(GELU_C0, GELU_C0));   // u'(x)
  const uint64_t sech2 = fma2(mul2(t, pk2(-1.f, -1.f)), t, pk2(1.f, 1.f));                     // 1 - t^2
  const uint64_t hx = mul2(x, pk2(0.5f, 0.5f));
  const uint64_t d = fma2(mul2(hx, sech2), du, fma2(t, pk2(0.5f, 0.5f), pk2(0.5f, 0.5f)));     // 0.5(1+t) + 0.5 x (1-t^2) u'
  up2(mul2(pk2(ga, gb), d), ga, gb);
}
__device__ __forceinline__ float gelu_act(float x) {
  float a = x, b = 0.f;
  gelu_pair(a, b);
  return a;
}
__device__ __forceinline__ float gelu_act_grad(float x) {
  float a = 1.f, b = 0.f;
  gelu_grad_mul_pair(x, 0.f, a, b);
  return a;
}

// 8 bf16 <-> 8 floats through ONE 16-byte access (uint4: a struct of __nv_bfloat162 is copied
// member-wise by nvcc and degenerates into four 4-byte accesses).
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ void load8(const bf16* p, float* f) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16(r.x), b = unpack_bf16(r.y), c = unpack_bf16(r.z), d = unpack_bf16(r.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ void store8(bf16* p, const float* f) { *reinterpret_cast<uint4*>(p) = pack8(f); }
__device__ __forceinline__ void load8(const float* p, float* f) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ---------------------------------------------------------------- PTX: cp.async (LDGSTS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t sz = pred ? 16u : 0u;   // src-size 0: the 16 bytes are zero-filled, the address is not dereferenced
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t sz = pred ? 4u : 0u;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------- PTX: mbarrier / TMA / tcgen05
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (reported as a CUDA error) instead of a hang
// that would take the GPU box down.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("kit: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// Same load, delivered to the same CTA-relative shared-memory offset (and mbarrier) of every CTA in cta_mask.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// The barrier before a CTA of a pair exits (its peer may still read its shared memory / signal its barriers): pure execution
// ordering, no data is handed over, so the arrive is relaxed -- the release form is MEMBAR.ALL.GPU + ERRBAR per thread, which at
// the end of a kernel waits for every outstanding global store of the epilogue (7-10 % of the samples of a K = 256 GEMM).
// The barrier after set-up: what the peer must see is (a) the mbarrier initialisation, published by
// fence.mbarrier_init.release.cluster, and (b) the all-ones operand tile, made visible to the tensor core by this thread's
// fence.proxy.async -- both issued before the arrive, so the arrive itself is relaxed (the release form starts every kernel
// with MEMBAR.ALL.GPU + ERRBAR: ~1.1 k of the 1.7 k set-up cycles of a GEMM launch).
__device__ __forceinline__ void cluster_sync_setup() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_exit() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// smem (generic-proxy writes made visible with fence.proxy.async) -> global tile, clipped at the tensor edge
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// global tile += smem tile (fp32), performed by the TMA unit at L2
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// Same, arriving on the barrier at this CTA-relative offset in every CTA of cta_mask (cluster multicast).
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// ---- cta_group::2: one tcgen05.mma spans the two CTAs of a cluster (256 x N tile: each CTA holds 128 accumulator rows in
// its own tensor memory and stages its own 128 rows of A plus HALF of the B tile) -- cute/arch/mma_sm100_umma.hpp
// SM100_MMA_F16BF16_2x1SM_SS, copy_sm100_tma.hpp SM100_TMA_2SM_LOAD_2D, cutlass/arch/barrier.h umma_arrive_multicast_2x1SM.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> CTA 0 of the pair
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this CTA-relative offset in every CTA of cta_mask once the pair's MMAs issued so far are done
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// tile -> OWN shared memory, completion bytes -> the mbarrier at the same offset in CTA 0 of the pair
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// relaxed remote arrive: the data the barrier guards is ordered by tcgen05 fences (TMEM) or by fence.proxy.async (this CTA's
// shared memory, read by the tensor core) -- a release fence at cluster scope per arrival costs ~1000 cycles per warp per arrival (MEMBAR.ALL.CTA + ERRBAR in the SASS)
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (base_lane + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 in
// [0,14), LBO>>4 in [16,30), SBO>>4 in [32,46), version=1 in [46,48), SWIZZLE_128B=2 in [61,64).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32 (InstrDescriptor in the same header).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace kit

// Thin inline-PTX wrappers for the sm_100a features used by the tensor-core kernels:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), proxy fences.
// Encodings follow the PTX ISA 8.7 and cross-checked against cute/arch/mma_sm100_desc.hpp.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tfgnn {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// (The suspend-time-hint form of try_wait compiles to the SAME SYNCS.PHASECHK.TRANS64.TRYWAIT as the plain one on sm_100a -
// checked in the SASS - so it is not used.)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// Same, with `ns` nanoseconds of nanosleep between probes (ns == 0: plain spin).  A waiting warp re-issues YIELD / TRYWAIT /
// BRA every 15-50 cycles: in the fused kernel the waiting roles (TMA, MMA, splitters, epilogue) issued 30 % of ALL warp
// instructions (ncu source page, gpurun r2n: 418 M probe triples per launch at H = 320), on the same schedulers as the gather
// warps.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns) {
  const uint32_t addr = smem_u32(bar);
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (ns) asm volatile("nanosleep.u32 %0;" ::"r"(ns));
  }
}

// 4x4 transpose of float4 over a lane quad in two shuffle rounds (xor 1, xor 2): every lane comes in with the four float4
// v[0..15] of ITS row and leaves with float4 number (lane & 3) of the quad's four rows (C_k = row k of the quad).  Epilogues
// use it to store 64 contiguous bytes per row and instruction straight from registers.  All 32 lanes must call it.
__device__ __forceinline__ float4 shfl_xor_f4(float4 x, int m) {
  return make_float4(__shfl_xor_sync(0xffffffffu, x.x, m), __shfl_xor_sync(0xffffffffu, x.y, m),
                     __shfl_xor_sync(0xffffffffu, x.z, m), __shfl_xor_sync(0xffffffffu, x.w, m));
}
__device__ __forceinline__ void quad_transpose_f4(const float (&v)[16], int lane, float4& C0, float4& C1, float4& C2,
                                                  float4& C3) {
  const bool b0 = lane & 1, b1 = lane & 2;
  const float4 A0 = make_float4(v[0], v[1], v[2], v[3]), A1 = make_float4(v[4], v[5], v[6], v[7]);
  const float4 A2 = make_float4(v[8], v[9], v[10], v[11]), A3 = make_float4(v[12], v[13], v[14], v[15]);
  // round 1 (pairs): afterwards B0,B1 = float4 number (lane & 1) of the pair's two rows, B2,B3 = number (lane & 1) + 2
  const float4 r0 = shfl_xor_f4(b0 ? A0 : A1, 1), r1 = shfl_xor_f4(b0 ? A2 : A3, 1);
  const float4 B0 = b0 ? r0 : A0, B1 = b0 ? A1 : r0, B2 = b0 ? r1 : A2, B3 = b0 ? A3 : r1;
  // round 2 (pairs of pairs): lanes 0,1 keep B0,B1 and get the other pair's B0,B1; lanes 2,3 keep B2,B3 and get B2,B3
  const float4 s0 = shfl_xor_f4(b1 ? B0 : B2, 2), s1 = shfl_xor_f4(b1 ? B1 : B3, 2);
  C0 = b1 ? s0 : B0; C1 = b1 ? s1 : B1; C2 = b1 ? B2 : s0; C3 = b1 ? B3 : s1;
}

// ---- proxy fences ----------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMA -------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tile load: coordinates (c0 = innermost/contiguous dim, c1 = row).  Completes on `bar` (tx bytes).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// Prefetch a 2-D tile into L2 (no shared-memory destination, no barrier): issued a whole tile ahead by the GEMM's producer so
// that the later cp.async.bulk.tensor finds its lines in L2 instead of paying a loaded-HBM round trip inside the pipeline.
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1)
               : "memory");
}

// 2-D tile load with an L2 cache-policy operand (createpolicy).
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                                 int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "l"(policy)
      : "memory");
}

// ---- L2 eviction policies (126 MB L2: keep the producer/consumer ring, stream everything else) ----
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// Drop a 128 B line from L2 WITHOUT writing it back (producer/consumer scratch that is dead after the read).
__device__ __forceinline__ void discard_l2_128(const void* ptr) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(ptr) : "memory");
}
// LDGSTS, L1 bypass.  (The .L2::cache_hint form of cp.async raised "illegal instruction" on B200 with CUDA 12.9 in
// tools/gather_ceiling.cu, so the gather ring copies without an eviction hint.)
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void sts_f4(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void st_f4_hint(float* ptr, float4 v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "l"(policy)
               : "memory");
}

// ---- tcgen05 / TMEM ----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// TMEM -> registers: this thread's lane (32*(warp%4)+laneid), 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Same without the wait: issue several loads, then ONE tmem_wait_ld() (each wait is a full round trip
// through the tcgen05 pipe, which also carries the queued MMAs).
__device__ __forceinline__ void tmem_ld_x16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ---------------------------------------------------------------------------
// Shared-memory operand descriptor, K-major tile with 128-byte swizzle: rows of 128 B (32 tf32),
// 8-row groups 1024 B apart (SBO), tile base 1024 B aligned.  (cute::UMMA::SmemDescriptor:
// start[0,14) LBO[16,30) SBO[32,46) version[46,48)=1 layout[61,64)=2 SWIZZLE_128B.)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;            // LBO (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO = 1024 B
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// Same for 64-byte rows (16 tf32 per K block): SWIZZLE_64B (layout code 4), 8-row groups 512 B apart.
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;   // SBO = 512 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;            // SWIZZLE_64B
  return d;
}
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t umma_desc_k(uint32_t smem_addr) {
  return ROW_BYTES == 128 ? umma_desc_k_sw128(smem_addr) : umma_desc_k_sw64(smem_addr);
}
// Instruction descriptor kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = n
// (cute::UMMA::InstrDescriptor: c_format[4,6)=1 a_format[7,10)=2 b_format[10,13)=2 n>>3 [17,23) m>>4 [24,29)).
__host__ __device__ __forceinline__ uint32_t umma_idesc_tf32_m128(uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// ---- thread-block clusters / CTA pairs (cta_group::2) -----------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {   // arrive on another CTA's mbarrier
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Cluster-scope release / acquire: the arriving thread's earlier writes (shared OR global) become visible to a thread of
// ANOTHER CTA of the cluster that observes the phase with mbar_wait_cluster (split-tile mode of the fused kernel).
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// TMEM allocation for a CTA pair: the same warp index of BOTH CTAs executes it with the same smem slot offset.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B over the CTA pair: M = 256 (128 rows of A in each CTA's smem), B's N rows split
// half/half between the two CTAs' smem at the same offsets.  Issued by ONE thread of the leader (rank 0) CTA.
__device__ __forceinline__ void mma_tf32_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of the pair's MMAs: arrives on the mbarrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
__host__ __device__ __forceinline__ uint32_t umma_idesc_tf32(uint32_t m, uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---- bf16-pair correction operands (round 2) -----------------------------------------------------------------
// 3xTF32 needs A_hi B_hi + (A_lo B_hi + A_hi B_lo).  The two small products only need ~8 bits each, so they run as ONE
// kind::f16 (bf16) MMA over a K-interleaved operand: the 32-bit word at the position of fp32 element k holds the bf16
// pair (A: bf16(a_k) | bf16(a_k - tf32(a_k)),  B: bf16(b_k - tf32(b_k)) | bf16(b_k)), low half first.  A row of 32 fp32
// words is then a row of 64 bf16 K-elements in the SAME bytes, swizzle and descriptors as the fp32 tile, and
//   sum_j A'_j B'_j = sum_k a_k lo(b_k) + lo(a_k) b_k.
// Tensor work per K block: 4 tf32 + 4 bf16 instructions instead of 12 tf32 (-33 %); the splitter writes one tile, not two.
__device__ __forceinline__ uint32_t pack_bf16x2(float upper, float lower) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
  return d;
}
// Instruction descriptor kind::f16 with bf16 operands, fp32 accumulate, A and B K-major
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// tf32 split of an fp32 value: hi keeps sign/exponent/10 mantissa bits (exact truncation, so the
// tensor core sees the same bits whether it truncates or rounds), lo = x - hi (exact in fp32).
__host__ __device__ __forceinline__ float tf32_hi(float x) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
#else
  union { float f; uint32_t u; } c; c.f = x; c.u &= 0xFFFFE000u; return c.f;
#endif
}

}  // namespace ptx
}  // namespace tfgnn

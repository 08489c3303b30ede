// The fused RGCN-style layer: ONE persistent kernel does gather(h[src]) -> segment-sum -> 1/(c+eps)
// -> 3xTF32 tcgen05 contraction with [W_0;..;W_{L-1}] -> row-norm / activation -> out.
//
//   out[v] = act( rn(v) * sum_l ( 1/(c_{v,l}+eps) * sum_{(u,v) in A_l} h_u ) W_l )       (rgcn.py:13-48)
//
// One CTA per SM owns 128-target tiles.  Inside the CTA, 16 GATHER warps produce, per edge type l, the
// normalised row sums A_l[128, D] (full 4*D-byte row reads from HBM, register accumulation, no atomics)
// into a small per-CTA ring in global memory that stays L2-resident (3 slots x 128 x D x 4 B per CTA,
// 57 MB chip-wide at D=256: the fp32-accurate operand does not fit the 227 KB of shared memory).  The
// TMA producer pulls each slot back in 128 B K-slices next to the pre-split weight tiles; splitter warps
// cut A into tf32 hi/lo; one thread issues the tcgen05 MMAs into TMEM (main + correction accumulators,
// double-buffered); epilogue warps drain TMEM through a staging tile to coalesced stores.  The gather
// of slot i+1..i+3 overlaps the MMAs of slot i, so the kernel runs at the HBM rate of the gather and the
// [V, L*D] intermediate never touches HBM.
//
// Warp roles (896 threads, by warpgroup so that setmaxnreg can move registers to the gather warps):
//   WG0: 0 TMA, 1 MMA, 2-3 idle | WG1 (4-7): A splitters | WG2 (8-11): epilogue | WG3-6 (12-27): gather.
// The gather is latency-bound on register-held loads (72 registers/thread cap at 896 threads), so each
// gather warp also runs a rolling L2 prefetch window ahead of its loads: DRAM requests in flight are then
// bounded by the memory system, not by registers.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "gemm.cuh"
#include "sm100_ptx.cuh"

namespace tfgnn {

constexpr int kFuBM = 128;
constexpr int kFuGatherWarps = 16;
constexpr int kFuFirstGatherWarp = 12;
constexpr int kFuThreads = 32 * (kFuFirstGatherWarp + kFuGatherWarps);
constexpr int kFuMaxSlots = 8;                  // ring slots per CTA: 3, or L+1 when the N dimension needs two passes
constexpr int kFuTmemCols = 512;
constexpr int kFuAccStride = 256;
constexpr int kFuSmemLimit = 227 * 1024;
constexpr int kFuEpiPitch = 36;
constexpr int kFuEpiBytes = 4 * 32 * kFuEpiPitch * 4;

struct FusedParams {
  // graph
  const float* h;
  int ldh;
  const int* row_ptr;
  const int* src;
  int V, L, D;
  int normalize;
  int discard_ring;     // discard.global.L2 on consumed ring slots
  int prefetch_warp;    // 1: a dedicated warp bulk-prefetches source rows into L2 a few row-waves ahead of the gather
  int debug_skip;       // timing experiments only (results invalid), bit mask: 1 = no edge gathers, 2 = one K block per slot, 4 = no epilogue work
  int prefetch_window;  // edges whose rows are L2-prefetched ahead of the register loads (0 = off)
  // ring
  float* ring;  // [grid * num_slots * 128, D]
  int num_slots;
  // GEMM
  int N, block_n, n_tiles;
  long long m_tiles;
  int kb_per_type, num_stages;
  float* C;
  int ldc;
  GemmEpilogue epi;
};

__device__ __forceinline__ float fu_row_norm(const GemmEpilogue& e, long long row) {
  if (e.row_norm == 0) return 1.0f;
  int cnt = 0;
  for (int l = 0; l < e.L; ++l) {
    const long long s = (long long)l * e.V + e.row0 + row;
    cnt += __ldg(e.row_ptr + s + 1) - __ldg(e.row_ptr + s);
  }
  const float c = (float)max(cnt, 1);
  return e.row_norm == 1 ? c : sqrtf(c);
}

// One warp reduces `nrows` CONSECUTIVE targets of edge type l: their CSR segments are contiguous, so the
// warp loads the source ids of all of them with coalesced loads and walks the edges in one flat loop with U
// row loads in flight per lane (the dependent row_ptr -> index chain is paid once per 8 rows, not per row).
// h rows are read with L2 evict_first (each row is used once per incoming edge, no reuse window); the ring
// rows are written with L2 evict_last so that they are still resident when the TMA reads them back.
template <int NV, int U>
__device__ __forceinline__ void gather_rows_batch(const FusedParams& p, int l, int v0, int nrows, int lane,
                                                  float* dst, uint64_t pol_stream, uint64_t pol_keep,
                                                  volatile uint32_t* progress) {
  const int C4 = p.D >> 2;
  const int rp = lane <= nrows ? __ldg(p.row_ptr + (long long)l * p.V + v0 + lane) : 0;
  const int e_begin = __shfl_sync(0xffffffffu, rp, 0);
  const int e_end = (p.debug_skip & 1) ? e_begin : __shfl_sync(0xffffffffu, rp, nrows);
  int row = 0;
  int seg_end = __shfl_sync(0xffffffffu, rp, 1);
  float4 acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto flush = [&](int r) {
    const int cnt = __shfl_sync(0xffffffffu, rp, r + 1) - __shfl_sync(0xffffffffu, rp, r);
    const float scale = p.normalize ? 1.0f / ((float)cnt + kSmallNumber) : 1.0f;
    float* d = dst + (size_t)r * p.D;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < C4)
        ptx::st_f4_hint(d + 4 * c4,
                        make_float4(acc[j].x * scale, acc[j].y * scale, acc[j].z * scale, acc[j].w * scale),
                        pol_keep);
      acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (lane == 0) atomicAdd(const_cast<uint32_t*>(progress), 1u);
  };

  for (int base = e_begin; base < e_end; base += 32) {
    const int n = min(32, e_end - base);
    const int my_src = lane < n ? __ldg(p.src + base + lane) : 0;
    const uint32_t row_bytes = (uint32_t)p.D * 4;
    // prime the L2 prefetch window: ONE bulk request per source row (cp.async.bulk.prefetch.L2)
    if (lane < min(n, p.prefetch_window)) ptx::bulk_prefetch_l2(p.h + (long long)my_src * p.ldh, row_bytes);
    for (int j0 = 0; j0 < n; j0 += U) {
      {
        // roll the window: rows of edges [j0 + W, j0 + W + U)
        const int w0 = j0 + p.prefetch_window;
        if (p.prefetch_window > 0 && lane >= w0 && lane < min(n, w0 + U))
          ptx::bulk_prefetch_l2(p.h + (long long)my_src * p.ldh, row_bytes);
      }
      float4 r[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = __shfl_sync(0xffffffffu, my_src, (j0 + u) & 31);
        const float* rowp = p.h + (long long)s * p.ldh;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c4 = lane + 32 * j;
          r[u][j] = (j0 + u < n && c4 < C4) ? ptx::ld_nc_f4_hint(rowp + 4 * c4, pol_stream)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j0 + u < n) {
          const int e = base + j0 + u;
          while (e >= seg_end) {   // warp-uniform: close finished (possibly empty) segments
            flush(row);
            ++row;
            seg_end = __shfl_sync(0xffffffffu, rp, row + 1);
          }
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            acc[j].x += r[u][j].x; acc[j].y += r[u][j].y; acc[j].z += r[u][j].z; acc[j].w += r[u][j].w;
          }
        }
      }
    }
  }
  while (row < nrows) {
    flush(row);
    ++row;
  }
}

// BK = floats per K block: 32 (128 B rows, SWIZZLE_128B) or 16 (64 B rows, SWIZZLE_64B).  The smaller block
// halves the bytes per pipeline stage, so twice as many stages fit: under the gather's L2 traffic a TMA round
// trip takes ~3 us, and it is bytes-in-flight / latency that bounds the operand feed of the tensor core.
template <int NV, int BK>
__global__ void __launch_bounds__(kFuThreads, 1)
fused_rgcn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kFuBK = BK;
  constexpr int kRowBytes = BK * 4;
  constexpr int kFuATileBytes = kFuBM * kRowBytes;
  const int S = p.num_stages;
  const int b_tile_bytes = p.block_n * kRowBytes;
  const int stage_bytes = 2 * kFuATileBytes + 2 * b_tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
  uint64_t* full = bars;
  uint64_t* split = bars + S;
  uint64_t* empty = bars + 2 * S;
  uint64_t* tmem_full = bars + 3 * S;
  uint64_t* tmem_empty = bars + 3 * S + 2;
  uint64_t* slot_ready = bars + 3 * S + 4;
  uint64_t* slot_free = bars + 3 * S + 4 + kFuMaxSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4 + 2 * kFuMaxSlots);
  volatile uint32_t* progress = tmem_slot + 1;   // rows finished by the gather warps (paces the prefetch warp)
  float* epi_stage = reinterpret_cast<float*>(bars + ((3 * S + 4 + 2 * kFuMaxSlots + 2 + 1) & ~1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total_tiles = p.m_tiles;      // one tile = 128 targets; the N dimension is covered in n_pass passes
  const int n_pass = p.n_tiles;
  const int kFuSlots = p.num_slots;
  const int kb_per_tile = p.L * p.kb_per_type;

  if (warp == 0 && lane == 0) {
    tmem_slot[1] = 0;
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&split[s], 128);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 128);
    }
    for (int r = 0; r < kFuMaxSlots; ++r) {
      ptx::mbar_init(&slot_ready[r], kFuGatherWarps);
      ptx::mbar_init(&slot_free[r], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, kFuTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t n_acc = p.block_n <= 128 ? 2 : 1;
  const uint32_t corr_off = p.block_n <= 128 ? 128 : 256;
  const int ring_row0 = blockIdx.x * kFuSlots * kFuBM;  // first ring row of this CTA

  // N passes: main + correction accumulators need 2*block_n <= 512 TMEM columns, so H in (256, 512] is covered
  // in two passes of block_n = H/2 columns over the SAME gathered ring slots (the ring then holds all L types of
  // the tile: num_slots = L + 1); every source row is still gathered exactly once.
  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0, slot_it = 0;
      const uint64_t pol_keep = ptx::policy_evict_last();   // ring slots and the 2 MB of weights stay in L2
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, slot_it += p.L) {
       for (int pass = 0; pass < n_pass; ++pass) {
        const int n0 = pass * p.block_n;
        for (int l = 0; l < p.L; ++l) {
          const uint32_t sq = slot_it + l;
          const int slot = sq % kFuSlots;
          ptx::mbar_wait(&slot_ready[slot], (sq / kFuSlots) & 1);
          for (int kb = 0; kb < p.kb_per_type; ++kb, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            ptx::mbar_wait(&empty[s], ph ^ 1);
            uint8_t* st = smem + (size_t)s * stage_bytes;
            ptx::mbar_arrive_expect_tx(&full[s], kFuATileBytes + 2 * b_tile_bytes);
            ptx::tma_load_2d_hint(st, &map_a, &full[s], kb * kFuBK, ring_row0 + slot * kFuBM, pol_keep);
            const int kcol = (l * p.kb_per_type + kb) * kFuBK;
            ptx::tma_load_2d_hint(st + 2 * kFuATileBytes, &map_b, &full[s], kcol, n0, pol_keep);
            ptx::tma_load_2d_hint(st + 2 * kFuATileBytes + b_tile_bytes, &map_b, &full[s], kcol, p.N + n0, pol_keep);
          }
        }
       }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = ptx::umma_idesc_tf32_m128((uint32_t)p.block_n);
    uint32_t it = 0, tile_count = 0;
    for (long long tp = (long long)blockIdx.x * n_pass; tp < total_tiles * n_pass;
         tp = (tp % n_pass == n_pass - 1) ? tp + (long long)(gridDim.x - 1) * n_pass + 1 : tp + 1, ++tile_count) {
      const uint32_t acc = tile_count % n_acc, acc_ph = (tile_count / n_acc) & 1;
      ptx::mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * kFuAccStride;
      const uint32_t c_tmem = d_tmem + corr_off;
      for (int kb = 0; kb < kb_per_tile; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::mbar_wait(&split[s], ph);
        ptx::tc_fence_after_sync();
        if (lane == 0) {
          const uint32_t st = ptx::smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t a_hi = ptx::umma_desc_k<kRowBytes>(st);
          const uint64_t a_lo = ptx::umma_desc_k<kRowBytes>(st + kFuATileBytes);
          const uint64_t b_hi = ptx::umma_desc_k<kRowBytes>(st + 2 * kFuATileBytes);
          const uint64_t b_lo = ptx::umma_desc_k<kRowBytes>(st + 2 * kFuATileBytes + b_tile_bytes);
#pragma unroll
          for (int k = 0; k < kFuBK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            ptx::mma_tf32_ss(c_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
            ptx::mma_tf32_ss(c_tmem, a_hi + adv, b_lo + adv, idesc, 1);
            ptx::mma_tf32_ss(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
          }
          ptx::mma_commit(&empty[s]);
          if (kb == kb_per_tile - 1) ptx::mma_commit(&tmem_full[acc]);
        }
        __syncwarp();
      }
    }
  } else if ((warp == 2 || warp == 3) && p.prefetch_warp) {
    // ================= L2 prefetch warps (the two otherwise idle warps of warpgroup 0) =================
    // They walk the same (tile, edge type) sequence as the gather warps, a few "row waves" ahead (wave w = row w
    // of each of the 16 gather warps' 8-row groups), and issue one bulk L2 prefetch per source row.  DRAM latency
    // is then paid by requests that hold no registers; the gather's own loads mostly hit L2.
    constexpr int kRowsPerWarpG = kFuBM / kFuGatherWarps;       // 8 rows per gather warp = 8 waves per slot
    constexpr int kLeadRows = 3 * kFuGatherWarps;              // stay <= ~3 waves ahead of the finished rows
    const uint32_t row_bytes = (uint32_t)p.D * 4;
    const int pw = warp - 2;
    uint32_t rows_base = 0;                                    // valid rows of all previous (tile, type) units
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (int)(tile * kFuBM);
      const int valid_rows = min(kFuBM, p.V - m0);
      auto wave_rows = [&](int w) {                             // valid rows in wave w: groups g with g*8 + w < valid
        const int c = (valid_rows - w + kRowsPerWarpG - 1) / kRowsPerWarpG;
        return c < 0 ? 0 : (c > kFuGatherWarps ? kFuGatherWarps : c);
      };
      for (int l = 0; l < p.L; ++l, rows_base += (uint32_t)valid_rows) {
        for (int w0 = 2 * pw; w0 < kRowsPerWarpG; w0 += 4) {   // two waves (32 rows) per iteration, one row per lane
          uint32_t before = rows_base;
          for (int w = 0; w < w0; ++w) before += (uint32_t)wave_rows(w);
          // pace against the rows the gather warps have finished (bounded: it is only a heuristic)
          for (int spin = 0; spin < 100000 && (int)(before - *progress) > kLeadRows; ++spin) __nanosleep(100);
          const int g = lane & 15, w = w0 + (lane >> 4);
          const int r = g * kRowsPerWarpG + w;
          int beg = 0, end = 0;
          if (r < valid_rows) {
            const long long seg = (long long)l * p.V + m0 + r;
            beg = __ldg(p.row_ptr + seg);
            end = __ldg(p.row_ptr + seg + 1);
          }
          for (int e = beg; e < end; e += 8) {
            int ids[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) ids[q] = e + q < end ? __ldg(p.src + e + q) : -1;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (ids[q] >= 0) ptx::bulk_prefetch_l2(p.h + (long long)ids[q] * p.ldh, row_bytes);
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================= A splitters =================
    const int tid = threadIdx.x - 128;
    uint32_t it = 0, slot_base_it = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, slot_base_it += p.L) {
     for (int pass = 0; pass < n_pass; ++pass) {
      for (int l = 0; l < p.L; ++l) {
        const uint32_t slot_it = slot_base_it + l;
        for (int kb = 0; kb < p.kb_per_type; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          ptx::mbar_wait(&full[s], ph);
          if (kb == p.kb_per_type - 1 && pass == n_pass - 1) {
            // all TMA reads of the slot have landed: its lines are dead.  Discard them from L2 so that they are
            // never written back to HBM (the ring is pure on-chip hand-off), then hand the slot back.
            if (p.discard_ring) {
              const char* sb = reinterpret_cast<const char*>(p.ring + ((size_t)ring_row0 + (size_t)(slot_it % kFuSlots) * kFuBM) * p.D);
              const int lines = kFuBM * p.D * 4 / 128;
              for (int i = tid; i < lines; i += 128) ptx::discard_l2_128(sb + (size_t)i * 128);
            }
            ptx::mbar_arrive(&slot_free[slot_it % kFuSlots]);
          }
          float4* a = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
          float4* lo = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + kFuATileBytes);
#pragma unroll 2
          for (int i = 0; i < kFuATileBytes / 16 / 128; ++i) {
            const int idx = tid + i * 128;
            const float4 x = a[idx];
            float4 hh, ll;
            hh.x = ptx::tf32_hi(x.x); hh.y = ptx::tf32_hi(x.y); hh.z = ptx::tf32_hi(x.z); hh.w = ptx::tf32_hi(x.w);
            ll.x = ptx::tf32_hi(x.x - hh.x); ll.y = ptx::tf32_hi(x.y - hh.y);
            ll.z = ptx::tf32_hi(x.z - hh.z); ll.w = ptx::tf32_hi(x.w - hh.w);
            a[idx] = hh;
            lo[idx] = ll;
          }
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(&split[s]);
        }
      }
     }
    }
  } else if (warp >= 8 && warp < kFuFirstGatherWarp) {
    // ================= epilogue (warps 8..11 -> TMEM lane quarters 0..3) =================
    const int q = warp & 3;
    uint32_t tile_count = 0;
    const uint64_t pol_stream = ptx::policy_evict_first();
    for (long long tp = (long long)blockIdx.x * n_pass; tp < total_tiles * n_pass;
         tp = (tp % n_pass == n_pass - 1) ? tp + (long long)(gridDim.x - 1) * n_pass + 1 : tp + 1, ++tile_count) {
      const long long m0 = (tp / n_pass) * kFuBM;
      const int n0 = (int)(tp % n_pass) * p.block_n;
      const uint32_t acc = tile_count % n_acc, acc_ph = (tile_count / n_acc) & 1;
      ptx::mbar_wait(&tmem_full[acc], acc_ph);
      ptx::tc_fence_after_sync();
      const long long row = m0 + q * 32 + lane;
      const bool row_ok = row < p.V;
      const float inv_rn = row_ok ? 1.0f / fu_row_norm(p.epi, row) : 1.0f;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kFuAccStride;
      float* stage = epi_stage + (size_t)(warp - 8) * 32 * kFuEpiPitch;
      for (int c0 = 0; c0 < ((p.debug_skip & 4) ? 0 : p.block_n); c0 += 16) {   // 16 columns per TMEM round trip
        const int ncols = 16;
        {
          // TMEM -> registers: main and correction accumulators of up to 32 columns, ONE wait
          uint32_t mv[2][16], cv[2][16];
          const int nh = ncols > 16 ? 2 : 1;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (half < nh) {
              ptx::tmem_ld_x16_nowait(taddr + c0 + half * 16, mv[half]);
              ptx::tmem_ld_x16_nowait(taddr + corr_off + c0 + half * 16, cv[half]);
            }
          }
          ptx::tmem_wait_ld();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (half < nh) {
              float v[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(mv[half][j]) + __uint_as_float(cv[half][j]);
              // uniform branches hoisted out of the element loops (predicated-off code still costs issue slots
              // and instruction-cache space: the epilogue was ~48 us per 128x256 tile before)
              if (p.epi.row_norm) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] *= inv_rn;
              }
              if (p.epi.bias) {
                const float4* bp = reinterpret_cast<const float4*>(p.epi.bias + n0 + c0 + half * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 bb = __ldg(bp + j);
                  v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
                }
              }
              apply_act_vec<16>(v, p.epi.act);
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(stage + lane * kFuEpiPitch + half * 16 + j) =
                    make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          }
        }
        __syncwarp();
        const int f4_per_row = ncols / 4;
        const int rows_per_it = 32 / f4_per_row;
        const int rr = lane / f4_per_row, cc = (lane % f4_per_row) * 4;
        for (int r0 = 0; r0 < 32; r0 += rows_per_it) {
          const int r = r0 + rr;
          const long long grow = m0 + q * 32 + r;
          if (grow < p.V) {
            const float4 val = *reinterpret_cast<const float4*>(stage + r * kFuEpiPitch + cc);
            ptx::st_f4_hint(p.C + grow * p.ldc + n0 + c0 + cc, val, pol_stream);
          }
        }
        __syncwarp();
      }
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(&tmem_empty[acc]);
    }
  } else if (warp >= kFuFirstGatherWarp) {
    // ================= gather warps =================
    const int gw = warp - kFuFirstGatherWarp;
    constexpr int kRowsPerWarp = kFuBM / kFuGatherWarps;
    constexpr int U = NV <= 1 ? 8 : (NV == 2 ? 4 : 2);
    const uint64_t pol_stream = ptx::policy_evict_first();
    const uint64_t pol_keep = ptx::policy_evict_last();
    uint32_t slot_it = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (int)(tile * kFuBM);
      const int v0 = m0 + gw * kRowsPerWarp;
      const int nrows = min(kRowsPerWarp, p.V - v0);   // <= 0 for warps past the last node
      for (int l = 0; l < p.L; ++l, ++slot_it) {
        const int slot = slot_it % kFuSlots;
        ptx::mbar_wait(&slot_free[slot], ((slot_it / kFuSlots) & 1) ^ 1);
        float* slot_base = p.ring + ((size_t)ring_row0 + (size_t)slot * kFuBM) * p.D;
        if (nrows > 0)
          gather_rows_batch<NV, U>(p, l, v0, nrows, lane, slot_base + (size_t)(gw * kRowsPerWarp) * p.D, pol_stream,
                                   pol_keep, progress);
        // generic-proxy global writes -> visible to the TMA (async proxy) reads of this CTA
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&slot_ready[slot]);
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kFuTmemCols);
  }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn fu_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

bool fused_rgcn_supported(long long V, int L, int D, int H, const float* h, const float* out, int ldo) {
  if (V < 1 || L < 1 || D % 32 != 0 || D > 512 || H % 16 != 0 || H < 16 || H > 512) return false;
  if (H > 256 && ((H / 2) % 16 != 0 || L + 1 > kFuMaxSlots)) return false;   // two N passes: ring holds L+1 slots
  if ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(out)) & 15) return false;
  if (ldo % 4 != 0) return false;
  // every source row is gathered exactly once; H <= 256 fits one accumulator pair (main + correction) in TMEM,
  // H <= 512 takes two passes over the ring.
  return gemm_tc_supported(V, H, L * D, h, D, out, ldo) && fu_encode_fn() != nullptr;
}

constexpr int kFuMaxGrid = 160;
static int fused_num_slots(int L, int H) {
  static const int env_slots = [] { const char* e = getenv("TFGNN_B200_RING_SLOTS"); return e ? atoi(e) : 0; }();
  if (H > 256) return L + 1;
  return (env_slots >= 2 && env_slots <= kFuMaxSlots) ? env_slots : 3;
}
size_t fused_rgcn_ring_bytes(int D, int L, int H) {
  return (size_t)kFuMaxGrid * fused_num_slots(L, H) * kFuBM * D * sizeof(float);
}

int launch_fused_rgcn(const float* h, int D, const int* row_ptr, const int* src, int V, int L, int normalize,
                      const float* packedB, int H, float* ring, float* out, int ldo, const GemmEpilogue& epi,
                      cudaStream_t st) {
  EncodeTiledFn encode = fu_encode_fn();
  if (!encode) {
    set_error(TFGNN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    return TFGNN_ERR_CUDA;
  }
  int dev = 0, sms = 148;
  TFGNN_CUDA(cudaGetDevice(&dev));
  TFGNN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  FusedParams p{};
  p.h = h; p.ldh = D; p.row_ptr = row_ptr; p.src = src; p.V = V; p.L = L; p.D = D; p.normalize = normalize;
  p.ring = ring;
  static const int pf_window = [] { const char* e = getenv("TFGNN_B200_PREFETCH_WINDOW"); return e ? atoi(e) : 32; }();
  p.prefetch_window = pf_window < 0 ? 0 : (pf_window > 32 ? 32 : pf_window);
  static const int discard_env = [] { const char* e = getenv("TFGNN_B200_RING_DISCARD"); return e ? atoi(e) : 1; }();
  p.discard_ring = discard_env;
  static const int pfw_env = [] { const char* e = getenv("TFGNN_B200_PREFETCH_WARP"); return e ? atoi(e) : 0; }();  // measured: no gain over the in-gather window
  p.prefetch_warp = pfw_env;
  if (p.prefetch_warp) p.prefetch_window = 0;   // the dedicated warp replaces the in-gather prefetch window
  static const int dbg_env = [] { const char* e = getenv("TFGNN_B200_DEBUG_SKIP"); return e ? atoi(e) : 0; }();
  p.debug_skip = dbg_env;
  p.N = H;
  p.n_tiles = H > 256 ? 2 : 1;            // N passes
  p.block_n = H / p.n_tiles;
  p.num_slots = fused_num_slots(L, H);
  p.m_tiles = ((long long)V + kFuBM - 1) / kFuBM;
  static const int bk_env = [] { const char* e = getenv("TFGNN_B200_FUSED_BK"); return e ? atoi(e) : 16; }();
  const int kFuBK = bk_env == 32 ? 32 : 16;
  const int kFuATileBytes = kFuBM * kFuBK * 4;
  p.kb_per_type = D / kFuBK;
  if (p.debug_skip & 2) p.kb_per_type = 1;
  const int stage_bytes = 2 * kFuATileBytes + 2 * p.block_n * kFuBK * 4;
  int stages = (kFuSmemLimit - 2048 - kFuEpiBytes) / stage_bytes;
  if (stages > 6) stages = 6;
  TFGNN_REQUIRE(stages >= 2, "fused RGCN: tile does not fit shared memory");
  p.num_stages = stages;
  p.C = out; p.ldc = ldo; p.epi = epi;
  if (sms > kFuMaxGrid) sms = kFuMaxGrid;
  const int grid = (int)(p.m_tiles < sms ? p.m_tiles : sms);

  const int Kp = L * D;
  CUtensorMap map_a, map_b;
  {
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)grid * p.num_slots * kFuBM};
    cuuint64_t strides[1] = {(cuuint64_t)D * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kFuBK, (cuuint32_t)kFuBM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ring, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, kFuBK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error(TFGNN_ERR_CUDA, "cuTensorMapEncodeTiled(ring) failed with code " + std::to_string((int)r));
      return TFGNN_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)(2 * H)};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kFuBK, (cuuint32_t)p.block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(packedB), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, kFuBK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error(TFGNN_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed with code " + std::to_string((int)r));
      return TFGNN_ERR_CUDA;
    }
  }
  const size_t smem_bytes = (size_t)stages * stage_bytes + (3 * stages + 4 + 2 * kFuMaxSlots + 4) * sizeof(uint64_t) +
                            kFuEpiBytes + 1024;
  const int nv = (D + 127) / 128;
  // L2 set-aside for the evict_last (persisting) lines: the ring + the packed weights.  Without a carve-out
  // the evict_last hint is advisory only and the ring gets written back to HBM (measured: +4 GB/layer).
  static std::once_flag l2_once;
  std::call_once(l2_once, [&] {
    int dev2 = 0, max_persist = 0;
    if (cudaGetDevice(&dev2) == cudaSuccess &&
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev2) == cudaSuccess &&
        max_persist > 0) {
      size_t want = (size_t)72 << 20;
      if (const char* e = getenv("TFGNN_B200_L2_PERSIST_MB")) want = (size_t)atoi(e) << 20;
      if (want > (size_t)max_persist) want = (size_t)max_persist;
      if (want > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
    }
    cudaGetLastError();
  });
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    cudaError_t e[8] = {
        cudaFuncSetAttribute(fused_rgcn_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemLimit),
        cudaFuncSetAttribute(fused_rgcn_kernel<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemLimit),
        cudaFuncSetAttribute(fused_rgcn_kernel<3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemLimit),
        cudaFuncSetAttribute(fused_rgcn_kernel<4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemLimit),
        cudaFuncSetAttribute(fused_rgcn_kernel<1, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemLimit),
        cudaFuncSetAttribute(fused_rgcn_kernel<2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemLimit),
        cudaFuncSetAttribute(fused_rgcn_kernel<3, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemLimit),
        cudaFuncSetAttribute(fused_rgcn_kernel<4, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemLimit)};
    for (cudaError_t x : e)
      if (x != cudaSuccess) attr_err = x;
  });
  TFGNN_CUDA(attr_err);
#define TFGNN_FU_LAUNCH(NV, BK) fused_rgcn_kernel<NV, BK><<<grid, kFuThreads, smem_bytes, st>>>(map_a, map_b, p)
  if (kFuBK == 32) {
    switch (nv) {
      case 1: TFGNN_FU_LAUNCH(1, 32); break;
      case 2: TFGNN_FU_LAUNCH(2, 32); break;
      case 3: TFGNN_FU_LAUNCH(3, 32); break;
      default: TFGNN_FU_LAUNCH(4, 32); break;
    }
  } else {
    switch (nv) {
      case 1: TFGNN_FU_LAUNCH(1, 16); break;
      case 2: TFGNN_FU_LAUNCH(2, 16); break;
      case 3: TFGNN_FU_LAUNCH(3, 16); break;
      default: TFGNN_FU_LAUNCH(4, 16); break;
    }
  }
#undef TFGNN_FU_LAUNCH
  TFGNN_LAUNCH_CHECK();
  return 0;
}

}  // namespace tfgnn

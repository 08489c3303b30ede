// The fused RGCN-style layer: ONE persistent kernel does gather(h[src]) -> segment-sum -> 1/(c+eps)
// -> 3xTF32 tcgen05 contraction with [W_0;..;W_{L-1}] -> row-norm / activation -> out.
//
//   out[v] = act( rn(v) * sum_l ( 1/(c_{v,l}+eps) * sum_{(u,v) in A_l} h_u ) W_l )       (rgcn.py:13-48)
//
// One CTA per SM owns 128-target tiles; the two CTAs of a cluster (one TPC) form a tcgen05 CTA pair.  Inside the
// CTA, 16 GATHER warps produce, per edge type l, the normalised row sums A_l[128, D] (full 4*D-byte row reads
// from HBM through a rolling cp.async ring in shared memory, register accumulation, no atomics) into a small
// per-CTA ring in global memory that stays L2-resident (4 slots x 128 x D x 4 B per CTA, 76 MB chip-wide at
// D=256: the fp32-accurate operand does not fit the 227 KB of shared memory).  The TMA producer pulls each slot
// back in 128 B K-slices next to this CTA's HALF of the pre-split weight tile; splitter warps cut A into tf32
// hi/lo; the leader CTA's MMA thread issues cta_group::2 MMAs of M = 256 into both CTAs' TMEM (main + correction
// accumulators); epilogue warps drain TMEM through a staging tile to coalesced stores.  The gather of the next
// slots overlaps the MMAs of slot i, and the [V, L*D] intermediate never touches HBM.
//
// Warp roles (896 threads):
//   0 TMA, 1 MMA, 2-3 idle | 4-7: A splitters | 8-11: epilogue | 12-27: gather.
// The gather warps hold no row data in registers while it is in flight: each keeps a rolling ring of cp.async
// copies into its own shared-memory slots (see GatherIssue), so DRAM requests in flight are bounded by the
// shared memory left over by the GEMM pipeline (64 KB for 16 warps x 4 rows at D = 256), not by registers.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

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
constexpr int kFuMaxQ = 8;                      // bulk row copies in flight per gather warp (upper bound)

struct FusedParams {
  // graph
  const float* h;
  int ldh;
  const int* row_ptr;
  const int* src;
  int V, L, D;
  int normalize;
  long long M;          // entries of src (bound for the 32-wide index block loads)
  int discard_ring;     // discard.global.L2 on consumed ring slots
  int gather_q;         // bulk row copies each gather warp keeps in flight (= its shared-memory row slots)
  int corr_bf16;        // 1: corrections as one bf16-pair MMA (sm100_ptx.cuh), 0: two tf32 MMAs
  int debug_skip;       // timing experiments only (results invalid), bit mask: 1 = no edge gathers, 2 = one K block per slot, 4 = no epilogue work, 8 = no weight-tile loads, 16 = no ring stores, 32 = raw fp32 tile as the hi operand (valid iff the MMA truncates)
  // ring
  float* ring;  // [grid * num_slots * 128, D]
  int num_slots;
  // GEMM
  int N, block_n, n_tiles;
  long long m_tiles;
  int kb_per_type, num_stages;
  float* C;
  int ldc;
  // fused all-gather (tfgnn_b200_rgcn_fwd_allgather): peer copies of the output table, NVLink-mapped; same row indexing as C
  float* C_peer[TFGNN_MAX_PEERS];
  int n_peer;
  float* C_mc;   // multicast mapping of all replicas (NVSwitch replicates one multimem.st), or null
  uint32_t sleep_crit, sleep_long;   // nanosleep between barrier probes (0 = spin): K-loop waits / once-per-tile waits
  int epi_direct;    // epilogue stores straight from registers: 1 = each thread its own row's 16 B pieces, 2 = lane pairs write
                     // 32 contiguous bytes, 3 = lane quads write 64 contiguous bytes per row; 0 = through the staging tile
  int epi_helpers;   // 1: single accumulator and direct stores -> the splitter warps drain the upper half of the columns
  long long* trace;   // debug (TFGNN_B200_FUSED_TRACE=file): kFuTraceSlots clock64 stamps per CTA, see fu_trace()
  GemmEpilogue epi;
};

// Debug timeline: slot 0 kernel entry, 1 set-up done, 2 exit, 3 / 4 globaltimer at entry / exit; 8+cc gather warp 0 finished call cc
// (cc < 40); 48+2t / 49+2t first MMA / last commit of tile t (t < 16); 80+2t / 81+2t epilogue start / end of tile t;
// 112+t TMA producer got the first slot of unit t.  SM-local clock64: comparable inside one CTA only.
constexpr int kFuTraceSlots = 128;
__device__ __forceinline__ void fu_trace(const FusedParams& p, int idx) {
  if (p.trace) p.trace[(size_t)blockIdx.x * kFuTraceSlots + idx] = clock64();
}

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

// ---- gather warps: rolling cp.async ring ---------------------------------------------------------------------
// Each gather warp owns 8 consecutive targets of every tile and walks their CSR segments, one edge type after
// the other ("calls": (tile, type) -> one contiguous edge range).  Source rows do not wait in registers: the warp
// keeps Q rows in flight AT ALL TIMES as cp.async (LDGSTS) copies into its own Q shared-memory row slots - every
// lane copies and later reads back only its own 16-byte columns, one commit group per row, so
// `cp.async.wait_group Q-1` means "the oldest row has landed" and no mbarrier or warp barrier is needed.  When a
// row has been added to the register accumulator its slot is refilled at once with the row Q edges ahead, across
// segment, call and tile boundaries: the issue cursor and the consume cursor walk the same edge sequence
// independently, so the pipeline never drains.  tools/gather_ceiling.cu: 16 warps x 4 slots (64 KB of shared
// memory at D = 256) read 1 KB random rows at 7.4 TB/s, 16 warps x 4 register-held rows at 5.3 TB/s.
// Summation order inside a segment is ascending CSR order, as in the unfused path.
struct GatherIssue {
  const int* row_ptr;
  const int* src;
  const float* h;        // + 4 * lane: this lane's first 16-byte column
  long long M;
  int V, L, ldh, skip, C4;
  int lane, gw, Q, ncalls;
  long long unit0, unit_step;
  int ctas, rank;
  int rpw, row_off;      // rows of the tile owned by this warp, first row of this CTA's share of the tile
  uint32_t buf_s;        // shared-window address of this warp's slot 0, + 16 * lane
  uint32_t row_bytes;
  int ic;                // call being issued
  int i_pos, i_end, in_blk, i_blk;
  int rpA, rpB, rpC;     // row_ptr[v0 + lane] of calls ic, ic+1, ic+2 (prefetched: the loads are two calls old when used)
  int ids, ids_next, ids0B;
  uint32_t ibuf;         // slot the next row goes to
  int islot;

  __device__ __forceinline__ void call_rows(int n, int& l, int& v0, int& nr) const {
    const int u = n / L;
    l = n - u * L;
    const long long tile = (unit0 + (long long)u * unit_step) * ctas + rank;
    v0 = (int)(tile * kFuBM) + row_off + gw * rpw;
    nr = V - v0;
    nr = nr < 0 ? 0 : (nr > rpw ? rpw : nr);
  }
  __device__ __forceinline__ int rp_of(int n) const {   // row_ptr of the call's rows, lane r -> first edge of row r
    if (n >= ncalls) return 0;
    int l, v0, nr;
    call_rows(n, l, v0, nr);
    if (nr == 0) return 0;
    const int* base = row_ptr + (long long)l * V + v0;
    return __ldg(base + (skip ? 0 : (lane <= nr ? lane : nr)));
  }
  __device__ __forceinline__ int nrows_of(int n) const {
    int l, v0, nr;
    call_rows(n, l, v0, nr);
    return nr;
  }
  __device__ __forceinline__ int load_ids(int first) const {
    const long long i = (long long)first + lane;
    return i < M ? __ldg(src + i) : 0;
  }
  __device__ __forceinline__ void next_call() {
    ++ic;
    rpA = rpB;
    rpB = rpC;
    rpC = rp_of(ic + 2);
    i_pos = __shfl_sync(0xffffffffu, rpA, 0);
    i_end = __shfl_sync(0xffffffffu, rpA, nrows_of(ic));
    ids = ids0B;
    in_blk = 0;
    i_blk = i_pos;
    ids_next = load_ids(i_blk + 32);
    ids0B = load_ids(__shfl_sync(0xffffffffu, rpB, 0));
  }
  // one commit group per call of issue(): the next edge's row if there is one, else an empty group (only after
  // the last edge of the last call, so groups and consumed rows stay in step)
  // QT > 0: ring depth known at compile time (the default 4); FULL: D == 128 * NV, no column predicate
  template <int NV, int QT = 0, bool FULL = false>
  __device__ __forceinline__ void issue() {
    const int Qc = QT > 0 ? QT : Q;
    while (i_pos == i_end && ic + 1 < ncalls) next_call();
    if (i_pos < i_end) {
      if (in_blk == 32) {
        in_blk = 0;
        i_blk += 32;
        ids = ids_next;
        ids_next = load_ids(i_blk + 32);
      }
      const int s = __shfl_sync(0xffffffffu, ids, in_blk);
      const float* rowp = h + (long long)s * ldh;
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (FULL || lane + 32 * j < C4) ptx::cp_async16(ibuf + 512u * j, rowp + 128 * j);
      ++in_blk;
      ++i_pos;
    }
    ptx::cp_async_commit();
    if (++islot == Qc) {
      islot = 0;
      ibuf = buf_s;
    } else {
      ibuf += row_bytes;
    }
  }
};

__device__ __forceinline__ void cp_async_wait_oldest(int Q) {   // at most Q-1 groups stay pending
  switch (Q) {
    case 1: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 7: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}

// split != 0 (split-tile mode): this CTA gathers only rows [split_rank*64, split_rank*64 + 64) of every tile; the slot
// is shared with the peer CTA of the cluster, whose slot_ready barrier gets a (cluster-scope release) arrival as well.
template <int NV, int QT = 0, bool FULL = false>
__device__ __forceinline__ void gather_warp_main(const FusedParams& p, int lane, int gw, int Q_in, uint8_t* bufs,
                                                 long long unit0, long long unit_step, long long total_units,
                                                 int ctas, int rank, int ring_row0, uint64_t* slot_ready,
                                                 uint64_t* slot_free, int split, int split_rank,
                                                 uint32_t peer_slot_ready0) {
  const int Q = QT > 0 ? QT : Q_in;
  const int kRowsPerWarp = split ? kFuBM / 2 / kFuGatherWarps : kFuBM / kFuGatherWarps;
  const int row_off = split ? split_rank * (kFuBM / 2) : 0;
  const int D = p.D, C4 = p.D >> 2, normalize = p.normalize;
  const bool no_store = p.debug_skip & 16;   // timing experiment: gathered rows are not written to the ring
  const int kFuSlots = p.num_slots;
  float* const ring = p.ring;
  const uint64_t pol_keep = ptx::policy_evict_last();
  const long long my_units = unit0 < total_units ? (total_units - unit0 + unit_step - 1) / unit_step : 0;
  GatherIssue g;
  g.row_ptr = p.row_ptr; g.src = p.src; g.h = p.h + 4 * lane; g.M = p.M; g.V = p.V; g.L = p.L; g.ldh = p.ldh;
  g.skip = p.debug_skip & 1; g.C4 = C4;
  g.lane = lane; g.gw = gw; g.Q = Q; g.ncalls = (int)(my_units * p.L);
  g.unit0 = unit0; g.unit_step = unit_step; g.ctas = ctas; g.rank = rank;
  g.rpw = kRowsPerWarp; g.row_off = row_off;
  g.buf_s = ptx::smem_u32(bufs) + (uint32_t)lane * 16u;
  g.row_bytes = (uint32_t)p.D * 4;
  g.ic = 0; g.islot = 0; g.in_blk = 0; g.ibuf = g.buf_s;
  g.rpA = g.rp_of(0); g.rpB = g.rp_of(1); g.rpC = g.rp_of(2);
  g.i_pos = __shfl_sync(0xffffffffu, g.rpA, 0);
  g.i_end = __shfl_sync(0xffffffffu, g.rpA, g.nrows_of(0));
  g.i_blk = g.i_pos;
  g.ids = g.load_ids(g.i_blk);
  g.ids_next = g.load_ids(g.i_blk + 32);
  g.ids0B = g.load_ids(__shfl_sync(0xffffffffu, g.rpB, 0));
  // consume side: its own row_ptr registers (the issue side may already be several calls ahead after the fill)
  int crp = g.rpA, crp_next = g.rpB;
  int cslot = 0;
  uint32_t cbuf = g.buf_s;
  for (int i = 0; i < Q; ++i) g.template issue<NV, QT, FULL>();   // fill the ring

  float4 acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int cc = 0; cc < g.ncalls; ++cc) {
    int l, v0, nrows;
    g.call_rows(cc, l, v0, nrows);
    const int rp = crp;
    crp = crp_next;
    crp_next = g.rp_of(cc + 2);
    const int slot = cc % kFuSlots;
    if (split) ptx::mbar_wait_cluster(&slot_free[slot], ((cc / kFuSlots) & 1) ^ 1);   // the peer's reads are done too
    else ptx::mbar_wait_backoff(&slot_free[slot], ((cc / kFuSlots) & 1) ^ 1, p.sleep_long);
    float* dst = ring + ((size_t)ring_row0 + (size_t)slot * kFuBM + (size_t)row_off + (size_t)gw * kRowsPerWarp) * D + 4 * lane;
    int row = 0;
    int seg_begin = __shfl_sync(0xffffffffu, rp, 0);
    int seg_end = __shfl_sync(0xffffffffu, rp, 1);
    const int e_end = __shfl_sync(0xffffffffu, rp, nrows);
    auto flush = [&]() {   // closes row `row`
      const float scale = normalize ? 1.0f / ((float)(seg_end - seg_begin) + kSmallNumber) : 1.0f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (lane + 32 * j < C4 && !no_store)
          ptx::st_f4_hint(dst + 128 * j,
                          make_float4(acc[j].x * scale, acc[j].y * scale, acc[j].z * scale, acc[j].w * scale),
                          pol_keep);
        acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      dst += D;
      ++row;
      seg_begin = seg_end;
      seg_end = __shfl_sync(0xffffffffu, rp, row + 1);
    };
    for (int e = seg_begin; e < e_end; ++e) {
      while (e >= seg_end) flush();   // warp-uniform: close finished (possibly empty) segments
      cp_async_wait_oldest(Q);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (FULL || lane + 32 * j < C4) {
          const float4 x = ptx::lds_f4(cbuf + 512u * j);
          acc[j].x += x.x; acc[j].y += x.y; acc[j].z += x.z; acc[j].w += x.w;
        }
      }
      if (++cslot == Q) {
        cslot = 0;
        cbuf = g.buf_s;
      } else {
        cbuf += g.row_bytes;
      }
      g.template issue<NV, QT, FULL>();   // refill the slot just read (same lane, same bytes: no cross-lane hazard)
    }
    while (row < nrows) flush();
    // generic-proxy global writes -> visible to the TMA (async proxy) reads of this CTA
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __syncwarp();
    if (lane == 0 && gw == 0 && cc < 40) fu_trace(p, 8 + cc);
    if (lane == 0) {
      if (split) {   // both CTAs of the cluster read the whole slot: tell the peer's TMA producer too
        ptx::mbar_arrive_cluster_release(ptx::mapa_shared(ptx::smem_u32(&slot_ready[slot]), (uint32_t)split_rank));
        ptx::mbar_arrive_cluster_release(peer_slot_ready0 + (uint32_t)slot * 8u);
      } else {
        ptx::mbar_arrive(&slot_ready[slot]);
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// BK = floats per K block: 32 (128 B rows, SWIZZLE_128B) or 16 (64 B rows, SWIZZLE_64B).  The smaller block
// halves the bytes per pipeline stage, so twice as many stages fit: under the gather's L2 traffic a TMA round
// trip takes ~3 us, and it is bytes-in-flight / latency that bounds the operand feed of the tensor core.
// CTAS = 2: the two CTAs of a cluster (one TPC) form a tcgen05 CTA pair.  Each CTA gathers and splits ITS 128
// targets exactly as before, but the MMA is one cta_group::2 instruction of M = 256 issued by the leader (rank 0):
// every CTA stages only HALF of the weight tile's N rows, so the weight stream from L2 (the largest on-chip
// traffic of the kernel: 2 MB per 128 targets at H = 256) and its shared-memory footprint are halved.
// SPLIT (small batches, fewer than SMs/2 tiles): the two CTAs of a cluster share ONE 128-target tile.  Each gathers half of
// its rows into the common ring slot and contracts the whole tile with its HALF of the output columns (cta_group::1 MMAs,
// N = H/2 per CTA): the gather (bound by one SM's L2 bandwidth when a tile is all an SM has) and the K loop are both
// spread over twice as many SMs.  Ring slots are released only when both CTAs have read them.
template <int NV, int BK, int CTAS, bool SPLIT = false>
__global__ void __launch_bounds__(kFuThreads, 1)
fused_rgcn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const FusedParams p) {
  static_assert(!(SPLIT && CTAS != 1), "split-tile mode uses single-CTA MMAs");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kFuBK = BK;
  constexpr int kRowBytes = BK * 4;
  constexpr int kFuATileBytes = kFuBM * kRowBytes;
  const int S = p.num_stages;
  const int b_rows = p.block_n / CTAS;               // N rows of the weight tile staged by this CTA
  const int b_tile_bytes = b_rows * kRowBytes;
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
  float* epi_stage = reinterpret_cast<float*>(bars + ((3 * S + 4 + 2 * kFuMaxSlots + 2 + 1) & ~1));
  uint8_t* gbuf = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(reinterpret_cast<uint8_t*>(epi_stage) + kFuEpiBytes) + 127) & ~uintptr_t(127));  // [16 * Q] rows

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CTAS == 2 ? ptx::cluster_ctarank() : 0u;
  const uint32_t srank = SPLIT ? ptx::cluster_ctarank() : 0u;      // split-tile mode: which half (rows to gather, columns to produce)
  constexpr int kClu = (CTAS == 2 || SPLIT) ? 2 : 1;               // CTAs per cluster
  // one unit = CTAS consecutive 128-target tiles (one per CTA of the pair); the N dimension is covered in n_pass passes
  const long long total_units = (p.m_tiles + CTAS - 1) / CTAS;
  const long long unit0 = blockIdx.x / kClu, unit_step = gridDim.x / kClu;
  const int n_pass = p.n_tiles;
  const int kFuSlots = p.num_slots;
  const int kb_per_tile = p.L * p.kb_per_type;

  if (threadIdx.x == 0 && p.trace) {
    fu_trace(p, 0);
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[(size_t)blockIdx.x * kFuTraceSlots + 3] = (long long)gt;
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&split[s], 4 * CTAS);      // one arrival per splitter warp (of both CTAs of a pair)
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], (p.epi_helpers ? 8 : 4) * CTAS);  // one arrival per epilogue (+ helper) warp
    }
    for (int r = 0; r < kFuMaxSlots; ++r) {
      ptx::mbar_init(&slot_ready[r], kFuGatherWarps * (SPLIT ? 2 : 1));   // split: the peer's gather warps arrive too
      ptx::mbar_init(&slot_free[r], 4 * (SPLIT ? 2 : 1));                  // one arrival per splitter warp (of both CTAs)
    }
    ptx::fence_barrier_init();
  }
  if (kClu == 2) {
    // both CTAs' barriers must exist before any remote arrive / multicast commit can land on them
    __syncthreads();
    ptx::cluster_sync_all();
  }
  if (warp == 1) {
    if (CTAS == 2) {
      ptx::tmem_alloc_pair(tmem_slot, kFuTmemCols);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_slot, kFuTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (kClu == 2) ptx::cluster_sync_all();   // both halves of the pair's TMEM are allocated before the first MMA
  ptx::tc_fence_after_sync();
  if (threadIdx.x == 0) fu_trace(p, 1);
  // the leader's barriers that collect arrivals from both CTAs (split[], tmem_empty[])
  const uint32_t split_remote0 = CTAS == 2 ? ptx::mapa_shared(ptx::smem_u32(&split[0]), 0) : 0u;
  const uint32_t tmem_empty_remote0 = CTAS == 2 ? ptx::mapa_shared(ptx::smem_u32(&tmem_empty[0]), 0) : 0u;
  const uint32_t tmem_base = *tmem_slot;
  // first ring row of this CTA (split-tile mode: of this CLUSTER, both CTAs fill and read the same slots)
  const int ring_row0 = (SPLIT ? blockIdx.x / 2 : blockIdx.x) * kFuSlots * kFuBM;
  const uint32_t peer_slot_ready0 = SPLIT ? ptx::mapa_shared(ptx::smem_u32(&slot_ready[0]), srank ^ 1u) : 0u;
  const uint32_t peer_slot_free0 = SPLIT ? ptx::mapa_shared(ptx::smem_u32(&slot_free[0]), srank ^ 1u) : 0u;

  // ---- epilogue of one (tile, column range) ----
  // Run by the epilogue warps and - when the tile has a single accumulator pair, so that the next tile's MMAs wait for the drain
  // anyway - by the splitter warps as well ("helpers": the upper half of the columns; same TMEM lane quarters, warp & 3).
  // Trace r2n (tools/fused_trace_summary.py): 17-20 us per 128x256 tile with 4 warps, a staging tile and the chunk in local
  // memory, against 47 us of K loop, strictly serialised.
  const uint32_t n_acc = p.block_n <= 128 ? 2 : 1;
  const uint32_t corr_off = p.block_n <= 128 ? 128 : 256;
  const bool helpers = p.epi_helpers != 0;
  const int epi_half = ((p.block_n / 16 + 1) / 2) * 16;
  auto epilogue_cols = [&](long long tp, uint32_t tile_count, int col_begin, int col_end, float* stage, bool traced) {
    const int q = warp & 3;
    const uint64_t pol_stream = ptx::policy_evict_first();
    const long long m0 = ((tp / n_pass) * CTAS + rank) * kFuBM;
    const int n0 = (int)(tp % n_pass) * p.block_n + (int)srank * p.block_n;
    const uint32_t acc = tile_count % n_acc, acc_ph = (tile_count / n_acc) & 1;
    ptx::mbar_wait_backoff(&tmem_full[acc], acc_ph, p.sleep_long);
    ptx::tc_fence_after_sync();
    if (traced && lane == 0 && tile_count < 16) fu_trace(p, 80 + 2 * (int)tile_count);
    const long long row = m0 + q * 32 + lane;
    const bool row_ok = row < p.V;
    const float inv_rn = row_ok ? 1.0f / fu_row_norm(p.epi, row) : 1.0f;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kFuAccStride;
    // one 16-column chunk of this thread's row: TMEM -> registers, row norm / bias / activation (all indices compile-time:
    // the chunk stays in registers)
    auto load_chunk = [&](int col, float (&v)[16]) {
      uint32_t mv[16], cv[16];
      ptx::tmem_ld_x16_nowait(taddr + col, mv);
      ptx::tmem_ld_x16_nowait(taddr + corr_off + col, cv);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(mv[j]) + __uint_as_float(cv[j]);
      // uniform branches hoisted out of the element loops (predicated-off code still costs issue slots
      // and instruction-cache space: the epilogue was ~48 us per 128x256 tile before)
      if (p.epi.row_norm) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= inv_rn;
      }
      if (p.epi.bias) {
        const float4* bp = reinterpret_cast<const float4*>(p.epi.bias + n0 + col);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 bb = __ldg(bp + j);
          v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
        }
      }
      apply_act_vec<16>(v, p.epi.act);
    };
    // Fused LayerNormalization (gnn.py:317-321 right after the message-passing layer): this thread owns one whole row
    // (single N pass, no helpers), so mean and variance are two extra sweeps over its TMEM columns - no [V,H] round trip
    // through HBM.  Two-pass variance (mean first) like Keras: no cancellation.
    float ln_mean = 0.f, ln_rstd = 1.f;
    const bool ln = p.epi.ln_gamma != nullptr;
    if (ln && !(p.debug_skip & 4)) {
      float s1 = 0.f;
      for (int col = 0; col < p.block_n; col += 16) {
        float v[16];
        load_chunk(col, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) s1 += v[j];
      }
      ln_mean = s1 / (float)p.block_n;
      float s2 = 0.f;
      for (int col = 0; col < p.block_n; col += 16) {
        float v[16];
        load_chunk(col, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) { const float dlt = v[j] - ln_mean; s2 = fmaf(dlt, dlt, s2); }
      }
      ln_rstd = rsqrtf(s2 / (float)p.block_n + p.epi.ln_eps);
    }
    auto ln_affine = [&](int col, float (&v)[16]) {
      const float4* gp = reinterpret_cast<const float4*>(p.epi.ln_gamma + n0 + col);
      const float4* bp = reinterpret_cast<const float4*>(p.epi.ln_beta + n0 + col);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 gg = __ldg(gp + j), bb = __ldg(bp + j);
        v[4 * j] = (v[4 * j] - ln_mean) * ln_rstd * gg.x + bb.x;
        v[4 * j + 1] = (v[4 * j + 1] - ln_mean) * ln_rstd * gg.y + bb.y;
        v[4 * j + 2] = (v[4 * j + 2] - ln_mean) * ln_rstd * gg.z + bb.z;
        v[4 * j + 3] = (v[4 * j + 3] - ln_mean) * ln_rstd * gg.w + bb.w;
      }
    };
    if (p.debug_skip & 4) col_end = col_begin;
    if (p.epi_direct) {
      // Straight from registers: each thread stores its own row's 64 B piece of the chunk (four 16 B stores; the L2 merges
      // the pieces of a line before it is written back).  No staging tile, no warp barrier, a third of the instructions.
      float* crow = p.C + row * p.ldc + n0;
      for (int col = col_begin; col < col_end; col += 16) {
        float v[16];
        load_chunk(col, v);
        if (ln) ln_affine(col, v);
        if (p.epi_direct == 3) {
          // 64 contiguous bytes per row and store instruction: 4x4 float4 transpose over lane quads (sm100_ptx.cuh)
          const int t = lane & 3;
          float4 C0, C1, C2, C3;
          ptx::quad_transpose_f4(v, lane, C0, C1, C2, C3);
          float* base = crow - (long long)t * p.ldc + col + 4 * t;
          const long long r_base = row - t;
          if (r_base < p.V) ptx::st_f4_hint(base, C0, pol_stream);
          if (r_base + 1 < p.V) ptx::st_f4_hint(base + p.ldc, C1, pol_stream);
          if (r_base + 2 < p.V) ptx::st_f4_hint(base + 2 * (long long)p.ldc, C2, pol_stream);
          if (r_base + 3 < p.V) ptx::st_f4_hint(base + 3 * (long long)p.ldc, C3, pol_stream);
        } else if (p.epi_direct == 2) {
          // full 32 B sectors per store instruction: lanes 2i / 2i+1 swap every second float4 and write 32 contiguous bytes
          // of one of their two rows (gemm_tc.cu: tc_store_pairwise)
          const bool odd = lane & 1;
          float* pe = (odd ? crow - p.ldc : crow) + col + (odd ? 4 : 0);
          float* po = (odd ? crow : crow + p.ldc) + col + (odd ? 4 : 0);
          const bool ok_e = (odd ? row - 1 : row) < p.V, ok_o = (odd ? row : row + 1) < p.V;
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            const float rx = __shfl_xor_sync(0xffffffffu, odd ? v[j] : v[j + 4], 1);
            const float ry = __shfl_xor_sync(0xffffffffu, odd ? v[j + 1] : v[j + 5], 1);
            const float rz = __shfl_xor_sync(0xffffffffu, odd ? v[j + 2] : v[j + 6], 1);
            const float rw = __shfl_xor_sync(0xffffffffu, odd ? v[j + 3] : v[j + 7], 1);
            const float4 mine = odd ? make_float4(v[j + 4], v[j + 5], v[j + 6], v[j + 7]) : make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            const float4 got = make_float4(rx, ry, rz, rw);
            if (ok_e) ptx::st_f4_hint(pe + j, odd ? got : mine, pol_stream);
            if (ok_o) ptx::st_f4_hint(po + j, odd ? mine : got, pol_stream);
          }
        } else if (row_ok) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            ptx::st_f4_hint(crow + col + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]), pol_stream);
        }
      }
    } else {
      // Through a padded staging tile to 64 B row pieces, 8 rows per store instruction: the form the NVLink stores of the
      // fused all-gather want.  (32-column phases - 128-byte row pieces - were tried for them: 1-2 % slower on one GPU in a
      // same-box A/B, gpurun r2l, so not kept.)
      for (int c0 = col_begin; c0 < col_end; c0 += 16) {
        {
          float v[16];
          load_chunk(c0, v);
          if (ln) ln_affine(c0, v);
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(stage + lane * kFuEpiPitch + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        __syncwarp();
        const int rr = lane >> 2, cc = (lane & 3) * 4;
        for (int r0 = 0; r0 < 32; r0 += 8) {
          const int r = r0 + rr;
          const long long grow = m0 + q * 32 + r;
          if (grow < p.V) {
            const float4 val = *reinterpret_cast<const float4*>(stage + r * kFuEpiPitch + cc);
            const long long off = grow * p.ldc + n0 + c0 + cc;
            if (p.C_mc) {
              // the all-gather of the sharded layer through the switch: ONE multimem.st, NVSwitch replicates the 16 bytes
              // into every GPU's copy of the table (this GPU's included)
              asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.C_mc + off), "f"(val.x),
                           "f"(val.y), "f"(val.z), "f"(val.w)
                           : "memory");
            } else {
              ptx::st_f4_hint(p.C + off, val, pol_stream);
              // the all-gather of the sharded layer, tile by tile: the same 16 bytes go to every peer's copy of the table
              // over NVLink (plain stores to P2P-mapped memory; the copies become the next layer's source table)
              for (int pr = 0; pr < p.n_peer; ++pr) *reinterpret_cast<float4*>(p.C_peer[pr] + off) = val;
            }
          }
        }
        __syncwarp();
      }
    }
    ptx::tc_fence_before_sync();
    __syncwarp();
    if (traced && lane == 0 && tile_count < 16) fu_trace(p, 81 + 2 * (int)tile_count);
    if (lane == 0) {
      if (CTAS == 2 && rank != 0) ptx::mbar_arrive_cluster(tmem_empty_remote0 + acc * 8u);
      else ptx::mbar_arrive(&tmem_empty[acc]);
    }
  };

  // N passes: main + correction accumulators need 2*block_n <= 512 TMEM columns, so H in (256, 512] is covered
  // in two passes of block_n = H/2 columns over the SAME gathered ring slots (the ring then holds all L types of
  // the tile: num_slots = L + 1); every source row is still gathered exactly once.
  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0, slot_it = 0;
      const uint64_t pol_keep = ptx::policy_evict_last();   // ring slots and the 2 MB of weights stay in L2
      for (long long unit = unit0; unit < total_units; unit += unit_step, slot_it += p.L) {
       for (int pass = 0; pass < n_pass; ++pass) {
        const int n0 = pass * p.block_n + (int)rank * b_rows + (int)srank * p.block_n;
        for (int l = 0; l < p.L; ++l) {
          const uint32_t sq = slot_it + l;
          const int slot = sq % kFuSlots;
          if (SPLIT) ptx::mbar_wait_cluster(&slot_ready[slot], (sq / kFuSlots) & 1);   // half of the rows come from the peer SM
          else ptx::mbar_wait_backoff(&slot_ready[slot], (sq / kFuSlots) & 1, p.sleep_long);
          if (l == 0 && pass == 0 && slot_it / p.L < 16) fu_trace(p, 112 + (int)(slot_it / p.L));
          for (int kb = 0; kb < p.kb_per_type; ++kb, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            ptx::mbar_wait_backoff(&empty[s], ph ^ 1, p.sleep_crit);
            uint8_t* st = smem + (size_t)s * stage_bytes;
            const bool skip_b = p.debug_skip & 8;   // timing experiment: no weight stream (results invalid)
            ptx::mbar_arrive_expect_tx(&full[s], kFuATileBytes + (skip_b ? 0 : 2 * b_tile_bytes));
            ptx::tma_load_2d_hint(st, &map_a, &full[s], kb * kFuBK, ring_row0 + slot * kFuBM, pol_keep);
            const int kcol = (l * p.kb_per_type + kb) * kFuBK;
            if (!skip_b) {
              ptx::tma_load_2d_hint(st + 2 * kFuATileBytes, &map_b, &full[s], kcol, n0, pol_keep);
              ptx::tma_load_2d_hint(st + 2 * kFuATileBytes + b_tile_bytes, &map_b, &full[s], kcol, p.N + n0, pol_keep);
            }
          }
        }
       }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (the leader CTA of a pair issues for both) =================
    const uint32_t idesc = ptx::umma_idesc_tf32(128u * CTAS, (uint32_t)p.block_n);
    const uint32_t idesc_bf = ptx::umma_idesc_bf16(128u * CTAS, (uint32_t)p.block_n);
    uint32_t it = 0, tile_count = 0;
    for (long long tp = unit0 * n_pass; rank == 0 && tp < total_units * n_pass;
         tp = (tp % n_pass == n_pass - 1) ? tp + (unit_step - 1) * n_pass + 1 : tp + 1, ++tile_count) {
      const uint32_t acc = tile_count % n_acc, acc_ph = (tile_count / n_acc) & 1;
      ptx::mbar_wait_backoff(&tmem_empty[acc], acc_ph ^ 1, p.sleep_long);
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * kFuAccStride;
      const uint32_t c_tmem = d_tmem + corr_off;
      for (int kb = 0; kb < kb_per_tile; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        ptx::mbar_wait_backoff(&full[s], ph, p.sleep_crit);
        ptx::mbar_wait_backoff(&split[s], ph, p.sleep_crit);   // pair: arrivals of both CTAs' splitter warps = both operand halves staged
        ptx::tc_fence_after_sync();
        if (lane == 0) {
          if (kb == 0 && tile_count < 16) fu_trace(p, 48 + 2 * (int)tile_count);
          const uint32_t st = ptx::smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t a_hi = ptx::umma_desc_k<kRowBytes>(st);
          const uint64_t a_lo = ptx::umma_desc_k<kRowBytes>(st + kFuATileBytes);
          const uint64_t b_hi = ptx::umma_desc_k<kRowBytes>(st + 2 * kFuATileBytes);
          const uint64_t b_lo = ptx::umma_desc_k<kRowBytes>(st + 2 * kFuATileBytes + b_tile_bytes);
#pragma unroll
          for (int k = 0; k < kFuBK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            if (CTAS == 2) {
              if (p.corr_bf16) {
                ptx::mma_bf16_ss_pair(c_tmem, a_lo + adv, b_lo + adv, idesc_bf, (kb | k) != 0);   // a lo(b) + lo(a) b
              } else {
                ptx::mma_tf32_ss_pair(c_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
                ptx::mma_tf32_ss_pair(c_tmem, a_hi + adv, b_lo + adv, idesc, 1);
              }
              ptx::mma_tf32_ss_pair(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
            } else {
              if (p.corr_bf16) {
                ptx::mma_bf16_ss(c_tmem, a_lo + adv, b_lo + adv, idesc_bf, (kb | k) != 0);
              } else {
                ptx::mma_tf32_ss(c_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
                ptx::mma_tf32_ss(c_tmem, a_hi + adv, b_lo + adv, idesc, 1);
              }
              ptx::mma_tf32_ss(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
            }
          }
          if (CTAS == 2) {
            ptx::mma_commit_pair(&empty[s], 3);     // frees stage s in BOTH CTAs
            if (kb == kb_per_tile - 1) ptx::mma_commit_pair(&tmem_full[acc], 3);
          } else {
            ptx::mma_commit(&empty[s]);
            if (kb == kb_per_tile - 1) ptx::mma_commit(&tmem_full[acc]);
          }
          if (kb == kb_per_tile - 1 && tile_count < 16) fu_trace(p, 49 + 2 * (int)tile_count);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================= A splitters =================
    const int tid = threadIdx.x - 128;
    const bool raw_hi = (p.debug_skip & 32) || p.corr_bf16;   // the fp32 tile stays in place as the hi operand
    const bool pair = p.corr_bf16;
    uint32_t it = 0, slot_base_it = 0, tile_count = 0;
    for (long long unit = unit0; unit < total_units; unit += unit_step, slot_base_it += p.L) {
     for (int pass = 0; pass < n_pass; ++pass) {
      for (int l = 0; l < p.L; ++l) {
        const uint32_t slot_it = slot_base_it + l;
        for (int kb = 0; kb < p.kb_per_type; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          ptx::mbar_wait_backoff(&full[s], ph, p.sleep_crit);
          if (kb == p.kb_per_type - 1 && pass == n_pass - 1) {
            // all TMA reads of the slot have landed: its lines are dead.  Discard them from L2 so that they are
            // never written back to HBM (the ring is pure on-chip hand-off), then hand the slot back.
            if (p.discard_ring && !SPLIT) {   // split-tile mode: the peer may still be reading the slot
              const char* sb = reinterpret_cast<const char*>(p.ring + ((size_t)ring_row0 + (size_t)(slot_it % kFuSlots) * kFuBM) * p.D);
              const int lines = kFuBM * p.D * 4 / 128;
              for (int i = tid; i < lines; i += 128) ptx::discard_l2_128(sb + (size_t)i * 128);
            }
            __syncwarp();
            if (lane == 0) {
              ptx::mbar_arrive(&slot_free[slot_it % kFuSlots]);
              if (SPLIT) ptx::mbar_arrive_cluster_release(peer_slot_free0 + (uint32_t)(slot_it % kFuSlots) * 8u);
            }
          }
          // all 16-byte loads of this thread first, through explicit shared-space instructions (generic pointers made the
          // compiler keep every load behind the previous iteration's stores: one shared-memory round trip per 16 bytes)
          const uint32_t a_s = ptx::smem_u32(smem + (size_t)s * stage_bytes) + (uint32_t)tid * 16u;
          const uint32_t lo_s = a_s + kFuATileBytes;
          constexpr int kIt = kFuATileBytes / 16 / 128;
          float4 xs[kIt];
#pragma unroll
          for (int i = 0; i < kIt; ++i) xs[i] = ptx::lds_f4(a_s + (uint32_t)i * 2048u);
#pragma unroll
          for (int i = 0; i < kIt; ++i) {
            const float4 x = xs[i];
            float4 hh, ll;
            hh.x = ptx::tf32_hi(x.x); hh.y = ptx::tf32_hi(x.y); hh.z = ptx::tf32_hi(x.z); hh.w = ptx::tf32_hi(x.w);
            if (pair) {   // one tile of bf16 pairs (a | a - tf32(a)) in the bytes of the lo tile
              uint4 w;
              w.x = ptx::pack_bf16x2(x.x - hh.x, x.x); w.y = ptx::pack_bf16x2(x.y - hh.y, x.y);
              w.z = ptx::pack_bf16x2(x.z - hh.z, x.z); w.w = ptx::pack_bf16x2(x.w - hh.w, x.w);
              ptx::sts_u4(lo_s + (uint32_t)i * 2048u, w);
            } else {
              ll.x = ptx::tf32_hi(x.x - hh.x); ll.y = ptx::tf32_hi(x.y - hh.y);
              ll.z = ptx::tf32_hi(x.z - hh.z); ll.w = ptx::tf32_hi(x.w - hh.w);
              if (!raw_hi) ptx::sts_f4(a_s + (uint32_t)i * 2048u, hh);   // raw_hi: the tensor core itself drops the low 13 mantissa bits
              ptx::sts_f4(lo_s + (uint32_t)i * 2048u, ll);
            }
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {   // one (possibly remote) arrival per warp: remote mbarrier arrives are DSMEM round trips
            if (CTAS == 2 && rank != 0) ptx::mbar_arrive_cluster(split_remote0 + (uint32_t)s * 8u);
            else ptx::mbar_arrive(&split[s]);
          }
        }
      }
      // single accumulator: nothing to split until this tile is drained - drain the upper half of its columns
      if (helpers) epilogue_cols(unit * n_pass + pass, tile_count, epi_half, p.block_n, nullptr, false);
      ++tile_count;
     }
    }
  } else if (warp >= 8 && warp < kFuFirstGatherWarp) {
    // ================= epilogue (warps 8..11 -> TMEM lane quarters 0..3) =================
    uint32_t tile_count = 0;
    const int col_end = helpers ? epi_half : p.block_n;
    for (long long tp = unit0 * n_pass; tp < total_units * n_pass;
         tp = (tp % n_pass == n_pass - 1) ? tp + (unit_step - 1) * n_pass + 1 : tp + 1, ++tile_count)
      epilogue_cols(tp, tile_count, 0, col_end, epi_stage + (size_t)(warp - 8) * 32 * kFuEpiPitch, warp == 8);
  } else if (warp >= kFuFirstGatherWarp) {
    // ================= gather warps =================
    const int gw = warp - kFuFirstGatherWarp;
    // common case (ring depth 4, D a multiple of 128): depth and column predicate resolved at compile time - the gather
    // loop runs once per EDGE and the kernel is co-limited by issue slots (58 %, ncu r2c)
    uint8_t* my_bufs = gbuf + (size_t)gw * p.gather_q * ((size_t)p.D * 4);
#define TFGNN_FU_GATHER(QT, FULL)                                                                                      \
    gather_warp_main<NV, QT, FULL>(p, lane, gw, p.gather_q, my_bufs, unit0, unit_step, total_units, CTAS, (int)rank,     \
                                   ring_row0, slot_ready, slot_free, SPLIT ? 1 : 0, (int)srank, peer_slot_ready0)
    // same-box A/B (gpurun r2l): the compile-time variant wins at D = 256 (4.56 vs 4.77 ms) but loses at D = 320, where the
    // third column group is half empty (6.84 with the generic loop vs 6.95 with a <4, false> instantiation)
    if (p.gather_q == 4 && p.D == 128 * NV) TFGNN_FU_GATHER(4, true);
    else TFGNN_FU_GATHER(0, false);
#undef TFGNN_FU_GATHER
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0 && p.trace) {
    fu_trace(p, 2);
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[(size_t)blockIdx.x * kFuTraceSlots + 4] = (long long)gt;
  }
  if (kClu == 2) ptx::cluster_sync_all();   // the leader's MMAs read the peer's shared memory / remote arrivals: leave together
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    if (CTAS == 2) ptx::tmem_dealloc_pair(tmem_base, kFuTmemCols);
    else ptx::tmem_dealloc(tmem_base, kFuTmemCols);
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

// ---- device-global state: the persisting-L2 carve-out ---------------------------------------------------------
// cudaLimitPersistingL2CacheSize is a property of the DEVICE CONTEXT the host framework shares with this library, so
// it is handled explicitly (include/tfgnn_b200.h, "Device-global state"): the first fused launch on a device saves
// the current limit and raises it to the configured size (default 72 MB: ring + packed weights);
// tfgnn_b200_set_l2_persist_mb(0) / TFGNN_B200_L2_PERSIST_MB=0 opt out (the kernel stays correct, the ring then
// spills ~4 GB per layer to HBM at cfg2); tfgnn_b200_release_device_state() restores the saved limit.
static std::mutex g_l2_mu;
static int g_l2_want_mb = -1;                 // -1: default / environment
static bool g_l2_set[64] = {};
static size_t g_l2_saved[64] = {};

int set_l2_persist_mb(int mb) {
  std::lock_guard<std::mutex> lock(g_l2_mu);
  g_l2_want_mb = mb;
  for (bool& b : g_l2_set) b = false;         // re-applied by the next fused launch
  return 0;
}

static size_t g_l2_effective[64] = {};   // persisting bytes in effect on the device since the last ensure()

// need_bytes: what this launch wants resident (its ring + packed weights).  Default policy: max(72 MB, need + 4 MB), clamped
// to the device maximum and only ever RAISED (H = 320 at 1M nodes: the ring is 97 MB, and with a 72 MB carve-out 2.7 GB of
// it per layer were written back to HBM, ncu r2n); an explicit size (API / environment) is taken as is.
static int ensure_l2_persist_carveout(size_t need_bytes) {
  int dev = 0;
  TFGNN_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return 0;
  std::lock_guard<std::mutex> lock(g_l2_mu);
  long long want_mb = g_l2_want_mb;
  if (want_mb < 0) {
    static const int env_mb = [] { const char* e = getenv("TFGNN_B200_L2_PERSIST_MB"); return e ? atoi(e) : -1; }();
    want_mb = env_mb;
  }
  size_t want = want_mb >= 0 ? (size_t)want_mb << 20 : std::max((size_t)72 << 20, need_bytes + ((size_t)4 << 20));
  if (g_l2_set[dev] && (want_mb >= 0 || g_l2_effective[dev] >= want)) return 0;
  g_l2_set[dev] = true;
  if (want == 0) return 0;
  int max_persist = 0;
  if (cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev) != cudaSuccess || max_persist <= 0) {
    cudaGetLastError();
    return 0;
  }
  if (want > (size_t)max_persist) want = (size_t)max_persist;
  size_t cur = 0;
  if (cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) != cudaSuccess) { cudaGetLastError(); return 0; }
  g_l2_effective[dev] = cur;
  if (cur >= want) {                          // the host (or an earlier launch) already reserves at least as much
    if (want == (size_t)max_persist) g_l2_effective[dev] = (size_t)-1;   // nothing more to get: stop asking
    return 0;
  }
  if (!g_l2_saved[dev]) g_l2_saved[dev] = cur + 1;   // +1: "saved" marker (0 = nothing to restore)
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
  cudaGetLastError();
  g_l2_effective[dev] = want == (size_t)max_persist ? (size_t)-1 : want;
  return 0;
}

void restore_l2_persist_carveout() {
  std::lock_guard<std::mutex> lock(g_l2_mu);
  int cur_dev = 0;
  if (cudaGetDevice(&cur_dev) != cudaSuccess) { cudaGetLastError(); return; }
  for (int d = 0; d < 64; ++d) {
    if (!g_l2_saved[d]) continue;
    if (cudaSetDevice(d) == cudaSuccess) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, g_l2_saved[d] - 1);
    g_l2_saved[d] = 0;
    g_l2_set[d] = false;
    g_l2_effective[d] = 0;
  }
  cudaSetDevice(cur_dev);
  cudaGetLastError();
}

constexpr int kFuMaxGrid = 160;
static int fused_num_slots(int L, int H, bool split = false) {
  static const int env_slots = [] { const char* e = getenv("TFGNN_B200_RING_SLOTS"); return e ? atoi(e) : 0; }();
  if (H > 256 && !split) return L + 1;
  return (env_slots >= 2 && env_slots <= kFuMaxSlots) ? env_slots : 4;   // cfg2: 4 slots 4.61 ms, 3: 4.63-4.88, 5: 4.70, 2: 5.05
}
size_t fused_rgcn_ring_bytes(int D, int L, int H) {
  const int slots = fused_num_slots(L, H) > 4 ? fused_num_slots(L, H) : 4;   // covers the split-tile mode (4 slots) as well
  return (size_t)kFuMaxGrid * slots * kFuBM * D * sizeof(float);
}

int launch_fused_rgcn(const float* h, int D, const int* row_ptr, const int* src, long long M, int V, int L,
                      int normalize, const float* packedB, int corr_bf16, int H, float* ring, float* out, int ldo,
                      const GemmEpilogue& epi, cudaStream_t st, float* const* peer_out, int n_peer_out, float* mc_out) {
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
  p.M = M;
  static const int discard_env = [] { const char* e = getenv("TFGNN_B200_RING_DISCARD"); return e ? atoi(e) : 1; }();
  p.discard_ring = discard_env;
  static const int dbg_env = [] { const char* e = getenv("TFGNN_B200_DEBUG_SKIP"); return e ? atoi(e) : 0; }();
  p.debug_skip = dbg_env;
  p.corr_bf16 = corr_bf16;
  p.n_peer = 0;
  for (int r = 0; r < n_peer_out && r < TFGNN_MAX_PEERS; ++r) p.C_peer[p.n_peer++] = peer_out[r];
  p.C_mc = mc_out;
  p.N = H;
  p.m_tiles = ((long long)V + kFuBM - 1) / kFuBM;
  if (sms > kFuMaxGrid) sms = kFuMaxGrid;
  // Split-tile mode (small batches): fewer tiles than half the SMs -> two CTAs share each tile (rows of the gather,
  // columns of the contraction).  TFGNN_B200_FUSED_SPLIT: 0 = never, 1 = default rule (read per call: the tests sweep it)
  const char* split_str = getenv("TFGNN_B200_FUSED_SPLIT");
  const int split_env = split_str ? atoi(split_str) : 1;
  const char* pair_str = getenv("TFGNN_B200_FUSED_PAIR");
  const int pair_env = pair_str ? atoi(pair_str) : 1;
  const bool want_ln = epi.ln_gamma != nullptr;   // the row statistics need the whole row in one CTA: single N pass, no split
  TFGNN_REQUIRE(!want_ln || H <= 256, "fused LayerNorm needs hidden_dim <= 256 (one N pass)");
  // Large batches at H > 256 (TFGNN_B200_FUSED_SPLIT=2, experiment): the split form covers all H columns in ONE pass (each CTA
  // holds main + correction accumulators of H/2 columns) and its clusters share their ring slots, so the ring is half as large
  // (H = 320 at 1M nodes: 47 MB instead of 97 MB, which no longer fits the 79 MB persisting carve-out and spilled ~4 GB per layer).
  const bool split_big = split_env == 2 && H > 256;
  const bool split = !want_ln && split_env != 0 && pair_env != 2 && (2 * p.m_tiles <= sms || split_big) && H % 32 == 0 &&
                     H / 2 >= 16 && H / 2 <= 256;
  p.n_tiles = split ? 1 : (H > 256 ? 2 : 1);            // N passes (per CTA)
  p.block_n = split ? H / 2 : H / p.n_tiles;
  p.num_slots = fused_num_slots(L, H, split);
  // K block: 32 floats (128 B rows, SWIZZLE_128B) since the CTA-pair kernel; measured on cfg2 4.63 ms vs 4.88 ms with
  // 16 floats, H=320 6.47 vs 6.61 ms (half as many barrier round trips per byte; 2 stages of 64 KB still fit)
  static const int bk_env = [] { const char* e = getenv("TFGNN_B200_FUSED_BK"); return e ? atoi(e) : 32; }();
  p.C = out; p.ldc = ldo; p.epi = epi;
  // Epilogue form.  TFGNN_B200_EPI_DIRECT=0: always through the staging tile; TFGNN_B200_EPI_HELPERS=0: epilogue warps only
  // (read per call: the tests sweep both).
  {
    const char* ed = getenv("TFGNN_B200_EPI_DIRECT");
    const char* eh = getenv("TFGNN_B200_EPI_HELPERS");
    p.epi_direct = (p.n_peer == 0 && p.C_mc == nullptr && !(ed && atoi(ed) == 0)) ? (ed ? atoi(ed) : 3) : 0;   // 1: own 16 B pieces, 2: lane pairs (32 B), 3: lane quads (64 B)
    p.epi_helpers = (p.epi_direct && p.block_n > 128 && !want_ln && !(eh && atoi(eh) == 0)) ? 1 : 0;
    const char* sc = getenv("TFGNN_B200_SLEEP_CRIT");
    const char* sl = getenv("TFGNN_B200_SLEEP_LONG");
    p.sleep_crit = sc ? (uint32_t)atoi(sc) : 0u;
    p.sleep_long = sl ? (uint32_t)atoi(sl) : 0u;
  }
  // CTA pairs (cta_group::2) when there is at least one 128-target tile per SM; tiny batches keep single CTAs
  // TFGNN_B200_FUSED_PAIR: 0 = never, 1 = default rule, 2 = whenever there are two tiles (tests); read per call
  const bool pair_ok = !split && (p.block_n / 2) % 8 == 0 && sms >= 2;
  const int ctas = (pair_ok && ((pair_env == 1 && p.m_tiles >= sms) || (pair_env == 2 && p.m_tiles >= 2))) ? 2 : 1;
  int grid = split ? (int)(2 * p.m_tiles < (sms & ~1) ? 2 * p.m_tiles : (sms & ~1)) : (int)(p.m_tiles < sms ? p.m_tiles : sms);
  if (ctas == 2) grid &= ~1;
  // shared memory: S pipeline stages + epilogue staging + Q row slots for each of the 16 gather warps.
  // Q = 4 rolling copies per warp saturate HBM in isolation (tools/gather_ceiling.cu); the pipeline gets what is left.
  static const int stage_env = [] { const char* e = getenv("TFGNN_B200_FUSED_STAGES"); return e ? atoi(e) : 0; }();
  const char* q_str = getenv("TFGNN_B200_GATHER_Q");   // read per call: the tests sweep it
  const int q_env = q_str ? atoi(q_str) : 0;
  const int fixed_bytes = 2048 + kFuEpiBytes + 128 + 1024;   // barriers, epilogue staging, alignment slack
  const int row_bytes = D * 4;
  const int want_q = q_env >= 1 && q_env <= kFuMaxQ ? q_env : 4;
  int kFuBK = 0, stage_bytes = 0, stages = 0, q = 0;
  auto plan = [&](int bk) {   // most stages (<= 4) that still leave want_q row slots per gather warp; at least 2
    kFuBK = bk;
    stage_bytes = 2 * kFuBM * bk * 4 + 2 * (p.block_n / ctas) * bk * 4;
    auto q_for = [&](int s_) { return (kFuSmemLimit - fixed_bytes - s_ * stage_bytes) / (kFuGatherWarps * row_bytes); };
    stages = stage_env >= 2 ? stage_env : 4;
    while (stages > 2 && q_for(stages) < want_q) --stages;
    q = q_for(stages);
    if (q > want_q) q = want_q;
  };
  // K block of 32 floats (128 B rows, SWIZZLE_128B) when two 64-96 KB stages still leave >= 3 row slots (CTA pairs at
  // H <= 256: cfg2 4.63 ms vs 4.88 ms, H=320 6.47 vs 6.61 ms), else 16 floats (64 B rows, SWIZZLE_64B).
  plan(bk_env == 16 ? 16 : 32);
  if (kFuBK == 32 && q < (want_q < 3 ? want_q : 3)) plan(16);
  TFGNN_REQUIRE(q >= 1 && stages >= 2, "fused RGCN: tile does not fit shared memory");
  const int kFuATileBytes = kFuBM * kFuBK * 4;
  p.kb_per_type = D / kFuBK;
  if (p.debug_skip & 2) p.kb_per_type = 1;
  p.num_stages = stages;
  p.gather_q = q;

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
    cuuint32_t box[2] = {(cuuint32_t)kFuBK, (cuuint32_t)(p.block_n / ctas)};
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
                            kFuEpiBytes + 128 + (size_t)kFuGatherWarps * q * row_bytes + 1024;
  TFGNN_REQUIRE(smem_bytes <= (size_t)kFuSmemLimit, "fused RGCN: shared memory budget exceeded");
  const int nv = (D + 127) / 128;
  // L2 set-aside for the evict_last (persisting) lines: the ring + the packed weights.  Without a carve-out
  // the evict_last hint is advisory only and the ring gets written back to HBM (measured: +4 GB/layer).
  {
    const int rc_l2 = ensure_l2_persist_carveout((size_t)grid * p.num_slots * kFuBM * D * sizeof(float) +
                                                 (size_t)2 * H * L * D * sizeof(float));
    if (rc_l2) return rc_l2;
  }
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    auto set = [&](const void* fn) {
      cudaError_t x = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemLimit);
      if (x != cudaSuccess) attr_err = x;
    };
#define TFGNN_FU_SET(NV) \
    set((const void*)fused_rgcn_kernel<NV, 16, 1>); set((const void*)fused_rgcn_kernel<NV, 32, 1>); \
    set((const void*)fused_rgcn_kernel<NV, 16, 2>); set((const void*)fused_rgcn_kernel<NV, 32, 2>); \
    set((const void*)fused_rgcn_kernel<NV, 16, 1, true>); set((const void*)fused_rgcn_kernel<NV, 32, 1, true>);
    TFGNN_FU_SET(1) TFGNN_FU_SET(2) TFGNN_FU_SET(3) TFGNN_FU_SET(4)
#undef TFGNN_FU_SET
  });
  TFGNN_CUDA(attr_err);
  p.trace = nullptr;
  if (const char* tr = getenv("TFGNN_B200_FUSED_TRACE"); tr && *tr) {
    TFGNN_CUDA(cudaMalloc(&p.trace, (size_t)grid * kFuTraceSlots * sizeof(long long)));
    TFGNN_CUDA(cudaMemset(p.trace, 0, (size_t)grid * kFuTraceSlots * sizeof(long long)));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kFuThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)(split ? 2 : ctas);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define TFGNN_FU_LAUNCH(NV, BK)                                                                             \
  TFGNN_CUDA(split ? cudaLaunchKernelEx(&cfg, fused_rgcn_kernel<NV, BK, 1, true>, map_a, map_b, p)          \
             : ctas == 2 ? cudaLaunchKernelEx(&cfg, fused_rgcn_kernel<NV, BK, 2>, map_a, map_b, p)          \
                         : cudaLaunchKernelEx(&cfg, fused_rgcn_kernel<NV, BK, 1>, map_a, map_b, p))
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
  if (p.trace) {   // debug only: synchronous dump of the LAST launch's stamps (one binary file, grid x kFuTraceSlots int64)
    std::vector<long long> host((size_t)grid * kFuTraceSlots);
    TFGNN_CUDA(cudaStreamSynchronize(st));
    TFGNN_CUDA(cudaMemcpy(host.data(), p.trace, host.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    if (FILE* f = fopen(getenv("TFGNN_B200_FUSED_TRACE"), "wb")) {
      const long long hdr[4] = {grid, kFuTraceSlots, split ? 1 : 0, ctas};
      fwrite(hdr, sizeof(long long), 4, f);
      fwrite(host.data(), sizeof(long long), host.size(), f);
      fclose(f);
    }
  }
  return 0;
}

}  // namespace tfgnn

// Node-level GEMM on the 5th-generation tensor cores with fp32-level accuracy ("3xTF32"):
//     C[M,N] = epi( A[M,K] * B[K,N] ),   x = x_hi + x_lo (tf32 split),
//     A*B ~= A_hi*B_hi + (A_lo*B_hi + A_hi*B_lo)     (dropped term ~2^-22 relative)
// so the 1e-5 parity bar of the fp32 reference holds (plain TF32 would miss it by 100x) while the
// contraction runs at tensor-core rate instead of the ~70 TFLOP/s FFMA ceiling.
// Measured on B200: the tensor core adds each K=8 product into the fp32 accumulator with
// truncation, a bias that grows with the number of accumulate steps (9e-6 relative at K=1024 with
// one accumulator).  The two small correction products therefore go to their OWN TMEM accumulator
// (their truncation error is 2^-11 smaller) and are added to the main one in the epilogue with
// round-to-nearest: 3x fewer biased steps on the large-magnitude sum.
//
// Persistent warp-specialised kernel, one CTA per SM, 128 x BLOCK_N output tiles, K in 32-float
// (128 B = one swizzle row) blocks through an mbarrier ring:
//   warp 0      TMA producer: raw fp32 A tile + pre-split B_hi / B_lo tiles (K-major, SWIZZLE_128B)
//   warps 2-5   splitters: A tile -> A_hi (in place) and A_lo (position-preserving, so layout-agnostic)
//   warp 1      MMA issuer: 12 tcgen05.mma.kind::tf32 per K block, fp32 accumulators in TMEM (2 stages)
//   warps 6-13  epilogue, two groups of four: tcgen05.ld -> row normalisation / bias / activation (or the chained / GRU /
//               RGAT-score forms) -> global, each thread its own row's 64 B pieces straight from registers
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "gemm.cuh"
#include "sm100_ptx.cuh"

namespace tfgnn {

constexpr int kTcBM = 128;
constexpr int kTcBK = 32;                       // floats per K block = 128 bytes
constexpr int kTcATileBytes = kTcBM * 128;      // 16 KiB
constexpr int kTcThreads = 448;               // TMA, MMA, 4 splitter warps, 8 epilogue warps
constexpr int kTcTmemCols = 512;
constexpr int kTcAccStride = 256;               // TMEM columns per accumulator stage (main + correction)
constexpr int kTcSmemLimit = 227 * 1024;

struct TcParams {
  long long M;
  int N, K;
  int block_n, n_tiles;
  long long m_tiles, total_tiles;
  int num_k_blocks, num_stages;
  int corr_bf16;   // 1: corrections as one bf16-pair MMA (sm100_ptx.cuh), 0: two tf32 MMAs (round-1 scheme)
  int prefetch_a;  // 1: L2-prefetch the A tile of this CTA's NEXT output tile while the current one is loaded
  float* C;
  int ldc;
  // K blocks [0, kb_split) come from map_a, the rest from map_a2 (A = [A1 | A2] without materialising the concatenation)
  int kb_split;
  int store_quad;     // epilogue stores: 1 = 64 B per row and instruction (lane quads), 0 = 32 B (lane pairs)
  long long* trace;   // debug (TFGNN_B200_GEMM_TRACE=file): kTcTraceSlots clock64 stamps per CTA, see tc_trace()
  // GRU mode (launch_gemm_tc_gru): the N tile of 128 columns holds the pre-activations [z | r | x_h | h_h] of 32 hidden units;
  // the epilogue applies the Keras GRUCell gate math (reset_after) and writes the 32 new states of each row.
  int gru;
  const float* gru_h;   // previous state rows [M, gru_ldh]
  int gru_ldh;
  const float* gru_bias;   // [N] in the tile layout: b0_z + b1_z | b0_r + b1_r | b0_h | b1_h per 32 units
  GemmEpilogue epi;
};

// Debug timeline of the first kTcTraceKb K blocks of a CTA (SM-local clock64): per K block i
//   4i+0 TMA issued, 4i+1 splitter saw the bytes, 4i+2 MMA thread saw bytes + split and issues, 4i+3 TMA saw the stage free again
// (slot of the block that reuses the stage); 240+ : 240 kernel entry, 241 first tile's epilogue start, 242 its end, 243 exit.
constexpr int kTcTraceSlots = 256;
constexpr int kTcTraceKb = 60;
__device__ __forceinline__ void tc_trace(const TcParams& p, int idx) {
  if (p.trace) p.trace[(size_t)blockIdx.x * kTcTraceSlots + idx] = clock64();
}

// Epilogue stores straight from registers, full 32 B sectors per instruction: every thread holds NF4 consecutive float4 of ITS
// row; lanes 2i / 2i+1 swap every second float4, so that in each store instruction the pair writes 32 contiguous bytes of one
// of its two rows.  (Each lane storing its own 16 B pieces wrote half sectors: the wide-output shapes lost 10-14 %,
// [500k,128]x[128,384] 0.456 -> 0.507 ms, gpurun r2v.)  Must be called by all 32 lanes.
template <int NF4>
__device__ __forceinline__ void tc_store_pairwise(float* own, long long ld, const float (&v)[4 * NF4], long long row,
                                                  long long M, int lane) {
  static_assert(NF4 % 2 == 0, "pairs of float4");
  const bool odd = lane & 1;
  float* pe = odd ? own - ld : own;            // row of the even lane of the pair
  float* po = odd ? own : own + ld;            // row of the odd lane
  const bool ok_e = (odd ? row - 1 : row) < M, ok_o = (odd ? row : row + 1) < M;
  const int off = odd ? 4 : 0;
#pragma unroll
  for (int j = 0; j < NF4; j += 2) {
    // even lane gives its float4 j+1 and gets the partner's float4 j; odd lane the other way round
    float sx = odd ? v[4 * j] : v[4 * j + 4], sy = odd ? v[4 * j + 1] : v[4 * j + 5];
    float sz = odd ? v[4 * j + 2] : v[4 * j + 6], sw = odd ? v[4 * j + 3] : v[4 * j + 7];
    const float rx = __shfl_xor_sync(0xffffffffu, sx, 1), ry = __shfl_xor_sync(0xffffffffu, sy, 1);
    const float rz = __shfl_xor_sync(0xffffffffu, sz, 1), rw = __shfl_xor_sync(0xffffffffu, sw, 1);
    const float4 mine = odd ? make_float4(v[4 * j + 4], v[4 * j + 5], v[4 * j + 6], v[4 * j + 7])
                            : make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    const float4 got = make_float4(rx, ry, rz, rw);
    if (ok_e) *reinterpret_cast<float4*>(pe + 4 * j + off) = odd ? got : mine;
    if (ok_o) *reinterpret_cast<float4*>(po + 4 * j + off) = odd ? mine : got;
  }
}

// The same idea over lane QUADS for the 16-column chunks: a 4x4 transpose of float4 in two shuffle rounds (xor 1, xor 2), after
// which lane t of a quad holds float4 number t of all four rows of the quad - every store instruction then writes 64 contiguous
// bytes per row (8 rows per instruction), half as many L2 requests as the pairwise form.  TFGNN_B200_GEMM_STORE=2 (default) /
// 1 (pairwise): wide outputs ([1M,256]x[256,1024]) measured 3.09 ms with the old 128 B staging-tile stores, 3.36 pairwise.
__device__ __forceinline__ void tc_store_quadwise(float* own, long long ld, const float (&v)[16], long long row, long long M,
                                                  int lane) {
  const int t = lane & 3;
  float4 C0, C1, C2, C3;
  ptx::quad_transpose_f4(v, lane, C0, C1, C2, C3);
  float* base = own - (long long)t * ld + 4 * t;     // row of the quad's lane 0, this lane's 16-byte column
  const long long r_base = row - t;
  if (r_base < M) *reinterpret_cast<float4*>(base) = C0;
  if (r_base + 1 < M) *reinterpret_cast<float4*>(base + ld) = C1;
  if (r_base + 2 < M) *reinterpret_cast<float4*>(base + 2 * ld) = C2;
  if (r_base + 3 < M) *reinterpret_cast<float4*>(base + 3 * ld) = C3;
}

__device__ __forceinline__ float tc_row_norm(const GemmEpilogue& e, long long row) {
  if (e.row_norm == 0) return 1.0f;
  int cnt = 0;
  for (int l = 0; l < e.L; ++l) {
    const long long s = (long long)l * e.V + e.row0 + row;
    cnt += __ldg(e.row_ptr + s + 1) - __ldg(e.row_ptr + s);
  }
  const float c = (float)max(cnt, 1);
  return e.row_norm == 1 ? c : sqrtf(c);
}

__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
               const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.num_stages;
  const int b_tile_bytes = p.block_n * 128;
  const int stage_bytes = 2 * kTcATileBytes + 2 * b_tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
  uint64_t* full = bars;                 // TMA bytes landed
  uint64_t* split = bars + S;            // A_hi / A_lo written
  uint64_t* empty = bars + 2 * S;        // MMAs reading the stage retired
  uint64_t* tmem_full = bars + 3 * S;    // accumulator complete
  uint64_t* tmem_empty = bars + 3 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) tc_trace(p, 240);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_a2);
    ptx::prefetch_tensormap(&map_b);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&split[s], 128);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], p.block_n <= 128 ? 128 : 256);   // one group per stage / both groups on the one stage
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, kTcTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int m0 = (int)((tile / p.n_tiles) * kTcBM);
        const int n0 = (int)(tile % p.n_tiles) * p.block_n;
        for (int kb = 0; kb < p.num_k_blocks; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          if (it < kTcTraceKb) { tc_trace(p, 4 * it + 3); tc_trace(p, 4 * it); }
          uint8_t* st = smem + (size_t)s * stage_bytes;
          ptx::mbar_arrive_expect_tx(&full[s], kTcATileBytes + 2 * b_tile_bytes);
          if (p.prefetch_a) {
            // The pipeline holds only 3 stages of 50-64 KB, so its k-block rate is (stages / TMA latency): ~1 us per k-block
            // when A comes from HBM under load.  The next tile's A lines are requested into L2 now, ~10 k-blocks early.
            const long long nt = tile + gridDim.x;
            if (nt < p.total_tiles && (nt / p.n_tiles) != (tile / p.n_tiles) && kb < p.kb_split)
              ptx::tma_prefetch_2d(&map_a, kb * kTcBK, (int)((nt / p.n_tiles) * kTcBM));
          }
          if (kb < p.kb_split) ptx::tma_load_2d(st, &map_a, &full[s], kb * kTcBK, m0);
          else ptx::tma_load_2d(st, &map_a2, &full[s], (kb - p.kb_split) * kTcBK, m0);
          ptx::tma_load_2d(st + 2 * kTcATileBytes, &map_b, &full[s], kb * kTcBK, n0);
          ptx::tma_load_2d(st + 2 * kTcATileBytes + b_tile_bytes, &map_b, &full[s], kb * kTcBK, p.N + n0);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = ptx::umma_idesc_tf32_m128((uint32_t)p.block_n);
    const uint32_t idesc_bf = ptx::umma_idesc_bf16(128u, (uint32_t)p.block_n);
    uint32_t it = 0, tile_count = 0;
    const uint32_t n_acc = p.block_n <= 128 ? 2 : 1;       // accumulator stages that fit 512 TMEM columns
    const uint32_t corr_off = p.block_n <= 128 ? 128 : 256;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tile_count) {
      const uint32_t acc = tile_count % n_acc, acc_ph = (tile_count / n_acc) & 1;
      ptx::mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * kTcAccStride;
      const uint32_t c_tmem = d_tmem + corr_off;
      for (int kb = 0; kb < p.num_k_blocks; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::mbar_wait(&split[s], ph);
        ptx::tc_fence_after_sync();
        if (lane == 0) {
          if (it < kTcTraceKb) tc_trace(p, 4 * it + 2);
          const uint32_t st = ptx::smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t a_hi = ptx::umma_desc_k_sw128(st);
          const uint64_t a_lo = ptx::umma_desc_k_sw128(st + kTcATileBytes);
          const uint64_t b_hi = ptx::umma_desc_k_sw128(st + 2 * kTcATileBytes);
          const uint64_t b_lo = ptx::umma_desc_k_sw128(st + 2 * kTcATileBytes + b_tile_bytes);
#pragma unroll
          for (int k = 0; k < kTcBK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);  // 8 tf32 (or 16 bf16) = 32 B along K inside the swizzle row
            if (p.corr_bf16) {
              ptx::mma_bf16_ss(c_tmem, a_lo + adv, b_lo + adv, idesc_bf, (kb | k) != 0);   // pair tiles: a lo(b) + lo(a) b
            } else {
              ptx::mma_tf32_ss(c_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
              ptx::mma_tf32_ss(c_tmem, a_hi + adv, b_lo + adv, idesc, 1);
            }
            ptx::mma_tf32_ss(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
          }
          ptx::mma_commit(&empty[s]);
          if (kb == p.num_k_blocks - 1) ptx::mma_commit(&tmem_full[acc]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ================= A splitters (128 threads) =================
    const int tid = threadIdx.x - 64;
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.num_k_blocks; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        ptx::mbar_wait(&full[s], ph);
        if (tid == 0 && it < kTcTraceKb) tc_trace(p, 4 * it + 1);
        // All eight 16-byte loads first, through explicit shared-space instructions: with generic pointers the compiler
        // could not hoist a load above the previous iteration's stores (possible aliasing), so every thread paid one
        // shared-memory round trip per 16 bytes: ~20 % of all stall samples sat on the LOP3 waiting for its LD (ncu r2c).
        const uint32_t a_s = ptx::smem_u32(smem + (size_t)s * stage_bytes) + (uint32_t)tid * 16u;
        const uint32_t lo_s = a_s + kTcATileBytes;
        constexpr int kIt = kTcATileBytes / 16 / 128;
        float4 xs[kIt];
#pragma unroll
        for (int i = 0; i < kIt; ++i) xs[i] = ptx::lds_f4(a_s + (uint32_t)i * 2048u);
#pragma unroll
        for (int i = 0; i < kIt; ++i) {
          const float4 x = xs[i];
          float4 h, l;
          h.x = ptx::tf32_hi(x.x); h.y = ptx::tf32_hi(x.y); h.z = ptx::tf32_hi(x.z); h.w = ptx::tf32_hi(x.w);
          if (p.corr_bf16) {
            // the raw fp32 tile stays in place as the main operand (kind::tf32 ignores the low 13 mantissa bits:
            // tools/exp_rawhi.py); one tile of bf16 pairs (a | a - tf32(a)) feeds the correction MMA
            uint4 w;
            w.x = ptx::pack_bf16x2(x.x - h.x, x.x); w.y = ptx::pack_bf16x2(x.y - h.y, x.y);
            w.z = ptx::pack_bf16x2(x.z - h.z, x.z); w.w = ptx::pack_bf16x2(x.w - h.w, x.w);
            ptx::sts_u4(lo_s + (uint32_t)i * 2048u, w);
          } else {
            // the raw fp32 tile stays in place as the hi operand here too (bitwise identical results) - the
            // shared-memory LSU pipe is the busiest unit of this kernel (ncu r2c: 67 % at [2M,320]x[320,320])
            l.x = ptx::tf32_hi(x.x - h.x); l.y = ptx::tf32_hi(x.y - h.y);
            l.z = ptx::tf32_hi(x.z - h.z); l.w = ptx::tf32_hi(x.w - h.w);
            ptx::sts_f4(lo_s + (uint32_t)i * 2048u, l);
          }
        }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&split[s]);
      }
    }
  } else {
    // ================= epilogue: warps 6..13, two groups of four (TMEM lane quarter = warp & 3) =================
    // Two accumulator stages (block_n <= 128): group g drains the tiles that use stage g, so a tile's drain may take as long as
    // two K loops (the GRU gate epilogue takes 8 us against 6.4 us of K loop: gpurun r2u trace).  One stage: the next tile's
    // MMAs wait for the drain, and the groups split its columns.  Every thread owns one row and stores its 64 B pieces
    // straight from registers: no staging tile (its 18 KB buy a 4th pipeline stage at 80-column tiles), no warp barriers.
    const int q = warp & 3;
    const int grp = (warp - 6) >> 2;
    uint32_t tile_count = 0;
    const uint32_t n_acc = p.block_n <= 128 ? 2 : 1;
    const uint32_t corr_off = p.block_n <= 128 ? 128 : 256;
    const int half_cols = ((p.block_n / 16 + 1) / 2) * 16;
    const int col_lo = n_acc == 2 ? 0 : (grp == 0 ? 0 : half_cols);
    const int col_hi = n_acc == 2 ? p.block_n : (grp == 0 ? half_cols : p.block_n);
    const bool chained = p.epi.mul != nullptr || p.epi.accumulate || !p.epi.finalize;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tile_count) {
      if (n_acc == 2 && (int)(tile_count & 1) != grp) continue;
      const long long m0 = (tile / p.n_tiles) * kTcBM;
      const int n0 = (int)(tile % p.n_tiles) * p.block_n;
      const uint32_t acc = tile_count % n_acc, acc_ph = (tile_count / n_acc) & 1;
      ptx::mbar_wait(&tmem_full[acc], acc_ph);
      ptx::tc_fence_after_sync();
      if (tile_count == 0 && warp == 6 && lane == 0) tc_trace(p, 241);
      const long long row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const float inv_rn = row_ok ? 1.0f / tc_row_norm(p.epi, row) : 1.0f;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kTcAccStride;
      if (p.gru) {
        // Keras GRUCell, reset_after (ggnn.py:84-87) on the 32 hidden units of this tile, 8 at a time, straight from TMEM:
        //   z = sigmoid(acc_z + bz), r = sigmoid(acc_r + br), hh = tanh(acc_x + bx + r * (acc_h + bh)), h' = z h + (1 - z) hh
        // (acc_z / acc_r already hold agg K + h U: one contraction over K = [agg | h]); each thread stores its row's 32 B pieces.
        const int u0 = (int)(tile % p.n_tiles) * 32;
        const float* hrow = p.gru_h + row * p.gru_ldh + u0;
        float* orow = p.C + row * p.ldc + u0;
        for (int g = 0; g < 32; g += 8) {
          uint32_t m[4][8], c[4][8];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            ptx::tmem_ld_x8_nowait(taddr + 32 * t + g, m[t]);
            ptx::tmem_ld_x8_nowait(taddr + corr_off + 32 * t + g, c[t]);
          }
          ptx::tmem_wait_ld();
          float pre[4][8];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float4* bp = reinterpret_cast<const float4*>(p.gru_bias + n0 + 32 * t + g);
            const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) pre[t][j] = (__uint_as_float(m[t][j]) + __uint_as_float(c[t][j])) + bb[j];
          }
          float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (row_ok) {
            const float4 h0 = __ldg(reinterpret_cast<const float4*>(hrow + g));
            const float4 h1 = __ldg(reinterpret_cast<const float4*>(hrow + g + 4));
            const float hp[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float z = 1.0f / (1.0f + expf(-pre[0][j]));
              const float r = 1.0f / (1.0f + expf(-pre[1][j]));
              const float hh = tanhf(pre[2][j] + r * pre[3][j]);
              o[j] = z * hp[j] + (1.0f - z) * hh;
            }
          }
          tc_store_pairwise<2>(orow + g, p.ldc, o, row, p.M, lane);
        }
        ptx::tc_fence_before_sync();
        if (tile_count == 0 && warp == 6 && lane == 0) tc_trace(p, 242);
        ptx::mbar_arrive(&tmem_empty[acc]);
        continue;
      }
      float sc_s = 0.f, sc_t = 0.f;   // RGAT score accumulators of the current head (epi.score_src)
      float* crow = p.C + row * p.ldc + n0;
      const float* mrow = p.epi.mul ? p.epi.mul + row * p.epi.ldm + n0 : nullptr;
      for (int col = col_lo; col < col_hi; col += 16) {
        uint32_t mv[16], cv[16];
        ptx::tmem_ld_x16_nowait(taddr + col, mv);
        ptx::tmem_ld_x16_nowait(taddr + corr_off + col, cv);
        ptx::tmem_wait_ld();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(mv[j]) + __uint_as_float(cv[j]);
        if (p.epi.score_src) {
          // RGAT: this row's attention score halves, head by head, while its 16-column chunks pass (a chunk lies inside one
          // head: d % 16 == 0; a tile - and each group's column range - holds whole heads, visited in ascending column order)
          const int gc = n0 + col;
          const int l = gc / p.epi.score_H, ct = gc - l * p.epi.score_H;
          const int k = ct / p.epi.score_d, i0 = ct - k * p.epi.score_d;
          const float* a = reinterpret_cast<const float*>(p.epi.score_att.p[l]) + (size_t)k * 2 * p.epi.score_d + i0;
          if (i0 == 0) { sc_s = 0.f; sc_t = 0.f; }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 as = __ldg(reinterpret_cast<const float4*>(a) + j);
            const float4 at = __ldg(reinterpret_cast<const float4*>(a + p.epi.score_d) + j);
            sc_s = fmaf(as.x, v[4 * j], sc_s); sc_s = fmaf(as.y, v[4 * j + 1], sc_s);
            sc_s = fmaf(as.z, v[4 * j + 2], sc_s); sc_s = fmaf(as.w, v[4 * j + 3], sc_s);
            sc_t = fmaf(at.x, v[4 * j], sc_t); sc_t = fmaf(at.y, v[4 * j + 1], sc_t);
            sc_t = fmaf(at.z, v[4 * j + 2], sc_t); sc_t = fmaf(at.w, v[4 * j + 3], sc_t);
          }
          if (i0 + 16 == p.epi.score_d && row_ok) {
            const long long o = row * (long long)(p.N / p.epi.score_d) + (long long)l * p.epi.score_K + k;
            p.epi.score_src[o] = sc_s;
            p.epi.score_tgt[o] = sc_t;
          }
        }
        if (chained && row_ok) {
          // FiLM-style chained contractions: val = (accumulate ? C_old : 0) + (mul ? mul * acc : acc)
          if (mrow) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 m = __ldg(reinterpret_cast<const float4*>(mrow + col) + j);
              v[4 * j] *= m.x; v[4 * j + 1] *= m.y; v[4 * j + 2] *= m.z; v[4 * j + 3] *= m.w;
            }
          }
          if (p.epi.accumulate) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 c = *(reinterpret_cast<const float4*>(crow + col) + j);
              v[4 * j] += c.x; v[4 * j + 1] += c.y; v[4 * j + 2] += c.z; v[4 * j + 3] += c.w;
            }
          }
        }
        if (!chained || p.epi.finalize) {
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
        }
        if (p.store_quad) tc_store_quadwise(crow + col, p.ldc, v, row, p.M, lane);
        else tc_store_pairwise<4>(crow + col, p.ldc, v, row, p.M, lane);
      }
      ptx::tc_fence_before_sync();
      if (tile_count == 0 && warp == 6 && lane == 0) tc_trace(p, 242);
      ptx::mbar_arrive(&tmem_empty[acc]);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) tc_trace(p, 243);
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTcTmemCols);
  }
}

// Correction scheme of the 3xTF32 contractions (rows [N, 2N) of the packed weights differ between the two):
//   0 = two tf32 MMAs  A_lo B_hi + A_hi B_lo   (round 1; ~3e-7 of the pre-activation scale in exact accumulation)
//   1 = ONE bf16-pair MMA (sm100_ptx.cuh)      (-33 % tensor work; ~1.5e-6 of the pre-activation scale)
// Measured on B200 (gpurun r2a): the fused RGCN kernel gains 5 % from the pair scheme (cfg2 4.59 -> 4.37 ms) and its
// error at K = L*D >= 768 is dominated by the tensor core's truncating accumulate either way (3.56e-6 vs 3.0e-6 of
// max|out| at cfg2).  The generic GEMM is latency-bound at its shapes, gains nothing, and feeds saturating
// activations (tanh / sigmoid of Dense layers and GRU gates) that turn an error relative to the PRE-activation scale
// into one relative to 1: with the pair scheme five parity tests landed at 1.2-2.1e-5.  So: the GEMM keeps scheme 0;
// the fused kernel uses scheme 1 unless its activation is tanh.  TFGNN_B200_CORR_BF16 = 0/1 forces the FUSED kernel's
// scheme, TFGNN_B200_CORR_BF16_GEMM = 1 the GEMM's (experiments).
int gemm_corr_bf16() {
  static const int v = [] { const char* e = getenv("TFGNN_B200_CORR_BF16_GEMM"); return e && atoi(e) != 0 ? 1 : 0; }();
  return v;
}
int fused_corr_bf16(int activation) {
  static const int v = [] { const char* e = getenv("TFGNN_B200_CORR_BF16"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
  if (v >= 0) return v;
  return activation == TFGNN_ACT_TANH ? 0 : 1;
}

// B [K,N] row-major -> packed [2N, Kp] K-major: rows [0,N) = tf32 hi of B^T, rows [N,2N) = the correction operand:
// tf32 lo (two-MMA scheme) or the bf16 pair (b - tf32(b) | b) (pair scheme, sm100_ptx.cuh).
__global__ void pack_weights_tc_kernel(const float* __restrict__ B, int ldb, int K, int N, int Kp, int corr_bf16,
                                       float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * Kp;
  if (idx >= total) return;
  const int k = (int)(idx % Kp);
  const int n = (int)(idx / Kp);
  const float x = k < K ? __ldg(B + (long long)k * ldb + n) : 0.f;
  const float h = ptx::tf32_hi(x);
  out[idx] = h;
  if (corr_bf16) reinterpret_cast<uint32_t*>(out)[total + idx] = ptx::pack_bf16x2(x, x - h);   // low half: lo(b), high half: b
  else out[total + idx] = ptx::tf32_hi(x - h);
}

// The same packing straight from the L per-type matrices W_l [D, H] stacked vertically (K = L*D): one launch instead
// of pack_vertical + pack_weights_tc on the per-layer path of the fused kernel.
__global__ void pack_weights_tc_table_kernel(PtrTable W, int L, int D, int H, int Kp, int corr_bf16,
                                             float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)H * Kp;
  if (idx >= total) return;
  const int k = (int)(idx % Kp);
  const int n = (int)(idx / Kp);
  float x = 0.f;
  if (k < L * D) {
    const int l = k / D, d = k - l * D;
    x = __ldg(reinterpret_cast<const float*>(W.p[l]) + (long long)d * H + n);
  }
  const float h = ptx::tf32_hi(x);
  out[idx] = h;
  if (corr_bf16) reinterpret_cast<uint32_t*>(out)[total + idx] = ptx::pack_bf16x2(x, x - h);
  else out[total + idx] = ptx::tf32_hi(x - h);
}

// GRU weights in the tile layout of the gate epilogue: output column n = 128 t + 32 gate + j belongs to hidden unit
// u = 32 t + j; rows k < H of the [2H, 4H] operand multiply the aggregated messages (GRUCell.kernel [H, 3H], gate order
// z | r | h as in Keras), rows k >= H the previous state (recurrent_kernel): gate 2 ("x_h") has no recurrent part, gate 3
// ("h_h") no input part.  Written directly in the packed K-major hi / lo form; the bias vector in the same column order.
__global__ void pack_gru_weights_kernel(const float* __restrict__ Kx, const float* __restrict__ Uh,
                                        const float* __restrict__ bias, int H, int Kp, float* __restrict__ out,
                                        float* __restrict__ bias_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = 4 * H;
  const long long total = (long long)N * Kp;
  if (idx >= total) return;
  const int k = (int)(idx % Kp);
  const int n = (int)(idx / Kp);
  const int t = n / 128, gate = (n % 128) / 32, u = 32 * t + (n % 32);
  float x = 0.f;
  if (k < H) {
    if (gate < 3) x = __ldg(Kx + (long long)k * 3 * H + gate * H + u);
  } else if (k < 2 * H) {
    const int gsrc = gate == 3 ? 2 : gate;
    if (gate != 2) x = __ldg(Uh + (long long)(k - H) * 3 * H + gsrc * H + u);
  }
  const float h = ptx::tf32_hi(x);
  out[idx] = h;
  out[total + idx] = ptx::tf32_hi(x - h);
  if (k == 0) {
    const float* b0 = bias;           // input bias  [3H]
    const float* b1 = bias + 3 * H;   // recurrent bias [3H]
    bias_out[n] = gate == 0 ? b0[u] + b1[u] : gate == 1 ? b0[H + u] + b1[H + u] : gate == 2 ? b0[2 * H + u] : b1[2 * H + u];
  }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

static int pick_block_n(int N) {
  if (N % 16 != 0 || N < 16) return 0;
  if (N <= 128) return N;
  // <=128 columns: main + correction accumulators fit twice in the 512 TMEM columns, so the epilogue of
  // one tile overlaps the MMAs of the next; the A tile is then read once per N tile (L2 hit).
  for (int bn = 128; bn >= 64; bn -= 16)
    if (N % bn == 0) return bn;
  if (N <= 256) return N;
  for (int bn = 256; bn > 128; bn -= 16)
    if (N % bn == 0) return bn;
  return 0;
}

static bool device_is_sm100() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cached = (major == 10) ? 1 : 0;
  }
  return cached == 1;
}

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

bool gemm_tc_supported(long long M, int N, int K, const float* A, int lda, const float* C, int ldc) {
  if (M < 1 || K < 4 || K % 4 != 0 || lda % 4 != 0 || ldc % 4 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(C)) & 15) return false;
  if (pick_block_n(N) == 0) return false;
  if (M > (1ll << 31) - 256) return false;
  return device_is_sm100() && get_encode_fn() != nullptr;
}

bool gemm_tc_scores_supported(int N, int H, int d) {
  const int bn = pick_block_n(N);
  if (!(bn > 0 && d % 16 == 0 && H % d == 0 && bn % d == 0 && N % H == 0 && (bn % H == 0 || H % bn == 0))) return false;
  return bn <= 128 || (((bn / 16 + 1) / 2) * 16) % d == 0;   // one accumulator stage: the two epilogue groups split the columns
}
size_t gemm_tc_packed_bytes(int N, int K) { return (size_t)2 * N * round_up(K, kTcBK) * sizeof(float); }

int launch_pack_weights_tc_table(const PtrTable& W, int L, int D, int H, int corr_bf16, float* packed, cudaStream_t st) {
  const int Kp = round_up(L * D, kTcBK);
  const long long total = (long long)H * Kp;
  pack_weights_tc_table_kernel<<<ceil_div(total, 256), 256, 0, st>>>(W, L, D, H, Kp, corr_bf16, packed);
  TFGNN_LAUNCH_CHECK();
  return 0;
}

int launch_pack_weights_tc(const float* B, int ldb, int K, int N, float* packed, cudaStream_t st) {
  const int Kp = round_up(K, kTcBK);
  const long long total = (long long)N * Kp;
  pack_weights_tc_kernel<<<ceil_div(total, 256), 256, 0, st>>>(B, ldb, K, N, Kp, gemm_corr_bf16(), packed);
  TFGNN_LAUNCH_CHECK();
  return 0;
}

struct TcExtra {             // optional second A operand / GRU epilogue (launch_gemm_tc_gru)
  const float* A2 = nullptr;
  int lda2 = 0, K1 = 0;      // A = [A1 (K1 columns) | A2 (K - K1 columns)]
  const float* gru_h = nullptr;
  int gru_ldh = 0;
  const float* gru_bias = nullptr;
};

static int launch_gemm_tc_impl(const float* A, int lda, const float* packedB, float* C, int ldc, long long M, int N, int K,
                               const GemmEpilogue& epi, const TcExtra& ex, cudaStream_t st) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) {
    set_error(TFGNN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    return TFGNN_ERR_CUDA;
  }
  const int Kp = round_up(K, kTcBK);
  TcParams p{};
  p.M = M; p.N = N; p.K = K;
  p.block_n = ex.gru_h ? 128 : pick_block_n(N);
  p.n_tiles = N / p.block_n;
  p.m_tiles = (M + kTcBM - 1) / kTcBM;
  p.total_tiles = p.m_tiles * p.n_tiles;
  p.num_k_blocks = Kp / kTcBK;
  const int stage_bytes = 2 * kTcATileBytes + 2 * p.block_n * 128;
  int stages = (kTcSmemLimit - 2048) / stage_bytes;
  if (stages > 4) stages = 4;
  TFGNN_REQUIRE(stages >= 2, "tcgen05 GEMM: tile does not fit shared memory");
  p.num_stages = stages;
  p.corr_bf16 = gemm_corr_bf16();
  {
    const char* e = getenv("TFGNN_B200_GEMM_PREFETCH");   // read per call (A/B experiments)
    p.prefetch_a = e ? (atoi(e) != 0) : 1;
  }
  p.C = C; p.ldc = ldc; p.epi = epi;
  const int K1 = ex.A2 ? ex.K1 : K;
  p.kb_split = ex.A2 ? K1 / kTcBK : p.num_k_blocks;
  p.gru = ex.gru_h ? 1 : 0;
  {
    const char* e = getenv("TFGNN_B200_GEMM_STORE");   // read per call (A/B experiments)
    p.store_quad = e ? (atoi(e) == 2) : 1;
  }
  if (p.gru) p.corr_bf16 = 0;   // pack_gru_weights_kernel writes the two-MMA (tf32 lo) correction operand
  p.gru_h = ex.gru_h; p.gru_ldh = ex.gru_ldh; p.gru_bias = ex.gru_bias;

  CUtensorMap map_a, map_a2, map_b;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K1, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)lda * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)kTcBM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error(TFGNN_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed with code " + std::to_string((int)r));
      return TFGNN_ERR_CUDA;
    }
  }
  map_a2 = map_a;
  if (ex.A2) {
    cuuint64_t dims[2] = {(cuuint64_t)(K - K1), (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)ex.lda2 * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)kTcBM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map_a2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ex.A2), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error(TFGNN_ERR_CUDA, "cuTensorMapEncodeTiled(A2) failed with code " + std::to_string((int)r));
      return TFGNN_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)(2 * N)};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)p.block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(packedB), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error(TFGNN_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed with code " + std::to_string((int)r));
      return TFGNN_ERR_CUDA;
    }
  }
  const size_t smem_bytes = (size_t)stages * stage_bytes + (3 * stages + 8) * sizeof(uint64_t) + 1024;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemLimit);
  });
  TFGNN_CUDA(attr_err);
  int dev = 0, sms = 148;
  TFGNN_CUDA(cudaGetDevice(&dev));
  TFGNN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = (int)(p.total_tiles < sms ? p.total_tiles : sms);
  p.trace = nullptr;
  const char* tr = getenv("TFGNN_B200_GEMM_TRACE");
  if (tr && *tr) {
    TFGNN_CUDA(cudaMalloc(&p.trace, (size_t)grid * kTcTraceSlots * sizeof(long long)));
    TFGNN_CUDA(cudaMemset(p.trace, 0, (size_t)grid * kTcTraceSlots * sizeof(long long)));
  }
  gemm_tc_kernel<<<grid, kTcThreads, smem_bytes, st>>>(map_a, map_a2, map_b, p);
  TFGNN_LAUNCH_CHECK();
  if (p.trace) {   // debug only: synchronous dump of the LAST launch (grid x kTcTraceSlots int64 after a 4-word header)
    std::vector<long long> host((size_t)grid * kTcTraceSlots);
    TFGNN_CUDA(cudaStreamSynchronize(st));
    TFGNN_CUDA(cudaMemcpy(host.data(), p.trace, host.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    if (FILE* f = fopen(tr, "wb")) {
      const long long hdr[4] = {grid, kTcTraceSlots, p.num_k_blocks, p.num_stages};
      fwrite(hdr, sizeof(long long), 4, f);
      fwrite(host.data(), sizeof(long long), host.size(), f);
      fclose(f);
    }
  }
  return 0;
}

int launch_gemm_tc(const float* A, int lda, const float* packedB, float* C, int ldc, long long M, int N, int K,
                   const GemmEpilogue& epi, cudaStream_t st) {
  return launch_gemm_tc_impl(A, lda, packedB, C, ldc, M, N, K, epi, TcExtra{}, st);
}

// The whole GRU update of GGNN as ONE contraction with a gate epilogue (ggnn.py:84-87):
//   [agg | h] [2H] x W_gru [2H, 4H]  ->  per hidden unit z, r, x_h, h_h pre-activations  ->  h' = z h + (1 - z) tanh(x_h + r h_h)
// instead of two [V,H]x[H,3H] GEMMs writing gx / gh (2 x 768 MB at cfg4), and a gate kernel reading them back.
bool gemm_tc_gru_supported(long long V, int H, const float* agg, int lda, const float* h, int ldh, const float* out, int ldo) {
  if (H % 32 != 0 || H < 32) return false;
  if ((reinterpret_cast<uintptr_t>(h) & 15) || ldh % 4 != 0) return false;
  return gemm_tc_supported(V, 4 * H, 2 * H, agg, lda, out, ldo);
}
size_t gemm_tc_gru_packed_bytes(int H) { return gemm_tc_packed_bytes(4 * H, 2 * H) + (size_t)4 * H * sizeof(float); }
int launch_gemm_tc_gru(const float* agg, int lda, const float* h, int ldh, const float* gru_kernel,
                       const float* gru_recurrent_kernel, const float* gru_bias, float* packed, float* out, int ldo,
                       long long V, int H, cudaStream_t st) {
  const int N = 4 * H, K = 2 * H;
  const long long total = (long long)N * K;   // K is a multiple of 32 already
  float* bias_tiles = packed + 2 * total;
  pack_gru_weights_kernel<<<ceil_div(total, 256), 256, 0, st>>>(gru_kernel, gru_recurrent_kernel, gru_bias, H, K, packed,
                                                               bias_tiles);
  TFGNN_LAUNCH_CHECK();
  TcExtra ex;
  ex.A2 = h; ex.lda2 = ldh; ex.K1 = H;
  ex.gru_h = h; ex.gru_ldh = ldh; ex.gru_bias = bias_tiles;
  return launch_gemm_tc_impl(agg, lda, packed, out, ldo, V, N, K, GemmEpilogue{}, ex, st);
}

}  // namespace tfgnn

#pragma once
#include "common.cuh"

namespace tfgnn {

// Epilogue of the node-level contractions: out = act( rownorm(v) * (A B)[v,:] + bias ).
struct GemmEpilogue {
  int act = TFGNN_ACT_NONE;
  const float* bias = nullptr;   // [N] or null
  // per-row normalisation for mean / sqrt_n aggregation (counts over all edge types of the CSR)
  const int* row_ptr = nullptr;
  int V = 0, L = 0;
  long long row0 = 0;  // global node id of output row 0 (chunked launches)
  int row_norm = 0;  // 0 none, 1 mean (/max(cnt,1)), 2 sqrt_n (/sqrt(max(cnt,1)))
  // FiLM-style chained contractions (variants.cu): val = (accumulate ? C_old : 0) + (mul ? mul[row, n] * acc : acc);
  // row-norm / bias / activation are applied only when `finalize` (the last contraction of the chain).
  // LayerNormalization over the finished row, fused into the fused RGCN kernel's epilogue (gnn.py:317-321 directly after the
  // message-passing layer): y = (x - mean) * rsqrt(var + eps) * gamma + beta over all N columns; single N pass only
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  float ln_eps = 0.f;
  const float* mul = nullptr;   // [M, ldm] elementwise multiplier (gamma), rows/cols aligned with C
  int ldm = 0;
  int accumulate = 0;
  int finalize = 1;
  // RGAT (tcgen05 GEMM only): while the projected row P[v, l*H + k*d + i] passes through the epilogue, its per-head attention
  // score halves s_src[v,l,k] = a_l[k,:d] . P_l[v,k,:], s_tgt[v,l,k] = a_l[k,d:] . P_l[v,k,:] (rgat.py:111-121) are written as
  // well: no second pass over P.  Needs d % 16 == 0 and N tiles that hold whole heads (gemm_tc_scores_supported).
  float* score_src = nullptr;   // [M, L*K]
  float* score_tgt = nullptr;
  PtrTable score_att{};         // a_l [K, 2d] per type
  int score_H = 0, score_K = 0, score_d = 0;
};

// fp32 SIMT GEMM, any shape / alignment (universal fallback).  C[M,N] = epi(A[M,K] B[K,N]).
int launch_gemm_simt(const float* A, int lda, const float* B, int ldb, float* C, int ldc, long long M,
                     int N, int K, const GemmEpilogue& epi, cudaStream_t st);

// 3xTF32 tcgen05 GEMM (sm_100a tensor cores, fp32-level accuracy).  Requirements are checked by
// gemm_tc_supported(); B is given K-major-packed by pack_weights_tc (see gemm_tc.cu).
bool gemm_tc_supported(long long M, int N, int K, const float* A, int lda, const float* C, int ldc);
size_t gemm_tc_packed_bytes(int N, int K);
bool gemm_tc_scores_supported(int N, int H, int d);   // RGAT scores in the epilogue of an [M, N = L*H] projection
int launch_pack_weights_tc(const float* B, int ldb, int K, int N, float* packed, cudaStream_t st);
int launch_pack_weights_tc_table(const PtrTable& W, int L, int D, int H, int corr_bf16, float* packed, cudaStream_t st);
int launch_gemm_tc(const float* A, int lda, const float* packedB, float* C, int ldc, long long M, int N,
                   int K, const GemmEpilogue& epi, cudaStream_t st);

// GGNN's GRU update as one contraction over [agg | h] with the gate math in the epilogue (gemm_tc.cu)
bool gemm_tc_gru_supported(long long V, int H, const float* agg, int lda, const float* h, int ldh, const float* out, int ldo);
size_t gemm_tc_gru_packed_bytes(int H);
int launch_gemm_tc_gru(const float* agg, int lda, const float* h, int ldh, const float* gru_kernel,
                       const float* gru_recurrent_kernel, const float* gru_bias, float* packed, float* out, int ldo,
                       long long V, int H, cudaStream_t st);

// Fused RGCN-style layer (fused_rgcn.cu): gather -> segment-sum -> 3xTF32 tcgen05 -> epilogue in one kernel.
bool fused_rgcn_supported(long long V, int L, int D, int H, const float* h, const float* out, int ldo);
size_t fused_rgcn_ring_bytes(int D, int L, int H);
int launch_fused_rgcn(const float* h, int D, const int* row_ptr, const int* src, long long M, int V, int L, int normalize,
                      const float* packedB, int corr_bf16, int H, float* ring, float* out, int ldo, const GemmEpilogue& epi,
                      cudaStream_t st, float* const* peer_out = nullptr, int n_peer_out = 0, float* mc_out = nullptr);

int gemm_corr_bf16();                 // correction scheme of the 3xTF32 contractions (gemm_tc.cu)
int fused_corr_bf16(int activation);
int set_l2_persist_mb(int mb);
void restore_l2_persist_carveout();
void pool_trim_all();

// Weight packing helpers: gather per-type matrices into one node-level operand.
//  vertical:   dst[(blk*rows + r), :] = src_blk[row0 + r, :]          (aggregate-then-transform)
//  horizontal: dst[r, blk*cols + c]  = src_blk[row0 + r, c]           (transform-then-aggregate)
int launch_pack_vertical(const PtrTable& src, int nblk, int row0, int rows, int cols, int ld_src, float* dst,
                         int ld_dst, int dst_row0, cudaStream_t st);
int launch_pack_horizontal(const PtrTable& src, int nblk, int row0, int rows, int cols, int ld_src, float* dst,
                           int ld_dst, cudaStream_t st);

}  // namespace tfgnn

#pragma once
#include "common.cuh"
#include "edge_reduce.cuh"
#include "gemm.cuh"

namespace tfgnn {
int unsupported(const std::string& msg);
bool valid_act(int a);
bool valid_agg(int a);
int agg_row_norm(int aggregation);
int node_gemm(const float* A, int lda, const float* B, int ldb, float* C, int ldc, long long M, int N, int K,
              const GemmEpilogue& epi, int path, tfgnn_batch* batch, int tc_slot, cudaStream_t st);
int edge_mlp_core(tfgnn_batch* b, const float* h, int D, const float* const* mlp_weights, int n_hidden, int H,
                  uint32_t flags, int aggregation, int activation, int path, float* out, int ldo, cudaStream_t st);
// literal per-edge path (literal.cu); FB = optional FiLM table [V, L*2H] (gamma | beta per type)
int edge_mlp_literal(tfgnn_batch* b, const float* h, int D, const float* const* mlp_weights, int n_hidden, int H,
                     uint32_t flags, int aggregation, int activation, const float* FB, int ldf, int path, float* out,
                     int ldo, cudaStream_t st);
// RGAT edge-level aggregation (rgat.cu): warp per target + chunked hub path; needs (H/K) % 4 == 0
int launch_rgat_aggregate(tfgnn_batch* b, const float* P, const float* s_src, const float* s_tgt, int K, int d,
                          int activation, float* out, cudaStream_t st);
}  // namespace tfgnn

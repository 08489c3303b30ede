// Edge-level kernels: gather source-node rows along the (type,target)-sorted CSR and reduce
// them per target.  HBM-bound: one 8 B index pair (amortised into the 4 B sorted source id) and
// 4*C bytes of node-table row per edge; no [E,D] gather, [E,H] message or [M,H] concat buffer
// is ever materialised (the reference materialises all three: message_passing.py:197-206,
// gnn_edge_mlp.py:100, message_passing.py:166-167).
//
// One warp owns one segment (PER_TYPE mode: segment (l,v) -> out[v, l*stride + :]) or one
// target node (MERGED mode: all L segments of v reduced into out[v, :]), so the reduction is a
// register accumulation in CSR order: no atomics, run-to-run deterministic.
#include <cstdlib>

#include "edge_reduce.cuh"

namespace tfgnn {

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int NV>
struct RowAcc {
  float4 v[NV];
};

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_max(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float4 f4_scale(float4 a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 f4_fill(float s) { return make_float4(s, s, s, s); }

// Per-edge message transform on pre-projected rows (everything that is NOT a matmul in
// gnn_edge_mlp.py:84-107 / gnn_film.py:99-107 / message_passing.py:169-170).
struct EdgeFn {
  float4 t, g, b;  // target-side additive term, FiLM gamma / beta for this (v,l) and column group
  float scale;
  bool has_t, hidden_relu, scale_per_edge, film;
  int edge_act;
  __device__ __forceinline__ float4 operator()(float4 x) const {
    if (has_t) x = f4_add(x, t);
    if (hidden_relu) x = f4_max(x, f4_fill(0.f));
    if (scale_per_edge) x = f4_scale(x, scale);
    if (film) x = make_float4(g.x * x.x + b.x, g.y * x.y + b.y, g.z * x.z + b.z, g.w * x.w + b.w);
    if (edge_act != TFGNN_ACT_NONE)
      x = make_float4(apply_act(x.x, edge_act), apply_act(x.y, edge_act), apply_act(x.z, edge_act),
                      apply_act(x.w, edge_act));
    return x;
  }
};

// NV = float4 column groups per lane (C <= 128*NV).  PLAIN: identity message + sum, scale at end.
// U = edges loaded per round (loads in flight per lane = U*NV).  The lean PLAIN variant trades unroll
// depth for occupancy (<= 40 registers -> 6 CTAs/SM): the gather is bound by the number of independent
// row_ptr -> index -> row dependency chains in flight, i.e. by resident warps, not by loads per warp.
template <int NV, bool MERGED, bool PLAIN, int U = 4, int MINB = 1>
__global__ void __launch_bounds__(256, MINB) edge_reduce_kernel(const EdgeReduceParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long warp_stride = ((long long)gridDim.x * blockDim.x) >> 5;
  const int vcount = p.v_count;
  const long long num_items = MERGED ? (long long)vcount : (long long)p.L * vcount;
  const int C4 = p.C >> 2;
  const bool use_max = (!PLAIN) && p.reduce_max;
  for (long long item = warp_global; item < num_items; item += warp_stride) {
  int l_first, l_last, v;
  if (MERGED) {
    v = p.v_begin + (int)item;
    l_first = 0;
    l_last = p.L;
  } else {
    l_first = (int)(item / vcount);
    l_last = l_first + 1;
    v = p.v_begin + (int)(item - (long long)l_first * vcount);
  }
  float4 acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = f4_fill(use_max ? kLowestFloat : 0.f);
  int total_cnt = 0;


  for (int l = l_first; l < l_last; ++l) {
    const long long seg = (long long)l * p.V + v;
    const int beg = __ldg(p.row_ptr + seg), end = __ldg(p.row_ptr + seg + 1);
    const int cnt = end - beg;
    total_cnt += cnt;
    const float scale = p.normalize ? 1.0f / ((float)cnt + kSmallNumber) : 1.0f;
    const float* __restrict__ xbase = p.X + (long long)l * p.x_type_stride;
    EdgeFn fn[NV];
    if (!PLAIN) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c4 = lane + 32 * j;
        fn[j].has_t = p.T != nullptr;
        fn[j].hidden_relu = p.hidden_relu;
        fn[j].film = p.G != nullptr;
        fn[j].edge_act = p.edge_act;
        fn[j].scale = scale;
        fn[j].scale_per_edge = p.normalize && (p.G != nullptr || p.edge_act != TFGNN_ACT_NONE || use_max);
        fn[j].t = fn[j].g = fn[j].b = f4_fill(0.f);
        if (c4 < C4 && cnt > 0) {
          if (p.T) fn[j].t = ldg_f4(p.T + (long long)v * p.ldt + (long long)l * p.t_type_stride + 4 * c4);
          if (p.G) {
            const float* gp = p.G + (long long)v * p.ldg + (long long)l * p.g_type_stride + 4 * c4;
            fn[j].g = ldg_f4(gp);
            fn[j].b = ldg_f4(gp + p.beta_off);
          }
        }
      }
    }
    float4 part[NV];  // per-type partial (sum mode) so the end-of-segment scale is per type
#pragma unroll
    for (int j = 0; j < NV; ++j) part[j] = f4_fill(0.f);

    for (int base = beg; base < end; base += 32) {
      const int n = min(32, end - base);
      const int my_src = lane < n ? __ldg(p.src + base + lane) : 0;
      int e = 0;
      for (; e + U <= n; e += U) {
        float4 r[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int s = __shfl_sync(0xffffffffu, my_src, e + u);
          const float* row = xbase + (long long)s * p.ldx;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const int c4 = lane + 32 * j;
            r[u][j] = c4 < C4 ? ldg_f4(row + 4 * c4) : f4_fill(0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            if (PLAIN) {
              part[j] = f4_add(part[j], r[u][j]);
            } else {
              float4 y = fn[j](r[u][j]);
              if (use_max) acc[j] = f4_max(acc[j], y);
              else part[j] = f4_add(part[j], y);
            }
          }
      }
      for (; e < n; ++e) {
        const int s = __shfl_sync(0xffffffffu, my_src, e);
        const float* row = xbase + (long long)s * p.ldx;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c4 = lane + 32 * j;
          float4 x = c4 < C4 ? ldg_f4(row + 4 * c4) : f4_fill(0.f);
          if (PLAIN) {
            part[j] = f4_add(part[j], x);
          } else {
            float4 y = fn[j](x);
            if (use_max) acc[j] = f4_max(acc[j], y);
            else part[j] = f4_add(part[j], y);
          }
        }
      }
    }
    if (!use_max) {
      const bool scale_at_end = p.normalize && (PLAIN || !fn[0].scale_per_edge);
#pragma unroll
      for (int j = 0; j < NV; ++j)
        acc[j] = f4_add(acc[j], scale_at_end ? f4_scale(part[j], scale) : part[j]);
    }
  }

  // epilogue
  float rn = 1.0f;
  bool divide = false;
  if (MERGED && !PLAIN) {
    if (p.row_norm == 1) { rn = (float)max(total_cnt, 1); divide = true; }
    else if (p.row_norm == 2) { rn = sqrtf((float)max(total_cnt, 1)); divide = true; }
  }
  float* orow = p.out + (long long)(v - p.v_begin) * p.ldo + (MERGED ? 0 : (long long)l_first * p.out_type_stride);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c4 = lane + 32 * j;
    if (c4 < C4) {
      float4 y = acc[j];
      if (divide) y = make_float4(y.x / rn, y.y / rn, y.z / rn, y.w / rn);
      if (!PLAIN && p.final_act != TFGNN_ACT_NONE)
        y = make_float4(apply_act(y.x, p.final_act), apply_act(y.y, p.final_act),
                        apply_act(y.z, p.final_act), apply_act(y.w, p.final_act));
      *reinterpret_cast<float4*>(orow + 4 * c4) = y;
    }
  }
  }  // item loop
}

// Scalar fallback for column counts / leading dimensions that are not multiples of 4 (doctest
// sizes such as D=3, H=7): one thread per (item, column).  Same semantics, no vector loads.
template <bool MERGED>
__global__ void edge_reduce_scalar_kernel(const EdgeReduceParams p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long num_items = MERGED ? (long long)p.v_count : (long long)p.L * p.v_count;
  if (idx >= num_items * p.C) return;
  const long long item = idx / p.C;
  const int c = (int)(idx - item * p.C);
  int l_first, l_last, v;
  if (MERGED) { v = p.v_begin + (int)item; l_first = 0; l_last = p.L; }
  else { l_first = (int)(item / p.v_count); l_last = l_first + 1; v = p.v_begin + (int)(item - (long long)l_first * p.v_count); }
  const bool use_max = p.reduce_max;
  float acc = use_max ? kLowestFloat : 0.f;
  int total_cnt = 0;
  for (int l = l_first; l < l_last; ++l) {
    const long long seg = (long long)l * p.V + v;
    const int beg = p.row_ptr[seg], end = p.row_ptr[seg + 1];
    const int cnt = end - beg;
    total_cnt += cnt;
    const float scale = p.normalize ? 1.0f / ((float)cnt + kSmallNumber) : 1.0f;
    const bool film = p.G != nullptr;
    const bool scale_per_edge = p.normalize && (film || p.edge_act != TFGNN_ACT_NONE || use_max);
    float t = 0.f, g = 0.f, b = 0.f;
    if (cnt > 0) {
      if (p.T) t = p.T[(long long)v * p.ldt + (long long)l * p.t_type_stride + c];
      if (film) {
        const float* gp = p.G + (long long)v * p.ldg + (long long)l * p.g_type_stride + c;
        g = gp[0];
        b = gp[p.beta_off];
      }
    }
    float part = 0.f;
    for (int e = beg; e < end; ++e) {
      float x = p.X[(long long)p.src[e] * p.ldx + (long long)l * p.x_type_stride + c];
      if (p.T) x += t;
      if (p.hidden_relu) x = fmaxf(x, 0.f);
      if (scale_per_edge) x *= scale;
      if (film) x = g * x + b;
      x = apply_act(x, p.edge_act);
      if (use_max) acc = fmaxf(acc, x);
      else part += x;
    }
    if (!use_max) acc += (p.normalize && !scale_per_edge) ? part * scale : part;
  }
  if (MERGED) {
    if (p.row_norm == 1) acc = acc / (float)max(total_cnt, 1);
    else if (p.row_norm == 2) acc = acc / sqrtf((float)max(total_cnt, 1));
  }
  acc = apply_act(acc, p.final_act);
  p.out[(long long)(v - p.v_begin) * p.ldo + (MERGED ? 0 : (long long)l_first * p.out_type_stride) + c] = acc;
}

// Target-state term of a 0-hidden-layer edge MLP with use_target_state_as_input
// (gnn_edge_mlp.py:93-98): sum_e (h_v W^t) / (c+eps) = (c/(c+eps)) * h_v W^t, so the per-(v,l)
// input row of the node-level contraction is coeff(v,l) * h_v.
__global__ void target_term_kernel(const float* __restrict__ h, int ldh, const int* __restrict__ row_ptr,
                                   int V, int L, int D, int normalize, float* __restrict__ out, int ldo,
                                   int col0) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)V * L * D;
  if (idx >= total) return;
  const int c = (int)(idx % D);
  const long long vl = idx / D;
  const int l = (int)(vl % L);
  const int v = (int)(vl / L);
  const long long seg = (long long)l * V + v;
  const float cnt = (float)(row_ptr[seg + 1] - row_ptr[seg]);
  const float coeff = normalize ? cnt * (1.0f / (cnt + kSmallNumber)) : cnt;
  out[(long long)v * ldo + col0 + (long long)l * D + c] = coeff * h[(long long)v * ldh + c];
}

// float4 version (D, ldh, ldo, col0 multiples of 4, aligned bases): one warp per node, the node's row is read ONCE and
// written L times (the scalar kernel above re-read it L times through a div/mod per element: 17 ms for 15 GB at the
// GNN-FiLM 1/8 shard, 0.9 TB/s).
__global__ void __launch_bounds__(256) target_term_vec_kernel(const float* __restrict__ h, int ldh,
                                                              const int* __restrict__ row_ptr, int V, int L, int D,
                                                              int normalize, float* __restrict__ out, int ldo, int col0) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int C4 = D >> 2;
  for (long long v = warp; v < V; v += nwarps) {
    const float* hr = h + v * ldh;
    float* orow = out + v * ldo + col0;
    for (int c4 = lane; c4 < C4; c4 += 32) {
      const float4 x = ldg_f4(hr + 4 * c4);
      for (int l = 0; l < L; ++l) {
        const long long seg = (long long)l * V + v;
        const float cnt = (float)(__ldg(row_ptr + seg + 1) - __ldg(row_ptr + seg));
        const float coeff = normalize ? cnt * (1.0f / (cnt + kSmallNumber)) : cnt;
        *reinterpret_cast<float4*>(orow + (long long)l * D + 4 * c4) =
            make_float4(coeff * x.x, coeff * x.y, coeff * x.z, coeff * x.w);
      }
    }
  }
}

// Per-edge red.global.add path (TFGNN_PATH_ATOMIC): the stock-TF-GPU formulation
// (UnsortedSegmentSum = atomicAdd).  Kept as the measured alternative to the CSR path.
__global__ void edge_scatter_atomic_kernel(const int2* __restrict__ edges, long long E, int l, int V,
                                           const float* __restrict__ X, int ldx, int C,
                                           const int* __restrict__ row_ptr, int normalize,
                                           float* __restrict__ out, int ldo, int col0) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long num_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int C4 = C >> 2;
  for (long long e = warp; e < E; e += num_warps) {
    const int2 st = __ldg(edges + e);
    if ((unsigned)st.x >= (unsigned)V || (unsigned)st.y >= (unsigned)V) continue;
    const long long seg = (long long)l * V + st.y;
    const float scale = normalize ? 1.0f / ((float)(__ldg(row_ptr + seg + 1) - __ldg(row_ptr + seg)) + kSmallNumber) : 1.0f;
    const float* row = X + (long long)st.x * ldx;
    float* orow = out + (long long)st.y * ldo + col0;
    for (int c4 = lane; c4 < C4; c4 += 32) {
      float4 x = f4_scale(ldg_f4(row + 4 * c4), scale);
      atomicAdd(reinterpret_cast<float4*>(orow + 4 * c4), x);  // red.global.add.v4.f32 (sm_90+)
    }
  }
}


int launch_edge_reduce(const EdgeReduceParams& p_in, bool merged, cudaStream_t st, int max_blocks) {
  EdgeReduceParams p = p_in;
  if (p.v_count <= 0) {  // full range
    p.v_begin = 0;
    p.v_count = p.V;
  }
  const long long items = merged ? (long long)p.v_count : (long long)p.L * p.v_count;
  if (items == 0 || p.C == 0) return 0;
  const bool plain = !merged && !p.T && !p.G && !p.hidden_relu && p.edge_act == TFGNN_ACT_NONE &&
                     !p.reduce_max && p.final_act == TFGNN_ACT_NONE;
  bool vec = (p.C % 4 == 0) && (p.ldx % 4 == 0) && (p.x_type_stride % 4 == 0) && (p.ldo % 4 == 0) &&
             (p.out_type_stride % 4 == 0) && aligned16(p.X) && aligned16(p.out) && p.C <= 128 * 4;
  if (p.T) vec = vec && (p.ldt % 4 == 0) && (p.t_type_stride % 4 == 0) && aligned16(p.T);
  if (p.G) vec = vec && (p.ldg % 4 == 0) && (p.g_type_stride % 4 == 0) && (p.beta_off % 4 == 0) && aligned16(p.G);
  if (!vec) {
    const long long total = items * p.C;
    const int blocks = ceil_div(total, 256);
    if (merged) edge_reduce_scalar_kernel<true><<<blocks, 256, 0, st>>>(p);
    else edge_reduce_scalar_kernel<false><<<blocks, 256, 0, st>>>(p);
    TFGNN_LAUNCH_CHECK();
    return 0;
  }
  const int nv = (p.C + 127) / 128;
  static const bool lean = [] { const char* e = getenv("TFGNN_B200_GATHER_LEAN"); return !e || atoi(e) != 0; }();
  int blocks = ceil_div(items * 32, 256);
  if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;
#define TFGNN_ER_LAUNCH(NV)                                                              \
  do {                                                                                   \
    if (merged) edge_reduce_kernel<NV, true, false><<<blocks, 256, 0, st>>>(p);          \
    else if (plain && lean) edge_reduce_kernel<NV, false, true, 2, (NV == 1 ? 6 : 5)><<<blocks, 256, 0, st>>>(p); \
    else if (plain) edge_reduce_kernel<NV, false, true><<<blocks, 256, 0, st>>>(p);      \
    else edge_reduce_kernel<NV, false, false><<<blocks, 256, 0, st>>>(p);                \
  } while (0)
  switch (nv) {
    case 1: TFGNN_ER_LAUNCH(1); break;
    case 2: TFGNN_ER_LAUNCH(2); break;
    case 3: TFGNN_ER_LAUNCH(3); break;
    default: TFGNN_ER_LAUNCH(4); break;
  }
#undef TFGNN_ER_LAUNCH
  TFGNN_LAUNCH_CHECK();
  return 0;
}

int launch_target_term(const float* h, int ldh, const int* row_ptr, int V, int L, int D, int normalize,
                       float* out, int ldo, int col0, cudaStream_t st) {
  const long long total = (long long)V * L * D;
  if (total == 0) return 0;
  if (D % 4 == 0 && ldh % 4 == 0 && ldo % 4 == 0 && col0 % 4 == 0 && aligned16(h) && aligned16(out)) {
    int blocks = ceil_div((long long)V * 32, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    target_term_vec_kernel<<<blocks, 256, 0, st>>>(h, ldh, row_ptr, V, L, D, normalize, out, ldo, col0);
    TFGNN_LAUNCH_CHECK();
    return 0;
  }
  target_term_kernel<<<ceil_div(total, 256), 256, 0, st>>>(h, ldh, row_ptr, V, L, D, normalize, out, ldo, col0);
  TFGNN_LAUNCH_CHECK();
  return 0;
}

int launch_edge_scatter_atomic(const tfgnn_batch* b, const float* X, int ldx, int C, int normalize,
                               float* out, int ldo, int type_stride, cudaStream_t st) {
  TFGNN_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && type_stride % 4 == 0 && aligned16(X) && aligned16(out),
                "atomic path needs 16-byte aligned rows (column counts divisible by 4)");
  for (int l = 0; l < b->L; ++l) {
    if (b->E[l] == 0) continue;
    int blocks = ceil_div(b->E[l] * 32, 256);
    if (blocks > 148 * 64) blocks = 148 * 64;
    edge_scatter_atomic_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const int2*>(b->adj[l]), b->E[l], l,
                                                        (int)b->V, X, ldx, C, b->row_ptr, normalize, out, ldo,
                                                        l * type_stride);
    TFGNN_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace tfgnn

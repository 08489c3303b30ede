// RGAT aggregation (rgat.py:125-163): softmax over ALL incoming edges of a target (all edge types jointly),
// per head, then the attention-weighted sum of the projected source rows.
//
// Node-level part (variants.cu): P_l = h W_l for every node once, and the per-edge score is split by linearity
// of the einsum (rgat.py:115-121) into node-level halves s_src[u,l,k] + s_tgt[v,l,k].
//
// Edge-level part (this file), HBM-bound: per edge 4 B index + 4*K B of source scores (twice) + 4*H B of P row.
//   * regular targets (in-degree <= kHubThreshold): one WARP per target, lanes over float4 column groups,
//     two passes over the CSR segments (running max, then exp-sum + weighted accumulate): coalesced 4*H-byte
//     row reads, no atomics, deterministic.
//   * hub targets (power-law graphs: up to 1e5+ incoming edges): the joint edge list is cut into chunks of
//     kHubChunk edges, one warp per chunk; chunk results are combined with float atomics (max, then sum)
//     and normalised in a final pass.  Without this a single warp would walk a hub serially for ~50 ms.
#include "layers.cuh"

namespace tfgnn {

constexpr int kHubThreshold = 2048;
constexpr int kHubChunk = 1024;

__device__ __forceinline__ float rgat_leaky(float x) { return x > 0.f ? x : kLeakyReluAlpha * x; }

__device__ __forceinline__ void rgat_atomic_max(float* addr, float val) {
  if (__float_as_int(val) >= 0) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(val));   // sign bit: -0.0f goes to the atomicMin branch
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(val));
}

struct RgatParams {
  const float* P;       // [Vs, L*H]
  const float* s_src;   // [Vs, L*K]
  const float* s_tgt;   // [Vs, L*K]
  const int* row_ptr;
  const int* src;
  long long V, tgt_off;
  int L, K, d, H, act;
  float* out;           // [V, H]
  // hub machinery
  float* hub_max;       // [V, K]
  float* hub_den;       // [V, K]
  int2* items;          // (target, chunk)
  int* item_count;
};

// One pass over the edges [e_lo, e_hi) of segment (l, v) for this lane's column group (head k):
//   PASS 0: m = max(m, score)      PASS 1: w = exp(score - m); den += w; acc += w * P_l[src, c..c+3]
//   PASS 2 (regular targets): ONE pass with a running maximum ("online softmax"): when the maximum grows, the sums so
//   far are rescaled by exp(m_old - m_new).  Every edge's score and P row are read once instead of twice (round 1 walked
//   each target's edges twice: 13 ms of the 22 ms cfg3 layer).
template <int PASS>
__device__ __forceinline__ void rgat_walk(const RgatParams& p, int l, int e_lo, int e_hi, float st, int k, int c,
                                          int lane, bool col_ok, float& m, float& den, float4& acc) {
  const long long LK = (long long)p.L * p.K, LH = (long long)p.L * p.H;
  for (int base = e_lo; base < e_hi; base += 32) {
    const int n = min(32, e_hi - base);
    const int my_src = lane < n ? __ldg(p.src + base + lane) : 0;
    for (int j0 = 0; j0 < n; j0 += 4) {
      float sc[4];
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long s = __shfl_sync(0xffffffffu, my_src, (j0 + u) & 31);
        const bool ok = (j0 + u < n) && col_ok;
        sc[u] = ok ? __ldg(p.s_src + s * LK + l * p.K + k) : 0.f;
        if (PASS >= 1) x[u] = ok ? ldg_f4(p.P + s * LH + (long long)l * p.H + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j0 + u < n) {
          const float score = rgat_leaky(sc[u] + st);
          if (PASS == 0) {
            m = fmaxf(m, score);
          } else if (PASS == 2) {
            // one expf per edge: either the maximum grows (weight of this edge is exp(0) = 1, the sums so far are rescaled;
            // the very first edge has m = -FLT_MAX: rescale = 0) or it stays (no rescale)
            if (score > m) {
              const float rescale = expf(m - score);
              m = score;
              den = fmaf(den, rescale, 1.0f);
              acc.x = fmaf(acc.x, rescale, x[u].x); acc.y = fmaf(acc.y, rescale, x[u].y);
              acc.z = fmaf(acc.z, rescale, x[u].z); acc.w = fmaf(acc.w, rescale, x[u].w);
            } else {
              const float w = expf(score - m);
              den += w;
              acc.x = fmaf(w, x[u].x, acc.x); acc.y = fmaf(w, x[u].y, acc.y);
              acc.z = fmaf(w, x[u].z, acc.z); acc.w = fmaf(w, x[u].w, acc.w);
            }
          } else {
            const float w = expf(score - m);
            den += w;
            acc.x = fmaf(w, x[u].x, acc.x); acc.y = fmaf(w, x[u].y, acc.y);
            acc.z = fmaf(w, x[u].z, acc.z); acc.w = fmaf(w, x[u].w, acc.w);
          }
        }
      }
    }
  }
}

// Regular targets: one warp per (target, 32-column-group block).  NVW = number of 128-column blocks (H <= 128*NVW
// handled by blockIdx.y).  Hubs are skipped here.
__global__ void __launch_bounds__(256) rgat_warp_kernel(const RgatParams p) {
  const int lane = threadIdx.x & 31;
  const long long v = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= p.V) return;
  const int c = (blockIdx.y * 32 + lane) * 4;
  const bool col_ok = c < p.H;
  const int k = col_ok ? c / p.d : 0;
  int deg = 0;
  for (int l = 0; l < p.L; ++l) {
    const long long seg = (long long)l * p.V + v;
    deg += __ldg(p.row_ptr + seg + 1) - __ldg(p.row_ptr + seg);
  }
  if (deg > kHubThreshold) return;   // handled by the hub kernels
  const long long LK = (long long)p.L * p.K;
  float m = kLowestFloat, den = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = 0; l < p.L; ++l) {
    const long long seg = (long long)l * p.V + v;
    const float st = __ldg(p.s_tgt + (v + p.tgt_off) * LK + l * p.K + k);
    rgat_walk<2>(p, l, __ldg(p.row_ptr + seg), __ldg(p.row_ptr + seg + 1), st, k, c, lane, col_ok, m, den, acc);
  }
  if (col_ok) {
    const float inv = den > 0.f ? 1.0f / den : 0.f;
    float o[4] = {acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv};
    apply_act_vec<4>(o, p.act);
    *reinterpret_cast<float4*>(p.out + v * p.H + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Hub discovery: one thread per target; a hub gets ceil(deg / kHubChunk) work items and its accumulators reset.
__global__ void rgat_hub_scan_kernel(const RgatParams p) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < p.V;
       v += (long long)gridDim.x * blockDim.x) {
    int deg = 0;
    for (int l = 0; l < p.L; ++l) {
      const long long seg = (long long)l * p.V + v;
      deg += p.row_ptr[seg + 1] - p.row_ptr[seg];
    }
    if (deg <= kHubThreshold) continue;
    const int nchunks = (deg + kHubChunk - 1) / kHubChunk;
    const int base = atomicAdd(p.item_count, nchunks);
    for (int c = 0; c < nchunks; ++c) p.items[base + c] = make_int2((int)v, c);
    for (int k = 0; k < p.K; ++k) {
      p.hub_max[v * p.K + k] = kLowestFloat;
      p.hub_den[v * p.K + k] = 0.f;
    }
    for (int c = 0; c < p.H; ++c) p.out[v * p.H + c] = 0.f;
  }
}

// Hub chunk pass: warp per (item, 128-column block).  PASS 0: chunk max -> atomic max.  PASS 1: chunk exp-sum and
// weighted sum against the final max -> atomic adds into hub_den and out (un-normalised).
template <int PASS>
__global__ void __launch_bounds__(256) rgat_hub_chunk_kernel(const RgatParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int count = *p.item_count;
  const int c = (blockIdx.y * 32 + lane) * 4;
  const bool col_ok = c < p.H;
  const int k = col_ok ? c / p.d : 0;
  const long long LK = (long long)p.L * p.K;
  for (long long it = warp; it < count; it += nwarps) {
    const int2 item = p.items[it];
    const long long v = item.x;
    const int r_lo = item.y * kHubChunk, r_hi = r_lo + kHubChunk;   // rank range in the joint edge list of v
    float m = PASS == 0 ? kLowestFloat : (col_ok ? p.hub_max[v * p.K + k] : 0.f);
    float den = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int rank0 = 0;
    for (int l = 0; l < p.L; ++l) {
      const long long seg = (long long)l * p.V + v;
      const int beg = __ldg(p.row_ptr + seg), end = __ldg(p.row_ptr + seg + 1);
      const int lo = max(r_lo - rank0, 0), hi = min(r_hi - rank0, end - beg);
      if (lo < hi) {
        const float st = __ldg(p.s_tgt + (v + p.tgt_off) * LK + l * p.K + k);
        rgat_walk<PASS>(p, l, beg + lo, beg + hi, st, k, c, lane, col_ok, m, den, acc);
      }
      rank0 += end - beg;
    }
    if (!col_ok) continue;
    if (PASS == 0) {
      if ((c % p.d) == 0) rgat_atomic_max(p.hub_max + v * p.K + k, m);
    } else {
      if ((c % p.d) == 0) atomicAdd(p.hub_den + v * p.K + k, den);
      float* o = p.out + v * p.H + c;
      atomicAdd(o, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
    }
  }
}

__global__ void rgat_hub_finalize_kernel(const RgatParams p) {
  const int count = *p.item_count;
  for (int it = blockIdx.x; it < count; it += gridDim.x) {
    const int2 item = p.items[it];
    if (item.y != 0) continue;           // once per hub
    const long long v = item.x;
    for (int c = threadIdx.x; c < p.H; c += blockDim.x) {
      const float den = p.hub_den[v * p.K + c / p.d];
      float o[1] = {den > 0.f ? p.out[v * p.H + c] / den : 0.f};
      apply_act_vec<1>(o, p.act);
      p.out[v * p.H + c] = o[0];
    }
  }
}

int launch_rgat_aggregate(tfgnn_batch* b, const float* P, const float* s_src, const float* s_tgt, int K, int d,
                          int activation, float* out, cudaStream_t st) {
  RgatParams p{};
  p.P = P; p.s_src = s_src; p.s_tgt = s_tgt; p.row_ptr = b->row_ptr; p.src = b->src_sorted;
  p.V = b->V; p.tgt_off = b->tgt_off; p.L = b->L; p.K = K; p.d = d; p.H = K * d; p.act = activation; p.out = out;
  const long long max_items = b->M_in / kHubChunk + b->M_in / kHubThreshold + 2;
  void *hm = nullptr, *hd = nullptr, *items = nullptr;
  int rc = batch_scratch(b, 8, (size_t)p.V * K * sizeof(float), &hm);
  if (rc) return rc;
  rc = batch_scratch(b, 9, (size_t)p.V * K * sizeof(float), &hd);
  if (rc) return rc;
  rc = batch_scratch(b, 10, (size_t)max_items * sizeof(int2) + 16, &items);
  if (rc) return rc;
  p.hub_max = (float*)hm; p.hub_den = (float*)hd;
  p.item_count = (int*)items;
  p.items = reinterpret_cast<int2*>(reinterpret_cast<char*>(items) + 16);
  TFGNN_CUDA(cudaMemsetAsync(p.item_count, 0, sizeof(int), st));
  const int col_blocks = (p.H / 4 + 31) / 32;
  int scan_blocks = ceil_div(p.V, 256);
  if (scan_blocks > 148 * 16) scan_blocks = 148 * 16;
  rgat_hub_scan_kernel<<<scan_blocks, 256, 0, st>>>(p);
  TFGNN_LAUNCH_CHECK();
  {
    dim3 grid((unsigned)ceil_div(p.V * 32, 256), col_blocks);
    rgat_warp_kernel<<<grid, 256, 0, st>>>(p);
    TFGNN_LAUNCH_CHECK();
  }
  {
    dim3 grid(148 * 4, col_blocks);
    rgat_hub_chunk_kernel<0><<<grid, 256, 0, st>>>(p);
    TFGNN_LAUNCH_CHECK();
    rgat_hub_chunk_kernel<1><<<grid, 256, 0, st>>>(p);
    TFGNN_LAUNCH_CHECK();
    rgat_hub_finalize_kernel<<<148, 128, 0, st>>>(p);
    TFGNN_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace tfgnn

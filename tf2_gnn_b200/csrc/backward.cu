// Backward pass of the RGCN-style layer: the reference differentiates through the layer
// with tf.GradientTape (tf2_gnn/models/graph_task_model.py:338-365); here the gradients of
//     out = act( rn(v) * sum_l A_l W_l ),   A_l[v] = s_{v,l} * sum_{(u,v) in A_l} h_u,   s = 1/(c_{v,l}+eps) or 1
// are computed with the same building blocks as the forward pass:
//   dZ      = dOut * act'(out) * rn(v)                        (elementwise, from the saved OUTPUT)
//   dW_l    = A_l^T dZ                                        (A recomputed by the CSR reduce; TN GEMM, reduction
//                                                              over the V nodes in fixed chunks -> deterministic)
//   dA      = dZ [W_0;..;W_{L-1}]^T, then dA_l[v] *= s_{v,l}   (3xTF32 tcgen05 GEMM + row/type scale)
//   dh[u]   = sum_l sum_{(u,v) in A_l} dA_l[v]                (CSR reduce over the SOURCE-keyed CSR: no atomics)
// Supported: 0 hidden layers, source or source+target state input, sum / mean / sqrt_n aggregation, activation after the
// aggregation, every activation of the reference's table (gelu through a recomputed pre-activation).
#include "layers.cuh"

namespace tfgnn {

__device__ __forceinline__ float act_grad_from_output(float y, int act) {
  switch (act) {
    case TFGNN_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case TFGNN_ACT_TANH: return 1.f - y * y;
    case TFGNN_ACT_LEAKY_RELU: return y > 0.f ? 1.f : kLeakyReluAlpha;
    case TFGNN_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;                               // d/dx (e^x - 1) = y + 1
    case TFGNN_ACT_SELU: return y > 0.f ? kSeluScale : y + kSeluScale * kSeluAlpha;   // scale*alpha*e^x = y + scale*alpha
    case TFGNN_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}
// gelu (utils/activation.py:7-14, tanh approximation) is not invertible from its output: derivative from the
// recomputed PRE-activation x.
__device__ __forceinline__ float gelu_grad_from_input(float x) {
  const float c = 0.7978845608028654f;
  const float t = tanhf(c * (x + 0.044715f * x * x * x));
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * c * (1.0f + 3.0f * 0.044715f * x * x);
}

__global__ void act_grad_kernel(const float* __restrict__ g, const float* __restrict__ out, long long V, int H,
                                int act, const int* __restrict__ row_ptr, int L, int row_norm,
                                float* __restrict__ dz) {
  const long long total = V * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / H;
    float s = 1.f;
    if (row_norm) {
      int cnt = 0;
      for (int l = 0; l < L; ++l) cnt += row_ptr[(long long)l * V + v + 1] - row_ptr[(long long)l * V + v];
      const float n = (float)max(cnt, 1);
      s = 1.f / (row_norm == 1 ? n : sqrtf(n));
    }
    // for gelu `out` holds the recomputed pre-activation
    dz[i] = g[i] * (act == TFGNN_ACT_GELU ? gelu_grad_from_input(out[i]) : act_grad_from_output(out[i], act)) * s;
  }
}

// dA[v, l*D + c] *= 1/(c_{v,l}+eps)   (first L*D columns of rows with leading dimension ld)
__global__ void scale_by_type_kernel(float* __restrict__ dA, int ld, long long V, int L, int D,
                                     const int* __restrict__ row_ptr) {
  const long long total = V * L * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / ((long long)L * D);
    const int col = (int)(i - v * (long long)L * D);
    const int l = col / D;
    const long long seg = (long long)l * V + v;
    dA[v * ld + col] *= 1.0f / ((float)(row_ptr[seg + 1] - row_ptr[seg]) + kSmallNumber);
  }
}

// use_target_state_as_input (gnn_edge_mlp.py:93-98): the target half of the node-level operand is
// T[v, l*D + c] = coeff(v,l) * h_v[c], coeff = c/(c+eps) or c; its gradient flows straight back to h_v:
// grad_h[v, c] += sum_l coeff(v,l) * dT[v, l*D + c]
__global__ void target_term_bwd_kernel(const float* __restrict__ dT, int ld, const int* __restrict__ row_ptr, long long V,
                                       int L, int D, int normalize, float* __restrict__ grad_h) {
  const long long total = V * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / D;
    const int c = (int)(i - v * D);
    float s = 0.f;
    for (int l = 0; l < L; ++l) {
      const long long seg = (long long)l * V + v;
      const float cnt = (float)(row_ptr[seg + 1] - row_ptr[seg]);
      const float coeff = normalize ? cnt * (1.0f / (cnt + kSmallNumber)) : cnt;
      s += coeff * dT[v * ld + (long long)l * D + c];
    }
    grad_h[i] += s;
  }
}

// WcatT[hh, col0 + l*D + d] = W_l[row0 + d, hh]   (operand of dA = dZ Wcat^T; ld_out = columns of WcatT)
__global__ void pack_transposed_kernel(PtrTable W, int L, int D, int H, float* __restrict__ out, int ld_out = 0,
                                       int col0 = 0, int row0 = 0) {
  const long long total = (long long)L * D * H;
  if (ld_out == 0) ld_out = L * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % ((long long)L * D));
    const int hh = (int)(i / ((long long)L * D));
    const int l = col / D, d = col - l * D;
    out[(long long)hh * ld_out + col0 + col] = reinterpret_cast<const float*>(W.p[l])[(long long)(row0 + d) * H + hh];
  }
}

// TN GEMM with the reduction over the (huge) node dimension: Cpart[chunk][k][n] = sum_{m in chunk} A[m,k] B[m,n].
// 128x128 output tile, 256 threads, 8x8 per thread; both operands are read row-wise (row m contiguous), so no
// transposition is needed in shared memory.  fp32 FFMA (deterministic; a tcgen05 version is future work).
constexpr int kTnTile = 128, kTnMB = 16, kTnChunk = 8192;

__global__ void __launch_bounds__(256, 2)
gemm_tn_partial_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, long long M,
                       int Kd, int N, float* __restrict__ Cpart) {
  // 16-byte row loads need 4-float leading dimensions and aligned bases (rgcn_bwd guarantees them; the Dense backward
  // of e.g. a 50-feature input layer does not)
  const bool vecA = (lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
  const bool vecB = (ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0;
  __shared__ __align__(16) float As[2][kTnMB][kTnTile];
  __shared__ __align__(16) float Bs[2][kTnMB][kTnTile];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int k0 = blockIdx.x * kTnTile, n0 = blockIdx.y * kTnTile;
  const long long m_begin = (long long)blockIdx.z * kTnChunk;
  const long long m_end = m_begin + kTnChunk < M ? m_begin + kTnChunk : M;
  const int lr = t >> 5, lc = (t & 31) * 4;     // load coordinates: rows lr, lr+8 ; 4 consecutive columns
  float4 ra[2], rb[2];
  auto load = [&](long long m0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long m = m0 + lr + 8 * i;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (m < m_end) {
        const float* pa = A + m * lda + k0 + lc;
        const float* pb = B + m * ldb + n0 + lc;
        if (vecA && k0 + lc + 3 < Kd) va = __ldg(reinterpret_cast<const float4*>(pa));
        else {
          if (k0 + lc < Kd) va.x = pa[0];
          if (k0 + lc + 1 < Kd) va.y = pa[1];
          if (k0 + lc + 2 < Kd) va.z = pa[2];
          if (k0 + lc + 3 < Kd) va.w = pa[3];
        }
        if (vecB && n0 + lc + 3 < N) vb = __ldg(reinterpret_cast<const float4*>(pb));
        else {
          if (n0 + lc < N) vb.x = pb[0];
          if (n0 + lc + 1 < N) vb.y = pb[1];
          if (n0 + lc + 2 < N) vb.z = pb[2];
          if (n0 + lc + 3 < N) vb.w = pb[3];
        }
      }
      ra[i] = va;
      rb[i] = vb;
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      *reinterpret_cast<float4*>(&As[buf][lr + 8 * i][lc]) = ra[i];
      *reinterpret_cast<float4*>(&Bs[buf][lr + 8 * i][lc]) = rb[i];
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  load(m_begin);
  store(0);
  __syncthreads();
  int buf = 0;
  for (long long m0 = m_begin; m0 < m_end; m0 += kTnMB, buf ^= 1) {
    if (m0 + kTnMB < m_end) load(m0 + kTnMB);
#pragma unroll
    for (int mm = 0; mm < kTnMB; ++mm) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][mm][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][mm][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][mm][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][mm][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (m0 + kTnMB < m_end) {
      store(buf ^ 1);
      __syncthreads();
    }
  }
  float* cp = Cpart + (long long)blockIdx.z * Kd * N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (k >= Kd) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n < N) cp[(long long)k * N + n] = acc[i][j];
    }
  }
}

// dW_l[d, :] = sum over chunks of Cpart[chunk][l*D + d, :]   (fixed order: deterministic)
// Cpart rows [k0, k0 + L*D) of K_total rows per chunk (K_total = 0: L*D) go to rows [row0, row0 + D) of dW_l.
__global__ void reduce_partials_kernel(const float* __restrict__ Cpart, int chunks, int L, int D, int N,
                                       PtrTable dW, int K_total = 0, int k0 = 0, int row0 = 0) {
  const long long total = (long long)L * D * N;
  if (K_total == 0) K_total = L * D;
  const long long chunk_stride = (long long)K_total * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += Cpart[(long long)c * chunk_stride + (long long)k0 * N + i];
    const int row = (int)(i / N), n = (int)(i - (long long)row * N);
    const int l = row / D, d = row - l * D;
    reinterpret_cast<float*>(const_cast<void*>(dW.p[l]))[(long long)(row0 + d) * N + n] = s;
  }
}

// ---- GGNN: Keras GRUCell (reset_after=True) backward, ggnn.py:84-87 -------------------------------------------------
// forward: z = sig(gx_z+gh_z), r = sig(gx_r+gh_r), hh = tanh(gx_h + r*gh_h), h' = z*h + (1-z)*hh.
// In place: gx <- dL/dgx, gh <- dL/dgh (each thread reads its six pre-activations before it writes),
// dh_direct = dL/dh' * z (the path of h through the convex combination).
__global__ void gru_gate_bwd_kernel(float* __restrict__ gx, float* __restrict__ gh, const float* __restrict__ h, int ldh,
                                    const float* __restrict__ grad_out, long long V, int H,
                                    float* __restrict__ dh_direct) {
  const long long total = V * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / H;
    const int c = (int)(i - v * H);
    float* x = gx + v * 3 * H;
    float* y = gh + v * 3 * H;
    const float ghh = y[2 * H + c];
    const float z = 1.0f / (1.0f + expf(-(x[c] + y[c])));
    const float r = 1.0f / (1.0f + expf(-(x[H + c] + y[H + c])));
    const float hh = tanhf(x[2 * H + c] + r * ghh);
    const float g = grad_out[i];
    const float da = g * (1.0f - z) * (1.0f - hh * hh);   // d/d(pre-tanh)
    const float daz = g * (h[v * ldh + c] - hh) * z * (1.0f - z);
    const float dar = da * ghh * r * (1.0f - r);
    x[c] = daz;          y[c] = daz;
    x[H + c] = dar;      y[H + c] = dar;
    x[2 * H + c] = da;   y[2 * H + c] = da * r;
    dh_direct[i] = g * z;
  }
}

// column sums over the node dimension in fixed 8192-row chunks (deterministic): partial[chunk][n]
__global__ void colsum_partial_kernel(const float* __restrict__ X, long long V, int N, float* __restrict__ partial) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long m0 = (long long)blockIdx.y * kTnChunk;
  const long long m1 = m0 + kTnChunk < V ? m0 + kTnChunk : V;
  float s = 0.f;
  for (long long m = m0; m < m1; ++m) s += X[m * N + n];
  partial[(long long)blockIdx.y * N + n] = s;
}
__global__ void colsum_reduce_kernel(const float* __restrict__ partial, int chunks, int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += partial[(long long)c * N + n];
  out[n] = s;
}

// grad_h = grad_h (message path) + a + b
__global__ void add3_kernel(float* __restrict__ acc, const float* __restrict__ a, const float* __restrict__ b, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc[i] = acc[i] + a[i] + b[i];
}

static int grid_cap(long long n) {
  int g = ceil_div(n, 256);
  return g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g);
}

}  // namespace tfgnn

using namespace tfgnn;

extern "C" int tfgnn_b200_rgcn_bwd(tfgnn_batch_t* b, tfgnn_batch_t* bt, const float* h, int32_t D,
                                   const float* const* W, int32_t H, uint32_t flags, int32_t aggregation,
                                   int32_t activation, const float* out, const float* grad_out, float* grad_h,
                                   float* const* grad_W, void* stream) {
  TFGNN_REQUIRE(b != nullptr && bt != nullptr, "batch / transposed batch is NULL");
  TFGNN_REQUIRE(D > 0 && H > 0, "D and H must be positive");
  TFGNN_REQUIRE(valid_act(activation) && valid_agg(aggregation), "unknown activation / aggregation code");
  const long long V = b->V;
  const int L = b->L;
  TFGNN_REQUIRE(bt->V == V && bt->L == L && b->V_src == V && bt->V_src == V,
                "forward and transposed batches must describe the same (unsharded) graph");
  if (flags & TFGNN_FLAG_ACT_BEFORE_AGGREGATION)
    return unsupported("rgcn_bwd: activation-before-aggregation is not built yet");
  const bool use_target = flags & TFGNN_FLAG_USE_TARGET_STATE;   // W_l is then [2D, H]: rows [0,D) source, [D,2D) target
  if (aggregation == TFGNN_AGG_MAX) return unsupported("rgcn_bwd: max aggregation is not built yet");
  if (D % 4 != 0 || H % 4 != 0) return unsupported("rgcn_bwd needs D and H to be multiples of 4");
  if (V == 0) return 0;
  TFGNN_REQUIRE(h && out && grad_out, "NULL pointer");
  TFGNN_REQUIRE(L == 0 || (W && grad_W), "weight / weight-gradient table is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const bool normalize = flags & TFGNN_FLAG_NORMALIZE_BY_NUM_INCOMING;
  const int LD = L * D;
  const int K = use_target ? 2 * LD : LD;
  if (L == 0) {
    if (grad_h) TFGNN_CUDA(cudaMemsetAsync(grad_h, 0, (size_t)V * D * sizeof(float), st));
    return 0;
  }
  PtrTable wt{}, gwt{};
  for (int l = 0; l < L; ++l) {
    TFGNN_REQUIRE(W[l] && grad_W[l], "a weight pointer is NULL");
    wt.p[l] = W[l];
    gwt.p[l] = grad_W[l];
  }
  void *dz = nullptr, *A = nullptr, *WT = nullptr, *part = nullptr;
  int rc = batch_enter(b, st);
  if (rc) return rc;
  rc = batch_enter(bt, st);
  if (rc) return rc;
  // 1a. gelu: act'(pre-activation).  The pre-activation is recomputed by the forward kernel without activation
  // BEFORE any other scratch pointer of this function is taken: the nested forward call may re-grow (= free and
  // re-allocate) the slots it shares with this function (2, 3, 6), which would leave them dangling.
  if (activation == TFGNN_ACT_GELU) {
    void* z = nullptr;
    rc = batch_scratch(b, 12, (size_t)V * H * sizeof(float), &z);
    if (rc) return rc;
    rc = edge_mlp_core(b, h, D, W, 0, H, flags, aggregation, TFGNN_ACT_NONE, TFGNN_PATH_AUTO, (float*)z, H, st);
    if (rc) return rc;
    out = (const float*)z;
  }
  rc = batch_scratch(b, 8, (size_t)V * H * sizeof(float), &dz);
  if (rc) return rc;
  rc = batch_scratch(b, 2, (size_t)V * K * sizeof(float), &A);     // A (forward operand), then dA
  if (rc) return rc;
  rc = batch_scratch(b, 3, (size_t)K * H * sizeof(float), &WT);
  if (rc) return rc;
  const int chunks = (int)((V + kTnChunk - 1) / kTnChunk);
  rc = batch_scratch(b, 9, (size_t)chunks * K * H * sizeof(float), &part);
  if (rc) return rc;

  // 1b. dZ = dOut * act'(out) * rn(v)
  act_grad_kernel<<<grid_cap(V * H), 256, 0, st>>>(grad_out, out, V, H, activation, b->row_ptr, L,
                                                   agg_row_norm(aggregation), (float*)dz);
  TFGNN_LAUNCH_CHECK();
  // 2. A_l (recomputed, normalised) and dW = A^T dZ
  {
    EdgeReduceParams p;
    p.X = h; p.ldx = D; p.x_type_stride = 0;
    p.row_ptr = b->row_ptr; p.src = b->src_sorted;
    p.out = (float*)A; p.ldo = K; p.out_type_stride = D;
    p.V = (int)V; p.L = L; p.C = D; p.normalize = normalize;
    rc = launch_edge_reduce(p, /*merged=*/false, st);
    if (rc) return rc;
    if (use_target) {
      rc = launch_target_term(h, D, b->row_ptr, (int)V, L, D, normalize, (float*)A, K, LD, st);
      if (rc) return rc;
    }
    dim3 grid((K + kTnTile - 1) / kTnTile, (H + kTnTile - 1) / kTnTile, chunks);
    gemm_tn_partial_kernel<<<grid, 256, 0, st>>>((const float*)A, K, (const float*)dz, H, V, K, H, (float*)part);
    TFGNN_LAUNCH_CHECK();
    reduce_partials_kernel<<<grid_cap((long long)LD * H), 256, 0, st>>>((const float*)part, chunks, L, D, H, gwt, K, 0, 0);
    TFGNN_LAUNCH_CHECK();
    if (use_target) {
      reduce_partials_kernel<<<grid_cap((long long)LD * H), 256, 0, st>>>((const float*)part, chunks, L, D, H, gwt, K, LD,
                                                                        D);
      TFGNN_LAUNCH_CHECK();
    }
  }
  if (!grad_h) return 0;
  // 3. dA = dZ Wcat^T (overwrites A), scaled per (v,l)
  pack_transposed_kernel<<<grid_cap((long long)LD * H), 256, 0, st>>>(wt, L, D, H, (float*)WT, K, 0, 0);
  TFGNN_LAUNCH_CHECK();
  if (use_target) {
    pack_transposed_kernel<<<grid_cap((long long)LD * H), 256, 0, st>>>(wt, L, D, H, (float*)WT, K, LD, D);
    TFGNN_LAUNCH_CHECK();
  }
  GemmEpilogue none;
  rc = node_gemm((const float*)dz, H, (const float*)WT, K, (float*)A, K, V, K, H, none, TFGNN_PATH_AUTO, b, 6, st);
  if (rc) return rc;
  if (normalize) {
    scale_by_type_kernel<<<grid_cap(V * LD), 256, 0, st>>>((float*)A, K, V, L, D, b->row_ptr);
    TFGNN_LAUNCH_CHECK();
  }
  // 4. dh[u] = sum over the edges LEAVING u (source-keyed CSR), all types merged
  {
    EdgeReduceParams p;
    p.X = (const float*)A; p.ldx = K; p.x_type_stride = D;
    p.row_ptr = bt->row_ptr; p.src = bt->src_sorted;
    p.out = grad_h; p.ldo = D;
    p.V = (int)V; p.L = L; p.C = D;
    rc = launch_edge_reduce(p, /*merged=*/true, st);
    if (rc) return rc;
  }
  if (use_target) {   // 5. the target half: grad_h[v] += sum_l coeff(v,l) * dT_l[v]
    target_term_bwd_kernel<<<grid_cap(V * D), 256, 0, st>>>((const float*)A + LD, K, b->row_ptr, V, L, D, normalize,
                                                            grad_h);
    TFGNN_LAUNCH_CHECK();
  }
  return 0;
}

// GGNN backward: gradient of tfgnn_b200_ggnn_fwd w.r.t. the node states, the per-type message
// weights and the GRU parameters.  Everything is recomputed from h (nothing but h is saved by the forward pass):
//   agg = sum_l s A_l W_l (forward kernel), gx = agg K + b0, gh = h U + b1 (tensor-core GEMMs),
//   gate backward in place -> dgx, dgh, dh_direct;  db = column sums;  dK = agg^T dgx, dU = h^T dgh (TN GEMM, fixed-order
//   partials);  dagg = dgx K^T, dh_rec = dgh U^T (tensor-core GEMMs);  messages: tfgnn_b200_rgcn_bwd with dagg.
extern "C" int tfgnn_b200_ggnn_bwd(tfgnn_batch_t* b, tfgnn_batch_t* bt, const float* h, int32_t D,
                                   const float* const* W, int32_t H, uint32_t flags, int32_t aggregation,
                                   const float* gru_kernel, const float* gru_recurrent_kernel, const float* gru_bias,
                                   const float* grad_out, float* grad_h, float* const* grad_W, float* grad_gru_kernel,
                                   float* grad_gru_recurrent_kernel, float* grad_gru_bias, void* stream) {
  TFGNN_REQUIRE(b != nullptr && bt != nullptr, "batch / transposed batch is NULL");
  TFGNN_REQUIRE(D == H, "GGNN needs node embedding dimension == hidden_dim (ggnn.py:30)");
  TFGNN_REQUIRE(valid_agg(aggregation), "unknown aggregation code");
  const long long V = b->V;
  const int L = b->L;
  if (flags & TFGNN_FLAG_USE_TARGET_STATE) return unsupported("ggnn_bwd: target-state input is not built yet");
  if (aggregation == TFGNN_AGG_MAX) return unsupported("ggnn_bwd: max aggregation is not built yet");
  if (H % 4 != 0) return unsupported("ggnn_bwd needs hidden_dim to be a multiple of 4");
  if (V == 0) return 0;
  TFGNN_REQUIRE(h && grad_out && grad_h, "NULL pointer");
  TFGNN_REQUIRE(gru_kernel && gru_recurrent_kernel && gru_bias, "GRU weight pointer is NULL");
  TFGNN_REQUIRE(grad_gru_kernel && grad_gru_recurrent_kernel && grad_gru_bias, "GRU gradient pointer is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const int N3 = 3 * H;
  const int chunks = (int)((V + kTnChunk - 1) / kTnChunk);
  void *agg = nullptr, *gx = nullptr, *gh = nullptr, *dagg = nullptr, *wT = nullptr, *part = nullptr, *dhd = nullptr,
       *tmp = nullptr;
  int rc = batch_enter(b, st);
  if (rc) return rc;
  rc = batch_enter(bt, st);
  if (rc) return rc;
  rc = batch_scratch(b, 11, (size_t)V * H * sizeof(float), &agg);
  if (rc) return rc;
  rc = batch_scratch(b, 12, (size_t)V * N3 * sizeof(float), &gx);
  if (rc) return rc;
  rc = batch_scratch(b, 13, (size_t)V * N3 * sizeof(float), &gh);
  if (rc) return rc;
  rc = batch_scratch(b, 14, (size_t)V * H * sizeof(float), &dagg);
  if (rc) return rc;
  rc = batch_scratch(b, 7, (size_t)N3 * H * sizeof(float), &wT);
  if (rc) return rc;
  rc = batch_scratch(b, 10, (size_t)chunks * ((size_t)H * N3 + N3) * sizeof(float), &part);
  if (rc) return rc;
  rc = batch_scratch(b, 4, (size_t)V * H * sizeof(float), &dhd);
  if (rc) return rc;
  rc = batch_scratch(b, 5, (size_t)V * H * sizeof(float), &tmp);
  if (rc) return rc;
  // 1. forward quantities: agg (no message activation, ggnn.py:68-83), gx, gh
  rc = edge_mlp_core(b, h, D, W, 0, H, flags & ~TFGNN_FLAG_ACT_BEFORE_AGGREGATION, aggregation, TFGNN_ACT_NONE,
                     TFGNN_PATH_AUTO, (float*)agg, H, st);
  if (rc) return rc;
  GemmEpilogue e0, e1, none;
  e0.bias = gru_bias;
  e1.bias = gru_bias + N3;
  rc = node_gemm((const float*)agg, H, gru_kernel, N3, (float*)gx, N3, V, N3, H, e0, TFGNN_PATH_AUTO, b, 6, st);
  if (rc) return rc;
  rc = node_gemm(h, D, gru_recurrent_kernel, N3, (float*)gh, N3, V, N3, H, e1, TFGNN_PATH_AUTO, b, 6, st);
  if (rc) return rc;
  // 2. gates
  gru_gate_bwd_kernel<<<grid_cap(V * H), 256, 0, st>>>((float*)gx, (float*)gh, h, D, grad_out, V, H, (float*)dhd);
  TFGNN_LAUNCH_CHECK();
  // 3. bias gradients: rows 0 / 1 of gru_bias belong to gx / gh
  float* cpart = (float*)part + (size_t)chunks * H * N3;
  for (int which = 0; which < 2; ++which) {
    dim3 grid((N3 + 127) / 128, chunks);
    colsum_partial_kernel<<<grid, 128, 0, st>>>((const float*)(which ? gh : gx), V, N3, cpart);
    TFGNN_LAUNCH_CHECK();
    colsum_reduce_kernel<<<(N3 + 127) / 128, 128, 0, st>>>(cpart, chunks, N3, grad_gru_bias + (size_t)which * N3);
    TFGNN_LAUNCH_CHECK();
  }
  // 4. dK = agg^T dgx, dU = h^T dgh
  for (int which = 0; which < 2; ++which) {
    PtrTable gt{};
    gt.p[0] = which ? grad_gru_recurrent_kernel : grad_gru_kernel;
    dim3 grid((H + kTnTile - 1) / kTnTile, (N3 + kTnTile - 1) / kTnTile, chunks);
    gemm_tn_partial_kernel<<<grid, 256, 0, st>>>(which ? h : (const float*)agg, which ? D : H,
                                                 (const float*)(which ? gh : gx), N3, V, H, N3, (float*)part);
    TFGNN_LAUNCH_CHECK();
    reduce_partials_kernel<<<grid_cap((long long)H * N3), 256, 0, st>>>((const float*)part, chunks, 1, H, N3, gt);
    TFGNN_LAUNCH_CHECK();
  }
  // 5. dagg = dgx K^T, dh_rec = dgh U^T
  for (int which = 0; which < 2; ++which) {
    PtrTable wt{};
    wt.p[0] = which ? gru_recurrent_kernel : gru_kernel;
    pack_transposed_kernel<<<grid_cap((long long)H * N3), 256, 0, st>>>(wt, 1, H, N3, (float*)wT);   // [3H, H]
    TFGNN_LAUNCH_CHECK();
    rc = node_gemm((const float*)(which ? gh : gx), N3, (const float*)wT, H, (float*)(which ? tmp : dagg), H, V, H, N3,
                   none, TFGNN_PATH_AUTO, b, 6, st);
    if (rc) return rc;
  }
  // 6. messages: dagg -> grad_h (through the edges) and grad_W
  rc = tfgnn_b200_rgcn_bwd(b, bt, h, D, W, H, flags & ~TFGNN_FLAG_ACT_BEFORE_AGGREGATION, aggregation, TFGNN_ACT_NONE,
                           (const float*)dagg, (const float*)dagg, grad_h, grad_W, stream);
  if (rc) return rc;
  // 7. grad_h += dh_direct + dh_rec
  add3_kernel<<<grid_cap(V * H), 256, 0, st>>>(grad_h, (const float*)dhd, (const float*)tmp, V * H);
  TFGNN_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================================
// Node-level glue of GNN._internal_call under training (gnn.py:279-327): backward of the bias-free / biased Dense layers,
// LayerNormalization, and the Philox dropout shared by forward and backward.  The reference gets all of these from
// tf.GradientTape (models/graph_task_model.py:338-365).
// =====================================================================================================================
namespace tfgnn {

// dZ = dY * act'(.) for a Dense layer: derivative from the OUTPUT (every activation but gelu) or from the recomputed
// pre-activation (gelu).
__global__ void dense_act_grad_kernel(const float* __restrict__ g, const float* __restrict__ y, long long n, int act,
                                      float* __restrict__ dz) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dz[i] = g[i] * (act == TFGNN_ACT_GELU ? gelu_grad_from_input(y[i]) : act_grad_from_output(y[i], act));
}

// LayerNormalization backward, one warp per row:
//   xhat = (x - mean) * rstd;  dxhat = g * gamma;
//   dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat));   t[v,c] = g * xhat  (column-summed into dgamma)
__global__ void layer_norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                      const float* __restrict__ g, long long V, int H, float eps,
                                      float* __restrict__ dx, float* __restrict__ t) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= V) return;
  const float* xr = x + row * H;
  const float* gr = g + row * H;
  float s = 0.f;
  for (int c = lane; c < H; c += 32) s += xr[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)H;
  float q = 0.f;
  for (int c = lane; c < H; c += 32) {
    const float d = xr[c] - mean;
    q += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)H + eps);
  float a = 0.f, b = 0.f;   // sum dxhat, sum dxhat * xhat
  for (int c = lane; c < H; c += 32) {
    const float xhat = (xr[c] - mean) * rstd;
    const float dxh = gr[c] * gamma[c];
    a += dxh;
    b += dxh * xhat;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  a /= (float)H;
  b /= (float)H;
  for (int c = lane; c < H; c += 32) {
    const float xhat = (xr[c] - mean) * rstd;
    const float dxh = gr[c] * gamma[c];
    if (dx) dx[row * H + c] = rstd * (dxh - a - xhat * b);
    t[row * H + c] = gr[c] * xhat;
  }
}

// Philox4x32-10 (Salmon et al. 2011), the counter-based generator TensorFlow's stateless random ops use as well.
// One counter value yields 4 uniform 32-bit words: element i takes word i & 3 of counter i >> 2, so the mask of an
// element depends only on (seed, offset, i): the backward pass regenerates it instead of storing it.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// tf.nn.dropout(x, rate): keep with probability 1 - rate, scale kept values by 1 / (1 - rate)   (gnn.py:285-289)
__global__ void dropout_kernel(const float* __restrict__ x, long long n, float rate, unsigned long long seed,
                               unsigned long long offset, float* __restrict__ out) {
  const float scale = 1.0f / (1.0f - rate);
  const long long groups = (n + 3) >> 2;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < groups;
       gi += (long long)gridDim.x * blockDim.x) {
    const unsigned long long c = (unsigned long long)gi + offset;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = gi * 4 + j;
      if (i < n) {
        const float u = (float)(w[j] >> 8) * (1.0f / 16777216.0f);   // uniform in [0, 1)
        out[i] = u >= rate ? x[i] * scale : 0.0f;
      }
    }
  }
}

__global__ void axpby_kernel(const float* __restrict__ a, float alpha, const float* __restrict__ b, float beta,
                             long long n, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = b ? alpha * a[i] + beta * b[i] : alpha * a[i];
}

// column sums of X [V, N] in fixed-order chunks -> out [N]   (bias / gamma / beta gradients)
static int column_sums(const float* X, long long V, int N, float* out, cudaStream_t st) {
  const int chunks = (int)((V + kTnChunk - 1) / kTnChunk);
  void* part = nullptr;
  int rc = pool_alloc(&part, (size_t)chunks * N * sizeof(float), st);
  if (rc) return rc;
  dim3 grid((N + 127) / 128, chunks);
  colsum_partial_kernel<<<grid, 128, 0, st>>>(X, V, N, (float*)part);
  TFGNN_LAUNCH_CHECK();
  colsum_reduce_kernel<<<(N + 127) / 128, 128, 0, st>>>((const float*)part, chunks, N, out);
  TFGNN_LAUNCH_CHECK();
  pool_free(part, st);
  return 0;
}

}  // namespace tfgnn

// Backward of out = act(x W + bias): grad_x = dZ W^T (tensor-core GEMM), grad_W = x^T dZ (TN GEMM, fixed-order partials),
// grad_bias = column sums of dZ, dZ = grad_out * act'.  `out` is the saved forward output.
extern "C" int tfgnn_b200_dense_bwd(const float* x, const float* W, const float* bias, const float* out,
                                    const float* grad_out, int64_t V, int32_t K, int32_t N, int32_t activation,
                                    float* grad_x, float* grad_W, float* grad_bias, void* stream) {
  TFGNN_REQUIRE(V >= 0 && K > 0 && N > 0, "bad dense shape");
  TFGNN_REQUIRE(valid_act(activation), "unknown activation code");
  cudaStream_t st = (cudaStream_t)stream;
  if (V == 0) {
    if (grad_W) TFGNN_CUDA(cudaMemsetAsync(grad_W, 0, (size_t)K * N * sizeof(float), st));
    if (grad_bias) TFGNN_CUDA(cudaMemsetAsync(grad_bias, 0, (size_t)N * sizeof(float), st));
    return 0;
  }
  TFGNN_REQUIRE(x && W && out && grad_out, "NULL pointer");
  void *dz = nullptr, *pre = nullptr, *wT = nullptr, *part = nullptr;
  int rc = pool_alloc(&dz, (size_t)V * N * sizeof(float), st);
  if (rc) return rc;
  const float* y = out;
  if (activation == TFGNN_ACT_GELU) {   // derivative needs the pre-activation: recompute x W + bias
    rc = pool_alloc(&pre, (size_t)V * N * sizeof(float), st);
    if (!rc) rc = tfgnn_b200_dense_bias_fwd(x, W, bias, (float*)pre, V, K, N, TFGNN_ACT_NONE, TFGNN_PATH_AUTO, stream);
    if (rc) { pool_free(dz, st); pool_free(pre, st); return rc; }
    y = (const float*)pre;
  }
  dense_act_grad_kernel<<<grid_cap(V * N), 256, 0, st>>>(grad_out, y, V * N, activation, (float*)dz);
  g_launch_count.fetch_add(1);
  if (grad_W) {
    const int chunks = (int)((V + kTnChunk - 1) / kTnChunk);
    rc = pool_alloc(&part, (size_t)chunks * K * N * sizeof(float), st);
    if (!rc) {
      PtrTable gt{};
      gt.p[0] = grad_W;
      dim3 grid((K + kTnTile - 1) / kTnTile, (N + kTnTile - 1) / kTnTile, chunks);
      gemm_tn_partial_kernel<<<grid, 256, 0, st>>>(x, K, (const float*)dz, N, V, K, N, (float*)part);
      g_launch_count.fetch_add(1);
      reduce_partials_kernel<<<grid_cap((long long)K * N), 256, 0, st>>>((const float*)part, chunks, 1, K, N, gt);
      g_launch_count.fetch_add(1);
    }
  }
  if (!rc && grad_bias) rc = column_sums((const float*)dz, V, N, grad_bias, st);
  if (!rc && grad_x) {
    rc = pool_alloc(&wT, (size_t)N * K * sizeof(float), st);
    if (!rc) {
      PtrTable wt{};
      wt.p[0] = W;
      pack_transposed_kernel<<<grid_cap((long long)K * N), 256, 0, st>>>(wt, 1, K, N, (float*)wT);   // [N, K]
      g_launch_count.fetch_add(1);
      rc = tfgnn_b200_dense_fwd((const float*)dz, (const float*)wT, grad_x, V, N, K, TFGNN_ACT_NONE, TFGNN_PATH_AUTO,
                                stream);
    }
  }
  if (!rc) rc = check_cuda(cudaGetLastError(), "dense_bwd kernels", __FILE__, __LINE__);
  pool_free(dz, st);
  pool_free(pre, st);
  pool_free(wT, st);
  pool_free(part, st);
  return rc;
}

extern "C" int tfgnn_b200_layer_norm_bwd(const float* x, const float* gamma, const float* grad_out, int64_t V, int32_t H,
                                         float epsilon, float* grad_x, float* grad_gamma, float* grad_beta,
                                         void* stream) {
  TFGNN_REQUIRE(V >= 0 && H > 0, "bad layer_norm shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (V == 0) {
    if (grad_gamma) TFGNN_CUDA(cudaMemsetAsync(grad_gamma, 0, (size_t)H * sizeof(float), st));
    if (grad_beta) TFGNN_CUDA(cudaMemsetAsync(grad_beta, 0, (size_t)H * sizeof(float), st));
    return 0;
  }
  TFGNN_REQUIRE(x && gamma && grad_out, "NULL pointer");
  void* t = nullptr;
  int rc = pool_alloc(&t, (size_t)V * H * sizeof(float), st);
  if (rc) return rc;
  layer_norm_bwd_kernel<<<ceil_div(V * 32, 256), 256, 0, st>>>(x, gamma, grad_out, V, H, epsilon, grad_x, (float*)t);
  g_launch_count.fetch_add(1);
  rc = check_cuda(cudaGetLastError(), "layer_norm_bwd_kernel", __FILE__, __LINE__);
  if (!rc && grad_gamma) rc = column_sums((const float*)t, V, H, grad_gamma, st);
  if (!rc && grad_beta) rc = column_sums(grad_out, V, H, grad_beta, st);
  pool_free(t, st);
  return rc;
}

// tf.nn.dropout (gnn.py:285-289, graph_global_exchange.py:98-101).  The mask is a pure function of (seed, offset, element
// index), so the backward pass calls the same entry on the incoming gradient.
extern "C" int tfgnn_b200_dropout(const float* x, int64_t n, float rate, uint64_t seed, uint64_t offset, float* out,
                                  void* stream) {
  TFGNN_REQUIRE(n >= 0 && rate >= 0.0f && rate < 1.0f, "dropout rate must lie in [0, 1)");
  if (n == 0) return 0;
  TFGNN_REQUIRE(x && out, "NULL pointer");
  dropout_kernel<<<grid_cap((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(x, n, rate, seed, offset, out);
  TFGNN_LAUNCH_CHECK();
  return 0;
}

extern "C" int tfgnn_b200_axpby(const float* a, float alpha, const float* b, float beta, int64_t n, float* out,
                                void* stream) {
  TFGNN_REQUIRE(n >= 0, "negative size");
  if (n == 0) return 0;
  TFGNN_REQUIRE(a && out, "NULL pointer");
  axpby_kernel<<<grid_cap(n), 256, 0, (cudaStream_t)stream>>>(a, alpha, b, beta, n, out);
  TFGNN_LAUNCH_CHECK();
  return 0;
}

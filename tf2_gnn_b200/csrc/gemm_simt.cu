// fp32 SIMT node-level GEMM with fused epilogue — the always-correct path for arbitrary shapes
// (doctest sizes like D=3/H=7, K not a multiple of 32, ...).  The tcgen05 path (gemm_tc.cu) takes
// over for tile-friendly shapes.  128x128x16 tiles, 256 threads, 8x8 register block per thread,
// register-staged double buffering.
#include "gemm.cuh"

namespace tfgnn {

constexpr int kBM = 128, kBN = 128, kBK = 16, kPad = 4;

__device__ __forceinline__ float row_norm_factor(const GemmEpilogue& e, long long row) {
  if (e.row_norm == 0) return 1.0f;
  int cnt = 0;
  for (int l = 0; l < e.L; ++l) {
    const long long s = (long long)l * e.V + e.row0 + row;
    cnt += e.row_ptr[s + 1] - e.row_ptr[s];
  }
  const float c = (float)max(cnt, 1);
  return e.row_norm == 1 ? c : sqrtf(c);
}

__global__ void __launch_bounds__(256, 2)
gemm_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                 float* __restrict__ C, int ldc, long long M, int N, int K, int n_tiles,
                 GemmEpilogue epi, int vecA, int vecB, int vecC) {
  __shared__ __align__(16) float As[2][kBK][kBM + kPad];
  __shared__ __align__(16) float Bs[2][kBK][kBN];
  const int t = threadIdx.x;
  const long long m_tile = blockIdx.x / n_tiles;
  const int n_tile = blockIdx.x % n_tiles;
  const long long m0 = m_tile * kBM;
  const int n0 = n_tile * kBN;
  const int tx = t & 15, ty = t >> 4;

  // global->register staging coordinates
  const int a_row = t >> 2, a_k = (t & 3) * 4;   // rows a_row, a_row+64
  const int b_k = t >> 5, b_n = (t & 31) * 4;    // k rows b_k, b_k+8
  float4 ra[2], rb[2];

  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long r = m0 + a_row + 64 * i;
      const int k = k0 + a_k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < M) {
        const float* p = A + r * lda + k;
        if (vecA && k + 3 < K) v = __ldg(reinterpret_cast<const float4*>(p));
        else {
          if (k < K) v.x = __ldg(p);
          if (k + 1 < K) v.y = __ldg(p + 1);
          if (k + 2 < K) v.z = __ldg(p + 2);
          if (k + 3 < K) v.w = __ldg(p + 3);
        }
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = k0 + b_k + 8 * i;
      const int n = n0 + b_n;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) {
        const float* p = B + (long long)k * ldb + n;
        if (vecB && n + 3 < N) v = __ldg(reinterpret_cast<const float4*>(p));
        else {
          if (n < N) v.x = __ldg(p);
          if (n + 1 < N) v.y = __ldg(p + 1);
          if (n + 2 < N) v.z = __ldg(p + 2);
          if (n + 3 < N) v.w = __ldg(p + 3);
        }
      }
      rb[i] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = a_row + 64 * i;
      As[buf][a_k + 0][r] = ra[i].x;
      As[buf][a_k + 1][r] = ra[i].y;
      As[buf][a_k + 2][r] = ra[i].z;
      As[buf][a_k + 3][r] = ra[i].w;
      *reinterpret_cast<float4*>(&Bs[buf][b_k + 8 * i][b_n]) = rb[i];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int num_k_tiles = (K + kBK - 1) / kBK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < num_k_tiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < num_k_tiles) load_tile((kt + 1) * kBK);
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < num_k_tiles) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
    const float rn = row_norm_factor(epi, r);
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[i][jh * 4 + j];
        if (n + j < N) {
          if (epi.mul) v *= __ldg(epi.mul + r * epi.ldm + n + j);
          if (epi.accumulate) v += C[r * ldc + n + j];
        }
        if (epi.finalize) {
          if (epi.row_norm) v = v / rn;
          if (epi.bias && n + j < N) v += __ldg(epi.bias + n + j);
          v = apply_act(v, epi.act);
        }
        o[j] = v;
      }
      float* cp = C + r * ldc + n;
      if (vecC && n + 3 < N) {
        *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) cp[j] = o[j];
      }
    }
  }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int launch_gemm_simt(const float* A, int lda, const float* B, int ldb, float* C, int ldc, long long M,
                     int N, int K, const GemmEpilogue& epi, cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  const int n_tiles = (N + kBN - 1) / kBN;
  const long long m_tiles = (M + kBM - 1) / kBM;
  const long long blocks = m_tiles * n_tiles;
  TFGNN_REQUIRE(blocks < (1ll << 31), "GEMM grid too large");
  const int vecA = (lda % 4 == 0) && al16(A);
  const int vecB = (ldb % 4 == 0) && al16(B);
  const int vecC = (ldc % 4 == 0) && al16(C);
  gemm_simt_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, n_tiles, epi, vecA,
                                                     vecB, vecC);
  TFGNN_LAUNCH_CHECK();
  return 0;
}

// ---- weight packing -------------------------------------------------------------------------
__global__ void pack_vertical_kernel(PtrTable src, int nblk, int row0, int rows, int cols, int ld_src,
                                     float* __restrict__ dst, int ld_dst, int dst_row0) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)nblk * rows * cols;
  if (idx >= total) return;
  const int c = (int)(idx % cols);
  const long long br = idx / cols;
  const int r = (int)(br % rows);
  const int blk = (int)(br / rows);
  const float* s = reinterpret_cast<const float*>(src.p[blk]);
  dst[((long long)dst_row0 + (long long)blk * rows + r) * ld_dst + c] = s[(long long)(row0 + r) * ld_src + c];
}

__global__ void pack_horizontal_kernel(PtrTable src, int nblk, int row0, int rows, int cols, int ld_src,
                                       float* __restrict__ dst, int ld_dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)nblk * rows * cols;
  if (idx >= total) return;
  const int c = (int)(idx % cols);
  const long long rb = idx / cols;
  const int blk = (int)(rb % nblk);
  const int r = (int)(rb / nblk);
  const float* s = reinterpret_cast<const float*>(src.p[blk]);
  dst[(long long)r * ld_dst + (long long)blk * cols + c] = s[(long long)(row0 + r) * ld_src + c];
}

int launch_pack_vertical(const PtrTable& src, int nblk, int row0, int rows, int cols, int ld_src, float* dst,
                         int ld_dst, int dst_row0, cudaStream_t st) {
  const long long total = (long long)nblk * rows * cols;
  if (total == 0) return 0;
  pack_vertical_kernel<<<ceil_div(total, 256), 256, 0, st>>>(src, nblk, row0, rows, cols, ld_src, dst, ld_dst,
                                                            dst_row0);
  TFGNN_LAUNCH_CHECK();
  return 0;
}

int launch_pack_horizontal(const PtrTable& src, int nblk, int row0, int rows, int cols, int ld_src, float* dst,
                           int ld_dst, cudaStream_t st) {
  const long long total = (long long)nblk * rows * cols;
  if (total == 0) return 0;
  pack_horizontal_kernel<<<ceil_div(total, 256), 256, 0, st>>>(src, nblk, row0, rows, cols, ld_src, dst, ld_dst);
  TFGNN_LAUNCH_CHECK();
  return 0;
}

}  // namespace tfgnn

// Micro-benchmark: what random-row gather-reduce rate can a B200 sustain?  (measurement tool, not product code)
//
// The RGCN hot path reads one full source row (4*D bytes, D = 256 or 320) per edge from a node table far larger
// than L2.  The HBM "copy" peak in MEASURED_PEAKS.json is a streaming number; this tool measures the ceiling of
// the actual access pattern so that DESIGN.md can say how far the fused kernel is from what the memory system
// can deliver for 1 KB random rows.  Variants:
//   ldg   : warp-per-segment register gather (float4 per lane, U rows in flight), W warps per SM
//   bulk  : cp.async.bulk global->shared row copies (no registers held by loads in flight), reduced from smem
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/gather_ceiling tools/gather_ceiling.cu
// Run  : tools/gather_ceiling [rows=1000000] [D=256] [edges=20000000] [deg=20]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));      \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ld_nc_f4_hint(const float* ptr, uint64_t policy) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

// ---- variant 1: register gather -----------------------------------------------------------------
template <int NV, int U>
__global__ void ldg_gather(const float* __restrict__ h, int D, const int* __restrict__ src, long long nseg, int deg,
                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const uint64_t pol = policy_evict_first();
  const int C4 = D >> 2;
  for (long long seg = warp; seg < nseg; seg += nwarps) {
    const int* ids = src + seg * deg;
    const int my = lane < deg ? __ldg(ids + lane) : 0;
    float4 acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = make_float4(0, 0, 0, 0);
    for (int e0 = 0; e0 < deg; e0 += U) {
      float4 r[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = __shfl_sync(0xffffffffu, my, (e0 + u) & 31);
        const float* rp = h + (long long)s * D;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c4 = lane + 32 * j;
          r[u][j] = (e0 + u < deg && c4 < C4) ? ld_nc_f4_hint(rp + 4 * c4, pol) : make_float4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          acc[j].x += r[u][j].x; acc[j].y += r[u][j].y; acc[j].z += r[u][j].z; acc[j].w += r[u][j].w;
        }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < C4) reinterpret_cast<float4*>(out + seg * D)[c4] = acc[j];
    }
  }
}

// ---- variant 2: bulk-async gather into shared memory ----------------------------------------------
// Each warp owns 2 batches x R row buffers; lanes 0..R-1 issue one bulk copy each (a batch = R edges of the
// segment stream), the whole warp reduces a landed batch from shared memory while the next one is in flight.
template <int NV, int R>
__global__ void bulk_gather(const float* __restrict__ h, int D, const int* __restrict__ src, long long nseg, int deg,
                            float* __restrict__ out, int warps_per_cta) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)D * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                 // 2 per warp
  uint8_t* bufs = smem + 1024 + (size_t)w * 2 * R * row_bytes;       // [2][R][row_bytes]
  uint64_t* bar = bars + 2 * w;
  if (lane == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint64_t pol = policy_evict_first();
  const long long warp = (long long)blockIdx.x * warps_per_cta + w;
  const long long nwarps = (long long)gridDim.x * warps_per_cta;
  // this warp's edge stream: segments warp, warp+nwarps, ... ; flatten into batches of R edges
  const long long my_segs = warp < nseg ? (nseg - warp + nwarps - 1) / nwarps : 0;
  const long long total_edges = my_segs * deg;
  const long long nbatch = (total_edges + R - 1) / R;
  auto edge_src = [&](long long k) {   // k-th edge of this warp's stream
    const long long sidx = k / deg;
    const int off = (int)(k - sidx * deg);
    return __ldg(src + (warp + sidx * nwarps) * deg + off);
  };
  auto issue = [&](long long b) {
    const int buf = (int)(b & 1);
    const long long k = b * R + lane;
    const int n = (int)min((long long)R, total_edges - b * R);
    if (lane == 0) mbar_expect_tx(&bar[buf], (uint32_t)n * row_bytes);
    __syncwarp();
    if (lane < n) {
      const int s = edge_src(k);
      bulk_g2s(bufs + ((size_t)buf * R + lane) * row_bytes, h + (long long)s * D, row_bytes, &bar[buf], pol);
    }
  };
  const int C4 = D >> 2;
  float4 acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = make_float4(0, 0, 0, 0);
  if (nbatch > 0) issue(0);
  long long edge = 0;
  for (long long b = 0; b < nbatch; ++b) {
    if (b + 1 < nbatch) issue(b + 1);
    const int buf = (int)(b & 1);
    mbar_wait(&bar[buf], (uint32_t)((b >> 1) & 1));
    const int n = (int)min((long long)R, total_edges - b * R);
    for (int i = 0; i < n; ++i, ++edge) {
      const float4* rowp = reinterpret_cast<const float4*>(bufs + ((size_t)buf * R + i) * row_bytes);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c4 = lane + 32 * j;
        if (c4 < C4) {
          const float4 x = rowp[c4];
          acc[j].x += x.x; acc[j].y += x.y; acc[j].z += x.z; acc[j].w += x.w;
        }
      }
      if ((edge + 1) % deg == 0) {
        const long long seg = warp + (edge / deg) * nwarps;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int c4 = lane + 32 * j;
          if (c4 < C4) reinterpret_cast<float4*>(out + seg * D)[c4] = acc[j];
          acc[j] = make_float4(0, 0, 0, 0);
        }
      }
    }
    __syncwarp();   // all lanes done reading `buf` before it is refilled two iterations later
  }
}

// ---- variant 3: rolling bulk ring ------------------------------------------------------------------
// Each warp owns Q row slots with one mbarrier each and keeps Q bulk copies in flight at all times: after a
// landed row has been added to the accumulator its slot is immediately refilled with the row Q edges ahead.
// Warps process CONTIGUOUS segment ranges, so source ids are fetched 32 at a time with coalesced loads.
template <int NV>
__global__ void roll_gather(const float* __restrict__ h, int D, const int* __restrict__ src, long long nseg, int deg,
                            float* __restrict__ out, int warps_per_cta, int Q) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)D * 4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem) + (size_t)w * Q;
  uint8_t* bufs = smem + 4096 + (size_t)w * Q * row_bytes;
  if (lane < Q) mbar_init(&bar[lane], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const uint64_t pol = policy_evict_first();
  const long long warp = (long long)blockIdx.x * warps_per_cta + w;
  const long long nwarps = (long long)gridDim.x * warps_per_cta;
  const long long per = (nseg + nwarps - 1) / nwarps;
  const long long s0 = warp * per, s1 = min(nseg, s0 + per);
  if (s0 >= s1) return;
  const long long e0 = s0 * deg, e1 = s1 * deg;
  const int C4 = D >> 2;
  float4 acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = make_float4(0, 0, 0, 0);
  long long issue = e0, cons = e0;
  long long blk = e0;                       // first edge of the id block held in `ids`
  int ids = blk + lane < e1 ? __ldg(src + blk + lane) : 0;
  int ids_next = blk + 32 + lane < e1 ? __ldg(src + blk + 32 + lane) : 0;
  int islot = 0, cslot = 0, in_blk = 0, left = deg;
  uint32_t cphase = 0;
  long long seg = s0;
  auto issue_one = [&]() {                  // warp-uniform; lane 0 issues
    if (in_blk == 32) {
      in_blk = 0;
      blk += 32;
      ids = ids_next;
      ids_next = blk + 32 + lane < e1 ? __ldg(src + blk + 32 + lane) : 0;
    }
    const int s = __shfl_sync(0xffffffffu, ids, in_blk);
    if (lane == 0) {
      mbar_expect_tx(&bar[islot], row_bytes);
      bulk_g2s(bufs + (size_t)islot * row_bytes, h + (long long)s * D, row_bytes, &bar[islot], pol);
    }
    islot = islot + 1 == Q ? 0 : islot + 1;
    ++in_blk;
    ++issue;
  };
  for (int i = 0; i < Q && issue < e1; ++i) issue_one();
  while (cons < e1) {
    mbar_wait(&bar[cslot], cphase);
    const float4* rowp = reinterpret_cast<const float4*>(bufs + (size_t)cslot * row_bytes);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c4 = lane + 32 * j;
      if (c4 < C4) {
        const float4 x = rowp[c4];
        acc[j].x += x.x; acc[j].y += x.y; acc[j].z += x.z; acc[j].w += x.w;
      }
    }
    if (++cslot == Q) { cslot = 0; cphase ^= 1; }
    ++cons;
    __syncwarp();
    if (issue < e1) issue_one();
    if (--left == 0) {
      left = deg;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c4 = lane + 32 * j;
        if (c4 < C4) reinterpret_cast<float4*>(out + seg * D)[c4] = acc[j];
        acc[j] = make_float4(0, 0, 0, 0);
      }
      ++seg;
    }
  }
}

// ---- variant 4: rolling ring fed by cp.async (LDGSTS), completion by wait_group ---------------------
// Q row slots per warp; every lane copies its own 16 B columns of a row (NV cp.async per row), one commit group
// per row; cp.async.wait_group Q-1 returns when the oldest row has landed.  No mbarrier, no uniform-register
// bulk-copy issue: the per-row critical path is LDGSTS issue + wait_group + LDS + FADD.
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, uint64_t pol) {
  (void)pol;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
template <int NV, int Q>
__global__ void rollcp_gather(const float* __restrict__ h, int D, const int* __restrict__ src, long long nseg, int deg,
                              float* __restrict__ out, int warps_per_cta) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)D * 4;
  const uint32_t buf0 = smem_u32(smem) + (uint32_t)w * Q * row_bytes + (uint32_t)lane * 16u;
  const uint64_t pol = policy_evict_first();
  const long long warp = (long long)blockIdx.x * warps_per_cta + w;
  const long long nwarps = (long long)gridDim.x * warps_per_cta;
  const long long per = (nseg + nwarps - 1) / nwarps;
  const long long s0 = warp * per, s1 = min(nseg, s0 + per);
  if (s0 >= s1) return;
  const long long e0 = s0 * deg, e1 = s1 * deg;
  const int C4 = D >> 2;
  float4 acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = make_float4(0, 0, 0, 0);
  long long issue = e0, cons = e0, blk = e0;
  int ids = blk + lane < e1 ? __ldg(src + blk + lane) : 0;
  int ids_next = blk + 32 + lane < e1 ? __ldg(src + blk + 32 + lane) : 0;
  int islot = 0, cslot = 0, in_blk = 0, left = deg;
  uint32_t ibuf = buf0, cbuf = buf0;
  long long seg = s0;
  auto issue_one = [&]() {
    if (issue < e1) {
      if (in_blk == 32) {
        in_blk = 0;
        blk += 32;
        ids = ids_next;
        ids_next = blk + 32 + lane < e1 ? __ldg(src + blk + 32 + lane) : 0;
      }
      const int s = __shfl_sync(0xffffffffu, ids, in_blk);
      const float* rp = h + (long long)s * D + 4 * lane;
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (lane + 32 * j < C4) cp_async16(ibuf + 512u * j, rp + 128 * j, pol);
      ++in_blk;
      ++issue;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // (possibly empty) group: keeps the group count in step
    if (++islot == Q) { islot = 0; ibuf = buf0; } else ibuf += row_bytes;
  };
#pragma unroll 1
  for (int i = 0; i < Q; ++i) issue_one();
  while (cons < e1) {
    asm volatile("cp.async.wait_group %0;" ::"n"(Q - 1) : "memory");
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (lane + 32 * j < C4) {
        float4 x;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(cbuf + 512u * j) : "memory");
        acc[j].x += x.x; acc[j].y += x.y; acc[j].z += x.z; acc[j].w += x.w;
      }
    }
    if (++cslot == Q) { cslot = 0; cbuf = buf0; } else cbuf += row_bytes;
    ++cons;
    issue_one();
    if (--left == 0) {
      left = deg;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c4 = lane + 32 * j;
        if (c4 < C4) reinterpret_cast<float4*>(out + seg * D)[c4] = acc[j];
        acc[j] = make_float4(0, 0, 0, 0);
      }
      ++seg;
    }
  }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

int main(int argc, char** argv) {
  const long long V = argc > 1 ? atoll(argv[1]) : 1000000;
  const int D = argc > 2 ? atoi(argv[2]) : 256;
  const long long E = argc > 3 ? atoll(argv[3]) : 20000000;
  const int deg = argc > 4 ? atoi(argv[4]) : 20;
  const long long nseg = E / deg;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float *h, *out;
  int* src;
  CK(cudaMalloc(&h, (size_t)V * D * 4));
  CK(cudaMalloc(&out, (size_t)nseg * D * 4));
  CK(cudaMalloc(&src, (size_t)E * 4));
  CK(cudaMemset(h, 0, (size_t)V * D * 4));
  {
    std::vector<int> s((size_t)E);
    uint64_t x = 88172645463325252ull;
    for (long long i = 0; i < E; ++i) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      s[(size_t)i] = (int)(x % (uint64_t)V);
    }
    CK(cudaMemcpy(src, s.data(), (size_t)E * 4, cudaMemcpyHostToDevice));
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const double bytes = (double)E * D * 4 + (double)nseg * D * 4 + (double)E * 4;
  printf("# V=%lld D=%d (row %d B) E=%lld deg=%d  bytes/launch %.2f GB  SMs=%d\n", V, D, D * 4, E, deg, bytes / 1e9, sms);
  printf("# variant, warps/SM, in-flight rows per warp, ms, GB/s\n");
  const int reps = 5;
  const char* only = argc > 5 ? argv[5] : nullptr;
  auto want = [&](const char* name) { return only == nullptr || strcmp(only, name) == 0; };
  auto report = [&](const char* name, int wps, int inflight, float ms) {
    printf("%s, %d, %d, %.3f, %.0f\n", name, wps, inflight, ms, bytes / (ms * 1e-3) / 1e9);
    fflush(stdout);
  };
  const int nv = (D + 127) / 128;
#define RUN_LDG(NV, U, WPS)                                                                      \
  {                                                                                              \
    const int threads = 256;                                                                     \
    const int grid = sms * (WPS) / 8;                                                            \
    ldg_gather<NV, U><<<grid, threads>>>(h, D, src, nseg, deg, out);                             \
    CK(cudaDeviceSynchronize());                                                                 \
    CK(cudaEventRecord(e0));                                                                     \
    for (int r = 0; r < reps; ++r) ldg_gather<NV, U><<<grid, threads>>>(h, D, src, nseg, deg, out); \
    CK(cudaEventRecord(e1));                                                                     \
    CK(cudaEventSynchronize(e1));                                                                \
    report("ldg", WPS, U, time_ms(e0, e1) / reps);                                               \
  }
  if (!want("ldg")) {
  } else if (nv == 2) {
    RUN_LDG(2, 2, 16) RUN_LDG(2, 4, 16) RUN_LDG(2, 8, 16)
    RUN_LDG(2, 2, 32) RUN_LDG(2, 4, 32) RUN_LDG(2, 8, 32)
    RUN_LDG(2, 2, 64) RUN_LDG(2, 4, 64)
  } else if (nv == 3) {
    RUN_LDG(3, 2, 16) RUN_LDG(3, 4, 16)
    RUN_LDG(3, 2, 32) RUN_LDG(3, 4, 32)
    RUN_LDG(3, 2, 64)
  }
#define RUN_BULK(NV, R, W)                                                                       \
  {                                                                                              \
    const size_t smem = 1024 + (size_t)(W) * 2 * (R) * D * 4;                                    \
    if (smem <= 227 * 1024) {                                                                    \
      CK(cudaFuncSetAttribute(bulk_gather<NV, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      bulk_gather<NV, R><<<sms, 32 * (W), smem>>>(h, D, src, nseg, deg, out, W);                 \
      CK(cudaDeviceSynchronize());                                                               \
      CK(cudaEventRecord(e0));                                                                   \
      for (int r = 0; r < reps; ++r) bulk_gather<NV, R><<<sms, 32 * (W), smem>>>(h, D, src, nseg, deg, out, W); \
      CK(cudaEventRecord(e1));                                                                   \
      CK(cudaEventSynchronize(e1));                                                              \
      report("bulk", W, 2 * (R), time_ms(e0, e1) / reps);                                        \
    }                                                                                            \
  }
  if (!want("bulk")) {
  } else if (nv == 2) {
    RUN_BULK(2, 4, 8) RUN_BULK(2, 8, 8) RUN_BULK(2, 4, 16) RUN_BULK(2, 6, 16) RUN_BULK(2, 2, 32) RUN_BULK(2, 3, 32)
    RUN_BULK(2, 2, 16) RUN_BULK(2, 1, 32)
  } else if (nv == 3) {
    RUN_BULK(3, 4, 8) RUN_BULK(3, 8, 8) RUN_BULK(3, 4, 16) RUN_BULK(3, 2, 32)
  }
#define RUN_ROLL(NV, Q, W)                                                                       \
  {                                                                                              \
    const size_t smem = 4096 + (size_t)(W) * (Q) * D * 4;                                        \
    if (smem <= 227 * 1024) {                                                                    \
      CK(cudaFuncSetAttribute(roll_gather<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      roll_gather<NV><<<sms, 32 * (W), smem>>>(h, D, src, nseg, deg, out, W, Q);                 \
      CK(cudaDeviceSynchronize());                                                               \
      CK(cudaEventRecord(e0));                                                                   \
      for (int r = 0; r < reps; ++r) roll_gather<NV><<<sms, 32 * (W), smem>>>(h, D, src, nseg, deg, out, W, Q); \
      CK(cudaEventRecord(e1));                                                                   \
      CK(cudaEventSynchronize(e1));                                                              \
      CK(cudaGetLastError());                                                                    \
      report("roll", W, Q, time_ms(e0, e1) / reps);                                              \
    }                                                                                            \
  }
  if (!want("roll")) {
  } else if (nv == 2) {
    RUN_ROLL(2, 4, 16) RUN_ROLL(2, 6, 16) RUN_ROLL(2, 8, 16) RUN_ROLL(2, 12, 16)
    RUN_ROLL(2, 8, 8) RUN_ROLL(2, 12, 8) RUN_ROLL(2, 16, 8) RUN_ROLL(2, 24, 8)
    RUN_ROLL(2, 24, 4) RUN_ROLL(2, 32, 4) RUN_ROLL(2, 32, 6)
  } else if (nv == 3) {
    RUN_ROLL(3, 4, 16) RUN_ROLL(3, 6, 16) RUN_ROLL(3, 8, 16) RUN_ROLL(3, 10, 8) RUN_ROLL(3, 16, 8) RUN_ROLL(3, 32, 4)
  }
#define RUN_ROLLCP(NV, Q, W)                                                                     \
  {                                                                                              \
    const size_t smem = (size_t)(W) * (Q) * D * 4;                                               \
    if (smem <= 227 * 1024) {                                                                    \
      CK(cudaFuncSetAttribute(rollcp_gather<NV, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      rollcp_gather<NV, Q><<<sms, 32 * (W), smem>>>(h, D, src, nseg, deg, out, W);               \
      CK(cudaDeviceSynchronize());                                                               \
      CK(cudaEventRecord(e0));                                                                   \
      for (int r = 0; r < reps; ++r) rollcp_gather<NV, Q><<<sms, 32 * (W), smem>>>(h, D, src, nseg, deg, out, W); \
      CK(cudaEventRecord(e1));                                                                   \
      CK(cudaEventSynchronize(e1));                                                              \
      CK(cudaGetLastError());                                                                    \
      report("rollcp", W, Q, time_ms(e0, e1) / reps);                                            \
    }                                                                                            \
  }
  if (!want("rollcp")) {
  } else if (nv == 2) {
    RUN_ROLLCP(2, 4, 16) RUN_ROLLCP(2, 6, 16) RUN_ROLLCP(2, 8, 16)
    RUN_ROLLCP(2, 4, 8) RUN_ROLLCP(2, 8, 8) RUN_ROLLCP(2, 16, 8)
    RUN_ROLLCP(2, 16, 4) RUN_ROLLCP(2, 32, 4)
  } else if (nv == 3) {
    RUN_ROLLCP(3, 4, 16) RUN_ROLLCP(3, 6, 16) RUN_ROLLCP(3, 8, 8) RUN_ROLLCP(3, 16, 8)
  }
  return 0;
}

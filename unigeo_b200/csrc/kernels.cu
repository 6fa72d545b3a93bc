// HBM-bound kernels (see kernels.cuh).  128-bit accesses, fp32 math, 16-bit storage.
#include "kernels.cuh"
#include "ptx.cuh"

// bisecting aid: -DUG_NOTRIG_GN / _LN / _ATT / _MISC build variants in which a kernel family never releases its
// dependents early (they start when it completes)
#ifdef UG_NOTRIG_GN
#define UG_TRIGGER_GN() do {} while (0)
#else
#define UG_TRIGGER_GN() pdl_launch_dependents()
#endif
#ifdef UG_NOTRIG_LN
#define UG_TRIGGER_LN() do {} while (0)
#else
#define UG_TRIGGER_LN() pdl_launch_dependents()
#endif
#ifdef UG_NOTRIG_ATT
#define UG_TRIGGER_ATT() do {} while (0)
#else
#define UG_TRIGGER_ATT() pdl_launch_dependents()
#endif
#ifdef UG_NOTRIG_MISC
#define UG_TRIGGER_MISC() do {} while (0)
#else
#define UG_TRIGGER_MISC() pdl_launch_dependents()
#endif

#include <atomic>
#include <map>
#include <mutex>

namespace ug {
namespace {

#define UG_DISPATCH_FMT(fmt, ...)                         \
  do {                                                    \
    if ((fmt) == 1) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { using T = __half; __VA_ARGS__; }               \
  } while (0)

template <typename T> __device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = Elem<T>::unpack2(u.x), b = Elem<T>::unpack2(u.y), c = Elem<T>::unpack2(u.z), d = Elem<T>::unpack2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <typename T> __device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = Elem<T>::pack2(f[0], f[1]); u.y = Elem<T>::pack2(f[2], f[3]);
  u.z = Elem<T>::pack2(f[4], f[5]); u.w = Elem<T>::pack2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ float silu_f(float x) { return x * rcp_approx(1.0f + __expf(-x)); }
// pairs on the packed fp32 pipe (FFMA2 / FMUL2 / FADD2, sm_100): the normalisation kernels are instruction-issue
// bound before they are HBM bound (ncu: 47-54 % issue slots at 15-20 % DRAM), so halving the FP instruction count
// is worth more than any further load tuning
template <typename T> __device__ __forceinline__ void unpack4x2(const uint4& u, float2 (&p)[4]) {
  p[0] = Elem<T>::unpack2(u.x); p[1] = Elem<T>::unpack2(u.y); p[2] = Elem<T>::unpack2(u.z); p[3] = Elem<T>::unpack2(u.w);
}
template <typename T> __device__ __forceinline__ uint4 pack4x2(const float2 (&p)[4]) {
  uint4 u;
  u.x = Elem<T>::pack2(p[0].x, p[0].y); u.y = Elem<T>::pack2(p[1].x, p[1].y);
  u.z = Elem<T>::pack2(p[2].x, p[2].y); u.w = Elem<T>::pack2(p[3].x, p[3].y);
  return u;
}
__device__ __forceinline__ float2 silu2(float2 v) {        // v / (1 + 2^(-v log2 e))
  const float2 t = __fmul2_rn(v, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 d = __fadd2_rn(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(1.0f, 1.0f));
  return __fmul2_rn(v, make_float2(rcp_approx(d.x), rcp_approx(d.y)));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// vector c8 of the (virtually concatenated) row
__device__ __forceinline__ const uint4* cat_ptr(const void* x1, int nv1, const void* x2, int nv2, long long row,
                                                int c8) {
  return c8 < nv1 ? reinterpret_cast<const uint4*>(x1) + row * nv1 + c8
                  : reinterpret_cast<const uint4*>(x2) + row * nv2 + (c8 - nv1);
}

// ------------------------------------------------------------------ GroupNorm statistics
// Deterministic two-level reduction (no atomics): every CTA reduces its row chunk in a fixed
// order and writes one (sum, sumsq) pair per group to part[set][chunk][G][2]; the apply
// kernel adds the chunks in index order.  Reruns are bit-identical.
template <typename T>
__global__ void __launch_bounds__(320, 3) gn_stats_kernel(const void* __restrict__ x1, int nv1, const void* __restrict__ x2, int nv2,
                                long long rows_per_set, long long chunk_rows, int G, int cs,
                                float* __restrict__ part, float* __restrict__ mr, unsigned int* __restrict__ counters,
                                float inv_cnt, float eps) {
  UG_TRIGGER_GN();
  pdl_wait();
  extern __shared__ float s_part[];   // [rpb][C] sums, then [rpb][C] sumsq
  __shared__ int s_last;
  const int nvec = nv1 + nv2;
  const int C = nvec * 8;
  const int rpb = blockDim.x / nvec;
  const int c8 = threadIdx.x % nvec;
  const int rr = threadIdx.x / nvec;
  const long long set = blockIdx.y;
  const long long r_begin = (long long)blockIdx.x * chunk_rows;
  long long r_end = r_begin + chunk_rows;
  if (r_end > rows_per_set) r_end = rows_per_set;
  float a[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = 0.f; q[j] = 0.f; }
  if (rr < rpb) {
    long long r = r_begin + rr;
    // eight independent 16-byte loads in flight per thread (the kernel is latency bound: a thread only
    // makes a few trips); rows past the end are predicated off and contribute zeros
    for (; r < r_end; r += 8 * rpb) {
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const long long rk = r + (long long)k * rpb;
        u[k] = rk < r_end ? ld_act(cat_ptr(x1, nv1, x2, nv2, set * rows_per_set + rk, c8)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float f[8];
        unpack8<T>(u[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
      }
    }
    float* ps = s_part + (size_t)rr * C + c8 * 8;
    float* pq = s_part + (size_t)(rpb + rr) * C + c8 * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) { ps[j] = a[j]; pq[j] = q[j]; }
  }
  __syncthreads();
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    float sa = 0.f, sq = 0.f;
    for (int r = 0; r < rpb; ++r)
      for (int cc = g * cs; cc < (g + 1) * cs; ++cc) {
        sa += s_part[(size_t)r * C + cc];
        sq += s_part[(size_t)(rpb + r) * C + cc];
      }
    float* o = part + ((set * gridDim.x + blockIdx.x) * G + g) * 2;
    o[0] = sa;
    o[1] = sq;
  }
  // ---- the last CTA of this set folds the partials into (mean, rstd): fixed summation order, so
  // the result does not depend on which CTA happens to be last
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(&counters[set], 1u);
    s_last = (ticket == gridDim.x - 1);
    if (s_last) counters[set] = 0u;          // ready for the next GroupNorm on this stream
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int slices = blockDim.x / G;         // >= 2
  const int g = threadIdx.x % G, k = threadIdx.x / G;
  float* red = s_part;                       // reuse: [slices][G][2]
  if (k < slices) {
    float sa = 0.f, sq = 0.f;                // eight independent L2 loads in flight, additions in chunk order
    const int n = (int)gridDim.x;
    const float2* base = reinterpret_cast<const float2*>(part) + (set * n) * G + g;
    int ch = k;
    for (; ch + 7 * slices < n; ch += 8 * slices) {
      float2 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldcg(base + (long long)(ch + j * slices) * G);
#pragma unroll
      for (int j = 0; j < 8; ++j) { sa += v[j].x; sq += v[j].y; }
    }
    for (; ch < n; ch += slices) {
      const float2 v = __ldcg(base + (long long)ch * G);
      sa += v.x;
      sq += v.y;
    }
    red[(k * G + g) * 2 + 0] = sa;
    red[(k * G + g) * 2 + 1] = sq;
  }
  __syncthreads();
  if (threadIdx.x < G) {
    float a = 0.f, q = 0.f;
    for (int i = 0; i < slices; ++i) { a += red[(i * G + g) * 2]; q += red[(i * G + g) * 2 + 1]; }
    const float mean = a * inv_cnt;
    const float var = fmaxf(q * inv_cnt - mean * mean, 0.f);
    mr[(set * G + g) * 2 + 0] = mean;
    mr[(set * G + g) * 2 + 1] = rsqrtf(var + eps);
  }
}

// ------------------------------------------------------------------ GroupNorm, one launch
// statistics + apply in ONE kernel: every CTA reduces its row chunk (phase 1), the last CTA of a set (integer
// ticket) folds the partials in index order and publishes (mean, rstd) with a release store; all CTAs of the set
// wait on that flag (acquire) and then normalise the SAME chunk (phase 2), which they just pulled through L2.
// Saves a launch and the second HBM read of x per GroupNorm (210 GroupNorms per cfg2 denoising step).
// Needs the whole grid co-resident: the launcher sizes it from the occupancy query; earlier kernels on the
// stream never depend on this one, so CTAs that start late only delay the flag, they cannot deadlock it.
// counters: [0,kGnMaxSets) tickets, [kGnMaxSets, 2k) flags, [2k, 3k) done counts; all zero between launches.
template <typename T>
__global__ void __launch_bounds__(320, 3)
gn_fused_kernel(const void* __restrict__ x1, int nv1, const void* __restrict__ x2, int nv2, long long rows_per_set,
                long long chunk_rows, int G, int cs, float* __restrict__ part, float* __restrict__ mr,
                unsigned int* __restrict__ counters, float inv_cnt, float eps, const float* __restrict__ gamma,
                const float* __restrict__ beta, int silu, void* __restrict__ y) {
  UG_TRIGGER_GN();
  pdl_wait();
  extern __shared__ float s_dyn[];
  __shared__ int s_last;
  const int nvec = nv1 + nv2;
  const int C = nvec * 8;
  const int rpb = blockDim.x / nvec;
  const int c8 = threadIdx.x % nvec;
  const int rr = threadIdx.x / nvec;
  const long long set = blockIdx.y;
  const long long r_begin = (long long)blockIdx.x * chunk_rows;
  long long r_end = r_begin + chunk_rows;
  if (r_end > rows_per_set) r_end = rows_per_set;
  // ---- phase 1: partial (sum, sumsq) per group of this chunk
  {
    float* s_part = s_dyn;               // [rpb][C] sums, then [rpb][C] sumsq
    float2 a[4], q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { a[j] = make_float2(0.f, 0.f); q[j] = make_float2(0.f, 0.f); }
    if (rr < rpb) {
      for (long long r = r_begin + rr; r < r_end; r += 8 * rpb) {
        uint4 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const long long rk = r + (long long)k * rpb;
          u[k] = rk < r_end ? ld_act(cat_ptr(x1, nv1, x2, nv2, set * rows_per_set + rk, c8)) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float2 f[4];
          unpack4x2<T>(u[k], f);
#pragma unroll
          for (int j = 0; j < 4; ++j) { a[j] = __fadd2_rn(a[j], f[j]); q[j] = __ffma2_rn(f[j], f[j], q[j]); }
        }
      }
      float2* ps = reinterpret_cast<float2*>(s_part + (size_t)rr * C + c8 * 8);
      float2* pq = reinterpret_cast<float2*>(s_part + (size_t)(rpb + rr) * C + c8 * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) { ps[j] = a[j]; pq[j] = q[j]; }
    }
    __syncthreads();
    if (threadIdx.x < G) {
      const int g = threadIdx.x;
      float sa = 0.f, sq = 0.f;
      for (int r = 0; r < rpb; ++r)
        for (int cc = g * cs; cc < (g + 1) * cs; ++cc) {
          sa += s_part[(size_t)r * C + cc];
          sq += s_part[(size_t)(rpb + r) * C + cc];
        }
      float* o = part + ((set * gridDim.x + blockIdx.x) * G + g) * 2;
      o[0] = sa;
      o[1] = sq;
    }
  }
  __threadfence();
  __syncthreads();
  unsigned int* flag = counters + kGnMaxSets + set;
  unsigned int* done = counters + 2 * kGnMaxSets + set;
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(&counters[set], 1u);
    s_last = (ticket == gridDim.x - 1);
    if (s_last) counters[set] = 0u;
  }
  __syncthreads();
  if (s_last) {                          // fixed summation order: the result does not depend on which CTA is last
    __threadfence();
    const int slices = blockDim.x / G;
    const int g = threadIdx.x % G, k = threadIdx.x / G;
    float* red = s_dyn;
    if (k < slices) {
      // eight independent L2 loads in flight (a plain loop serialises one L2 round trip per chunk: 12-25 us for
      // 300-600 chunks); additions stay in chunk order, so the result is unchanged and deterministic
      float sa = 0.f, sq = 0.f;
      const int n = (int)gridDim.x;
      const float2* base = reinterpret_cast<const float2*>(part) + (set * n) * G + g;
      int ch = k;
      for (; ch + 7 * slices < n; ch += 8 * slices) {
        float2 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcg(base + (long long)(ch + j * slices) * G);
#pragma unroll
        for (int j = 0; j < 8; ++j) { sa += v[j].x; sq += v[j].y; }
      }
      for (; ch < n; ch += slices) {
        const float2 v = __ldcg(base + (long long)ch * G);
        sa += v.x;
        sq += v.y;
      }
      red[(k * G + g) * 2 + 0] = sa;
      red[(k * G + g) * 2 + 1] = sq;
    }
    __syncthreads();
    if (threadIdx.x < G) {
      float a = 0.f, q = 0.f;
      for (int i = 0; i < slices; ++i) { a += red[(i * G + g) * 2]; q += red[(i * G + g) * 2 + 1]; }
      const float mean = a * inv_cnt;
      const float var = fmaxf(q * inv_cnt - mean * mean, 0.f);
      mr[(set * G + g) * 2 + 0] = mean;
      mr[(set * G + g) * 2 + 1] = rsqrtf(var + eps);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(1u) : "memory");
  }
  // gamma / beta of this thread's channels are fetched before the wait: one L2 round trip off the critical path
  constexpr int kGB = 8;                                  // C <= 2560, blockDim >= 320 there: at most 8 channels each
  float gpre[kGB], bpre[kGB];
#pragma unroll
  for (int i = 0; i < kGB; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    gpre[i] = c < C ? __ldg(gamma + c) : 0.f;
    bpre[i] = c < C ? __ldg(beta + c) : 0.f;
  }
  if (threadIdx.x == 0) {
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if (v == 0u) __nanosleep(32);
    } while (v == 0u);
  }
  __syncthreads();
  // ---- phase 2: y = act((x - mean) * rstd * gamma + beta) over the same chunk (x comes back from L2)
  {
    float* s_a = s_dyn;
    float* s_b = s_dyn + C;
    float* s_mean = s_b + C;
    float* s_rstd = s_mean + G;
    if (threadIdx.x < G) {
      s_mean[threadIdx.x] = __ldcg(mr + (set * G + threadIdx.x) * 2 + 0);
      s_rstd[threadIdx.x] = __ldcg(mr + (set * G + threadIdx.x) * 2 + 1);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kGB; ++i) {
      const int c = threadIdx.x + i * blockDim.x;
      if (c < C) {
        const int g = c / cs;
        const float ga = gpre[i] * s_rstd[g];
        s_a[c] = ga;
        s_b[c] = bpre[i] - s_mean[g] * ga;
      }
    }
    __syncthreads();
    if (rr < rpb) {
      float2 sa[4], sb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sa[j] = reinterpret_cast<const float2*>(s_a + c8 * 8)[j];
        sb[j] = reinterpret_cast<const float2*>(s_b + c8 * 8)[j];
      }
      uint4* yo = reinterpret_cast<uint4*>(y);
      for (long long r = r_begin + rr; r < r_end; r += 4 * rpb) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const long long rk = r + (long long)k * rpb;
          u[k] = rk < r_end ? ld_act(cat_ptr(x1, nv1, x2, nv2, set * rows_per_set + rk, c8)) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const long long rk = r + (long long)k * rpb;
          float2 f0[4];
          unpack4x2<T>(u[k], f0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 v0 = __ffma2_rn(f0[j], sa[j], sb[j]);
            f0[j] = silu ? silu2(v0) : v0;
          }
          if (rk < r_end) yo[(set * rows_per_set + rk) * nvec + c8] = pack4x2<T>(f0);
        }
      }
    }
  }
  // ---- the last CTA to finish re-arms flag and counters for the next GroupNorm on this stream
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int d = atomicAdd(done, 1u);
    if (d == gridDim.x - 1) {
      *done = 0u;
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(0u) : "memory");
    }
  }
}

// ------------------------------------------------------------------ GroupNorm over a thread-block cluster
// For statistics sets that fit the shared memory of one cluster (<= 8 CTAs x 64 KB of rows: the per-frame GroupNorms
// of the UNet's two lowest resolutions): each CTA pulls its rows from HBM ONCE into shared memory while it accumulates
// its partial sums, the cluster exchanges the per-group partials through distributed shared memory (every CTA folds
// them in rank order: deterministic, identical on all ranks), and the rows are normalised straight out of shared
// memory.  No global partials, no tickets or flags, no second read of x: the one-launch kernel above spends 19-33 us
// on a 3-25 MB tensor in its launch -> partials -> ticket -> fold -> flag -> apply chain.
template <typename T>
__global__ void __launch_bounds__(320)
gn_cluster_kernel(const void* __restrict__ x1, int nv1, const void* __restrict__ x2, int nv2, long long rows_per_set,
                  int chunk_rows, int G, int cs, float inv_cnt, float eps, const float* __restrict__ gamma,
                  const float* __restrict__ beta, int silu, void* __restrict__ y) {
  UG_TRIGGER_GN();
  extern __shared__ __align__(16) unsigned char s_cl[];
  const int nvec = nv1 + nv2;
  const int C = nvec * 8;
  const int rpb = blockDim.x / nvec;
  const int c8 = threadIdx.x % nvec;
  const int rr = threadIdx.x / nvec;
  const unsigned int S = gridDim.x;                     // the cluster spans the grid's x extent
  const unsigned int rank = blockIdx.x;
  const long long set = blockIdx.y;
  const long long r_begin = (long long)rank * chunk_rows;
  long long r_stop = r_begin + chunk_rows;
  if (r_stop > rows_per_set) r_stop = rows_per_set;
  const int nrows = r_stop > r_begin ? (int)(r_stop - r_begin) : 0;
  uint4* tile = reinterpret_cast<uint4*>(s_cl);                                   // [chunk_rows][nvec]
  float* s_part = reinterpret_cast<float*>(s_cl + (size_t)chunk_rows * nvec * 16); // [2][rpb][C]; later s_a | s_b
  float* s_stat = s_part + (size_t)2 * rpb * C;                                   // [G][2], read by the peers
  float* s_mean = s_stat + 2 * G;
  float* s_rstd = s_mean + G;
  // the affine parameters are weights, not the previous kernel's output: fetch them ahead of the dependency wait
  constexpr int kGB = 8;
  float gpre[kGB], bpre[kGB];
#pragma unroll
  for (int i = 0; i < kGB; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    gpre[i] = c < C ? __ldg(gamma + c) : 0.f;
    bpre[i] = c < C ? __ldg(beta + c) : 0.f;
  }
  pdl_wait();
  // ---- phase 1: rows -> shared memory, partial (sum, sumsq) per channel
  {
    float2 a[4], q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { a[j] = make_float2(0.f, 0.f); q[j] = make_float2(0.f, 0.f); }
    if (rr < rpb) {
      for (int r = rr; r < nrows; r += 8 * rpb) {
        uint4 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rk = r + k * rpb;
          u[k] = rk < nrows ? ld_act(cat_ptr(x1, nv1, x2, nv2, set * rows_per_set + r_begin + rk, c8))
                            : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rk = r + k * rpb;
          if (rk < nrows) tile[(size_t)rk * nvec + c8] = u[k];
          float2 f[4];
          unpack4x2<T>(u[k], f);
#pragma unroll
          for (int j = 0; j < 4; ++j) { a[j] = __fadd2_rn(a[j], f[j]); q[j] = __ffma2_rn(f[j], f[j], q[j]); }
        }
      }
      float2* ps = reinterpret_cast<float2*>(s_part + (size_t)rr * C + c8 * 8);
      float2* pq = reinterpret_cast<float2*>(s_part + (size_t)(rpb + rr) * C + c8 * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) { ps[j] = a[j]; pq[j] = q[j]; }
    }
  }
  __syncthreads();
  {
    // L lanes per group (fixed strided split + xor tree: the same order on every run)
    int L = 8;
    while (L > 1 && G * L > (int)blockDim.x) L >>= 1;
    const int g = threadIdx.x / L, l = threadIdx.x % L;
    float sa = 0.f, sq = 0.f;
    if (g < G) {
      const int per_group = rpb * cs;
      for (int e = l; e < per_group; e += L) {
        const int r = e / cs, cc = g * cs + (e - r * cs);
        sa += s_part[(size_t)r * C + cc];
        sq += s_part[(size_t)(rpb + r) * C + cc];
      }
    }
    for (int o = 4; o > 0; o >>= 1) {
      const float ta = __shfl_xor_sync(0xffffffffu, sa, o), tq = __shfl_xor_sync(0xffffffffu, sq, o);
      if (o < L) { sa += ta; sq += tq; }
    }
    if (g < G && l == 0) {
      s_stat[2 * g] = sa;
      s_stat[2 * g + 1] = sq;
    }
  }
  cluster_arrive_release();                             // partials of this CTA are visible to the cluster ...
  cluster_wait_acquire();                               // ... and everybody else's are visible here
  if (threadIdx.x < G) {
    float2 v[8];                                        // all remote reads in flight at once (S <= 8)
#pragma unroll
    for (unsigned int r = 0; r < 8; ++r)
      v[r] = r < S ? ld_dsmem_f2(s_stat + 2 * threadIdx.x, r) : make_float2(0.f, 0.f);
    float sa = 0.f, sq = 0.f;
#pragma unroll
    for (unsigned int r = 0; r < 8; ++r) {              // rank order on every CTA: one result, bit for bit
      sa += v[r].x;
      sq += v[r].y;
    }
    const float mean = sa * inv_cnt;
    const float var = fmaxf(sq * inv_cnt - mean * mean, 0.f);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + eps);
  }
  cluster_arrive_release();                             // done reading the peers (waited for before exit)
  __syncthreads();
  // ---- phase 2: y = act((x - mean) * rstd * gamma + beta) out of shared memory
  float* s_a = s_part;
  float* s_b = s_part + C;
#pragma unroll
  for (int i = 0; i < kGB; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < C) {
      const int g = c / cs;
      const float ga = gpre[i] * s_rstd[g];
      s_a[c] = ga;
      s_b[c] = bpre[i] - s_mean[g] * ga;
    }
  }
  __syncthreads();
  if (rr < rpb) {
    float2 sa[4], sb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sa[j] = reinterpret_cast<const float2*>(s_a + c8 * 8)[j];
      sb[j] = reinterpret_cast<const float2*>(s_b + c8 * 8)[j];
    }
    uint4* yo = reinterpret_cast<uint4*>(y) + (set * rows_per_set + r_begin) * nvec + c8;
    for (int r = rr; r < nrows; r += rpb) {
      float2 f0[4];
      unpack4x2<T>(tile[(size_t)r * nvec + c8], f0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 v0 = __ffma2_rn(f0[j], sa[j], sb[j]);
        f0[j] = silu ? silu2(v0) : v0;
      }
      yo[(size_t)r * nvec] = pack4x2<T>(f0);
    }
  }
  cluster_wait_acquire();                               // a CTA's shared memory must outlive its peers' reads
}

// ------------------------------------------------------------------ GroupNorm finalize
// one CTA per statistics set: (mean, rstd) per group from the per-chunk partials, summed in a fixed
// order (thread k adds chunks k, k+8, ...; the 8 lanes of a group are then folded in lane order).
__global__ void gn_finalize_kernel(const float* __restrict__ part, int chunks, int G, float inv_cnt, float eps,
                                   float* __restrict__ mr) {
  __shared__ float s_s[8][64], s_q[8][64];
  const long long set = blockIdx.x;
  const int g = threadIdx.x % G;
  const int k = threadIdx.x / G;          // 0..7
  float sa = 0.f, sq = 0.f;
  for (int ch = k; ch < chunks; ch += 8) {
    const float* o = part + ((set * chunks + ch) * G + g) * 2;
    sa += o[0];
    sq += o[1];
  }
  s_s[k][g] = sa;
  s_q[k][g] = sq;
  __syncthreads();
  if (k == 0) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += s_s[i][g]; q += s_q[i][g]; }
    const float mean = a * inv_cnt;
    const float var = fmaxf(q * inv_cnt - mean * mean, 0.f);
    mr[(set * G + g) * 2 + 0] = mean;
    mr[(set * G + g) * 2 + 1] = rsqrtf(var + eps);
  }
}

// ------------------------------------------------------------------ GroupNorm apply (+SiLU)
template <typename T>
__global__ void __launch_bounds__(320, 3) gn_apply_kernel(const void* __restrict__ x1, int nv1, const void* __restrict__ x2, int nv2,
                                long long rows_per_set, long long chunk_rows, int G, int cs,
                                const float* __restrict__ part /* [sets][G] (mean, rstd) */,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                void* __restrict__ y) {
  UG_TRIGGER_GN();
  pdl_wait();
  extern __shared__ float s_ab[];   // [C] scale, [C] shift, [G] mean, [G] rstd
  const int nvec = nv1 + nv2;
  const int C = nvec * 8;
  float* s_a = s_ab;
  float* s_b = s_ab + C;
  float* s_mean = s_b + C;
  float* s_rstd = s_mean + G;
  const long long set = blockIdx.y;
  if (threadIdx.x < G) {
    s_mean[threadIdx.x] = ld_act(part + (set * G + threadIdx.x) * 2 + 0);      // written by the previous kernel
    s_rstd[threadIdx.x] = ld_act(part + (set * G + threadIdx.x) * 2 + 1);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cs;
    const float ga = gamma[c] * s_rstd[g];
    s_a[c] = ga;
    s_b[c] = beta[c] - s_mean[g] * ga;
  }
  __syncthreads();
  const long long r_begin = (long long)blockIdx.x * chunk_rows;
  long long r_end = r_begin + chunk_rows;
  if (r_end > rows_per_set) r_end = rows_per_set;
  // thread = (row slot rr, channel vector c8): scale/shift of its 8 channels stay in registers
  const int rpb = blockDim.x / nvec;
  const int c8 = threadIdx.x % nvec;
  const int rr = threadIdx.x / nvec;
  if (rr >= rpb) return;
  float sa[8], sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sa[j] = s_a[c8 * 8 + j]; sb[j] = s_b[c8 * 8 + j]; }
  uint4* yo = reinterpret_cast<uint4*>(y);
  long long r = r_begin + rr;
  // six rows in flight per thread; rows past the end are predicated off
  for (; r < r_end; r += 6 * rpb) {
    uint4 u[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const long long rk = r + (long long)k * rpb;
      u[k] = rk < r_end ? ld_act(cat_ptr(x1, nv1, x2, nv2, set * rows_per_set + rk, c8)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const long long rk = r + (long long)k * rpb;
      float f0[8];
      unpack8<T>(u[k], f0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v0 = fmaf(f0[j], sa[j], sb[j]);
        f0[j] = silu ? silu_f(v0) : v0;
      }
      if (rk < r_end) yo[(set * rows_per_set + rk) * nvec + c8] = pack8<T>(f0);
    }
  }
}

// ------------------------------------------------------------------ LayerNorm
// LPR lanes share a row (32 / LPR rows per warp at once); lane l of a row owns the 16-byte vectors
// l, l + LPR, ... (VPL of them), so every lane is busy for C = 320 / 640 / 1280 (LPR = 8 / 16 / 32,
// VPL = 5) and a load instruction touches whole 128-byte lines.  The kernel is instruction-issue bound
// before it is HBM bound (ncu: 300 instructions per row in the one-warp-per-row form), hence: values are
// unpacked once and stay in registers, reductions take log2(LPR) shuffles, statistics are two-pass fp32.
template <typename T, int LPR, int VPL>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ x, long long rows, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, const float* __restrict__ add, int add_div,
                 void* __restrict__ y) {
  UG_TRIGGER_LN();
  pdl_wait();
  constexpr int RPW = 32 / LPR;                       // rows per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool rok = row < rows;
  const int nvec = C >> 3;
  const uint4* xb = reinterpret_cast<const uint4*>(x) + (rok ? row : 0) * nvec;
  uint4 raw[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = sub + LPR * k;
    raw[k] = make_uint4(0u, 0u, 0u, 0u);
    if (rok && i < nvec) raw[k] = ld_act(xb + i);
  }
  float2 f[VPL][4];
  const float* addr = (add && rok) ? add + (row / add_div) * C : nullptr;
  float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = sub + LPR * k;
    unpack4x2<T>(raw[k], f[k]);
    if (addr && i < nvec) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(addr + i * 8));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(addr + i * 8) + 1);
      f[k][0] = __fadd2_rn(f[k][0], make_float2(a0.x, a0.y)); f[k][1] = __fadd2_rn(f[k][1], make_float2(a0.z, a0.w));
      f[k][2] = __fadd2_rn(f[k][2], make_float2(a1.x, a1.y)); f[k][3] = __fadd2_rn(f[k][3], make_float2(a1.z, a1.w));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) s2 = __fadd2_rn(s2, f[k][j]);      // padding vectors are zero
  }
  float s = s2.x + s2.y;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float2 v2 = make_float2(0.f, 0.f);
  const float2 nm2 = make_float2(-mean, -mean);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    if (sub + LPR * k < nvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { f[k][j] = __fadd2_rn(f[k][j], nm2); v2 = __ffma2_rn(f[k][j], f[k][j], v2); }
    }
  }
  float v = v2.x + v2.y;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / (float)C + eps);
  if (!rok) return;
  uint4* yb = reinterpret_cast<uint4*>(y) + row * nvec;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = sub + LPR * k;
    if (i < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + i * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + i * 8) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + i * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + i * 8) + 1);
      const float2 r2 = make_float2(rstd, rstd);
      float2 o[4];
      o[0] = __ffma2_rn(__fmul2_rn(f[k][0], r2), make_float2(g0.x, g0.y), make_float2(b0.x, b0.y));
      o[1] = __ffma2_rn(__fmul2_rn(f[k][1], r2), make_float2(g0.z, g0.w), make_float2(b0.z, b0.w));
      o[2] = __ffma2_rn(__fmul2_rn(f[k][2], r2), make_float2(g1.x, g1.y), make_float2(b1.x, b1.y));
      o[3] = __ffma2_rn(__fmul2_rn(f[k][3], r2), make_float2(g1.z, g1.w), make_float2(b1.z, b1.w));
      yb[i] = pack4x2<T>(o);
    }
  }
}

// Row statistics only: what a GEMM with a folded LayerNorm (TapGemmArgs::ln_stat) needs of its A rows when the
// producing GEMM could not leave them behind.  Same lane mapping and two-pass arithmetic as layernorm_kernel; one
// read of x, 8 bytes written per row.
template <typename T, int LPR, int VPL>
__global__ void __launch_bounds__(256)
row_stats_kernel(const void* x, long long rows, int C, float eps, float2* __restrict__ stat) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  constexpr int RPW = 32 / LPR;
  const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool rok = row < rows;
  const int nvec = C >> 3;
  const uint4* xb = reinterpret_cast<const uint4*>(x) + (rok ? row : 0) * nvec;
  uint4 raw[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = sub + LPR * k;
    raw[k] = make_uint4(0u, 0u, 0u, 0u);
    if (rok && i < nvec) raw[k] = ld_act(xb + i);
  }
  float2 f[VPL][4];
  float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    unpack4x2<T>(raw[k], f[k]);
#pragma unroll
    for (int j = 0; j < 4; ++j) s2 = __fadd2_rn(s2, f[k][j]);      // padding vectors are zero
  }
  float s = s2.x + s2.y;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float2 v2 = make_float2(0.f, 0.f);
  const float2 nm2 = make_float2(-mean, -mean);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    if (sub + LPR * k < nvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 d = __fadd2_rn(f[k][j], nm2); v2 = __ffma2_rn(d, d, v2); }
    }
  }
  float v = v2.x + v2.y;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (rok && sub == 0) stat[row] = make_float2(mean, rsqrtf(v / (float)C + eps));
}

// ------------------------------------------------------------------ row softmax (block per row)
template <typename T, int VPT>
__global__ void softmax_rows_kernel(void* __restrict__ s, long long rows, int n, int n_valid, float scale_log2e) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const int nvec = n >> 3;
  uint4* p = reinterpret_cast<uint4*>(s) + row * nvec;
  float f[VPT][8];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int i = threadIdx.x + k * blockDim.x;
    if (i < nvec) {
      unpack8<T>(p[i], f[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[k][j] = (i * 8 + j < n_valid) ? f[k][j] * scale_log2e : -INFINITY;   // padded keys get probability 0
        m = fmaxf(m, f[k][j]);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < nw; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int i = threadIdx.x + k * blockDim.x;
    if (i < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { f[k][j] = exp2f(f[k][j] - m); sum += f[k][j]; }
    }
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < nw; ++w) sum += red[w];
  const float inv = 1.0f / sum;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int i = threadIdx.x + k * blockDim.x;
    if (i < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[k][j] *= inv;
      p[i] = pack8<T>(f[k]);
    }
  }
}

// ------------------------------------------------------------------ temporal attention
// One warp per (pixel, head): Q, K, V of its T <= 64 frames (64 dims) are staged in XOR-swizzled smem
// with cp.async, S = Q K^T and O = P V run on mma.sync m16n8k16 (fp32 accumulate) 16 queries at a
// time, softmax stays in the accumulator registers.  The problem is HBM bound (12 flop/byte): the
// tensor-core path only exists to get the arithmetic out of the way of the loads.
template <typename T> struct MmaK16;
template <> struct MmaK16<__half> {
  __device__ static __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
template <> struct MmaK16<__nv_bfloat16> {
  __device__ static __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// byte offset of 16-byte chunk `c` (0..7) of row `row` in a [rows][128 B] tile, XOR swizzled
__device__ __forceinline__ uint32_t swz(int row, int c) { return (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4); }

template <typename T, int TPAD>
__global__ void __launch_bounds__(128)
temporal_attn_kernel(const void* __restrict__ qkv, void* __restrict__ out, int Tn, long long P, int C,
                     float scale_log2e) {
  UG_TRIGGER_ATT();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t sm_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int heads = C >> 6;
  const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (item >= P * heads) return;
  const long long p = item / heads;
  const int h = (int)(item % heads);
  uint8_t* sq = sm_raw + (size_t)warp * 3 * TPAD * 128;
  const uint32_t aQ = smem_u32(sq), aK = aQ + TPAD * 128, aV = aK + TPAD * 128;
  const uint16_t* base = reinterpret_cast<const uint16_t*>(qkv);
  const long long row_elems = (long long)3 * C;

  // ---- stage Q | K | V rows (16-byte chunks), zero the padding rows
  for (int i = lane; i < 3 * TPAD * 8; i += 32) {
    const int c = i & 7, row = (i >> 3) % TPAD, which = i / (8 * TPAD);
    const uint32_t dst = aQ + (uint32_t)which * TPAD * 128 + swz(row, c);
    if (row < Tn) {
      const uint16_t* src = base + ((long long)row * P + p) * row_elems + which * C + h * 64 + c * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    } else {
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  constexpr int NT = TPAD / 8;                      // key tiles of 8
  const int qr = lane >> 2, qc = (lane & 3) * 2;    // accumulator row / column pair of this lane
#pragma unroll 1
  for (int mt = 0; mt * 16 < Tn; ++mt) {
    float sacc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) { sacc[n][0] = sacc[n][1] = sacc[n][2] = sacc[n][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {                 // 16 dims per step
      uint32_t af[4];
      ldsm_x4(aQ + swz(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)), af);
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {          // two key tiles per ldmatrix.x4
        uint32_t bf[4];
        ldsm_x4(aK + swz(np * 16 + (lane & 7) + (lane >> 4) * 8, ks * 2 + ((lane >> 3) & 1)), bf);
        MmaK16<T>::mma(sacc[2 * np], af, bf[0], bf[1]);
        MmaK16<T>::mma(sacc[2 * np + 1], af, bf[2], bf[3]);
      }
    }
    // ---- softmax over keys (rows qr and qr + 8 of this 16-query tile)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = n * 8 + qc + e < Tn;
        sacc[n][e] = ok ? sacc[n][e] * scale_log2e : -INFINITY;
        sacc[n][2 + e] = ok ? sacc[n][2 + e] * scale_log2e : -INFINITY;
        m0 = fmaxf(m0, sacc[n][e]);
        m1 = fmaxf(m1, sacc[n][2 + e]);
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sacc[n][e] = exp2f(sacc[n][e] - m0); s0 += sacc[n][e];
        sacc[n][2 + e] = exp2f(sacc[n][2 + e] - m1); s1 += sacc[n][2 + e];
      }
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float i0 = 1.0f / s0, i1 = 1.0f / s1;
    // ---- O = P V ; probabilities are normalised, then rounded to the storage type (A operand)
    float oacc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < TPAD / 16; ++kk) {         // 16 keys per step
      uint32_t pf[4];
      pf[0] = Elem<T>::pack2(sacc[2 * kk][0] * i0, sacc[2 * kk][1] * i0);
      pf[1] = Elem<T>::pack2(sacc[2 * kk][2] * i1, sacc[2 * kk][3] * i1);
      pf[2] = Elem<T>::pack2(sacc[2 * kk + 1][0] * i0, sacc[2 * kk + 1][1] * i0);
      pf[3] = Elem<T>::pack2(sacc[2 * kk + 1][2] * i1, sacc[2 * kk + 1][3] * i1);
#pragma unroll
      for (int np = 0; np < 4; ++np) {               // two dim tiles (16 dims) per ldmatrix.x4.trans
        uint32_t vf[4];
        ldsm_x4_trans(aV + swz(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 2 + (lane >> 4)), vf);
        MmaK16<T>::mma(oacc[2 * np], pf, vf[0], vf[1]);
        MmaK16<T>::mma(oacc[2 * np + 1], pf, vf[2], vf[3]);
      }
    }
    // ---- park the 16 x 64 output tile over this tile's (consumed) Q rows
    __syncwarp();
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const uint32_t w0 = Elem<T>::pack2(oacc[n][0], oacc[n][1]), w1 = Elem<T>::pack2(oacc[n][2], oacc[n][3]);
      const uint32_t a0 = aQ + swz(mt * 16 + qr, n) + qc * 2, a1 = aQ + swz(mt * 16 + qr + 8, n) + qc * 2;
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0), "r"(w0) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1), "r"(w1) : "memory");
    }
  }
  __syncwarp();
  // ---- coalesced 16-byte stores of the T output rows
  uint16_t* ob = reinterpret_cast<uint16_t*>(out);
  for (int i = lane; i < Tn * 8; i += 32) {
    const int c = i & 7, row = i >> 3;
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(aQ + swz(row, c)));
    *reinterpret_cast<uint4*>(ob + ((long long)row * P + p) * C + h * 64 + c * 8) = v;
  }
}


// ------------------------------------------------------------------ cross-attention (short context)
// Attention of N queries per frame against a SHORT key/value set (Lk <= 128 tokens: the 77 CLIP text
// tokens of StableNormal's fixed prompt).  CTA = 64 queries of one (frame, head): K and V of the head are
// staged once in XOR-swizzled smem and shared by 4 warps of 16 queries each; S = Q K^T and O = P V run on
// mma.sync m16n8k16 with the softmax in accumulator registers (same scheme as temporal attention).
// HBM bound: Q read once, O written once, K/V come from L2.
//   q  [F*N][ldq]   (head h at columns h*64..)          kv [Fk*Lk][2C]  (K | V; Fk = 1 when shared)
template <typename T, int KPAD>
__global__ void __launch_bounds__(128)
cross_attn_kernel(const void* __restrict__ q, int ldq, const void* __restrict__ kv, void* __restrict__ out, int N,
                  int C, int Lk, long long kv_frame_rows, float scale_log2e) {
  UG_TRIGGER_ATT();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t sm_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = blockIdx.y, f = blockIdx.z;
  const int q0 = blockIdx.x * 64 + warp * 16;
  const uint32_t aK = smem_u32(sm_raw), aV = aK + KPAD * 128, aQ = aV + KPAD * 128 + (uint32_t)warp * 16 * 128;
  const uint16_t* kvb = reinterpret_cast<const uint16_t*>(kv) + (long long)f * kv_frame_rows * 2 * C;
  const uint16_t* qb = reinterpret_cast<const uint16_t*>(q);

  for (int i = threadIdx.x; i < 2 * KPAD * 8; i += 128) {        // K | V rows of this head
    const int c = i & 7, row = (i >> 3) % KPAD, which = i / (8 * KPAD);
    const uint32_t dst = aK + (uint32_t)which * KPAD * 128 + swz(row, c);
    if (row < Lk) {
      const uint16_t* src = kvb + (long long)row * 2 * C + which * C + h * 64 + c * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    } else {
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
    }
  }
  for (int i = lane; i < 16 * 8; i += 32) {                       // this warp's 16 query rows
    const int c = i & 7, row = i >> 3;
    const uint32_t dst = aQ + swz(row, c);
    if (q0 + row < N) {
      const uint16_t* src = qb + ((long long)f * N + q0 + row) * ldq + h * 64 + c * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    } else {
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (q0 >= N) return;

  constexpr int NT = KPAD / 8;
  const int qr = lane >> 2, qc = (lane & 3) * 2;
  float sacc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) { sacc[n][0] = sacc[n][1] = sacc[n][2] = sacc[n][3] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t af[4];
    ldsm_x4(aQ + swz((lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)), af);
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t bf[4];
      ldsm_x4(aK + swz(np * 16 + (lane & 7) + (lane >> 4) * 8, ks * 2 + ((lane >> 3) & 1)), bf);
      MmaK16<T>::mma(sacc[2 * np], af, bf[0], bf[1]);
      MmaK16<T>::mma(sacc[2 * np + 1], af, bf[2], bf[3]);
    }
  }
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool ok = n * 8 + qc + e < Lk;
      sacc[n][e] = ok ? sacc[n][e] * scale_log2e : -INFINITY;
      sacc[n][2 + e] = ok ? sacc[n][2 + e] * scale_log2e : -INFINITY;
      m0 = fmaxf(m0, sacc[n][e]);
      m1 = fmaxf(m1, sacc[n][2 + e]);
    }
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      sacc[n][e] = exp2f(sacc[n][e] - m0); s0 += sacc[n][e];
      sacc[n][2 + e] = exp2f(sacc[n][2 + e] - m1); s1 += sacc[n][2 + e];
    }
  }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float i0 = 1.0f / s0, i1 = 1.0f / s1;
  float oacc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) { oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < KPAD / 16; ++kk) {
    uint32_t pf[4];
    pf[0] = Elem<T>::pack2(sacc[2 * kk][0] * i0, sacc[2 * kk][1] * i0);
    pf[1] = Elem<T>::pack2(sacc[2 * kk][2] * i1, sacc[2 * kk][3] * i1);
    pf[2] = Elem<T>::pack2(sacc[2 * kk + 1][0] * i0, sacc[2 * kk + 1][1] * i0);
    pf[3] = Elem<T>::pack2(sacc[2 * kk + 1][2] * i1, sacc[2 * kk + 1][3] * i1);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t vf[4];
      ldsm_x4_trans(aV + swz(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 2 + (lane >> 4)), vf);
      MmaK16<T>::mma(oacc[2 * np], pf, vf[0], vf[1]);
      MmaK16<T>::mma(oacc[2 * np + 1], pf, vf[2], vf[3]);
    }
  }
  __syncwarp();
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const uint32_t w0 = Elem<T>::pack2(oacc[n][0], oacc[n][1]), w1 = Elem<T>::pack2(oacc[n][2], oacc[n][3]);
    const uint32_t a0 = aQ + swz(qr, n) + qc * 2, a1 = aQ + swz(qr + 8, n) + qc * 2;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0), "r"(w0) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1), "r"(w1) : "memory");
  }
  __syncwarp();
  uint16_t* ob = reinterpret_cast<uint16_t*>(out);
  for (int i = lane; i < 16 * 8; i += 32) {
    const int c = i & 7, row = i >> 3;
    if (q0 + row >= N) continue;
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(aQ + swz(row, c)));
    *reinterpret_cast<uint4*>(ob + ((long long)f * N + q0 + row) * C + h * 64 + c * 8) = v;
  }
}

// ------------------------------------------------------------------ 2-D (StableNormal) scheduler glue
// DDIM step, prediction_type = "sample", eta = 0, in place on fp32 latents:
//   x <- c_x0 * x0 + c_x * x   with  c_x = sqrt((1-a_prev)/(1-a_t)),  c_x0 = sqrt(a_prev) - sqrt(a_t) * c_x
__global__ void axpby_kernel(float* x, const float* x0, float c_x0, float c_x, long long n) {
  UG_TRIGGER_MISC();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = c_x0 * ld_act(x0 + i) + c_x * x[i];
}
// fp32 [tokens][Cs] -> 16-bit [tokens][Cd] (Cd >= Cs, zero padded): the UNet input of the 2-D path
template <typename T>
__global__ void f32_to_tokens_kernel(const float* x, int Cs, int Cd, long long tokens, T* __restrict__ y) {
  UG_TRIGGER_MISC();
  pdl_wait();
  const long long n = tokens * Cd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / Cd;
    const int c = (int)(i % Cd);
    y[i] = Elem<T>::from_f(c < Cs ? ld_act(x + t * Cs + c) : 0.f);
  }
}
// decoded normals: 16-bit NHWC [N][HW][Cs] (3 valid) -> unit vectors, clip, 8-bit HWC [N][HW][3]
template <typename T>
__global__ void normals_to_u8_kernel(const T* x, int Cs, long long pixels, uint8_t* y) {   // x: no read-only path (PDL)
  UG_TRIGGER_MISC();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
    const float a = Elem<T>::to_f(x[i * Cs]), b = Elem<T>::to_f(x[i * Cs + 1]), c = Elem<T>::to_f(x[i * Cs + 2]);
    const float inv = 1.0f / fmaxf(sqrtf(a * a + b * b + c * c), 1e-6f);
    const float v[3] = {a * inv, b * inv, c * inv};
#pragma unroll
    for (int k = 0; k < 3; ++k) y[i * 3 + k] = (uint8_t)((fminf(fmaxf(v[k], -1.f), 1.f) + 1.f) * 0.5f * 255.f);
  }
}

// ------------------------------------------------------------------ layout kernels
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W,
                                  int nvec) {
  const long long total = (long long)N * 2 * H * 2 * W * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % nvec);
    long long r = i / nvec;
    const int ox = (int)(r % (2 * W)); r /= (2 * W);
    const int oy = (int)(r % (2 * H));
    const long long n = r / (2 * H);
    y[i] = ld_act(x + ((n * H + (oy >> 1)) * W + (ox >> 1)) * nvec + c8);
  }
}
__global__ void concat_kernel(const void* __restrict__ x1, int nv1, const void* __restrict__ x2, int nv2,
                              long long rows, uint4* __restrict__ y) {
  UG_TRIGGER_MISC();
  pdl_wait();
  const int nvec = nv1 + nv2;
  const long long total = rows * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = ld_act(cat_ptr(x1, nv1, x2, nv2, i / nvec, (int)(i % nvec)));
}
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ noise, float ns,
                                    float scale, float shift, int N, int Cs, long long HW, int Cd,
                                    T* __restrict__ y) {
  const long long total = (long long)N * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i % HW;
    for (int c = 0; c < Cd; ++c) {
      float v = 0.f;
      if (c < Cs) {
        const long long idx = (n * Cs + c) * HW + p;
        v = x[idx] * scale + shift;
        if (noise) v += ns * noise[idx];
      }
      y[i * Cd + c] = Elem<T>::from_f(v);
    }
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, int N, long long HW, int Cs, int Cd, float scale,
                                    float shift, int clamp01, float* __restrict__ y) {
  const long long total = (long long)N * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i % HW;
    for (int c = 0; c < Cd; ++c) {
      float v = Elem<T>::to_f(x[i * Cs + c]) * scale + shift;
      if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
      y[(n * Cd + c) * HW + p] = v;
    }
  }
}

__global__ void f32_swap_kernel(const float* __restrict__ x, int N, long long HW, int C, float scale,
                                float* __restrict__ y, int to_nhwc) {
  const long long total = (long long)N * HW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // i indexes the NHWC tensor
    const int c = (int)(i % C);
    const long long p = (i / C) % HW;
    const long long n = i / (C * HW);
    const long long j = (n * C + c) * HW + p;
    if (to_nhwc) y[i] = x[j] * scale;
    else y[j] = x[i] * scale;
  }
}

// ------------------------------------------------------------------ small dense helpers
template <typename T>
__global__ void gemv_kernel(const void* __restrict__ W, const float* __restrict__ b,
                            const float* __restrict__ addend, const float* __restrict__ x,
                            float* __restrict__ out, int N, int K, int silu_in, int silu_out) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int m = blockIdx.y;
  if (n >= N) return;
  const uint4* w = reinterpret_cast<const uint4*>(W) + (long long)n * (K >> 3);
  const float* xr = x + (long long)m * K;
  float acc = 0.f;
  for (int i = lane; i < (K >> 3); i += 32) {
    float f[8];
    unpack8<T>(__ldg(w + i), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float xv = xr[i * 8 + j];
      if (silu_in) xv = silu_f(xv);
      acc = fmaf(f[j], xv, acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc + (b ? b[n] : 0.f) + (addend ? addend[n] : 0.f);
    out[(long long)m * N + n] = silu_out ? silu_f(v) : v;
  }
}
__global__ void sinusoid_vals_kernel(float4 vals, int n, int dim, float* __restrict__ out) {
  const int half = dim >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int r = i / half, j = i % half;
  const float v = r == 0 ? vals.x : r == 1 ? vals.y : r == 2 ? vals.z : vals.w;
  const float freq = expf(-9.210340371976184f * (float)j / (float)half);
  const float ang = v * freq;
  out[(long long)r * dim + j] = cosf(ang);
  out[(long long)r * dim + half + j] = sinf(ang);
}
__global__ void iota_kernel(float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)i;
}
__global__ void sinusoid_kernel(const float* __restrict__ vals, int n, int dim, float* __restrict__ out) {
  const int half = dim >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int r = i / half, j = i % half;
  const float freq = expf(-9.210340371976184f * (float)j / (float)half);   // ln(10000)
  const float ang = vals[r] * freq;
  out[(long long)r * dim + j] = cosf(ang);
  out[(long long)r * dim + half + j] = sinf(ang);
}
template <typename T>
__global__ void build_unet_input_kernel(const float4* __restrict__ lat, const uint2* __restrict__ cond,
                                        float inv, long long tokens, uint4* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tokens) return;
  const float4 l = lat[i];
  const uint2 c = cond[i];
  uint4 u;
  u.x = Elem<T>::pack2(l.x * inv, l.y * inv);
  u.y = Elem<T>::pack2(l.z * inv, l.w * inv);
  u.z = c.x;
  u.w = c.y;
  y[i] = u;
}
__global__ void euler_step_kernel(float* __restrict__ lat, const float* __restrict__ v, float c_v, float c_x,
                                  float inv_sigma, float dsigma, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = lat[i];
  const float x0 = v[i] * c_v + x * c_x;
  const float d = (x - x0) * inv_sigma;
  lat[i] = x + d * dsigma;
}
template <typename T>
__global__ void convert_weight_kernel(const void* __restrict__ src, int sd, T* __restrict__ dst, int Cout,
                                      int Cin, int CinPad, int taps) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    long long r = i / Cin;
    const int co = (int)(r % Cout);
    const int tap = (int)(r / Cout);
    const long long s = ((long long)co * Cin + ci) * taps + tap;
    float v = sd == 2 ? reinterpret_cast<const float*>(src)[s]
            : sd == 1 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[s])
                      : __half2float(reinterpret_cast<const __half*>(src)[s]);
    dst[((long long)tap * Cout + co) * CinPad + ci] = Elem<T>::from_f(v);
  }
}
__global__ void convert_f32_kernel(const void* __restrict__ src, int sd, float* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = sd == 2 ? reinterpret_cast<const float*>(src)[i]
           : sd == 1 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i])
                     : __half2float(reinterpret_cast<const __half*>(src)[i]);
}

// Split-K reduce (tapgemm ksplit > 1): out[m][n] = epilogue(sum_s part[s][m][n]) with the partials added in index order
// (deterministic) and the epilogue of the fused path in the same order: * scale, + bias, + frame bias, GELU, + residual,
// blend.  4 columns per thread (N % 4 == 0).
template <typename T>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* part, int S, long long M, int N, const float* __restrict__ bias,
                     const float* __restrict__ fbias, int fbias_ld, int fbias_div, const T* res, long long ldr,
                     const T* blend, long long ldb, float alpha, float scale, int act, void* out, long long ldc,
                     int out_fp32) {
  const int nq = N >> 2;
  const long long total = M * nq, slab = M * (long long)N;
  UG_TRIGGER_MISC();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / nq;
    const int n = (int)(i - m * nq) << 2;
    const float* p = part + m * N + n;
    float4 acc = ld_act(reinterpret_cast<const float4*>(p));
    for (int s_ = 1; s_ < S; ++s_) {
      const float4 v = ld_act(reinterpret_cast<const float4*>(p + (long long)s_ * slab));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float f[4] = {acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale};
    if (bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] += bias[n + j];
    }
    if (fbias != nullptr) {
      const float* fb = fbias + (m / fbias_div) * fbias_ld + n;
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] += fb[j];
    }
    if (act == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] = 0.5f * f[j] * (1.0f + erff(f[j] * 0.70710678118654752f));
    }
    if (res != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] += Elem<T>::to_f(res[m * ldr + n + j]);
    }
    if (blend != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] = alpha * Elem<T>::to_f(blend[m * ldb + n + j]) + (1.0f - alpha) * f[j];
    }
    if (out_fp32) {
      float* o = reinterpret_cast<float*>(out) + m * ldc + n;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = f[j];
    } else {
      T* o = reinterpret_cast<T*>(out) + m * ldc + n;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = Elem<T>::from_f(f[j]);
    }
  }
}

// Whole state dicts in ONE launch (ug_ctx_load_weights): block b serves tensor t with first_block[t] <= b <
// first_block[t + 1]; matrices go [cout][cin][taps] -> [tap][cout][cin_pad] 16-bit, vectors -> fp32.
__global__ void __launch_bounds__(256) convert_batch_kernel(const ConvertDesc* __restrict__ d, int n) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {                                   // last tensor whose first block is <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (d[mid].first_block <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const ConvertDesc t = d[lo];
  const long long nb = (lo + 1 < n ? d[lo + 1].first_block : (long long)gridDim.x) - t.first_block;
  const long long start = ((long long)blockIdx.x - t.first_block) * 256 + threadIdx.x, step = nb * 256;
  auto load = [&](long long s_) -> float {
    return t.src_dtype == 2 ? reinterpret_cast<const float*>(t.src)[s_]
         : t.src_dtype == 1 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(t.src)[s_])
                            : __half2float(reinterpret_cast<const __half*>(t.src)[s_]);
  };
  if (t.dst_fmt == 2) {
    for (long long i = start; i < t.total; i += step) reinterpret_cast<float*>(t.dst)[i] = load(i);
    return;
  }
  for (long long i = start; i < t.total; i += step) {
    const int ci = (int)(i % t.cin);
    const long long r = i / t.cin;
    const int co = (int)(r % t.cout);
    const int tap = (int)(r / t.cout);
    const float v = load(((long long)co * t.cin + ci) * t.taps + tap);
    const long long o = ((long long)tap * t.cout + co) * t.cin_pad + ci;
    if (t.dst_fmt == 1) reinterpret_cast<__nv_bfloat16*>(t.dst)[o] = __float2bfloat16_rn(v);
    else reinterpret_cast<__half*>(t.dst)[o] = __float2half_rn(v);
  }
}

inline int grid_for(long long total, int block, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}
inline int last_err() { return (int)cudaGetLastError(); }

struct GnGeom { int threads; int chunks; long long chunk_rows; };
inline GnGeom gn_geom(int nvec, long long rows_per_set, long long sets) {
  GnGeom g;
  int rpb = 256 / nvec;
  if (rpb < 1) rpb = 1;
  g.threads = nvec * rpb;   // >= 64 whenever C >= 64 * 8 / rpb ... padded below for tiny C
  if (g.threads < 64) g.threads = 64;
  long long want = (148LL * 4 + sets - 1) / sets;
  long long maxc = (rows_per_set + rpb * 8 - 1) / (rpb * 8);   // one batch of 8 loads per thread at least: fewer,
                                                               // fatter chunks shorten the partial fold
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  g.chunk_rows = (rows_per_set + want - 1) / want;
  g.chunks = (int)((rows_per_set + g.chunk_rows - 1) / g.chunk_rows);
  return g;
}

}  // namespace

long long gn_partial_floats(int C, long long rows, long long rows_per_set, int G) {
  const long long sets = rows / rows_per_set;
  GnGeom g = gn_geom(C / 8, rows_per_set, sets);
  return sets * g.chunks * G * 2 + sets * G * 2;     // per-chunk partials, then (mean, rstd) per set
}

int launch_gn_stats(const void* x1, int C1, const void* x2, int C2, long long rows, long long rows_per_set,
                    int G, float eps, float* stats, unsigned int* counters, int fmt, cudaStream_t st) {
  const int C = C1 + C2;
  if ((C1 & 7) || (C2 & 7) || C % G || G > 64 || C / 8 > 512) return (int)cudaErrorInvalidValue;
  const long long sets = rows / rows_per_set;
  GnGeom g = gn_geom(C / 8, rows_per_set, sets);
  dim3 grid(g.chunks, (unsigned)sets);
  const int rpb = g.threads / (C / 8);
  size_t smem = (size_t)2 * rpb * C * sizeof(float);
  const size_t smem_fin = (size_t)(g.threads / G) * G * 2 * sizeof(float);
  if (smem < smem_fin) smem = smem_fin;
  if (sets > kGnMaxSets || g.threads < 2 * G) return (int)cudaErrorInvalidValue;
  float* mr = stats + sets * g.chunks * G * 2;
  const float inv_cnt = 1.0f / ((float)rows_per_set * (float)(C / G));
  cudaError_t err;
  UG_DISPATCH_FMT(fmt, (err = launch_pdl_tag("gn_stats", gn_stats_kernel<T>, grid, dim3(g.threads), smem, st, x1, C1 / 8, x2, C2 / 8,
                                         rows_per_set, g.chunk_rows, G, C / G, stats, mr, counters, inv_cnt, eps)));
  return (int)err;
}

// one-launch GroupNorm; returns cudaErrorNotSupported when the grid cannot be made co-resident (caller falls back
// to stats + apply).  counters: 3 * kGnMaxSets zero-initialised uints.
int launch_gn_fused(const void* x1, int C1, const void* x2, int C2, long long rows, long long rows_per_set, int G,
                    float eps, float* stats, unsigned int* counters, const float* gamma, const float* beta, int silu,
                    void* y, int fmt, cudaStream_t st) {
  const int C = C1 + C2;
  if ((C1 & 7) || (C2 & 7) || C % G || G > 64 || C / 8 > 320) return (int)cudaErrorInvalidValue;
  const long long sets = rows / rows_per_set;
  GnGeom g = gn_geom(C / 8, rows_per_set, sets);
  const int rpb = g.threads / (C / 8);
  size_t smem = (size_t)2 * rpb * C * sizeof(float);
  const size_t smem_fin = (size_t)(g.threads / G) * G * 2 * sizeof(float);
  const size_t smem_apply = (size_t)(C * 2 + G * 2) * sizeof(float);
  if (smem < smem_fin) smem = smem_fin;
  if (smem < smem_apply) smem = smem_apply;
  if (sets > kGnMaxSets || g.threads < 2 * G || g.threads > 320) return (int)cudaErrorInvalidValue;
  // co-residency: CTAs per SM for this block shape x SM count, cached per DEVICE under a mutex (contexts of several
  // devices / threads share these statics)
  static std::mutex mu;
  static std::map<int, int> sms_of;
  static std::map<unsigned long long, int> occ_cache;
  int dev = 0;
  cudaGetDevice(&dev);
  int sms = 0, occ = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    auto sit = sms_of.find(dev);
    if (sit == sms_of.end()) {
      int n = 0;
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
      sit = sms_of.emplace(dev, n).first;
    }
    sms = sit->second;
    const unsigned long long key = ((unsigned long long)dev << 48) | ((unsigned long long)fmt << 40) |
                                   ((unsigned long long)g.threads << 24) | (unsigned long long)smem;
    auto it = occ_cache.find(key);
    if (it == occ_cache.end()) {
      int nb = 0;
      cudaError_t e;
      if (fmt == 1) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gn_fused_kernel<__nv_bfloat16>, g.threads, smem);
      else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gn_fused_kernel<__half>, g.threads, smem);
      if (e != cudaSuccess) return (int)e;
      it = occ_cache.emplace(key, nb).first;
    }
    occ = it->second;
  }
  const long long resident = (long long)occ * sms;
  if (resident < sets) return (int)cudaErrorNotSupported;
  long long chunks = resident / sets;
  if (chunks > g.chunks) chunks = g.chunks;
  const long long chunk_rows = (rows_per_set + chunks - 1) / chunks;
  chunks = (rows_per_set + chunk_rows - 1) / chunk_rows;
  dim3 grid((unsigned)chunks, (unsigned)sets);
  float* mr = stats + sets * chunks * G * 2;
  const float inv_cnt = 1.0f / ((float)rows_per_set * (float)(C / G));
  // The kernel's CTAs wait for each other (grid-wide flag), so they must all be resident.  A COOPERATIVE launch makes the
  // driver guarantee that (or fail the launch) even when other streams / processes share the GPU; a plain launch sized
  // from the occupancy query only assumes it.  UG_GN_COOP=0 selects the plain launch; a cooperative launch that the
  // driver refuses (e.g. in combination with programmatic dependent launch) falls back to it once and for all.
  static std::atomic<int> coop{[] { const char* e = getenv("UG_GN_COOP"); return e ? atoi(e) : 1; }()};
  cudaError_t err = cudaSuccess;
  if (coop.load() != 0) {
    static const bool no_pdl = pdl_off("gn_fused");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(g.threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (no_pdl || coop.load() == 2) ? 1 : 2;       // 2: cooperative without the PDL attribute
    UG_DISPATCH_FMT(fmt, (err = cudaLaunchKernelEx(&cfg, gn_fused_kernel<T>, x1, C1 / 8, x2, C2 / 8, rows_per_set,
                                                   chunk_rows, G, C / G, stats, mr, counters, inv_cnt, eps, gamma, beta,
                                                   silu, y)));
    if (err == cudaSuccess) return 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap != cudaStreamCaptureStatusNone) return (int)err;   // never retry inside a capture: surface the error
    cudaGetLastError();
    coop.store(0);
  }
  UG_DISPATCH_FMT(fmt, (err = launch_pdl_tag("gn_fused", gn_fused_kernel<T>, grid, dim3(g.threads), smem, st, x1, C1 / 8, x2, C2 / 8,
                                         rows_per_set, chunk_rows, G, C / G, stats, mr, counters, inv_cnt, eps, gamma,
                                         beta, silu, y)));
  return (int)err;
}

int launch_gn_finalize(int C, long long rows, long long rows_per_set, int G, float eps, float* stats,
                       cudaStream_t st) {
  const long long sets = rows / rows_per_set;
  GnGeom g = gn_geom(C / 8, rows_per_set, sets);
  float* mr = stats + sets * g.chunks * G * 2;
  const float inv_cnt = 1.0f / ((float)rows_per_set * (float)(C / G));
  gn_finalize_kernel<<<(unsigned)sets, 8 * G, 0, st>>>(stats, g.chunks, G, inv_cnt, eps, mr);
  return last_err();
}

int launch_gn_apply(const void* x1, int C1, const void* x2, int C2, long long rows, long long rows_per_set,
                    int G, const float* stats, const float* gamma, const float* beta, float eps, int silu,
                    void* y, int fmt, cudaStream_t st) {
  const int C = C1 + C2;
  if ((C1 & 7) || (C2 & 7) || C % G) return (int)cudaErrorInvalidValue;
  const long long sets = rows / rows_per_set;
  GnGeom g = gn_geom(C / 8, rows_per_set, sets);
  dim3 grid(g.chunks, (unsigned)sets);
  const size_t smem = (size_t)(C * 2 + G * 2) * sizeof(float);
  const float* mr = stats + sets * g.chunks * G * 2;
  (void)eps;
  cudaError_t err;
  UG_DISPATCH_FMT(fmt, (err = launch_pdl_tag("gn_apply", gn_apply_kernel<T>, grid, dim3(g.threads), smem, st, x1, C1 / 8, x2, C2 / 8,
                                         rows_per_set, g.chunk_rows, G, C / G, mr, gamma, beta, silu, y)));
  return (int)err;
}

// GroupNorm over clusters of <= 8 CTAs (gn_cluster_kernel); cudaErrorNotSupported when a statistics set does not fit
// the shared memory of one cluster (caller uses launch_gn_fused).  UG_GN_CLUSTER=0 disables.
int launch_gn_cluster(const void* x1, int C1, const void* x2, int C2, long long rows, long long rows_per_set, int G,
                      float eps, const float* gamma, const float* beta, int silu, void* y, int fmt, cudaStream_t st) {
  static const bool off = [] { const char* e = getenv("UG_GN_CLUSTER"); return e && atoi(e) == 0; }();
  const int C = C1 + C2;
  const int nvec = C / 8;
  if (off || (C1 & 7) || (C2 & 7) || C % G || G > 64 || nvec > 320 || nvec < 1 || rows % rows_per_set)
    return (int)cudaErrorNotSupported;
  const long long sets = rows / rows_per_set;
  const int rpb = 320 / nvec;
  const int threads = nvec * rpb;
  if (threads < G || sets > 65535 || rows_per_set > (1 << 24)) return (int)cudaErrorNotSupported;
  // measured (tools/ab_gn.py): the cluster form wins while a CTA's rows stay <= 64 KB (17 -> 13, 23 -> 17, 25 -> 20 us
  // at the L3 / L2 shapes) and loses at 123+ KB per CTA (one 320-thread CTA per SM cannot keep enough loads in flight)
  constexpr size_t kBudget = 96 * 1024, kTileMax = 64 * 1024;
  const size_t fixed = ((size_t)2 * rpb * C + 4 * G) * sizeof(float);
  int S = 8;
  while (S > 1 && S > rows_per_set) S >>= 1;
  const int chunk_rows = (int)((rows_per_set + S - 1) / S);
  const size_t tile_bytes = (size_t)chunk_rows * nvec * 16;
  const size_t smem = tile_bytes + fixed;
  if (tile_bytes > kTileMax || smem > kBudget) return (int)cudaErrorNotSupported;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gn_cluster_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kBudget);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gn_cluster_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)kBudget);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)S, (unsigned)sets);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  static const bool no_pdl = pdl_off("gn_cluster");
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 1 : 2;
  const float inv_cnt = 1.0f / ((float)rows_per_set * (float)(C / G));
  cudaError_t err;
  UG_DISPATCH_FMT(fmt, (err = cudaLaunchKernelEx(&cfg, gn_cluster_kernel<T>, x1, C1 / 8, x2, C2 / 8, rows_per_set,
                                                 chunk_rows, G, C / G, inv_cnt, eps, gamma, beta, silu, y)));
  return (int)err;
}

template <int LPR, int VPL>
int launch_ln_t(const void* x, long long rows, int C, const float* gamma, const float* beta, float eps,
                const float* add, int add_div, void* y, int fmt, cudaStream_t st) {
  const int wpb = 8;
  const long long rows_per_block = (long long)wpb * (32 / LPR);
  const unsigned grid = (unsigned)((rows + rows_per_block - 1) / rows_per_block);
  cudaError_t err;
  UG_DISPATCH_FMT(fmt, (err = launch_pdl_tag("layernorm", layernorm_kernel<T, LPR, VPL>, dim3(grid), dim3(wpb * 32), 0, st, x, rows, C,
                                         gamma, beta, eps, add, add_div > 0 ? add_div : 1, y)));
  return (int)err;
}

template <typename T>
__global__ void ln_fold_weights_kernel(const T* __restrict__ W, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ bias,
                                       T* __restrict__ Wf, float* __restrict__ colsum, float* __restrict__ bias_out,
                                       int N, int K) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float cs = 0.f, bb = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = static_cast<float>(W[(size_t)n * K + k]);
    const T wf = static_cast<T>(gamma[k] * w);
    Wf[(size_t)n * K + k] = wf;
    cs += static_cast<float>(wf);            // of the ROUNDED weights: a constant row cancels exactly in the epilogue
    bb = fmaf(beta[k], w, bb);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cs += __shfl_xor_sync(0xffffffffu, cs, o);
    bb += __shfl_xor_sync(0xffffffffu, bb, o);
  }
  if (lane == 0) {
    colsum[n] = cs;
    bias_out[n] = bb + (bias != nullptr ? bias[n] : 0.f);
  }
}

__global__ void scale_f32_kernel(const float* __restrict__ x, float a, float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * x[i];
}

template <int LPR, int VPL>
int launch_rs_t(const void* x, long long rows, int C, float eps, float2* stat, int fmt, cudaStream_t st) {
  const int wpb = 8;
  const long long rows_per_block = (long long)wpb * (32 / LPR);
  const unsigned grid = (unsigned)((rows + rows_per_block - 1) / rows_per_block);
  cudaError_t err;
  UG_DISPATCH_FMT(fmt, (err = launch_pdl_tag("layernorm", row_stats_kernel<T, LPR, VPL>, dim3(grid), dim3(wpb * 32), 0, st, x,
                                         rows, C, eps, stat)));
  return (int)err;
}

int launch_ln_fold_weights(const void* W, const float* gamma, const float* beta, const float* bias, void* Wf,
                           float* colsum, float* bias_out, int N, int K, int fmt, cudaStream_t st) {
  if (N <= 0 || K <= 0) return (int)cudaErrorInvalidValue;
  const unsigned grid = (unsigned)((N + 7) / 8);
  UG_DISPATCH_FMT(fmt, (ln_fold_weights_kernel<T><<<grid, 256, 0, st>>>(reinterpret_cast<const T*>(W), gamma, beta, bias,
                                                                      reinterpret_cast<T*>(Wf), colsum, bias_out, N, K)));
  return (int)cudaGetLastError();
}

int launch_scale_f32(const float* x, float a, float* y, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  scale_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, a, y, n);
  return (int)cudaGetLastError();
}

int launch_row_stats(const void* x, long long rows, int C, float eps, float2* stat, int fmt, cudaStream_t st) {
  if ((C & 7) || C > 2048 || C < 8) return (int)cudaErrorInvalidValue;
  const int nvec = C / 8;
#define UG_RS(LPR, VPL) return launch_rs_t<LPR, VPL>(x, rows, C, eps, stat, fmt, st)
  if (nvec <= 8) UG_RS(8, 1);
  if (nvec <= 16) UG_RS(16, 1);
  if (nvec <= 32) UG_RS(32, 1);
  if (nvec <= 40) UG_RS(8, 5);
  if (nvec <= 64) UG_RS(32, 2);
  if (nvec <= 80) UG_RS(16, 5);
  if (nvec <= 128) UG_RS(32, 4);
  if (nvec <= 160) UG_RS(32, 5);
  UG_RS(32, 8);
#undef UG_RS
}

int launch_layernorm(const void* x, long long rows, int C, const float* gamma, const float* beta, float eps,
                     const float* add, int add_div, void* y, int fmt, cudaStream_t st) {
  if ((C & 7) || C > 2048 || C < 8) return (int)cudaErrorInvalidValue;
  const int nvec = C / 8;
#define UG_LN(LPR, VPL) return launch_ln_t<LPR, VPL>(x, rows, C, gamma, beta, eps, add, add_div, y, fmt, st)
  if (nvec <= 8) UG_LN(8, 1);          // C <= 64
  if (nvec <= 16) UG_LN(16, 1);        // C <= 128
  if (nvec <= 32) UG_LN(32, 1);        // C <= 256
  if (nvec <= 40) UG_LN(8, 5);         // C = 320
  if (nvec <= 64) UG_LN(32, 2);        // C <= 512
  if (nvec <= 80) UG_LN(16, 5);        // C = 640
  if (nvec <= 128) UG_LN(32, 4);       // C <= 1024
  if (nvec <= 160) UG_LN(32, 5);       // C = 1280
  UG_LN(32, 8);                        // C <= 2048
#undef UG_LN
}

int launch_softmax_rows(void* s, long long rows, int n, int n_valid, float scale, int fmt, cudaStream_t st) {
  if ((n & 7) || n_valid < 1 || n_valid > n) return (int)cudaErrorInvalidValue;
  const float sl = scale * 1.4426950408889634f;
  const int nvec = n / 8;
  if (nvec <= 32) {
    UG_DISPATCH_FMT(fmt, (softmax_rows_kernel<T, 1><<<(unsigned)rows, 32, 0, st>>>(s, rows, n, n_valid, sl)));
  } else if (nvec <= 128) {
    UG_DISPATCH_FMT(fmt, (softmax_rows_kernel<T, 1><<<(unsigned)rows, 128, 0, st>>>(s, rows, n, n_valid, sl)));
  } else if (nvec <= 512) {
    UG_DISPATCH_FMT(fmt, (softmax_rows_kernel<T, 2><<<(unsigned)rows, 256, 0, st>>>(s, rows, n, n_valid, sl)));
  } else if (nvec <= 2048) {
    UG_DISPATCH_FMT(fmt, (softmax_rows_kernel<T, 8><<<(unsigned)rows, 256, 0, st>>>(s, rows, n, n_valid, sl)));
  } else {
    return (int)cudaErrorInvalidValue;
  }
  return last_err();
}

int launch_temporal_attention(const void* qkv, void* out, int Tn, long long P, int C, float scale, int fmt,
                              cudaStream_t st) {
  if ((C & 63) || Tn > 64 || Tn < 1) return (int)cudaErrorInvalidValue;
  const int wpb = 4;
  const long long items = P * (C / 64);
  const unsigned grid = (unsigned)((items + wpb - 1) / wpb);
  const float sl = scale * 1.4426950408889634f;
  if (Tn <= 32) {
    const size_t smem = (size_t)wpb * 3 * 32 * 128;
    cudaError_t err;
    UG_DISPATCH_FMT(fmt, (err = launch_pdl_tag("tattn", temporal_attn_kernel<T, 32>, dim3(grid), dim3(wpb * 32), smem, st, qkv, out,
                                           Tn, P, C, sl)));
    return (int)err;
  } else {
    const size_t smem = (size_t)wpb * 3 * 64 * 128;
    static bool configured = false;      // once, outside any stream capture (the first call of a loop is eager)
    if (!configured) {
      cudaFuncSetAttribute(temporal_attn_kernel<__half, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(temporal_attn_kernel<__nv_bfloat16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      configured = true;
    }
    cudaError_t err;
    UG_DISPATCH_FMT(fmt, (err = launch_pdl_tag("tattn", temporal_attn_kernel<T, 64>, dim3(grid), dim3(wpb * 32), smem, st, qkv, out,
                                           Tn, P, C, sl)));
    return (int)err;
  }
}

int launch_cross_attention(const void* q, int ldq, const void* kv, void* out, int F, int N, int C, int Lk,
                           int kv_per_frame, float scale, int fmt, cudaStream_t st) {
  if ((C & 63) || (ldq & 7) || Lk < 1 || Lk > 128 || N < 1) return (int)cudaErrorInvalidValue;
  const int kpad = (Lk + 15) & ~15;
  const dim3 grid((N + 63) / 64, C / 64, F);
  const size_t smem = (size_t)(2 * kpad + 64) * 128;
  const float sl = scale * 1.4426950408889634f;
  const long long fr = kv_per_frame ? Lk : 0;
  cudaError_t err = cudaErrorInvalidValue;
#define UG_XATTN(KP)                                                                                             \
  case KP:                                                                                                       \
    UG_DISPATCH_FMT(fmt, (err = launch_pdl_tag("cattn", cross_attn_kernel<T, KP>, grid, dim3(128), smem, st, q, ldq, kv, out, N, C, \
                                           Lk, fr, sl)));                                                        \
    break;
  switch (kpad) {
    UG_XATTN(16) UG_XATTN(32) UG_XATTN(48) UG_XATTN(64) UG_XATTN(80) UG_XATTN(96) UG_XATTN(112) UG_XATTN(128)
    default: break;
  }
#undef UG_XATTN
  return (int)err;
}

int launch_axpby(float* x, const float* x0, float c_x0, float c_x, long long n, cudaStream_t st) {
  return (int)launch_pdl(axpby_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, x, x0, c_x0, c_x, n);
}
int launch_f32_to_tokens(const float* x, int Cs, int Cd, long long tokens, void* y, int fmt, cudaStream_t st) {
  cudaError_t err;
  UG_DISPATCH_FMT(fmt, (err = launch_pdl(f32_to_tokens_kernel<T>, dim3(grid_for(tokens * Cd, 256)), dim3(256), 0, st, x,
                                         Cs, Cd, tokens, reinterpret_cast<T*>(y))));
  return (int)err;
}
int launch_normals_to_u8(const void* x, int Cs, long long pixels, unsigned char* y, int fmt, cudaStream_t st) {
  cudaError_t err;
  UG_DISPATCH_FMT(fmt, (err = launch_pdl(normals_to_u8_kernel<T>, dim3(grid_for(pixels, 256)), dim3(256), 0, st,
                                         reinterpret_cast<const T*>(x), Cs, pixels, y)));
  return (int)err;
}

int launch_upsample2x(const void* x, void* y, int N, int H, int W, int C, cudaStream_t st) {
  if (C & 7) return (int)cudaErrorInvalidValue;
  const long long total = (long long)N * 4 * H * W * (C / 8);
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(x),
                                                           reinterpret_cast<uint4*>(y), N, H, W, C / 8);
  return last_err();
}

int launch_concat(const void* x1, int C1, const void* x2, int C2, long long rows, void* y, cudaStream_t st) {
  if ((C1 & 7) || (C2 & 7)) return (int)cudaErrorInvalidValue;
  const long long total = rows * ((C1 + C2) / 8);
  return (int)launch_pdl_tag("concat", concat_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, x1, C1 / 8, x2, C2 / 8, rows,
                         reinterpret_cast<uint4*>(y));
}

int launch_nchw_to_nhwc(const float* x, const float* noise, float noise_scale, float scale, float shift, int N,
                        int Csrc, int H, int W, int Cdst, void* y, int fmt, cudaStream_t st) {
  const long long total = (long long)N * H * W;
  UG_DISPATCH_FMT(fmt, (nchw_to_nhwc_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(
                           x, noise, noise_scale, scale, shift, N, Csrc, (long long)H * W, Cdst,
                           reinterpret_cast<T*>(y))));
  return last_err();
}

int launch_nhwc_to_nchw(const void* x, int N, int H, int W, int Csrc, int Cdst, float scale, float shift,
                        int clamp01, float* y, int fmt, cudaStream_t st) {
  const long long total = (long long)N * H * W;
  UG_DISPATCH_FMT(fmt, (nhwc_to_nchw_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(
                           reinterpret_cast<const T*>(x), N, (long long)H * W, Csrc, Cdst, scale, shift, clamp01,
                           y)));
  return last_err();
}

int launch_f32_nchw_to_nhwc(const float* x, int N, long long HW, int C, float scale, float* y, cudaStream_t st) {
  f32_swap_kernel<<<grid_for((long long)N * HW * C, 256), 256, 0, st>>>(x, N, HW, C, scale, y, 1);
  return last_err();
}
int launch_f32_nhwc_to_nchw(const float* x, int N, long long HW, int C, float scale, float* y, cudaStream_t st) {
  f32_swap_kernel<<<grid_for((long long)N * HW * C, 256), 256, 0, st>>>(x, N, HW, C, scale, y, 0);
  return last_err();
}

int launch_gemv(const void* W, const float* b, const float* addend, const float* x, float* out, int M, int N,
                int K, int silu_in, int silu_out, int fmt, cudaStream_t st) {
  if (K & 7) return (int)cudaErrorInvalidValue;
  const int wpb = 4;
  dim3 grid((N + wpb - 1) / wpb, M);
  UG_DISPATCH_FMT(fmt, (gemv_kernel<T><<<grid, wpb * 32, 0, st>>>(W, b, addend, x, out, N, K, silu_in, silu_out)));
  return last_err();
}

int launch_sinusoid(const float* vals, int n, int dim, float* out, cudaStream_t st) {
  const int total = n * (dim / 2);
  sinusoid_kernel<<<(total + 127) / 128, 128, 0, st>>>(vals, n, dim, out);
  return last_err();
}

int launch_sinusoid_vals(float v0, float v1, float v2, float v3, int n, int dim, float* out, cudaStream_t st) {
  const int total = n * (dim / 2);
  sinusoid_vals_kernel<<<(total + 127) / 128, 128, 0, st>>>(make_float4(v0, v1, v2, v3), n, dim, out);
  return last_err();
}
int launch_iota(float* out, int n, cudaStream_t st) {
  iota_kernel<<<(n + 127) / 128, 128, 0, st>>>(out, n);
  return last_err();
}

int launch_build_unet_input(const float* latents, const void* cond, float sigma, long long tokens, void* y,
                            int fmt, cudaStream_t st) {
  const float inv = 1.0f / sqrtf(sigma * sigma + 1.0f);
  UG_DISPATCH_FMT(fmt, (build_unet_input_kernel<T><<<(unsigned)((tokens + 255) / 256), 256, 0, st>>>(
                           reinterpret_cast<const float4*>(latents), reinterpret_cast<const uint2*>(cond), inv,
                           tokens, reinterpret_cast<uint4*>(y))));
  return last_err();
}

int launch_euler_step(float* latents, const float* v, float sigma, float sigma_next, long long n,
                      cudaStream_t st) {
  const float c_v = -sigma / sqrtf(sigma * sigma + 1.0f);
  const float c_x = 1.0f / (sigma * sigma + 1.0f);
  euler_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(latents, v, c_v, c_x, 1.0f / sigma,
                                                                  sigma_next - sigma, n);
  return last_err();
}

int launch_convert_weight(const void* src, int src_dtype, void* dst, int Cout, int Cin, int CinPad, int taps,
                          int fmt, cudaStream_t st) {
  const long long total = (long long)Cout * Cin * taps;
  UG_DISPATCH_FMT(fmt, (convert_weight_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(
                           src, src_dtype, reinterpret_cast<T*>(dst), Cout, Cin, CinPad, taps)));
  return last_err();
}

int launch_splitk_reduce(const float* part, int S, long long M, int N, const float* bias, const float* fbias, int fbias_ld,
                         int fbias_div, const void* res, long long ldr, const void* blend, long long ldb, float alpha,
                         float scale, int act, void* out, long long ldc, int out_fp32, int fmt, cudaStream_t st) {
  if ((N & 3) || S < 1) return (int)cudaErrorInvalidValue;
  const long long total = M * (N >> 2);
  cudaError_t err;
  UG_DISPATCH_FMT(fmt, (err = launch_pdl_tag("splitk", splitk_reduce_kernel<T>, dim3(grid_for(total, 256, 148 * 8)), dim3(256), 0, st, part, S, M,
                                         N, bias, fbias, fbias_ld, fbias_div > 0 ? fbias_div : 1,
                                         reinterpret_cast<const T*>(res), ldr, reinterpret_cast<const T*>(blend), ldb, alpha,
                                         scale, act, out, ldc, out_fp32)));
  return (int)err;
}

int launch_convert_batch(const ConvertDesc* dev_descs, int n, long long total_blocks, cudaStream_t st) {
  if (n <= 0) return 0;
  convert_batch_kernel<<<(unsigned)total_blocks, 256, 0, st>>>(dev_descs, n);
  return last_err();
}

int launch_convert_f32(const void* src, int src_dtype, float* dst, long long n, cudaStream_t st) {
  convert_f32_kernel<<<grid_for(n, 256), 256, 0, st>>>(src, src_dtype, dst, n);
  return last_err();
}

}  // namespace ug

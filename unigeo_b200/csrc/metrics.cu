// The two metric functions eval.py applies to the plugin's outputs, on the device:
//   depth_evaluation(pred, gt, custom_mask, align_with_lstsq)   /root/reference/metrics/eval_depth.py:6-246
//     + align_with_lstsq_torch                                   /root/reference/metrics/alignment.py:150-167
//   normal_evaluation / compute_normal_metrics                   /root/reference/metrics/eval_normal.py:4-72
// as called at /root/reference/eval.py:49 and :54 (1.4 s + 0.34 s per clip on the CPU there).
//
// Per-pixel arithmetic follows the reference's fp32 operation order; sums are accumulated in fp64 by a fixed-order
// two-level reduction (deterministic), the scale/shift fit is the closed-form solution of the 2x2 normal equations
// in fp64 (the reference's fp32 SVD lstsq agrees to ~1e-6 relative), the median is an exact radix select of the
// lower middle value (torch.median semantics, App. B.15).  The unmodified reference functions stay the yardstick:
// tests compare against their bit-exact restatement with |delta| <= 1e-5.
#include "kernels.cuh"

namespace ug {
namespace {

constexpr int kBlocks = 148 * 4, kThreads = 256;

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* __restrict__ part) {
  __shared__ double s_red[NV][kThreads / 32];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) s_red[k][warp] = v[k];
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += s_red[threadIdx.x][w];
    part[(size_t)blockIdx.x * NV + threadIdx.x] = t;
  }
}

// fold [blocks][NV] partials -> out[NV]: warp w owns column w, lane l adds blocks l, l + 32, ... in order, then a
// fixed xor tree over the lanes (deterministic; a one-thread loop over 592 blocks cost 37 us of serial L2 round trips)
template <int NV>
__global__ void __launch_bounds__(32 * NV) fold_kernel(const double* __restrict__ part, int blocks,
                                                        double* __restrict__ out) {
  const int col = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double t = 0.0;
  for (int b = lane; b < blocks; b += 32) t += part[(size_t)b * NV + col];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) out[col] = t;
}

// ---- depth, pass 1: normal equations of [p 1] [s t]^T ~ g over valid = (gt > 0) & (gt < max_depth)
__global__ void __launch_bounds__(kThreads)
depth_fit_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long long n, float max_depth,
                 double* __restrict__ part) {
  double v[5] = {0, 0, 0, 0, 0};      // n, sum p, sum g, sum pp, sum pg
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const float g = gt[i];
    if (g > 0.f && g < max_depth) {
      const double p = pred[i], gd = g;
      v[0] += 1.0; v[1] += p; v[2] += gd; v[3] += p * p; v[4] += p * gd;
    }
  }
  block_reduce_store<5>(v, part);
}
__global__ void depth_solve_kernel(const double* __restrict__ sums, float* __restrict__ st) {
  const double n = sums[0], sp = sums[1], sg = sums[2], spp = sums[3], spg = sums[4];
  const double det = n * spp - sp * sp;
  const double s = det != 0.0 ? (n * spg - sp * sg) / det : 0.0;
  const double t = n > 0.0 ? (sg - s * sp) / n : 0.0;
  st[0] = (float)s;                    // the reference keeps the solution in float32 (alignment.py:160-163)
  st[1] = (float)t;
}
// ---- depth, pass 2: error terms of p' = s p + t on valid & custom mask
__global__ void __launch_bounds__(kThreads)
depth_err_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const unsigned char* __restrict__ mask,
                 long long n, float max_depth, const float* __restrict__ st, double* __restrict__ part,
                 float* __restrict__ err_map, float* __restrict__ pred_aligned, float* __restrict__ gt_valid) {
  const float s = st[0], t = st[1];
  double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // n, abs_rel, sq_rel, sq, log_sq, d<1, d<1.25, d<1.25^2, d<1.25^3
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const float g = gt[i];
    const bool valid = g > 0.f && g < max_depth;
    float p = __fadd_rn(__fmul_rn(s, pred[i]), t);
    const float d = __fsub_rn(p, g);
    // the three full-size maps the reference returns next to the dict (eval_depth.py:166-213), over `valid` only
    if (err_map != nullptr) err_map[i] = valid ? __fdiv_rn(fabsf(d), g) : 0.f;
    if (pred_aligned != nullptr) pred_aligned[i] = p;
    if (gt_valid != nullptr) gt_valid[i] = valid ? g : 0.f;
    if (!valid) continue;
    if (mask != nullptr && mask[i] == 0) continue;
    v[0] += 1.0;
    v[1] += (double)__fdiv_rn(fabsf(d), g);
    v[2] += (double)__fdiv_rn(__fmul_rn(d, d), g);
    v[3] += (double)__fmul_rn(d, d);
    p = fmaxf(p, 1e-5f);
    const float l = __fsub_rn(logf(p), logf(g));
    v[4] += (double)__fmul_rn(l, l);
    const float ratio = fmaxf(__fdiv_rn(p, g), __fdiv_rn(g, p));
    v[5] += ratio < 1.0f ? 1.0 : 0.0;
    v[6] += ratio < 1.25f ? 1.0 : 0.0;
    v[7] += ratio < (float)(1.25 * 1.25) ? 1.0 : 0.0;
    v[8] += ratio < (float)(1.25 * 1.25 * 1.25) ? 1.0 : 0.0;
  }
  block_reduce_store<9>(v, part);
}

// ---- normals: angular error per pixel (fp32, reference order) -> err[] (masked-out pixels: +inf bits sentinel),
// sums n, e, e^2, and the five threshold counts
__global__ void __launch_bounds__(kThreads)
normal_err_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const unsigned char* __restrict__ mask,
                  long long n, unsigned int* __restrict__ err_bits, double* __restrict__ part,
                  float* __restrict__ err_deg) {
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const bool keep = mask == nullptr || mask[i] != 0;
    if (!keep && err_deg == nullptr) { err_bits[i] = 0xffffffffu; continue; }
    const float px = pred[3 * i], py = pred[3 * i + 1], pz = pred[3 * i + 2];
    const float gx = gt[3 * i], gy = gt[3 * i + 1], gz = gt[3 * i + 2];
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(px, gx), __fmul_rn(py, gy)), __fmul_rn(pz, gz));
    const float np = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
    const float ng = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz)));
    float c = __fdiv_rn(dot, __fadd_rn(__fmul_rn(np, ng), 1e-6f));
    c = fminf(fmaxf(c, -1.0f), 1.0f);
    const float e = __fdiv_rn(__fmul_rn(acosf(c), 180.0f), 3.14159265358979323846f);
    if (err_deg != nullptr) err_deg[i] = e;      // the per-pixel map the reference builds before masking (:12-18)
    if (!keep) { err_bits[i] = 0xffffffffu; continue; }
    err_bits[i] = __float_as_uint(e);            // e >= 0: the bit pattern orders like the value
    v[0] += 1.0; v[1] += (double)e; v[2] += (double)__fmul_rn(e, e);
    v[3] += e < 5.0f ? 1.0 : 0.0; v[4] += e < 7.5f ? 1.0 : 0.0; v[5] += e < 11.25f ? 1.0 : 0.0;
    v[6] += e < 22.5f ? 1.0 : 0.0; v[7] += e < 30.0f ? 1.0 : 0.0;
  }
  block_reduce_store<8>(v, part);
}

// ---- exact k-th smallest of the valid entries by radix select: 4 passes of 8 bits, most significant first.
// state: [0] prefix value, [1] prefix mask, [2..3] remaining rank k (64-bit)
__global__ void __launch_bounds__(kThreads)
select_hist_kernel(const unsigned int* __restrict__ bits, long long n, int shift, const unsigned int* __restrict__ state,
                   unsigned long long* __restrict__ hist) {
  __shared__ unsigned int s_h[256];
  s_h[threadIdx.x] = 0u;
  __syncthreads();
  const unsigned int prefix = state[0], pmask = state[1];
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const unsigned int b = bits[i];
    if (b != 0xffffffffu && (b & pmask) == prefix) atomicAdd(&s_h[(b >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (s_h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)s_h[threadIdx.x]);
}
// one 256-thread CTA: inclusive scan of the histogram, the bin holding rank k extends the prefix; re-zeroes hist
__global__ void __launch_bounds__(256)
select_step_kernel(unsigned long long* __restrict__ hist, int shift, unsigned int* __restrict__ state) {
  __shared__ unsigned long long s_c[256];
  const int b = threadIdx.x;
  const unsigned long long k = (unsigned long long)state[2] | ((unsigned long long)state[3] << 32);
  const unsigned int prefix = state[0], pmask = state[1];
  const unsigned long long mine = hist[b];
  s_c[b] = mine;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {                  // Hillis-Steele inclusive scan
    const unsigned long long add = b >= o ? s_c[b - o] : 0ull;
    __syncthreads();
    s_c[b] += add;
    __syncthreads();
  }
  const unsigned long long incl = s_c[b], excl = incl - mine;
  hist[b] = 0ull;
  // exactly one bin satisfies excl <= k < incl when k < total; bin 255 takes the (impossible) overflow
  const bool hit = (k >= excl && k < incl) || (b == 255 && k >= incl);
  if (hit) {
    const unsigned long long rem = k >= incl ? 0ull : k - excl;
    state[0] = prefix | ((unsigned int)b << shift);
    state[1] = pmask | (255u << shift);
    state[2] = (unsigned int)(rem & 0xffffffffull);
    state[3] = (unsigned int)(rem >> 32);
  }
}

}  // namespace

// workspace (bytes) for n pixels: partials + sums + select state / histogram + the per-pixel error buffer
long long metrics_workspace_bytes(long long n) {
  return (long long)kBlocks * 9 * 8 + 64 * 8 + 256 * 8 + 64 + n * 4 + 1024;
}

// out[11]: Abs Rel, Sq Rel, RMSE, Log RMSE, delta<1, delta<1.25, delta<1.25^2, delta<1.25^3, valid_pixels, scale, shift
int launch_depth_metrics(const float* pred, const float* gt, const unsigned char* mask, long long n, float max_depth,
                         void* ws, double* out_host, float* err_map, float* pred_aligned, float* gt_valid,
                         cudaStream_t st) {
  double* part = reinterpret_cast<double*>(ws);
  double* sums = part + (size_t)kBlocks * 9;
  float* stv = reinterpret_cast<float*>(sums + 16);
  const int blocks = (int)((n + kThreads - 1) / kThreads < kBlocks ? (n + kThreads - 1) / kThreads : kBlocks);
  depth_fit_kernel<<<blocks, kThreads, 0, st>>>(pred, gt, n, max_depth, part);
  fold_kernel<5><<<1, 32 * 5, 0, st>>>(part, blocks, sums);
  depth_solve_kernel<<<1, 1, 0, st>>>(sums, stv);
  depth_err_kernel<<<blocks, kThreads, 0, st>>>(pred, gt, mask, n, max_depth, stv, part, err_map, pred_aligned,
                                                gt_valid);
  fold_kernel<9><<<1, 32 * 9, 0, st>>>(part, blocks, sums);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  double h[9];
  float sth[2];
  e = cudaMemcpyAsync(h, sums, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyAsync(sth, stv, sizeof(sth), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return (int)e;
  const double cnt = h[0];
  out_host[9] = sth[0];
  out_host[10] = sth[1];
  if (cnt <= 0.0) {                       // eval_depth.py:217-227: every metric 0 when nothing is valid
    for (int i = 0; i < 9; ++i) out_host[i] = 0.0;
    return 0;
  }
  out_host[0] = h[1] / cnt;
  out_host[1] = h[2] / cnt;
  out_host[2] = sqrt(h[3] / cnt);
  out_host[3] = sqrt(h[4] / cnt);
  for (int i = 0; i < 4; ++i) out_host[4 + i] = h[5 + i] / cnt;
  out_host[8] = cnt;
  return 0;
}

// out[8]: normal mean, median, rmse, angle<5, <7.5, <11.25, <22.5, <30 (percent)
int launch_normal_metrics(const float* pred, const float* gt, const unsigned char* mask, long long n, void* ws,
                          double* out_host, float* err_deg, cudaStream_t st) {
  double* part = reinterpret_cast<double*>(ws);
  double* sums = part + (size_t)kBlocks * 9;
  unsigned long long* hist = reinterpret_cast<unsigned long long*>(sums + 64);
  unsigned int* state = reinterpret_cast<unsigned int*>(hist + 256);
  unsigned int* bits = state + 16;
  const int blocks = (int)((n + kThreads - 1) / kThreads < kBlocks ? (n + kThreads - 1) / kThreads : kBlocks);
  normal_err_kernel<<<blocks, kThreads, 0, st>>>(pred, gt, mask, n, bits, part, err_deg);
  fold_kernel<8><<<1, 32 * 8, 0, st>>>(part, blocks, sums);
  double h[8];
  cudaError_t e = cudaMemcpyAsync(h, sums, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return (int)e;
  const double cnt = h[0];
  if (cnt <= 0.0) {
    for (int i = 0; i < 8; ++i) out_host[i] = nan("");
    return 0;
  }
  // torch.median: the lower of the two middle values -> rank (cnt - 1) / 2 (0-based)
  const unsigned long long k = (unsigned long long)((cnt - 1.0) / 2.0);
  const unsigned int init[4] = {0u, 0u, (unsigned int)(k & 0xffffffffull), (unsigned int)(k >> 32)};
  e = cudaMemcpyAsync(state, init, sizeof(init), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(hist, 0, 256 * 8, st);
  if (e != cudaSuccess) return (int)e;
  for (int shift = 24; shift >= 0; shift -= 8) {
    select_hist_kernel<<<blocks, kThreads, 0, st>>>(bits, n, shift, state, hist);
    select_step_kernel<<<1, 256, 0, st>>>(hist, shift, state);
  }
  unsigned int med_bits = 0;
  e = cudaMemcpyAsync(&med_bits, state, 4, cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return (int)e;
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  float med;
  memcpy(&med, &med_bits, 4);
  out_host[0] = h[1] / cnt;
  out_host[1] = (double)med;
  out_host[2] = sqrt(h[2] / cnt);
  for (int i = 0; i < 5; ++i) out_host[3 + i] = 100.0 * h[3 + i] / cnt;
  return 0;
}

}  // namespace ug

// Per-clip post-processing of the DepthCrafter adapter on the device.
// Replaces, in order (reference file:line):
//   disparity -> depth      /root/reference/model/depthcrafter.py:92-97   channel mean, clip-wide min-max, 1/(x+0.1)
//   backprojection          /root/reference/utils/geometry_utils.py:246-253 (float64, then .float())
//   plane-fit normals       /root/reference/utils/geometry_utils.py:9-70   5x5 un-normalised box sums, +1e-6 I, orientation
//   OpenCV -> OpenGL flip   /root/reference/model/depthcrafter.py:59
// and the frame-layout glue either side of the VAE (x*2-1 + noise on the way in,
// (x/2+0.5).clamp(0,1) -> [T,H,W,3] on the way out: [UPSTREAM] pipeline steps 2,4 and postprocess_video).
//
// Numerics: the depth chain is fp32 with the reference's operation order (bit-exact against the
// golden fixture).  The reference solves the per-pixel 3x3 systems with an fp32 lstsq, which is not
// reproducible run to run; here box sums and the closed-form
// (adjugate) solve run in fp64 -- the exact solution the lstsq approximates (<= 0.04 deg apart).
// HBM bound: one read of the frames, one write of depth + normals; the 5x5 neighbourhood is staged in
// shared memory (tile + 2-pixel halo), so every disparity value is read from HBM ~1.3 times.
#include "kernels.cuh"
#include "ptx.cuh"

namespace ug {
namespace {

constexpr int kTW = 32, kTH = 8, kHalo = 2;
constexpr int kSW = kTW + 2 * kHalo, kSH = kTH + 2 * kHalo;

// res = ((f0 + f1) + f2) / 3 per pixel; per-block (min, max) partials
__global__ void __launch_bounds__(256)
disparity_kernel(const float* __restrict__ frames, long long pixels, float* __restrict__ res,
                 float* __restrict__ part) {
  __shared__ float s_lo[8], s_hi[8];
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels;
       i += (long long)gridDim.x * blockDim.x) {
    const float a = frames[3 * i], b = frames[3 * i + 1], c = frames[3 * i + 2];
    const float r = __fdiv_rn(__fadd_rn(__fadd_rn(a, b), c), 3.0f);
    res[i] = r;
    lo = fminf(lo, r);
    hi = fmaxf(hi, r);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) { lo = fminf(lo, s_lo[k]); hi = fmaxf(hi, s_hi[k]); }
    part[2 * blockIdx.x] = lo;
    part[2 * blockIdx.x + 1] = hi;
  }
}

__global__ void minmax_final_kernel(float* __restrict__ part, int n) {
  __shared__ float s_lo[32], s_hi[32];
  float lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { lo = fminf(lo, part[2 * i]); hi = fmaxf(hi, part[2 * i + 1]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { lo = fminf(lo, s_lo[k]); hi = fmaxf(hi, s_hi[k]); }
    part[2 * n] = lo;          // result slot behind the partials
    part[2 * n + 1] = hi;
  }
}

// depth = 1 / ((res - lo) / (hi - lo) + 0.1) (fp32, reference order)
__device__ __forceinline__ float depth_of(float r, float lo, float range) {
  return __fdiv_rn(1.0f, __fadd_rn(__fdiv_rn(__fsub_rn(r, lo), range), 0.1f));
}

// one thread = one pixel of a 32x8 tile; camera points of tile + halo staged in smem as fp32 (the
// reference rounds the float64 backprojection to float32 before the plane fit)
__global__ void __launch_bounds__(kTW * kTH)
depth_normals_kernel(const float* __restrict__ res, const float* __restrict__ minmax, const float* __restrict__ K,
                     int H, int W, float* __restrict__ depth, float* __restrict__ normals) {
  __shared__ float s_x[kSH][kSW], s_y[kSH][kSW], s_z[kSH][kSW];
  const int t = blockIdx.z;
  const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
  const float lo = minmax[0], range = __fsub_rn(minmax[1], minmax[0]);
  const float* Kt = K + 9 * t;
  const double fx = Kt[0], cx = Kt[2], fy = Kt[4], cy = Kt[5];
  const float* rt = res + (long long)t * H * W;
  for (int i = threadIdx.x; i < kSH * kSW; i += kTW * kTH) {
    const int sy = i / kSW, sx = i % kSW;
    const int gx = x0 + sx - kHalo, gy = y0 + sy - kHalo;
    float px = 0.f, py = 0.f, pz = 0.f;               // zero padding of the box filter
    if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
      const float z = depth_of(rt[(long long)gy * W + gx], lo, range);
      px = (float)(((double)gx - cx) * (double)z / fx);
      py = (float)(((double)gy - cy) * (double)z / fy);
      pz = z;
    }
    s_x[sy][sx] = px; s_y[sy][sx] = py; s_z[sy][sx] = pz;
  }
  __syncthreads();
  const int lx = threadIdx.x % kTW, ly = threadIdx.x / kTW;
  const int gx = x0 + lx, gy = y0 + ly;
  if (gx >= W || gy >= H) return;
  double a = 0, b = 0, c = 0, d = 0, e = 0, f = 0, r0 = 0, r1 = 0, r2 = 0;
#pragma unroll
  for (int dy = 0; dy < 5; ++dy)
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) {
      const double x = s_x[ly + dy][lx + dx], y = s_y[ly + dy][lx + dx], z = s_z[ly + dy][lx + dx];
      a += x * x; b += x * y; c += x * z; d += y * y; e += y * z; f += z * z;
      r0 += x; r1 += y; r2 += z;
    }
  a += 1e-6; d += 1e-6; f += 1e-6;
  // adjugate solve of [[a,b,c],[b,d,e],[c,e,f]] n = r
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double c11 = a * f - c * c, c12 = b * c - a * e, c22 = a * d - b * b;
  const double det = a * c00 + b * c01 + c * c02;
  double nx = (c00 * r0 + c01 * r1 + c02 * r2) / det;
  double ny = (c01 * r0 + c11 * r1 + c12 * r2) / det;
  double nz = (c02 * r0 + c12 * r1 + c22 * r2) / det;
  const double inv = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
  nx *= inv; ny *= inv; nz *= inv;
  const double px = s_x[ly + kHalo][lx + kHalo], py = s_y[ly + kHalo][lx + kHalo], pz = s_z[ly + kHalo][lx + kHalo];
  if (nx * px + ny * py + nz * pz > 0.0) { nx = -nx; ny = -ny; nz = -nz; }   // face the camera
  const long long o = ((long long)t * H + gy) * W + gx;
  depth[o] = (float)pz;
  normals[3 * o] = (float)nx;                          // OpenCV -> OpenGL: y, z negated
  normals[3 * o + 1] = (float)(-ny);
  normals[3 * o + 2] = (float)(-nz);
}

// frames fp32 [T][H][W][3] in [0,1] (+ noise fp32 NCHW * ns) -> 16-bit [T][HW][8] (3 valid): f*2-1 (+ns*noise)
template <typename T>
__global__ void frames_in_kernel(const float* __restrict__ frames, const float* __restrict__ noise, float ns,
                                 long long HW, long long pixels, T* __restrict__ y, float* __restrict__ video_nchw) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i % HW;
    T o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = Elem<T>::from_f(0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = __fsub_rn(__fmul_rn(frames[3 * i + c], 2.0f), 1.0f);
      if (video_nchw) video_nchw[(n * 3 + c) * HW + p] = v;
      if (noise) v = __fadd_rn(v, __fmul_rn(ns, noise[(n * 3 + c) * HW + p]));
      o[c] = Elem<T>::from_f(v);
    }
    *reinterpret_cast<uint4*>(y + 8 * i) = *reinterpret_cast<const uint4*>(o);
  }
}

// images fp32 [T][3][HW] in 0..255 -> frames fp32 [T][HW][3] = float(uint8(v)) / 255: the uint8 TRUNCATION and
// the /255 of model/depthcrafter.py:43-44, bit-exact (values outside [0,256) saturate like a C cast on x86
// would not -- the dataset contract is 0..255, dataset/Readme.md:22-33)
__global__ void images_in_kernel(const float* __restrict__ img, long long HW, long long pixels,
                                 float* __restrict__ frames) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i % HW;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = img[(n * 3 + c) * HW + p];
      const unsigned int u = (unsigned int)fminf(fmaxf(v, 0.f), 255.f);      // cvt.rzi: truncation
      frames[3 * i + c] = __fdiv_rn((float)u, 255.0f);
    }
  }
}

// 16-bit [pixels][8] (3 valid) -> fp32 [pixels][3]: clamp(x / 2 + 0.5, 0, 1)
template <typename T>
__global__ void frames_out_kernel(const T* __restrict__ x, long long pixels, float* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + 8 * i);
    const float2 a = Elem<T>::unpack2(u.x), b = Elem<T>::unpack2(u.y);
    y[3 * i] = fminf(fmaxf(__fadd_rn(__fmul_rn(a.x, 0.5f), 0.5f), 0.f), 1.f);
    y[3 * i + 1] = fminf(fmaxf(__fadd_rn(__fmul_rn(a.y, 0.5f), 0.5f), 0.f), 1.f);
    y[3 * i + 2] = fminf(fmaxf(__fadd_rn(__fmul_rn(b.x, 0.5f), 0.5f), 0.f), 1.f);
  }
}

inline int blocks_for(long long n, int cap = 148 * 8) {
  long long g = (n + 255) / 256;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

long long post_workspace_floats(long long pixels) { return pixels + 2 * (long long)(148 * 8 + 1); }

int launch_depth_postprocess(const float* frames, const float* K, int T, int H, int W, float* depth, float* normals,
                             float* ws, cudaStream_t st) {
  const long long pixels = (long long)T * H * W;
  float* res = ws;
  float* part = ws + pixels;
  const int nb = blocks_for(pixels);
  disparity_kernel<<<nb, 256, 0, st>>>(frames, pixels, res, part);
  minmax_final_kernel<<<1, 1024, 0, st>>>(part, nb);
  const dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH, T);
  depth_normals_kernel<<<grid, kTW * kTH, 0, st>>>(res, part + 2 * nb, K, H, W, depth, normals);
  return (int)cudaGetLastError();
}

int launch_frames_in(const float* frames, const float* noise, float ns, int T, long long HW, void* y,
                     float* video_nchw, int fmt, cudaStream_t st) {
  const long long pixels = (long long)T * HW;
  if (fmt == 1)
    frames_in_kernel<__nv_bfloat16><<<blocks_for(pixels, 148 * 16), 256, 0, st>>>(
        frames, noise, ns, HW, pixels, reinterpret_cast<__nv_bfloat16*>(y), video_nchw);
  else
    frames_in_kernel<__half><<<blocks_for(pixels, 148 * 16), 256, 0, st>>>(frames, noise, ns, HW, pixels,
                                                                            reinterpret_cast<__half*>(y), video_nchw);
  return (int)cudaGetLastError();
}

int launch_images_in(const float* img, int T, long long HW, float* frames, cudaStream_t st) {
  const long long pixels = (long long)T * HW;
  images_in_kernel<<<blocks_for(pixels, 148 * 16), 256, 0, st>>>(img, HW, pixels, frames);
  return (int)cudaGetLastError();
}

int launch_frames_out(const void* x, long long pixels, float* y, int fmt, cudaStream_t st) {
  if (fmt == 1)
    frames_out_kernel<__nv_bfloat16><<<blocks_for(pixels, 148 * 16), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), pixels, y);
  else
    frames_out_kernel<__half><<<blocks_for(pixels, 148 * 16), 256, 0, st>>>(reinterpret_cast<const __half*>(x), pixels, y);
  return (int)cudaGetLastError();
}

}  // namespace ug

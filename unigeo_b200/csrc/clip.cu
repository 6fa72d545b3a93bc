// CLIP image encoder of the DepthCrafter pipeline: replaces [UPSTREAM]
//   encode_video = _resize_with_antialiasing(video, (224, 224)) -> (x + 1) / 2 -> CLIP normalise ->
//                  CLIPVisionModelWithProjection(...).image_embeds            (pipeline step 3, App. A.1)
// reached from /root/reference/model/depthcrafter.py:80-90.  Weight keys: "clip." + the transformers
// state_dict names (vision_model.embeddings.*, vision_model.encoder.layers.N.*, visual_projection.weight).
//
// One kernel does the whole preprocessing: Gaussian blur (reflect border; sigma, taps as upstream), bicubic
// resample (A = -0.75, align_corners = True, clamped taps), CLIP normalisation, and writes the result directly
// in PATCH-MATRIX layout [frames * patches][3 * P * P (+ pad)], so the patch embedding is one plain GEMM.
// The transformer runs on the same tcgen05 GEMM / LayerNorm kernels as the UNet.  Two paddings keep every shape
// TMA / UMMA friendly and are exact: tokens 257 -> 264 per frame (padded keys are masked in the softmax) and
// head_dim 80 -> 128 (zero rows in the fused q|k|v weight, zero columns in out_proj).
#include <cmath>

#include "model.cuh"
#include "ptx.cuh"

namespace ug {
namespace {

const std::string CP = "clip.";
const std::string VM = "clip.vision_model.";

__device__ __forceinline__ int reflect_idx(int i, int n) {      // F.pad(mode="reflect")
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}
__device__ __forceinline__ float cubic1(float x) {              // |x| <= 1, A = -0.75
  return ((1.25f * x - 2.25f) * x) * x + 1.0f;
}
__device__ __forceinline__ float cubic2(float x) {              // 1 < |x| < 2
  return ((-0.75f * x + 3.75f) * x - 6.0f) * x + 3.0f;
}

// video fp32 [F][3][H][W] in [-1,1] -> patch matrix 16-bit [F * np * np][Kp]; k = c*P*P + py*P + px
template <typename T>
__global__ void clip_preprocess_kernel(const float* __restrict__ video, int F, int H, int W, int S, int P, int Kp,
                                       int kh, int kw, float sig_h, float sig_w, T* __restrict__ out) {
  __shared__ float s_kh[32], s_kw[32];
  if (threadIdx.x == 0) {
    float sum = 0.f;
    for (int i = 0; i < kh; ++i) { const float x = (float)i - (kh - 1) * 0.5f; s_kh[i] = expf(-0.5f * (x / sig_h) * (x / sig_h)); sum += s_kh[i]; }
    for (int i = 0; i < kh; ++i) s_kh[i] /= sum;
    sum = 0.f;
    for (int i = 0; i < kw; ++i) { const float x = (float)i - (kw - 1) * 0.5f; s_kw[i] = expf(-0.5f * (x / sig_w) * (x / sig_w)); sum += s_kw[i]; }
    for (int i = 0; i < kw; ++i) s_kw[i] /= sum;
  }
  __syncthreads();
  const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
  const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  const int np = S / P;
  const long long total = (long long)F * 3 * S * S;
  const float ry = S > 1 ? (float)(H - 1) / (float)(S - 1) : 0.f, rx = S > 1 ? (float)(W - 1) / (float)(S - 1) : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % S), y = (int)((i / S) % S), c = (int)((i / ((long long)S * S)) % 3);
    const long long f = i / ((long long)3 * S * S);
    const float* img = video + (f * 3 + c) * (long long)H * W;
    const float sy = ry * y, sx = rx * x;
    const int iy = (int)floorf(sy), ix = (int)floorf(sx);
    const float ty = sy - iy, tx = sx - ix;
    const float wy[4] = {cubic2(ty + 1.f), cubic1(ty), cubic1(1.f - ty), cubic2(2.f - ty)};
    const float wx[4] = {cubic2(tx + 1.f), cubic1(tx), cubic1(1.f - tx), cubic2(2.f - tx)};
    float acc = 0.f;
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), H - 1);
      float row = 0.f;
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), W - 1);
        // blurred(yy, xx): vertical taps first, then horizontal (the upstream conv order)
        float v = 0.f;
        for (int q = 0; q < kw; ++q) {
          const int xs = reflect_idx(xx + q - kw / 2, W);
          float col = 0.f;
          for (int p = 0; p < kh; ++p) col += s_kh[p] * __ldg(img + (long long)reflect_idx(yy + p - kh / 2, H) * W + xs);
          v += s_kw[q] * col;
        }
        row += wx[b] * v;
      }
      acc += wy[a] * row;
    }
    const float pix = ((acc + 1.0f) * 0.5f - mean[c]) / stdv[c];
    const long long prow = (f * np + y / P) * np + x / P;
    const int k = (c * P + y % P) * P + x % P;
    out[prow * Kp + k] = Elem<T>::from_f(pix);
  }
}

// tokens[f][0] = cls + pos[0]; tokens[f][1 + p] = patch_embed[f][p] + pos[1 + p]; rows >= 1 + np2 are zero
template <typename T>
__global__ void clip_assemble_kernel(const T* __restrict__ patches, const float* __restrict__ cls,
                                     const float* __restrict__ pos, int F, int np2, int Npad, int C, T* __restrict__ tok) {
  const long long total = (long long)F * Npad * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), n = (int)((i / C) % Npad);
    const long long f = i / ((long long)C * Npad);
    float v = 0.f;
    if (n == 0) v = cls[c] + pos[c];
    else if (n <= np2) v = Elem<T>::to_f(patches[(f * np2 + (n - 1)) * C + c]) + pos[(long long)n * C + c];
    tok[i] = Elem<T>::from_f(v);
  }
}

// fused, padded q|k|v: rows [which][head][dhp] <- to_{q,k,v} rows [head][dh]; pad rows zero
template <typename T>
__global__ void clip_pad_qkv_kernel(const T* __restrict__ wq, const T* __restrict__ wk, const T* __restrict__ wv,
                                    const float* __restrict__ bq, const float* __restrict__ bk, const float* __restrict__ bv,
                                    int heads, int dh, int dhp, int C, T* __restrict__ W, float* __restrict__ B) {
  const long long rows = (long long)3 * heads * dhp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * C; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int d = (int)(r % dhp), h = (int)((r / dhp) % heads), which = (int)(r / ((long long)dhp * heads));
    const T* src = which == 0 ? wq : which == 1 ? wk : wv;
    const float* bs = which == 0 ? bq : which == 1 ? bk : bv;
    W[i] = d < dh ? src[((long long)h * dh + d) * C + c] : Elem<T>::from_f(0.f);
    if (c == 0) B[r] = d < dh ? bs[h * dh + d] : 0.f;
  }
}
// out_proj columns [head][dh] -> [head][dhp] (zero pad): Wp [C][heads*dhp]
template <typename T>
__global__ void clip_pad_out_kernel(const T* __restrict__ wo, int heads, int dh, int dhp, int C, T* __restrict__ Wp) {
  const int Kp = heads * dhp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)C * Kp; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    const long long r = i / Kp;
    const int d = k % dhp, h = k / dhp;
    Wp[i] = d < dh ? wo[r * ((long long)heads * dh) + h * dh + d] : Elem<T>::from_f(0.f);
  }
}

void add_w(Ctx& c, const std::string& key, void* p, bool f32, int cout, int cin) {
  Weight w;
  w.p = p; w.is_f32 = f32; w.taps = 1; w.cout = cout; w.cin = cin; w.cin_pad = cin;
  w.numel = (long long)cout * cin;
  c.weights[key] = w;
}
inline int grid1d(long long n) { long long g = (n + 255) / 256; return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g)); }
inline int dh_pad(const ug_clip_cfg& g) { return ((g.hidden / g.heads) + 63) / 64 * 64; }

}  // namespace

void clip_finalize(Ctx& c, cudaStream_t st) {
  const ug_clip_cfg& g = c.cfg_clip;
  if (g.layers == 0 || !c.has(VM + "embeddings.patch_embedding.weight")) return;
  if (c.has(VM + "encoder.layers.0.self_attn.qkv_pad.weight")) return;
  const int C = g.hidden, H = g.heads, dh = C / H, dhp = dh_pad(g);
  UG_CHECK(C % H == 0 && (C % 8) == 0, UG_ERR_INVALID, "CLIP hidden size must divide into heads");
  for (int l = 0; l < g.layers; ++l) {
    const std::string k = VM + "encoder.layers." + std::to_string(l) + ".self_attn";
    void* W = c.dmalloc((size_t)3 * H * dhp * C * 2);
    float* B = reinterpret_cast<float*>(c.dmalloc((size_t)3 * H * dhp * 4));
    void* Wo = c.dmalloc((size_t)C * H * dhp * 2);
    if (c.fmt == 1) {
      using T = __nv_bfloat16;
      clip_pad_qkv_kernel<T><<<grid1d((long long)3 * H * dhp * C), 256, 0, st>>>(
          (const T*)c.M(k + ".q_proj.weight"), (const T*)c.M(k + ".k_proj.weight"), (const T*)c.M(k + ".v_proj.weight"),
          c.F(k + ".q_proj.bias"), c.F(k + ".k_proj.bias"), c.F(k + ".v_proj.bias"), H, dh, dhp, C, (T*)W, B);
      clip_pad_out_kernel<T><<<grid1d((long long)C * H * dhp), 256, 0, st>>>((const T*)c.M(k + ".out_proj.weight"), H, dh, dhp, C, (T*)Wo);
    } else {
      using T = __half;
      clip_pad_qkv_kernel<T><<<grid1d((long long)3 * H * dhp * C), 256, 0, st>>>(
          (const T*)c.M(k + ".q_proj.weight"), (const T*)c.M(k + ".k_proj.weight"), (const T*)c.M(k + ".v_proj.weight"),
          c.F(k + ".q_proj.bias"), c.F(k + ".k_proj.bias"), c.F(k + ".v_proj.bias"), H, dh, dhp, C, (T*)W, B);
      clip_pad_out_kernel<T><<<grid1d((long long)C * H * dhp), 256, 0, st>>>((const T*)c.M(k + ".out_proj.weight"), H, dh, dhp, C, (T*)Wo);
    }
    UG_CUDA(cudaGetLastError());
    add_w(c, k + ".qkv_pad.weight", W, false, 3 * H * dhp, C);
    add_w(c, k + ".qkv_pad.bias", B, true, 3 * H * dhp, 1);
    add_w(c, k + ".out_pad.weight", Wo, false, C, H * dhp);
  }
}

// video fp32 [F][3][H][W] in [-1,1] -> image_embeds fp32 [F][proj_dim]
void clip_embed(Ctx& c, const float* video, int F, int H, int W, float* enc) {
  const ug_clip_cfg& g = c.cfg_clip;
  UG_CHECK(g.layers > 0, UG_ERR_STATE, "ug_ctx_set_clip_cfg was not called");
  const int C = g.hidden, S = g.image_size, P = g.patch, np = S / P, np2 = np * np;
  const int ntok = np2 + 1, Npad = (ntok + 7) & ~7;
  const int heads = g.heads, dh = C / heads, dhp = dh_pad(g);
  const int K = 3 * P * P, Kp = (K + 7) & ~7;
  const long long rows = (long long)F * Npad;
  // upstream blur parameters
  const float fh = (float)H / S, fw = (float)W / S;
  const float sh = fmaxf((fh - 1.0f) * 0.5f, 0.001f), sw = fmaxf((fw - 1.0f) * 0.5f, 0.001f);
  int kh = (int)(2.0f * 2.0f * sh) | 1, kw = (int)(2.0f * 2.0f * sw) | 1;
  if (kh < 3) kh = 3;
  if (kw < 3) kw = 3;
  UG_CHECK(kh <= 31 && kw <= 31 && kh / 2 < H && kw / 2 < W, UG_ERR_INVALID, "CLIP resize: frame too large / small");

  void* pm = c.alloc16((long long)F * np2 * Kp);
  void* pe = c.alloc16((long long)F * np2 * C);
  void* x = c.alloc16(rows * C);
  void* x2 = c.alloc16(rows * C);
  void* n = c.alloc16(rows * C);
  void* big = c.alloc16(rows * (3 * heads * dhp > g.mlp ? 3 * heads * dhp : g.mlp));
  void* ao = c.alloc16(rows * heads * dhp);
  void* cls = c.alloc16((long long)F * C);
  void* cls_n = c.alloc16((long long)F * C);
  if (!c.dry) {
    if (Kp != K) UG_CUDA(cudaMemsetAsync(pm, 0, (size_t)F * np2 * Kp * 2, c.stream));
    const long long total = (long long)F * 3 * S * S;
    if (c.fmt == 1)
      clip_preprocess_kernel<__nv_bfloat16><<<grid1d(total), 256, 0, c.stream>>>(video, F, H, W, S, P, Kp, kh, kw, sh, sw, (__nv_bfloat16*)pm);
    else
      clip_preprocess_kernel<__half><<<grid1d(total), 256, 0, c.stream>>>(video, F, H, W, S, P, Kp, kh, kw, sh, sw, (__half*)pm);
    op_check(c, (int)cudaGetLastError(), "clip_preprocess", 0.0, 4.0 * F * 3 * (double)H * W + 2.0 * F * np2 * Kp);
  }
  { Epi e; e.out = pe; e.ldc = C;
    op_linear(c, pm, (long long)F * np2, Kp, Kp, c.M(VM + "embeddings.patch_embedding.weight"), C, e); }
  if (!c.dry) {
    const long long total = rows * C;
    if (c.fmt == 1)
      clip_assemble_kernel<__nv_bfloat16><<<grid1d(total), 256, 0, c.stream>>>((const __nv_bfloat16*)pe, c.F(VM + "embeddings.class_embedding"),
          c.F(VM + "embeddings.position_embedding.weight"), F, np2, Npad, C, (__nv_bfloat16*)n);
    else
      clip_assemble_kernel<__half><<<grid1d(total), 256, 0, c.stream>>>((const __half*)pe, c.F(VM + "embeddings.class_embedding"),
          c.F(VM + "embeddings.position_embedding.weight"), F, np2, Npad, C, (__half*)n);
    op_check(c, (int)cudaGetLastError(), "clip_assemble", 0.0, 4.0 * rows * C);
  }
  op_layernorm(c, n, rows, C, c.F(VM + "pre_layrnorm.weight"), c.F(VM + "pre_layrnorm.bias"), g.ln_eps, nullptr, 1, x);
  const float scale = 1.0f / sqrtf((float)dh);
  for (int l = 0; l < g.layers; ++l) {
    const std::string L = VM + "encoder.layers." + std::to_string(l);
    op_layernorm(c, x, rows, C, c.F(L + ".layer_norm1.weight"), c.F(L + ".layer_norm1.bias"), g.ln_eps, nullptr, 1, n);
    { Epi e; e.out = big; e.ldc = 3 * heads * dhp; e.bias = c.F(L + ".self_attn.qkv_pad.bias");
      op_linear(c, n, rows, C, C, c.M(L + ".self_attn.qkv_pad.weight"), 3 * heads * dhp, e); }
    op_spatial_attention(c, big, F, Npad, heads * dhp, dhp, ao, ntok, scale);
    { Epi e; e.out = x2; e.ldc = C; e.bias = c.F(L + ".self_attn.out_proj.bias"); e.res = x; e.ldr = C;
      op_linear(c, ao, rows, heads * dhp, heads * dhp, c.M(L + ".self_attn.out_pad.weight"), C, e); }
    op_layernorm(c, x2, rows, C, c.F(L + ".layer_norm2.weight"), c.F(L + ".layer_norm2.bias"), g.ln_eps, nullptr, 1, n);
    { Epi e; e.out = big; e.ldc = g.mlp; e.bias = c.F(L + ".mlp.fc1.bias"); e.act = 1;
      op_linear(c, n, rows, C, C, c.M(L + ".mlp.fc1.weight"), g.mlp, e); }
    { Epi e; e.out = x; e.ldc = C; e.bias = c.F(L + ".mlp.fc2.bias"); e.res = x2; e.ldr = C;
      op_linear(c, big, rows, g.mlp, g.mlp, c.M(L + ".mlp.fc2.weight"), C, e); }
  }
  // pooled output = post_layernorm(CLS token) -> visual_projection (no bias)
  if (!c.dry)
    UG_CUDA(cudaMemcpy2DAsync(cls, (size_t)C * 2, x, (size_t)Npad * C * 2, (size_t)C * 2, F, cudaMemcpyDeviceToDevice, c.stream));
  op_layernorm(c, cls, F, C, c.F(VM + "post_layernorm.weight"), c.F(VM + "post_layernorm.bias"), g.ln_eps, nullptr, 1, cls_n);
  { Epi e; e.out = enc; e.ldc = g.proj_dim; e.out_fp32 = 1;
    op_linear(c, cls_n, F, C, C, c.M(CP + "visual_projection.weight"), g.proj_dim, e); }
}

}  // namespace ug

// Library context: device, element format, weight store, bump-arena workspace and the
// op wrappers that the model graphs (unet.cu / vae.cu) are written in.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/unigeo_b200.h"
#include "kernels.cuh"
#include "tapgemm.cuh"

namespace ug {

struct UgError : std::runtime_error {
  int code;
  UgError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define UG_CUDA(expr)                                                                            \
  do {                                                                                           \
    int _e = (int)(expr);                                                                        \
    if (_e != 0)                                                                                 \
      throw ::ug::UgError(UG_ERR_CUDA, std::string(#expr) + " failed: " +                        \
                                           cudaGetErrorString((cudaError_t)_e) + " (" +          \
                                           std::to_string(_e) + ")");                            \
  } while (0)

#define UG_CHECK(cond, code, msg)                                        \
  do {                                                                   \
    if (!(cond)) throw ::ug::UgError((code), std::string(msg));          \
  } while (0)

// Bump allocator over one cudaMalloc'd slab.  In dry mode nothing is backed: addresses
// are offsets from a fake base and only the high-water mark matters.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool dry = false;
  void* alloc(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    off = a + bytes;
    if (off > peak) peak = off;
    if (!dry && off > cap)
      throw UgError(UG_ERR_WORKSPACE, "workspace arena exhausted (" + std::to_string(off) + " > " +
                                          std::to_string(cap) + " bytes)");
    return (dry ? reinterpret_cast<char*>(uintptr_t(4096)) : base) + a;
  }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

struct Weight {
  void* p = nullptr;             // device storage (library-owned)
  int is_f32 = 0;                // 1: fp32 vector/scalars; 0: 16-bit matrix in ctx fmt
  int taps = 1, cout = 0, cin = 0, cin_pad = 0;   // matrix: [taps][cout][cin_pad]
  long long numel = 0;
  std::vector<float> host;       // small tensors (mix_factor) mirrored on the host
};

struct Epi {                     // epilogue description for tapgemm-backed ops
  void* out = nullptr;
  long long ldc = 0;
  int out_fp32 = 0;
  const float* bias = nullptr;
  const float* fbias = nullptr;
  int fbias_ld = 0, fbias_div = 1;
  const void* res = nullptr;
  long long ldr = 0;
  const void* blend = nullptr;
  long long ldb = 0;
  float alpha = 0.f;
  float scale = 1.f;
  int geglu = 0;
  int act = 0;                   // 1: GELU(acc + bias) before residual
  // LayerNorm folded into the GEMM (TapGemmArgs::ln_stat): x is the RAW row, Wm = gamma (.) W, bias = bias + W beta
  const float2* ln_stat = nullptr;
  int ln_parts = 0;              // 0: (mean, rstd) per row; P: (sum, sum of squares) partials per row
  float ln_inv_c = 0.f, ln_eps = 0.f;
  const float* ln_colsum = nullptr;
  // producer side: per-row (sum, sum of squares) partials of the output for the LayerNorm the next GEMM folds in.
  // *stat_parts receives the partial count per row (op_linear falls back to one row_stats pass, = 0, for split-K launches)
  float2* stat_out = nullptr;
  int* stat_parts = nullptr;
  int stat_cap = 0;              // partial slots per row the stat_out buffer has room for
  float stat_eps = 0.f;          // epsilon of that LayerNorm (used by the row_stats fallback, which stores rstd)
};

struct UNetModel;
struct VaeModel;
struct Nets2D;

struct Ctx {
  int device = 0;
  int fmt = 1;                   // 0 fp16, 1 bf16
  int enc_fmt = -1;              // storage / compute format of the VAE ENCODER ("vae.encoder.", "vae.quant_conv."
                                 // matrices and its activations); -1 = same as fmt (ug_ctx_set_vae_encode_dtype)
  ug_model_cfg cfg{};
  std::unordered_map<std::string, Weight> weights;
  Arena ws;                      // activations / temporaries
  bool dry = false;              // size-only pass: no launches
  cudaStream_t stream = nullptr; // stream of the current API call
  long long launches = 0;        // kernels launched since the last reset (bench "gpu_launches")
  UNetModel* unet = nullptr;
  VaeModel* vae = nullptr;
  ug_unet2d_cfg cfg2d{};         // StableNormal path (ug_ctx_set_unet2d_cfg); num_blocks == 0: not configured
  Nets2D* nets2d = nullptr;
  ug_clip_cfg cfg_clip{};        // CLIP image encoder (ug_ctx_set_clip_cfg); layers == 0: not configured
  bool finalized = false;
  unsigned int* gn_counters = nullptr;   // "last CTA" tickets of the fused GroupNorm finalize
  bool attn_materialized = false; // true: head_dim-64 attention through QK^T / softmax / PV GEMMs (A/B debug)
  // prepared clip shape
  int T = 0, h = 0, w = 0;
  std::vector<void*> owned;      // cudaMalloc'd blocks to free at destroy
  unsigned long long ptr_epoch = 0;      // bumped whenever library-owned device addresses may have changed
                                         // (allocation, workspace growth, weights): invalidates captured graphs
  // ---- per-launch profiling (ug_ctx_profile): one event after every launch; a launch's time
  // is the gap to the previous event on the (single) stream
  struct ProfRec { const char* name; double flops, bytes; cudaEvent_t ev; };
  bool profile = false;
  bool profile_shapes = false;           // tapgemm rows keyed by shape (ug_ctx_profile(ctx, 2))
  std::set<std::string> prof_names;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  void prof_mark(const char* name, double flops, double bytes);

  void* dmalloc(size_t bytes);
  const Weight& W(const std::string& key) const;
  bool has(const std::string& key) const { return weights.count(key) != 0; }
  const float* F(const std::string& key) const;     // fp32 vector
  const void* M(const std::string& key) const;      // 16-bit matrix
  void* alloc16(long long elems) { return ws.alloc((size_t)elems * 2); }
  float* allocf(long long elems) { return reinterpret_cast<float*>(ws.alloc((size_t)elems * 4)); }
  void ensure_workspace(size_t bytes);
};

// ---- op wrappers (all enqueue on ctx.stream; no-ops in dry mode) ----------------------
void op_linear(Ctx& c, const void* x, long long M, int K, long long ldx, const void* Wm, int N, const Epi& e);
// x: [Nf][H][W][C] dense.  stride 1: pad 1.  stride 2: pad 1 (asym=0) or pad (0,1,0,1) (asym=1).
void op_conv3x3(Ctx& c, const void* x, int Nf, int H, int W, int C, const void* Wm, int Cout, int stride,
                int asym, const Epi& e);
// x: [T][P][C]; (3,1,1) conv over T with zero padding inside each chunk of `chunk` frames.
void op_tconv3(Ctx& c, const void* x, int T, long long P, int C, const void* Wm, int Cout, int chunk,
               const Epi& e);
// self-attention over N tokens per frame from a fused [F*N][3C] q|k|v buffer, head_dim dh -> out [F*N][C]
// n_valid < N: keys >= n_valid are padding (masked); scale <= 0: 1/sqrt(dh).  head_dim 64 runs the fused flash
// kernel (needs n_valid == N); other multiples of 64 materialise the scores.
void op_spatial_attention(Ctx& c, const void* qkv, int F, int N, int C, int dh, void* out, int n_valid = 0,
                          float scale = 0.f);

void op_gn(Ctx& c, const void* x1, int C1, const void* x2, int C2, long long rows, long long rows_per_set,
           const float* gamma, const float* beta, float eps, int silu, void* y);
void op_layernorm(Ctx& c, const void* x, long long rows, int C, const float* g, const float* b, float eps,
                  const float* add, int add_div, void* y);
// (mean, rstd) of every row of x [rows][C] -> stat [rows] (the statistics a LayerNorm-folding GEMM consumes)
void op_row_stats(Ctx& c, const void* x, long long rows, int C, float eps, float2* stat);
void op_temporal_attention(Ctx& c, const void* qkv, void* out, int T, long long P, int C);
// q [F*N][ldq] against kv [Fk*Lk][2C] (K | V), head_dim 64 -> out [F*N][C]
void op_cross_attention(Ctx& c, const void* q, int ldq, const void* kv, void* out, int F, int N, int C, int Lk,
                        int kv_per_frame);
void op_upsample2x(Ctx& c, const void* x, void* y, int N, int H, int W, int C);
void op_concat(Ctx& c, const void* x1, int C1, const void* x2, int C2, long long rows, void* y);
void op_gemv(Ctx& c, const void* Wm, const float* b, const float* addend, const float* x, float* out, int M,
             int N, int K, int silu_in, int silu_out);
void op_check(Ctx& c, int err, const char* what, double flops = 0.0, double bytes = 0.0);

}  // namespace ug

// HBM-bound kernels of the hot path (CUDA cores, 128-bit coalesced accesses).
// All activations are 16-bit channels-last: [frames][pixels][C] ("token-major").
// fmt: 0 = fp16, 1 = bf16.  Every launcher returns a cudaError_t as int.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ug {

// ---- GroupNorm over channels-last tokens ------------------------------------------
// Statistics set s covers rows [s*rows_per_set, (s+1)*rows_per_set): one frame for a
// spatial GroupNorm, the whole clip for the temporal-resnet GroupNorm.  The input may be
// the channel concatenation [x1 (C1) | x2 (C2)] (UNet skip connections); C = C1 + C2.
// stats: per-CTA partial (sum, sumsq) pairs, gn_partial_floats(...) floats; no zeroing needed,
// no atomics: the result is bit-identical from run to run.
long long gn_partial_floats(int C, long long rows, long long rows_per_set, int G);
// The last CTA of every set also folds the partials into (mean, rstd) (stored behind the partials);
// `counters` = kGnMaxSets zero-initialised uints owned by the caller (left zero again on exit).
constexpr int kGnMaxSets = 1024;
int launch_gn_stats(const void* x1, int C1, const void* x2, int C2, long long rows, long long rows_per_set,
                    int G, float eps, float* stats, unsigned int* counters, int fmt, cudaStream_t st);
// statistics + apply in ONE launch (grid-wide flag between the phases; needs 3 * kGnMaxSets zeroed counters).
// Returns cudaErrorNotSupported (801) when the grid cannot be co-resident: use launch_gn_stats + launch_gn_apply.
int launch_gn_fused(const void* x1, int C1, const void* x2, int C2, long long rows, long long rows_per_set, int G,
                    float eps, float* stats, unsigned int* counters, const float* gamma, const float* beta, int silu,
                    void* y, int fmt, cudaStream_t st);
// the same in one launch over clusters of <= 8 CTAs whose shared memory holds a whole statistics set (x read from
// HBM once, partial sums exchanged through distributed shared memory).  cudaErrorNotSupported when a set does not fit.
int launch_gn_cluster(const void* x1, int C1, const void* x2, int C2, long long rows, long long rows_per_set, int G,
                      float eps, const float* gamma, const float* beta, int silu, void* y, int fmt, cudaStream_t st);
// folds the partials into (mean, rstd) per (set, group), stored behind the partials in `stats`
int launch_gn_finalize(int C, long long rows, long long rows_per_set, int G, float eps, float* stats,
                       cudaStream_t st);
// y = act((x - mean) * rstd * gamma + beta), act = SiLU when silu != 0; y is [rows][C] dense.
int launch_gn_apply(const void* x1, int C1, const void* x2, int C2, long long rows, long long rows_per_set,
                    int G, const float* stats, const float* gamma, const float* beta, float eps, int silu,
                    void* y, int fmt, cudaStream_t st);

// ---- LayerNorm over C, optional per-frame vector added first: LN(x + add[row / add_div]) ----
// (mean, rstd) of every row of x [rows][C] (two-pass fp32, the LayerNorm kernel's own statistics) -> stat [rows]
int launch_row_stats(const void* x, long long rows, int C, float eps, float2* stat, int fmt, cudaStream_t st);
// Finalize-time fold of a LayerNorm's affine into the linear layer behind it (one warp per output row n):
//   Wf[n][k] = round16(gamma[k] W[n][k]);  colsum[n] = sum_k Wf[n][k];  bias_out[n] = bias[n] (nullable) + sum_k beta[k] W[n][k]
int launch_ln_fold_weights(const void* W, const float* gamma, const float* beta, const float* bias, void* Wf,
                           float* colsum, float* bias_out, int N, int K, int fmt, cudaStream_t st);
// y[i] = a * x[i] (fp32 vectors; shape constants at prepare time)
int launch_scale_f32(const float* x, float a, float* y, long long n, cudaStream_t st);
int launch_layernorm(const void* x, long long rows, int C, const float* gamma, const float* beta, float eps,
                     const float* add, int add_div, void* y, int fmt, cudaStream_t st);

// ---- row softmax in place over [rows][n] 16-bit; columns >= n_valid (padded keys) get probability 0 ----
int launch_softmax_rows(void* s, long long rows, int n, int n_valid, float scale, int fmt, cudaStream_t st);

// ---- temporal self-attention: sequences of T tokens per (pixel, head), head_dim 64 ----
// qkv: [T][P][3C] (q | k | v), out: [T][P][C]; heads = C / 64.
int launch_temporal_attention(const void* qkv, void* out, int T, long long P, int C, float scale, int fmt,
                              cudaStream_t st);

// ---- cross-attention against a short context (Lk <= 128 keys), head_dim 64 --------------
// q [F*N][ldq]; kv [Fk*Lk][2C] (K | V), Fk = F when kv_per_frame else 1; out [F*N][C].
int launch_cross_attention(const void* q, int ldq, const void* kv, void* out, int F, int N, int C, int Lk,
                           int kv_per_frame, float scale, int fmt, cudaStream_t st);
// x <- c_x0 * x0 + c_x * x (DDIM "sample"-prediction step on fp32 latents)
int launch_axpby(float* x, const float* x0, float c_x0, float c_x, long long n, cudaStream_t st);
// fp32 [tokens][Cs] -> 16-bit [tokens][Cd], zero padded channels
int launch_f32_to_tokens(const float* x, int Cs, int Cd, long long tokens, void* y, int fmt, cudaStream_t st);
// 16-bit [pixels][Cs] (3 valid) -> unit normals -> uint8 [pixels][3]
int launch_normals_to_u8(const void* x, int Cs, long long pixels, unsigned char* y, int fmt, cudaStream_t st);

// ---- layout / elementwise -------------------------------------------------------------
int launch_upsample2x(const void* x, void* y, int N, int H, int W, int C, cudaStream_t st);   // nearest
int launch_concat(const void* x1, int C1, const void* x2, int C2, long long rows, void* y, cudaStream_t st);
// fp32 NCHW [N][Csrc][H][W] -> 16-bit NHWC [N][H][W][Cdst] (Cdst >= Csrc, zero padded), v*scale+shift (+noise*ns)
int launch_nchw_to_nhwc(const float* x, const float* noise, float noise_scale, float scale, float shift, int N,
                        int Csrc, int H, int W, int Cdst, void* y, int fmt, cudaStream_t st);
// uint8/float HWC frames handled on the host side of the ABI; this one: 16-bit NHWC -> fp32 NCHW, first Cdst ch.
int launch_nhwc_to_nchw(const void* x, int N, int H, int W, int Csrc, int Cdst, float scale, float shift,
                        int clamp01, float* y, int fmt, cudaStream_t st);
// fp32 layout swaps for the (small) latent tensors: y = x * scale
int launch_f32_nchw_to_nhwc(const float* x, int N, long long HW, int C, float scale, float* y, cudaStream_t st);
int launch_f32_nhwc_to_nchw(const float* x, int N, long long HW, int C, float scale, float* y, cudaStream_t st);
// out[n] = act(sum_k W[n][k] * in_act(x[k]) + b[n]); W 16-bit [N][K]; x,b,out fp32.  M rows of x/out.
// addend (nullable, [N]) is added before the output activation.
int launch_gemv(const void* W, const float* b, const float* addend, const float* x, float* out, int M, int N,
                int K, int silu_in, int silu_out, int fmt, cudaStream_t st);
// sinusoid embedding rows: out[i][0:dim/2] = cos(v_i * f_j), out[i][dim/2:] = sin(v_i * f_j)
int launch_sinusoid(const float* vals, int n, int dim, float* out, cudaStream_t st);
// same, values passed by value (n <= 4): timestep / added time ids known on the host
int launch_sinusoid_vals(float v0, float v1, float v2, float v3, int n, int dim, float* out, cudaStream_t st);
// vals[i] = i
int launch_iota(float* out, int n, cudaStream_t st);

// ---- scheduler glue (K10) -------------------------------------------------------------
// UNet input [T][P][8] 16-bit  <-  [ latents/sqrt(sigma^2+1) (4 ch) | cond latents (4 ch) ]
// latents fp32 [T][P][4] (channels-last), cond 16-bit [T][P][4].
int launch_build_unet_input(const float* latents, const void* cond, float sigma, long long tokens, void* y,
                            int fmt, cudaStream_t st);
// Euler v-prediction step in place on fp32 latents; v fp32 [tokens][4].
int launch_euler_step(float* latents, const float* v, float sigma, float sigma_next, long long n,
                      cudaStream_t st);

// ---- adapter-side glue (post.cu) ---------------------------------------------------------
// depth post-processing of one clip: frames fp32 [T][H][W][3] in [0,1], K fp32 [T][3][3] ->
// depth fp32 [T][H][W], normals fp32 [T][H][W][3] (OpenGL frame); ws: post_workspace_floats(T*H*W) floats
long long post_workspace_floats(long long pixels);
int launch_depth_postprocess(const float* frames, const float* K, int T, int H, int W, float* depth, float* normals,
                             float* ws, cudaStream_t st);
// ---- scene stitch (stitch.cu): 2-parameter maps between consecutive clips on their shared frames, chained into clip 0's
// frame; buf = all-gathered overlap frames [world][per_rank][2 (head, tail)][n] fp32; chain = [num_clips][2] doubles (S, T)
long long stitch_workspace_bytes(int num_clips);
int launch_stitch_fit(const float* buf, int world, int per_rank, int num_clips, long long n, int space, float offset,
                      void* ws, double* chain, cudaStream_t st);
int launch_stitch_apply(const float* clip, long long elems, const float* prev_tail, long long n_ov, long long frame_elems,
                        int overlap, const double* chain, int k, int space, float offset, float* out, cudaStream_t st);
// ---- metric kernels (metrics.cu): eval.py:49 / :54 on the device; results land in host doubles (the launchers
// synchronise `st`).  ws: metrics_workspace_bytes(n) bytes.  mask nullable (uint8, 0 = excluded).
long long metrics_workspace_bytes(long long n);
// out[11]: Abs Rel, Sq Rel, RMSE, Log RMSE, delta<1, <1.25, <1.25^2, <1.25^3, valid_pixels, scale, shift;
// err_map / pred_aligned / gt_valid: nullable fp32 [n] maps (eval_depth.py:166-213)
int launch_depth_metrics(const float* pred, const float* gt, const unsigned char* mask, long long n, float max_depth,
                         void* ws, double* out_host, float* err_map, float* pred_aligned, float* gt_valid,
                         cudaStream_t st);
// out[8]: normal mean, median, rmse, angle<5, <7.5, <11.25, <22.5, <30 (percent); err_deg nullable fp32 [n]
int launch_normal_metrics(const float* pred, const float* gt, const unsigned char* mask, long long n, void* ws,
                          double* out_host, float* err_deg, cudaStream_t st);
// frames fp32 [T][HW][3] in [0,1] (+ noise fp32 [T][3][HW] * ns) -> 16-bit [T][HW][8]; video_nchw (nullable)
// receives frames*2-1 as fp32 [T][3][HW] (the CLIP branch's input)
int launch_frames_in(const float* frames, const float* noise, float ns, int T, long long HW, void* y,
                     float* video_nchw, int fmt, cudaStream_t st);
// images fp32 [T][3][HW] 0..255 -> frames fp32 [T][HW][3] = float(uint8(v)) / 255
int launch_images_in(const float* img, int T, long long HW, float* frames, cudaStream_t st);
// 16-bit [pixels][8] -> fp32 [pixels][3] = clamp(x/2+0.5, 0, 1)
int launch_frames_out(const void* x, long long pixels, float* y, int fmt, cudaStream_t st);

// weights: src fp32/16-bit [Cout][Cin][taps] -> dst 16-bit [taps][Cout][CinPad] (dst pre-zeroed when padded)
int launch_convert_weight(const void* src, int src_dtype /*0 f16,1 bf16,2 f32*/, void* dst, int Cout, int Cin,
                          int CinPad, int taps, int fmt, cudaStream_t st);
// split-K reduce of tapgemm partials [S][M][N] fp32 -> out (16-bit or fp32, row stride ldc) with the fused epilogue's
// operations in the fused path's order (N % 4 == 0)
int launch_splitk_reduce(const float* part, int S, long long M, int N, const float* bias, const float* fbias, int fbias_ld,
                         int fbias_div, const void* res, long long ldr, const void* blend, long long ldb, float alpha,
                         float scale, int act, void* out, long long ldc, int out_fp32, int fmt, cudaStream_t st);
// one launch for a whole state dict: descriptors (device memory) sorted by first_block; dst_fmt 0 fp16 / 1 bf16 matrix
// ([cout][cin][taps] -> [tap][cout][cin_pad]), 2 fp32 vector copy
struct ConvertDesc {
  const void* src;
  void* dst;
  long long total;         // source elements
  long long first_block;   // first CTA serving this tensor
  int src_dtype, dst_fmt;
  int cout, cin, cin_pad, taps;
};
int launch_convert_batch(const ConvertDesc* dev_descs, int n, long long total_blocks, cudaStream_t st);
int launch_convert_f32(const void* src, int src_dtype, float* dst, long long n, cudaStream_t st);

}  // namespace ug

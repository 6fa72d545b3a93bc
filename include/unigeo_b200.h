/* unigeo_b200 -- C ABI of the B200-native hot path of SunYangtian/UniGeo.
 *
 * The reference has no FFI of its own: its hot path is the Python call
 *     self.pipeline(frames, ...)            /root/reference/model/depthcrafter.py:80-90
 * into [UPSTREAM] diffusers/DepthCrafter modules.  Each entry
 * point below names the reference-side call it replaces; INTEGRATION.md shows the
 * ctypes binding a maintainer adds to model/depthcrafter.py.
 *
 * Conventions: every function returns 0 (UG_OK) or a negative ug_status and never throws
 * across the ABI; ug_last_error() gives the message (thread-local, valid until the next
 * call on that thread).  The CALLER allocates all inputs/outputs (device pointers, e.g.
 * torch.Tensor.data_ptr()); the library owns only its weight copies and its workspace.
 * `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 * A ug_ctx is bound to one device and is not thread-safe; distinct contexts are
 * independent (one per GPU rank).  Batch size is 1 clip per call, as in the reference
 * (eval.py:33-39 feeds one clip at a time).
 *
 * Tensor layouts at the ABI are the upstream ones (fp32, NCHW per frame) so the parity
 * tests read like upstream code; inside, activations are 16-bit channels-last.
 */
#ifndef UNIGEO_B200_H
#define UNIGEO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UG_VERSION 102

typedef struct ug_ctx ug_ctx;

typedef enum ug_status {
  UG_OK = 0,
  UG_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  UG_ERR_CUDA = -2,      /* CUDA runtime or driver error */
  UG_ERR_WEIGHT = -3,    /* missing / mis-shaped weight */
  UG_ERR_WORKSPACE = -4, /* workspace arena exhausted */
  UG_ERR_STATE = -5      /* call order (finalize / prepare / set_clip_context missing) */
} ug_status;

typedef enum ug_dtype { UG_F16 = 0, UG_BF16 = 1, UG_F32 = 2 } ug_dtype;

/* Architecture description (mirrors unigeo_b200/config.py). */
typedef struct ug_model_cfg {
  int32_t dtype; /* storage/compute type of activations and matrices: UG_F16 or UG_BF16 */
  /* spatio-temporal UNet */
  int32_t unet_in_channels, unet_out_channels;
  int32_t unet_num_blocks;
  int32_t unet_block_out[4];
  int32_t unet_heads[4];
  int32_t unet_layers_per_block;
  int32_t cross_attention_dim;
  int32_t addition_time_embed_dim;
  int32_t num_added_ids;
  int32_t norm_groups;
  float eps_cross_attn_block, eps_plain_block, eps_plain_up_block, eps_transformer_norm, eps_out_norm, ln_eps;
  /* temporal-decoder VAE */
  int32_t vae_in_channels, vae_latent_channels;
  int32_t vae_num_blocks;
  int32_t vae_block_out[4];
  int32_t vae_layers_per_block;
  int32_t vae_norm_groups;
  float vae_eps, vae_temporal_eps, vae_scaling_factor;
  /* Euler / Karras scheduler */
  float sigma_min, sigma_max, rho;
} ug_model_cfg;

/* 2-D conditional UNet / ControlNet of the StableNormal path (SD-2.1 class; SURVEY.md App. A.5,
 * reference call site model/stablenormal.py:16,39).  Optional: set before ug_ctx_finalize when weights
 * of such a network are loaded.  The 2-D VAE ("vae2d." keys) shares the vae_* fields of ug_model_cfg. */
typedef struct ug_unet2d_cfg {
  int32_t in_channels, out_channels;
  int32_t num_blocks;
  int32_t block_out[4];
  int32_t heads[4];
  int32_t layers_per_block;
  int32_t cross_attention_dim;
  float eps_resnet, eps_transformer_norm, ln_eps;
  /* DDIM scheduler: scaled-linear betas */
  int32_t num_train_timesteps;
  float beta_start, beta_end;
} ug_unet2d_cfg;

/* CLIP image encoder (ViT, transformers CLIPVisionModelWithProjection layout) of the DepthCrafter pipeline's
 * encode_video step.  Optional: set before ug_ctx_finalize when "clip." weights are
 * loaded.  ViT-H/14: hidden 1280, 32 layers, 16 heads, mlp 5120, patch 14, image 224, proj_dim 1024. */
typedef struct ug_clip_cfg {
  int32_t hidden, layers, heads, mlp, patch, image_size, proj_dim;
  float ln_eps;
} ug_clip_cfg;

int ug_version(void);
const char* ug_last_error(void);

/* Replaces DepthCrafter.__init__'s from_pretrained + .to(cuda) (model/depthcrafter.py:18-31). */
int ug_ctx_create(ug_ctx** out, int device, const ug_model_cfg* cfg);
int ug_ctx_destroy(ug_ctx* ctx);
/* One tensor of a diffusers state_dict ("unet." / "vae." + diffusers key), any of f16/bf16/f32,
 * original diffusers shape.  Copied and re-laid-out into library storage. */
int ug_ctx_load_weight(ug_ctx* ctx, const char* key, const void* dev_ptr, int dtype, const int64_t* shape,
                       int rank, void* stream);
/* A whole state dict in one call and ONE conversion launch (n tensors; shapes holds 5 int64 per tensor, the first
 * ranks[i] of them valid).  Same result as n ug_ctx_load_weight calls; the source tensors may be released on return. */
int ug_ctx_load_weights(ug_ctx* ctx, int n, const char* const* keys, const void* const* dev_ptrs, const int* dtypes,
                        const int64_t* shapes, const int* ranks, void* stream);
/* Storage / compute type of the VAE ENCODER alone (UG_F16, UG_BF16, or -1 = the context's dtype; call before the
 * "vae.encoder." / "vae.quant_conv." (and "vae2d." ditto) weights are loaded).  [UPSTREAM] encode_vae_video upcasts the VAE to fp32
 * (force_upcast) because SVD's encoder activations leave the fp16 range; UG_BF16 keeps the
 * fp32 exponent range on the 16-bit tensor-core path while the rest of an fp16 context stays fp16. */
int ug_ctx_set_vae_encode_dtype(ug_ctx* ctx, int dtype);
/* Build fused matrices (QKV, GEGLU-interleaved FF, stacked time-embedding projections). */
int ug_ctx_finalize(ug_ctx* ctx, void* stream);
/* Size the workspace for clips of T frames with h x w latents (H = 8h, W = 8w) and precompute
 * shape-only constants.  Must be called outside CUDA-graph capture. */
int ug_ctx_prepare(ug_ctx* ctx, int T, int h, int w, void* stream);

/* Per-clip constants of the single-token cross-attentions (SURVEY.md App. A.3 "exact
 * simplification"): enc fp32 [T][cross_attention_dim], the CLIP image embeddings upstream
 * passes as encoder_hidden_states (pipeline step 3). */
int ug_set_clip_context(ug_ctx* ctx, const float* enc, void* stream);

/* Replaces unet(x_in, t, encoder_hidden_states, added_time_ids)[0] (pipeline step 8):
 * x fp32 [1][T][8][h][w]; added_time_ids HOST float[3]; out fp32 [1][T][4][h][w].
 * Needs ug_set_clip_context for this clip. */
int ug_unet_st_forward(ug_ctx* ctx, const float* x, float timestep, const float* added_time_ids_host,
                       float* out, void* stream);

/* Replaces the whole denoising loop (scheduler.set_timesteps / scale_model_input / unet / step,
 * pipeline steps 7-8): cond_lat, init_noise, lat_out fp32 [T][4][h][w]; init_noise is the raw
 * N(0,1) draw (scaled by init_noise_sigma inside). */
int ug_denoise_clip(ug_ctx* ctx, const float* cond_lat, const float* init_noise,
                    const float* added_time_ids_host, int steps, float* lat_out, void* stream);

/* Replaces vae.encode(video).latent_dist.mode() (pipeline steps 4-5):
 * img fp32 [N][3][H][W] in [-1,1]; noise (nullable) fp32 same shape, added * noise_strength;
 * lat_mean fp32 [N][4][H/8][W/8] (not multiplied by the scaling factor). */
int ug_vae_encode(ug_ctx* ctx, const float* img, const float* noise, float noise_strength, int N, int H,
                  int W, float* lat_mean, void* stream);

/* Replaces decode_latents (pipeline step 9): lat fp32 [T][4][h][w] -> img fp32 [T][3][8h][8w];
 * divides by the scaling factor, decodes in chunks of `chunk` frames. */
int ug_vae_decode_temporal(ug_ctx* ctx, const float* lat, int T, int h, int w, int chunk, float* img,
                           void* stream);

/* ---- StableNormal path: 2-D UNet (+ ControlNet) over per-frame latents ---------------------------
 * Networks are addressed by the key prefix their weights were loaded under ("unet2d.", "controlnet.",
 * "yoso_unet.", ...); ug_ctx_finalize discovers every 2-D network among the loaded weights. */
int ug_ctx_set_unet2d_cfg(ug_ctx* ctx, const ug_unet2d_cfg* cfg);
/* encoder_hidden_states of network `net_prefix`: tokens fp32 [frames][len][cross_attention_dim], frames = 1
 * (one prompt shared by all frames, StableNormal's fixed prompt) or the frame count; len <= 128.
 * Precomputes K | V of every cross-attention (they do not depend on the latents or the timestep). */
int ug_set_text_context(ug_ctx* ctx, const char* net_prefix, const float* tokens, int frames, int len,
                        void* stream);
/* Replaces unet(sample, t, encoder_hidden_states, down_block_additional_residuals=controlnet(...)...)[0]:
 * x, controlnet_sample fp32 [F][in_channels][h][w]; out fp32 [F][out_channels][h][w];
 * controlnet_prefix / controlnet_sample nullable (plain UNet). */
int ug_unet2d_forward(ug_ctx* ctx, const char* unet_prefix, const float* x, int F, int h, int w, float timestep,
                      const char* controlnet_prefix, const float* controlnet_sample, float* out, void* stream);
/* Replaces the refinement loop of the hub predictor (DDIM, prediction_type "sample", eta 0, trailing
 * spacing from t_start, or from num_train_timesteps-1 when t_start < 0): image_latent, latents_in,
 * latents_out fp32 [F][4][h][w] (scaled latents). */
int ug_refine_frames_2d(ug_ctx* ctx, const char* unet_prefix, const char* controlnet_prefix,
                        const float* image_latent, const float* latents_in, int F, int h, int w, int steps,
                        int t_start, float* latents_out, void* stream);
/* 2-D AutoencoderKL ("vae2d." keys): encode = latent_dist.mode() * out_scale; decode: lat / scaling_factor ->
 * post_quant_conv -> decoder.  img (nullable) fp32 [N][3][8h][8w]; normals_u8 (nullable) uint8 [N][8h][8w][3] =
 * unit-normalised, clipped, ((n+1)/2*255) truncated -- the 8-bit image the predictor returns
 * (model/stablenormal.py:39-40). */
int ug_vae2d_encode(ug_ctx* ctx, const float* img, int N, int H, int W, float out_scale, float* lat,
                    void* stream);
int ug_vae2d_decode(ug_ctx* ctx, const float* lat, int N, int h, int w, float* img, unsigned char* normals_u8,
                    void* stream);

/* ---- adapter-side glue of the DepthCrafter plugin, on the device ------------------------------------
 * ug_vae_encode_frames = ug_vae_encode on the pipeline's own input layout: frames fp32 [T][H][W][3] in [0,1]
 * (model/depthcrafter.py:39-45 output), x*2-1 and the noise augmentation (noise fp32 [T][3][H][W], nullable)
 * fused into the layout conversion; video_nchw (nullable) receives x*2-1 as fp32 [T][3][H][W] for the CLIP branch.
 * ug_vae_decode_frames = ug_vae_decode_temporal + postprocess_video("np"): frames fp32 [T][8h][8w][3] =
 * clamp(img/2+0.5, 0, 1), the `.frames[0]` the reference reads at model/depthcrafter.py:80-90. */
/* Replaces DepthCrafter.prepare_input (model/depthcrafter.py:39-45) on the device: images fp32 [T][3][H][W]
 * in 0..255 (the dataset dict's `images`, stacked) -> frames fp32 [T][H][W][3] = float(uint8(v)) / 255. */
int ug_prepare_frames(ug_ctx* ctx, const float* images, int T, int H, int W, float* frames, void* stream);
int ug_vae_encode_frames(ug_ctx* ctx, const float* frames, const float* noise, float noise_strength, int T, int H,
                         int W, float* video_nchw, float* lat_mean, void* stream);
int ug_vae_decode_frames(ug_ctx* ctx, const float* lat, int T, int h, int w, int chunk, float* frames, void* stream);
/* Replaces model/depthcrafter.py:92-97 (channel mean, clip-wide min-max, 1/(x+0.1)) and :48-69 (backprojection
 * utils/geometry_utils.py:246-253, plane-fit normals :9-70, OpenCV->OpenGL flip): frames fp32 [T][H][W][3],
 * intrinsics fp32 [T][3][3] (device) -> depths fp32 [T][H][W], normals fp32 [T][H][W][3]. */
int ug_depth_postprocess(ug_ctx* ctx, const float* frames, const float* intrinsics, int T, int H, int W,
                         float* depths, float* normals, void* stream);

/* ---- scene stitch (BASELINE cfg4): the one exchange step of the clip-sharded path.  Clips of a scene
 * (dataset/scannetpp/scannetpp.py:42-48, clip_overlap shared frames) are min-max normalised one by one
 * (model/depthcrafter.py:95), so consecutive clips disagree by a 2-parameter map on their shared frames.  After ONE
 * all-gather of every clip's first / last `overlap` frames (NCCL; unigeo_b200/sharding.py) every rank calls
 * ug_stitch_fit: overlap_frames fp32 [world][per_rank][2 (head, tail)][n] as gathered (clip k on rank k % world, slot
 * k / world), n = overlap * H * W; space 1 fits on x = 1 / depth - offset (the normalised disparity of
 * model/depthcrafter.py:96, where the map IS affine), 0 on the values as given; chain_dev (device) receives
 * [num_clips][2] doubles (S, T) mapping clip k into clip 0's frame (fp64 fixed-order sums, a constant overlap keeps
 * the scale and matches the means).  ug_stitch_apply maps one clip (elems = T * H * W values) and ramps its first
 * `overlap` frames from prev_tail (the previous clip's last frames, NULL for clip 0) with linspace(0, 1, overlap).
 * Not in the reference (it never stitches: window_size = len(frames), model/depthcrafter.py:87): an ADDITIONAL output. */
int ug_stitch_fit(ug_ctx* ctx, const float* overlap_frames, int world, int per_rank, int num_clips, long long n,
                  int space, float offset, double* chain_dev, void* stream);
int ug_stitch_apply(ug_ctx* ctx, const float* clip, long long elems, const float* prev_tail, long long n_overlap,
                    long long frame_elems, int overlap, const double* chain_dev, int k, int space, float offset,
                    float* out, void* stream);

/* ---- consumer side of the plugin boundary: the two metric functions eval.py applies to the outputs, on the device
 * (SURVEY.md 8(f)-3), so that depths / normals can be scored where they were produced.
 * ug_depth_metrics replaces depth_evaluation(pred, gt, custom_mask=mask, align_with_lstsq=True)
 * (metrics/eval_depth.py:6-246 with metrics/alignment.py:150-167; call site eval.py:49): pred, gt fp32 [n] on the
 * device (any [Nf][H][W] flattened), mask uint8 [n] (nullable, 0 = excluded), valid = gt > 0 && gt < max_depth;
 * scale/shift fitted over `valid`, errors over valid && mask.  out11 (HOST doubles): Abs Rel, Sq Rel, RMSE,
 * Log RMSE, delta<1, delta<1.25, delta<1.25^2, delta<1.25^3, valid_pixels, scale, shift (all 0 when nothing is
 * valid, eval_depth.py:217-227).  err_map / pred_aligned / gt_valid: nullable fp32 [n] device outputs = the three
 * full-size maps the reference returns beside the dict (eval_depth.py:166-213).
 * ug_normal_metrics replaces normal_evaluation (metrics/eval_normal.py:4-72; call site eval.py:54): pred, gt fp32
 * [n][3], mask uint8 [n] (nullable).  out8 (HOST doubles): normal mean, median (torch.median: lower middle value,
 * exact), rmse, angle<5, <7.5, <11.25, <22.5, <30 in percent (NaN when the mask is empty); err_deg: nullable fp32
 * [n] device output = the per-pixel angular error in degrees before masking (eval_normal.py:12-18).
 * Both synchronise `stream` before returning (they return host scalars). */
int ug_depth_metrics(ug_ctx* ctx, const float* pred, const float* gt, const unsigned char* mask, long long n,
                     float max_depth, double* out11, float* err_map, float* pred_aligned, float* gt_valid,
                     void* stream);
int ug_normal_metrics(ug_ctx* ctx, const float* pred, const float* gt, const unsigned char* mask, long long n,
                      double* out8, float* err_deg, void* stream);

/* Replaces [UPSTREAM] encode_video (antialiased 224x224 resize, CLIP normalisation, ViT + projection):
 * video fp32 [F][3][H][W] in [-1,1] (ug_vae_encode_frames' video_nchw) -> image embeddings fp32 [F][proj_dim],
 * the encoder_hidden_states ug_set_clip_context takes.  Weight keys: "clip." + transformers state_dict names;
 * patch_embedding.weight is loaded reshaped to [hidden][3*patch*patch], position_embedding.weight flattened. */
int ug_ctx_set_clip_cfg(ug_ctx* ctx, const ug_clip_cfg* cfg);
int ug_clip_embed(ug_ctx* ctx, const float* video, int F, int H, int W, float* enc, void* stream);

/* Host-only schedule tables (no device needed), exactly what the two step loops iterate over.
 * ug_karras_schedule: [UPSTREAM] EulerDiscreteScheduler.set_timesteps (Karras sigmas, "leading"): sigmas[steps+1]
 * (last 0), UNet timesteps[steps] = 0.25 ln sigma (nullable), init_noise_sigma (nullable).
 * ug_ddim_schedule: trailing-spaced DDIM timesteps[steps] and the per-step update x <- c_x0 x0 + c_x x
 * (prediction_type "sample", eta 0; both nullable). */
int ug_karras_schedule(const ug_model_cfg* cfg, int steps, double* sigmas, double* timesteps, double* init_noise_sigma);
int ug_ddim_schedule(const ug_unet2d_cfg* cfg, int steps, int t_start, int* timesteps, double* c_x0, double* c_x);

/* Host-only: the per-CTA unit lists tapgemm uses when the last N tile of a launch is ragged (balanced by list
 * scheduling under the tile picker's cost model; every unit exactly once).  m_units = M tiles (pairs of M tiles when
 * ctas == 2), slots = CTAs (CTA pairs) of the persistent grid, k_iters = taps x 64-channel chunks.  units_out
 * (nullable) receives a [slots][len] table (-1 = none), cap = its capacity in ints; returns len (>= 1) or a negative
 * ug_status.  max_cost / rr_max_cost (nullable): modelled cost of the heaviest CTA under this assignment / under
 * round-robin.  No device needed. */
int ug_tile_schedule(int m_units, int n_total, int bn_tile, int batch, int n_fastest, int ctas, int slots, int k_iters,
                     int* units_out, int cap, long long* max_cost, long long* rr_max_cost);

/* Kernels launched on behalf of this context since the last reset (bench "gpu_launches"). */
long long ug_ctx_launch_count(ug_ctx* ctx, int reset);
/* Bytes of workspace currently reserved. */
long long ug_ctx_workspace_bytes(ug_ctx* ctx);
/* CUDA graphs currently instantiated: ug_denoise_clip / ug_refine_frames_2d run their step loop eagerly on the first
 * call with a given signature, capture it on the second and replay it afterwards (UG_NO_GRAPH=1 keeps them eager). */
long long ug_ctx_graph_count(ug_ctx* ctx);
/* Host-only: n / d the way the persistent GEMM kernel's per-tile index decoding computes it (multiply-high by a launch
 * constant + shift, tapgemm.cuh FastDiv), for 1 <= d, 0 <= n < 2^31; negative ug_status otherwise.  Exists so that the
 * CPU suite can pin the arithmetic against integer division.  No device needed. */
long long ug_fastdiv(int d, int n);

/* Per-launch profiling for bench.py's roofline: while enabled, one CUDA event is recorded on the
 * call's stream after every kernel launch; ug_ctx_profile_read synchronises and aggregates by kernel
 * name (returns the number of rows written, or a negative ug_status).  Off by default. */
int ug_ctx_profile(ug_ctx* ctx, int enable);
int ug_ctx_profile_read(ug_ctx* ctx, int cap, char* names /* [cap][64] */, long long* launches, double* ms,
                        double* flops, double* bytes);

/* ---- single-op entry points (kernel parity tests; same kernels the graphs above use) ----
 * 16-bit tensors are in `dtype` (UG_F16 / UG_BF16), channels-last, dense. */
/* y[M][N] = x[M][K] W[N][K]^T (+bias) (+res) ; geglu: N = 2*Nout, W rows already interleaved by 64 */
int ug_op_linear(int dtype, const void* x, long long M, int K, const void* W, int N, const float* bias,
                 const void* res, int geglu, int out_fp32, void* y, void* stream);
/* y = alpha * blend + (1 - alpha) * (x W^T + bias (+res)): the AlphaBlender of a TransformerSpatioTemporalModel fused
 * into the temporal block's last linear layer ([UPSTREAM] diffusers, reached from model/depthcrafter.py:80-90); blend
 * [M][N] is a different tensor than res here (both arrive by TMA) */
int ug_op_linear_blend(int dtype, const void* x, long long M, int K, const void* W, int N, const float* bias,
                       const void* res, const void* blend, float alpha, void* y, void* stream);
/* x [Nf][H][W][C], Wt [9][Cout][C] -> y [Nf][H/stride][W/stride][Cout] */
int ug_op_conv3x3(int dtype, const void* x, int Nf, int H, int W, int C, const void* Wt, int Cout, int stride,
                  int asym_pad, const float* bias, const void* res, void* y, void* stream);
/* x [T][P][C], Wt [3][Cout][C]; blend != NULL: y = alpha*blend + (1-alpha)*(conv + bias + res) */
int ug_op_tconv3(int dtype, const void* x, int T, long long P, int C, const void* Wt, int Cout, int chunk,
                 const float* bias, const void* res, const void* blend, float alpha, void* y, void* stream);
/* GroupNorm(+SiLU) over [rows][C1+C2] (x2 nullable), sets of rows_per_set rows */
int ug_op_groupnorm(int dtype, const void* x1, int C1, const void* x2, int C2, long long rows,
                    long long rows_per_set, int groups, const float* gamma, const float* beta, float eps,
                    int silu, void* y, void* stream);
int ug_op_layernorm(int dtype, const void* x, long long rows, int C, const float* gamma, const float* beta,
                    float eps, const float* add, int add_div, void* y, void* stream);
/* y[M][N] = LayerNorm(x[M][K]; gamma, beta, eps) W[N][K]^T (+bias) in the FOLDED form the UNet graph uses for every
 * LayerNorm of a transformer block (replaces [UPSTREAM] diffusers BasicTransformerBlock / TemporalBasicTransformerBlock
 * norm1 / norm3 / norm_in + to_q|k|v / ff.net.0.proj, reached from /root/reference/model/depthcrafter.py:80-90): the
 * GEMM reads the raw rows against gamma-scaled weights and its epilogue applies rstd * (acc - mean * colsum) + (bias +
 * W beta).  x0 == NULL: x is given and one row_stats pass provides (mean, rstd).  x0 != NULL: x = x0[M][K0] W0[K][K0]^T
 * (+bias0) (+res0) is computed first (and written to x) and the row statistics come out of THAT GEMM's epilogue. */
int ug_op_ln_linear(int dtype, const void* x0, int K0, const void* W0, const float* bias0, const void* res0, void* x,
                    long long M, int K, const float* gamma, const float* beta, float eps, const void* W, int N,
                    const float* bias, int geglu, void* y, void* stream);
/* qkv [F*N][3C] -> y [F*N][C]; softmax(q k^T / sqrt(dh)) v per frame and head (dh = 64 or C) */
int ug_op_spatial_attention(int dtype, const void* qkv, int F, int N, int C, int dh, void* y, void* stream);
/* qkv [T][P][3C] -> y [T][P][C]; attention over T per pixel and 64-wide head */
int ug_op_temporal_attention(int dtype, const void* qkv, int T, long long P, int C, void* y, void* stream);
/* q [F*N][C], kv [Fk*Lk][2C] (K | V; Fk = F if kv_per_frame else 1) -> y [F*N][C]; head_dim 64, Lk <= 128 */
int ug_op_cross_attention(int dtype, const void* q, const void* kv, int F, int N, int C, int Lk, int kv_per_frame,
                          void* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIGEO_B200_H */

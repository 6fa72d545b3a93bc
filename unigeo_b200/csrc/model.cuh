// Model graphs written over the op wrappers: the spatio-temporal UNet (unet.cu) and the
// temporal-decoder VAE (vae.cu).  Weight keys are "unet." / "vae." + the diffusers key.
#pragma once
#include "ctx.cuh"

namespace ug {

struct Act {          // [frames][H*W][C] 16-bit channels-last activation
  void* p = nullptr;
  int C = 0, H = 0, W = 0;
};

struct UNetModel {
  // per-clip constants of the single-token cross-attentions (ug_set_clip_context)
  std::unordered_map<std::string, float*> attn2_spatial;   // key -> [T][C]
  std::unordered_map<std::string, float*> attn2_temporal;  // key -> [C]
  std::unordered_map<std::string, float*> time_pos;        // key -> [T][C]  (shape constant)
  std::unordered_map<std::string, float*> time_pos_blend; // key -> [T][C] = -alpha / (1 - alpha) * time_pos (LayerNorm fold)
  bool ln_fold = false;                                    // LayerNorms folded into the GEMMs behind them (unet_finalize)
  std::unordered_map<std::string, int> temb_offset;        // resnet key -> offset in temb_out
  int temb_total = 0;
  float* temb_out = nullptr;                                // [temb_total] per-step conv1 biases
  float* scratch = nullptr;                                 // small fp32 scratch for embeddings
  bool clip_context_set = false;
  int prepared_T = 0;
};

struct VaeModel {
  int dummy = 0;
};

// One 2-D conditional UNet or ControlNet (StableNormal path), addressed by its weight-key prefix.
struct Net2D {
  std::string prefix;                                   // "unet2d." / "controlnet." / ...
  bool controlnet = false;
  std::unordered_map<std::string, int> temb_offset;     // resnet key -> offset in temb_out
  int temb_total = 0;
  float* temb_out = nullptr;                            // per-step conv1 biases of every resnet
  float* scratch = nullptr;
  std::unordered_map<std::string, void*> kv;            // transformer key -> K | V of the context [Fk*Lk][2C]
  int ctx_len = 0, ctx_frames = 0;
};
struct Nets2D {
  std::unordered_map<std::string, Net2D> nets;
};

// shared blocks -------------------------------------------------------------------------
// ResnetBlock2D on x = [x1 | x2] (x2 optional): returns new activation [frames][HW][cout].
// bias1: per-step (time-embedding) bias for conv1, or nullptr -> conv1.bias.
Act resnet2d(Ctx& c, const std::string& key, const Act& x1, const Act* x2, int frames, int cout,
             const float* bias1, float eps);
// SpatioTemporalResBlock; chunk = frames per temporal clip (T for the UNet, <=8 for VAE decode)
Act st_resblock(Ctx& c, const std::string& key, const Act& x1, const Act* x2, int frames, int chunk, int cout,
                const float* bias1_s, const float* bias1_t, float eps, float teps, bool switch_mix);

// finalize-time weight fusion
void fuse_qkv(Ctx& c, const std::string& attn_key, cudaStream_t st);     // -> attn_key + ".to_qkv.weight"(.bias)
void fuse_geglu(Ctx& c, const std::string& ff_key, cudaStream_t st);     // -> ff_key + ".net.0.proj.geglu.*"

void unet_finalize(Ctx& c, cudaStream_t st);
void unet_prepare(Ctx& c, int T, cudaStream_t st);
void unet_set_clip_context(Ctx& c, const float* enc, cudaStream_t st);
// x16: [T][hw][8] 16-bit; v_out: fp32 [T][hw][4]
void unet_forward(Ctx& c, const void* x16, float timestep, const float ids[3], float* v_out);

std::string norm_prefix(const char* p);   // "unet2d" / "unet2d." -> "unet2d."
void unet2d_finalize(Ctx& c, cudaStream_t st);
// tokens fp32 [frames][len][cross_attention_dim] -> K | V of every cross-attention of network `prefix`
void unet2d_set_context(Ctx& c, const std::string& prefix, const float* tokens, int frames, int len);
// x16 / ctrl_x16: [F][hw][8] 16-bit (in_channels valid); out_tokens fp32 [F][hw][out_channels]
void unet2d_forward(Ctx& c, const std::string& prefix, const void* x16, int F, int h, int w, float timestep,
                    const std::string* ctrl_prefix, const void* ctrl_x16, float* out_tokens);

void vae_finalize(Ctx& c, cudaStream_t st);
// V = "vae." (SVD temporal-decoder VAE) or "vae2d." (AutoencoderKL of the StableNormal path)
// img16 [N][H][W][8] 16-bit (3 valid channels) -> lat fp32 NCHW [N][4][H/8][W/8] * out_scale
void vae_encode(Ctx& c, const std::string& V, const void* img16, int N, int H, int W, float out_scale,
                float* lat_nchw);
// 2-D decoder ("vae2d."): z16 [N][h*w][8] (4 valid, already / scaling) -> img fp32 NCHW [N][3][8h][8w] (nullable)
// and / or 8-bit unit normals [N][8h][8w][3] (nullable)
void vae2d_decode(Ctx& c, const void* z16, int N, int h, int w, float* img_nchw, unsigned char* normals_u8);
// z16 [T][h*w][8] (4 valid, already / scaling) -> img fp32 NCHW [T][3][8h][8w] (nullable) and / or
// frames fp32 [T][8h][8w][3] = clamp(img/2+0.5, 0, 1) (nullable; postprocess_video "np" layout)
void vae_decode(Ctx& c, const void* z16, int T, int h, int w, int chunk, float* img_nchw,
                float* frames_hwc = nullptr);

// CLIP image encoder (clip.cu): padded q|k|v / out_proj weights; video fp32 [F][3][H][W] -> enc fp32 [F][proj_dim]
void clip_finalize(Ctx& c, cudaStream_t st);
void clip_embed(Ctx& c, const float* video, int F, int H, int W, float* enc);

float sigmoidf_host(float x);

}  // namespace ug

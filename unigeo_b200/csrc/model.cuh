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

void vae_finalize(Ctx& c, cudaStream_t st);
// img16 [N][H][W][8] 16-bit (3 valid channels) -> lat fp32 NCHW [N][4][H/8][W/8]
void vae_encode(Ctx& c, const void* img16, int N, int H, int W, float* lat_nchw);
// z16 [T][h*w][8] (4 valid, already / scaling) -> img fp32 NCHW [T][3][8h][8w]
void vae_decode(Ctx& c, const void* z16, int T, int h, int w, int chunk, float* img_nchw);

float sigmoidf_host(float x);

}  // namespace ug

// AutoencoderKLTemporalDecoder graphs: 2-D encoder (latent_dist.mode())
// and the temporal decoder, run chunk by chunk like [UPSTREAM] decode_latents does.
// Replaces vae.encode(...) / vae.decode(..., num_frames) reached from
// /root/reference/model/depthcrafter.py:80-90.
#include "model.cuh"

namespace ug {

namespace {
const std::string V = "vae.";
const std::string V2 = "vae2d.";

// diffusers Attention(heads=1, norm_num_groups=32, residual_connection=True, bias=True)
Act mid_attention(Ctx& c, const std::string& key, const Act& x, int frames) {
  const int C = x.C, hw = x.H * x.W;
  const long long rows = (long long)frames * hw;
  Act out{c.alloc16(rows * C), C, x.H, x.W};
  const size_t m = c.ws.mark();
  void* n = c.alloc16(rows * C);
  op_gn(c, x.p, C, nullptr, 0, rows, hw, c.F(key + ".group_norm.weight"), c.F(key + ".group_norm.bias"),
        c.cfg.vae_eps, 0, n);
  void* qkv = c.alloc16(rows * 3 * C);
  { Epi e; e.out = qkv; e.ldc = 3 * C; e.bias = c.F(key + ".to_qkv.bias");
    op_linear(c, n, rows, C, C, c.M(key + ".to_qkv.weight"), 3 * C, e); }
  op_spatial_attention(c, qkv, frames, hw, C, C, n);
  { Epi e; e.out = out.p; e.ldc = C; e.bias = c.F(key + ".to_out.0.bias"); e.res = x.p; e.ldr = C;
    op_linear(c, n, rows, C, C, c.M(key + ".to_out.0.weight"), C, e); }
  c.ws.release(m);
  return out;
}
}  // namespace

void vae_finalize(Ctx& c, cudaStream_t st) {
  for (const std::string& P : {V, V2}) {
    for (const char* half : {"encoder", "decoder"}) {
      const std::string k = P + half + ".mid_block.attentions.0";
      if (c.has(k + ".to_q.weight") && !c.has(k + ".to_qkv.weight")) fuse_qkv(c, k, st);
    }
  }
  if (!c.vae) c.vae = new VaeModel();
}

void vae_encode(Ctx& c, const std::string& V, const void* img16, int N, int H, int W, float out_scale,
                float* lat_nchw) {
  const ug_model_cfg& g = c.cfg;
  const int nb = g.vae_num_blocks;
  const float eps = g.vae_eps;
  Act x{nullptr, g.vae_block_out[0], H, W};
  x.p = c.alloc16((long long)N * H * W * x.C);
  { Epi e; e.out = x.p; e.ldc = x.C; e.bias = c.F(V + "encoder.conv_in.bias");
    op_conv3x3(c, img16, N, H, W, 8, c.M(V + "encoder.conv_in.weight"), x.C, 1, 0, e); }
  for (int i = 0; i < nb; ++i) {
    const std::string b = V + "encoder.down_blocks." + std::to_string(i);
    for (int j = 0; j < g.vae_layers_per_block; ++j)
      x = resnet2d(c, b + ".resnets." + std::to_string(j), x, nullptr, N, g.vae_block_out[i], nullptr, eps);
    if (i < nb - 1) {
      Act d{c.alloc16((long long)N * (x.H / 2) * (x.W / 2) * x.C), x.C, x.H / 2, x.W / 2};
      Epi e; e.out = d.p; e.ldc = x.C; e.bias = c.F(b + ".downsamplers.0.conv.bias");
      op_conv3x3(c, x.p, N, x.H, x.W, x.C, c.M(b + ".downsamplers.0.conv.weight"), x.C, 2, 1, e);
      x = d;
    }
  }
  const int cm = g.vae_block_out[nb - 1];
  x = resnet2d(c, V + "encoder.mid_block.resnets.0", x, nullptr, N, cm, nullptr, eps);
  x = mid_attention(c, V + "encoder.mid_block.attentions.0", x, N);
  x = resnet2d(c, V + "encoder.mid_block.resnets.1", x, nullptr, N, cm, nullptr, eps);
  const long long rows = (long long)N * x.H * x.W;
  void* n = c.alloc16(rows * cm);
  op_gn(c, x.p, cm, nullptr, 0, rows, (long long)x.H * x.W, c.F(V + "encoder.conv_norm_out.weight"),
        c.F(V + "encoder.conv_norm_out.bias"), eps, 1, n);
  const int L2 = 2 * g.vae_latent_channels;
  UG_CHECK(L2 == 8, UG_ERR_INVALID, "encoder moments must have 8 channels");
  void* mom = c.alloc16(rows * L2);
  { Epi e; e.out = mom; e.ldc = L2; e.bias = c.F(V + "encoder.conv_out.bias");
    op_conv3x3(c, n, N, x.H, x.W, cm, c.M(V + "encoder.conv_out.weight"), L2, 1, 0, e); }
  // quant_conv 1x1; latent_dist.mode() = mean = first latent_channels outputs
  void* q = c.alloc16(rows * L2);
  { Epi e; e.out = q; e.ldc = L2; e.bias = c.F(V + "quant_conv.bias");
    op_linear(c, mom, rows, L2, L2, c.M(V + "quant_conv.weight"), L2, e); }
  if (!c.dry)
    op_check(c, launch_nhwc_to_nchw(q, N, x.H, x.W, L2, g.vae_latent_channels, out_scale, 0.f, 0, lat_nchw, c.fmt,
                                    c.stream),
             "nhwc_to_nchw");
}

void vae_decode(Ctx& c, const void* z16, int T, int h, int w, int chunk, float* img_nchw, float* frames_hwc) {
  const ug_model_cfg& g = c.cfg;
  const int nb = g.vae_num_blocks;
  const float eps = g.vae_eps, teps = g.vae_temporal_eps;
  const int cm = g.vae_block_out[nb - 1];
  const int H = h << (nb - 1), W = w << (nb - 1);
  for (int t0 = 0; t0 < T; t0 += chunk) {
    const int F = (T - t0) < chunk ? (T - t0) : chunk;
    const size_t m = c.ws.mark();
    auto st = [&](const std::string& key, const Act& a, int cout) {
      return st_resblock(c, key, a, nullptr, F, F, cout, nullptr, nullptr, eps, teps, true);
    };
    Act x{c.alloc16((long long)F * h * w * cm), cm, h, w};
    { Epi e; e.out = x.p; e.ldc = cm; e.bias = c.F(V + "decoder.conv_in.bias");
      op_conv3x3(c, reinterpret_cast<const char*>(z16) + (size_t)t0 * h * w * 8 * 2, F, h, w, 8,
                 c.M(V + "decoder.conv_in.weight"), cm, 1, 0, e); }
    x = st(V + "decoder.mid_block.resnets.0", x, cm);
    x = mid_attention(c, V + "decoder.mid_block.attentions.0", x, F);
    x = st(V + "decoder.mid_block.resnets.1", x, cm);
    for (int i = 0; i < nb; ++i) {
      const int co = g.vae_block_out[nb - 1 - i];
      const std::string b = V + "decoder.up_blocks." + std::to_string(i);
      for (int j = 0; j < g.vae_layers_per_block + 1; ++j) x = st(b + ".resnets." + std::to_string(j), x, co);
      if (i < nb - 1) {
        Act u{c.alloc16((long long)F * x.H * 2 * x.W * 2 * x.C), x.C, x.H * 2, x.W * 2};
        op_upsample2x(c, x.p, u.p, F, x.H, x.W, x.C);
        Act d{c.alloc16((long long)F * u.H * u.W * x.C), x.C, u.H, u.W};
        Epi e; e.out = d.p; e.ldc = x.C; e.bias = c.F(b + ".upsamplers.0.conv.bias");
        op_conv3x3(c, u.p, F, u.H, u.W, x.C, c.M(b + ".upsamplers.0.conv.weight"), x.C, 1, 0, e);
        x = d;
      }
    }
    const long long hw = (long long)x.H * x.W, rows = (long long)F * hw;
    void* n = c.alloc16(rows * x.C);
    op_gn(c, x.p, x.C, nullptr, 0, rows, hw, c.F(V + "decoder.conv_norm_out.weight"),
          c.F(V + "decoder.conv_norm_out.bias"), eps, 1, n);
    // conv_out (C -> 3) lands in an 8-channel buffer (zeroed) so time_conv_out can read it via TMA
    void* rgb = c.alloc16(rows * 8);
    void* rgb2 = c.alloc16(rows * 8);
    if (!c.dry) UG_CUDA(cudaMemsetAsync(rgb, 0, (size_t)rows * 16, c.stream));
    { Epi e; e.out = rgb; e.ldc = 8; e.bias = c.F(V + "decoder.conv_out.bias");
      op_conv3x3(c, n, F, x.H, x.W, x.C, c.M(V + "decoder.conv_out.weight"), g.vae_in_channels, 1, 0, e); }
    { Epi e; e.out = rgb2; e.ldc = 8; e.bias = c.F(V + "decoder.time_conv_out.bias");
      op_tconv3(c, rgb, F, hw, 8, c.M(V + "decoder.time_conv_out.weight"), g.vae_in_channels, F, e); }
    if (!c.dry && img_nchw)
      op_check(c, launch_nhwc_to_nchw(rgb2, F, H, W, 8, g.vae_in_channels, 1.f, 0.f, 0,
                                      img_nchw + (size_t)t0 * g.vae_in_channels * hw, c.fmt, c.stream),
               "nhwc_to_nchw");
    if (!c.dry && frames_hwc)
      op_check(c, launch_frames_out(rgb2, rows, frames_hwc + (size_t)t0 * 3 * hw, c.fmt, c.stream), "frames_out");
    c.ws.release(m);
  }
}

// AutoencoderKL.decode of the StableNormal path: post_quant_conv -> Decoder (mid Res-Attn-Res,
// 4 up blocks of 3 resnets, nearest x2 + conv) -- the temporal decoder above without its temporal half.
void vae2d_decode(Ctx& c, const void* z16, int N, int h, int w, float* img_nchw, unsigned char* normals_u8) {
  const ug_model_cfg& g = c.cfg;
  const int nb = g.vae_num_blocks;
  const float eps = g.vae_eps;
  const int cm = g.vae_block_out[nb - 1];
  const int H = h << (nb - 1), W = w << (nb - 1);
  const long long lrows = (long long)N * h * w;
  // post_quant_conv (1x1, 4 -> 4) into a zeroed 8-channel buffer so conv_in can read it via TMA
  void* pq = c.alloc16(lrows * 8);
  if (!c.dry) UG_CUDA(cudaMemsetAsync(pq, 0, (size_t)lrows * 16, c.stream));
  { Epi e; e.out = pq; e.ldc = 8; e.bias = c.F(V2 + "post_quant_conv.bias");
    op_linear(c, z16, lrows, 8, 8, c.M(V2 + "post_quant_conv.weight"), g.vae_latent_channels, e); }
  Act x{c.alloc16(lrows * cm), cm, h, w};
  { Epi e; e.out = x.p; e.ldc = cm; e.bias = c.F(V2 + "decoder.conv_in.bias");
    op_conv3x3(c, pq, N, h, w, 8, c.M(V2 + "decoder.conv_in.weight"), cm, 1, 0, e); }
  x = resnet2d(c, V2 + "decoder.mid_block.resnets.0", x, nullptr, N, cm, nullptr, eps);
  x = mid_attention(c, V2 + "decoder.mid_block.attentions.0", x, N);
  x = resnet2d(c, V2 + "decoder.mid_block.resnets.1", x, nullptr, N, cm, nullptr, eps);
  for (int i = 0; i < nb; ++i) {
    const int co = g.vae_block_out[nb - 1 - i];
    const std::string b = V2 + "decoder.up_blocks." + std::to_string(i);
    for (int j = 0; j < g.vae_layers_per_block + 1; ++j)
      x = resnet2d(c, b + ".resnets." + std::to_string(j), x, nullptr, N, co, nullptr, eps);
    if (i < nb - 1) {
      Act u{c.alloc16((long long)N * x.H * 2 * x.W * 2 * x.C), x.C, x.H * 2, x.W * 2};
      op_upsample2x(c, x.p, u.p, N, x.H, x.W, x.C);
      Act d{c.alloc16((long long)N * u.H * u.W * x.C), x.C, u.H, u.W};
      Epi e; e.out = d.p; e.ldc = x.C; e.bias = c.F(b + ".upsamplers.0.conv.bias");
      op_conv3x3(c, u.p, N, u.H, u.W, x.C, c.M(b + ".upsamplers.0.conv.weight"), x.C, 1, 0, e);
      x = d;
    }
  }
  const long long hw = (long long)x.H * x.W, rows = (long long)N * hw;
  void* n = c.alloc16(rows * x.C);
  op_gn(c, x.p, x.C, nullptr, 0, rows, hw, c.F(V2 + "decoder.conv_norm_out.weight"),
        c.F(V2 + "decoder.conv_norm_out.bias"), eps, 1, n);
  void* rgb = c.alloc16(rows * 8);
  { Epi e; e.out = rgb; e.ldc = 8; e.bias = c.F(V2 + "decoder.conv_out.bias");
    op_conv3x3(c, n, N, x.H, x.W, x.C, c.M(V2 + "decoder.conv_out.weight"), g.vae_in_channels, 1, 0, e); }
  if (c.dry) return;
  if (img_nchw)
    op_check(c, launch_nhwc_to_nchw(rgb, N, H, W, 8, g.vae_in_channels, 1.f, 0.f, 0, img_nchw, c.fmt, c.stream),
             "nhwc_to_nchw");
  if (normals_u8) op_check(c, launch_normals_to_u8(rgb, 8, rows, normals_u8, c.fmt, c.stream), "normals_to_u8");
}

}  // namespace ug

// 2-D conditional UNet + ControlNet graphs of the StableNormal path (SD-2.1 class topology,
// SURVEY.md App. A.5).  Replaces, per refinement step of the hub predictor the reference calls at
// /root/reference/model/stablenormal.py:39,
//   down, mid = controlnet(image_latent, t, encoder_hidden_states=prompt)          [UPSTREAM]
//   x0 = unet(latents, t, encoder_hidden_states=prompt, down_block_additional_residuals=down,
//             mid_block_additional_residual=mid)[0]
// Frames are the batch: every activation is [F][h*w][C] 16-bit channels-last, the same layout and the
// same kernels as the spatial half of the spatio-temporal UNet (unet.cu).  Differences from that
// graph: no temporal blocks, and a REAL cross-attention against the text tokens (77 keys), whose
// K | V depend only on the prompt and are computed once by ug_set_text_context.
// Fusions (exact): GEGLU / bias / residual in GEMM epilogues; the per-step time embedding of all
// resnets through one stacked GEMV; ControlNet's 1x1 "zero" convs write skip + residual directly.
#include <cmath>

#include "model.cuh"

namespace ug {

namespace {

struct Topo2D {
  std::vector<std::string> resnets;       // keys relative to the network prefix
  std::vector<int> resnet_cout;
  std::vector<std::string> transformers;
  std::vector<int> transformer_c;
};

Topo2D topo2d(const ug_unet2d_cfg& g, bool controlnet) {
  Topo2D t;
  const int nb = g.num_blocks, L = g.layers_per_block;
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < L; ++j) {
      const std::string b = "down_blocks." + std::to_string(i);
      t.resnets.push_back(b + ".resnets." + std::to_string(j));
      t.resnet_cout.push_back(g.block_out[i]);
      if (i < nb - 1) {
        t.transformers.push_back(b + ".attentions." + std::to_string(j));
        t.transformer_c.push_back(g.block_out[i]);
      }
    }
  t.resnets.push_back("mid_block.resnets.0"); t.resnet_cout.push_back(g.block_out[nb - 1]);
  t.resnets.push_back("mid_block.resnets.1"); t.resnet_cout.push_back(g.block_out[nb - 1]);
  t.transformers.push_back("mid_block.attentions.0"); t.transformer_c.push_back(g.block_out[nb - 1]);
  if (!controlnet)
    for (int i = 0; i < nb; ++i)
      for (int j = 0; j < L + 1; ++j) {
        const std::string b = "up_blocks." + std::to_string(i);
        const int co = g.block_out[nb - 1 - i];
        t.resnets.push_back(b + ".resnets." + std::to_string(j));
        t.resnet_cout.push_back(co);
        if (i > 0) {
          t.transformers.push_back(b + ".attentions." + std::to_string(j));
          t.transformer_c.push_back(co);
        }
      }
  return t;
}

void add_weight2d(Ctx& c, const std::string& key, void* p, bool f32, int cout, int cin) {
  Weight w;
  w.p = p; w.is_f32 = f32; w.taps = 1; w.cout = cout; w.cin = cin; w.cin_pad = cin;
  w.numel = (long long)cout * cin;
  c.weights[key] = w;
}

// attn2: to_k | to_v stacked -> one [2C][D] projection of the context
void fuse_kv(Ctx& c, const std::string& k, cudaStream_t st) {
  if (c.has(k + ".to_kv.weight")) return;
  const Weight& wk = c.W(k + ".to_k.weight");
  const Weight& wv = c.W(k + ".to_v.weight");
  UG_CHECK(wk.cin == wv.cin && wk.cout == wv.cout, UG_ERR_WEIGHT, "fuse_kv: to_k / to_v must agree: " + k);
  UG_CHECK(!c.has(k + ".to_k.bias"), UG_ERR_WEIGHT, "cross-attention projections with bias are not supported: " + k);
  const size_t bytes = (size_t)wk.cout * wk.cin * 2;
  char* dst = reinterpret_cast<char*>(c.dmalloc(2 * bytes));
  UG_CUDA(cudaMemcpyAsync(dst, wk.p, bytes, cudaMemcpyDeviceToDevice, st));
  UG_CUDA(cudaMemcpyAsync(dst + bytes, wv.p, bytes, cudaMemcpyDeviceToDevice, st));
  add_weight2d(c, k + ".to_kv.weight", dst, false, 2 * wk.cout, wk.cin);
}

void finalize_net(Ctx& c, Net2D& n, cudaStream_t st) {
  const ug_unet2d_cfg& g = c.cfg2d;
  const std::string& P = n.prefix;
  const Topo2D t = topo2d(g, n.controlnet);
  for (const std::string& k : t.transformers) {
    const std::string b = P + k + ".transformer_blocks.0";
    fuse_qkv(c, b + ".attn1", st);
    fuse_kv(c, b + ".attn2", st);
    fuse_geglu(c, b + ".ff", st);
  }
  // every resnet's time_emb_proj stacked into one GEMV, conv1.bias folded in
  const int E = g.block_out[0] * 4;
  int total = 0;
  for (int co : t.resnet_cout) total += co;
  char* Wall = reinterpret_cast<char*>(c.dmalloc((size_t)total * E * 2));
  std::vector<float> ball(total), tmp;
  int off = 0;
  UG_CUDA(cudaStreamSynchronize(st));
  for (size_t i = 0; i < t.resnets.size(); ++i) {
    const std::string rk = P + t.resnets[i];
    const int co = t.resnet_cout[i];
    const Weight& w = c.W(rk + ".time_emb_proj.weight");
    UG_CHECK(w.cout == co && w.cin == E, UG_ERR_WEIGHT, "time_emb_proj shape: " + rk);
    UG_CUDA(cudaMemcpy(Wall + (size_t)off * E * 2, w.p, (size_t)co * E * 2, cudaMemcpyDeviceToDevice));
    tmp.resize(co);
    UG_CUDA(cudaMemcpy(tmp.data(), c.F(rk + ".time_emb_proj.bias"), co * 4, cudaMemcpyDeviceToHost));
    for (int j = 0; j < co; ++j) ball[off + j] = tmp[j];
    UG_CUDA(cudaMemcpy(tmp.data(), c.F(rk + ".conv1.bias"), co * 4, cudaMemcpyDeviceToHost));
    for (int j = 0; j < co; ++j) ball[off + j] += tmp[j];
    n.temb_offset[rk] = off;
    off += co;
  }
  float* bdev = reinterpret_cast<float*>(c.dmalloc((size_t)total * 4));
  UG_CUDA(cudaMemcpy(bdev, ball.data(), (size_t)total * 4, cudaMemcpyHostToDevice));
  add_weight2d(c, P + "__temb_all.weight", Wall, false, total, E);
  add_weight2d(c, P + "__temb_all.bias", bdev, true, total, 1);
  n.temb_total = total;
  n.temb_out = reinterpret_cast<float*>(c.dmalloc((size_t)total * 4));
  n.scratch = reinterpret_cast<float*>(c.dmalloc((size_t)(4 * E + 4096) * 4));
}

// Transformer2DModel (linear projections, one BasicTransformerBlock)
Act transformer2d(Ctx& c, Net2D& n, const std::string& key, const Act& x, int F, int heads) {
  const ug_unet2d_cfg& g = c.cfg2d;
  const int C = x.C, hw = x.H * x.W;
  const long long rows = (long long)F * hw;
  UG_CHECK(C == heads * 64, UG_ERR_INVALID, "2-D UNet attention needs head_dim 64: " + key);
  Act out{c.alloc16(rows * C), C, x.H, x.W};
  const size_t mk = c.ws.mark();
  const std::string b = key + ".transformer_blocks.0";
  void* nrm = c.alloc16(rows * C);
  void* h = c.alloc16(rows * C);
  void* h1 = c.alloc16(rows * C);
  void* big = c.alloc16(rows * 4 * C);
  void* ao = c.alloc16(rows * C);

  op_gn(c, x.p, C, nullptr, 0, rows, hw, c.F(key + ".norm.weight"), c.F(key + ".norm.bias"), g.eps_transformer_norm, 0,
        nrm);
  { Epi e; e.out = h; e.ldc = C; e.bias = c.F(key + ".proj_in.bias");
    op_linear(c, nrm, rows, C, C, c.M(key + ".proj_in.weight"), C, e); }
  // self-attention
  op_layernorm(c, h, rows, C, c.F(b + ".norm1.weight"), c.F(b + ".norm1.bias"), g.ln_eps, nullptr, 1, nrm);
  { Epi e; e.out = big; e.ldc = 3 * C;
    op_linear(c, nrm, rows, C, C, c.M(b + ".attn1.to_qkv.weight"), 3 * C, e); }
  op_spatial_attention(c, big, F, hw, C, 64, ao);
  { Epi e; e.out = h1; e.ldc = C; e.bias = c.F(b + ".attn1.to_out.0.bias"); e.res = h; e.ldr = C;
    op_linear(c, ao, rows, C, C, c.M(b + ".attn1.to_out.0.weight"), C, e); }
  // cross-attention against the prompt tokens (K | V precomputed per context)
  op_layernorm(c, h1, rows, C, c.F(b + ".norm2.weight"), c.F(b + ".norm2.bias"), g.ln_eps, nullptr, 1, nrm);
  { Epi e; e.out = big; e.ldc = C;
    op_linear(c, nrm, rows, C, C, c.M(b + ".attn2.to_q.weight"), C, e); }
  if (!c.dry) {
    auto it = n.kv.find(key);
    UG_CHECK(it != n.kv.end() && n.ctx_len > 0, UG_ERR_STATE, "ug_set_text_context must precede the 2-D UNet forward");
    UG_CHECK(n.ctx_frames == 1 || n.ctx_frames == F, UG_ERR_INVALID, "text context frames must be 1 or F");
    op_cross_attention(c, big, C, it->second, ao, F, hw, C, n.ctx_len, n.ctx_frames > 1);
  }
  { Epi e; e.out = h; e.ldc = C; e.bias = c.F(b + ".attn2.to_out.0.bias"); e.res = h1; e.ldr = C;
    op_linear(c, ao, rows, C, C, c.M(b + ".attn2.to_out.0.weight"), C, e); }
  // feed-forward (GEGLU)
  op_layernorm(c, h, rows, C, c.F(b + ".norm3.weight"), c.F(b + ".norm3.bias"), g.ln_eps, nullptr, 1, nrm);
  { Epi e; e.out = big; e.ldc = 4 * C; e.bias = c.F(b + ".ff.net.0.proj.geglu.bias"); e.geglu = 1;
    op_linear(c, nrm, rows, C, C, c.M(b + ".ff.net.0.proj.geglu.weight"), 8 * C, e); }
  { Epi e; e.out = h1; e.ldc = C; e.bias = c.F(b + ".ff.net.2.bias"); e.res = h; e.ldr = C;
    op_linear(c, big, rows, 4 * C, 4 * C, c.M(b + ".ff.net.2.weight"), C, e); }
  { Epi e; e.out = out.p; e.ldc = C; e.bias = c.F(key + ".proj_out.bias"); e.res = x.p; e.ldr = C;
    op_linear(c, h1, rows, C, C, c.M(key + ".proj_out.weight"), C, e); }
  c.ws.release(mk);
  return out;
}

// time embedding -> conv1 bias vectors of every resnet of network n (one GEMV)
void time_embed(Ctx& c, Net2D& n, float timestep) {
  if (c.dry) return;
  const ug_unet2d_cfg& g = c.cfg2d;
  const std::string& P = n.prefix;
  const int C0 = g.block_out[0], E = 4 * C0;
  UG_CHECK(C0 <= 4096, UG_ERR_INVALID, "embedding dims");
  float* t_sin = n.scratch;            // [C0]
  float* t_h = t_sin + 4096;           // [E]
  float* emb = t_h + E;                // [E]
  op_check(c, launch_sinusoid_vals(timestep, 0, 0, 0, 1, C0, t_sin, c.stream), "sinusoid(t)");
  op_gemv(c, c.M(P + "time_embedding.linear_1.weight"), c.F(P + "time_embedding.linear_1.bias"), nullptr, t_sin, t_h,
          1, E, C0, 0, 1);
  op_gemv(c, c.M(P + "time_embedding.linear_2.weight"), c.F(P + "time_embedding.linear_2.bias"), nullptr, t_h, emb, 1,
          E, E, 0, 0);
  op_gemv(c, c.M(P + "__temb_all.weight"), c.F(P + "__temb_all.bias"), nullptr, emb, n.temb_out, 1, n.temb_total, E, 1,
          0);
}

// conv_in + down blocks + mid block; x16 [F][hw][8]
Act encoder_half(Ctx& c, Net2D& n, const void* x16, int F, int h, int w, std::vector<Act>& skips) {
  const ug_unet2d_cfg& g = c.cfg2d;
  const std::string& P = n.prefix;
  const int nb = g.num_blocks, L = g.layers_per_block;
  auto res = [&](const std::string& rel, const Act& a, const Act* b2, int cout) {
    const std::string key = P + rel;
    return resnet2d(c, key, a, b2, F, cout, n.temb_out + n.temb_offset.at(key), g.eps_resnet);
  };
  Act x{c.alloc16((long long)F * h * w * g.block_out[0]), g.block_out[0], h, w};
  { Epi e; e.out = x.p; e.ldc = x.C; e.bias = c.F(P + "conv_in.bias");
    op_conv3x3(c, x16, F, h, w, 8, c.M(P + "conv_in.weight"), x.C, 1, 0, e); }
  skips.push_back(x);
  for (int i = 0; i < nb; ++i) {
    const std::string b = "down_blocks." + std::to_string(i);
    for (int j = 0; j < L; ++j) {
      x = res(b + ".resnets." + std::to_string(j), x, nullptr, g.block_out[i]);
      if (i < nb - 1) x = transformer2d(c, n, P + b + ".attentions." + std::to_string(j), x, F, g.heads[i]);
      skips.push_back(x);
    }
    if (i < nb - 1) {
      UG_CHECK(x.H % 2 == 0 && x.W % 2 == 0, UG_ERR_INVALID, "latent size must be divisible by 8");
      Act d{c.alloc16((long long)F * (x.H / 2) * (x.W / 2) * x.C), x.C, x.H / 2, x.W / 2};
      Epi e; e.out = d.p; e.ldc = x.C; e.bias = c.F(P + b + ".downsamplers.0.conv.bias");
      op_conv3x3(c, x.p, F, x.H, x.W, x.C, c.M(P + b + ".downsamplers.0.conv.weight"), x.C, 2, 0, e);
      x = d;
      skips.push_back(x);
    }
  }
  const int cm = g.block_out[nb - 1];
  x = res("mid_block.resnets.0", x, nullptr, cm);
  x = transformer2d(c, n, P + "mid_block.attentions.0", x, F, g.heads[nb - 1]);
  x = res("mid_block.resnets.1", x, nullptr, cm);
  return x;
}

Net2D& find_net(Ctx& c, const std::string& prefix) {
  UG_CHECK(c.nets2d != nullptr, UG_ERR_STATE, "no 2-D network finalized (ug_ctx_set_unet2d_cfg + ug_ctx_finalize)");
  auto it = c.nets2d->nets.find(prefix);
  UG_CHECK(it != c.nets2d->nets.end(), UG_ERR_WEIGHT, "no 2-D network loaded under prefix '" + prefix + "'");
  return it->second;
}

}  // namespace

std::string norm_prefix(const char* p) {
  std::string s = p ? p : "";
  if (!s.empty() && s.back() != '.') s += '.';
  return s;
}

void unet2d_finalize(Ctx& c, cudaStream_t st) {
  if (c.cfg2d.num_blocks == 0) return;
  const std::string suffix = "mid_block.attentions.0.transformer_blocks.0.attn2.to_k.weight";
  std::vector<std::string> prefixes;
  for (const auto& kv : c.weights) {
    const std::string& k = kv.first;
    if (k.size() < suffix.size() || k.compare(k.size() - suffix.size(), suffix.size(), suffix) != 0) continue;
    const std::string P = k.substr(0, k.size() - suffix.size());
    if (c.has(P + "mid_block.attentions.0.temporal_transformer_blocks.0.attn1.to_q.weight")) continue;   // ST UNet
    prefixes.push_back(P);
  }
  if (prefixes.empty()) return;
  UG_CHECK(c.cfg2d.num_blocks >= 2 && c.cfg2d.num_blocks <= 4, UG_ERR_INVALID, "2-D UNet needs 2..4 blocks");
  if (!c.nets2d) c.nets2d = new Nets2D();
  for (const std::string& P : prefixes) {
    if (c.nets2d->nets.count(P)) continue;
    Net2D& n = c.nets2d->nets[P];
    n.prefix = P;
    n.controlnet = c.has(P + "controlnet_mid_block.weight");
    finalize_net(c, n, st);
  }
}

void unet2d_set_context(Ctx& c, const std::string& prefix, const float* tokens, int frames, int len) {
  Net2D& n = find_net(c, prefix);
  const ug_unet2d_cfg& g = c.cfg2d;
  const int D = g.cross_attention_dim;
  UG_CHECK(len >= 1 && len <= 128 && frames >= 1, UG_ERR_INVALID, "text context: 1 <= len <= 128 tokens");
  const long long rows = (long long)frames * len;
  const Topo2D t = topo2d(g, n.controlnet);
  const size_t mk = c.ws.mark();
  void* tok16 = c.alloc16(rows * D);
  if (!c.dry) op_check(c, launch_f32_to_tokens(tokens, D, D, rows, tok16, c.fmt, c.stream), "context tokens");
  const bool regrow = n.ctx_len * n.ctx_frames < rows;
  for (size_t i = 0; i < t.transformers.size(); ++i) {
    const std::string key = n.prefix + t.transformers[i];
    const int C = t.transformer_c[i];
    void* out = tok16;                       // size-only pass: any aligned address (op_linear sizes its split-K partials)
    if (!c.dry) {
      void*& buf = n.kv[key];
      if (buf == nullptr || regrow) buf = c.dmalloc((size_t)rows * 2 * C * 2);
      out = buf;
    }
    Epi e; e.out = out; e.ldc = 2 * C;
    op_linear(c, tok16, rows, D, D, c.M(key + ".transformer_blocks.0.attn2.to_kv.weight"), 2 * C, e);
  }
  if (!c.dry) {
    n.ctx_len = len;
    n.ctx_frames = frames;
  }
  c.ws.release(mk);
}

void unet2d_forward(Ctx& c, const std::string& prefix, const void* x16, int F, int h, int w, float timestep,
                    const std::string* ctrl_prefix, const void* ctrl_x16, float* out_tokens) {
  const ug_unet2d_cfg& g = c.cfg2d;
  UG_CHECK(g.num_blocks > 0, UG_ERR_STATE, "ug_ctx_set_unet2d_cfg was not called");
  Net2D& n = find_net(c, prefix);
  UG_CHECK(!n.controlnet, UG_ERR_INVALID, "'" + prefix + "' is a ControlNet, not a UNet");
  const std::string& P = n.prefix;
  const int nb = g.num_blocks, L = g.layers_per_block;

  time_embed(c, n, timestep);
  std::vector<Act> skips;
  Act x = encoder_half(c, n, x16, F, h, w, skips);

  if (ctrl_prefix) {
    Net2D& cn = find_net(c, *ctrl_prefix);
    UG_CHECK(cn.controlnet, UG_ERR_INVALID, "'" + *ctrl_prefix + "' is not a ControlNet");
    const std::string& Q = cn.prefix;
    time_embed(c, cn, timestep);
    std::vector<Act> cskips;
    Act cmid = encoder_half(c, cn, ctrl_x16, F, h, w, cskips);
    UG_CHECK(cskips.size() == skips.size(), UG_ERR_WEIGHT, "ControlNet / UNet topology mismatch");
    // skip_i <- skip_i + zero_conv_i(controlnet skip_i): the add rides in the 1x1 conv's epilogue
    for (size_t i = 0; i < skips.size(); ++i) {
      const Act& s = skips[i];
      const long long rows = (long long)F * s.H * s.W;
      const std::string zk = Q + "controlnet_down_blocks." + std::to_string(i);
      Act r{c.alloc16(rows * s.C), s.C, s.H, s.W};
      Epi e; e.out = r.p; e.ldc = s.C; e.bias = c.F(zk + ".bias"); e.res = s.p; e.ldr = s.C;
      op_linear(c, cskips[i].p, rows, s.C, s.C, c.M(zk + ".weight"), s.C, e);
      skips[i] = r;
    }
    const long long rows = (long long)F * x.H * x.W;
    Act r{c.alloc16(rows * x.C), x.C, x.H, x.W};
    Epi e; e.out = r.p; e.ldc = x.C; e.bias = c.F(Q + "controlnet_mid_block.bias"); e.res = x.p; e.ldr = x.C;
    op_linear(c, cmid.p, rows, x.C, x.C, c.M(Q + "controlnet_mid_block.weight"), x.C, e);
    x = r;
  }

  for (int i = 0; i < nb; ++i) {
    const int co = g.block_out[nb - 1 - i];
    const std::string b = "up_blocks." + std::to_string(i);
    for (int j = 0; j < L + 1; ++j) {
      Act sk = skips.back();
      skips.pop_back();
      const std::string key = P + b + ".resnets." + std::to_string(j);
      x = resnet2d(c, key, x, &sk, F, co, n.temb_out + n.temb_offset.at(key), g.eps_resnet);
      if (i > 0) x = transformer2d(c, n, P + b + ".attentions." + std::to_string(j), x, F, g.heads[nb - 1 - i]);
    }
    if (i < nb - 1) {
      Act u{c.alloc16((long long)F * x.H * 2 * x.W * 2 * x.C), x.C, x.H * 2, x.W * 2};
      op_upsample2x(c, x.p, u.p, F, x.H, x.W, x.C);
      Act d{c.alloc16((long long)F * u.H * u.W * x.C), x.C, u.H, u.W};
      Epi e; e.out = d.p; e.ldc = x.C; e.bias = c.F(P + b + ".upsamplers.0.conv.bias");
      op_conv3x3(c, u.p, F, u.H, u.W, x.C, c.M(P + b + ".upsamplers.0.conv.weight"), x.C, 1, 0, e);
      x = d;
    }
  }
  const long long rows = (long long)F * x.H * x.W;
  void* nrm = c.alloc16(rows * x.C);
  op_gn(c, x.p, x.C, nullptr, 0, rows, (long long)x.H * x.W, c.F(P + "conv_norm_out.weight"),
        c.F(P + "conv_norm_out.bias"), g.eps_resnet, 1, nrm);
  Epi e; e.out = out_tokens; e.ldc = g.out_channels; e.out_fp32 = 1; e.bias = c.F(P + "conv_out.bias");
  op_conv3x3(c, nrm, F, x.H, x.W, x.C, c.M(P + "conv_out.weight"), g.out_channels, 1, 0, e);
}

}  // namespace ug

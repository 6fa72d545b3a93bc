// Spatio-temporal UNet graph (SVD-XT topology with DepthCrafter's per-frame image embeddings).
// Follows SURVEY.md App. A.3; replaces the [UPSTREAM] call
//   unet(x_in, t, encoder_hidden_states=enc, added_time_ids=ids)[0]
// made once per denoising step from the pipeline the reference invokes at
// /root/reference/model/depthcrafter.py:80-90.
//
// Layout: every activation is [T][H*W][C] 16-bit, so the (B*T,HW,C) <-> (B*HW,T,C)
// permutes of the upstream code vanish: spatial ops see frames as the batch, temporal ops
// stride by H*W*C.  Fusions (all exact): GEGLU, bias, residual, AlphaBlender and the
// time-embedding add live in GEMM epilogues; both single-token cross-attentions collapse to
// per-frame / per-clip vectors computed once per clip (ug_set_clip_context).
#include <cmath>

#include "model.cuh"

namespace ug {

float sigmoidf_host(float x) { return 1.0f / (1.0f + std::exp(-x)); }

namespace {

const std::string U = "unet.";

struct Topo {   // enumerates block keys of the UNet in execution order
  std::vector<std::string> resnets;        // SpatioTemporalResBlock keys
  std::vector<int> resnet_cout;
  std::vector<std::string> transformers;   // TransformerSpatioTemporalModel keys
  std::vector<int> transformer_c;
};

Topo unet_topo(const ug_model_cfg& g) {
  Topo t;
  const int nb = g.unet_num_blocks, L = g.unet_layers_per_block;
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < L; ++j) {
      const std::string b = U + "down_blocks." + std::to_string(i);
      t.resnets.push_back(b + ".resnets." + std::to_string(j));
      t.resnet_cout.push_back(g.unet_block_out[i]);
      if (i < nb - 1) {
        t.transformers.push_back(b + ".attentions." + std::to_string(j));
        t.transformer_c.push_back(g.unet_block_out[i]);
      }
    }
  t.resnets.push_back(U + "mid_block.resnets.0"); t.resnet_cout.push_back(g.unet_block_out[nb - 1]);
  t.resnets.push_back(U + "mid_block.resnets.1"); t.resnet_cout.push_back(g.unet_block_out[nb - 1]);
  t.transformers.push_back(U + "mid_block.attentions.0"); t.transformer_c.push_back(g.unet_block_out[nb - 1]);
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < L + 1; ++j) {
      const std::string b = U + "up_blocks." + std::to_string(i);
      const int co = g.unet_block_out[nb - 1 - i];
      t.resnets.push_back(b + ".resnets." + std::to_string(j));
      t.resnet_cout.push_back(co);
      if (i > 0) {
        t.transformers.push_back(b + ".attentions." + std::to_string(j));
        t.transformer_c.push_back(co);
      }
    }
  return t;
}

void add_weight(Ctx& c, const std::string& key, void* p, bool f32, int taps, int cout, int cin) {
  Weight w;
  w.p = p; w.is_f32 = f32; w.taps = taps; w.cout = cout; w.cin = cin; w.cin_pad = cin;
  w.numel = (long long)taps * cout * cin;
  c.weights[key] = w;
}

}  // namespace

// ------------------------------------------------------------------ weight fusion
void fuse_qkv(Ctx& c, const std::string& k, cudaStream_t st) {
  if (c.has(k + ".to_qkv.weight")) return;       // ug_ctx_finalize may run again after more weights were loaded
  const Weight& q = c.W(k + ".to_q.weight");
  const Weight& kk = c.W(k + ".to_k.weight");
  const Weight& v = c.W(k + ".to_v.weight");
  UG_CHECK(q.cin == kk.cin && q.cin == v.cin && q.cout == kk.cout && q.cout == v.cout, UG_ERR_WEIGHT,
           "fuse_qkv: self-attention projections must agree: " + k);
  const size_t bytes = (size_t)q.cout * q.cin * 2;
  char* dst = reinterpret_cast<char*>(c.dmalloc(3 * bytes));
  UG_CUDA(cudaMemcpyAsync(dst, q.p, bytes, cudaMemcpyDeviceToDevice, st));
  UG_CUDA(cudaMemcpyAsync(dst + bytes, kk.p, bytes, cudaMemcpyDeviceToDevice, st));
  UG_CUDA(cudaMemcpyAsync(dst + 2 * bytes, v.p, bytes, cudaMemcpyDeviceToDevice, st));
  add_weight(c, k + ".to_qkv.weight", dst, false, 1, 3 * q.cout, q.cin);
  if (c.has(k + ".to_q.bias")) {
    const size_t bb = (size_t)q.cout * 4;
    char* bd = reinterpret_cast<char*>(c.dmalloc(3 * bb));
    UG_CUDA(cudaMemcpyAsync(bd, c.F(k + ".to_q.bias"), bb, cudaMemcpyDeviceToDevice, st));
    UG_CUDA(cudaMemcpyAsync(bd + bb, c.F(k + ".to_k.bias"), bb, cudaMemcpyDeviceToDevice, st));
    UG_CUDA(cudaMemcpyAsync(bd + 2 * bb, c.F(k + ".to_v.bias"), bb, cudaMemcpyDeviceToDevice, st));
    add_weight(c, k + ".to_qkv.bias", bd, true, 1, 3 * q.cout, 1);
  }
}

void fuse_geglu(Ctx& c, const std::string& k, cudaStream_t st) {
  if (c.has(k + ".net.0.proj.geglu.weight")) return;
  const Weight& w = c.W(k + ".net.0.proj.weight");
  const int n = w.cout, K = w.cin, half = n / 2;
  UG_CHECK(half % 128 == 0, UG_ERR_WEIGHT, "fuse_geglu: inner dim must be a multiple of 128: " + k);
  // tile t of 256 output columns = [128 value rows t*128.. | 128 gate rows half + t*128..]
  char* dst = reinterpret_cast<char*>(c.dmalloc((size_t)n * K * 2));
  const char* src = reinterpret_cast<const char*>(w.p);
  const size_t blk = (size_t)128 * K * 2;
  UG_CUDA(cudaMemcpy2DAsync(dst, 2 * blk, src, blk, blk, half / 128, cudaMemcpyDeviceToDevice, st));
  UG_CUDA(cudaMemcpy2DAsync(dst + blk, 2 * blk, src + (size_t)half * K * 2, blk, blk, half / 128,
                            cudaMemcpyDeviceToDevice, st));
  add_weight(c, k + ".net.0.proj.geglu.weight", dst, false, 1, n, K);
  const char* bs = reinterpret_cast<const char*>(c.F(k + ".net.0.proj.bias"));
  char* bd = reinterpret_cast<char*>(c.dmalloc((size_t)n * 4));
  UG_CUDA(cudaMemcpy2DAsync(bd, 1024, bs, 512, 512, half / 128, cudaMemcpyDeviceToDevice, st));
  UG_CUDA(cudaMemcpy2DAsync(bd + 512, 1024, bs + (size_t)half * 4, 512, 512, half / 128, cudaMemcpyDeviceToDevice,
                            st));
  add_weight(c, k + ".net.0.proj.geglu.bias", bd, true, 1, n, 1);
}

// LayerNorm `norm` followed by the linear layer `lin` (weight key lin + ".weight", optional lin + ".bias") ->
// lin + ".lnf.weight" = gamma (.) W, ".lnf.colsum", ".lnf.bias" = bias + W beta: what a tapgemm launch with
// TapGemmArgs::ln_stat consumes (the GEMM then reads the un-normalised rows)
void fold_layernorm(Ctx& c, const std::string& norm, const std::string& lin, cudaStream_t st) {
  if (c.has(lin + ".lnf.weight")) return;
  const Weight& w = c.W(lin + ".weight");
  UG_CHECK(w.taps == 1 && !w.is_f32 && w.cin_pad == w.cin, UG_ERR_WEIGHT, "fold_layernorm: plain matrix expected: " + lin);
  const int N = w.cout, K = w.cin;
  void* wf = c.dmalloc((size_t)N * K * 2);
  float* cs = reinterpret_cast<float*>(c.dmalloc((size_t)N * 4));
  float* bo = reinterpret_cast<float*>(c.dmalloc((size_t)N * 4));
  const float* bias = c.has(lin + ".bias") ? c.F(lin + ".bias") : nullptr;
  const int r = launch_ln_fold_weights(w.p, c.F(norm + ".weight"), c.F(norm + ".bias"), bias, wf, cs, bo, N, K, c.fmt, st);
  UG_CHECK(r == 0, UG_ERR_CUDA, "ln_fold_weights: " + lin);
  add_weight(c, lin + ".lnf.weight", wf, false, 1, N, K);
  add_weight(c, lin + ".lnf.colsum", cs, true, 1, N, 1);
  add_weight(c, lin + ".lnf.bias", bo, true, 1, N, 1);
}

void unet_finalize(Ctx& c, cudaStream_t st) {
  if (!c.has(U + "conv_in.weight")) return;   // VAE-only context
  if (!c.unet) c.unet = new UNetModel();
  UNetModel& m = *c.unet;
  const Topo t = unet_topo(c.cfg);
  for (const std::string& k : t.transformers) {
    fuse_qkv(c, k + ".transformer_blocks.0.attn1", st);
    fuse_qkv(c, k + ".temporal_transformer_blocks.0.attn1", st);
    fuse_geglu(c, k + ".transformer_blocks.0.ff", st);
    fuse_geglu(c, k + ".temporal_transformer_blocks.0.ff_in", st);
    fuse_geglu(c, k + ".temporal_transformer_blocks.0.ff", st);
  }
  // LayerNorms folded into the GEMMs that consume them: opt-in (UG_LN_FOLD=1, read at every finalize).  Measured at cfg2
  // (profiles/r02_ln_fold.txt): 250 fewer launches and 10 fewer activation passes per transformer, but the consumers'
  // epilogues are instruction-issue bound at K = 320 and the two extra packed FMAs per column pair cost what the
  // LayerNorm launches did -- 28.4 ms per step either way.
  { const char* e = getenv("UG_LN_FOLD"); m.ln_fold = e != nullptr && atoi(e) != 0; }
  if (m.ln_fold)
    for (const std::string& k : t.transformers) {
      const std::string sb = k + ".transformer_blocks.0", tb = k + ".temporal_transformer_blocks.0";
      fold_layernorm(c, sb + ".norm1", sb + ".attn1.to_qkv", st);
      fold_layernorm(c, sb + ".norm3", sb + ".ff.net.0.proj.geglu", st);
      fold_layernorm(c, tb + ".norm_in", tb + ".ff_in.net.0.proj.geglu", st);
      fold_layernorm(c, tb + ".norm1", tb + ".attn1.to_qkv", st);
      fold_layernorm(c, tb + ".norm3", tb + ".ff.net.0.proj.geglu", st);
    }
  if (c.has(U + "__temb_all.weight")) return;
  // stack every time_emb_proj (spatial + temporal resnets) into one GEMV; fold conv1.bias in
  const int E = c.cfg.unet_block_out[0] * 4;
  int total = 0;
  for (size_t i = 0; i < t.resnets.size(); ++i) total += 2 * t.resnet_cout[i];
  char* Wall = reinterpret_cast<char*>(c.dmalloc((size_t)total * E * 2));
  std::vector<float> ball(total), tmp;
  int off = 0;
  UG_CUDA(cudaStreamSynchronize(st));
  for (size_t i = 0; i < t.resnets.size(); ++i) {
    for (int part = 0; part < 2; ++part) {
      const std::string rk = t.resnets[i] + (part == 0 ? ".spatial_res_block" : ".temporal_res_block");
      const int co = t.resnet_cout[i];
      const Weight& w = c.W(rk + ".time_emb_proj.weight");
      UG_CHECK(w.cout == co && w.cin == E, UG_ERR_WEIGHT, "time_emb_proj shape: " + rk);
      UG_CUDA(cudaMemcpy(Wall + (size_t)off * E * 2, w.p, (size_t)co * E * 2, cudaMemcpyDeviceToDevice));
      tmp.resize(co);
      UG_CUDA(cudaMemcpy(tmp.data(), c.F(rk + ".time_emb_proj.bias"), co * 4, cudaMemcpyDeviceToHost));
      for (int j = 0; j < co; ++j) ball[off + j] = tmp[j];
      UG_CUDA(cudaMemcpy(tmp.data(), c.F(rk + ".conv1.bias"), co * 4, cudaMemcpyDeviceToHost));
      for (int j = 0; j < co; ++j) ball[off + j] += tmp[j];
      m.temb_offset[rk] = off;
      off += co;
    }
  }
  float* bdev = reinterpret_cast<float*>(c.dmalloc((size_t)total * 4));
  UG_CUDA(cudaMemcpy(bdev, ball.data(), (size_t)total * 4, cudaMemcpyHostToDevice));
  add_weight(c, U + "__temb_all.weight", Wall, false, 1, total, E);
  add_weight(c, U + "__temb_all.bias", bdev, true, 1, total, 1);
  m.temb_total = total;
  m.temb_out = reinterpret_cast<float*>(c.dmalloc((size_t)total * 4));
  m.scratch = reinterpret_cast<float*>(c.dmalloc((size_t)(16 * E + 4096) * 4));
}

// shape-only constants: frame positional embeddings of every transformer
void unet_prepare(Ctx& c, int T, cudaStream_t st) {
  if (!c.unet) return;
  UNetModel& m = *c.unet;
  if (m.prepared_T == T) return;
  const Topo t = unet_topo(c.cfg);
  c.stream = st;
  int cmax = 0;
  for (int cc : t.transformer_c) cmax = cc > cmax ? cc : cmax;
  float* idx = reinterpret_cast<float*>(c.dmalloc((size_t)T * 4));
  float* sinb = reinterpret_cast<float*>(c.dmalloc((size_t)T * cmax * 4));
  float* hid = reinterpret_cast<float*>(c.dmalloc((size_t)T * cmax * 4 * 4));
  op_check(c, launch_iota(idx, T, st), "iota");
  for (size_t i = 0; i < t.transformers.size(); ++i) {
    const std::string& k = t.transformers[i];
    const int C = t.transformer_c[i];
    float* tp = reinterpret_cast<float*>(c.dmalloc((size_t)T * C * 4));
    float* a2s = reinterpret_cast<float*>(c.dmalloc((size_t)T * C * 4));
    float* a2t = reinterpret_cast<float*>(c.dmalloc((size_t)C * 4));
    op_check(c, launch_sinusoid(idx, T, C, sinb, st), "sinusoid");
    op_gemv(c, c.M(k + ".time_pos_embed.linear_1.weight"), c.F(k + ".time_pos_embed.linear_1.bias"), nullptr,
            sinb, hid, T, 4 * C, C, 0, 1);
    op_gemv(c, c.M(k + ".time_pos_embed.linear_2.weight"), c.F(k + ".time_pos_embed.linear_2.bias"), nullptr,
            hid, tp, T, C, 4 * C, 0, 0);
    m.time_pos[k] = tp;
    {
      // LayerNorm fold: the spatial block stores h + emb_t, and the AlphaBlender input alpha * h = alpha * (h + emb_t) -
      // alpha * emb_t gets the correction as a per-frame bias of the last temporal GEMM: (1 - alpha) * (-alpha / (1 - alpha)) emb_t
      const float alpha = sigmoidf_host(c.W(k + ".time_mixer.mix_factor").host.at(0));
      float* tpb = reinterpret_cast<float*>(c.dmalloc((size_t)T * C * 4));
      const float a = (1.0f - alpha) > 1e-4f ? -alpha / (1.0f - alpha) : 0.0f;
      op_check(c, launch_scale_f32(tp, a, tpb, (long long)T * C, st), "scale");
      m.time_pos_blend[k] = tpb;
    }
    m.attn2_spatial[k] = a2s;
    m.attn2_temporal[k] = a2t;
  }
  m.prepared_T = T;
  m.clip_context_set = false;
}

// attn2(.) with ONE context token: softmax over a single key is 1, so the block output is
// W_o (W_v ctx) + b_o -- per frame for the spatial block, first-frame for the temporal one.
void unet_set_clip_context(Ctx& c, const float* enc, cudaStream_t st) {
  UG_CHECK(c.unet && c.unet->prepared_T > 0, UG_ERR_STATE, "ug_ctx_prepare must precede ug_set_clip_context");
  UNetModel& m = *c.unet;
  const Topo t = unet_topo(c.cfg);
  const int T = m.prepared_T, D = c.cfg.cross_attention_dim;
  c.stream = st;
  c.ws.off = 0;
  for (size_t i = 0; i < t.transformers.size(); ++i) {
    const std::string& k = t.transformers[i];
    const int C = t.transformer_c[i];
    const size_t mk = c.ws.mark();
    float* tmp = c.allocf((long long)T * C);
    const std::string s = k + ".transformer_blocks.0.attn2";
    op_gemv(c, c.M(s + ".to_v.weight"), nullptr, nullptr, enc, tmp, T, C, D, 0, 0);
    op_gemv(c, c.M(s + ".to_out.0.weight"), c.F(s + ".to_out.0.bias"), nullptr, tmp, m.attn2_spatial[k], T, C, C,
            0, 0);
    const std::string tt = k + ".temporal_transformer_blocks.0.attn2";
    op_gemv(c, c.M(tt + ".to_v.weight"), nullptr, nullptr, enc, tmp, 1, C, D, 0, 0);
    op_gemv(c, c.M(tt + ".to_out.0.weight"), c.F(tt + ".to_out.0.bias"), nullptr, tmp, m.attn2_temporal[k], 1, C,
            C, 0, 0);
    c.ws.release(mk);
  }
  m.clip_context_set = true;
}

// ------------------------------------------------------------------ blocks
Act resnet2d(Ctx& c, const std::string& key, const Act& x1, const Act* x2, int frames, int cout,
             const float* bias1, float eps) {
  const long long rows = (long long)frames * x1.H * x1.W;
  const long long hw = (long long)x1.H * x1.W;
  const int cin = x1.C + (x2 ? x2->C : 0);
  Act out{c.alloc16(rows * cout), cout, x1.H, x1.W};
  const size_t m = c.ws.mark();
  const void* xin = x1.p;
  if (x2) {
    void* xc = c.alloc16(rows * cin);
    op_concat(c, x1.p, x1.C, x2->p, x2->C, rows, xc);
    xin = xc;
  }
  void* h = c.alloc16(rows * cin);
  op_gn(c, xin, cin, nullptr, 0, rows, hw, c.F(key + ".norm1.weight"), c.F(key + ".norm1.bias"), eps, 1, h);
  void* h1 = c.alloc16(rows * cout);
  Epi e1;
  e1.out = h1; e1.ldc = cout;
  e1.bias = bias1 ? bias1 : c.F(key + ".conv1.bias");
  op_conv3x3(c, h, frames, x1.H, x1.W, cin, c.M(key + ".conv1.weight"), cout, 1, 0, e1);
  void* h2 = c.alloc16(rows * cout);
  op_gn(c, h1, cout, nullptr, 0, rows, hw, c.F(key + ".norm2.weight"), c.F(key + ".norm2.bias"), eps, 1, h2);
  const void* sc = xin;
  if (c.has(key + ".conv_shortcut.weight")) {
    void* s = c.alloc16(rows * cout);
    Epi es;
    es.out = s; es.ldc = cout; es.bias = c.F(key + ".conv_shortcut.bias");
    op_linear(c, xin, rows, cin, cin, c.M(key + ".conv_shortcut.weight"), cout, es);
    sc = s;
  } else {
    UG_CHECK(cin == cout, UG_ERR_WEIGHT, "resnet without conv_shortcut needs cin == cout: " + key);
  }
  Epi e2;
  e2.out = out.p; e2.ldc = cout; e2.bias = c.F(key + ".conv2.bias");
  e2.res = sc; e2.ldr = cout;
  op_conv3x3(c, h2, frames, x1.H, x1.W, cout, c.M(key + ".conv2.weight"), cout, 1, 0, e2);
  c.ws.release(m);
  return out;
}

Act st_resblock(Ctx& c, const std::string& key, const Act& x1, const Act* x2, int frames, int /*chunk*/,
                int cout, const float* bias1_s, const float* bias1_t, float eps, float teps, bool switch_mix) {
  const long long hw = (long long)x1.H * x1.W;
  const long long rows = (long long)frames * hw;
  Act out{c.alloc16(rows * cout), cout, x1.H, x1.W};
  const size_t m = c.ws.mark();
  Act s = resnet2d(c, key + ".spatial_res_block", x1, x2, frames, cout, bias1_s, eps);
  const std::string tk = key + ".temporal_res_block";
  void* g1 = c.alloc16(rows * cout);
  op_gn(c, s.p, cout, nullptr, 0, rows, rows, c.F(tk + ".norm1.weight"), c.F(tk + ".norm1.bias"), teps, 1, g1);
  void* t1 = c.alloc16(rows * cout);
  Epi e1;
  e1.out = t1; e1.ldc = cout;
  e1.bias = bias1_t ? bias1_t : c.F(tk + ".conv1.bias");
  op_tconv3(c, g1, frames, hw, cout, c.M(tk + ".conv1.weight"), cout, frames, e1);
  op_gn(c, t1, cout, nullptr, 0, rows, rows, c.F(tk + ".norm2.weight"), c.F(tk + ".norm2.bias"), teps, 1, g1);
  float alpha = sigmoidf_host(c.W(key + ".time_mixer.mix_factor").host.at(0));
  if (switch_mix) alpha = 1.0f - alpha;
  Epi e2;
  e2.out = out.p; e2.ldc = cout; e2.bias = c.F(tk + ".conv2.bias");
  e2.res = s.p; e2.ldr = cout;            // temporal resnet: x + h
  e2.blend = s.p; e2.ldb = cout;          // AlphaBlender: alpha * x_spatial + (1 - alpha) * x_temporal
  e2.alpha = alpha;
  op_tconv3(c, g1, frames, hw, cout, c.M(tk + ".conv2.weight"), cout, frames, e2);
  c.ws.release(m);
  return out;
}

namespace {

Act st_transformer(Ctx& c, const std::string& key, const Act& x, int T, int heads) {
  UNetModel& m = *c.unet;
  const int C = x.C;
  const int hw = x.H * x.W;
  const long long rows = (long long)T * hw;
  const float ln_eps = c.cfg.ln_eps;
  Act out{c.alloc16(rows * C), C, x.H, x.W};
  const size_t mk = c.ws.mark();
  const std::string sb = key + ".transformer_blocks.0";
  const std::string tb = key + ".temporal_transformer_blocks.0";
  void* n = c.alloc16(rows * C);           // normalised scratch (reused)
  void* h = c.alloc16(rows * C);
  void* h1 = c.alloc16(rows * C);
  void* big = c.alloc16(rows * 4 * C);     // qkv (3C) and GEGLU (4C) scratch
  void* ao = c.alloc16(rows * C);

  op_gn(c, x.p, C, nullptr, 0, rows, hw, c.F(key + ".norm.weight"), c.F(key + ".norm.bias"),
        c.cfg.eps_transformer_norm, 0, n);

  const float alpha_mix = sigmoidf_host(c.W(key + ".time_mixer.mix_factor").host.at(0));
  if (m.ln_fold && (1.0f - alpha_mix) > 1e-4f) {
    // ---- every LayerNorm folded into the GEMM behind it: the GEMM reads the raw rows with gamma-scaled weights and
    // its epilogue applies rstd * (acc - mean * colsum) + (bias + W beta); the row statistics come out of the epilogue
    // of the GEMM that PRODUCED the rows (or one row_stats pass where that launch cannot provide them).  No normalised
    // tensor is ever written: 5 LayerNorm launches and 10 activation passes per transformer disappear.
    const int cap = 4;     // partials a consumer keeps in registers one tile ahead; wider producers run a row_stats pass
    float2* stat[2] = {reinterpret_cast<float2*>(c.allocf(rows * cap * 2)), reinterpret_cast<float2*>(c.allocf(rows * cap * 2))};
    int parts = 0, si = 0;
    auto produce = [&](Epi& e) {
      e.stat_out = stat[si]; e.stat_parts = &parts; e.stat_cap = cap; e.stat_eps = ln_eps;
    };
    auto consume = [&](Epi& e, const std::string& lin) {
      e.ln_stat = stat[si]; e.ln_parts = parts; e.ln_inv_c = 1.0f / (float)C; e.ln_eps = ln_eps;
      e.ln_colsum = c.F(lin + ".lnf.colsum"); e.bias = c.F(lin + ".lnf.bias");
      si ^= 1;
    };
    const float* tp = m.time_pos.at(key);
    { Epi e; e.out = h; e.ldc = C; e.bias = c.F(key + ".proj_in.bias"); produce(e);
      op_linear(c, n, rows, C, C, c.M(key + ".proj_in.weight"), C, e); }
    // spatial block: norm1 -> qkv
    { Epi e; e.out = big; e.ldc = 3 * C; consume(e, sb + ".attn1.to_qkv");
      op_linear(c, h, rows, C, C, c.M(sb + ".attn1.to_qkv.lnf.weight"), 3 * C, e); }
    op_spatial_attention(c, big, T, hw, C, C / heads, ao);
    { Epi e; e.out = h1; e.ldc = C; e.bias = c.F(sb + ".attn1.to_out.0.bias");
      e.res = h; e.ldr = C;
      e.fbias = m.attn2_spatial.at(key); e.fbias_ld = C; e.fbias_div = hw;     // + attn2 (collapsed)
      produce(e);
      op_linear(c, ao, rows, C, C, c.M(sb + ".attn1.to_out.0.weight"), C, e); }
    // norm3 -> GEGLU
    { Epi e; e.out = big; e.ldc = 4 * C; e.geglu = 1; consume(e, sb + ".ff.net.0.proj.geglu");
      op_linear(c, h1, rows, C, C, c.M(sb + ".ff.net.0.proj.geglu.lnf.weight"), 8 * C, e); }
    // hp = spatial block output + emb_t[frame]: the temporal block's input stream
    { Epi e; e.out = h; e.ldc = C; e.bias = c.F(sb + ".ff.net.2.bias"); e.res = h1; e.ldr = C;
      e.fbias = tp; e.fbias_ld = C; e.fbias_div = hw;
      produce(e);
      op_linear(c, big, rows, 4 * C, 4 * C, c.M(sb + ".ff.net.2.weight"), C, e); }
    // temporal block: norm_in -> ff_in (GEGLU) -> + hp
    { Epi e; e.out = big; e.ldc = 4 * C; e.geglu = 1; consume(e, tb + ".ff_in.net.0.proj.geglu");
      op_linear(c, h, rows, C, C, c.M(tb + ".ff_in.net.0.proj.geglu.lnf.weight"), 8 * C, e); }
    { Epi e; e.out = h1; e.ldc = C; e.bias = c.F(tb + ".ff_in.net.2.bias"); e.res = h; e.ldr = C;
      produce(e);
      op_linear(c, big, rows, 4 * C, 4 * C, c.M(tb + ".ff_in.net.2.weight"), C, e); }
    // norm1 -> temporal qkv
    { Epi e; e.out = big; e.ldc = 3 * C; consume(e, tb + ".attn1.to_qkv");
      op_linear(c, h1, rows, C, C, c.M(tb + ".attn1.to_qkv.lnf.weight"), 3 * C, e); }
    op_temporal_attention(c, big, ao, T, hw, C);
    void* h2 = n;   // the GroupNorm output is dead after proj_in
    { Epi e; e.out = h2; e.ldc = C; e.bias = c.F(tb + ".attn1.to_out.0.bias"); e.res = h1; e.ldr = C;
      e.fbias = m.attn2_temporal.at(key); e.fbias_ld = C; e.fbias_div = (int)rows;  // + attn2 (one vector)
      produce(e);
      op_linear(c, ao, rows, C, C, c.M(tb + ".attn1.to_out.0.weight"), C, e); }
    // norm3 -> GEGLU -> ff.net.2 + h2, mixed with x_spatial = hp - emb_t
    { Epi e; e.out = big; e.ldc = 4 * C; e.geglu = 1; consume(e, tb + ".ff.net.0.proj.geglu");
      op_linear(c, h2, rows, C, C, c.M(tb + ".ff.net.0.proj.geglu.lnf.weight"), 8 * C, e); }
    { Epi e; e.out = h1; e.ldc = C; e.bias = c.F(tb + ".ff.net.2.bias"); e.res = h2; e.ldr = C;
      e.blend = h; e.ldb = C; e.alpha = alpha_mix;
      e.fbias = m.time_pos_blend.at(key); e.fbias_ld = C; e.fbias_div = hw;
      op_linear(c, big, rows, 4 * C, 4 * C, c.M(tb + ".ff.net.2.weight"), C, e); }
    { Epi e; e.out = out.p; e.ldc = C; e.bias = c.F(key + ".proj_out.bias"); e.res = x.p; e.ldr = C;
      op_linear(c, h1, rows, C, C, c.M(key + ".proj_out.weight"), C, e); }
    c.ws.release(mk);
    return out;
  }

  { Epi e; e.out = h; e.ldc = C; e.bias = c.F(key + ".proj_in.bias");
    op_linear(c, n, rows, C, C, c.M(key + ".proj_in.weight"), C, e); }

  // ---- spatial BasicTransformerBlock
  op_layernorm(c, h, rows, C, c.F(sb + ".norm1.weight"), c.F(sb + ".norm1.bias"), ln_eps, nullptr, 1, n);
  { Epi e; e.out = big; e.ldc = 3 * C;
    op_linear(c, n, rows, C, C, c.M(sb + ".attn1.to_qkv.weight"), 3 * C, e); }
  op_spatial_attention(c, big, T, hw, C, C / heads, ao);
  { Epi e; e.out = h1; e.ldc = C; e.bias = c.F(sb + ".attn1.to_out.0.bias");
    e.res = h; e.ldr = C;
    e.fbias = m.attn2_spatial.at(key); e.fbias_ld = C; e.fbias_div = hw;     // + attn2 (collapsed)
    op_linear(c, ao, rows, C, C, c.M(sb + ".attn1.to_out.0.weight"), C, e); }
  op_layernorm(c, h1, rows, C, c.F(sb + ".norm3.weight"), c.F(sb + ".norm3.bias"), ln_eps, nullptr, 1, n);
  { Epi e; e.out = big; e.ldc = 4 * C; e.bias = c.F(sb + ".ff.net.0.proj.geglu.bias"); e.geglu = 1;
    op_linear(c, n, rows, C, C, c.M(sb + ".ff.net.0.proj.geglu.weight"), 8 * C, e); }
  { Epi e; e.out = h; e.ldc = C; e.bias = c.F(sb + ".ff.net.2.bias"); e.res = h1; e.ldr = C;
    op_linear(c, big, rows, 4 * C, 4 * C, c.M(sb + ".ff.net.2.weight"), C, e); }
  // h = spatial block output (x_spatial of the mixer)

  // ---- TemporalBasicTransformerBlock on h + emb_t[frame]
  const float* tp = m.time_pos.at(key);
  op_layernorm(c, h, rows, C, c.F(tb + ".norm_in.weight"), c.F(tb + ".norm_in.bias"), ln_eps, tp, hw, n);
  { Epi e; e.out = big; e.ldc = 4 * C; e.bias = c.F(tb + ".ff_in.net.0.proj.geglu.bias"); e.geglu = 1;
    op_linear(c, n, rows, C, C, c.M(tb + ".ff_in.net.0.proj.geglu.weight"), 8 * C, e); }
  { Epi e; e.out = h1; e.ldc = C; e.bias = c.F(tb + ".ff_in.net.2.bias"); e.res = h; e.ldr = C;
    e.fbias = tp; e.fbias_ld = C; e.fbias_div = hw;                           // residual is (h + emb_t)
    op_linear(c, big, rows, 4 * C, 4 * C, c.M(tb + ".ff_in.net.2.weight"), C, e); }
  op_layernorm(c, h1, rows, C, c.F(tb + ".norm1.weight"), c.F(tb + ".norm1.bias"), ln_eps, nullptr, 1, n);
  { Epi e; e.out = big; e.ldc = 3 * C;
    op_linear(c, n, rows, C, C, c.M(tb + ".attn1.to_qkv.weight"), 3 * C, e); }
  op_temporal_attention(c, big, ao, T, hw, C);
  void* h2 = n;   // n is free after the qkv projection
  { Epi e; e.out = h2; e.ldc = C; e.bias = c.F(tb + ".attn1.to_out.0.bias"); e.res = h1; e.ldr = C;
    e.fbias = m.attn2_temporal.at(key); e.fbias_ld = C; e.fbias_div = (int)rows;  // + attn2 (one vector)
    op_linear(c, ao, rows, C, C, c.M(tb + ".attn1.to_out.0.weight"), C, e); }
  op_layernorm(c, h2, rows, C, c.F(tb + ".norm3.weight"), c.F(tb + ".norm3.bias"), ln_eps, nullptr, 1, ao);
  { Epi e; e.out = big; e.ldc = 4 * C; e.bias = c.F(tb + ".ff.net.0.proj.geglu.bias"); e.geglu = 1;
    op_linear(c, ao, rows, C, C, c.M(tb + ".ff.net.0.proj.geglu.weight"), 8 * C, e); }
  const float alpha = sigmoidf_host(c.W(key + ".time_mixer.mix_factor").host.at(0));
  { Epi e; e.out = h1; e.ldc = C; e.bias = c.F(tb + ".ff.net.2.bias"); e.res = h2; e.ldr = C;
    e.blend = h; e.ldb = C; e.alpha = alpha;                                  // mix with x_spatial
    op_linear(c, big, rows, 4 * C, 4 * C, c.M(tb + ".ff.net.2.weight"), C, e); }
  { Epi e; e.out = out.p; e.ldc = C; e.bias = c.F(key + ".proj_out.bias"); e.res = x.p; e.ldr = C;
    op_linear(c, h1, rows, C, C, c.M(key + ".proj_out.weight"), C, e); }
  c.ws.release(mk);
  return out;
}

}  // namespace

void unet_forward(Ctx& c, const void* x16, float timestep, const float ids[3], float* v_out) {
  UG_CHECK(c.unet != nullptr, UG_ERR_STATE, "no UNet weights loaded / ug_ctx_finalize not called");
  UNetModel& m = *c.unet;
  UG_CHECK(c.dry || m.clip_context_set, UG_ERR_STATE, "ug_set_clip_context must precede the UNet forward");
  const ug_model_cfg& g = c.cfg;
  const int T = c.T, nb = g.unet_num_blocks, L = g.unet_layers_per_block;
  const int E = g.unet_block_out[0] * 4;

  // ---- time embedding -> one bias vector per resnet conv1 (K9)
  if (!c.dry) {
    float* s = m.scratch;
    float* t_sin = s;                      // [C0]
    float* t_h = s + 2048;                 // [E]
    float* emb_t = t_h + E;                // [E]
    float* id_sin = emb_t + E;             // [3 * add_dim]
    float* a_h = id_sin + 4096;            // [E]
    float* emb = a_h + E;                  // [E]
    const int C0 = g.unet_block_out[0], AD = g.addition_time_embed_dim;
    UG_CHECK(C0 <= 2048 && 3 * AD <= 4096 && g.num_added_ids == 3, UG_ERR_INVALID, "embedding dims");
    op_check(c, launch_sinusoid_vals(timestep, 0, 0, 0, 1, C0, t_sin, c.stream), "sinusoid(t)");
    op_gemv(c, c.M(U + "time_embedding.linear_1.weight"), c.F(U + "time_embedding.linear_1.bias"), nullptr, t_sin,
            t_h, 1, E, C0, 0, 1);
    op_gemv(c, c.M(U + "time_embedding.linear_2.weight"), c.F(U + "time_embedding.linear_2.bias"), nullptr, t_h,
            emb_t, 1, E, E, 0, 0);
    op_check(c, launch_sinusoid_vals(ids[0], ids[1], ids[2], 0, 3, AD, id_sin, c.stream), "sinusoid(ids)");
    op_gemv(c, c.M(U + "add_embedding.linear_1.weight"), c.F(U + "add_embedding.linear_1.bias"), nullptr, id_sin,
            a_h, 1, E, 3 * AD, 0, 1);
    op_gemv(c, c.M(U + "add_embedding.linear_2.weight"), c.F(U + "add_embedding.linear_2.bias"), emb_t, a_h, emb,
            1, E, E, 0, 0);
    op_gemv(c, c.M(U + "__temb_all.weight"), c.F(U + "__temb_all.bias"), nullptr, emb, m.temb_out, 1,
            m.temb_total, E, 1, 0);
  }
  auto tb = [&](const std::string& rk) -> const float* { return m.temb_out + m.temb_offset.at(rk); };
  auto stres = [&](const std::string& key, const Act& a, const Act* b, int cout, float eps) {
    return st_resblock(c, key, a, b, T, T, cout, tb(key + ".spatial_res_block"), tb(key + ".temporal_res_block"),
                       eps, eps, false);
  };

  Act x{nullptr, g.unet_block_out[0], c.h, c.w};
  x.p = c.alloc16((long long)T * c.h * c.w * x.C);
  { Epi e; e.out = x.p; e.ldc = x.C; e.bias = c.F(U + "conv_in.bias");
    op_conv3x3(c, x16, T, c.h, c.w, g.unet_in_channels, c.M(U + "conv_in.weight"), x.C, 1, 0, e); }
  std::vector<Act> skips;
  skips.push_back(x);

  for (int i = 0; i < nb; ++i) {
    const bool attn = i < nb - 1;
    const float eps = attn ? g.eps_cross_attn_block : g.eps_plain_block;
    const std::string b = U + "down_blocks." + std::to_string(i);
    for (int j = 0; j < L; ++j) {
      x = stres(b + ".resnets." + std::to_string(j), x, nullptr, g.unet_block_out[i], eps);
      if (attn) x = st_transformer(c, b + ".attentions." + std::to_string(j), x, T, g.unet_heads[i]);
      skips.push_back(x);
    }
    if (i < nb - 1) {
      UG_CHECK(x.H % 2 == 0 && x.W % 2 == 0, UG_ERR_INVALID, "latent size must be divisible by 8");
      Act d{c.alloc16((long long)T * (x.H / 2) * (x.W / 2) * x.C), x.C, x.H / 2, x.W / 2};
      Epi e; e.out = d.p; e.ldc = x.C; e.bias = c.F(b + ".downsamplers.0.conv.bias");
      op_conv3x3(c, x.p, T, x.H, x.W, x.C, c.M(b + ".downsamplers.0.conv.weight"), x.C, 2, 0, e);
      x = d;
      skips.push_back(x);
    }
  }

  const int cm = g.unet_block_out[nb - 1];
  x = stres(U + "mid_block.resnets.0", x, nullptr, cm, g.eps_plain_block);
  x = st_transformer(c, U + "mid_block.attentions.0", x, T, g.unet_heads[nb - 1]);
  x = stres(U + "mid_block.resnets.1", x, nullptr, cm, g.eps_plain_block);

  for (int i = 0; i < nb; ++i) {
    const bool attn = i > 0;
    const float eps = attn ? g.eps_cross_attn_block : g.eps_plain_up_block;
    const int co = g.unet_block_out[nb - 1 - i];
    const std::string b = U + "up_blocks." + std::to_string(i);
    for (int j = 0; j < L + 1; ++j) {
      Act sk = skips.back();
      skips.pop_back();
      x = stres(b + ".resnets." + std::to_string(j), x, &sk, co, eps);
      if (attn) x = st_transformer(c, b + ".attentions." + std::to_string(j), x, T, g.unet_heads[nb - 1 - i]);
    }
    if (i < nb - 1) {
      Act u{c.alloc16((long long)T * x.H * 2 * x.W * 2 * x.C), x.C, x.H * 2, x.W * 2};
      op_upsample2x(c, x.p, u.p, T, x.H, x.W, x.C);
      Act d{c.alloc16((long long)T * u.H * u.W * x.C), x.C, u.H, u.W};
      Epi e; e.out = d.p; e.ldc = x.C; e.bias = c.F(b + ".upsamplers.0.conv.bias");
      op_conv3x3(c, u.p, T, u.H, u.W, x.C, c.M(b + ".upsamplers.0.conv.weight"), x.C, 1, 0, e);
      x = d;
    }
  }

  const long long rows = (long long)T * x.H * x.W;
  void* n = c.alloc16(rows * x.C);
  op_gn(c, x.p, x.C, nullptr, 0, rows, (long long)x.H * x.W, c.F(U + "conv_norm_out.weight"),
        c.F(U + "conv_norm_out.bias"), g.eps_out_norm, 1, n);
  Epi e; e.out = v_out; e.ldc = g.unet_out_channels; e.out_fp32 = 1; e.bias = c.F(U + "conv_out.bias");
  op_conv3x3(c, n, T, x.H, x.W, x.C, c.M(U + "conv_out.weight"), g.unet_out_channels, 1, 0, e);
}

}  // namespace ug

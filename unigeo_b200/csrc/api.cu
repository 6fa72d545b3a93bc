// extern "C" boundary of libunigeo_b200.so (declared in include/unigeo_b200.h).
#include <cmath>
#include <cstring>
#include <functional>

#include "model.cuh"

using namespace ug;

// A launch-bound inner loop (denoising / refinement steps) captured once and replayed: hundreds of launches
// per step, each with host-side tensor-map encoding, become one cudaGraphLaunch.
struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  unsigned long long epoch = 0;    // Ctx::ptr_epoch the graph was captured under
  long long launches = 0;          // kernels in the graph (bench "gpu_launches" accounting)
  int calls = 0;                   // first call runs eagerly (one-time attribute / counter setup), second captures
  bool disabled = false;           // capture or instantiation failed once: stay eager
};

struct ug_ctx {
  Ctx c;
  std::unordered_map<std::string, size_t> ws_need;   // call signature -> workspace bytes
  std::unordered_map<std::string, GraphEntry> graphs;
  cudaStream_t gstream = nullptr;                    // capture needs a non-default stream (torch's default is 0)
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
};

namespace {

thread_local std::string g_err;

int guard(const std::function<void()>& f) {
  try {
    f();
    g_err.clear();
    return UG_OK;
  } catch (const UgError& e) {
    g_err = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_err = e.what();
    return UG_ERR_INVALID;
  }
}

// Runs `body` twice: once dry to learn the workspace high-water mark (cached per signature),
// then for real on `stream`.
void run_sized(ug_ctx* u, const std::string& sig, void* stream, const std::function<void(Ctx&)>& body) {
  Ctx& c = u->c;
  UG_CUDA(cudaSetDevice(c.device));
  auto it = u->ws_need.find(sig);
  if (it == u->ws_need.end()) {
    c.dry = true; c.ws.dry = true; c.ws.off = 0; c.ws.peak = 0;
    try { body(c); } catch (...) { c.dry = false; c.ws.dry = false; throw; }
    c.dry = false; c.ws.dry = false;
    it = u->ws_need.emplace(sig, c.ws.peak + (1 << 20)).first;
  }
  c.ensure_workspace(it->second);
  c.ws.off = 0;
  c.stream = reinterpret_cast<cudaStream_t>(stream);
  if (c.profile) c.prof_mark("(start)", 0.0, 0.0);
  body(c);
}

// Runs `body` (which must only enqueue kernels on c.stream, at fixed workspace addresses) through a cached CUDA
// graph keyed by `key`.  The graph lives on the context's own stream, fenced against the caller's stream with
// events, so it also works when the caller uses the legacy default stream.  UG_NO_GRAPH=1 disables.
void run_graphed(ug_ctx* u, const std::string& key, const std::function<void(Ctx&)>& body) {
  static const bool off = getenv("UG_NO_GRAPH") != nullptr;
  Ctx& c = u->c;
  if (u->graphs.size() > 16 && u->graphs.find(key) == u->graphs.end()) {   // bound the cache: drop everything, re-learn
    for (auto& kv : u->graphs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    u->graphs.clear();
  }
  GraphEntry& g = u->graphs[key];
  if (off || c.profile || g.disabled) { body(c); return; }
  if (g.exec && g.epoch != c.ptr_epoch) {
    cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
    g.calls = 0;
  }
  if (!g.exec && g.calls++ < 1) { body(c); return; }
  cudaStream_t user = c.stream;
  if (!u->gstream) {
    UG_CUDA(cudaStreamCreateWithFlags(&u->gstream, cudaStreamNonBlocking));
    UG_CUDA(cudaEventCreateWithFlags(&u->ev_in, cudaEventDisableTiming));
    UG_CUDA(cudaEventCreateWithFlags(&u->ev_out, cudaEventDisableTiming));
  }
  if (!g.exec) {
    const long long l0 = c.launches;
    const unsigned long long e0 = c.ptr_epoch;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(u->gstream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      c.stream = u->gstream;
      try {
        body(c);
      } catch (...) {
        c.stream = user;
        cudaStreamEndCapture(u->gstream, &graph);
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        throw;
      }
      c.stream = user;
      ok = cudaStreamEndCapture(u->gstream, &graph) == cudaSuccess && graph != nullptr && c.ptr_epoch == e0;
      if (ok) ok = cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess;
      if (graph) cudaGraphDestroy(graph);
    }
    if (!ok) {                       // e.g. a driver that cannot capture these launches: stay eager for this key
      cudaGetLastError();
      g.exec = nullptr;
      g.disabled = true;
      c.launches = l0;
      body(c);
      return;
    }
    g.epoch = c.ptr_epoch;
    g.launches = c.launches - l0;
    c.launches = l0;
  }
  UG_CUDA(cudaEventRecord(u->ev_in, user));
  UG_CUDA(cudaStreamWaitEvent(u->gstream, u->ev_in, 0));
  UG_CUDA(cudaGraphLaunch(g.exec, u->gstream));
  UG_CUDA(cudaEventRecord(u->ev_out, u->gstream));
  UG_CUDA(cudaStreamWaitEvent(user, u->ev_out, 0));
  c.launches += g.launches;
}

std::vector<double> karras_sigmas(const ug_model_cfg& g, int steps) {
  std::vector<double> s;
  const double lo = std::pow((double)g.sigma_min, 1.0 / g.rho), hi = std::pow((double)g.sigma_max, 1.0 / g.rho);
  for (int i = 0; i < steps; ++i) {
    const double r = steps > 1 ? (double)i / (steps - 1) : 0.0;
    s.push_back(std::pow(hi + r * (lo - hi), (double)g.rho));
  }
  s.push_back(0.0);
  return s;
}

// scaled-linear betas -> cumulative alpha products (double, like the host-side scheduler upstream)
std::vector<double> ddim_alphas_cumprod(int NT, double beta_start, double beta_end) {
  std::vector<double> ac(NT);
  const double b0 = std::sqrt(beta_start), b1 = std::sqrt(beta_end);
  double prod = 1.0;
  for (int i = 0; i < NT; ++i) {
    const double sb = NT > 1 ? b0 + (b1 - b0) * i / (NT - 1) : b0;
    prod *= 1.0 - sb * sb;
    ac[i] = prod;
  }
  return ac;
}
// trailing spacing from t_start (or NT - 1 when negative): round(top - k top / steps) - 1
std::vector<int> ddim_timesteps(int NT, int steps, int t_start) {
  const int top = t_start < 0 ? NT : t_start + 1;
  std::vector<int> ts(steps);
  for (int k = 0; k < steps; ++k) ts[k] = (int)std::nearbyint((double)top - (double)k * top / steps) - 1;
  return ts;
}

// matrices of the VAE encoder live in the encoder's own format (ug_ctx_set_vae_encode_dtype)
bool is_vae_encoder_key(const char* key) {
  return std::strncmp(key, "vae.encoder.", 12) == 0 || std::strncmp(key, "vae.quant_conv.", 15) == 0 ||
         std::strncmp(key, "vae2d.encoder.", 14) == 0 || std::strncmp(key, "vae2d.quant_conv.", 17) == 0;
}
int weight_fmt(const Ctx& c, const char* key) { return (c.enc_fmt >= 0 && is_vae_encoder_key(key)) ? c.enc_fmt : c.fmt; }

// the VAE encoder graph runs with the encoder's format as the context format
struct FmtScope {
  Ctx& c;
  int saved;
  FmtScope(Ctx& ctx, int fmt) : c(ctx), saved(ctx.fmt) { if (fmt >= 0) c.fmt = fmt; }
  ~FmtScope() { c.fmt = saved; }
};

thread_local ug_ctx* g_scratch[2] = {nullptr, nullptr};
ug_ctx* scratch_ctx(int dtype) {
  UG_CHECK(dtype == UG_F16 || dtype == UG_BF16, UG_ERR_INVALID, "dtype must be UG_F16 or UG_BF16");
  if (!g_scratch[dtype]) {
    g_scratch[dtype] = new ug_ctx();
    Ctx& c = g_scratch[dtype]->c;
    c.fmt = dtype;
    c.cfg.dtype = dtype;
    c.cfg.norm_groups = 32;
    UG_CUDA(cudaGetDevice(&c.device));
  }
  return g_scratch[dtype];
}

}  // namespace

extern "C" {

int ug_version(void) { return UG_VERSION; }
const char* ug_last_error(void) { return g_err.c_str(); }

int ug_ctx_create(ug_ctx** out, int device, const ug_model_cfg* cfg) {
  return guard([&] {
    UG_CHECK(out && cfg, UG_ERR_INVALID, "null argument");
    UG_CHECK(cfg->dtype == UG_F16 || cfg->dtype == UG_BF16, UG_ERR_INVALID, "cfg.dtype must be UG_F16 or UG_BF16");
    UG_CHECK(cfg->unet_num_blocks <= 4 && cfg->vae_num_blocks <= 4, UG_ERR_INVALID, "at most 4 blocks");
    UG_CUDA(cudaSetDevice(device));
    int major = 0;
    UG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    UG_CHECK(major == 10, UG_ERR_CUDA, "unigeo_b200 kernels are built for sm_100a only (B200)");
    ug_ctx* u = new ug_ctx();
    u->c.device = device;
    u->c.cfg = *cfg;
    u->c.fmt = cfg->dtype;
    *out = u;
  });
}

int ug_ctx_destroy(ug_ctx* u) {
  return guard([&] {
    if (!u) return;
    cudaSetDevice(u->c.device);
    cudaDeviceSynchronize();
    for (auto& kv : u->graphs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (u->gstream) cudaStreamDestroy(u->gstream);
    if (u->ev_in) cudaEventDestroy(u->ev_in);
    if (u->ev_out) cudaEventDestroy(u->ev_out);
    for (void* p : u->c.owned) cudaFree(p);
    if (u->c.ws.base) cudaFree(u->c.ws.base);
    delete u->c.unet;
    delete u->c.vae;
    delete u->c.nets2d;
    delete u;
  });
}

int ug_ctx_load_weight(ug_ctx* u, const char* key, const void* dev_ptr, int dtype, const int64_t* shape, int rank,
                       void* stream) {
  return guard([&] {
    UG_CHECK(u && key && dev_ptr && shape, UG_ERR_INVALID, "null argument");
    UG_CHECK(rank >= 1 && rank <= 5, UG_ERR_WEIGHT, std::string("unsupported weight rank: ") + key);
    UG_CHECK(dtype >= 0 && dtype <= 2, UG_ERR_INVALID, "bad dtype");
    Ctx& c = u->c;
    UG_CUDA(cudaSetDevice(c.device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Weight w;
    long long n = 1;
    for (int i = 0; i < rank; ++i) n *= shape[i];
    w.numel = n;
    if (rank == 1) {
      w.is_f32 = 1;
      w.cout = (int)shape[0];
      w.p = c.dmalloc((size_t)n * 4);
      UG_CUDA(launch_convert_f32(dev_ptr, dtype, reinterpret_cast<float*>(w.p), n, st));
      if (n <= 16) {
        w.host.resize(n);
        UG_CUDA(cudaStreamSynchronize(st));
        UG_CUDA(cudaMemcpy(w.host.data(), w.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
      }
    } else {
      w.cout = (int)shape[0];
      w.cin = (int)shape[1];
      w.taps = 1;
      for (int i = 2; i < rank; ++i) w.taps *= (int)shape[i];
      UG_CHECK(w.taps == 1 || w.taps == 3 || w.taps == 9, UG_ERR_WEIGHT, std::string("unsupported kernel size: ") + key);
      w.cin_pad = (w.cin + 7) & ~7;
      const size_t bytes = (size_t)w.taps * w.cout * w.cin_pad * 2;
      w.p = c.dmalloc(bytes);
      if (w.cin_pad != w.cin) UG_CUDA(cudaMemsetAsync(w.p, 0, bytes, st));
      UG_CUDA(launch_convert_weight(dev_ptr, dtype, w.p, w.cout, w.cin, w.cin_pad, w.taps, weight_fmt(c, key), st));
      w.cin = w.cin_pad;
    }
    c.weights[key] = w;
    c.finalized = false;
    ++c.ptr_epoch;
  });
}

int ug_ctx_load_weights(ug_ctx* u, int n, const char* const* keys, const void* const* dev_ptrs, const int* dtypes,
                        const int64_t* shapes, const int* ranks, void* stream) {
  return guard([&] {
    UG_CHECK(u && keys && dev_ptrs && dtypes && shapes && ranks, UG_ERR_INVALID, "null argument");
    UG_CHECK(n >= 0, UG_ERR_INVALID, "n must be non-negative");
    if (n == 0) return;
    Ctx& c = u->c;
    UG_CUDA(cudaSetDevice(c.device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    std::vector<ConvertDesc> descs((size_t)n);
    std::vector<Weight> ws((size_t)n);
    long long blocks = 0;
    for (int i = 0; i < n; ++i) {
      const char* key = keys[i];
      const int64_t* shape = shapes + (size_t)i * 5;
      const int rank = ranks[i];
      UG_CHECK(key && dev_ptrs[i], UG_ERR_INVALID, "null tensor");
      UG_CHECK(rank >= 1 && rank <= 5, UG_ERR_WEIGHT, std::string("unsupported weight rank: ") + key);
      UG_CHECK(dtypes[i] >= 0 && dtypes[i] <= 2, UG_ERR_INVALID, "bad dtype");
      Weight& w = ws[i];
      long long numel = 1;
      for (int k = 0; k < rank; ++k) numel *= shape[k];
      w.numel = numel;
      ConvertDesc& d = descs[i];
      d.src = dev_ptrs[i]; d.total = numel; d.src_dtype = dtypes[i]; d.first_block = blocks;
      if (rank == 1) {
        w.is_f32 = 1;
        w.cout = (int)shape[0];
        w.p = c.dmalloc((size_t)numel * 4);
        d.dst_fmt = 2; d.cout = w.cout; d.cin = d.cin_pad = d.taps = 1;
      } else {
        w.cout = (int)shape[0];
        w.cin = (int)shape[1];
        w.taps = 1;
        for (int k = 2; k < rank; ++k) w.taps *= (int)shape[k];
        UG_CHECK(w.taps == 1 || w.taps == 3 || w.taps == 9, UG_ERR_WEIGHT, std::string("unsupported kernel size: ") + key);
        w.cin_pad = (w.cin + 7) & ~7;
        const size_t bytes = (size_t)w.taps * w.cout * w.cin_pad * 2;
        w.p = c.dmalloc(bytes);
        if (w.cin_pad != w.cin) UG_CUDA(cudaMemsetAsync(w.p, 0, bytes, st));
        d.dst_fmt = weight_fmt(c, key); d.cout = w.cout; d.cin = w.cin; d.cin_pad = w.cin_pad; d.taps = w.taps;
        w.cin = w.cin_pad;
      }
      d.dst = w.p;
      long long nb = (numel + 256 * 16 - 1) / (256 * 16);       // 16 elements per thread, 1..2048 CTAs per tensor
      blocks += nb < 1 ? 1 : (nb > 2048 ? 2048 : nb);
    }
    ConvertDesc* dd = nullptr;
    UG_CUDA(cudaMalloc(&dd, descs.size() * sizeof(ConvertDesc)));
    cudaError_t e = cudaMemcpyAsync(dd, descs.data(), descs.size() * sizeof(ConvertDesc), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = (cudaError_t)launch_convert_batch(dd, n, blocks, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(dd);
    UG_CUDA(e);
    c.launches++;
    for (int i = 0; i < n; ++i) {
      Weight& w = ws[i];
      if (w.is_f32 && w.numel <= 16) {
        w.host.resize(w.numel);
        UG_CUDA(cudaMemcpy(w.host.data(), w.p, (size_t)w.numel * 4, cudaMemcpyDeviceToHost));
      }
      c.weights[keys[i]] = w;
    }
    c.finalized = false;
    ++c.ptr_epoch;
  });
}

int ug_ctx_set_vae_encode_dtype(ug_ctx* u, int dtype) {
  return guard([&] {
    UG_CHECK(u, UG_ERR_INVALID, "null ctx");
    UG_CHECK(dtype == -1 || dtype == UG_F16 || dtype == UG_BF16, UG_ERR_INVALID, "dtype must be -1, UG_F16 or UG_BF16");
    for (const auto& kv : u->c.weights)
      UG_CHECK(kv.second.is_f32 || !is_vae_encoder_key(kv.first.c_str()), UG_ERR_STATE,
               "ug_ctx_set_vae_encode_dtype must precede the encoder's weights");
    u->c.enc_fmt = dtype == u->c.fmt ? -1 : dtype;
  });
}

int ug_ctx_finalize(ug_ctx* u, void* stream) {
  return guard([&] {
    UG_CHECK(u, UG_ERR_INVALID, "null ctx");
    Ctx& c = u->c;
    UG_CUDA(cudaSetDevice(c.device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unet_finalize(c, st);
    unet2d_finalize(c, st);
    vae_finalize(c, st);
    clip_finalize(c, st);
    UG_CUDA(cudaStreamSynchronize(st));
    c.finalized = true;
  });
}

int ug_ctx_prepare(ug_ctx* u, int T, int h, int w, void* stream) {
  return guard([&] {
    UG_CHECK(u && u->c.finalized, UG_ERR_STATE, "ug_ctx_finalize must precede ug_ctx_prepare");
    UG_CHECK(T >= 1 && T <= 64 && h >= 1 && w >= 1, UG_ERR_INVALID, "need 1 <= T <= 64 frames");
    Ctx& c = u->c;
    UG_CUDA(cudaSetDevice(c.device));
    if (c.T != T || c.h != h || c.w != w) u->ws_need.clear();
    c.T = T; c.h = h; c.w = w;
    c.stream = reinterpret_cast<cudaStream_t>(stream);
    unet_prepare(c, T, c.stream);
    UG_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int ug_set_clip_context(ug_ctx* u, const float* enc, void* stream) {
  return guard([&] {
    UG_CHECK(u && enc, UG_ERR_INVALID, "null argument");
    run_sized(u, "clipctx", stream, [&](Ctx& c) {
      if (c.dry) { c.allocf((long long)c.T * 4096); return; }
      unet_set_clip_context(c, enc, c.stream);
    });
  });
}

int ug_unet_st_forward(ug_ctx* u, const float* x, float timestep, const float* ids, float* out, void* stream) {
  return guard([&] {
    UG_CHECK(u && x && ids && out, UG_ERR_INVALID, "null argument");
    UG_CHECK(u->c.T > 0, UG_ERR_STATE, "ug_ctx_prepare must precede ug_unet_st_forward");
    run_sized(u, "unet", stream, [&](Ctx& c) {
      const long long hw = (long long)c.h * c.w, tok = hw * c.T;
      const int Ci = c.cfg.unet_in_channels, Co = c.cfg.unet_out_channels;
      UG_CHECK(Ci == 8 && Co == 4, UG_ERR_INVALID, "UNet I/O must be 8 -> 4 channels");
      void* x16 = c.alloc16(tok * Ci);
      float* v = c.allocf(tok * Co);
      if (!c.dry)
        op_check(c, launch_nchw_to_nhwc(x, nullptr, 0.f, 1.f, 0.f, c.T, Ci, c.h, c.w, Ci, x16, c.fmt, c.stream),
                 "nchw_to_nhwc");
      unet_forward(c, x16, timestep, ids, v);
      if (!c.dry) op_check(c, launch_f32_nhwc_to_nchw(v, c.T, hw, Co, 1.f, out, c.stream), "f32 swap");
    });
  });
}

int ug_denoise_clip(ug_ctx* u, const float* cond_lat, const float* init_noise, const float* ids, int steps,
                    float* lat_out, void* stream) {
  return guard([&] {
    UG_CHECK(u && cond_lat && init_noise && ids && lat_out, UG_ERR_INVALID, "null argument");
    UG_CHECK(steps >= 1 && steps <= 1000, UG_ERR_INVALID, "steps out of range");
    UG_CHECK(u->c.T > 0, UG_ERR_STATE, "ug_ctx_prepare must precede ug_denoise_clip");
    const std::vector<double> sig = karras_sigmas(u->c.cfg, steps);
    run_sized(u, "denoise", stream, [&](Ctx& c) {
      const long long hw = (long long)c.h * c.w, tok = hw * c.T;
      void* cond16 = c.alloc16(tok * 4);
      float* lat = c.allocf(tok * 4);
      void* x16 = c.alloc16(tok * 8);
      float* v = c.allocf(tok * 4);
      const size_t mk = c.ws.mark();
      if (!c.dry) {
        op_check(c, launch_nchw_to_nhwc(cond_lat, nullptr, 0.f, 1.f, 0.f, c.T, 4, c.h, c.w, 4, cond16, c.fmt,
                                        c.stream), "cond latents");
        const float s0 = (float)std::sqrt(sig[0] * sig[0] + 1.0);   // init_noise_sigma ("leading" spacing)
        op_check(c, launch_f32_nchw_to_nhwc(init_noise, c.T, hw, 4, s0, lat, c.stream), "init latents");
      }
      auto loop = [&](Ctx& cc) {
        const int n_iter = cc.dry ? 1 : steps;
        for (int i = 0; i < n_iter; ++i) {
          cc.ws.release(mk);
          const float sg = (float)sig[i], sn = (float)sig[i + 1];
          if (!cc.dry) op_check(cc, launch_build_unet_input(lat, cond16, sg, tok, x16, cc.fmt, cc.stream), "unet input");
          unet_forward(cc, x16, (float)(0.25 * std::log(sig[i])), ids, v);
          if (!cc.dry) op_check(cc, launch_euler_step(lat, v, sg, sn, tok * 4, cc.stream), "euler step");
        }
      };
      if (c.dry) {
        loop(c);
      } else {
        char key[160];
        snprintf(key, sizeof(key), "denoise:%dx%dx%d:%d:%a:%a:%a", c.T, c.h, c.w, steps, ids[0], ids[1], ids[2]);
        run_graphed(u, key, loop);
      }
      if (!c.dry) op_check(c, launch_f32_nhwc_to_nchw(lat, c.T, hw, 4, 1.f, lat_out, c.stream), "latents out");
    });
  });
}

int ug_vae_encode(ug_ctx* u, const float* img, const float* noise, float noise_strength, int N, int H, int W,
                  float* lat_mean, void* stream) {
  return guard([&] {
    UG_CHECK(u && img && lat_mean, UG_ERR_INVALID, "null argument");
    UG_CHECK(u->c.finalized, UG_ERR_STATE, "ug_ctx_finalize must precede ug_vae_encode");
    UG_CHECK(N >= 1 && H % 8 == 0 && W % 8 == 0, UG_ERR_INVALID, "H and W must be multiples of 8");
    const std::string sig = "enc:" + std::to_string(N) + "x" + std::to_string(H) + "x" + std::to_string(W);
    run_sized(u, sig, stream, [&](Ctx& c) {
      FmtScope fs(c, c.enc_fmt);
      void* img16 = c.alloc16((long long)N * H * W * 8);
      if (!c.dry)
        op_check(c, launch_nchw_to_nhwc(img, noise, noise_strength, 1.f, 0.f, N, c.cfg.vae_in_channels, H, W, 8,
                                        img16, c.fmt, c.stream), "image in");
      vae_encode(c, "vae.", img16, N, H, W, 1.f, lat_mean);
    });
  });
}

int ug_ctx_set_clip_cfg(ug_ctx* u, const ug_clip_cfg* cfg) {
  return guard([&] {
    UG_CHECK(u && cfg, UG_ERR_INVALID, "null argument");
    UG_CHECK(cfg->layers >= 1 && cfg->heads >= 1 && cfg->hidden % cfg->heads == 0 && cfg->hidden % 8 == 0 &&
                 cfg->mlp % 8 == 0 && cfg->patch >= 1 && cfg->image_size % cfg->patch == 0 && cfg->proj_dim >= 1,
             UG_ERR_INVALID, "bad CLIP config");
    u->c.cfg_clip = *cfg;
    u->c.finalized = false;
  });
}

int ug_clip_embed(ug_ctx* u, const float* video, int F, int H, int W, float* enc, void* stream) {
  return guard([&] {
    UG_CHECK(u && video && enc, UG_ERR_INVALID, "null argument");
    UG_CHECK(u->c.finalized, UG_ERR_STATE, "ug_ctx_finalize must precede ug_clip_embed");
    UG_CHECK(F >= 1 && H >= 2 && W >= 2, UG_ERR_INVALID, "bad shape");
    const std::string sig = "clip:" + std::to_string(F) + "x" + std::to_string(H) + "x" + std::to_string(W);
    run_sized(u, sig, stream, [&](Ctx& c) { clip_embed(c, video, F, H, W, enc); });
  });
}

int ug_prepare_frames(ug_ctx* u, const float* images, int T, int H, int W, float* frames, void* stream) {
  return guard([&] {
    UG_CHECK(u && images && frames, UG_ERR_INVALID, "null argument");
    UG_CHECK(T >= 1 && H >= 1 && W >= 1, UG_ERR_INVALID, "bad shape");
    Ctx& c = u->c;
    UG_CUDA(cudaSetDevice(c.device));
    c.stream = reinterpret_cast<cudaStream_t>(stream);
    op_check(c, launch_images_in(images, T, (long long)H * W, frames, c.stream), "images_in");
  });
}

int ug_vae_encode_frames(ug_ctx* u, const float* frames, const float* noise, float noise_strength, int T, int H,
                         int W, float* video_nchw, float* lat_mean, void* stream) {
  return guard([&] {
    UG_CHECK(u && frames && lat_mean, UG_ERR_INVALID, "null argument");
    UG_CHECK(u->c.finalized, UG_ERR_STATE, "ug_ctx_finalize must precede ug_vae_encode_frames");
    UG_CHECK(T >= 1 && H % 8 == 0 && W % 8 == 0, UG_ERR_INVALID, "H and W must be multiples of 8");
    UG_CHECK(u->c.cfg.vae_in_channels == 3, UG_ERR_INVALID, "frames are RGB");
    const std::string sig = "enc:" + std::to_string(T) + "x" + std::to_string(H) + "x" + std::to_string(W);
    run_sized(u, sig, stream, [&](Ctx& c) {
      FmtScope fs(c, c.enc_fmt);
      void* img16 = c.alloc16((long long)T * H * W * 8);
      if (!c.dry)
        op_check(c, launch_frames_in(frames, noise, noise_strength, T, (long long)H * W, img16, video_nchw, c.fmt,
                                     c.stream), "frames_in");
      vae_encode(c, "vae.", img16, T, H, W, 1.f, lat_mean);
    });
  });
}

int ug_vae_decode_frames(ug_ctx* u, const float* lat, int T, int h, int w, int chunk, float* frames, void* stream) {
  return guard([&] {
    UG_CHECK(u && lat && frames, UG_ERR_INVALID, "null argument");
    UG_CHECK(u->c.finalized, UG_ERR_STATE, "ug_ctx_finalize must precede ug_vae_decode_frames");
    UG_CHECK(T >= 1 && chunk >= 1, UG_ERR_INVALID, "T and chunk must be positive");
    const std::string sig = "dec:" + std::to_string(T) + "x" + std::to_string(h) + "x" + std::to_string(w) + "/" +
                            std::to_string(chunk);
    run_sized(u, sig, stream, [&](Ctx& c) {
      void* z16 = c.alloc16((long long)T * h * w * 8);
      if (!c.dry)
        op_check(c, launch_nchw_to_nhwc(lat, nullptr, 0.f, 1.0f / c.cfg.vae_scaling_factor, 0.f, T,
                                        c.cfg.vae_latent_channels, h, w, 8, z16, c.fmt, c.stream), "latents in");
      vae_decode(c, z16, T, h, w, chunk, nullptr, frames);
    });
  });
}

int ug_depth_postprocess(ug_ctx* u, const float* frames, const float* intrinsics, int T, int H, int W, float* depths,
                         float* normals, void* stream) {
  return guard([&] {
    UG_CHECK(u && frames && intrinsics && depths && normals, UG_ERR_INVALID, "null argument");
    UG_CHECK(T >= 1 && H >= 1 && W >= 1, UG_ERR_INVALID, "bad shape");
    const std::string sig = "post:" + std::to_string(T) + "x" + std::to_string(H) + "x" + std::to_string(W);
    run_sized(u, sig, stream, [&](Ctx& c) {
      float* ws = c.allocf(post_workspace_floats((long long)T * H * W));
      if (!c.dry) {
        op_check(c, launch_depth_postprocess(frames, intrinsics, T, H, W, depths, normals, ws, c.stream),
                 "depth_postprocess", 0.0, 28.0 * T * H * W);
        c.launches += 2;   // three kernels behind one launcher
      }
    });
  });
}

#ifdef UG_TAPGEMM_TRACE
// trace builds only (tools/trace_tapgemm.py; not part of include/unigeo_b200.h): device buffer the following tapgemm
// launches stamp -- [grid units][4 roles][16 tiles][4] uint64 -- or NULL to stop
extern "C" int ug_debug_tapgemm_trace(unsigned long long* dev_buf) {
  tapgemm_set_trace(dev_buf);
  return 0;
}
#endif

long long ug_fastdiv(int d, int n) {
  if (d < 1 || n < 0) return (long long)UG_ERR_INVALID;
  return (long long)fastdiv_host(make_fastdiv(d), n);
}

int ug_tile_schedule(int m_units, int n_total, int bn_tile, int batch, int n_fastest, int ctas, int slots, int k_iters,
                     int* units_out, int cap, long long* max_cost, long long* rr_max_cost) {
  int ret = 0;
  const int rc = guard([&] {
    UG_CHECK(m_units >= 1 && n_total >= 1 && batch >= 1 && slots >= 1 && k_iters >= 1, UG_ERR_INVALID, "bad extents");
    UG_CHECK((ctas == 1 || ctas == 2) && bn_tile >= 16 * ctas && bn_tile <= 256 && bn_tile % (16 * ctas) == 0,
             UG_ERR_INVALID, "bn_tile must be a multiple of 16 * ctas, <= 256");
    const int n_tiles = (n_total + bn_tile - 1) / bn_tile;
    UG_CHECK((long long)m_units * n_tiles * batch <= (1 << 22), UG_ERR_INVALID, "too many units");
    std::vector<int> table;
    const int len = tapgemm_build_schedule(m_units, n_tiles, batch, n_fastest ? 1 : 0, n_total, bn_tile, ctas, slots,
                                           k_iters, &table, max_cost, rr_max_cost);
    if (units_out != nullptr) {
      UG_CHECK((long long)cap >= (long long)slots * len, UG_ERR_INVALID, "units_out too small");
      std::memcpy(units_out, table.data(), table.size() * sizeof(int));
    }
    ret = len;
  });
  return rc != UG_OK ? rc : ret;
}

int ug_depth_metrics(ug_ctx* u, const float* pred, const float* gt, const unsigned char* mask, long long n,
                     float max_depth, double* out11, float* err_map, float* pred_aligned, float* gt_valid,
                     void* stream) {
  return guard([&] {
    UG_CHECK(u && pred && gt && out11, UG_ERR_INVALID, "null argument");
    UG_CHECK(n >= 1, UG_ERR_INVALID, "n must be positive");
    run_sized(u, "dmetrics:" + std::to_string(n), stream, [&](Ctx& c) {
      void* ws = c.ws.alloc((size_t)metrics_workspace_bytes(n));
      if (!c.dry) {
        op_check(c, launch_depth_metrics(pred, gt, mask, n, max_depth, ws, out11, err_map, pred_aligned, gt_valid,
                                         c.stream), "depth_metrics", 0.0, (mask ? 17.0 : 16.0) * n);
        c.launches += 4;   // fit, fold, solve, errors, fold behind one launcher
      }
    });
  });
}

int ug_normal_metrics(ug_ctx* u, const float* pred, const float* gt, const unsigned char* mask, long long n,
                      double* out8, float* err_deg, void* stream) {
  return guard([&] {
    UG_CHECK(u && pred && gt && out8, UG_ERR_INVALID, "null argument");
    UG_CHECK(n >= 1, UG_ERR_INVALID, "n must be positive");
    run_sized(u, "nmetrics:" + std::to_string(n), stream, [&](Ctx& c) {
      void* ws = c.ws.alloc((size_t)metrics_workspace_bytes(n));
      if (!c.dry) {
        op_check(c, launch_normal_metrics(pred, gt, mask, n, ws, out8, err_deg, c.stream), "normal_metrics", 0.0,
                 (28.0 + 4.0 * 4.0 + (mask ? 1.0 : 0.0)) * n);
        c.launches += 9;   // errors, fold, 4 x (histogram, step) behind one launcher
      }
    });
  });
}

int ug_stitch_fit(ug_ctx* u, const float* overlap_frames, int world, int per_rank, int num_clips, long long n, int space,
                  float offset, double* chain_dev, void* stream) {
  return guard([&] {
    UG_CHECK(u && overlap_frames && chain_dev, UG_ERR_INVALID, "null argument");
    UG_CHECK(world >= 1 && per_rank >= 1 && num_clips >= 1 && num_clips <= world * per_rank && n >= 1, UG_ERR_INVALID,
             "bad extents");
    UG_CHECK(space == 0 || space == 1, UG_ERR_INVALID, "space must be 0 (as given) or 1 (1 / v - offset)");
    run_sized(u, "stitch:" + std::to_string(num_clips), stream, [&](Ctx& c) {
      void* ws = c.ws.alloc((size_t)stitch_workspace_bytes(num_clips));
      if (!c.dry) {
        op_check(c, launch_stitch_fit(overlap_frames, world, per_rank, num_clips, n, space, offset, ws, chain_dev, c.stream),
                 "stitch_fit", 0.0, 8.0 * n * (num_clips - 1));
        if (num_clips > 1) c.launches += 1;   // fit + chain behind one launcher
      }
    });
  });
}

int ug_stitch_apply(ug_ctx* u, const float* clip, long long elems, const float* prev_tail, long long n_overlap,
                    long long frame_elems, int overlap, const double* chain_dev, int k, int space, float offset, float* out,
                    void* stream) {
  return guard([&] {
    UG_CHECK(u && clip && chain_dev && out, UG_ERR_INVALID, "null argument");
    UG_CHECK(elems >= 1 && k >= 0 && frame_elems >= 1 && overlap >= 0 && n_overlap == (long long)overlap * frame_elems &&
                 n_overlap <= elems, UG_ERR_INVALID, "bad extents");
    UG_CHECK(k > 0 || prev_tail == nullptr, UG_ERR_INVALID, "clip 0 has no predecessor");
    Ctx& c = u->c;
    UG_CUDA(cudaSetDevice(c.device));
    c.stream = reinterpret_cast<cudaStream_t>(stream);
    op_check(c, launch_stitch_apply(clip, elems, prev_tail, n_overlap, frame_elems, overlap, chain_dev, k, space, offset,
                                    out, c.stream), "stitch_apply", 0.0, 8.0 * elems);
  });
}

int ug_vae_decode_temporal(ug_ctx* u, const float* lat, int T, int h, int w, int chunk, float* img, void* stream) {
  return guard([&] {
    UG_CHECK(u && lat && img, UG_ERR_INVALID, "null argument");
    UG_CHECK(u->c.finalized, UG_ERR_STATE, "ug_ctx_finalize must precede ug_vae_decode_temporal");
    UG_CHECK(T >= 1 && chunk >= 1, UG_ERR_INVALID, "T and chunk must be positive");
    const std::string sig = "dec:" + std::to_string(T) + "x" + std::to_string(h) + "x" + std::to_string(w) + "/" +
                            std::to_string(chunk);
    run_sized(u, sig, stream, [&](Ctx& c) {
      void* z16 = c.alloc16((long long)T * h * w * 8);
      if (!c.dry)
        op_check(c, launch_nchw_to_nhwc(lat, nullptr, 0.f, 1.0f / c.cfg.vae_scaling_factor, 0.f, T,
                                        c.cfg.vae_latent_channels, h, w, 8, z16, c.fmt, c.stream), "latents in");
      vae_decode(c, z16, T, h, w, chunk, img);
    });
  });
}

// ------------------------------------------------------------------ StableNormal path (2-D UNet)
int ug_ctx_set_unet2d_cfg(ug_ctx* u, const ug_unet2d_cfg* cfg) {
  return guard([&] {
    UG_CHECK(u && cfg, UG_ERR_INVALID, "null argument");
    UG_CHECK(cfg->num_blocks >= 2 && cfg->num_blocks <= 4, UG_ERR_INVALID, "2-D UNet needs 2..4 blocks");
    UG_CHECK(cfg->in_channels >= 1 && cfg->in_channels <= 8 && cfg->out_channels >= 1 && cfg->out_channels <= 8,
             UG_ERR_INVALID, "2-D UNet I/O channels must be 1..8");
    UG_CHECK(cfg->num_train_timesteps >= 1 && cfg->num_train_timesteps <= 100000, UG_ERR_INVALID,
             "num_train_timesteps out of range");
    u->c.cfg2d = *cfg;
    u->c.finalized = false;
  });
}

int ug_set_text_context(ug_ctx* u, const char* net_prefix, const float* tokens, int frames, int len, void* stream) {
  return guard([&] {
    UG_CHECK(u && net_prefix && tokens, UG_ERR_INVALID, "null argument");
    UG_CHECK(u->c.finalized, UG_ERR_STATE, "ug_ctx_finalize must precede ug_set_text_context");
    const std::string P = norm_prefix(net_prefix);
    run_sized(u, "textctx:" + std::to_string(frames) + "x" + std::to_string(len), stream,
              [&](Ctx& c) { unet2d_set_context(c, P, tokens, frames, len); });
  });
}

namespace {
// fp32 NCHW [F][Cin][h][w] -> 16-bit tokens [F][hw][8]
void* latents_to_tokens(Ctx& c, const float* x, int F, int h, int w, int Cin) {
  void* x16 = c.alloc16((long long)F * h * w * 8);
  if (!c.dry)
    op_check(c, launch_nchw_to_nhwc(x, nullptr, 0.f, 1.f, 0.f, F, Cin, h, w, 8, x16, c.fmt, c.stream), "nchw_to_nhwc");
  return x16;
}
}  // namespace

int ug_unet2d_forward(ug_ctx* u, const char* unet_prefix, const float* x, int F, int h, int w, float timestep,
                      const char* controlnet_prefix, const float* controlnet_sample, float* out, void* stream) {
  return guard([&] {
    UG_CHECK(u && unet_prefix && x && out, UG_ERR_INVALID, "null argument");
    UG_CHECK(u->c.finalized, UG_ERR_STATE, "ug_ctx_finalize must precede ug_unet2d_forward");
    UG_CHECK(F >= 1 && h >= 1 && w >= 1, UG_ERR_INVALID, "bad shape");
    UG_CHECK((controlnet_prefix == nullptr) == (controlnet_sample == nullptr), UG_ERR_INVALID,
             "controlnet_prefix and controlnet_sample go together");
    const std::string P = norm_prefix(unet_prefix), Q = norm_prefix(controlnet_prefix);
    const std::string sig = "unet2d:" + P + Q + ":" + std::to_string(F) + "x" + std::to_string(h) + "x" + std::to_string(w);
    run_sized(u, sig, stream, [&](Ctx& c) {
      const int Ci = c.cfg2d.in_channels, Co = c.cfg2d.out_channels;
      const long long hw = (long long)h * w;
      void* x16 = latents_to_tokens(c, x, F, h, w, Ci);
      void* c16 = controlnet_sample ? latents_to_tokens(c, controlnet_sample, F, h, w, Ci) : nullptr;
      float* v = c.allocf(F * hw * Co);
      unet2d_forward(c, P, x16, F, h, w, timestep, controlnet_prefix ? &Q : nullptr, c16, v);
      if (!c.dry) op_check(c, launch_f32_nhwc_to_nchw(v, F, hw, Co, 1.f, out, c.stream), "f32 swap");
    });
  });
}

int ug_refine_frames_2d(ug_ctx* u, const char* unet_prefix, const char* controlnet_prefix, const float* image_latent,
                        const float* latents_in, int F, int h, int w, int steps, int t_start, float* latents_out,
                        void* stream) {
  return guard([&] {
    UG_CHECK(u && unet_prefix && latents_in && latents_out, UG_ERR_INVALID, "null argument");
    UG_CHECK(u->c.finalized, UG_ERR_STATE, "ug_ctx_finalize must precede ug_refine_frames_2d");
    UG_CHECK((controlnet_prefix == nullptr) || image_latent, UG_ERR_INVALID, "a ControlNet needs image_latent");
    const ug_unet2d_cfg& g = u->c.cfg2d;
    UG_CHECK(g.num_blocks > 0, UG_ERR_STATE, "ug_ctx_set_unet2d_cfg was not called");
    UG_CHECK(g.in_channels == 4 && g.out_channels == 4, UG_ERR_INVALID, "refinement needs a 4 -> 4 channel UNet");
    const int NT = g.num_train_timesteps;
    UG_CHECK(steps >= 1 && steps <= NT && t_start < NT, UG_ERR_INVALID, "steps / t_start out of range");
    const std::vector<double> ac = ddim_alphas_cumprod(NT, (double)g.beta_start, (double)g.beta_end);
    const std::vector<int> ts = ddim_timesteps(NT, steps, t_start);
    const std::string P = norm_prefix(unet_prefix), Q = norm_prefix(controlnet_prefix);
    const std::string sig = "refine2d:" + P + Q + ":" + std::to_string(F) + "x" + std::to_string(h) + "x" + std::to_string(w);
    run_sized(u, sig, stream, [&](Ctx& c) {
      const long long hw = (long long)h * w, tok = hw * F;
      float* lat = c.allocf(tok * 4);
      float* x0 = c.allocf(tok * 4);
      void* x16 = c.alloc16(tok * 8);
      void* c16 = controlnet_prefix ? latents_to_tokens(c, image_latent, F, h, w, 4) : nullptr;
      if (!c.dry) op_check(c, launch_f32_nchw_to_nhwc(latents_in, F, hw, 4, 1.f, lat, c.stream), "latents in");
      const size_t mk = c.ws.mark();
      auto loop = [&](Ctx& cc) {
        const int n_iter = cc.dry ? 1 : steps;
        for (int i = 0; i < n_iter; ++i) {
          cc.ws.release(mk);
          const int t = ts[i], tp = i + 1 < steps ? ts[i + 1] : -1;
          UG_CHECK(t >= 0 && t < NT, UG_ERR_INVALID, "timestep out of range");
          if (!cc.dry) op_check(cc, launch_f32_to_tokens(lat, 4, 8, tok, x16, cc.fmt, cc.stream), "unet input");
          unet2d_forward(cc, P, x16, F, h, w, (float)t, controlnet_prefix ? &Q : nullptr, c16, x0);
          const double a_t = ac[t], a_p = tp >= 0 ? ac[tp] : 1.0;
          const double cx = std::sqrt((1.0 - a_p) / (1.0 - a_t));
          const double cx0 = std::sqrt(a_p) - std::sqrt(a_t) * cx;
          if (!cc.dry) op_check(cc, launch_axpby(lat, x0, (float)cx0, (float)cx, tok * 4, cc.stream), "ddim step");
        }
      };
      if (c.dry) loop(c);
      else run_graphed(u, sig + ":" + std::to_string(steps) + ":" + std::to_string(t_start), loop);
      if (!c.dry) op_check(c, launch_f32_nhwc_to_nchw(lat, F, hw, 4, 1.f, latents_out, c.stream), "latents out");
    });
  });
}

int ug_vae2d_encode(ug_ctx* u, const float* img, int N, int H, int W, float out_scale, float* lat, void* stream) {
  return guard([&] {
    UG_CHECK(u && img && lat, UG_ERR_INVALID, "null argument");
    UG_CHECK(u->c.finalized, UG_ERR_STATE, "ug_ctx_finalize must precede ug_vae2d_encode");
    UG_CHECK(N >= 1 && H % 8 == 0 && W % 8 == 0, UG_ERR_INVALID, "H and W must be multiples of 8");
    const std::string sig = "enc2d:" + std::to_string(N) + "x" + std::to_string(H) + "x" + std::to_string(W);
    run_sized(u, sig, stream, [&](Ctx& c) {
      FmtScope fs(c, c.enc_fmt);
      void* img16 = c.alloc16((long long)N * H * W * 8);
      if (!c.dry)
        op_check(c, launch_nchw_to_nhwc(img, nullptr, 0.f, 1.f, 0.f, N, c.cfg.vae_in_channels, H, W, 8, img16, c.fmt,
                                        c.stream), "image in");
      vae_encode(c, "vae2d.", img16, N, H, W, out_scale, lat);
    });
  });
}

int ug_vae2d_decode(ug_ctx* u, const float* lat, int N, int h, int w, float* img, unsigned char* normals_u8,
                    void* stream) {
  return guard([&] {
    UG_CHECK(u && lat && (img || normals_u8), UG_ERR_INVALID, "null argument");
    UG_CHECK(u->c.finalized, UG_ERR_STATE, "ug_ctx_finalize must precede ug_vae2d_decode");
    UG_CHECK(N >= 1 && h >= 1 && w >= 1, UG_ERR_INVALID, "bad shape");
    const std::string sig = "dec2d:" + std::to_string(N) + "x" + std::to_string(h) + "x" + std::to_string(w);
    run_sized(u, sig, stream, [&](Ctx& c) {
      void* z16 = c.alloc16((long long)N * h * w * 8);
      if (!c.dry)
        op_check(c, launch_nchw_to_nhwc(lat, nullptr, 0.f, 1.0f / c.cfg.vae_scaling_factor, 0.f, N,
                                        c.cfg.vae_latent_channels, h, w, 8, z16, c.fmt, c.stream), "latents in");
      vae2d_decode(c, z16, N, h, w, img, normals_u8);
    });
  });
}

// ---- host-only schedule tables (no device needed): what the step loops above iterate over
int ug_karras_schedule(const ug_model_cfg* cfg, int steps, double* sigmas, double* timesteps, double* init_noise_sigma) {
  return guard([&] {
    UG_CHECK(cfg && sigmas && steps >= 1 && steps <= 1000, UG_ERR_INVALID, "bad argument");
    const std::vector<double> sig = karras_sigmas(*cfg, steps);
    for (int i = 0; i <= steps; ++i) sigmas[i] = sig[i];
    if (timesteps)
      for (int i = 0; i < steps; ++i) timesteps[i] = 0.25 * std::log(sig[i]);
    if (init_noise_sigma) *init_noise_sigma = std::sqrt(sig[0] * sig[0] + 1.0);
  });
}

int ug_ddim_schedule(const ug_unet2d_cfg* cfg, int steps, int t_start, int* timesteps, double* c_x0, double* c_x) {
  return guard([&] {
    UG_CHECK(cfg && timesteps && steps >= 1 && steps <= cfg->num_train_timesteps && t_start < cfg->num_train_timesteps,
             UG_ERR_INVALID, "bad argument");
    const int NT = cfg->num_train_timesteps;
    const std::vector<double> ac = ddim_alphas_cumprod(NT, (double)cfg->beta_start, (double)cfg->beta_end);
    const std::vector<int> ts = ddim_timesteps(NT, steps, t_start);
    for (int i = 0; i < steps; ++i) {
      const int t = ts[i], tp = i + 1 < steps ? ts[i + 1] : -1;
      UG_CHECK(t >= 0 && t < NT, UG_ERR_INVALID, "timestep out of range");
      timesteps[i] = t;
      const double a_t = ac[t], a_p = tp >= 0 ? ac[tp] : 1.0;
      const double cx = std::sqrt((1.0 - a_p) / (1.0 - a_t));
      if (c_x) c_x[i] = cx;
      if (c_x0) c_x0[i] = std::sqrt(a_p) - std::sqrt(a_t) * cx;
    }
  });
}

long long ug_ctx_launch_count(ug_ctx* u, int reset) {
  if (!u) return -1;
  const long long n = u->c.launches;
  if (reset) u->c.launches = 0;
  return n;
}
long long ug_ctx_workspace_bytes(ug_ctx* u) { return u ? (long long)u->c.ws.cap : -1; }
long long ug_ctx_graph_count(ug_ctx* u) {
  if (!u) return -1;
  long long n = 0;
  for (const auto& kv : u->graphs) n += kv.second.exec != nullptr;
  return n;
}

int ug_ctx_profile(ug_ctx* u, int enable) {
  return guard([&] {
    UG_CHECK(u, UG_ERR_INVALID, "null ctx");
    u->c.profile = enable != 0;
    u->c.profile_shapes = enable == 2;
    u->c.prof.clear();
  });
}

// Aggregates the launches recorded since ug_ctx_profile(ctx, 1) by kernel name.  Writes up to `cap`
// rows: names (64 bytes each, NUL padded), launches, milliseconds, algorithmic flops, algorithmic bytes.
int ug_ctx_profile_read(ug_ctx* u, int cap, char* names, long long* counts, double* ms, double* flops,
                        double* bytes) {
  try {
    if (!u) return UG_ERR_INVALID;
    Ctx& c = u->c;
    UG_CUDA(cudaSetDevice(c.device));
    UG_CUDA(cudaDeviceSynchronize());
    std::vector<std::string> order;
    std::unordered_map<std::string, int> idx;
    int n = 0;
    for (size_t i = 1; i < c.prof.size(); ++i) {      // record 0 is the opening marker
      const Ctx::ProfRec& r = c.prof[i];
      if (std::strcmp(r.name, "(start)") == 0) continue;   // time origin of an API call
      float t = 0.f;
      UG_CUDA(cudaEventElapsedTime(&t, c.prof[i - 1].ev, r.ev));
      auto it = idx.find(r.name);
      int k;
      if (it == idx.end()) {
        if (n >= cap) continue;
        k = n++;
        idx[r.name] = k;
        std::memset(names + 64 * k, 0, 64);
        std::strncpy(names + 64 * k, r.name, 63);
        counts[k] = 0; ms[k] = 0; flops[k] = 0; bytes[k] = 0;
      } else {
        k = it->second;
      }
      counts[k] += 1; ms[k] += t; flops[k] += r.flops; bytes[k] += r.bytes;
    }
    g_err.clear();
    return n;
  } catch (const std::exception& e) {
    g_err = e.what();
    return UG_ERR_CUDA;
  }
}

// ------------------------------------------------------------------ single ops
int ug_op_linear(int dtype, const void* x, long long M, int K, const void* W, int N, const float* bias,
                 const void* res, int geglu, int out_fp32, void* y, void* stream) {
  return guard([&] {
    ug_ctx* u = scratch_ctx(dtype);
    Epi e;
    const int nout = geglu ? N / 2 : N;
    e.out = y; e.ldc = nout; e.out_fp32 = out_fp32; e.bias = bias; e.res = res; e.ldr = nout; e.geglu = geglu;
    // sized run: a launch the tile picker splits along K needs workspace for its fp32 partials
    const std::string sig = "lin:" + std::to_string(M) + ":" + std::to_string(K) + ":" + std::to_string(N) + ":" +
                            std::to_string(geglu) + std::to_string(out_fp32) + (res ? "r" : "");
    run_sized(u, sig, stream, [&](Ctx& c) { op_linear(c, x, M, K, K, W, N, e); });
  });
}

int ug_op_linear_blend(int dtype, const void* x, long long M, int K, const void* W, int N, const float* bias,
                       const void* res, const void* blend, float alpha, void* y, void* stream) {
  return guard([&] {
    UG_CHECK(x && W && blend && y, UG_ERR_INVALID, "null argument");
    ug_ctx* u = scratch_ctx(dtype);
    Epi e;
    e.out = y; e.ldc = N; e.bias = bias; e.res = res; e.ldr = N; e.blend = blend; e.ldb = N; e.alpha = alpha;
    const std::string sig = "linb:" + std::to_string(M) + ":" + std::to_string(K) + ":" + std::to_string(N) + (res ? "r" : "");
    run_sized(u, sig, stream, [&](Ctx& c) { op_linear(c, x, M, K, K, W, N, e); });
  });
}

int ug_op_conv3x3(int dtype, const void* x, int Nf, int H, int W, int C, const void* Wt, int Cout, int stride,
                  int asym_pad, const float* bias, const void* res, void* y, void* stream) {
  return guard([&] {
    ug_ctx* u = scratch_ctx(dtype);
    Epi e;
    e.out = y; e.ldc = Cout; e.bias = bias; e.res = res; e.ldr = Cout;
    const std::string sig = "conv:" + std::to_string(Nf) + ":" + std::to_string(H) + ":" + std::to_string(W) + ":" +
                            std::to_string(C) + ":" + std::to_string(Cout) + ":" + std::to_string(stride) + (res ? "r" : "");
    run_sized(u, sig, stream, [&](Ctx& c) { op_conv3x3(c, x, Nf, H, W, C, Wt, Cout, stride, asym_pad, e); });
  });
}

int ug_op_tconv3(int dtype, const void* x, int T, long long P, int C, const void* Wt, int Cout, int chunk,
                 const float* bias, const void* res, const void* blend, float alpha, void* y, void* stream) {
  return guard([&] {
    ug_ctx* u = scratch_ctx(dtype);
    Epi e;
    e.out = y; e.ldc = Cout; e.bias = bias; e.res = res; e.ldr = Cout; e.blend = blend; e.ldb = Cout; e.alpha = alpha;
    const std::string sig = "tconv:" + std::to_string(T) + ":" + std::to_string(P) + ":" + std::to_string(C) + ":" +
                            std::to_string(Cout) + ":" + std::to_string(chunk) + (res ? "r" : "") + (blend ? "b" : "");
    run_sized(u, sig, stream, [&](Ctx& c) { op_tconv3(c, x, T, P, C, Wt, Cout, chunk, e); });
  });
}

int ug_op_groupnorm(int dtype, const void* x1, int C1, const void* x2, int C2, long long rows,
                    long long rows_per_set, int groups, const float* gamma, const float* beta, float eps, int silu,
                    void* y, void* stream) {
  return guard([&] {
    ug_ctx* u = scratch_ctx(dtype);
    u->c.cfg.norm_groups = groups;
    run_sized(u, "gn:" + std::to_string(rows / rows_per_set) + ":" + std::to_string(groups), stream, [&](Ctx& c) {
      op_gn(c, x1, C1, x2, C2, rows, rows_per_set, gamma, beta, eps, silu, y);
    });
  });
}

int ug_op_layernorm(int dtype, const void* x, long long rows, int C, const float* gamma, const float* beta,
                    float eps, const float* add, int add_div, void* y, void* stream) {
  return guard([&] {
    ug_ctx* u = scratch_ctx(dtype);
    u->c.stream = reinterpret_cast<cudaStream_t>(stream);
    op_layernorm(u->c, x, rows, C, gamma, beta, eps, add, add_div, y);
  });
}

int ug_op_ln_linear(int dtype, const void* x0, int K0, const void* W0, const float* bias0, const void* res0, void* x,
                    long long M, int K, const float* gamma, const float* beta, float eps, const void* W, int N,
                    const float* bias, int geglu, void* y, void* stream) {
  return guard([&] {
    UG_CHECK(x && gamma && beta && W && y, UG_ERR_INVALID, "null argument");
    ug_ctx* u = scratch_ctx(dtype);
    Ctx& cc = u->c;
    UG_CUDA(cudaSetDevice(cc.device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // what ug_ctx_finalize does once per (LayerNorm, linear) pair: gamma into the weights, beta into the bias
    void* wf = nullptr; float* cs = nullptr; float* bo = nullptr; float2* stat = nullptr;
    const int cap = 4;
    UG_CUDA(cudaMalloc(&wf, (size_t)N * K * 2));
    UG_CUDA(cudaMalloc(&cs, (size_t)N * 4));
    UG_CUDA(cudaMalloc(&bo, (size_t)N * 4));
    UG_CUDA(cudaMalloc(&stat, (size_t)M * cap * sizeof(float2)));
    auto cleanup = [&] { cudaStreamSynchronize(st); cudaFree(wf); cudaFree(cs); cudaFree(bo); cudaFree(stat); };
    try {
      UG_CHECK(launch_ln_fold_weights(W, gamma, beta, bias, wf, cs, bo, N, K, cc.fmt, st) == 0, UG_ERR_CUDA, "ln_fold_weights");
      const int nout = geglu ? N / 2 : N;
      const std::string sig = "lnlin:" + std::to_string(M) + ":" + std::to_string(K0) + ":" + std::to_string(K) + ":" +
                              std::to_string(N) + ":" + std::to_string(geglu) + (x0 ? "p" : "") + (res0 ? "r" : "");
      run_sized(u, sig, stream, [&](Ctx& c) {
        int parts = 0;
        if (x0 != nullptr) {              // the rows come out of a GEMM whose epilogue leaves their statistics behind
          Epi p;
          p.out = x; p.ldc = K; p.bias = bias0; p.res = res0; p.ldr = K;
          p.stat_out = stat; p.stat_parts = &parts; p.stat_cap = cap; p.stat_eps = eps;
          op_linear(c, x0, M, K0, K0, W0, K, p);
        } else {
          op_row_stats(c, x, M, K, eps, stat);
        }
        Epi e;
        e.out = y; e.ldc = nout; e.geglu = geglu; e.bias = bo;
        e.ln_stat = stat; e.ln_parts = parts; e.ln_inv_c = 1.0f / (float)K; e.ln_eps = eps; e.ln_colsum = cs;
        op_linear(c, x, M, K, K, wf, N, e);
      });
    } catch (...) { cleanup(); throw; }
    cleanup();
  });
}

int ug_op_spatial_attention(int dtype, const void* qkv, int F, int N, int C, int dh, void* y, void* stream) {
  return guard([&] {
    ug_ctx* u = scratch_ctx(dtype);
    const std::string sig = "attn:" + std::to_string(F) + ":" + std::to_string(N) + ":" + std::to_string(C) + ":" +
                            std::to_string(dh);
    run_sized(u, sig, stream, [&](Ctx& c) { op_spatial_attention(c, qkv, F, N, C, dh, y); });
  });
}

int ug_op_temporal_attention(int dtype, const void* qkv, int T, long long P, int C, void* y, void* stream) {
  return guard([&] {
    ug_ctx* u = scratch_ctx(dtype);
    u->c.stream = reinterpret_cast<cudaStream_t>(stream);
    op_temporal_attention(u->c, qkv, y, T, P, C);
  });
}

int ug_op_cross_attention(int dtype, const void* q, const void* kv, int F, int N, int C, int Lk, int kv_per_frame,
                          void* y, void* stream) {
  return guard([&] {
    ug_ctx* u = scratch_ctx(dtype);
    u->c.stream = reinterpret_cast<cudaStream_t>(stream);
    op_cross_attention(u->c, q, C, kv, y, F, N, C, Lk, kv_per_frame);
  });
}

}  // extern "C"

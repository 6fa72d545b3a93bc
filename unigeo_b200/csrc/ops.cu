// Op wrappers: turn layer-level calls into tensor maps + tapgemm launches / kernel launches.
#include "ctx.cuh"
#include "fmha.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace ug {

// ------------------------------------------------------------------ Ctx
void* Ctx::dmalloc(size_t bytes) {
  void* p = nullptr;
  UG_CUDA(cudaMalloc(&p, bytes < 256 ? 256 : bytes));
  owned.push_back(p);
  ++ptr_epoch;
  return p;
}
const Weight& Ctx::W(const std::string& key) const {
  auto it = weights.find(key);
  if (it == weights.end()) throw UgError(UG_ERR_WEIGHT, "missing weight: " + key);
  return it->second;
}
const float* Ctx::F(const std::string& key) const {
  const Weight& w = W(key);
  UG_CHECK(w.is_f32, UG_ERR_WEIGHT, "weight is not an fp32 vector: " + key);
  return reinterpret_cast<const float*>(w.p);
}
const void* Ctx::M(const std::string& key) const {
  const Weight& w = W(key);
  UG_CHECK(!w.is_f32, UG_ERR_WEIGHT, "weight is not a 16-bit matrix: " + key);
  return w.p;
}
void Ctx::ensure_workspace(size_t bytes) {
  if (bytes <= ws.cap) return;
  if (ws.base) {
    UG_CUDA(cudaDeviceSynchronize());
    UG_CUDA(cudaFree(ws.base));
    ws.base = nullptr;
    ws.cap = 0;
  }
  void* p = nullptr;
  UG_CUDA(cudaMalloc(&p, bytes));
  ws.base = reinterpret_cast<char*>(p);
  ws.cap = bytes;
  ++ptr_epoch;
}

void Ctx::prof_mark(const char* name, double flops, double bytes) {
  cudaEvent_t ev;
  if (ev_pool.size() > prof.size()) {
    ev = ev_pool[prof.size()];
  } else {
    UG_CUDA(cudaEventCreate(&ev));
    ev_pool.push_back(ev);
  }
  UG_CUDA(cudaEventRecord(ev, stream));
  prof.push_back({name, flops, bytes, ev});
}

// profile row name: `what` or, in by-shape mode, `what` + " " + shape (interned for the run)
static const char* prof_name(Ctx& c, const char* what, const std::string& shape) {
  if (!(c.profile && c.profile_shapes)) return what;
  return c.prof_names.insert(std::string(what) + " " + shape).first->c_str();
}

void op_check(Ctx& c, int err, const char* what, double flops, double bytes) {
  if (err != 0)
    throw UgError(UG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString((cudaError_t)err) + " (" +
                                   std::to_string(err) + ")");
  c.launches++;
  if (c.profile) c.prof_mark(what, flops, bytes);
}

// ------------------------------------------------------------------ tapgemm plumbing
namespace {

void fill_epi(TapGemmArgs& a, const Epi& e, int fmt) {
  a.fmt = fmt;
  a.out = e.out;
  a.ldc = e.ldc;
  a.out_fp32 = e.out_fp32;
  a.bias = e.bias;
  a.fbias = e.fbias;
  a.fbias_ld = e.fbias_ld;
  a.fbias_div = e.fbias_div > 0 ? e.fbias_div : 1;
  a.res = e.res;
  a.ldr = e.ldr;
  a.blend = e.blend;
  a.ldb = e.ldb;
  a.alpha = e.alpha;
  a.scale = e.scale;
  a.geglu = e.geglu;
  a.act = e.act;
  a.ln_stat = e.ln_stat;
  a.ln_parts = e.ln_parts;
  a.ln_inv_c = e.ln_inv_c;
  a.ln_eps = e.ln_eps;
  a.ln_colsum = e.ln_colsum;
  a.stat_out = nullptr;          // set by op_linear once the tile width (= partial count) is known
}

void base_args(TapGemmArgs& a) {
  std::memset(&a, 0, sizeof(a));
  a.dim_x = a.dim_y = a.dim_n = a.dim_z0 = a.dim_z1 = 5;
  a.zdiv = 1;
  a.tiles_x = a.tiles_y = a.tiles_n = 1;
  a.bw = a.bh = a.bn = 1;
  a.W = a.H = a.N = 1;
  a.num_taps = 1;
  a.kchunks = 1;
  a.scale = 1.f;
  a.fbias_div = 1;
}

// A: rank-5 map (unused dims = 1).  dims[0] = channels.
void make_a_map(CUtensorMap* m, const void* p, int fmt, const unsigned long long dims[5],
                const unsigned long long strides_bytes[4], int box1, int box2, int box3, int box4,
                bool chunk32 = false) {
  TmapDesc d;
  d.ptr = p;
  d.elem_fmt = fmt;
  d.rank = 5;
  for (int i = 0; i < 5; ++i) d.dims[i] = dims[i];
  for (int i = 0; i < 4; ++i) d.strides[i] = strides_bytes[i];
  d.box[0] = chunk32 ? 32 : 64;            // 32-column SWIZZLE_64B chunks: the residual-by-TMA epilogue
  d.swizzle64 = chunk32 ? 1 : 0;
  d.box[1] = box1; d.box[2] = box2; d.box[3] = box3; d.box[4] = box4;
  int r = encode_tmap(m, d);
  if (r != 0) throw UgError(UG_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: " + std::to_string(r));
}
// B: rank-2 map [rows][cols], box (64 cols, box_rows)
void make_b_map(CUtensorMap* m, const void* p, int fmt, unsigned long long cols, unsigned long long rows,
                unsigned long long row_stride_bytes, int box_rows) {
  TmapDesc d;
  d.ptr = p;
  d.elem_fmt = fmt;
  d.rank = 2;
  d.dims[0] = cols; d.dims[1] = rows;
  d.strides[0] = row_stride_bytes;
  d.box[0] = 64; d.box[1] = box_rows;
  int r = encode_tmap(m, d);
  if (r != 0) throw UgError(UG_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: " + std::to_string(r));
}

// TMA tile stores need a 16-bit output whose rows start on 16-byte boundaries
inline bool can_tma_store(const Epi& e) {
  static const bool off = getenv("UG_NO_TMA_STORE") != nullptr;
  return !off && !e.out_fp32 && (e.ldc % 8) == 0 && (reinterpret_cast<uintptr_t>(e.out) % 16) == 0;
}

// residual tiles by TMA (in-place 32-column chunks): 16-bit residual with 16-byte aligned rows, N % 32 == 0
inline bool can_tma_res(const Epi& e, int n_out) {
  static const bool off = getenv("UG_NO_TMA_RES") != nullptr;
  return !off && can_tma_store(e) && e.res != nullptr && !e.geglu && (n_out % 32) == 0 && (e.ldr % 8) == 0 &&
         (reinterpret_cast<uintptr_t>(e.res) % 16) == 0;
}

// split-K is available to launches whose epilogue the reduce kernel can reproduce (no GEGLU, N % 4 == 0)
inline bool can_split(const Epi& e, int n_out) { return !e.geglu && (n_out & 3) == 0; }

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline int imin(int a, int b) { return a < b ? a : b; }

// algorithmic work of one tapgemm launch: 2*M*N*K flops over the REAL (unpadded) extents;
// bytes = A read once + output written once (16-bit), weights ignored (SURVEY.md §8 convention)
void launch(Ctx& c, const CUtensorMap& ma, const CUtensorMap& mb, const TapGemmArgs& a, int batch,
            const char* what, long long k_real, const CUtensorMap* mc = nullptr, const CUtensorMap* mr = nullptr,
            const CUtensorMap* mbl = nullptr);

// Split-K launch: S partial GEMMs into fp32 workspace slabs + one reduce pass carrying the epilogue `e`.
// `a` holds the geometry (tiling, taps, n_total, bn_tile, ctas); its epilogue fields are overwritten here.
void launch_split(Ctx& c, const CUtensorMap& ma, const CUtensorMap& mb, TapGemmArgs a, int S, const Epi& e,
                  const char* what, long long k_real) {
  const long long M = (long long)a.W * a.H * a.N;
  const size_t mk = c.ws.mark();
  float* part = c.allocf((long long)S * M * a.n_total);
  Epi pe;
  pe.out = part; pe.ldc = a.n_total; pe.out_fp32 = 1;
  fill_epi(a, pe, c.fmt);
  a.fbias_uniform = 0;
  a.tma_store = 0;
  a.res_tma = 0;
  a.blend_tma = 0;
  a.ksplit = S;
  a.out_z1stride = M * a.n_total;
  a.out_z0stride = 0;
  launch(c, ma, mb, a, S, what, k_real);
  op_check(c, launch_splitk_reduce(part, S, M, a.n_total, e.bias, e.fbias, e.fbias_ld, e.fbias_div, e.res, e.ldr, e.blend,
                                   e.ldb, e.alpha, e.scale, e.act, e.out, e.ldc, e.out_fp32, c.fmt, c.stream),
           "splitk_reduce", 0.0, (double)M * a.n_total * (4.0 * S + 2.0));
  c.ws.release(mk);
}

void launch(Ctx& c, const CUtensorMap& ma, const CUtensorMap& mb, const TapGemmArgs& a, int batch,
            const char* what, long long k_real, const CUtensorMap* mc, const CUtensorMap* mr, const CUtensorMap* mbl) {
  const double M = (double)a.W * a.H * a.N * (a.ksplit > 1 ? 1 : batch);
  const double flops = 2.0 * M * a.n_total * (double)k_real * a.num_taps;
  const double bytes = M * ((double)k_real + (a.geglu ? a.n_total / 2 : a.n_total)) * 2.0;
  const char* name = what;
  if (c.profile && c.profile_shapes) {   // per-shape rows in the profile table
    const std::string full = std::string(what) + " M" + std::to_string((long long)M) + " N" + std::to_string(a.n_total) +
                             " K" + std::to_string(k_real * a.num_taps) + " bn" + std::to_string(a.bn_tile) + "x" +
                             std::to_string(a.ctas) + (a.res ? " +res" : "") + (a.blend ? " +blend" : "") +
                             (a.fbias ? " +fbias" : "") + (a.tma_store ? "" : " direct") + (a.res_tma ? " rtma" : "") + (a.blend_tma ? " btma" : "") +
                             (a.ln_stat ? " lnfold" : "") + (a.stat_out ? " +stats" : "") +
                             (a.ksplit > 1 ? " splitk" + std::to_string(a.ksplit) : "");
    name = c.prof_names.insert(full).first->c_str();
  }
  op_check(c, launch_tapgemm(ma, mb, mc, a, batch, c.stream, mr, mbl), name, flops, bytes);
}

}  // namespace

void op_linear(Ctx& c, const void* x, long long M, int K, long long ldx, const void* Wm, int N, const Epi& e) {
  UG_CHECK((K % 8) == 0 && (ldx % 8) == 0, UG_ERR_INVALID, "linear: K and ldx must be multiples of 8");
  TapGemmArgs a;
  base_args(a);
  a.dim_x = 1;
  a.bw = 128;
  a.W = (int)M;
  a.tiles_x = cdiv(M, 128);
  a.kchunks = cdiv(K, 64);
  a.n_total = N;
  fill_epi(a, e, c.fmt);
  a.fbias_uniform = (e.fbias != nullptr && (a.fbias_div % 128) == 0) ? 1 : 0;   // tiles = 128 consecutive tokens
  a.tma_store = can_tma_store(e);
  a.res_tma = can_tma_res(e, N) ? 1 : 0;
  // the AlphaBlender input by TMA as well when it is a different tensor than the residual (UG_NO_TMA_BLEND: per-thread loads)
  static const bool no_tma_blend = getenv("UG_NO_TMA_BLEND") != nullptr;
  a.blend_tma = (a.res_tma && !no_tma_blend && e.blend != nullptr && !(e.blend == e.res && e.ldb == e.ldr) &&
                 (e.ldb % 8) == 0 && (reinterpret_cast<uintptr_t>(e.blend) % 16) == 0) ? 1 : 0;
  int ksplit = 1;
  // a LayerNorm-folding launch keeps its epilogue (the split-K reduce pass does not know the fold)
  a.bn_tile = tapgemm_pick_tile(a, 1, &a.ctas, (can_split(e, N) && e.ln_stat == nullptr) ? &ksplit : nullptr);
  // row statistics for the next GEMM's folded LayerNorm: out of this launch's epilogue when its partials fit the
  // caller's buffer and nothing splits K, else one row_stats pass over the finished output
  static const bool no_epi_stats = getenv("UG_NO_EPI_STATS") != nullptr;
  const int parts = 2 * cdiv(N, a.bn_tile);
  const bool epi_stats = e.stat_out != nullptr && !no_epi_stats && ksplit == 1 && parts <= e.stat_cap && !e.geglu &&
                         !e.out_fp32;
  if (e.stat_out != nullptr) {
    UG_CHECK(e.stat_parts != nullptr && e.stat_cap >= 1 && e.ldc == N, UG_ERR_INVALID, "linear: stat_out needs stat_parts / dense rows");
    *e.stat_parts = epi_stats ? parts : 0;
  }
  if (c.dry) {                                    // size-only pass: the split-K partials are the op's only workspace
    if (ksplit > 1) { const size_t mk = c.ws.mark(); c.allocf((long long)ksplit * M * N); c.ws.release(mk); }
    return;
  }
  if (epi_stats) a.stat_out = e.stat_out;
  const int bn = a.bn_tile / a.ctas;
  CUtensorMap ma, mb, mc, mr, mbl;
  unsigned long long dims[5] = {(unsigned long long)K, (unsigned long long)M, 1, 1, 1};
  unsigned long long st[4] = {(unsigned long long)ldx * 2, (unsigned long long)ldx * 2 * M,
                              (unsigned long long)ldx * 2 * M, (unsigned long long)ldx * 2 * M};
  make_a_map(&ma, x, c.fmt, dims, st, 128, 1, 1, 1);
  make_b_map(&mb, Wm, c.fmt, K, N, (unsigned long long)K * 2, bn);
  if (ksplit > 1) {
    launch_split(c, ma, mb, a, ksplit, e, "tapgemm.linear", K);
    if (e.stat_out != nullptr) op_row_stats(c, e.out, M, N, e.stat_eps, e.stat_out);
    return;
  }
  if (a.tma_store) {
    const unsigned long long rb = (unsigned long long)e.ldc * 2;
    unsigned long long od[5] = {(unsigned long long)(e.geglu ? N / 2 : N), (unsigned long long)M, 1, 1, 1};
    unsigned long long os[4] = {rb, rb * M, rb * M, rb * M};
    make_a_map(&mc, e.out, c.fmt, od, os, 128, 1, 1, 1, a.res_tma);
    if (a.res_tma) {
      const unsigned long long rr = (unsigned long long)e.ldr * 2;
      unsigned long long rs[4] = {rr, rr * M, rr * M, rr * M};
      make_a_map(&mr, e.res, c.fmt, od, rs, 128, 1, 1, 1, true);
      if (a.blend_tma) {
        const unsigned long long br = (unsigned long long)e.ldb * 2;
        unsigned long long bs[4] = {br, br * M, br * M, br * M};
        make_a_map(&mbl, e.blend, c.fmt, od, bs, 128, 1, 1, 1, true);
      }
    }
  }
  launch(c, ma, mb, a, 1, e.geglu ? "tapgemm.linear_geglu" : "tapgemm.linear", K, a.tma_store ? &mc : nullptr,
         a.res_tma ? &mr : nullptr, a.blend_tma ? &mbl : nullptr);
  if (e.stat_out != nullptr && !epi_stats) op_row_stats(c, e.out, M, N, e.stat_eps, e.stat_out);
}

void op_conv3x3(Ctx& c, const void* x, int Nf, int H, int W, int C, const void* Wm, int Cout, int stride,
                int asym, const Epi& e) {
  UG_CHECK((C % 8) == 0, UG_ERR_INVALID, "conv3x3: C must be a multiple of 8");
  TapGemmArgs a;
  base_args(a);
  const int Ho = H / stride, Wo = W / stride;
  a.W = Wo; a.H = Ho; a.N = Nf;
  a.bw = imin(Wo, 128);
  // 128-row tile = bw x bh x bn pixels: pick the (rows, frames) split that covers the clip with the fewest
  // tiles (e.g. 16x12 frames: 4 rows x 2 frames -> 39 tiles instead of 8 rows x 1 frame -> 50)
  {
    const int rows_left = 128 / a.bw;
    long long best = -1;
    for (int bh = 1; bh <= rows_left && bh <= Ho; ++bh) {
      const int bn = imin(Nf, rows_left / bh);
      const long long tiles = (long long)cdiv(Ho, bh) * cdiv(Nf, bn);
      if (best < 0 || tiles < best || (tiles == best && bh > a.bh)) {
        best = tiles;
        a.bh = bh;
        a.bn = bn;
      }
    }
  }
  a.tiles_x = cdiv(Wo, a.bw);
  a.tiles_y = cdiv(Ho, a.bh);
  a.tiles_n = cdiv(Nf, a.bn);
  a.num_taps = 9;
  a.kchunks = cdiv(C, 64);
  a.b_tap_rows = Cout;
  a.n_total = Cout;
  fill_epi(a, e, c.fmt);
  a.tma_store = can_tma_store(e);
  a.res_tma = can_tma_res(e, Cout) ? 1 : 0;
  int ksplit = 1;
  a.bn_tile = tapgemm_pick_tile(a, 1, &a.ctas, can_split(e, Cout) ? &ksplit : nullptr);
  if (c.dry) {
    if (ksplit > 1) { const size_t mk = c.ws.mark(); c.allocf((long long)ksplit * Nf * Ho * Wo * Cout); c.ws.release(mk); }
    return;
  }
  const int bn = a.bn_tile / a.ctas;
  CUtensorMap ma, mb, mc, mr;
  if (a.tma_store && ksplit == 1) {
    const unsigned long long rb = (unsigned long long)e.ldc * 2;
    unsigned long long od[5] = {(unsigned long long)Cout, (unsigned long long)Wo, (unsigned long long)Ho,
                                (unsigned long long)Nf, 1};
    unsigned long long os[4] = {rb, rb * Wo, rb * Wo * Ho, rb * Wo * Ho * Nf};
    make_a_map(&mc, e.out, c.fmt, od, os, a.bw, a.bh, a.bn, 1, a.res_tma);
    if (a.res_tma) {
      const unsigned long long rr = (unsigned long long)e.ldr * 2;
      unsigned long long rs[4] = {rr, rr * Wo, rr * Wo * Ho, rr * Wo * Ho * Nf};
      make_a_map(&mr, e.res, c.fmt, od, rs, a.bw, a.bh, a.bn, 1, true);
    }
  }
  const unsigned long long rowb = (unsigned long long)C * 2;
  if (stride == 1) {
    a.dim_x = 1; a.dim_y = 2; a.dim_n = 3;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        int* o = a.tap_off[ky * 3 + kx];
        o[1] = kx - 1;
        o[2] = ky - 1;
      }
    unsigned long long dims[5] = {(unsigned long long)C, (unsigned long long)W, (unsigned long long)H,
                                  (unsigned long long)Nf, 1};
    unsigned long long st[4] = {rowb, rowb * W, rowb * W * H, rowb * W * H * Nf};
    make_a_map(&ma, x, c.fmt, dims, st, a.bw, a.bh, a.bn, 1);
  } else {
    UG_CHECK(stride == 2 && (H % 2) == 0 && (W % 2) == 0, UG_ERR_INVALID, "conv3x3: stride 2 needs even H, W");
    // space-to-depth VIEW of x (no copy): dims (2C [px,c], W/2, py, H/2, N)
    a.dim_x = 1; a.dim_y = 3; a.dim_n = 4;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        int* o = a.tap_off[ky * 3 + kx];
        const int dy = asym ? ky : ky - 1;       // input row = 2*oy + dy
        const int dx = asym ? kx : kx - 1;
        const int py = ((dy % 2) + 2) % 2, px = ((dx % 2) + 2) % 2;
        o[0] = px * C;
        o[1] = (dx - px) / 2;
        o[2] = py;
        o[3] = (dy - py) / 2;
      }
    unsigned long long dims[5] = {(unsigned long long)2 * C, (unsigned long long)W / 2, 2,
                                  (unsigned long long)H / 2, (unsigned long long)Nf};
    unsigned long long st[4] = {rowb * 2, rowb * W, rowb * W * 2, rowb * W * H};
    make_a_map(&ma, x, c.fmt, dims, st, a.bw, 1, a.bh, a.bn);
  }
  make_b_map(&mb, Wm, c.fmt, C, (unsigned long long)9 * Cout, rowb, bn);
  if (ksplit > 1) {
    launch_split(c, ma, mb, a, ksplit, e, "tapgemm.conv3x3", C);
    return;
  }
  launch(c, ma, mb, a, 1, "tapgemm.conv3x3", C, a.tma_store ? &mc : nullptr, a.res_tma ? &mr : nullptr);
}

void op_tconv3(Ctx& c, const void* x, int T, long long P, int C, const void* Wm, int Cout, int chunk,
               const Epi& e) {
  UG_CHECK((C % 8) == 0, UG_ERR_INVALID, "tconv3: C must be a multiple of 8");
  const unsigned long long rowb = (unsigned long long)C * 2;
  // chunks are independent zero-padded clips (VAE decode): run one launch per chunk so that
  // the TMA bounds ARE the chunk bounds.
  for (int t0 = 0; t0 < T; t0 += chunk) {
    const int Tc = imin(chunk, T - t0);
    TapGemmArgs a;
    base_args(a);
    a.dim_x = 1; a.dim_y = 2;
    a.W = (int)P; a.H = Tc;
    a.bw = (int)(P < 128 ? P : 128);
    a.bh = imin(Tc, 128 / a.bw);
    a.tiles_x = cdiv(P, a.bw);
    a.tiles_y = cdiv(Tc, a.bh);
    a.num_taps = 3;
    a.kchunks = cdiv(C, 64);
    a.b_tap_rows = Cout;
    a.n_total = Cout;
    for (int kt = 0; kt < 3; ++kt) a.tap_off[kt][2] = kt - 1;
    Epi ec = e;
    const long long tok0 = (long long)t0 * P;
    const int osz = e.out_fp32 ? 4 : 2;
    ec.out = reinterpret_cast<char*>(e.out) + tok0 * e.ldc * osz;
    if (e.res) ec.res = reinterpret_cast<const char*>(e.res) + tok0 * e.ldr * 2;
    if (e.blend) ec.blend = reinterpret_cast<const char*>(e.blend) + tok0 * e.ldb * 2;
    if (e.fbias) ec.fbias = e.fbias + (tok0 / ec.fbias_div) * e.fbias_ld;
    fill_epi(a, ec, c.fmt);
    a.tma_store = can_tma_store(ec);
    a.res_tma = can_tma_res(ec, Cout) ? 1 : 0;
    int ksplit = 1;
    a.bn_tile = tapgemm_pick_tile(a, 1, &a.ctas, can_split(ec, Cout) ? &ksplit : nullptr);
    if (c.dry) {
      if (ksplit > 1) { const size_t mk = c.ws.mark(); c.allocf((long long)ksplit * Tc * P * Cout); c.ws.release(mk); }
      continue;
    }
    const int bn = a.bn_tile / a.ctas;
    CUtensorMap ma, mb, mc, mr;
    if (a.tma_store && ksplit == 1) {
      const unsigned long long rb = (unsigned long long)ec.ldc * 2;
      unsigned long long od[5] = {(unsigned long long)Cout, (unsigned long long)P, (unsigned long long)Tc, 1, 1};
      unsigned long long os[4] = {rb, rb * P, rb * P * Tc, rb * P * Tc};
      make_a_map(&mc, ec.out, c.fmt, od, os, a.bw, a.bh, 1, 1, a.res_tma);
      if (a.res_tma) {
        const unsigned long long rr = (unsigned long long)ec.ldr * 2;
        unsigned long long rs[4] = {rr, rr * P, rr * P * Tc, rr * P * Tc};
        make_a_map(&mr, ec.res, c.fmt, od, rs, a.bw, a.bh, 1, 1, true);
      }
    }
    unsigned long long dims[5] = {(unsigned long long)C, (unsigned long long)P, (unsigned long long)Tc, 1, 1};
    unsigned long long st[4] = {rowb, rowb * P, rowb * P * Tc, rowb * P * Tc};
    make_a_map(&ma, reinterpret_cast<const char*>(x) + tok0 * rowb, c.fmt, dims, st, a.bw, a.bh, 1, 1);
    make_b_map(&mb, Wm, c.fmt, C, (unsigned long long)3 * Cout, rowb, bn);
    if (ksplit > 1) {
      launch_split(c, ma, mb, a, ksplit, ec, "tapgemm.tconv3", C);
      continue;
    }
    launch(c, ma, mb, a, 1, "tapgemm.tconv3", C, a.tma_store ? &mc : nullptr, a.res_tma ? &mr : nullptr);
  }
}

// v1 attention: S = Q K^T (tapgemm, batched over frame x head) -> row softmax -> O = P V
// (tapgemm with the MN-major B path reading V in place).  Scores live in the workspace.
void op_spatial_attention(Ctx& c, const void* qkv, int F, int N, int C, int dh, void* out, int n_valid, float scale) {
  UG_CHECK((N % 8) == 0 && (dh % 64) == 0 && (C % dh) == 0, UG_ERR_INVALID, "attention: N % 8, head_dim % 64, C % head_dim");
  if (n_valid <= 0) n_valid = N;
  if (scale <= 0.f) scale = 1.0f / sqrtf((float)dh);
  const int heads = C / dh;
  if (dh == 64 && n_valid == N && !c.attn_materialized) {
    // fused flash attention (tcgen05): no score tensor in HBM
    if (c.dry) return;
    TmapDesc d;
    d.ptr = qkv; d.elem_fmt = c.fmt; d.rank = 2;
    d.dims[0] = (unsigned long long)3 * C; d.dims[1] = (unsigned long long)F * N;
    d.strides[0] = (unsigned long long)3 * C * 2;
    d.box[0] = 64; d.box[1] = 128;
    CUtensorMap tm;
    int r = encode_tmap(&tm, d);
    if (r != 0) throw UgError(UG_ERR_CUDA, "cuTensorMapEncodeTiled(qkv) failed: " + std::to_string(r));
    FmhaArgs fa;
    fa.N = N; fa.C = C; fa.heads = heads; fa.F = F;
    fa.scale_log2 = 1.4426950408889634f * scale;
    fa.out = out; fa.fmt = c.fmt;
    op_check(c, launch_fmha_d64(tm, fa, c.stream),
             prof_name(c, "fmha_d64", "F" + std::to_string(F) + " N" + std::to_string(N) + " heads" + std::to_string(heads)),
             4.0 * F * heads * (double)N * N * dh,
             8.0 * F * (double)N * C);
    return;
  }
  const size_t m = c.ws.mark();
  void* S = c.alloc16((long long)F * heads * N * N);
  if (!c.dry) {
    const unsigned long long rowb = (unsigned long long)3 * C * 2;
    {  // S[z][N][N] = Q K^T
      TapGemmArgs a;
      base_args(a);
      a.dim_x = 1;
      a.bw = imin(N, 128);
      a.W = N;
      a.tiles_x = cdiv(N, a.bw);
      a.kchunks = dh / 64;
      a.zdiv = heads;
      a.dim_z1 = 1; a.a_z1step = N;       // frame -> token rows
      a.dim_z0 = 0; a.a_z0step = dh;      // head  -> q columns
      a.b_c0 = C; a.b_z0_cstep = dh; a.b_z1_rowstep = N;
      a.n_total = N;
      Epi e;
      e.out = S; e.ldc = N;
      fill_epi(a, e, c.fmt);
      a.out_z1stride = (long long)heads * N * N;
      a.out_z0stride = (long long)N * N;
      CUtensorMap ma, mb;
      unsigned long long dims[5] = {(unsigned long long)3 * C, (unsigned long long)F * N, 1, 1, 1};
      unsigned long long st[4] = {rowb, rowb * F * N, rowb * F * N, rowb * F * N};
      make_a_map(&ma, qkv, c.fmt, dims, st, a.bw, 1, 1, 1);
      a.bn_tile = tapgemm_pick_tile(a, F * heads, &a.ctas);
      make_b_map(&mb, qkv, c.fmt, (unsigned long long)3 * C, (unsigned long long)F * N, rowb, a.bn_tile / a.ctas);
      launch(c, ma, mb, a, F * heads, "tapgemm.attn_qk", dh);
    }
    op_check(c, launch_softmax_rows(S, (long long)F * heads * N, N, n_valid, scale, c.fmt, c.stream),
             "softmax_rows", 0.0, 4.0 * F * heads * (double)N * N);
    {  // O[z] = P[z] V[z]
      TapGemmArgs a;
      base_args(a);
      a.dim_x = 1;
      a.bw = imin(N, 128);
      a.W = N;
      a.tiles_x = cdiv(N, a.bw);
      a.kchunks = cdiv(N, 64);
      a.zdiv = heads;
      a.dim_z0 = 2; a.a_z0step = 1;
      a.dim_z1 = 3; a.a_z1step = 1;
      a.b_mn_major = 1;
      a.b_c0 = 2 * C; a.b_z0_cstep = dh; a.b_z1_rowstep = N;
      a.n_total = dh;
      Epi e;
      e.out = out; e.ldc = C;
      fill_epi(a, e, c.fmt);
      a.out_z1stride = (long long)N * C;
      a.out_z0stride = dh;
      CUtensorMap ma, mb;
      const unsigned long long srow = (unsigned long long)N * 2;
      unsigned long long dims[5] = {(unsigned long long)N, (unsigned long long)N, (unsigned long long)heads,
                                    (unsigned long long)F, 1};
      unsigned long long st[4] = {srow, srow * N, srow * N * heads, srow * N * heads * F};
      make_a_map(&ma, S, c.fmt, dims, st, a.bw, 1, 1, 1);
      a.bn_tile = tapgemm_pick_tile(a, F * heads, &a.ctas);   // 64, 1 CTA: MN-major boxes are [64 K][64 N]
      make_b_map(&mb, qkv, c.fmt, (unsigned long long)3 * C, (unsigned long long)F * N, rowb, 64);
      launch(c, ma, mb, a, F * heads, "tapgemm.attn_pv", N);
    }
  }
  c.ws.release(m);
}

// ------------------------------------------------------------------ simple wrappers
void op_gn(Ctx& c, const void* x1, int C1, const void* x2, int C2, long long rows, long long rows_per_set,
           const float* gamma, const float* beta, float eps, int silu, void* y) {
  const int G = c.cfg.norm_groups;
  const long long sets = rows / rows_per_set;
  const size_t m = c.ws.mark();
  float* stats = c.allocf(gn_partial_floats(C1 + C2, rows, rows_per_set, G));
  if (!c.dry) {
    if (!c.gn_counters) {
      c.gn_counters = reinterpret_cast<unsigned int*>(c.dmalloc(3 * kGnMaxSets * sizeof(unsigned int)));
      UG_CUDA(cudaMemset(c.gn_counters, 0, 3 * kGnMaxSets * sizeof(unsigned int)));
    }
    static const bool two_pass = getenv("UG_GN_TWO_PASS") != nullptr;
    const std::string shape = "rows" + std::to_string(rows) + " C" + std::to_string(C1 + C2) + " sets" + std::to_string(sets);
    if (!two_pass) {      // sets that fit one cluster's shared memory: x read once, no global flags
      const int rc = launch_gn_cluster(x1, C1, x2, C2, rows, rows_per_set, G, eps, gamma, beta, silu, y, c.fmt, c.stream);
      if (rc != (int)cudaErrorNotSupported) {
        op_check(c, rc, prof_name(c, "gn_cluster", shape), 0.0, 4.0 * rows * (C1 + C2));
        c.ws.release(m);
        return;
      }
    }
    int r = two_pass ? (int)cudaErrorNotSupported
                     : launch_gn_fused(x1, C1, x2, C2, rows, rows_per_set, G, eps, stats, c.gn_counters, gamma, beta, silu,
                                       y, c.fmt, c.stream);
    if (r == (int)cudaErrorNotSupported) {      // grid cannot be co-resident: two launches
      op_check(c, launch_gn_stats(x1, C1, x2, C2, rows, rows_per_set, G, eps, stats, c.gn_counters, c.fmt, c.stream),
               prof_name(c, "gn_stats", shape), 0.0, 2.0 * rows * (C1 + C2));
      op_check(c, launch_gn_apply(x1, C1, x2, C2, rows, rows_per_set, G, stats, gamma, beta, eps, silu, y, c.fmt,
                                  c.stream),
               prof_name(c, "gn_apply", shape), 0.0, 4.0 * rows * (C1 + C2));
    } else {
      op_check(c, r, prof_name(c, "gn_fused", shape), 0.0, 4.0 * rows * (C1 + C2));
    }
  }
  c.ws.release(m);
}
void op_layernorm(Ctx& c, const void* x, long long rows, int C, const float* g, const float* b, float eps,
                  const float* add, int add_div, void* y) {
  if (c.dry) return;
  op_check(c, launch_layernorm(x, rows, C, g, b, eps, add, add_div, y, c.fmt, c.stream),
           prof_name(c, "layernorm", "rows" + std::to_string(rows) + " C" + std::to_string(C)), 0.0, 4.0 * rows * C);
}
void op_row_stats(Ctx& c, const void* x, long long rows, int C, float eps, float2* stat) {
  if (c.dry) return;
  op_check(c, launch_row_stats(x, rows, C, eps, stat, c.fmt, c.stream),
           prof_name(c, "row_stats", "rows" + std::to_string(rows) + " C" + std::to_string(C)), 0.0, 2.0 * rows * C);
}
void op_temporal_attention(Ctx& c, const void* qkv, void* out, int T, long long P, int C) {
  if (c.dry) return;
  op_check(c, launch_temporal_attention(qkv, out, T, P, C, 0.125f, c.fmt, c.stream),
           prof_name(c, "temporal_attention", "P" + std::to_string(P) + " C" + std::to_string(C)),
           4.0 * T * T * 64.0 * P * (C / 64), 8.0 * T * P * C);
}
void op_cross_attention(Ctx& c, const void* q, int ldq, const void* kv, void* out, int F, int N, int C, int Lk,
                        int kv_per_frame) {
  if (c.dry) return;
  op_check(c, launch_cross_attention(q, ldq, kv, out, F, N, C, Lk, kv_per_frame, 0.125f, c.fmt, c.stream),
           prof_name(c, "cross_attention", "F" + std::to_string(F) + " N" + std::to_string(N) + " C" + std::to_string(C)),
           4.0 * F * (double)N * Lk * C, 4.0 * F * (double)N * C);
}
void op_upsample2x(Ctx& c, const void* x, void* y, int N, int H, int W, int C) {
  if (c.dry) return;
  op_check(c, launch_upsample2x(x, y, N, H, W, C, c.stream), "upsample2x", 0.0, 10.0 * N * H * W * C);
}
void op_concat(Ctx& c, const void* x1, int C1, const void* x2, int C2, long long rows, void* y) {
  if (c.dry) return;
  op_check(c, launch_concat(x1, C1, x2, C2, rows, y, c.stream), "concat", 0.0, 4.0 * rows * (C1 + C2));
}
void op_gemv(Ctx& c, const void* Wm, const float* b, const float* addend, const float* x, float* out, int M,
             int N, int K, int silu_in, int silu_out) {
  if (c.dry) return;
  op_check(c, launch_gemv(Wm, b, addend, x, out, M, N, K, silu_in, silu_out, c.fmt, c.stream), "gemv");
}

}  // namespace ug

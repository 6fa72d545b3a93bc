// tapgemm kernel + launcher (see tapgemm.cuh for what it computes and replaces).
//
// Persistent CTA = 320 threads walking 128 x BN output tiles (CTA pairs: 256 x BN with cta_group::2):
//   warp 0      TMA producer   (one elected lane; A box + B box per stage, mbarrier tx-count)
//   warp 1      TMEM owner + UMMA issuer (one elected lane; tcgen05.mma 128xBNx16, commit -> mbarrier)
//   warps 2..9  epilogue       (tcgen05.ld 32x32b: thread = output row, registers = columns; 4 TMEM lane quarters x
//                               2 column groups), accumulator double-buffered in TMEM
// smem: STAGES x (A 128x64 + B BNx64) 16-bit tiles in the TMA/UMMA SWIZZLE_128B layout + 4 x 16 KB store staging.
#include "tapgemm.cuh"
#include "ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <queue>
#include <vector>

namespace ug {

// Per-role time stamps of the persistent kernel (clock64 of the CTA's SM), compiled in only with -DUG_TAPGEMM_TRACE
// (tools/trace_tapgemm.py builds that variant into its own .so): role 0 producer, 1 UMMA issuer, 2 / 3 epilogue column
// groups; rank 0 of a pair only, first kTraceTiles units of every CTA.  Expands to nothing in the product build.
#ifdef UG_TAPGEMM_TRACE
#define UG_TRACE(role, tile, slot)                                                                             \
  do {                                                                                                         \
    if (a.trace != nullptr && rank == 0 && (tile) < kTraceTiles)                                               \
      a.trace[(((size_t)unit0 * 4 + (role)) * kTraceTiles + (tile)) * 4 + (slot)] = (unsigned long long)clock64(); \
  } while (0)
#else
#define UG_TRACE(role, tile, slot) do {} while (0)
#endif

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kATileBytes = BM * BK * 2;
constexpr int kBTileBytes = 256 * BK * 2;            // room for the widest N tile
constexpr int kStageBytes = kATileBytes + kBTileBytes;
constexpr int kSlabBytes = 128 * 128;                // 128 rows x 64 16-bit columns of finished output
constexpr int kStages = 3;
constexpr int kSmemBytes = kStages * kStageBytes + 4 * kSlabBytes + 256 /*barriers*/ + 2048 /*bias*/;
// CTA-pair mode: each CTA stages 128 rows of A and at most 128 rows (half) of B
constexpr int kStageBytes2 = kATileBytes + 128 * BK * 2;
constexpr int kStages2 = 5;
constexpr int kSmemBytes2 = kStages2 * kStageBytes2 + 4 * kSlabBytes + 256 + 2048;

#ifdef UG_GELU_ERF
// exact (erf) GELU with erf from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, far below 16-bit output
// resolution): 2 MUFU + ~12 FMA-pipe instructions instead of libm erff's ~30
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// Two gates at once on the packed fp32 pipe (FFMA2 / FMUL2 / FADD2, sm_100): the GEGLU epilogue is
// instruction-issue bound at K = 320 (ncu: 30 thread instructions per output element), and the packed
// forms halve the polynomial's share.  Same formula as gelu_erf: returns 0.5 x (1 + erf(x / sqrt 2)).
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  // t = 1 / (1 + 0.3275911 |x| / sqrt 2)   (|.| rides on the scalar FFMA as an operand modifier)
  float2 t;
  t.x = rcp_approx(fmaf(0.23164189f, fabsf(x.x), 1.0f));
  t.y = rcp_approx(fmaf(0.23164189f, fabsf(x.y), 1.0f));
  // q = -(a1 t + a2 t^2 + ... + a5 t^5)  (coefficients negated so that erf = 1 + q e)
  float2 q = __ffma2_rn(t, make_float2(-1.061405429f, -1.061405429f), make_float2(1.453152027f, 1.453152027f));
  q = __ffma2_rn(t, q, make_float2(-1.421413741f, -1.421413741f));
  q = __ffma2_rn(t, q, make_float2(0.284496736f, 0.284496736f));
  q = __ffma2_rn(t, q, make_float2(-0.254829592f, -0.254829592f));
  q = __fmul2_rn(q, t);
  // e = exp(-x^2 / 2) = 2^(-0.5 log2(e) x^2)
  const float2 w = __fmul2_rn(__fmul2_rn(x, x), make_float2(-0.72134752044448170f, -0.72134752044448170f));
  float2 e;
  e.x = ex2_approx(w.x);
  e.y = ex2_approx(w.y);
  const float2 erf_abs = __ffma2_rn(q, e, make_float2(1.0f, 1.0f));
  // 0.5 x (1 + sign(x) erf|.|) = h + |h| erf|.|,  h = x / 2
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return make_float2(fmaf(fabsf(h.x), erf_abs.x, h.x), fmaf(fabsf(h.y), erf_abs.y, h.y));
}
#endif

// GELU through ONE MUFU per element:  x Phi(x),  Phi(x) = (1 + tanh(x (c0 + c1 x^2 + c2 x^4))) / 2  with the three
// coefficients fitted (minimax, here) to the EXACT erf form: max |error| 2.5e-5 over all x -- 19 times tighter than the
// textbook two-coefficient tanh form, and below what rounding the gate to 16 bit first (as the reference's unfused
// proj -> gelu does) costs; MUFU.TANH adds <= 2^-12 x relative to Phi.  x^2 is clamped at 36 so that the quartic never
// turns over (tanh has saturated to 1 - 1e-9 there).  10 instructions per pair (4 FMA-pipe packed, 2 min, 2 MUFU, 2
// packed) against 16 with 4 MUFU for the erf form: the K = 320 GEGLU epilogue is bound by exactly these (trace: 3700 of a
// 5700-clock tile period in the gate math, MUFU 32 clk per pair and warp).  -DUG_GELU_ERF restores the erf form.
__device__ __forceinline__ float2 gelu_tanh2(float2 x) {
  float2 x2 = __fmul2_rn(x, x);
  x2.x = fminf(x2.x, 36.0f);
  x2.y = fminf(x2.y, 36.0f);
  float2 p = __ffma2_rn(x2, make_float2(-0.0003515167918521911f, -0.0003515167918521911f),
                        make_float2(0.03700564429163933f, 0.03700564429163933f));
  p = __ffma2_rn(p, x2, make_float2(0.7975078821182251f, 0.7975078821182251f));
  const float2 g = __fmul2_rn(p, x);
  const float2 t = make_float2(tanh_approx(g.x), tanh_approx(g.y));
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(h, t, h);
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float x2 = fminf(x * x, 36.0f);
  const float p = fmaf(fmaf(x2, -0.0003515167918521911f, 0.03700564429163933f), x2, 0.7975078821182251f);
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(p * x), h);
}
#ifdef UG_GELU_ERF
__device__ __forceinline__ float2 gelu2(float2 x) { return gelu_erf2(x); }
__device__ __forceinline__ float gelu1(float x) { return gelu_erf(x); }
#else
__device__ __forceinline__ float2 gelu2(float2 x) { return gelu_tanh2(x); }
__device__ __forceinline__ float gelu1(float x) { return gelu_tanh(x); }
#endif

__device__ __forceinline__ float2 unpack16x2(uint32_t u, int fmt) {
  return fmt ? Elem<__nv_bfloat16>::unpack2(u) : Elem<__half>::unpack2(u);
}
__device__ __forceinline__ uint32_t pack16x2(float a, float b, int fmt) {
  return fmt ? Elem<__nv_bfloat16>::pack2(a, b) : Elem<__half>::pack2(a, b);
}
__device__ __forceinline__ float load16(const void* base, long long idx, int fmt) {
  return fmt ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx])
             : __half2float(reinterpret_cast<const __half*>(base)[idx]);
}
__device__ __forceinline__ void store16(void* base, long long idx, float v, int fmt) {
  if (fmt) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
}

// (sum, sum of squares) of finished output values, for the LayerNorm the NEXT GEMM folds in (TapGemmArgs::stat_out)
template <int NC>
__device__ __forceinline__ void stat_accum(const float (&f)[NC], int ncols_valid, float& s1, float& s2) {
  if (ncols_valid >= NC) {
    float2 p1 = make_float2(f[0], f[1]);
    float2 p2 = __fmul2_rn(p1, p1);
#pragma unroll
    for (int j = 2; j < NC; j += 2) {
      const float2 v = make_float2(f[j], f[j + 1]);
      p1 = __fadd2_rn(p1, v);
      p2 = __ffma2_rn(v, v, p2);
    }
    s1 += p1.x + p1.y;
    s2 += p2.x + p2.y;
  } else {
#pragma unroll
    for (int j = 0; j < NC; ++j)
      if (j < ncols_valid) { s1 += f[j]; s2 = fmaf(f[j], f[j], s2); }
  }
}

// Finishes NC (16 or 32) consecutive output columns of one row and stores them.
//   f[]    accumulator values (already scaled / gated)
//   col0   first column in OUTPUT column space; ncols_valid = how many of the NC exist
//   LNF    the launch folds a LayerNorm in (TapGemmArgs::ln_stat): ln = (rstd, -mean * rstd) of the row,
//          sb points at the PAIRED per-tile vector {colsum[c], colsum[c+1], bias[c], bias[c+1]} per column pair
//          (2 floats per column) and the columns become rstd * f + (-mean * rstd) * colsum + bias
template <int NC, bool AUX = true, bool LNF = false>
__device__ __forceinline__ void finish_cols(float (&f)[NC], const TapGemmArgs& a, long long pix, long long fb_off,
                                            int col0, const float* sb, int ncols_valid, bool row_ok,
                                            float2 ln = make_float2(1.f, 0.f)) {
  if (!row_ok || ncols_valid <= 0) return;
  const int fmt = a.fmt;
  const bool full = ncols_valid >= NC;
  const bool vec_ok = full && ((col0 & 7) == 0);
  if constexpr (LNF) {     // folded LayerNorm: two packed FMAs per column pair (columns past n_total are staged as zeros)
    const float2 r2 = make_float2(ln.x, ln.x), u2 = make_float2(ln.y, ln.y);
#pragma unroll
    for (int j = 0; j < NC / 2; ++j) {
      const float4 q = reinterpret_cast<const float4*>(sb)[j];
      const float2 o = __ffma2_rn(make_float2(f[2 * j], f[2 * j + 1]), r2,
                                  __ffma2_rn(u2, make_float2(q.x, q.y), make_float2(q.z, q.w)));
      f[2 * j] = o.x;
      f[2 * j + 1] = o.y;
    }
  } else if (sb != nullptr) {     // per-tile bias (+ frame bias when uniform over the tile), staged in smem
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) {
      const float4 b = reinterpret_cast<const float4*>(sb)[j];
      f[4 * j + 0] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
    }
  }
  if (a.fbias != nullptr && !a.fbias_uniform) {
    const float* fb = a.fbias + fb_off + col0;        // fb_off = (pixel / fbias_div) * fbias_ld, once per tile
    if (full && ((col0 & 3) == 0) && ((a.fbias_ld & 3) == 0)) {
#pragma unroll
      for (int j = 0; j < NC / 4; ++j) {
        const float4 b = ld_act(reinterpret_cast<const float4*>(fb) + j);
        f[4 * j + 0] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (full || j < ncols_valid) f[j] += ld_act(fb + j);
    }
  }
  if (a.act == 1) {
#pragma unroll
    for (int j = 0; j < NC; j += 2) {
      const float2 o = gelu2(make_float2(f[j], f[j + 1]));
      f[j] = o.x;
      f[j + 1] = o.y;
    }
  }
  if (!AUX) return;      // residual / blend handled by the caller (prefetched)
  if (a.res != nullptr) {
    const long long off = pix * a.ldr + col0;
    if (vec_ok && ((a.ldr & 7) == 0)) {
      const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.res) + off);
#pragma unroll
      for (int j = 0; j < NC / 8; ++j) {
        uint4 u = ld_act(p + j);
        float2 t0 = unpack16x2(u.x, fmt), t1 = unpack16x2(u.y, fmt), t2 = unpack16x2(u.z, fmt),
               t3 = unpack16x2(u.w, fmt);
        f[8 * j + 0] += t0.x; f[8 * j + 1] += t0.y; f[8 * j + 2] += t1.x; f[8 * j + 3] += t1.y;
        f[8 * j + 4] += t2.x; f[8 * j + 5] += t2.y; f[8 * j + 6] += t3.x; f[8 * j + 7] += t3.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (full || j < ncols_valid) f[j] += load16(a.res, off + j, fmt);
    }
  }
  if (a.blend != nullptr) {
    const long long off = pix * a.ldb + col0;
    const float al = a.alpha, be = 1.0f - a.alpha;
    if (vec_ok && ((a.ldb & 7) == 0)) {
      const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.blend) + off);
#pragma unroll
      for (int j = 0; j < NC / 8; ++j) {
        uint4 u = ld_act(p + j);
        float2 t0 = unpack16x2(u.x, fmt), t1 = unpack16x2(u.y, fmt), t2 = unpack16x2(u.z, fmt),
               t3 = unpack16x2(u.w, fmt);
        f[8 * j + 0] = al * t0.x + be * f[8 * j + 0]; f[8 * j + 1] = al * t0.y + be * f[8 * j + 1];
        f[8 * j + 2] = al * t1.x + be * f[8 * j + 2]; f[8 * j + 3] = al * t1.y + be * f[8 * j + 3];
        f[8 * j + 4] = al * t2.x + be * f[8 * j + 4]; f[8 * j + 5] = al * t2.y + be * f[8 * j + 5];
        f[8 * j + 6] = al * t3.x + be * f[8 * j + 6]; f[8 * j + 7] = al * t3.y + be * f[8 * j + 7];
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (full || j < ncols_valid) f[j] = al * load16(a.blend, off + j, fmt) + be * f[j];
    }
  }
}

template <int NC>
__device__ __forceinline__ void store_cols(const float (&f)[NC], const TapGemmArgs& a, long long pix,
                                           long long zoff, int col0, int ncols_valid, bool row_ok) {
  if (!row_ok || ncols_valid <= 0) return;
  const int fmt = a.fmt;
  const bool full = ncols_valid >= NC;
  const bool vec_ok = full && ((col0 & 7) == 0);
  const long long ooff = zoff + pix * a.ldc + col0;
  if (a.out_fp32) {
    float* o = reinterpret_cast<float*>(a.out) + ooff;
    if (full && ((a.ldc & 3) == 0) && ((col0 & 3) == 0)) {
#pragma unroll
      for (int j = 0; j < NC / 4; ++j)
        reinterpret_cast<float4*>(o)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (full || j < ncols_valid) o[j] = f[j];
    }
  } else {
    if (vec_ok && ((a.ldc & 7) == 0)) {
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(a.out) + ooff);
#pragma unroll
      for (int j = 0; j < NC / 8; ++j) {
        uint4 u;
        u.x = pack16x2(f[8 * j + 0], f[8 * j + 1], fmt);
        u.y = pack16x2(f[8 * j + 2], f[8 * j + 3], fmt);
        u.z = pack16x2(f[8 * j + 4], f[8 * j + 5], fmt);
        u.w = pack16x2(f[8 * j + 6], f[8 * j + 7], fmt);
        o[j] = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (full || j < ncols_valid) store16(a.out, ooff + j, f[j], fmt);
    }
  }
}

template <int NC, bool LNF = false>
__device__ __forceinline__ void finish_and_store(float (&f)[NC], const TapGemmArgs& a, long long pix,
                                                 long long zoff, long long fb_off, int col0, const float* sb,
                                                 int ncols_valid, bool row_ok, float2 ln = make_float2(1.f, 0.f)) {
  finish_cols<NC, true, LNF>(f, a, pix, fb_off, col0, sb, ncols_valid, row_ok, ln);
  store_cols<NC>(f, a, pix, zoff, col0, ncols_valid, row_ok);
}

// 32 finished columns of row r -> 64 bytes of the 128-byte-swizzled output slab (TMA-store staging)
__device__ __forceinline__ void stage_cols32(const float (&f)[32], uint32_t slab, int r, int half, int fmt) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t c16 = (uint32_t)(half * 4 + j);
    const uint32_t addr = slab + (uint32_t)r * 128u + ((c16 ^ (uint32_t)(r & 7)) << 4);
    const uint32_t w0 = pack16x2(f[8 * j + 0], f[8 * j + 1], fmt), w1 = pack16x2(f[8 * j + 2], f[8 * j + 3], fmt);
    const uint32_t w2 = pack16x2(f[8 * j + 4], f[8 * j + 5], fmt), w3 = pack16x2(f[8 * j + 6], f[8 * j + 7], fmt);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                 : "memory");
  }
}

// 16 finished columns (quarter qt = 0..3 of the 64-column slab) of row r
__device__ __forceinline__ void stage_cols16(const float (&f)[16], uint32_t slab, int r, int qt, int fmt) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint32_t c16 = (uint32_t)(qt * 2 + j);
    const uint32_t addr = slab + (uint32_t)r * 128u + ((c16 ^ (uint32_t)(r & 7)) << 4);
    const uint32_t w0 = pack16x2(f[8 * j + 0], f[8 * j + 1], fmt), w1 = pack16x2(f[8 * j + 2], f[8 * j + 3], fmt);
    const uint32_t w2 = pack16x2(f[8 * j + 4], f[8 * j + 5], fmt), w3 = pack16x2(f[8 * j + 6], f[8 * j + 7], fmt);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                 : "memory");
  }
}

// Persistent kernel: one CTA per SM walks output tiles t, t + stride, ...  The three roles run ahead of
// each other across tiles: the TMA ring never drains between tiles, and the accumulator is
// double-buffered in TMEM (2 x 256 columns) so the MMAs of tile i+1 overlap the epilogue of tile i.
// BN (the UMMA N extent) is a RUNTIME multiple of 16 <= 256 chosen per layer.
//
// CTAS == 2: the two CTAs of a cluster (one TPC) compute a 256 x BN tile with cta_group::2 UMMAs.  Each
// CTA stages its own 128 rows of A and HALF of the B tile, so shared-memory fill + operand-read
// traffic per FLOP drops by 1/3 against the single-CTA 128 x 256 tile (the binding resource here).
// The leader (rank 0) issues every MMA; commits are multicast to both CTAs' barriers; each CTA's TMA
// counts its bytes on the leader's full barrier; both epilogues release the accumulator by arriving
// on the leader's tmem_empty barrier.
// LN: 0 plain; 1 the epilogue applies a folded LayerNorm (TapGemmArgs::ln_stat); 2 the epilogue leaves row statistics
// for the next GEMM's folded LayerNorm (TapGemmArgs::stat_out).  Compile-time, so that the unrolled epilogue bodies
// stay branch-free (a runtime switch inside them cost the GEGLU epilogue 44 branches and half its ILP).
template <int CTAS, int LN>
__global__ void __launch_bounds__(kThreads, 1)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmBl,
               const __grid_constant__ TapGemmArgs a) {
  constexpr int kSt = CTAS == 2 ? kStages2 : kStages;
  constexpr int kStB = CTAS == 2 ? kStageBytes2 : kStageBytes;
  constexpr int G16 = 16 * CTAS;                     // granularity of the N extent of one UMMA
  extern __shared__ __align__(1024) uint8_t smem[];  // no static smem in this kernel: the dynamic window starts
  if ((smem_u32(smem) & 1023u) != 0u) __trap();      // 1024-aligned (SWIZZLE_128B tiles need it); checked, not assumed
  uint8_t* slabs = smem + kSt * kStB;                // 4 x 16 KB output staging (2 column groups x 2 buffers)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slabs + 4 * kSlabBytes);
  uint64_t* empty_bar = full_bar + kSt;
  uint64_t* tmem_full_bar = empty_bar + kSt;         // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint64_t* res_bar = tmem_empty_bar + 2;            // [2 column groups][4 chunk buffers] residual chunk landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + 8);
  float* bias_s = reinterpret_cast<float*>(slabs + 4 * kSlabBytes + 256);   // [2][256] per-tile bias vectors

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CTAS == 2 ? (int)cluster_ctarank() : 0;
  const int BN = a.bn_tile;                          // tile width (the last N tile may be narrower)
  const int m_tiles = a.tiles_x * a.tiles_y * a.tiles_n;
  const int pm_tiles = (m_tiles + CTAS - 1) / CTAS;  // tiles along M per scheduling unit
  const int n_tiles = a.n_tiles;
  const int total_tiles = pm_tiles * n_tiles * a.batch;
  // unit index -> (M tile, N tile, batch).  Default order: M fastest (one weight tile stays hot while the
  // activations stream).  n_fastest: all N tiles of an M tile run at the same time on neighbouring CTAs, so an
  // activation matrix larger than L2 is read from HBM once instead of once per N tile; the N tile is rotated by
  // the M index so that ragged (e.g. 192 + 128) tiles alternate on every CTA.
  auto decode = [&](int t, int& m_lin, int& n_tile, int& z) {
    // (launch-constant divisors through FastDiv: see tapgemm.cuh)
    if (a.n_fastest) {
      int u;
      a.fd_perz.divmod(t, z, u);
      int nrem;
      a.fd_nt.divmod(u, m_lin, nrem);
      int nq;
      a.fd_nt.divmod(nrem + m_lin, nq, n_tile);
    } else {
      int rest;
      a.fd_pm.divmod(t, rest, m_lin);
      a.fd_nt.divmod(rest, z, n_tile);
    }
  };
  const int num_iters = a.num_taps * a.kchunks;
  // split-K (a.ksplit > 1; batch == ksplit): unit z accumulates the (tap, chunk) iterations [it0, it1) of its tile into
  // its own fp32 partial (out + z * out_z1stride); a reduce kernel sums the partials in index order afterwards
  auto k_range = [&](int z, int& it0, int& it1) {
    if (a.ksplit <= 1) { it0 = 0; it1 = num_iters; return; }
    it0 = (int)(((long long)num_iters * z) / a.ksplit);
    it1 = (int)(((long long)num_iters * (z + 1)) / a.ksplit);
  };
  const int unit0 = blockIdx.x / CTAS, unit_stride = gridDim.x / CTAS;
  // i-th unit of this CTA (pair): round-robin, or the host's balanced list (ragged N tiles cost less than full ones,
  // and a plain round-robin leaves the CTAs that drew one unit more holding the whole tail).  All three roles walk
  // the same sequence.
  auto unit_at = [&](int i) -> int {
    if (a.sched != nullptr) return i < a.sched_len ? __ldg(a.sched + (size_t)unit0 * a.sched_len + i) : -1;
    const long long t = (long long)unit0 + (long long)i * unit_stride;
    return t < total_tiles ? (int)t : -1;
  };
  // N extent of the tile starting at column n0: ragged last tile, rounded up to the UMMA granularity
  auto n_cur = [&](int n0) {
    const int rem = (a.n_total - n0 + G16 - 1) / G16 * G16;
    return rem < BN ? rem : BN;
  };

  pdl_launch_dependents();   // the next kernel may stage its prologue while this one drains
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (a.tma_store) tma_prefetch_desc(&tmC);
    if (a.res_tma) tma_prefetch_desc(&tmR);
    if (a.blend_tma) tma_prefetch_desc(&tmBl);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kSt; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tmem_full_bar[b], 1);
        mbar_init(&tmem_empty_bar[b], kEpiWarps * CTAS);
      }
      for (int b = 0; b < 8; ++b) mbar_init(&res_bar[b], 1);
      fence_mbar_init();
    }
    __syncwarp();
    if constexpr (CTAS == 2) {
      tmem_alloc_pair(tmem_ptr_smem, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr_smem, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();                // everything above (barriers, TMEM, descriptors) overlapped the previous kernel

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const uint32_t tx_bytes = ((uint32_t)(a.bw * a.bh * a.bn) * (BK * 2) +
                                 (uint32_t)(a.b_mn_major ? BK : BN / CTAS) * (BK * 2)) * CTAS;
      int it_g = 0;   // ring position, continuous across tiles
      for (int ui = 0, t; (t = unit_at(ui)) >= 0; ++ui) {
        UG_TRACE(0, ui, 0);
        int m_lin, n_tile_p, z;
        decode(t, m_lin, n_tile_p, z);
        int m_tile = m_lin * CTAS + rank;
        const int n0 = n_tile_p * BN;
        int z1, z0, tx, ty, tn, mrest;
        a.fd_zdiv.divmod(z, z1, z0);
        a.fd_tx.divmod(m_tile, mrest, tx);
        a.fd_ty.divmod(mrest, tn, ty);                 // tn >= tiles_n for the phantom half of an odd pair: OOB -> zeros
        int base[6] = {0, 0, 0, 0, 0, 0};   // slot 5 swallows unused roles
        base[a.dim_x] += tx * a.bw;
        base[a.dim_y] += ty * a.bh;
        base[a.dim_n] += tn * a.bn;
        base[a.dim_z1] += z1 * a.a_z1step;
        base[a.dim_z0] += z0 * a.a_z0step;
        const int bcol0 = a.b_c0 + z0 * a.b_z0_cstep;
        // pair mode: this CTA stages the rank-th half of the tile's (possibly ragged) B rows
        const int brow0 = z1 * a.b_z1_rowstep + (a.b_mn_major ? 0 : n0 + rank * (n_cur(n0) / CTAS));
        int it0, it1;
        k_range(z, it0, it1);
        for (int it = it0; it < it1; ++it, ++it_g) {
          const int s = it_g % kSt;
          const uint32_t ph = (uint32_t)(it_g / kSt) & 1u;
          int tap, kc;
          a.fd_kc.divmod(it, tap, kc);
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (it == it0) UG_TRACE(0, ui, 1);
          if (it == it1 - 1) UG_TRACE(0, ui, 2);
          uint8_t* sa = smem + s * kStB;
          uint8_t* sb = sa + kATileBytes;
          if constexpr (CTAS == 2) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
            tma_load_5d_pair(sa, &tmA, &full_bar[s], base[0] + a.tap_off[tap][0] + kc * BK,
                             base[1] + a.tap_off[tap][1], base[2] + a.tap_off[tap][2], base[3] + a.tap_off[tap][3],
                             base[4] + a.tap_off[tap][4]);
            tma_load_2d_pair(sb, &tmB, &full_bar[s], bcol0 + kc * BK, brow0 + tap * a.b_tap_rows);
          } else {
            mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
            tma_load_5d(sa, &tmA, &full_bar[s], base[0] + a.tap_off[tap][0] + kc * BK, base[1] + a.tap_off[tap][1],
                        base[2] + a.tap_off[tap][2], base[3] + a.tap_off[tap][3], base[4] + a.tap_off[tap][4]);
            if (a.b_mn_major)  // [64 K rows][64 N elements] box of a row-major [K, N] matrix
              tma_load_2d(sb, &tmB, &full_bar[s], bcol0 + n0, brow0 + kc * BK);
            else
              tma_load_2d(sb, &tmB, &full_bar[s], bcol0 + kc * BK, brow0 + tap * a.b_tap_rows);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer (leader CTA only in pair mode) =====================
    if (rank == 0) {
      const uint32_t bstep = a.b_mn_major ? 128u : 2u;
      int it_g = 0, tl = 0;
      for (int ui = 0, t; (t = unit_at(ui)) >= 0; ++ui, ++tl) {
        int m_lin_u, n_tile_u, z_u;
        decode(t, m_lin_u, n_tile_u, z_u);
        const int n0 = n_tile_u * BN;
        const uint32_t idesc = make_idesc_f16(BM * CTAS, n_cur(n0), a.fmt, a.b_mn_major);
        const int buf = tl & 1;
        const uint32_t use = (uint32_t)(tl >> 1);
        if (lane == 0) UG_TRACE(1, tl, 0);
        mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u);     // epilogue(s) drained this accumulator
        tc_fence_after();
        if (lane == 0) UG_TRACE(1, tl, 1);
        const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256u;
        int it0, it1;
        k_range(z_u, it0, it1);
        for (int it = it0; it < it1; ++it, ++it_g) {
          const int s = it_g % kSt;
          const uint32_t ph = (uint32_t)(it_g / kSt) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (lane == 0 && it == it0) UG_TRACE(1, tl, 2);
          if (lane == 0 && it == it1 - 1) UG_TRACE(1, tl, 3);
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + s * kStB);
            const uint64_t da = make_desc_kmajor_sw128(sa);
            // K-major: +32 B per K=16 slice inside the 128-byte swizzle row (start field += 2).
            // MN-major: a K=16 slice is 16 rows of 128 B (start field += 128).
            const uint64_t db = a.b_mn_major ? make_desc_mnmajor_sw128(sa + kATileBytes, 8192)
                                             : make_desc_kmajor_sw128(sa + kATileBytes);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if constexpr (CTAS == 2)
                umma_f16_pair(tmem_d, da + 2 * k, db + bstep * k, idesc, (it != it0 || k != 0) ? 1u : 0u);
              else
                umma_f16(tmem_d, da + 2 * k, db + bstep * k, idesc, (it != it0 || k != 0) ? 1u : 0u);
            }
            if constexpr (CTAS == 2) {
              umma_commit_pair(&empty_bar[s]);
              if (it == it1 - 1) umma_commit_pair(&tmem_full_bar[buf]);
            } else {
              umma_commit(&empty_bar[s]);
              if (it == it1 - 1) umma_commit(&tmem_full_bar[buf]);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue (8 warps: 4 lane quarters x 2 column groups) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int hsel = (warp - 2) >> 2;       // column group
    const int r = q * 32 + lane;            // output row inside this CTA's 128-row tile
    const int xi = r % a.bw;
    const int yi = (r / a.bw) % a.bh;
    const int ni = r / (a.bw * a.bh);
    const bool tracer = (q == 0) && lane == 0;            // (debug builds) one stamping thread per column group
    int trace_tl = 0;
    (void)tracer; (void)trace_tl;
    auto release = [&](int buf) {           // this warp no longer needs accumulator `buf`
      tc_fence_before();
      __syncwarp();
      if (tracer) UG_TRACE(2 + hsel, trace_tl, 2);
      if (lane == 0) {
        if constexpr (CTAS == 2) mbar_arrive_leader(&tmem_empty_bar[buf]);
        else mbar_arrive(&tmem_empty_bar[buf]);
      }
    };
    const bool issuer = (warp == 2 + 4 * hsel) && lane == 0;   // issues this group's TMA stores
    const uint32_t slab_base = smem_u32(slabs + hsel * 2 * kSlabBytes);
    int slab_it = 0;
    uint32_t ring_it = 0;                   // chunk ring position of this column group (res_tma path)
    int tl = 0;
    constexpr bool lnf = LN == 1;
    constexpr bool stats = LN == 2;
    // Per-tile vectors ONE TILE AHEAD.  Every tile needs, per thread, its column's bias (+ frame bias) and -- with a folded
    // LayerNorm -- the column sum and this thread's row statistics.  Requested at the start of the tile that uses them,
    // each is an exposed L2 round trip (~1000 clk under load) per tile in a loop whose period IS the epilogue's time; here
    // the loads of tile i + 1 are issued while tile i is processed and their values wait in registers.  (Needs the cheap
    // index decoding: with ten integer divisions per decode the second decode per tile cost more than the loads.)
    const int et = threadIdx.x - 64;
    const bool have_sb = a.bias != nullptr || a.fbias_uniform;
    float pf_b = 0.f, pf_fb = 0.f, pf_c = 0.f;
    float4 pf_s0 = make_float4(0.f, 0.f, 0.f, 0.f), pf_s1 = make_float4(0.f, 0.f, 0.f, 0.f);
    // (predicated loads whose destination IS the carried variable: nothing here may read a loaded value, or the in-order
    // warp waits for the L2 round trip at the point of issue)
    int pm = 0, pn = 0, pz = 0;                  // the prefetched tile's decoded index: decoded once, used by both
    auto prefetch_tile = [&](int tt) {
      decode(tt, pm, pn, pz);
      const int mt = pm * CTAS + rank;
      const int col = pn * BN + et;
      const bool col_ok = have_sb && col < a.n_total;
      pf_b = ld_act_pred(a.bias + col, col_ok && a.bias != nullptr);        // the time-embedding bias is rewritten every step
      pf_fb = ld_act_pred(a.fbias + (long long)a.fd_fb.div(mt * BM) * a.fbias_ld + col, col_ok && a.fbias_uniform);
      if constexpr (lnf) {
        pf_c = ldg_pred(a.ln_colsum + col, col < a.n_total);
        const long long prow = (long long)mt * BM + r;          // linear ops only: row = pixel
        ld_stats_pred(a.ln_stat + prow * (a.ln_parts > 0 ? a.ln_parts : 1), a.ln_parts, prow < (long long)a.W, pf_s0, pf_s1);
      }
    };
    int t = unit_at(0), t_next = -1;
    if (t >= 0) prefetch_tile(t);
    for (int ui = 0; t >= 0; ++ui, ++tl, t = t_next) {
      trace_tl = tl;
      if (tracer) UG_TRACE(2 + hsel, tl, 0);
      t_next = unit_at(ui + 1);
      const float cur_b = pf_b + pf_fb, cur_c = pf_c;
      const float4 cur_s0 = pf_s0, cur_s1 = pf_s1;
      const int m_lin = pm, n_tile = pn, z = pz;               // decoded when this tile's vectors were requested
      int m_tile = m_lin * CTAS + rank;
      const int m_tile_lin = m_tile;
      const int n0 = n_tile * BN;
      const int Ncur = n_cur(n0);
      int z1 = 0, z0 = 0, tx = m_tile, ty = 0, tn = 0, mrest;
      if (!a.flat) {                        // (plain GEMM rows skip this chain: it sits in front of every tile's epilogue)
        a.fd_zdiv.divmod(z, z1, z0);
        a.fd_tx.divmod(m_tile, mrest, tx);
        a.fd_ty.divmod(mrest, tn, ty);
      }
      const int x0 = tx * a.bw, y0 = ty * a.bh, nn0 = tn * a.bn;
      const bool row_ok = a.flat ? (x0 + r < a.W)
                                 : (ni < a.bn) && (x0 + xi < a.W) && (y0 + yi < a.H) && (nn0 + ni < a.N);
      const long long pix = a.flat ? (long long)(x0 + r)
                                   : ((long long)(nn0 + ni) * a.H + (y0 + yi)) * a.W + (x0 + xi);
      const long long zoff = a.flat ? 0ll : (long long)z1 * a.out_z1stride + (long long)z0 * a.out_z0stride;
      const long long fb_off = a.fbias ? (long long)a.fd_fb.div((int)pix) * a.fbias_ld : 0;
      const int buf = tl & 1;
      // LayerNorm folded into this GEMM: (rstd, -mean * rstd) of this thread's row, requested ahead of the accumulator
      // wait.  Partial form: the (sum, sum of squares) pieces the producing GEMM's epilogue left, folded in index order.
      float2 lnv = make_float2(0.f, 0.f);
      if (lnf && row_ok) {
        if (a.ln_parts == 0) {
          lnv = make_float2(cur_s0.y, -cur_s0.x * cur_s0.y);
        } else {                                                 // 2 or 4 (sum, sum of squares) partials, in index order
          const float s1 = (cur_s0.x + cur_s0.z) + (cur_s1.x + cur_s1.z);
          const float s2 = (cur_s0.y + cur_s0.w) + (cur_s1.y + cur_s1.w);
          const float mean = s1 * a.ln_inv_c;
          const float rstd = rsqrtf(fmaxf(fmaf(s2, a.ln_inv_c, -mean * mean), 0.f) + a.ln_eps);
          lnv = make_float2(rstd, -mean * rstd);
        }
      }
      const float2 ln = lnv;
      float st1 = 0.f, st2 = 0.f;                                        // producer side (stat_out)
      // bias (+ frame bias when every row of the tile shares the frame) of this tile's columns -> smem
      // folded LayerNorm: ONE buffer of paired {colsum, colsum, bias, bias} per column pair (the double buffer has no
      // room for two vectors), fenced by a second barrier so that no warp still reads the previous tile's vector
      float* sbt = lnf ? bias_s : bias_s + buf * 256;
#if defined(UG_TRACE_SPLIT) && UG_TRACE_SPLIT == 1
      if (tracer) UG_TRACE(2 + hsel, tl, 1);      // (debug variants) where inside the tile prologue the time goes
#endif
      if (have_sb) {
        if (lnf) {
          named_bar_sync(3, kEpiWarps * 32);
          sbt[(et >> 1) * 4 + (et & 1)] = cur_c;
          sbt[(et >> 1) * 4 + 2 + (et & 1)] = cur_b;
        } else {
          sbt[et] = cur_b;
        }
        named_bar_sync(3, kEpiWarps * 32);
      }
#if defined(UG_TRACE_SPLIT) && UG_TRACE_SPLIT == 2
      if (tracer) UG_TRACE(2 + hsel, tl, 1);
#endif
      if (t_next >= 0) prefetch_tile(t_next);
      (void)m_tile_lin;
      auto sb_at = [&](int c) -> const float* { return lnf ? sbt + 2 * c : (have_sb ? sbt + c : nullptr); };
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 256u;
      const int BNh = BN >> 1;              // GEGLU: [BNh value | BNh gate] accumulator columns
      const bool prefetching = a.res_tma || (a.tma_store && !a.geglu && (a.res != nullptr || a.blend != nullptr));
      if (!prefetching) {                   // (the prefetching path waits after issuing its first loads)
#if defined(UG_TRACE_SPLIT) && UG_TRACE_SPLIT == 3
        if (tracer) UG_TRACE(2 + hsel, tl, 1);    // (debug variant) stamp BEFORE the accumulator wait: isolates the tile prologue
#endif
        mbar_wait(&tmem_full_bar[buf], (uint32_t)(tl >> 1) & 1u);
#ifndef UG_TRACE_SPLIT
        if (tracer) UG_TRACE(2 + hsel, tl, 1);
#endif
        tc_fence_after();
      }

      if (a.res_tma) {
        // ---- residual by TMA, in place: the store staging memory of this column group is a ring of four
        // [128 rows x 32 columns] chunk buffers (SWIZZLE_64B).  A residual chunk lands in its buffer two chunks
        // ahead of use (the first two of a tile while the accumulator is still being produced), each thread adds
        // its accumulator row into its own 64 bytes, and the same buffer leaves through a TMA tile store.
        // Replaces 4 x LDG.128 per row and chunk, whose 32 distinct lines per request made the L1 tag stage the
        // bottleneck of every residual-carrying GEMM (ncu: 30 sectors per request, +20 us per launch at L0).
        const int out_w = Ncur, out_c0 = n0;
        const float al = a.alpha, be = 1.0f - a.alpha;
        const bool same_aux = a.blend != nullptr && a.blend == a.res && a.ldr == a.ldb;
        // A blend tensor that is not the residual: by TMA too where the launcher could build its map (pairm) -- the ring
        // then works as two (residual, blend) buffer pairs with one chunk of look-ahead -- else by per-thread loads
        // (4 x LDG.128 per row and chunk touch 32 lines per request: +37 us on the L0 ff.net.2 of the temporal block).
        const bool pairm = a.blend_tma != 0;
        const bool blend_regs = a.blend != nullptr && !same_aux && !pairm;
        // 32-column chunks alternate between the two column groups: a 192-wide tile splits 3 : 3 (64-column slabs
        // alternating gave 4 : 2 and the heavier group set the tile's time)
        auto chunk_col = [&](int qq) { return (2 * qq + hsel) * 32; };
        int nch = 0;
        while (nch < 8 && chunk_col(nch) < out_w) ++nch;
        uint8_t* ring = slabs + hsel * 2 * kSlabBytes;
        uint64_t* rbar = res_bar + hsel * 4;
        const uint32_t chunk_bytes = (uint32_t)(a.bw * a.bh * a.bn) * 64u;
        auto issue_load = [&](int qq) {
          if (pairm) {                                         // pair (it & 1): residual -> buffer 2p, blend -> 2p + 1
            const int b = (int)((ring_it + (uint32_t)qq) & 1u) * 2;
            mbar_arrive_expect_tx(&rbar[b], 2 * chunk_bytes);
            tma_load_5d(ring + b * 8192, &tmR, &rbar[b], out_c0 + chunk_col(qq), x0, y0, nn0, 0);
            tma_load_5d(ring + (b + 1) * 8192, &tmBl, &rbar[b], out_c0 + chunk_col(qq), x0, y0, nn0, 0);
            return;
          }
          const int b = (int)((ring_it + (uint32_t)qq) & 3u);
          mbar_arrive_expect_tx(&rbar[b], chunk_bytes);
          tma_load_5d(ring + b * 8192, &tmR, &rbar[b], out_c0 + chunk_col(qq), x0, y0, nn0, 0);
        };
        if (issuer && nch > 0) {
          if (pairm) {
            bulk_wait_group_read<0>();                         // the pair's residual buffer was the last one stored from
            issue_load(0);
          } else {
            bulk_wait_group_read<1>();                         // the stores that last read these buffers are done
            issue_load(0);
            if (nch > 1) issue_load(1);
          }
        }
        mbar_wait(&tmem_full_bar[buf], (uint32_t)(tl >> 1) & 1u);
        if (tracer) UG_TRACE(2 + hsel, tl, 1);
        tc_fence_after();
        bool released = false;
#pragma unroll 1
        for (int qq = 0; qq < nch; ++qq) {
          const uint32_t it = ring_it + (uint32_t)qq;
          const int b = pairm ? (int)(it & 1u) * 2 : (int)(it & 3u);
          const int c = chunk_col(qq);
          if (!pairm && issuer && qq + 2 < nch) {
            bulk_wait_group_read<1>();                         // buffer (it + 2) & 3 was stored two chunks ago
            issue_load(qq + 2);
          }
          uint4 bq4[4];
          if (blend_regs && row_ok) {
            const uint4* pb = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.blend) +
                                                            pix * a.ldb + out_c0 + c);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(bq4[j].x), "=r"(bq4[j].y), "=r"(bq4[j].z), "=r"(bq4[j].w) : "l"(pb + j));
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) bq4[j] = make_uint4(0u, 0u, 0u, 0u);
          }
          uint32_t v[32];
          tmem_ld_32x32(trow + c, v);
          tmem_ld_wait();
          if (qq == nch - 1) { release(buf); released = true; }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (a.scale != 1.0f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= a.scale;
          }
          finish_cols<32, false, lnf>(f, a, pix, fb_off, out_c0 + c, sb_at(c), a.n_total - (out_c0 + c), row_ok, ln);
          if (pairm && issuer && qq + 1 < nch) {               // the other pair: its residual buffer left with chunk qq - 1
            bulk_wait_group_read<0>();
            issue_load(qq + 1);
          }
          mbar_wait(&rbar[b], pairm ? (it >> 1) & 1u : (it >> 2) & 1u);
          const uint32_t rowaddr = smem_u32(ring + b * 8192) + (uint32_t)r * 64u;
          const uint32_t sw = (uint32_t)((r >> 1) & 3);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t addr = rowaddr + ((((uint32_t)j) ^ sw) << 4);
            uint4 u;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
            const float2 t0 = unpack16x2(u.x, a.fmt), t1 = unpack16x2(u.y, a.fmt), t2 = unpack16x2(u.z, a.fmt),
                         t3 = unpack16x2(u.w, a.fmt);
            f[8 * j + 0] += t0.x; f[8 * j + 1] += t0.y; f[8 * j + 2] += t1.x; f[8 * j + 3] += t1.y;
            f[8 * j + 4] += t2.x; f[8 * j + 5] += t2.y; f[8 * j + 6] += t3.x; f[8 * j + 7] += t3.y;
            if (a.blend != nullptr) {
              uint4 bw = same_aux ? u : bq4[j];
              if (pairm)                                       // the blend chunk sits in the pair's second buffer, same layout
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(bw.x), "=r"(bw.y), "=r"(bw.z), "=r"(bw.w) : "r"(addr + 8192u));
              const float2 s0 = unpack16x2(bw.x, a.fmt), s1 = unpack16x2(bw.y, a.fmt), s2 = unpack16x2(bw.z, a.fmt),
                           s3 = unpack16x2(bw.w, a.fmt);
              f[8 * j + 0] = al * s0.x + be * f[8 * j + 0]; f[8 * j + 1] = al * s0.y + be * f[8 * j + 1];
              f[8 * j + 2] = al * s1.x + be * f[8 * j + 2]; f[8 * j + 3] = al * s1.y + be * f[8 * j + 3];
              f[8 * j + 4] = al * s2.x + be * f[8 * j + 4]; f[8 * j + 5] = al * s2.y + be * f[8 * j + 5];
              f[8 * j + 6] = al * s3.x + be * f[8 * j + 6]; f[8 * j + 7] = al * s3.y + be * f[8 * j + 7];
            }
            const uint32_t w0 = pack16x2(f[8 * j + 0], f[8 * j + 1], a.fmt), w1 = pack16x2(f[8 * j + 2], f[8 * j + 3], a.fmt);
            const uint32_t w2 = pack16x2(f[8 * j + 4], f[8 * j + 5], a.fmt), w3 = pack16x2(f[8 * j + 6], f[8 * j + 7], a.fmt);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
          }
          if constexpr (stats) stat_accum<32>(f, 32, st1, st2);
          fence_proxy_async_smem();
          named_bar_sync(1 + hsel, 128);
          if (issuer) {
            tma_store_5d(&tmC, reinterpret_cast<const void*>(ring + b * 8192), out_c0 + c, x0, y0, nn0, 0);
            bulk_commit_group();
          }
        }
        ring_it += (uint32_t)nch;
        if (!released) release(buf);
      } else if (a.tma_store && !a.geglu && (a.res != nullptr || a.blend != nullptr)) {
        // ---- coalesced stores + software-pipelined residual / blend loads: the 64 bytes a row needs for
        // chunk k+1 are requested before chunk k is computed (and, for the first chunk, before the
        // accumulator is even ready), so their L2/HBM latency hides under the MMA wait and the math.
        const int out_w = Ncur, out_c0 = n0, n_out = a.n_total;
        const float al = a.alpha, be = 1.0f - a.alpha;
        // SpatioTemporalResBlock passes x_spatial as BOTH the temporal resnet's residual and the AlphaBlender's
        // spatial input: one load serves both (out = (al + be) * x + be * (acc + bias))
        const bool same_aux = a.res != nullptr && a.res == a.blend && a.ldr == a.ldb;
        uint4 rq[2][4], bq[2][4];
        auto prefetch = [&](int which, int c) {
          const bool ok = row_ok && (c < out_w) && (out_c0 + c + 32 <= n_out);
#pragma unroll
          for (int j = 0; j < 4; ++j) { rq[which][j] = make_uint4(0u, 0u, 0u, 0u); bq[which][j] = make_uint4(0u, 0u, 0u, 0u); }
          if (ok) {
            if (a.res != nullptr) {
              const uint4* pr = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.res) +
                                                              pix * a.ldr + out_c0 + c);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(rq[which][j].x), "=r"(rq[which][j].y), "=r"(rq[which][j].z), "=r"(rq[which][j].w)
                             : "l"(pr + j));
            }
            if (a.blend != nullptr && !same_aux) {
              const uint4* pb = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.blend) +
                                                              pix * a.ldb + out_c0 + c);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(bq[which][j].x), "=r"(bq[which][j].y), "=r"(bq[which][j].z), "=r"(bq[which][j].w)
                             : "l"(pb + j));
            }
          }
        };
        auto apply_aux = [&](int which, float (&f)[32], int c) {
          if (!(row_ok && out_c0 + c + 32 <= n_out)) return;     // ragged tail columns never reach memory
          if (a.res != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 t0 = unpack16x2(rq[which][j].x, a.fmt), t1 = unpack16x2(rq[which][j].y, a.fmt),
                           t2 = unpack16x2(rq[which][j].z, a.fmt), t3 = unpack16x2(rq[which][j].w, a.fmt);
              f[8 * j + 0] += t0.x; f[8 * j + 1] += t0.y; f[8 * j + 2] += t1.x; f[8 * j + 3] += t1.y;
              f[8 * j + 4] += t2.x; f[8 * j + 5] += t2.y; f[8 * j + 6] += t3.x; f[8 * j + 7] += t3.y;
            }
          }
          if (a.blend != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 bw = same_aux ? rq[which][j] : bq[which][j];
              const float2 t0 = unpack16x2(bw.x, a.fmt), t1 = unpack16x2(bw.y, a.fmt),
                           t2 = unpack16x2(bw.z, a.fmt), t3 = unpack16x2(bw.w, a.fmt);
              f[8 * j + 0] = al * t0.x + be * f[8 * j + 0]; f[8 * j + 1] = al * t0.y + be * f[8 * j + 1];
              f[8 * j + 2] = al * t1.x + be * f[8 * j + 2]; f[8 * j + 3] = al * t1.y + be * f[8 * j + 3];
              f[8 * j + 4] = al * t2.x + be * f[8 * j + 4]; f[8 * j + 5] = al * t2.y + be * f[8 * j + 5];
              f[8 * j + 6] = al * t3.x + be * f[8 * j + 6]; f[8 * j + 7] = al * t3.y + be * f[8 * j + 7];
            }
          }
        };
        auto chunk = [&](int which, int c, uint32_t slab, int half, bool last_ld, bool& released) {
          uint32_t v[32];
          tmem_ld_32x32(trow + c, v);
          tmem_ld_wait();
          if (last_ld) { release(buf); released = true; }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (a.scale != 1.0f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= a.scale;
          }
          if (out_c0 + c + 32 <= n_out) {
            finish_cols<32, false, lnf>(f, a, pix, fb_off, out_c0 + c, sb_at(c), n_out - (out_c0 + c), row_ok, ln);
            apply_aux(which, f, c);
          } else {                                               // ragged tail: guarded scalar loads
            finish_cols<32, true, lnf>(f, a, pix, fb_off, out_c0 + c, sb_at(c), n_out - (out_c0 + c), row_ok, ln);
          }
          if constexpr (stats) stat_accum<32>(f, n_out - (out_c0 + c), st1, st2);
          stage_cols32(f, slab, r, half, a.fmt);
        };
        prefetch(0, hsel * 64);                                  // overlaps the wait for the accumulator
        mbar_wait(&tmem_full_bar[buf], (uint32_t)(tl >> 1) & 1u);
        if (tracer) UG_TRACE(2 + hsel, tl, 1);
        tc_fence_after();
        bool released = false;
#pragma unroll 1
        for (int sl = hsel; sl * 64 < out_w; sl += 2, ++slab_it) {
          const uint32_t slab = slab_base + (uint32_t)(slab_it & 1) * kSlabBytes;
          if (issuer) bulk_wait_group_read<1>();
          named_bar_sync(1 + hsel, 128);
          const bool last_slab = (sl + 2) * 64 >= out_w;
          const int c0 = sl * 64, c1 = c0 + 32;
          prefetch(1, c1);                                       // second half of this slab
          chunk(0, c0, slab, 0, last_slab && c1 >= out_w, released);
          prefetch(0, c0 + 128);                                 // first half of this warp's next slab
          if (c1 < out_w) chunk(1, c1, slab, 1, last_slab, released);
          fence_proxy_async_smem();
          named_bar_sync(1 + hsel, 128);
          if (issuer) {
            tma_store_5d(&tmC, reinterpret_cast<const void*>(slabs + hsel * 2 * kSlabBytes + (slab_it & 1) * kSlabBytes),
                         out_c0 + sl * 64, x0, y0, nn0, 0);
            bulk_commit_group();
          }
        }
        if (!released) release(buf);
      } else if (a.tma_store) {
        // ---- coalesced path: finished 64-column slabs go through swizzled smem and a TMA tile store
        const int out_w = a.geglu ? BNh : Ncur;               // output columns of this tile
        const int out_c0 = a.geglu ? n_tile * BNh : n0;       // first output column
        const int n_out = a.geglu ? a.n_total / 2 : a.n_total;
        bool released = false;
#pragma unroll 1
        for (int sl = hsel; sl * 64 < out_w; sl += 2, ++slab_it) {
          const uint32_t slab = slab_base + (uint32_t)(slab_it & 1) * kSlabBytes;
          if (issuer) bulk_wait_group_read<1>();               // the store that last read this buffer is done
          named_bar_sync(1 + hsel, 128);
          const bool last_slab = (sl + 2) * 64 >= out_w;
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            const int c = sl * 64 + half * 32;
            if (c >= out_w) break;
            const bool last_ld = last_slab && (half == 1 || c + 32 >= out_w);
            if (a.geglu) {
              // value and gate columns 16 at a time: both accumulators in flight (tcgen05.ld x16), packed fp32 math
              uint32_t v[2][16], g[2][16];
              tmem_ld_32x16(trow + c, v[0]);
              tmem_ld_32x16(trow + BNh + c, g[0]);
              tmem_ld_32x16(trow + c + 16, v[1]);
              tmem_ld_32x16(trow + BNh + c + 16, g[1]);
              tmem_ld_wait();
              if (last_ld) { release(buf); released = true; }
#pragma unroll
              for (int hq = 0; hq < 2; ++hq) {
                const float2* bv = reinterpret_cast<const float2*>(sbt + c + 16 * hq);
                const float2* bg = reinterpret_cast<const float2*>(sbt + BNh + c + 16 * hq);
                // folded LayerNorm: paired {colsum, colsum, bias, bias} vectors of the value / gate columns
                const float4* qv4 = reinterpret_cast<const float4*>(sbt + 2 * (c + 16 * hq));
                const float4* qg4 = reinterpret_cast<const float4*>(sbt + 2 * (BNh + c + 16 * hq));
                const float2 lr2 = make_float2(lnv.x, lnv.x), lu2 = make_float2(lnv.y, lnv.y);
                float fq[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float2 val = make_float2(__uint_as_float(v[hq][2 * j]), __uint_as_float(v[hq][2 * j + 1]));
                  float2 gate = make_float2(__uint_as_float(g[hq][2 * j]), __uint_as_float(g[hq][2 * j + 1]));
                  if constexpr (lnf) {
                    const float4 qv = qv4[j], qg = qg4[j];
                    val = __ffma2_rn(val, lr2, __ffma2_rn(lu2, make_float2(qv.x, qv.y), make_float2(qv.z, qv.w)));
                    gate = __ffma2_rn(gate, lr2, __ffma2_rn(lu2, make_float2(qg.x, qg.y), make_float2(qg.z, qg.w)));
                  } else if (have_sb) {
                    val = __fadd2_rn(val, bv[j]);
                    gate = __fadd2_rn(gate, bg[j]);
                  }
                  const float2 o = __fmul2_rn(val, gelu2(gate));
                  fq[2 * j] = o.x;
                  fq[2 * j + 1] = o.y;
                }
                stage_cols16(fq, slab, r, half * 2 + hq, a.fmt);
              }
              continue;
            }
            float f[32];
            {
              uint32_t v[32];
              tmem_ld_32x32(trow + c, v);     // columns past Ncur are stale but never reach memory (map clips)
              tmem_ld_wait();
              if (last_ld) { release(buf); released = true; }
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
              if (a.scale != 1.0f) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] *= a.scale;
              }
              finish_cols<32, true, lnf>(f, a, pix, fb_off, out_c0 + c, sb_at(c), n_out - (out_c0 + c), row_ok, ln);
              if constexpr (stats) stat_accum<32>(f, n_out - (out_c0 + c), st1, st2);
            }
            stage_cols32(f, slab, r, half, a.fmt);
          }
          fence_proxy_async_smem();
          named_bar_sync(1 + hsel, 128);
          if (issuer) {
            tma_store_5d(&tmC, reinterpret_cast<const void*>(slabs + hsel * 2 * kSlabBytes + (slab_it & 1) * kSlabBytes),
                         out_c0 + sl * 64, x0, y0, nn0, 0);
            bulk_commit_group();
          }
        }
        if (!released) release(buf);
      } else if (a.geglu) {
        const int ocol_tile = n_tile * BNh;
#pragma unroll 1
        for (int c = hsel * 32; c < BNh; c += 64) {
          uint32_t v[32], g[32];
          tmem_ld_32x32(trow + c, v);
          tmem_ld_32x32(trow + BNh + c, g);
          tmem_ld_wait();
          if (c + 64 >= BNh) release(buf);   // last chunk of this warp
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int vc = n0 + c + j, gc = n0 + BNh + c + j;
            float val = __uint_as_float(v[j]);
            float gate = __uint_as_float(g[j]);
            if constexpr (lnf) {
              const int iv = ((c + j) >> 1) * 4 + ((c + j) & 1), ig = ((BNh + c + j) >> 1) * 4 + ((BNh + c + j) & 1);
              val = fmaf(val, lnv.x, fmaf(lnv.y, sbt[iv], sbt[iv + 2]));
              gate = fmaf(gate, lnv.x, fmaf(lnv.y, sbt[ig], sbt[ig + 2]));
            } else if (have_sb) {
              val += sbt[c + j];
              gate += sbt[BNh + c + j];
            }
            (void)vc; (void)gc;
            f[j] = val * gelu1(gate);
          }
          finish_and_store<32>(f, a, pix, zoff, fb_off, ocol_tile + c, nullptr, a.n_total / 2 - (ocol_tile + c), row_ok);
        }
      } else {
        bool released = false;
#pragma unroll 1
        for (int c = hsel * 32; c < Ncur; c += 64) {
          const bool last = (c + 64 >= Ncur);
          if (Ncur - c >= 32) {
            uint32_t v[32];
            tmem_ld_32x32(trow + c, v);
            tmem_ld_wait();
            if (last) { release(buf); released = true; }
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (a.scale != 1.0f) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] *= a.scale;
            }
            finish_and_store<32, lnf>(f, a, pix, zoff, fb_off, n0 + c, sb_at(c), a.n_total - (n0 + c), row_ok, ln);
            if constexpr (stats) stat_accum<32>(f, a.n_total - (n0 + c), st1, st2);
          } else {
            uint32_t v[16];
            tmem_ld_32x16(trow + c, v);
            tmem_ld_wait();
            if (last) { release(buf); released = true; }
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * a.scale;
            finish_and_store<16, lnf>(f, a, pix, zoff, fb_off, n0 + c, sb_at(c), a.n_total - (n0 + c), row_ok, ln);
            if constexpr (stats) stat_accum<16>(f, a.n_total - (n0 + c), st1, st2);
          }
        }
        if (!released) release(buf);         // this warp had no chunk in this tile (narrow tile)
      }
      // producer side of a folded LayerNorm: this row's partial over the columns this group finished in this tile
      // (zeros when the group had none), one slot per (N tile, column group), each written exactly once per launch
      if (stats && row_ok)
        a.stat_out[pix * a.stat_parts + n_tile * 2 + hsel] = make_float2(st1, st2);
      if (tracer) UG_TRACE(2 + hsel, tl, 3);
    }
    if (issuer && a.tma_store) bulk_wait_group<0>();            // all tile stores have landed
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTAS == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

}  // namespace

int encode_tmap(CUtensorMap* out, const TmapDesc& d) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) return -1;
  cuuint64_t dims[5];
  cuuint64_t strides[4];
  cuuint32_t box[5], estr[5];
  for (int i = 0; i < d.rank; ++i) {
    dims[i] = d.dims[i];
    box[i] = d.box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < d.rank; ++i) strides[i] = d.strides[i];
  CUresult r = fn(out, d.elem_fmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                  (cuuint32_t)d.rank, const_cast<void*>(d.ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, d.swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

int tapgemm_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// Tile shape: (CTAs per tile, N extent).  Cost model per k-iteration in SM cycles, from the measured
// behaviour that shared-memory bandwidth (128 B/clk: TMA fill + UMMA operand reads) binds before the
// tensor pipe does:   1 CTA: 256 + 2 BN      CTA pair: max(2 BN, 256 + BN)
// times the number of waves over the SMs (pairs: over SM pairs), plus a fixed per-tile cost.
int tapgemm_pick_tile(const TapGemmArgs& a, int batch, int* ctas_out, int* ksplit_out) {
  *ctas_out = 1;
  if (ksplit_out) *ksplit_out = 1;
  if (a.b_mn_major) return 64;           // MN-major B boxes are [64 K][64 N]
  const long long m_tiles = (long long)a.tiles_x * a.tiles_y * a.tiles_n;
  const int sms = tapgemm_num_sms();
  static const int force = [] {
    const char* e = getenv("UG_TAPGEMM_CTAS");
    return e ? atoi(e) : 0;
  }();
  static const int max_split = [] {
    const char* e = getenv("UG_SPLITK");            // 0 / 1 disables split-K, n caps it
    return e ? atoi(e) : 16;
  }();
  // Split-K candidates (only when the caller can take fp32 partials: ksplit_out != nullptr): a launch whose tiles
  // cannot fill the SMs (M = 64 .. 1200 rows against K = 11520) runs `waves * iters` iterations on a few CTAs; splitting
  // the (tap, chunk) iterations over S units per tile fills the machine at the price of one reduce pass over S fp32
  // partials.  Costs in SM clocks: per-iteration shared-memory model as before, epilogue 8 clk per column + 400, reduce
  // = its bytes at ~3700 B/clk chip-wide + ~3000 clk of launch latency.
  const long long iters = (long long)a.num_taps * a.kchunks;
  const double M = (double)a.W * a.H * a.N;
  int best = 16;
  double best_cost = -1.0;
  for (int ctas = 1; ctas <= 2; ++ctas) {
    if (force && ctas != force) continue;
    if (ctas == 2 && (m_tiles < 2 || (sms & 1))) continue;
    const int step = a.tma_store ? 64 : 16 * ctas;   // TMA-store slabs are 64 columns wide
    for (int bn = 256; bn >= step; bn -= step) {
      if (a.geglu && bn != 256) continue;      // [128 value | 128 gate] column tiles
      const long long tiles = ((m_tiles + ctas - 1) / ctas) * ((a.n_total + bn - 1) / bn) * batch;
      const long long slots = sms / ctas;
      const long long per = ctas == 1 ? 256 + 2 * bn : (2 * bn > 256 + bn ? 2 * bn : 256 + bn);
      const int s_max = (ksplit_out && batch == 1 && !a.geglu && max_split > 1) ? max_split : 1;
      for (int S = 1; S <= s_max; ++S) {
        if (S > 1 && (iters / S < 4 || tiles * S > 2 * slots || M * a.n_total * 4.0 * S > 96e6)) break;
        const long long units = tiles * S;
        const long long waves = (units + slots - 1) / slots;
        double cost = (double)waves * ((double)per * (double)((iters + S - 1) / S) + 8.0 * bn + 400.0);
        if (S > 1) cost += M * a.n_total * (4.0 * S + 4.0) / 3700.0 + 3000.0;
        if (best_cost < 0 || cost < best_cost) {
          best_cost = cost;
          best = bn;
          *ctas_out = ctas;
          if (ksplit_out) *ksplit_out = S;
        }
      }
    }
  }
  return best;
}

int tapgemm_pick_bn(const TapGemmArgs& a, int batch) {
  int ctas;
  return tapgemm_pick_tile(a, batch, &ctas, nullptr);
}

// Balanced unit lists for launches whose last N tile is ragged (cheaper than a full one) and that run more than one
// wave: units are dealt in their normal order (M- or N-fastest, so the L2 reuse pattern is unchanged) to the CTA
// (pair) that becomes free first under the picker's cost model -- what a dynamic tile counter would do, computed once
// per shape on the host and kept in device memory for the life of the process (stable pointers: CUDA-graph safe).
// Tiles are computed independently, so the assignment cannot change a single output bit (tested).  UG_SCHED=0 = off.
// Pure host part (no CUDA): per-slot unit lists as a [slots][len] table (-1 = none); returns len.  max_cost / rr_max_cost
// (nullable) receive the heaviest slot's modelled cost under this assignment and under plain round-robin.
int tapgemm_build_schedule(int pm_tiles, int n_tiles, int batch, int n_fastest, int n_total, int bn_tile, int ctas,
                           int slots, int iters, std::vector<int>* table, long long* max_cost, long long* rr_max_cost) {
  const long long units = (long long)pm_tiles * n_tiles * batch;
  const int g16 = 16 * ctas;
  auto unit_cost = [&](int t) -> long long {
    int n_tile;
    if (n_fastest) {
      const int per_z = pm_tiles * n_tiles;
      const int u = t % per_z, m_lin = u / n_tiles;
      n_tile = (u - m_lin * n_tiles + m_lin) % n_tiles;
    } else {
      n_tile = (t / pm_tiles) % n_tiles;
    }
    const int rem = (n_total - n_tile * bn_tile + g16 - 1) / g16 * g16;
    const long long bn = rem < bn_tile ? rem : bn_tile;
    const long long per = ctas == 1 ? 256 + 2 * bn : (2 * bn > 256 + bn ? 2 * bn : 256 + bn);
    const long long mma = (long long)iters * per, epi = 8 * bn + 200;      // epilogue overlaps the next tile
    return (mma > epi ? mma : epi) + 48;
  };
  std::vector<std::vector<int>> lists(slots);
  std::vector<long long> rr(slots, 0);
  // (finish time, slot) min-heap; ties go to the lowest slot, so equal costs reproduce the round-robin order
  std::priority_queue<std::pair<long long, int>, std::vector<std::pair<long long, int>>,
                      std::greater<std::pair<long long, int>>> heap;
  for (int s2 = 0; s2 < slots; ++s2) heap.push({0, s2});
  long long worst = 0;
  for (int t = 0; t < (int)units; ++t) {
    auto top = heap.top();
    heap.pop();
    lists[top.second].push_back(t);
    const long long c = unit_cost(t);
    heap.push({top.first + c, top.second});
    if (top.first + c > worst) worst = top.first + c;
    rr[t % slots] += c;
  }
  size_t len = 0;
  for (auto& l : lists) len = l.size() > len ? l.size() : len;
  table->assign((size_t)slots * len, -1);
  for (int s2 = 0; s2 < slots; ++s2)
    for (size_t i = 0; i < lists[s2].size(); ++i) (*table)[(size_t)s2 * len + i] = lists[s2][i];
  if (max_cost) *max_cost = worst;
  if (rr_max_cost) {
    *rr_max_cost = 0;
    for (long long v : rr) *rr_max_cost = v > *rr_max_cost ? v : *rr_max_cost;
  }
  return (int)len;
}

namespace {
struct SchedKey {
  int dev, pm_tiles, n_tiles, batch, n_fastest, n_total, bn, ctas, slots, iters;
  bool operator<(const SchedKey& o) const {
    return std::memcmp(this, &o, sizeof(SchedKey)) < 0;
  }
};
struct SchedVal { int* dev_ptr; int len; };

bool tapgemm_schedule(const TapGemmArgs& a, int ctas, int pm_tiles, int slots, cudaStream_t stream, SchedVal* out) {
  static const bool off = [] { const char* e = getenv("UG_SCHED"); return e && atoi(e) == 0; }();
  const long long units = (long long)pm_tiles * a.n_tiles * a.batch;
  if (off || units <= slots || a.n_total % a.bn_tile == 0 || units > (1 << 22)) return false;
  static std::mutex mu;
  static std::map<SchedKey, SchedVal> cache;
  SchedKey key;
  std::memset(&key, 0, sizeof(key));
  cudaGetDevice(&key.dev);
  key.pm_tiles = pm_tiles; key.n_tiles = a.n_tiles; key.batch = a.batch; key.n_fastest = a.n_fastest;
  key.n_total = a.n_total; key.bn = a.bn_tile; key.ctas = ctas; key.slots = slots;
  key.iters = a.num_taps * a.kchunks;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return it->second.dev_ptr != nullptr; }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return false;                        // never allocate / copy under capture; this launch stays round-robin
  }
  std::vector<int> table;
  const int len = tapgemm_build_schedule(pm_tiles, a.n_tiles, a.batch, a.n_fastest, a.n_total, a.bn_tile, ctas, slots,
                                         key.iters, &table, nullptr, nullptr);
  SchedVal v{nullptr, len};
  if (cudaMalloc(&v.dev_ptr, table.size() * sizeof(int)) != cudaSuccess ||
      cudaMemcpy(v.dev_ptr, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    v.dev_ptr = nullptr;
  }
  cache.emplace(key, v);
  *out = v;
  return v.dev_ptr != nullptr;
}
}  // namespace

#ifdef UG_TAPGEMM_TRACE
static unsigned long long* g_trace_buf = nullptr;
void tapgemm_set_trace(unsigned long long* dev_buf) { g_trace_buf = dev_buf; }
#endif

int launch_tapgemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmC,
                   const TapGemmArgs& args_in, int batch, cudaStream_t stream, const CUtensorMap* tmR,
                   const CUtensorMap* tmBl) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaSuccess;
#define UG_SMEM_ATTR(K, BYTES) if (e == cudaSuccess) e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES)
    UG_SMEM_ATTR((tapgemm_kernel<1, 0>), kSmemBytes);  UG_SMEM_ATTR((tapgemm_kernel<1, 1>), kSmemBytes);
    UG_SMEM_ATTR((tapgemm_kernel<1, 2>), kSmemBytes);  UG_SMEM_ATTR((tapgemm_kernel<2, 0>), kSmemBytes2);
    UG_SMEM_ATTR((tapgemm_kernel<2, 1>), kSmemBytes2); UG_SMEM_ATTR((tapgemm_kernel<2, 2>), kSmemBytes2);
#undef UG_SMEM_ATTR
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  TapGemmArgs args = args_in;
#ifdef UG_TAPGEMM_TRACE
  args.trace = g_trace_buf;
#endif
  const int ctas = args.ctas == 2 ? 2 : 1;
  if (args.bn_tile <= 0 || args.bn_tile > 256 || (args.bn_tile & (16 * ctas - 1))) return (int)cudaErrorInvalidValue;
  if (ctas == 2 && args.b_mn_major) return (int)cudaErrorInvalidValue;
  if (args.tma_store && (tmC == nullptr || (args.bn_tile & 63) || batch != 1 || args.out_fp32))
    return (int)cudaErrorInvalidValue;
  if (args.ksplit > 1 && (batch != args.ksplit || !args.out_fp32 || args.tma_store || args.geglu || args.b_mn_major ||
                          args.bias != nullptr || args.fbias != nullptr || args.res != nullptr || args.blend != nullptr ||
                          args.act != 0 || args.ksplit > args.num_taps * args.kchunks))
    return (int)cudaErrorInvalidValue;             // partial sums only: everything else happens in the reduce pass
  if (args.geglu && (args.bn_tile != 256 || (args.n_total & 255))) return (int)cudaErrorInvalidValue;
  if (args.geglu && (args.res != nullptr || args.blend != nullptr || args.fbias != nullptr || args.scale != 1.0f ||
                     args.act != 0))
    return (int)cudaErrorInvalidValue;             // the GEGLU epilogue is bias + gate only
  if (args.res_tma && (tmR == nullptr || !args.tma_store || args.res == nullptr || args.geglu || (args.n_total & 31)))
    return (int)cudaErrorInvalidValue;
  const bool rows_flat = args.tiles_y == 1 && args.tiles_n == 1 && args.bh == 1 && args.bn == 1 && batch == 1;
  if ((args.ln_stat != nullptr || args.stat_out != nullptr) && (!rows_flat || args.ksplit > 1))
    return (int)cudaErrorInvalidValue;             // LayerNorm fold: linear ops (row = pixel) only
  if (args.ln_stat != nullptr && (args.ln_colsum == nullptr || args.bias == nullptr || args.ln_parts < 0 ||
                                  args.ln_parts > 4 || (args.ln_parts & 1) || args.scale != 1.0f))
    return (int)cudaErrorInvalidValue;
  if (args.stat_out != nullptr && (args.geglu || args.out_fp32 || args.ln_stat != nullptr))
    return (int)cudaErrorInvalidValue;             // a launch is a consumer or a producer of row statistics, never both
  const CUtensorMap& mc = tmC ? *tmC : tmA;
  const CUtensorMap& mr = tmR ? *tmR : tmA;
  if (args.blend_tma && (tmBl == nullptr || !args.res_tma || args.blend == nullptr || args.blend == args.res))
    return (int)cudaErrorInvalidValue;
  const CUtensorMap& mbl = tmBl ? *tmBl : tmA;
  args.batch = batch;
  args.n_tiles = (args.n_total + args.bn_tile - 1) / args.bn_tile;
  args.stat_parts = 2 * args.n_tiles;
  {
    const long long mt = (long long)args.tiles_x * args.tiles_y * args.tiles_n;
    const int pm = (int)((mt + ctas - 1) / ctas);
    args.fd_pm = make_fastdiv(pm);
    args.fd_nt = make_fastdiv(args.n_tiles);
    args.fd_perz = make_fastdiv(pm * args.n_tiles);
    args.fd_tx = make_fastdiv(args.tiles_x);
    args.fd_ty = make_fastdiv(args.tiles_y);
    args.fd_zdiv = make_fastdiv(args.zdiv);
    args.fd_fb = make_fastdiv(args.fbias_div);
    args.fd_kc = make_fastdiv(args.kchunks);
    args.flat = (rows_flat && args.bw == BM && args.zdiv == 1 && args.ksplit <= 1) ? 1 : 0;
  }
  {
    // plain GEMMs whose activation matrix does not survive in L2 (126 MB, shared with the output stream) until the
    // next N pass.  Measured at cfg2: M76800 N320 K1280 100 -> 85 us, M19200 N640 K2560 87 -> 75 us; convolutions
    // (9 taps re-read each tile from L2 anyway, weights 20+ MB) are slower this way and keep the M-fastest order.
    static const char* force = getenv("UG_NFAST");
    const double a_bytes = (double)args.W * args.H * args.N * args.kchunks * 64.0 * 2.0;
    args.n_fastest = force ? atoi(force)
                           : (args.n_tiles > 1 && batch == 1 && args.num_taps == 1 && a_bytes > 40e6) ? 1 : 0;
  }
  const long long m_tiles = (long long)args.tiles_x * args.tiles_y * args.tiles_n;
  const long long units = ((m_tiles + ctas - 1) / ctas) * args.n_tiles * batch;
  if (units <= 0 || units > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  const int sms = tapgemm_num_sms();
  {
    SchedVal sv{nullptr, 0};
    const bool have = tapgemm_schedule(args, ctas, (int)((m_tiles + ctas - 1) / ctas), sms / ctas, stream, &sv);
    args.sched = have ? sv.dev_ptr : nullptr;
    args.sched_len = have ? sv.len : 0;
  }
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                        const TapGemmArgs);
  static const Kern kern1[3] = {tapgemm_kernel<1, 0>, tapgemm_kernel<1, 1>, tapgemm_kernel<1, 2>};
  static const Kern kern2[3] = {tapgemm_kernel<2, 0>, tapgemm_kernel<2, 1>, tapgemm_kernel<2, 2>};
  const int ln_mode = args.ln_stat != nullptr ? 1 : args.stat_out != nullptr ? 2 : 0;
  if (ctas == 1) {
    const int grid = (int)(units < sms ? units : sms);
    return (int)launch_pdl_tag("tapgemm", kern1[ln_mode], dim3(grid), dim3(kThreads), kSmemBytes,
                           stream, tmA, tmB, mc, mr, mbl, args);
  }
  const long long slots = sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * (units < slots ? units : slots)));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes2;
  cfg.stream = stream;
  static const bool no_pdl = pdl_off("tapgemm");
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 1 : 2;
  return (int)cudaLaunchKernelEx(&cfg, kern2[ln_mode], tmA, tmB, mc, mr, mbl, args);
}

}  // namespace ug

// fmha_d64: flash attention on tcgen05, two query tiles per CTA in ping-pong.
//
//   CTA = 256 queries (tiles A and B of 128) of one (frame, head); 320 threads, one CTA per SM:
//     warp 0      TMA producer: Q_A, Q_B once, then K_j / V_j blocks of 128 keys through 2-deep rings
//     warp 1      TMEM owner + UMMA issuer.  Tensor-pipe order per key block j:
//                   [P_A(j) ready] S_A(j+1) = Q_A K_{j+1}^T ; O_A += P_A(j) V_j ;
//                   [P_B(j) ready] S_B(j+1) = Q_B K_{j+1}^T ; O_B += P_B(j) V_j
//                 so a softmax group only ever waits for one 128x128x64 MMA, and the P.V MMAs run under
//                 the other group's exp2 phase
//     warps 2-5   softmax of tile A, warps 6-9 softmax of tile B (thread = query row)
//   TMEM (512 columns): S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384).
//   smem: Q 2x16K | K 2x16K | V 2x16K | P 2 tiles x 2 buffers x 32K = 224 KB.
//
//   Softmax reads S ONCE per block (TMEM read bandwidth, 64 B/clk, and MUFU exp2 are the binding
//   resources): p = 2^(s*scale - m_ref) against a LAZY reference maximum m_ref.  While no row of the warp
//   pushes its row sum past 2^14 (every p then is far inside the 16-bit range) nothing is rescaled; otherwise (and on the
//   first block) the warp takes the exact two-pass route and rescales O in TMEM (tcgen05.ld / st).  O / l is
//   independent of the reference, so results equal the textbook formulation.
//   P goes to smem as the 16-bit K-major A operand; V is consumed in place as an MN-major B operand.
#include "fmha.cuh"
#include "ptx.cuh"

#include <cstdio>
#include <type_traits>
#include <cstdlib>

namespace ug {
namespace {

constexpr int kTile = 16384;                       // 128 rows x 128 B
// K / V rings are 2 deep (measured with in-kernel cycle stamps: loads are never waited for); the smem
// goes to double-buffered P tiles instead, so a group's exp2 phase for block j starts as soon as S(j)
// lands, without waiting for P(j-1).V to finish reading the previous P tile.
constexpr int kNst = 2;
constexpr int kOffQ = 0;                           // Q_A, Q_B
constexpr int kOffK = 2 * kTile;                   // K ring
constexpr int kOffV = kOffK + kNst * kTile;        // V ring
constexpr int kOffP = kOffV + kNst * kTile;        // P[tile][buffer], 2 x kTile each
constexpr int kOffBar = kOffP + 8 * kTile;
constexpr int kSmem = kOffBar + 256;
constexpr int kThreads = 320;
constexpr uint32_t kColS = 0, kColO = 256;
// lazy rescale: a block is taken against the stale reference maximum while its row sums stay below this (every p then
// is below it too: far inside the 16-bit range of the P operand, 65504 in fp16); beyond it the warp takes the exact route
constexpr float kPSumMax = 16384.0f;
constexpr int kPolyDefault = 4;   // measured (L0, 25 x 3072 tokens): 508 us all-MUFU, 462 / 474 / 512 us with every 4th / 3rd / 2nd pair

// 2^x for a pair on the FMA / ALU pipes instead of the MUFU (4 ex2 per clock and SM partition is the kernel's
// co-binding resource next to the TMEM read port): round-to-nearest split x = r + f by the 1.5 * 2^23 trick, degree-3
// minimax polynomial of 2^f on [-0.5, 0.5] (max relative error 7.5e-5, below the 16-bit rounding P gets anyway: 4.9e-4 fp16,
// 3.9e-3 bf16), r added into the exponent field.  x is clamped to [-120, 126]: below, 2^x is 0 in 16 bit anyway;
// above, the result still exceeds the rescale threshold (kPSumMax), which sends the block down the exact route.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fminf(fmaxf(x.x, -120.0f), 126.0f);
  x.y = fminf(fmaxf(x.y, -120.0f), 126.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 j = __fadd2_rn(x, magic);                                   // low mantissa bits = round(x)
  const float2 r = __fadd2_rn(j, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __fadd2_rn(x, make_float2(-r.x, -r.y));
  float2 p = __ffma2_rn(f, make_float2(0.05517132207751274f, 0.05517132207751274f),
                        make_float2(0.24261054396629333f, 0.24261054396629333f));
  p = __ffma2_rn(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  p = __ffma2_rn(p, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(j.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(j.y) << 23)));
}

// POLY: every POLY-th pair of a chunk takes ex2_poly2 instead of two MUFU ex2 (0 = none)
template <int FMT, int POLY>
__global__ void __launch_bounds__(kThreads, 1)
fmha_d64_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ FmhaArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;              // [kNst]
  uint64_t* v_full = k_full + kNst;         // [kNst]
  uint64_t* k_empty = v_full + kNst;        // [kNst]
  uint64_t* v_empty = k_empty + kNst;       // [kNst]
  uint64_t* s_full = v_empty + kNst;        // [2] per query tile
  uint64_t* p_full = s_full + 2;            // [2]
  uint64_t* o_done = p_full + 2;            // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y;
  const int f = blockIdx.z;
  const int row_base = f * a.N;
  const int nb = (a.N + 127) >> 7;

  if ((smem_u32(smem) & 1023u) != 0u) __trap();   // swizzled tiles need a 1024-byte aligned base

  // NO griddepcontrol.launch_dependents here: this kernel's dependents start when it completes.  With the early trigger
  // the 49 x 576 x 1024 denoising step deviated from its own no-PDL result in ~1 run of 10 (max |d| 0.03 on O(1)
  // latents, every element slightly off; never at 25 x 384 x 512 in 120 runs).  Bisected with per-family switches
  // (UG_NO_PDL_K, build variants without the trigger): only the pair "fmha -> attn1.to_out GEMM starting during fmha's
  // last wave" reproduces it, 0 of 100 runs without this trigger; fences after the consumer's wait do not help.  The
  // mechanism is not understood (profiles/r02_pdl_race.txt); the cost of not triggering is the overlap of one GEMM
  // prologue per attention launch (16 per step).
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm);
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < kNst; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&k_empty[i], 1);
        mbar_init(&v_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 128);
        mbar_init(&o_done[i], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * kTile);
      tma_load_2d(smem + kOffQ, &tm, q_full, h * 64, row_base + q0);
      tma_load_2d(smem + kOffQ + kTile, &tm, q_full, h * 64, row_base + q0 + 128);
      for (int j = 0; j < nb; ++j) {
        const int s = j % kNst;
        const uint32_t ph = (uint32_t)(j / kNst) & 1u;
        mbar_wait(&k_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&k_full[s], kTile);
        tma_load_2d(smem + kOffK + s * kTile, &tm, &k_full[s], a.C + h * 64, row_base + j * 128);
        mbar_wait(&v_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&v_full[s], kTile);
        tma_load_2d(smem + kOffV + s * kTile, &tm, &v_full[s], 2 * a.C + h * 64, row_base + j * 128);
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    const uint32_t idesc_qk = make_idesc_f16(128, 128, a.fmt, 0);
    const uint32_t idesc_pv = make_idesc_f16(128, 64, a.fmt, 1);
    const uint32_t sQ = smem_u32(smem + kOffQ), sP = smem_u32(smem + kOffP);
    // S_x(j) = Q_x K_j^T into tile x's S columns
    auto issue_qk = [&](int x, int j) {
      if (elect_one()) {
        const uint64_t da = make_desc_kmajor_sw128(sQ + x * kTile);
        const uint64_t db = make_desc_kmajor_sw128(smem_u32(smem + kOffK + (j % kNst) * kTile));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + kColS + x * 128, da + 2 * k, db + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&s_full[x]);
        if (x == 1) umma_commit(&k_empty[j % kNst]);    // both tiles have consumed K_j
      }
      __syncwarp();
    };
    // O_x += P_x(j) V_j
    auto issue_pv = [&](int x, int j) {
      if (elect_one()) {
        const uint64_t dv = make_desc_mnmajor_sw128(smem_u32(smem + kOffV + (j % kNst) * kTile), 8192);
        const uint32_t pbase = sP + (uint32_t)(x * 2 + (j & 1)) * 2 * kTile;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A: P half (k >> 2), +32 B per 16 keys inside the swizzle row; B: 16 key rows = 2048 B
          const uint64_t dp = make_desc_kmajor_sw128(pbase + (k >> 2) * kTile) + 2 * (k & 3);
          umma_f16(tmem_base + kColO + x * 64, dp, dv + 128 * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&o_done[x]);
        if (x == 1) umma_commit(&v_empty[j % kNst]);    // both tiles have consumed V_j
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    issue_qk(0, 0);
    issue_qk(1, 0);
    for (int j = 0; j < nb; ++j) {
#pragma unroll 1
      for (int x = 0; x < 2; ++x) {
        mbar_wait(&p_full[x], (uint32_t)j & 1u);        // P_x(j) in smem, S_x drained, O_x rescaled
        tc_fence_after();
        if (j + 1 < nb) {
          if (x == 0) {
            mbar_wait(&k_full[(j + 1) % kNst], (uint32_t)((j + 1) / kNst) & 1u);
            tc_fence_after();
          }
          issue_qk(x, j + 1);
        }
        if (x == 0) {
          mbar_wait(&v_full[j % kNst], (uint32_t)(j / kNst) & 1u);
          tc_fence_after();
        }
        issue_pv(x, j);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (tile x = 0: warps 2-5, x = 1: warps 6-9) =====
    const int x = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;                        // query row in the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t colS = kColS + (uint32_t)x * 128u, colO = kColO + (uint32_t)x * 64u;
    const uint32_t sPx = smem_u32(smem + kOffP) + (uint32_t)x * 4 * kTile;
    const float sc = a.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    // p = 2^(s*scale - mref) for the 128 keys of the block, 32 columns at a time, written straight to the
    // P tile (K-major SWIZZLE_128B: row r, 16-byte chunk c16 of half hh at ((c16 ^ (r & 7)) * 16))
    // The TMEM load of chunk c+1 is in flight while chunk c goes through the MUFU (both cost ~256 clk per
    // chunk and SM; serialised they were the whole kernel time).
    // MASK = the block has fewer than 128 real keys (last block of a ragged sequence): a compile-time branch, so
    // that full blocks carry no per-element compare / select
    auto p_chunk = [&](auto mask_tag, const uint32_t (&v)[32], int c, uint32_t sP, float mref, int valid,
                       float& rowsum, float& pmax) {
      constexpr bool MASK = decltype(mask_tag)::value;
      uint32_t pk[16];
      // packed fp32 pipe (FFMA2 / FADD2): the softmax warps are issue bound (ncu: 52 % issue slots with 2.5 warps
      // per scheduler), so the scale-and-shift and the row sum take half an instruction per element
      const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-mref, -mref);
      float2 rs2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2);
        float2 pp;
        if (POLY > 0 && ((i >> 1) % (POLY > 0 ? POLY : 1)) == (POLY > 1 ? 1 : 0)) pp = ex2_poly2(t);
        else pp = make_float2(ex2_approx(t.x), ex2_approx(t.y));
        if constexpr (MASK) {
          if (c * 32 + i >= valid) pp.x = 0.f;
          if (c * 32 + i + 1 >= valid) pp.y = 0.f;
        }
        rs2 = __fadd2_rn(rs2, pp);
        pk[i >> 1] = FMT ? Elem<__nv_bfloat16>::pack2(pp.x, pp.y) : Elem<__half>::pack2(pp.x, pp.y);
      }
      rowsum += rs2.x + rs2.y;
      // The rescale test needs no maximum of its own: every p is >= 0, so the block's row sum bounds each of them
      // (and inf / nan survive the sum).  The caller compares the sum with kPSumMax; a packed max per pair was half an
      // instruction per score for the same decision.
      pmax = rowsum;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int c16 = c * 4 + q4;
        const uint32_t addr = sP + (uint32_t)(c16 >> 3) * kTile + (uint32_t)r * 128u +
                              (uint32_t)(((c16 & 7) ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[q4 * 4 + 0]),
                     "r"(pk[q4 * 4 + 1]), "r"(pk[q4 * 4 + 2]), "r"(pk[q4 * 4 + 3])
                     : "memory");
      }
    };
    auto compute_p_t = [&](auto mask_tag, uint32_t sP, float mref, int valid, float& rowsum, float& pmax) {
      rowsum = 0.f;
      pmax = 0.f;
      uint32_t va[32], vb[32];
      tmem_ld_32x32(lane_addr + colS, va);
      tmem_ld_wait();
      tmem_ld_32x32(lane_addr + colS + 32, vb);
      p_chunk(mask_tag, va, 0, sP, mref, valid, rowsum, pmax);
      tmem_ld_wait();
      tmem_ld_32x32(lane_addr + colS + 64, va);
      p_chunk(mask_tag, vb, 1, sP, mref, valid, rowsum, pmax);
      tmem_ld_wait();
      tmem_ld_32x32(lane_addr + colS + 96, vb);
      p_chunk(mask_tag, va, 2, sP, mref, valid, rowsum, pmax);
      tmem_ld_wait();
      p_chunk(mask_tag, vb, 3, sP, mref, valid, rowsum, pmax);
    };
    auto compute_p = [&](uint32_t sP, float mref, int valid, float& rowsum, float& pmax) {
      if (valid == 128) compute_p_t(std::false_type{}, sP, mref, valid, rowsum, pmax);
      else compute_p_t(std::true_type{}, sP, mref, valid, rowsum, pmax);
    };
    for (int j = 0; j < nb; ++j) {
      mbar_wait(&s_full[x], (uint32_t)j & 1u);
      tc_fence_after();
      // S_x(j) complete => every MMA issued before it has retired, in particular P_x(j-2) V_{j-2}: P buffer
      // (j & 1) is free.  P_x(j-1) V_{j-1} may still be running; only the rescale path waits for it.
      const uint32_t sP = sPx + (uint32_t)(j & 1) * 2 * kTile;
      const int valid = min(128, a.N - j * 128);        // keys of this block that exist
      float rowsum, pmax, alpha = 1.0f;
      bool exact = (j == 0);
      if (!exact) {
        compute_p(sP, m_ref, valid, rowsum, pmax);
        exact = __any_sync(0xffffffffu, !(pmax <= kPSumMax));   // also catches inf / nan
      }
      if (exact) {
        // ---- exact route: row max of this block first, then p against the updated reference
        float mraw = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(lane_addr + colS + c * 32, v);
          tmem_ld_wait();
          if (valid == 128) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mraw = fmaxf(mraw, __uint_as_float(v[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < valid) mraw = fmaxf(mraw, __uint_as_float(v[i]));
          }
        }
        const float mx = fmaxf(m_ref, mraw * sc);        // sc > 0
        alpha = ex2_approx(m_ref - mx);                   // m_ref = -inf on the first block -> 0
        m_ref = mx;
        compute_p(sP, m_ref, valid, rowsum, pmax);
        l = l * alpha + rowsum;
        if (j > 0) {                                     // O_x *= alpha once P_x(j-1) V_{j-1} has retired
          mbar_wait(&o_done[x], (uint32_t)(j - 1) & 1u);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(lane_addr + colO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(lane_addr + colO + c * 32, o);
          }
          tmem_st_wait();
        }
      } else {
        l += rowsum;
      }
      fence_proxy_async_smem();                          // P visible to the tensor core's async proxy
      tc_fence_before();
      mbar_arrive(&p_full[x]);
    }
    // ---- epilogue: O / l -> out
    mbar_wait(&o_done[x], (uint32_t)(nb - 1) & 1u);
    tc_fence_after();
    const float inv = 1.0f / l;
    const int qrow = q0 + x * 128 + r;
    const bool ok = qrow < a.N;
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(a.out) +
                                          ((long long)(row_base + qrow) * a.C + h * 64));
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(lane_addr + colO + c * 32, o);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float x0 = __uint_as_float(o[g * 8 + e * 2]) * inv, x1 = __uint_as_float(o[g * 8 + e * 2 + 1]) * inv;
            w[e] = FMT ? Elem<__nv_bfloat16>::pack2(x0, x1) : Elem<__half>::pack2(x0, x1);
          }
          u.x = w[0]; u.y = w[1]; u.z = w[2]; u.w = w[3];
          dst[c * 4 + g] = u;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int launch_fmha_d64(const CUtensorMap& tm, const FmhaArgs& args, cudaStream_t stream) {
  using Kern = void (*)(const CUtensorMap, const FmhaArgs);
  // UG_FMHA_POLY = n: every n-th pair of exponentials on the FMA pipe (0 = all on the MUFU); default kPolyDefault
  static const int poly = [] { const char* e = getenv("UG_FMHA_POLY"); return e ? atoi(e) : kPolyDefault; }();
  static Kern table[2][5] = {{fmha_d64_kernel<0, 0>, nullptr, fmha_d64_kernel<0, 2>, fmha_d64_kernel<0, 3>, fmha_d64_kernel<0, 4>},
                             {fmha_d64_kernel<1, 0>, nullptr, fmha_d64_kernel<1, 2>, fmha_d64_kernel<1, 3>, fmha_d64_kernel<1, 4>}};
  static bool configured = false;
  if (!configured) {
    for (int f = 0; f < 2; ++f)
      for (int q = 0; q < 5; ++q)
        if (table[f][q] != nullptr) {
          cudaError_t e = cudaFuncSetAttribute(table[f][q], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
          if (e != cudaSuccess) return (int)e;
        }
    configured = true;
  }
  if (args.C != args.heads * 64 || (args.C & 7)) return (int)cudaErrorInvalidValue;
  const int pq = (poly >= 2 && poly <= 4) ? poly : 0;
  dim3 grid((args.N + 255) / 256, args.heads, args.F);
  return (int)launch_pdl_tag("fmha", table[args.fmt ? 1 : 0][pq], grid, dim3(kThreads), kSmem, stream, tm, args);
}

}  // namespace ug

// fmha_d64: flash attention on tcgen05.
//   CTA = one 128-query tile of one (frame, head); 192 threads:
//     warp 0     TMA producer: Q once, then K_j / V_j blocks of 128 keys through 2-deep rings
//     warp 1     TMEM owner + UMMA issuer:  S = Q K_j^T (128x128x64)  and  O += P_j V_j (128x64x128)
//     warps 2-5  softmax: thread = query row; S read ONCE per block with tcgen05.ld (TMEM read bandwidth,
//                64 B/clk, is the binding resource), exp2 against a lazily updated reference maximum,
//                P_j written to smem as the 16-bit K-major A operand of the P.V MMA, O rescaled in TMEM
//                (tcgen05.ld / st) only when a row maximum jumps by more than 2^8
//   TMEM: S at columns [0,128), O at [128,192).  smem: Q 16K | K 16K | V 16K | P 32K = 80 KB, so
//   two CTAs share an SM: one CTA's exp2 phase (MUFU bound) overlaps the other's MMAs.
//   V is consumed in place as an MN-major B operand (no transpose anywhere).
#include "fmha.cuh"
#include "ptx.cuh"

#include <cstdio>
#include <cstdlib>

namespace ug {
namespace {

constexpr int kTile = 16384;                       // 128 rows x 128 B
// K and V are single-buffered: K_{j+1} is requested as soon as S_j = Q K_j^T has been issued and V_j
// as soon as P_{j-1} V_{j-1} retires, both a whole softmax phase before they are needed -- and the
// 80 KB footprint is what lets two CTAs share an SM.
constexpr int kNst = 1;
constexpr int kOffQ = 0, kOffK = kTile, kOffV = kOffK + kNst * kTile, kOffP = kOffV + kNst * kTile,
              kOffBar = kOffP + 2 * kTile;
constexpr int kSmem = kOffBar + 128;
constexpr int kThreads = 192;
constexpr uint32_t kColS = 0, kColO = 128;

__global__ void __launch_bounds__(kThreads, 2)
fmha_d64_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ FmhaArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* k_empty = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_full = bars + 10;
  uint64_t* o_done = bars + 11;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int f = blockIdx.z;
  const int row_base = f * a.N;
  const int nb = (a.N + 127) >> 7;

  if ((smem_u32(smem) & 1023u) != 0u) __trap();   // swizzled tiles need a 1024-byte aligned base

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm);
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&k_empty[i], 1);
        mbar_init(&v_empty[i], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(p_full, 128);
      mbar_init(o_done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, kTile);
      tma_load_2d(smem + kOffQ, &tm, q_full, h * 64, row_base + q0);
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
    auto issue_qk = [&](int j) {
      const int s = j % kNst;
      mbar_wait(&k_full[s], (uint32_t)(j / kNst) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = make_desc_kmajor_sw128(sQ);
        const uint64_t db = make_desc_kmajor_sw128(smem_u32(smem + kOffK + s * kTile));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base + kColS, da + 2 * k, db + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&k_empty[s]);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < nb; ++j) {
      const int s = j % kNst;
      mbar_wait(p_full, (uint32_t)j & 1u);              // P_j in smem, O rescaled
      mbar_wait(&v_full[s], (uint32_t)(j / kNst) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dv = make_desc_mnmajor_sw128(smem_u32(smem + kOffV + s * kTile), 8192);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A: P half (k >> 2), +32 B per 16 keys inside the swizzle row; B: 16 key rows = 2048 B
          const uint64_t dp = make_desc_kmajor_sw128(sP + (k >> 2) * kTile) + 2 * (k & 3);
          umma_f16(tmem_base + kColO, dp, dv + 128 * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[s]);
        umma_commit(o_done);
      }
      __syncwarp();
      if (j + 1 < nb) issue_qk(j + 1);
    }
  } else {
    // ===================== softmax / correction / epilogue =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;                        // query row in the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t sP = smem_u32(smem + kOffP);
    const float sc = a.scale_log2;
    // Online softmax with a LAZY reference maximum: p = 2^(s*scale - m_ref) where m_ref is the row maximum
    // known from earlier blocks.  As long as no row of the warp exceeds m_ref by more than 8 (p <= 256,
    // exact in 16 bit) the block needs ONE pass over S in TMEM and no rescale of O; only when a row jumps
    // past the threshold (and on the first block) the warp takes the exact two-pass route and rescales.
    // O / l at the end is independent of the reference, so results match the exact formulation.
    float m_ref = -INFINITY, l = 0.f;
    // p = 2^(s*scale - mref) for the 128 keys of the block, 32 columns at a time, written straight to the
    // P tile in smem (K-major SWIZZLE_128B: row r, 16-byte chunk c16 of half hh at ((c16 ^ (r & 7)) * 16))
    auto compute_p = [&](float mref, int valid, float& rowsum, float& pmax) {
      rowsum = 0.f;
      pmax = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + kColS + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), sc, -mref));
          float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, -mref));
          if (valid != 128) {
            if (c * 32 + i >= valid) p0 = 0.f;
            if (c * 32 + i + 1 >= valid) p1 = 0.f;
          }
          rowsum += p0 + p1;
          pmax = fmaxf(pmax, fmaxf(p0, p1));
          pk[i >> 1] = a.fmt ? Elem<__nv_bfloat16>::pack2(p0, p1) : Elem<__half>::pack2(p0, p1);
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int c16 = c * 4 + q4;
          const uint32_t addr = sP + (uint32_t)(c16 >> 3) * kTile + (uint32_t)r * 128u +
                                (uint32_t)(((c16 & 7) ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[q4 * 4 + 0]),
                       "r"(pk[q4 * 4 + 1]), "r"(pk[q4 * 4 + 2]), "r"(pk[q4 * 4 + 3])
                       : "memory");
        }
      }
    };
    for (int j = 0; j < nb; ++j) {
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      // S_j complete implies P_{j-1} V_{j-1} (issued before Q K_j^T) has retired: the P tile and O are free
      if (j > 0) {
        mbar_wait(o_done, (uint32_t)(j - 1) & 1u);
        tc_fence_after();
      }
      const int valid = min(128, a.N - j * 128);        // keys of this block that exist
      float rowsum, pmax, alpha = 1.0f;
      bool exact = (j == 0);
      if (!exact) {
        compute_p(m_ref, valid, rowsum, pmax);
        exact = __any_sync(0xffffffffu, !(pmax <= 256.0f));   // also catches inf / nan
      }
      if (exact) {
        // ---- exact route: row max of this block first, then p against the updated reference
        float mraw = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(lane_addr + kColS + c * 32, v);
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
        compute_p(m_ref, valid, rowsum, pmax);
        l = l * alpha + rowsum;
      } else {
        l += rowsum;
      }
      // ---- O *= alpha (only on the exact route; warp-uniform)
      if (exact && j > 0) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          tmem_ld_32x32(lane_addr + kColO + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32(lane_addr + kColO + c * 32, o);
        }
        tmem_st_wait();
      }
      fence_proxy_async_smem();                          // P visible to the tensor core's async proxy
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> out
    mbar_wait(o_done, (uint32_t)(nb - 1) & 1u);
    tc_fence_after();
    const float inv = 1.0f / l;
    const bool ok = (q0 + r) < a.N;
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(a.out) +
                                          ((long long)(row_base + q0 + r) * a.C + h * 64));
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(lane_addr + kColO + c * 32, o);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float x0 = __uint_as_float(o[g * 8 + e * 2]) * inv, x1 = __uint_as_float(o[g * 8 + e * 2 + 1]) * inv;
            w[e] = a.fmt ? Elem<__nv_bfloat16>::pack2(x0, x1) : Elem<__half>::pack2(x0, x1);
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
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

int launch_fmha_d64(const CUtensorMap& tm, const FmhaArgs& args, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fmha_d64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return (int)e;
    // two CTAs per SM need (almost) the whole 228 KB as shared memory
    e = cudaFuncSetAttribute(fmha_d64_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    if (getenv("UG_DEBUG")) {
      int nb = 0, dev = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0;
      cudaGetDevice(&dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fmha_d64_kernel, kThreads, kSmem);
      cudaDeviceGetAttribute(&v1, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
      cudaDeviceGetAttribute(&v2, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      cudaDeviceGetAttribute(&v3, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
      cudaDeviceGetAttribute(&v4, cudaDevAttrReservedSharedMemoryPerBlock, dev);
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, fmha_d64_kernel);
      fprintf(stderr, "[unigeo_b200] fmha_d64: %d CTAs/SM (smem %d B) smem/SM %d optin %d regs/SM %d reserved %d | "
              "kernel regs %d static smem %zu maxdyn %d\n", nb, kSmem, v1, v2, v3, v4, fa.numRegs,
              fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes);
      for (int sm = 0; sm <= 112 * 1024; sm += 16 * 1024) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fmha_d64_kernel, kThreads, sm);
        fprintf(stderr, "   dyn smem %6d -> %d CTAs/SM\n", sm, nb);
      }
    }
    configured = true;
  }
  if (args.C != args.heads * 64 || (args.C & 7)) return (int)cudaErrorInvalidValue;
  dim3 grid((args.N + 127) / 128, args.heads, args.F);
  fmha_d64_kernel<<<grid, kThreads, kSmem, stream>>>(tm, args);
  return (int)cudaGetLastError();
}

}  // namespace ug

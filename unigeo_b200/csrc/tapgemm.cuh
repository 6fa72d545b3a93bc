// tapgemm: the one tensor-core workhorse of the hot path.
//
//   D[pixel, cout] = sum_{tap} sum_{cin} A[pixel shifted by tap, cin] * W[tap][cout][cin]
//
// A is a 16-bit channels-last activation tensor described by a <=5-D TMA tensor map;
// every (tap, 64-channel chunk) is ONE box load whose out-of-bounds elements TMA
// zero-fills, which is exactly the zero padding of a 3x3 conv, of the (3,1,1)
// temporal conv at clip edges, and of the K tail.  With one tap it is a plain
// (batched) GEMM  D = A * W^T.  Accumulation runs on tcgen05 (UMMA 128 x BN x 16,
// fp32 in TMEM); the epilogue reads TMEM with tcgen05.ld and fuses bias,
// per-frame bias (time embedding / collapsed cross-attention), residual add,
// AlphaBlender mixing and GEGLU gating before the store.
//
// Replaces (reference -> [UPSTREAM] diffusers): cuDNN Conv2d/Conv3d
// implicit GEMMs and cuBLASLt linear layers of the SVD UNet / temporal VAE.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace ug {

constexpr int kMaxTaps = 9;

// Division of a non-negative int (< 2^31) by a launch constant without the ~100-clock integer division sequence: the
// persistent kernel decodes a unit index into (M tile, N tile, batch, x, y, n, frame) once per tile and role, ten
// dependent divisions that made up most of a 1650-clock tile prologue in front of an epilogue-bound 6300-clock period
// (tools/trace_tapgemm.py ... split).  q = (n * m) >> (31 + s), s = ceil(log2 d), m = ceil(2^(31+s) / d): exact for
// every n < 2^31 (checked on the host against n / d, tests/test_host_logic.py).
struct FastDiv {
  unsigned int d, m, sh;     // d <= 1: identity
#if defined(__CUDACC__)
  __device__ __forceinline__ int div(int n) const { return d <= 1u ? n : (int)(__umulhi((unsigned int)n, m) >> sh); }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const { q = div(n); r = n - q * (int)d; }
#endif
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d < 1 ? 1u : (unsigned int)d;
  f.m = 0u;
  f.sh = 0u;
  if (f.d > 1u) {
    unsigned int s = 0;
    while ((1ull << s) < f.d) ++s;
    f.m = (unsigned int)(((1ull << (31 + s)) + f.d - 1) / f.d);
    f.sh = s - 1;
  }
  return f;
}
inline int fastdiv_host(const FastDiv& f, int n) {       // the device arithmetic, for the host-side check
  return f.d <= 1u ? n : (int)((((unsigned long long)(unsigned int)n * f.m) >> 32) >> f.sh);
}

struct TapGemmArgs {
  // ---- M tiling over the OUTPUT pixel grid (n, y, x); pixel = (n*H + y)*W + x
  int tiles_x, tiles_y, tiles_n;
  int bw, bh, bn;          // box extent: rows per tile = bw*bh*bn <= 128
  int W, H, N;
  int dim_x, dim_y, dim_n; // which A-map coordinate (1..4) carries x / y / n (5 = unused)
  // batch index z = blockIdx.z splits as z1 = z / zdiv, z0 = z % zdiv (e.g. frame, head):
  //   A coordinate[dim_z1] += z1 * a_z1step ; A coordinate[dim_z0] += z0 * a_z0step
  // (coordinate 0 is the channel coordinate; use dim 5 for "none")
  int zdiv, dim_z1, a_z1step, dim_z0, a_z0step;
  // ---- K loop
  int num_taps, kchunks;   // iterations = num_taps * kchunks, 64 channels each
  int tap_off[kMaxTaps][5];
  int b_tap_rows;          // B row offset per tap
  int b_z1_rowstep;        // B row offset per z1
  int b_z0_cstep;          // B column (dim-0) offset per z0
  int b_c0;                // B column (dim-0) base offset
  int b_mn_major;          // B tile is [K rows][64 N elements] (e.g. V of attention): needs BN == 64
  // ---- epilogue
  int n_total;             // valid output columns (before GEGLU halving)
  int bn_tile;             // N extent of one tile: multiple of 16 * ctas, <= 256 (tapgemm_pick_tile)
  int ctas;                // 1: 128 x BN tile per CTA; 2: 256 x BN tile per CTA pair (cta_group::2)
  int tma_store;           // finished 64-column slabs leave through smem + TMA tile stores (coalesced);
                           // needs a 16-bit output, ldc % 8 == 0, batch 1 and the output tensor map
  int res_tma;             // residual tiles arrive by TMA (32-column chunks, in place in the store staging ring):
                           // tmC / tmR are 32-column SWIZZLE_64B maps of the output / residual tensor
  int blend_tma;           // (res_tma launches whose AlphaBlender input is a different tensor than the residual) the blend
                           // tile arrives by TMA as well: the ring works as two (residual, blend) buffer pairs, tmBl = its map
  int ksplit;              // > 1: split-K -- batch == ksplit units per tile, each accumulating its share of the (tap,
                           // chunk) iterations into an fp32 partial at out + z * out_z1stride (direct stores, no fusions)
  int n_tiles, batch;      // filled by launch_tapgemm
  int n_fastest;           // tile order (filled by launch_tapgemm): N tiles of one M tile run concurrently
#ifdef UG_TAPGEMM_TRACE
  unsigned long long* trace;   // debug builds (tools/trace_tapgemm.py): [grid units][4 roles][kTraceTiles][4] clock64 stamps
#endif
  // filled by launch_tapgemm: divisors of the per-tile index decoding
  FastDiv fd_pm, fd_nt, fd_perz;   // M units per N tile (pairs count once), N tiles, their product
  FastDiv fd_tx, fd_ty, fd_zdiv, fd_fb, fd_kc;   // tiles_x, tiles_y, zdiv, fbias_div, kchunks
  int flat;                // filled by launch_tapgemm: plain GEMM rows (linear ops, batch 1): pixel = M tile * 128 + row,
                           // the epilogue skips the (x, y, n, z) decoding of a tile
  const int* sched;        // filled by launch_tapgemm: [grid units][sched_len] unit indices (-1 = none) when the host
  int sched_len;           // balanced ragged-width tiles over the CTAs (list scheduling); nullptr = round-robin
  int fmt;                 // 0 = fp16, 1 = bf16 (operands and 16-bit outputs)
  int out_fp32;
  int act;                 // 1: exact (erf) GELU on (acc * scale + bias) before residual / blend (ViT MLPs)
  int geglu;               // columns [0,128) of each 256-wide tile gate-multiplied by gelu([128,256))
  void* out;
  long long ldc;
  long long out_z1stride, out_z0stride;   // element offsets per z1 / z0
  const float* bias;       // [n_total]
  const float* fbias;      // [frames][fbias_ld], frame = pixel / fbias_div
  int fbias_ld, fbias_div;
  int fbias_uniform;       // every 128-row tile lies inside one frame (linear ops with fbias_div % 128 == 0):
                           // the frame bias is folded into the per-tile bias vector staged in smem
  const void* res;         // 16-bit residual, same pixel indexing
  long long ldr;
  const void* blend;       // out = alpha * blend + (1 - alpha) * value
  long long ldb;
  float alpha;
  float scale;             // value = acc * scale + ...
  // ---- LayerNorm folded into this GEMM (linear ops; rows = pixels).  The A operand is the RAW row x, the weights are
  // W' = gamma (.) W, and the epilogue applies   rstd[row] * (acc - mean[row] * colsum[col]) + bias'[col]
  // with colsum[n] = sum_k W'[n][k] and bias' = bias + W beta: exactly LayerNorm(x) W^T + bias, without a normalised
  // copy of x ever existing.  ln_stat == nullptr: off.
  const float2* ln_stat;   // ln_parts == 0: [row] = (mean, rstd).  ln_parts = P > 0: [row * P + p] = (sum, sum of squares)
  int ln_parts;            //   partials of the row as its producing GEMM's epilogue left them, folded here in index order
  float ln_inv_c, ln_eps;  //   (1 / row length, epsilon) for the partial form
  const float* ln_colsum;  // [n_total] (GEGLU: in the interleaved column order of the weights)
  // ---- producer side of the same fusion: every finished output row leaves its (sum, sum of squares) over the columns
  // this CTA's column group handled at stat_out[row * stat_parts + n_tile * 2 + group] (fp32, pre-rounding values)
  float2* stat_out;
  int stat_parts;          // = 2 * n_tiles (filled by launch_tapgemm)
};

#ifdef UG_TAPGEMM_TRACE
constexpr int kTraceTiles = 16;
// the buffer every following launch stamps (nullptr = off); set through ug_debug_tapgemm_trace (trace builds only)
void tapgemm_set_trace(unsigned long long* dev_buf);
#endif

// Host-side helpers -----------------------------------------------------------------
struct TmapDesc {
  const void* ptr;
  int elem_fmt;                 // 0 fp16, 1 bf16
  int rank;                     // 2..5
  unsigned long long dims[5];
  unsigned long long strides[4];   // bytes, dims 1..rank-1
  unsigned int box[5];
  int swizzle64 = 0;            // 0: SWIZZLE_128B (64-element inner box), 1: SWIZZLE_64B (32-element inner box)
};

// Encodes a SWIZZLE_128B tiled tensor map; returns cudaSuccess-like 0 or non-zero.
int encode_tmap(CUtensorMap* out, const TmapDesc& d);

// Launch (persistent, one CTA per SM); args.bn_tile must be set (tapgemm_pick_bn) and must equal the
// row extent of the B tensor map's box.  Returns cudaError_t as int.
int launch_tapgemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmC,
                   const TapGemmArgs& args, int batch, cudaStream_t stream, const CUtensorMap* tmR = nullptr,
                   const CUtensorMap* tmBl = nullptr);

// needs tiles_*, n_total, geglu, b_mn_major filled in; returns bn_tile and the CTA count per tile.
// The B tensor map's box must have bn_tile / ctas rows.
// ksplit (nullable): where non-null the model may also split the K iterations (see TapGemmArgs::ksplit) and returns S.
int tapgemm_pick_tile(const TapGemmArgs& args, int batch, int* ctas, int* ksplit = nullptr);
// Balanced per-CTA unit lists for launches with ragged N tiles (pure host code): [slots][len] table, -1 = none;
// returns len.  max_cost / rr_max_cost (nullable): modelled cost of the heaviest slot, here and under round-robin.
int tapgemm_build_schedule(int pm_tiles, int n_tiles, int batch, int n_fastest, int n_total, int bn_tile, int ctas,
                           int slots, int iters, std::vector<int>* table, long long* max_cost, long long* rr_max_cost);
int tapgemm_pick_bn(const TapGemmArgs& args, int batch);
int tapgemm_num_sms();

}  // namespace ug

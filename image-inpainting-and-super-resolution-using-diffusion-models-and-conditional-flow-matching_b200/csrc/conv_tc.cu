// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
//   D[m, n] = sum_{segment, tap, c} A_seg[pixel(m) + tap, c] * W[kiter][n][c]      bf16 x bf16 -> fp32
//
// * M = output pixels.  A CTA tile is 128 pixels = a (bn x bh x bw) box of the NHWC tensor, so one
//   4-D TMA box load per (tap, 64-channel chunk) lands as 128 rows x 128 B in shared memory in
//   exactly the K-major SWIZZLE_128B layout tcgen05.mma consumes.  The 3x3 halo is produced by
//   shifting the box origin by the tap offset; out-of-bounds rows/cols (the zero padding) are
//   zero-filled by the TMA unit.  Stride-2 convs use the tensor map's element strides.
// * K runs over up to three "segments": the main 3x3 (or 1x1) operand and the ResBlock's 1x1 skip
//   operand(s), whose products accumulate into the same TMEM tile - so `skip_connection(x) + h`
//   and the never-materialised channel concat `cat([h, hs.pop()])` cost no extra pass.
// * Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
//   warps 2..5 = epilogue (TMEM -> registers -> +bias +emb +residual -> bf16 -> global).  Two TMEM
//   accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cstdio>
#include <cstring>
#include <map>
#include "engine.h"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cfm {

void* tensor_ptr(const Engine& e, int id, int B);

constexpr int TC_BLOCK_M = 128;         // rows of one UMMA; a CTA tile is 128 * mh rows (mh = 1 or 2 M-halves)
constexpr int TC_BLOCK_K = 64;          // widest K-iteration: 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;   // 16 KB per M-half
#ifndef CFM_TC_EPI_WARPS
#define CFM_TC_EPI_WARPS 8
#endif
constexpr int TC_EPI_WARPS = CFM_TC_EPI_WARPS;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;        // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue
constexpr int TC_MAX_COUT = 2048;                         // bias staged in smem
constexpr int TC_MAX_A = 8, TC_MAX_B = 16;                // ring depths (A slots, B slots)
constexpr int TC_RING_BYTES = 216 * 1024;                 // A ring + B ring
constexpr int TC_STAGE_BUF = 16 * 1024;                   // TMA-store epilogue: one 128-row x 64-channel bf16 box
constexpr int TC_STAGE_BYTES = 4 * TC_STAGE_BUF;          // two per epilogue warp group, taken from the top of the rings
constexpr int TC_STAGE_OFF = TC_RING_BYTES - TC_STAGE_BYTES;
// Fused GroupNorm epilogue (kGN): 12 KB taken from the top of the rings -
//   [0, 2K)   mean / rstd per (sample of the tile, group): float2 [8][32]
//   [2K, 4K)  gamma [256] | beta [256]
//   [4K, 8K)  per-warp partial sums: float2 [chunk <= 8][block of rows <= 8][granule 8]
//   [8K, 12K) per-(sample of the tile, channel) scale / shift of pass 2: float2 [2][256] (tiles of at most two samples)
constexpr int TC_GN_BYTES = 12 * 1024;
constexpr int TC_GN_OFF = TC_RING_BYTES - TC_GN_BYTES;
constexpr int TC_GN_MAX_COUT = 256;
constexpr int TC_BAR_BYTES = 512;
constexpr int TC_SMEM_BYTES = TC_RING_BYTES + TC_MAX_COUT * 4 + TC_BAR_BYTES + 1024 /*align*/;

// One K segment: the main 3x3 / 1x1 operand or a 1x1 skip operand.
//   halo = 0: one A tile per (tap, 64-channel chunk), shifted by the tap offset.
//   halo = 1: one A tile per (chunk, dx): a box of (tile rows + 2) image rows; the taps along y are row-shifted
//             views of that box (descriptor start + dy * bw * 128 B - a multiple of the 1024 B swizzle atom), so
//             a 3x3 conv pulls its activations from L2 three times instead of nine.  The operand stream from L2
//             (not the tensor pipe) is what bounds the N <= 128 layers: 96 B/clk/SM without this, ~57 with.
struct TcSeg { int map; int n_chunks; int ks; int stride; int halo; };

// n / d for n < 2^31 as one multiply-high and a shift (the divisors are tile counts and box extents known at launch;
// the epilogue of a short-K tile is only a few hundred instructions, so a handful of runtime divisions per tile show)
struct FastDiv {
  uint32_t d, mul, shr;
  __host__ void init(int dd) {
    d = (uint32_t)dd; mul = 0; shr = 0;
    if (dd > 1) {
      int lg = 0; while ((1u << lg) < d) ++lg;
      const int pw = 31 + lg;
      mul = (uint32_t)((((unsigned long long)1 << pw) + d - 1) / d);
      shr = (uint32_t)(pw - 32);
    }
  }
  __device__ __forceinline__ int div(int n) const { return d == 1 ? n : (int)(__umulhi((uint32_t)n, mul) >> shr); }
  __device__ __forceinline__ void divmod(int n, int* q, int* r) const { *q = div(n); *r = n - *q * (int)d; }
};

struct TcParams {
  int n_seg; TcSeg seg[3];
  int total_k;
  int B, H, W;                 // spatial size of the grid the M tiles walk (output size; source size for the folded upsample)
  int bw, bh, bn;              // tile box, bw*bh*bn == 128 * mh  (128 per CTA in the pair kernel)
  int mh;                      // M-halves per CTA tile (2: two UMMAs share one B tile)
  int tiles_w, tiles_h, tiles_b, tiles_n, n_tiles;
  FastDiv d_tiles_n, d_phase, d_tiles_w, d_tiles_h, d_bw, d_bh;
  int n_phase;                 // 4: nearest-x2 upsample folded into the conv as four 2x2 sub-pixel convs (H, W = source size)
  int block_n, Cout;
  int a_slot_bytes, b_slot_bytes, n_a, n_b;   // operand rings
  int a_tile_bytes, a_halo_bytes;             // bytes of a plain / halo A load
  int kc;                      // channels per K-iteration: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B)
  int valid_rows;              // bw*bh*bn: rows of the tile that carry pixels (< 128*mh for maps such as 28x28)
  int tma_store;               // pair kernel: epilogue stages bf16 rows in swizzled shared memory, TMA writes them out
  int st_dh, st_dn;            // ... box origin of a CTA's second 128-row half (kMH = 2): rows / samples to skip
  int b_stat;                  // pair kernel, short-K 1x1 convs: the weight tile of this pair's N block stays in shared memory
  int nswap;                   // halo boxes of a tile that holds TWO whole samples (8x8 maps): image rows of the two samples interleave
                               // (tensor-map dims ordered C, W, N, H), tile row = (h * 2 + n) * bw + w, so the y taps stay row-shifted views
  const float* bias;
  const float* emb; int emb_stride; const int* emb_row;
  const bf16* res0; const bf16* res1; int R0, R1;
  bf16* out;
  float* out_nchw; int cout_real;   // network head: fp32 NCHW output of the first cout_real channels
  float* out_f32;                   // fp32 NHWC output (head taps: 32 partial products per pixel)
  // kGN: GroupNorm32 (+ SiLU) of the OUTPUT applied in the epilogue - the ResBlock's `out_layers.0` folded into its first
  // conv (unet.py:307-308), so the un-normalised h is never written and the separate GroupNorm pass disappears.
  const float* gn_gamma; const float* gn_beta; float gn_eps; int gn_cpg;
  int gn_R, gn_R_log2, gn_cpg_log2; // rows of a CTA tile that belong to one sample: min(H * W, 128 * mh); powers of two
  int gn_seg;                       // rows of one warp that share a sample: 32, or 16 (4x4 maps: a half-warp per sample)
  int gn_ctas;                      // CTA tiles per sample (> 1: partial sums are exchanged through L2, see the kernel)
  int gn_hw;                        // pixels per sample
  int gn_silu;
  bf16* out2;                       // dual output: `out` receives the un-normalised tile as well (a block output that is both
                                    // a residual / skip operand and the input of the next GroupNorm), `out2` the normalised one
  uint2* gn_exch;                   // [B][gn_ctas][32][2]: {sum | sum of squares, epoch} of every group over one CTA tile
  const unsigned* gn_epoch;         // device counter, bumped once per NFE (setup_rows_kernel): marks the words of this NFE
};

// ------------------------------------------------------------------------------------------------
// epilogue helpers (shared by the single-CTA and the CTA-pair kernel)
// ------------------------------------------------------------------------------------------------
// One accumulator row (= output pixel) as the epilogue sees it.
struct EpiRow { bool valid; int n, h, w; long long pix; const float* embp; };   // embp: this row's embedding vector (or null)

struct TileCoord { int nt, ph, tw, th, tb; };
// tile index -> (N tile, sub-pixel phase, spatial tile).  N tile and phase vary fastest, so the CTAs that share an
// A tile run at the same time and the tile is fetched from DRAM once.
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int tile) {
  TileCoord c;
  p.d_tiles_n.divmod(tile, &tile, &c.nt);
  p.d_phase.divmod(tile, &tile, &c.ph);
  p.d_tiles_w.divmod(tile, &tile, &c.tw);
  p.d_tiles_h.divmod(tile, &c.tb, &c.th);
  return c;
}
// tap offset of K-iteration `tap` of a segment: plain ks x ks conv, or (ks == 2) phase `ph` of the sub-pixel
// decomposition of "nearest x2 upsample, then 3x3 conv": output (2y+py, 2x+px) reads source rows y-1+py, y+py.
__device__ __forceinline__ void tap_offset(int ks, int tap, int ph, int* dy, int* dx) {
  if (ks == 2) { *dy = (tap >> 1) - 1 + (ph >> 1); *dx = (tap & 1) - 1 + (ph & 1); }
  else { const int pad = ks >> 1; *dy = tap / ks - pad; *dx = tap % ks - pad; }
}
// CTA-pair kernel: a pair tile is two consecutive 128-row tiles; `rank` picks this CTA's half
__device__ __forceinline__ TileCoord decode_pair_tile(const TcParams& p, int pt, int rank) {
  TileCoord c;
  p.d_tiles_n.divmod(pt, &pt, &c.nt);
  p.d_phase.divmod(pt, &pt, &c.ph);
  int mt = pt * 2 + rank;
  p.d_tiles_w.divmod(mt, &mt, &c.tw);
  p.d_tiles_h.divmod(mt, &c.tb, &c.th);
  return c;
}
__device__ __forceinline__ EpiRow epi_decode_row(const TcParams& p, const TileCoord& c, int row) {
  EpiRow r;
  int wi, hi, ni, t;
  p.d_bw.divmod(row, &t, &wi);
  if (p.nswap) { ni = t & 1; hi = t >> 1; }
  else p.d_bh.divmod(t, &ni, &hi);
  r.n = c.tb * p.bn + ni; r.h = c.th * p.bh + hi; r.w = c.tw * p.bw + wi;
  r.valid = row < p.valid_rows && r.n < p.B && r.h < p.H && r.w < p.W;
  if (p.n_phase == 4) {   // rows index the source grid; this phase's outputs interleave into the 2H x 2W map
    r.h = 2 * r.h + (c.ph >> 1); r.w = 2 * r.w + (c.ph & 1);
    r.pix = ((long long)r.n * (2 * p.H) + r.h) * (2 * p.W) + r.w;
  } else {
    r.pix = ((long long)r.n * p.H + r.h) * p.W + r.w;
  }
  r.embp = nullptr;
  return r;
}
// The row's time / class embedding vector: resolved ONCE per tile, before the accumulator is awaited - the index load
// (row_of_sample[n]) and the vector loads were two dependent global round trips in front of every 32-column chunk.
__device__ __forceinline__ void epi_attach_emb(const TcParams& p, EpiRow& r) {
  if (p.emb && r.valid) r.embp = p.emb + (long long)p.emb_row[r.n] * p.emb_stride;
}
// rows[half] without dynamic indexing (which would put the two EpiRows in local memory): field-wise selects
__device__ __forceinline__ EpiRow epi_pick(const EpiRow& a, const EpiRow& b, bool second) {
  EpiRow r;
  r.valid = second ? b.valid : a.valid; r.n = second ? b.n : a.n; r.h = second ? b.h : a.h; r.w = second ? b.w : a.w;
  r.pix = second ? b.pix : a.pix; r.embp = second ? b.embp : a.embp;
  return r;
}
// issue the residual loads of one 32-channel chunk early (they are the only DRAM-latency operand of the epilogue):
// this lane's row, 64 bytes as two 256-bit loads
__device__ __forceinline__ void epi_load_res(const TcParams& p, const EpiRow& r, int lane, int cg, uint4 (&res)[4]) {
  if (p.res0 && r.valid) {
    const bf16* rp = (cg < p.R0) ? p.res0 + r.pix * p.R0 + cg : p.res1 + r.pix * p.R1 + (cg - p.R0);
    ldg256(rp, res[0], res[1]);
    ldg256(rp + 16, res[2], res[3]);
  }
}
// pull the residual rows of a LATER tile into L2 (no registers held): the epilogue's residual loads are its only
// DRAM-latency operand and, issued one chunk ahead, they left ~1 us exposed per chunk on the short-K layers
__device__ __forceinline__ void epi_prefetch_res(const TcParams& p, const EpiRow& r, int c_begin, int n_cols) {
  if (p.res0 && r.valid) {
    for (int c = 0; c < n_cols; c += 64) {
      const int cg = c_begin + c;
      const bf16* rp = (cg < p.R0) ? p.res0 + r.pix * p.R0 + cg : p.res1 + r.pix * p.R1 + (cg - p.R0);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
    }
  }
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
// Explicit shared-space accesses for the epilogue's tables: through a generic pointer the compiler emits LD.E / ST.E,
// which resolve the address space at run time and sit on the long scoreboard (8 % of a K = 18 layer's stall samples were
// FFMAs waiting for such loads of the scale / shift table).
__device__ __forceinline__ float2 lds_f2(uint32_t saddr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float lds_f1(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_f2(uint32_t saddr, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(saddr), "f"(x), "f"(y) : "memory");
}
// accumulator chunk (32 fp32 from TMEM) + bias + embedding vector, in fp32
__device__ __forceinline__ void epi_bias_emb(const TcParams& p, const EpiRow& r, int cg, const uint32_t (&v)[32], uint32_t s_bias_addr,
                                             float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b4 = lds_f4(s_bias_addr + (uint32_t)(cg + j) * 4u);
    f[j] = __uint_as_float(v[j]) + b4.x; f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
    f[j + 2] = __uint_as_float(v[j + 2]) + b4.z; f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
  }
  if (r.embp) {
    const float* embp = r.embp;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 e4 = __ldg((const float4*)(embp + cg + j));
      f[j] += e4.x; f[j + 1] += e4.y; f[j + 2] += e4.z; f[j + 3] += e4.w;
    }
  }
}
// 32 fp32 -> bf16, this lane's 64 B of its output row: two full 32 B sectors per store instruction
__device__ __forceinline__ void epi_add_res(float (&f)[32], const uint4 (&res)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162* rb = (const __nv_bfloat162*)&res[j];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float2 t2 = __bfloat1622float2(rb[q]); f[8 * j + 2 * q] += t2.x; f[8 * j + 2 * q + 1] += t2.y; }
  }
}
__device__ __forceinline__ void epi_store_bf16(const TcParams& p, const EpiRow& r, int cg, const float (&f)[32], bf16* dst) {
  uint4 o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162* ob = (__nv_bfloat162*)&o[j];
#pragma unroll
    for (int q = 0; q < 4; ++q) ob[q] = __floats2bfloat162_rn(f[8 * j + 2 * q], f[8 * j + 2 * q + 1]);
  }
  bf16* op = dst + r.pix * p.Cout + cg;
  stg256(op, o[0], o[1]);
  stg256(op + 16, o[2], o[3]);
}
// accumulator chunk (32 fp32 from TMEM) + bias + embedding vector + residual -> bf16 NHWC (or fp32 NCHW for the head).
__device__ __forceinline__ void epi_finish(const TcParams& p, const EpiRow& r, int lane, int cg, const uint32_t (&v)[32], uint4 (&res)[4],
                                           uint32_t s_bias_addr) {
  float f[32];
  epi_bias_emb(p, r, cg, v, s_bias_addr, f);
  if (p.out_nchw) {
    // head conv: channel-planar fp32; lanes are consecutive pixels -> coalesced per channel
    if (!r.valid) return;
    const long long hw = (long long)p.H * p.W;
    float* op = p.out_nchw + (long long)r.n * p.cout_real * hw + (long long)r.h * p.W + r.w;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (cg + j < p.cout_real) op[(long long)(cg + j) * hw] = f[j];
    return;
  }
  if (!r.valid) return;
  if (p.out_f32) {                  // 32 fp32 = 128 B of this row: four 256-bit stores
    float* op = p.out_f32 + r.pix * p.Cout + cg;
#pragma unroll
    for (int j = 0; j < 32; j += 8)
      stg256(op + j, make_uint4(__float_as_uint(f[j]), __float_as_uint(f[j + 1]), __float_as_uint(f[j + 2]), __float_as_uint(f[j + 3])),
             make_uint4(__float_as_uint(f[j + 4]), __float_as_uint(f[j + 5]), __float_as_uint(f[j + 6]), __float_as_uint(f[j + 7])));
    return;
  }
  if (p.res0) epi_add_res(f, res);
  epi_store_bf16(p, r, cg, f, p.out);
}
// ---- fused GroupNorm epilogue helpers ----
// Statistics are kept per GRANULE of 4 consecutive channels (a group is a whole number of granules: cpg % 4 == 0), so each
// lane folds its row's 32 columns to 8 (sum, sum of squares) pairs in registers before anything crosses lanes.
__device__ __forceinline__ void epi_granules(const float (&f)[32], float (&gs)[8], float (&gq)[8], bool accumulate) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float a = f[4 * g], b = f[4 * g + 1], c = f[4 * g + 2], d = f[4 * g + 3];
    const float s = (a + b) + (c + d);
    const float q = fmaf(a, a, fmaf(b, b, fmaf(c, c, d * d)));
    gs[g] = accumulate ? gs[g] + s : s;
    gq[g] = accumulate ? gq[g] + q : q;
  }
}
// Sum the 8 granule pairs over the rows of a warp that belong to one sample: reduce-scatter butterfly (4 + 2 + 1
// exchanges per quantity), then plain butterflies over the remaining lanes of the sample.
//   kMode 0: the 32 rows are one sample; lane l ends with granule l >> 2 in gs[0], gq[0] (4 copies).
//   kMode 1: each half-warp is a sample (16 pixels per sample, 4x4 maps): lane l ends with granule (l >> 1) & 7 of its
//            half-warp's rows (2 copies).
//   kMode 2: the rows of two samples interleave in runs of 8 (nswap tiles of 8x8 maps: lane bit 3 is the sample): lane l
//            ends with granule 4 * bit4(l) + 2 * bit2(l) + bit1(l) of sample bit3(l) (2 copies).
template <int kMode>
__device__ __forceinline__ void epi_granule_reduce(float (&gs)[8], float (&gq)[8], int lane) {
  constexpr int HS[3][3] = {{16, 8, 4}, {8, 4, 2}, {16, 4, 2}};
#pragma unroll
  for (int st = 0, n = 4; st < 3; ++st, n >>= 1) {
    const int H = HS[kMode][st];
    const bool up = (lane & H) != 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < n) {
        const float ss = up ? gs[j] : gs[j + n], ks = up ? gs[j + n] : gs[j];
        const float sq = up ? gq[j] : gq[j + n], kq = up ? gq[j + n] : gq[j];
        gs[j] = ks + __shfl_xor_sync(0xffffffffu, ss, H);
        gq[j] = kq + __shfl_xor_sync(0xffffffffu, sq, H);
      }
    }
  }
#pragma unroll
  for (int H = kMode == 0 ? 2 : 1; H >= 1; H >>= 1) {
    gs[0] += __shfl_xor_sync(0xffffffffu, gs[0], H);
    gq[0] += __shfl_xor_sync(0xffffffffu, gq[0], H);
  }
}
__device__ __forceinline__ void st_relaxed_gpu_v2(uint2* p, unsigned x, unsigned y) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint2 ld_relaxed_gpu_v2(const uint2* p) {
  uint2 v; asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v;
}
// TMA-store variant: the same arithmetic (bias + embedding + residual, one rounding to bf16), but the 64 bytes of this
// row go to the staging box in shared memory - 128-byte rows, 16-byte chunks XOR-swizzled with the row (SWIZZLE_128B) -
// `chunk0` = first of the four chunks (0 or 4).  Rows that are not pixels are written too: TMA clips them.
__device__ __forceinline__ void epi_finish_smem(const TcParams& p, const EpiRow& r, int cg, const uint32_t (&v)[32], uint4 (&res)[4],
                                                uint32_t s_bias_addr, uint32_t stage_row_addr, int row, int chunk0) {
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b4 = lds_f4(s_bias_addr + (uint32_t)(cg + j) * 4u);
    f[j] = __uint_as_float(v[j]) + b4.x; f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
    f[j + 2] = __uint_as_float(v[j + 2]) + b4.z; f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
  }
  if (r.embp) {
    const float* embp = r.embp;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 e4 = __ldg((const float4*)(embp + cg + j));
      f[j] += e4.x; f[j + 1] += e4.y; f[j + 2] += e4.z; f[j + 3] += e4.w;
    }
  }
  if (p.res0 && r.valid) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162* rb = (const __nv_bfloat162*)&res[j];
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float2 t2 = __bfloat1622float2(rb[q]); f[8 * j + 2 * q] += t2.x; f[8 * j + 2 * q + 1] += t2.y; }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 o;
    __nv_bfloat162* ob = (__nv_bfloat162*)&o;
#pragma unroll
    for (int q = 0; q < 4; ++q) ob[q] = __floats2bfloat162_rn(f[8 * j + 2 * q], f[8 * j + 2 * q + 1]);
    sts_u4(stage_row_addr + (uint32_t)(((chunk0 + j) ^ (row & 7)) << 4), o);
  }
}

// ------------------------------------------------------------------------------------------------
// K-loop walk shared by the producer and the MMA issuer
// ------------------------------------------------------------------------------------------------
struct Ring { int slot; uint32_t phase; int n; __device__ __forceinline__ void next() { if (++slot == n) { slot = 0; phase ^= 1; } } };

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
               const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ring_a = smem;
  uint8_t* ring_b = smem + p.n_a * p.a_slot_bytes;
  float* s_bias = (float*)(smem + TC_RING_BYTES);
  uint64_t* bars = (uint64_t*)(smem + TC_RING_BYTES + TC_MAX_COUT * 4);
  uint64_t* fullA = bars;                              // [TC_MAX_A]
  uint64_t* emptyA = bars + TC_MAX_A;                  // [TC_MAX_A]
  uint64_t* fullB = bars + 2 * TC_MAX_A;               // [TC_MAX_B]
  uint64_t* emptyB = bars + 2 * TC_MAX_A + TC_MAX_B;   // [TC_MAX_B]
  uint64_t* tfull_bar = bars + 2 * TC_MAX_A + 2 * TC_MAX_B;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                          // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_launch_dependents();      // the next kernel's CTAs may take over SMs as ours retire and run their prologue
  if (warp == 0) {
    // the 52 barriers are initialised by the 32 lanes in parallel (one thread doing all of them costs ~1 us per launch)
    if (lane == 0) {
      prefetch_tmap(&mapA0); prefetch_tmap(&mapB);
      if (p.n_seg > 1) prefetch_tmap(&mapA1);
      if (p.n_seg > 2) prefetch_tmap(&mapA2);
    }
    if (lane < TC_MAX_A) { mbar_init(&fullA[lane], 1); mbar_init(&emptyA[lane], 1); }
    if (lane < TC_MAX_B) { mbar_init(&fullB[lane], 1); mbar_init(&emptyB[lane], 1); }
    if (lane < 2) { mbar_init(&tfull_bar[lane], 1); mbar_init(&tempty_bar[lane], TC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) s_bias[i] = p.bias[i];   // weights: not produced by a predecessor
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                   // activations of the predecessor are complete and visible from here on
  const uint32_t tmem_base = *tmem_slot;
  const int acc_cols = p.block_n * p.mh;     // TMEM columns per accumulator stage

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      Ring ra{0, 0, p.n_a}, rb{0, 0, p.n_b};
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        const int w0 = tc.tw * p.bw, h0 = tc.th * p.bh, n0 = tc.tb * p.bn;
        int brow = tc.ph * p.total_k * p.Cout + tc.nt * p.block_n;      // row of the packed weights of the next K-iteration
        for (int s = 0; s < p.n_seg; ++s) {
          const TcSeg sg = p.seg[s];
          const CUtensorMap* map = sg.map == 0 ? &mapA0 : (sg.map == 1 ? &mapA1 : &mapA2);
          if (sg.halo) {
            const int yorg = h0 - 1 + (sg.ks == 2 ? (tc.ph >> 1) : 0);
            for (int ch = 0; ch < sg.n_chunks; ++ch)
              for (int xi = 0; xi < sg.ks; ++xi) {
                const int dx = xi - 1 + (sg.ks == 2 ? (tc.ph & 1) : 0);
                mbar_wait(&emptyA[ra.slot], ra.phase ^ 1);
                mbar_expect_tx(&fullA[ra.slot], p.a_halo_bytes);
                tma_load_4d(ring_a + ra.slot * p.a_slot_bytes, map, &fullA[ra.slot], ch * p.kc, w0 + dx, yorg, n0);
                ra.next();
                for (int yi = 0; yi < sg.ks; ++yi, brow += p.Cout) {
                  mbar_wait(&emptyB[rb.slot], rb.phase ^ 1);
                  mbar_expect_tx(&fullB[rb.slot], p.b_slot_bytes);
                  tma_load_2d(ring_b + rb.slot * p.b_slot_bytes, &mapB, &fullB[rb.slot], 0, brow);
                  rb.next();
                }
              }
          } else {
            const int taps = sg.ks * sg.ks;
            for (int tap = 0; tap < taps; ++tap) {
              int dy, dx; tap_offset(sg.ks, tap, tc.ph, &dy, &dx);
              for (int ch = 0; ch < sg.n_chunks; ++ch, brow += p.Cout) {
                mbar_wait(&emptyA[ra.slot], ra.phase ^ 1);
                mbar_expect_tx(&fullA[ra.slot], p.a_tile_bytes);
                tma_load_4d(ring_a + ra.slot * p.a_slot_bytes, map, &fullA[ra.slot], ch * p.kc, w0 * sg.stride + dx, h0 * sg.stride + dy, n0);
                ra.next();
                mbar_wait(&emptyB[rb.slot], rb.phase ^ 1);
                mbar_expect_tx(&fullB[rb.slot], p.b_slot_bytes);
                tma_load_2d(ring_b + rb.slot * p.b_slot_bytes, &mapB, &fullB[rb.slot], 0, brow);
                rb.next();
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(TC_BLOCK_M, p.block_n);
    const uint32_t row_step = (uint32_t)(p.bw * p.kc * 2) >> 4;   // one image row of a halo slot, in descriptor units
    const uint32_t half_step = (uint32_t)(TC_BLOCK_M * p.kc * 2) >> 4;   // second M-half of an A slot
    const int n_kk = p.kc >> 4;
    Ring ra{0, 0, p.n_a}, rb{0, 0, p.n_b};
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_cols);
      int kdone = 0;
      for (int s = 0; s < p.n_seg; ++s) {
        const TcSeg sg = p.seg[s];
        const int G = sg.halo ? sg.ks : 1;
        const int n_groups = sg.halo ? sg.n_chunks * sg.ks : sg.ks * sg.ks * sg.n_chunks;
        for (int g = 0; g < n_groups; ++g) {
          mbar_wait(&fullA[ra.slot], ra.phase);
          const uint32_t sa = smem_u32(ring_a + ra.slot * p.a_slot_bytes);
          for (int j = 0; j < G; ++j, ++kdone) {
            mbar_wait(&fullB[rb.slot], rb.phase);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t adesc = make_desc_k(sa, p.kc) + (uint64_t)(sg.halo ? (uint32_t)j * row_step : 0u);
              const uint64_t bdesc = make_desc_k(smem_u32(ring_b + rb.slot * p.b_slot_bytes), p.kc);
#pragma unroll
              for (int kk = 0; kk < TC_BLOCK_K / 16; ++kk) {
                if (kk < n_kk) {
                  const uint32_t accum = (kdone > 0 || kk > 0) ? 1u : 0u;
                  umma_bf16(d_tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, accum);
                  if (p.mh == 2)   // second M-half re-uses the same B tile: halves the B traffic per FLOP
                    umma_bf16(d_tmem + (uint32_t)p.block_n, adesc + (uint64_t)half_step + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, accum);
                }
              }
              umma_commit(&emptyB[rb.slot]);                      // B slot reusable once these MMAs retire
              if (j == G - 1) umma_commit(&emptyA[ra.slot]);      // last view of this A slot
              if (kdone == p.total_k - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete
            }
            __syncwarp();
            rb.next();
          }
          ra.next();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // warp%4 selects the TMEM lane quarter the hardware lets this warp read; the two warps sharing a
    // quarter split the (M-half, 32-column chunk) work items between them.  The residual operand of the
    // first item is fetched BEFORE waiting for the accumulator, and that of item i+1 while item i is
    // being finished, so no DRAM latency sits between "accumulator ready" and "accumulator released".
    const int quad = warp & 3;
    const int sub = (warp - 2) >> 2;           // 0 or 1
    const int chunks_per_half = p.block_n >> 5;
    const int n_items = chunks_per_half * p.mh;
    const uint32_t s_bias_addr = smem_u32(s_bias);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile);
      const int nt = tc.nt;
      EpiRow row0 = epi_decode_row(p, tc, quad * 32 + lane);
      epi_attach_emb(p, row0);
      EpiRow row1 = row0;
      if (p.mh == 2) { row1 = epi_decode_row(p, tc, 128 + quad * 32 + lane); epi_attach_emb(p, row1); }
      // (no L2 prefetch of the next tile's residual rows here: on the narrow MNIST layers this kernel serves, the extra
      //  tile decode in a K = 32..96 tile's epilogue cost more than the latency it hid: proj_out 0.171 -> 0.187 ms)
      uint4 res_cur[4], res_nxt[4];
      if (sub < n_items) {
        const int half = sub >= chunks_per_half ? 1 : 0;          // mh <= 2
        epi_load_res(p, epi_pick(row0, row1, half != 0), lane, nt * p.block_n + ((sub - half * chunks_per_half) << 5), res_nxt);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * acc_cols);
      for (int item = sub; item < n_items; item += TC_EPI_WARPS / 4) {
        const int half = item >= chunks_per_half ? 1 : 0;
        const int c0 = (item - half * chunks_per_half) << 5;
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)(half * p.block_n + c0), v);
#pragma unroll
        for (int j = 0; j < 4; ++j) res_cur[j] = res_nxt[j];
        const int nxt = item + TC_EPI_WARPS / 4;
        if (nxt < n_items) {
          const int nh = nxt >= chunks_per_half ? 1 : 0;
          epi_load_res(p, epi_pick(row0, row1, nh != 0), lane, nt * p.block_n + ((nxt - nh * chunks_per_half) << 5), res_nxt);
        }
        tmem_ld_wait();
        epi_finish(p, epi_pick(row0, row1, half != 0), lane, nt * p.block_n + c0, v, res_cur, s_bias_addr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two SMs of a cluster cooperate on a 256-row x N tile.  Each CTA loads
// its own 128-row A half and HALF of the weight tile (N/2 rows) - so the bytes pulled from L2 and
// written to / read from shared memory per FLOP drop by a third versus the single-CTA 128 x N tile.
// The leader CTA's MMA warp issues one 256 x N x 16 UMMA per K step; tcgen05.commit multicasts
// slot-free / accumulator-ready signals to both CTAs; both CTAs' epilogues drain their own 128 TMEM
// lanes and report back to the leader's accumulator-empty barrier.
// ------------------------------------------------------------------------------------------------
// kMH: 128-row M-halves per CTA, a compile-time constant so that the kMH = 1 instances (N >= 192 layers, most of
// them short-K and epilogue-bound) carry none of the two-half bookkeeping
// kGN: the epilogue also applies GroupNorm32 (+ SiLU) to the tile (see the epilogue branch below).
template <int kMH, bool kGN>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ CUtensorMap mapOut, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ring_a = smem;
  uint8_t* ring_b = smem + p.n_a * p.a_slot_bytes;
  float* s_bias = (float*)(smem + TC_RING_BYTES);
  uint64_t* bars = (uint64_t*)(smem + TC_RING_BYTES + TC_MAX_COUT * 4);
  uint64_t* fullA = bars;                              // leader's are used
  uint64_t* emptyA = bars + TC_MAX_A;                  // per CTA, signalled by the leader's commit multicast
  uint64_t* fullB = bars + 2 * TC_MAX_A;               // leader's are used
  uint64_t* emptyB = bars + 2 * TC_MAX_A + TC_MAX_B;   // per CTA
  uint64_t* tfull_bar = bars + 2 * TC_MAX_A + 2 * TC_MAX_B;      // [2] per CTA, commit multicast
  uint64_t* tempty_bar = tfull_bar + 2;                          // [2] leader's are used: 2 * TC_EPI_WARPS arrivals
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles128 = p.tiles_w * p.tiles_h * p.tiles_b;
  const int pair_tiles = ((tiles128 + 1) >> 1) * p.tiles_n * p.n_phase;

  pdl_launch_dependents();
  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&mapA0); prefetch_tmap(&mapB);
      if (p.n_seg > 1) prefetch_tmap(&mapA1);
      if (p.n_seg > 2) prefetch_tmap(&mapA2);
    }
    if (lane < TC_MAX_A) { mbar_init(&fullA[lane], 1); mbar_init(&emptyA[lane], 1); }
    if (lane < TC_MAX_B) { mbar_init(&fullB[lane], 1); mbar_init(&emptyB[lane], 1); }
    if (lane < 2) { mbar_init(&tfull_bar[lane], 1); mbar_init(&tempty_bar[lane], 2 * TC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
  // (bias / gamma / beta are staged by the epilogue warps themselves, below: their global-load latency used to sit in
  //  front of this block-wide barrier and so in front of the producer's first TMA load)
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // peer barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      Ring ra{0, 0, p.n_a}, rb{0, 0, p.n_b};
      for (int pt = pair; pt < pair_tiles; pt += n_pairs) {
        const bool load_b = !p.b_stat || pt == pair;               // stationary weights: loaded with the first tile only
        const TileCoord tc = decode_pair_tile(p, pt, (int)rank);   // tb may equal tiles_b for the odd tail: all-OOB loads
        const int w0 = tc.tw * p.bw, h0 = tc.th * p.bh, n0 = tc.tb * p.bn;
        int brow = tc.ph * p.total_k * p.Cout + tc.nt * p.block_n + (int)rank * (p.block_n >> 1);
        for (int s = 0; s < p.n_seg; ++s) {
          const TcSeg sg = p.seg[s];
          const CUtensorMap* map = sg.map == 0 ? &mapA0 : (sg.map == 1 ? &mapA1 : &mapA2);
          if (sg.halo) {
            const int yorg = h0 - 1 + (sg.ks == 2 ? (tc.ph >> 1) : 0);
            for (int ch = 0; ch < sg.n_chunks; ++ch)
              for (int xi = 0; xi < sg.ks; ++xi) {
                const int dx = xi - 1 + (sg.ks == 2 ? (tc.ph & 1) : 0);
                mbar_wait(&emptyA[ra.slot], ra.phase ^ 1);
                if (leader) mbar_expect_tx(&fullA[ra.slot], 2 * p.a_halo_bytes);      // bytes of BOTH CTAs
                tma_load_4d_2sm(ring_a + ra.slot * p.a_slot_bytes, map, mapa_u32(smem_u32(&fullA[ra.slot]), 0), ch * p.kc, w0 + dx,
                                p.nswap ? n0 : yorg, p.nswap ? yorg : n0);
                ra.next();
                for (int yi = 0; yi < sg.ks; ++yi, brow += p.Cout) {
                  mbar_wait(&emptyB[rb.slot], rb.phase ^ 1);
                  if (leader) mbar_expect_tx(&fullB[rb.slot], 2 * p.b_slot_bytes);
                  tma_load_2d_2sm(ring_b + rb.slot * p.b_slot_bytes, &mapB, mapa_u32(smem_u32(&fullB[rb.slot]), 0), 0, brow);
                  rb.next();
                }
              }
          } else {
            const int taps = sg.ks * sg.ks;
            for (int tap = 0; tap < taps; ++tap) {
              int dy, dx; tap_offset(sg.ks, tap, tc.ph, &dy, &dx);
              for (int ch = 0; ch < sg.n_chunks; ++ch, brow += p.Cout) {
                mbar_wait(&emptyA[ra.slot], ra.phase ^ 1);
                if (leader) mbar_expect_tx(&fullA[ra.slot], 2 * p.a_tile_bytes);
                tma_load_4d_2sm(ring_a + ra.slot * p.a_slot_bytes, map, mapa_u32(smem_u32(&fullA[ra.slot]), 0), ch * p.kc, w0 * sg.stride + dx,
                                p.nswap ? n0 : h0 * sg.stride + dy, p.nswap ? h0 * sg.stride + dy : n0);
                ra.next();
                if (load_b) {
                  mbar_wait(&emptyB[rb.slot], rb.phase ^ 1);
                  if (leader) mbar_expect_tx(&fullB[rb.slot], 2 * p.b_slot_bytes);
                  tma_load_2d_2sm(ring_b + rb.slot * p.b_slot_bytes, &mapB, mapa_u32(smem_u32(&fullB[rb.slot]), 0), 0, brow);
                  rb.next();
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint32_t idesc = make_idesc(256, p.block_n);
      const uint32_t row_step = (uint32_t)(p.bw * (p.nswap ? 2 : 1) * p.kc * 2) >> 4;   // one image row (of both samples: nswap)
      const uint32_t half_step = (uint32_t)(TC_BLOCK_M * p.kc * 2) >> 4;   // second M-half of each CTA's A slot
      const int n_kk = p.kc >> 4;
      const int acc_cols = p.block_n * kMH;
      Ring ra{0, 0, p.n_a}, rb{0, 0, p.n_b};
      int acc = 0; uint32_t acc_phase = 0;
      for (int pt = pair; pt < pair_tiles; pt += n_pairs) {
        const bool wait_b = !p.b_stat || pt == pair;      // stationary weights (n_b == total_k: slot == K step) arrive once
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_cols);
        int kdone = 0;
        for (int s = 0; s < p.n_seg; ++s) {
          const TcSeg sg = p.seg[s];
          const int G = sg.halo ? sg.ks : 1;
          const int n_groups = sg.halo ? sg.n_chunks * sg.ks : sg.ks * sg.ks * sg.n_chunks;
          for (int g = 0; g < n_groups; ++g) {
            mbar_wait(&fullA[ra.slot], ra.phase);
            const uint32_t sa = smem_u32(ring_a + ra.slot * p.a_slot_bytes);
            for (int j = 0; j < G; ++j, ++kdone) {
              if (wait_b) mbar_wait(&fullB[rb.slot], rb.phase);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t adesc = make_desc_k(sa, p.kc) + (uint64_t)(sg.halo ? (uint32_t)j * row_step : 0u);
                const uint64_t bdesc = make_desc_k(smem_u32(ring_b + rb.slot * p.b_slot_bytes), p.kc);
#pragma unroll
                for (int kk = 0; kk < TC_BLOCK_K / 16; ++kk)
                  if (kk < n_kk) {
                    const uint32_t accum = (kdone > 0 || kk > 0) ? 1u : 0u;
                    umma_bf16_2sm(d_tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, accum);
                    if (kMH == 2)   // each CTA's second 128 rows against the same (shared) weight tile
                      umma_bf16_2sm(d_tmem + (uint32_t)p.block_n, adesc + (uint64_t)half_step + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, accum);
                  }
                if (!p.b_stat) umma_commit_2sm(&emptyB[rb.slot], 3);          // both CTAs may refill this B slot
                if (j == G - 1) umma_commit_2sm(&emptyA[ra.slot], 3);         // last view of this A slot
                if (kdone == p.total_k - 1) umma_commit_2sm(&tfull_bar[acc], 3);   // both CTAs' epilogues may drain
              }
              __syncwarp();
              rb.next();
            }
            ra.next();
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9, both CTAs) =====================
    const int quad = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int chunks_per_half = p.block_n >> 5;
    const int n_items = chunks_per_half * kMH;
    const int acc_cols = p.block_n * kMH;
    const uint32_t s_bias_addr = smem_u32(s_bias);
    int acc = 0; uint32_t acc_phase = 0;
    {
      // stage the per-channel constants (weights of the layer: not produced by a predecessor) while the first tile's
      // operands are in flight; only the eight epilogue warps read them
      const int et0 = (warp - 2) * 32 + lane;
      for (int i = et0; i < p.Cout; i += 32 * TC_EPI_WARPS) s_bias[i] = p.bias[i];
      if (kGN) {
        float* s_gb = (float*)(smem + TC_GN_OFF + 2048);
        for (int i = et0; i < p.Cout; i += 32 * TC_EPI_WARPS) { s_gb[i] = p.gn_gamma[i]; s_gb[TC_GN_MAX_COUT + i] = p.gn_beta[i]; }
      }
      named_bar_sync(6, 32 * TC_EPI_WARPS);
    }
    if (kGN) {
      // ---- fused GroupNorm epilogue (ResBlock conv1 + out_layers.0/1): two passes over the accumulator stage.
      // Pass 1: every warp folds its rows to per-granule (sum, sum of squares), the eight warps combine them in a fixed
      //   order through shared memory into per-(sample, group) partial sums of this CTA tile.  When a sample spans
      //   several CTA tiles (32x32 maps: four 256-row tiles on two CTA pairs) the partials are exchanged through L2:
      //   each CTA stores its 32 pairs, publishes an epoch flag (st.release) and spins on the flags of the sample's
      //   other tiles.  Those CTAs are co-resident (persistent grid, one CTA per SM, consecutive pair tiles go to
      //   consecutive pairs) and never wait on a LATER tile, so the wait cannot deadlock.  Totals are summed in tile
      //   order: bit-reproducible, no atomics.
      // Pass 2: the accumulators are read again, normalised, SiLU-activated and stored as bf16.  The second TMEM stage
      //   keeps the MMAs of the next tile running underneath.
      // shared-space byte addresses of the epilogue's tables (explicit ld.shared / st.shared, see lds_f2)
      const uint32_t s_mr = smem_u32(smem + TC_GN_OFF);             // float2 [unit][group]: mean, rstd
      const uint32_t s_gamma = s_mr + 2048;                         // float [256]
      const uint32_t s_beta = s_gamma + TC_GN_MAX_COUT * 4;         // float [256]
      const uint32_t s_scr = s_mr + 4096;                           // float2 [chunk][block][granule]
      const uint32_t s_tab = s_mr + 8192;                           // float2 [unit <= 2][channel]: (scale, shift) of pass 2
      const int n_mine = chunks_per_half > sub ? ((chunks_per_half - sub + 1) >> 1) * kMH : 0;
      const bool sum_halves = kMH == 2 && p.gn_R == 256;            // both 128-row halves belong to the same sample
      const bool il = kMH == 1 && p.nswap;                          // two samples, rows interleaved in runs of 8 (gn_seg == 8)
      const int rows_blk = sum_halves ? 64 : (il ? 16 : p.gn_seg);  // tile rows one scratch block stands for
      const int n_blk = sum_halves ? 4 : (128 * kMH) / rows_blk;    // scratch blocks per chunk (<= 8)
      const int bpu = sum_halves ? 4 : p.gn_R / rows_blk;           // blocks per sample ("unit") of the tile
      const int n_units = n_blk / bpu;
      const int gpg = p.gn_cpg >> 2;                                // granules per group
      const int et = (warp - 2) * 32 + lane;                        // epilogue thread 0..255
      const unsigned epoch = p.gn_ctas > 1 ? *p.gn_epoch : 0u;
      for (int pt = pair; pt < pair_tiles; pt += n_pairs) {
        const TileCoord tc = decode_pair_tile(p, pt, (int)rank);
        const int cbase = tc.nt * p.block_n;                       // first channel of this N tile (whole groups per tile)
        const int groups_tile = p.block_n >> p.gn_cpg_log2;
        EpiRow row0 = epi_decode_row(p, tc, quad * 32 + lane);
        epi_attach_emb(p, row0);
        EpiRow row1 = row0;
        if (kMH == 2) { row1 = epi_decode_row(p, tc, 128 + quad * 32 + lane); epi_attach_emb(p, row1); }
        uint4 res_cur[4], res_nxt[4];
        if (p.res0) {
          if (sub == 0 && pt + n_pairs < pair_tiles) {      // next tile's residual rows -> L2
            const TileCoord tn = decode_pair_tile(p, pt + n_pairs, (int)rank);
            epi_prefetch_res(p, epi_decode_row(p, tn, quad * 32 + lane), tn.nt * p.block_n, p.block_n);
            if (kMH == 2) epi_prefetch_res(p, epi_decode_row(p, tn, 128 + quad * 32 + lane), tn.nt * p.block_n, p.block_n);
          }
          if (n_mine > 0) epi_load_res(p, row0, lane, cbase + (sub << 5), res_nxt);
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * acc_cols);
        // ---------------- pass 1: finished values (bias + embedding + residual) back into TMEM, statistics ----------------
        float run_s[8], run_q[8];
        for (int k = 0; k < n_mine; ++k) {
          const int half = kMH == 2 ? (k & 1) : 0;
          const int ci = sub + 2 * (kMH == 2 ? (k >> 1) : k);
          const EpiRow rr = epi_pick(row0, row1, half != 0);
          const uint32_t ta = t_addr + (uint32_t)(half * p.block_n + (ci << 5));
          uint32_t v[32];
          tmem_ld32(ta, v);
#pragma unroll
          for (int j = 0; j < 4; ++j) res_cur[j] = res_nxt[j];
          if (p.res0 && k + 1 < n_mine) {   // the next item's residual row travels while this item is being reduced
            const int nh = kMH == 2 ? ((k + 1) & 1) : 0;
            const int nc0 = (sub + 2 * (kMH == 2 ? ((k + 1) >> 1) : (k + 1))) << 5;
            epi_load_res(p, epi_pick(row0, row1, nh != 0), lane, cbase + nc0, res_nxt);
          }
          tmem_ld_wait();
          float f[32];
          epi_bias_emb(p, rr, cbase + (ci << 5), v, s_bias_addr, f);
          if (p.res0) epi_add_res(f, res_cur);
          if (!rr.valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(f[j]);
          tmem_st32(ta, v);                 // pass 2 reads the finished value back: no second bias / embedding / residual fetch
          epi_granules(f, run_s, run_q, sum_halves && half == 1);
          if (sum_halves && half == 0) continue;
          const uint32_t sc = s_scr + (uint32_t)ci * 512u;      // 64 float2 per chunk
          if (il) {                      // block = sample * 4 + warp quarter: a sample's four blocks are contiguous
            epi_granule_reduce<2>(run_s, run_q, lane);
            if ((lane & 1) == 0)
              sts_f2(sc + (uint32_t)((((lane >> 3) & 1) * 4 + quad) * 8 + ((lane >> 4) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)) * 8u, run_s[0], run_q[0]);
          } else if (p.gn_seg == 32) {
            epi_granule_reduce<0>(run_s, run_q, lane);
            const int blk = sum_halves ? quad : half * 4 + quad;
            if ((lane & 3) == 0) sts_f2(sc + (uint32_t)(blk * 8 + (lane >> 2)) * 8u, run_s[0], run_q[0]);
          } else {
            epi_granule_reduce<1>(run_s, run_q, lane);
            if ((lane & 1) == 0) sts_f2(sc + (uint32_t)((quad * 2 + (lane >> 4)) * 8 + ((lane >> 1) & 7)) * 8u, run_s[0], run_q[0]);
          }
        }
        tmem_st_wait();
        named_bar_sync(5, 32 * TC_EPI_WARPS);
        // ---------------- per-(sample, group) statistics of the tile ----------------
        if (et < n_units * 32 && (et & 31) < groups_tile) {
          const int u = et >> 5, g = et & 31;                   // a warp per unit, a lane per group of this N tile
          const int g0 = g * gpg;                                // first granule of the group; groups never straddle chunks
          const uint32_t sc = s_scr + (uint32_t)((g0 >> 3) * 64 + (g0 & 7)) * 8u;
          float ts = 0.f, tq = 0.f;
          for (int b = 0; b < bpu; ++b)
            for (int j = 0; j < gpg; ++j) { const float2 t2 = lds_f2(sc + (uint32_t)((u * bpu + b) * 8 + j) * 8u); ts += t2.x; tq += t2.y; }
          if (p.gn_ctas > 1) {
            // u == 0: the whole CTA tile lies inside sample n; `my` = its index among the sample's tiles.  Every partial
            // travels as two 8-byte words {value, epoch}: an aligned 8-byte store is single-copy atomic, so the word itself
            // says when it is valid - no fence, no release / acquire pair (and no L1 invalidation) on either side.
            const int n = tc.tb * p.bn;
            const int my = (tc.th * p.bh * p.W) / (128 * kMH);
            if (n < p.B) {
              uint2* ex = p.gn_exch + ((long long)n * p.gn_ctas) * 64;
              st_relaxed_gpu_v2(ex + (my * 32 + g) * 2, __float_as_uint(ts), epoch);       // (tiles that exchange span all channels: g is global)
              st_relaxed_gpu_v2(ex + (my * 32 + g) * 2 + 1, __float_as_uint(tq), epoch);
              ts = 0.f; tq = 0.f;
              for (int j0 = 0; j0 < p.gn_ctas; j0 += 4) {  // fixed order over the sample's tiles: bit-reproducible
                // four tiles' words per round trip to L2 (the other tiles of the sample have usually published already)
                uint2 a[4], b[4];
                bool ok;
                do {
                  ok = true;
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (j0 + j < p.gn_ctas) {
                      a[j] = ld_relaxed_gpu_v2(ex + ((j0 + j) * 32 + g) * 2);
                      b[j] = ld_relaxed_gpu_v2(ex + ((j0 + j) * 32 + g) * 2 + 1);
                    }
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (j0 + j < p.gn_ctas) ok = ok && a[j].y == epoch && b[j].y == epoch;
                } while (!ok);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (j0 + j < p.gn_ctas) { ts += __uint_as_float(a[j].x); tq += __uint_as_float(b[j].x); }
              }
            }
          }
          const float inv_n = 1.0f / (float)(p.gn_cpg * p.gn_hw);
          const float mean = ts * inv_n;
          const float var = fmaxf(tq * inv_n - mean * mean, 0.f);
          sts_f2(s_mr + (uint32_t)(u * 32 + g) * 8u, mean, rsqrtf(var + p.gn_eps));
        }
        named_bar_sync(5, 32 * TC_EPI_WARPS);
        const float hs = p.gn_silu ? 0.5f : 1.0f;                   // silu(y) = h * tanh(h) + h with h = y / 2
        const bool dual = p.out2 != nullptr;
        const bool table = n_units <= 2;
        if (table) {
          // per-(sample, channel) scale / shift: y = f * sc + sh,  sc = rstd * gamma * hs,  sh = (beta - mean * rstd * gamma) * hs
          for (int i = et; i < n_units * p.block_n; i += 32 * TC_EPI_WARPS) {
            const int u = i >= p.block_n ? 1 : 0, c = i - u * p.block_n;       // c: channel within this N tile
            const float2 m = lds_f2(s_mr + (uint32_t)(u * 32 + (c >> p.gn_cpg_log2)) * 8u);
            const float sc = m.y * lds_f1(s_gamma + (uint32_t)(cbase + c) * 4u) * hs;
            sts_f2(s_tab + (uint32_t)(u * TC_GN_MAX_COUT + c) * 8u, sc, fmaf(-m.x, sc, lds_f1(s_beta + (uint32_t)(cbase + c) * 4u) * hs));
          }
          named_bar_sync(5, 32 * TC_EPI_WARPS);
        }
        // ---------------- pass 2: normalise, activate, store ----------------
        for (int k = 0; k < n_mine; ++k) {
          const int half = kMH == 2 ? (k & 1) : 0;
          const int c0 = (sub + 2 * (kMH == 2 ? (k >> 1) : k)) << 5;
          const EpiRow rr = epi_pick(row0, row1, half != 0);
          uint32_t v[32];
          tmem_ld32(t_addr + (uint32_t)(half * p.block_n + c0), v);
          tmem_ld_wait();
          if (k == n_mine - 1) {       // last read of this accumulator stage
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
          }
          const int trow = half * 128 + quad * 32 + lane;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (dual && rr.valid) epi_store_bf16(p, rr, cbase + c0, f, p.out);
          if (table) {
            const uint32_t tb = s_tab + (uint32_t)((il ? ((trow >> 3) & 1) : (trow >> p.gn_R_log2)) * TC_GN_MAX_COUT + c0) * 8u;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float4 t4 = lds_f4(tb + (uint32_t)j * 8u);
              f[j] = fmaf(f[j], t4.x, t4.y); f[j + 1] = fmaf(f[j + 1], t4.z, t4.w);
            }
          } else {
            const uint32_t mr = s_mr + (uint32_t)((il ? ((trow >> 3) & 1) : (trow >> p.gn_R_log2)) * 32) * 8u;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float2 m = lds_f2(mr + (uint32_t)((c0 + j) >> p.gn_cpg_log2) * 8u);       // 4 | cpg: the four channels share a group
              const float4 ga = lds_f4(s_gamma + (uint32_t)(cbase + c0 + j) * 4u), be = lds_f4(s_beta + (uint32_t)(cbase + c0 + j) * 4u);
              const float r = m.y * hs;
              f[j] = fmaf((f[j] - m.x) * r, ga.x, be.x * hs); f[j + 1] = fmaf((f[j + 1] - m.x) * r, ga.y, be.y * hs);
              f[j + 2] = fmaf((f[j + 2] - m.x) * r, ga.z, be.z * hs); f[j + 3] = fmaf((f[j + 3] - m.x) * r, ga.w, be.w * hs);
            }
          }
          if (p.gn_silu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = silu_from_half(f[j]);
          }
          if (rr.valid) epi_store_bf16(p, rr, cbase + c0, f, dual ? p.out2 : p.out);
        }
        if (n_mine == 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    } else if (p.tma_store) {
      // ---- staged epilogue: each group of four warps (one per TMEM lane quadrant) owns two 16 KB staging boxes.  A unit
      // is one 128-row x 64-channel box of the output: two 32-column TMEM chunks per thread, written as bf16 into the
      // swizzled box, then ONE thread hands the box to TMA (full 128-byte lines, no per-lane sector stores).
      const int n_cb = p.block_n >> 6;
      const int n_units = n_cb * kMH;
      const uint32_t stage0 = smem_u32(smem + TC_STAGE_OFF) + (uint32_t)(sub * 2 * TC_STAGE_BUF);
      const int row = quad * 32 + lane;
      const bool issuer = quad == 0 && lane == 0;
      uint32_t n_store = 0;
      if (issuer) prefetch_tmap(&mapOut);
      for (int pt = pair; pt < pair_tiles; pt += n_pairs) {
        const TileCoord tc = decode_pair_tile(p, pt, (int)rank);
        const int nt = tc.nt;
        EpiRow row0 = epi_decode_row(p, tc, row);
        epi_attach_emb(p, row0);
        EpiRow row1 = row0;
        if (kMH == 2) { row1 = epi_decode_row(p, tc, 128 + row); epi_attach_emb(p, row1); }
        if (p.res0 && sub == 0 && pt + n_pairs < pair_tiles) {
          const TileCoord tn = decode_pair_tile(p, pt + n_pairs, (int)rank);
          epi_prefetch_res(p, epi_decode_row(p, tn, row), tn.nt * p.block_n, p.block_n);
          if (kMH == 2) epi_prefetch_res(p, epi_decode_row(p, tn, 128 + row), tn.nt * p.block_n, p.block_n);
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * acc_cols);
        for (int u = sub; u < n_units; u += TC_EPI_WARPS / 4) {
          const int half = (kMH == 2 && u >= n_cb) ? 1 : 0;
          const int cb = u - half * n_cb;
          const EpiRow rr = epi_pick(row0, row1, half != 0);
          const uint32_t buf = stage0 + (n_store & 1u) * (uint32_t)TC_STAGE_BUF;
          uint4 res[2][4];
          epi_load_res(p, rr, lane, nt * p.block_n + (cb << 6), res[0]);
          epi_load_res(p, rr, lane, nt * p.block_n + (cb << 6) + 32, res[1]);
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int c0 = (cb << 6) + (it << 5);
            uint32_t v[32];
            tmem_ld32(t_addr + (uint32_t)(half * p.block_n + c0), v);
            tmem_ld_wait();
            epi_finish_smem(p, rr, nt * p.block_n + c0, v, res[it], s_bias_addr, buf + (uint32_t)row * 128u, row, it * 4);
          }
          fence_proxy_async();                      // generic-proxy writes of this thread -> visible to the TMA engine
          if (issuer) bulk_wait_group_read0();      // the box stored one unit ago has been read: free after the barrier
          named_bar_sync(1 + sub, 128);
          if (issuer) {
            tma_store_4d(&mapOut, smem + TC_STAGE_OFF + (sub * 2 + (int)(n_store & 1u)) * TC_STAGE_BUF, nt * p.block_n + (cb << 6),
                         tc.tw * p.bw, tc.th * p.bh + half * p.st_dh, tc.tb * p.bn + half * p.st_dn);
            bulk_commit_group();
          }
          ++n_store;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (issuer) bulk_wait_group0();               // shared memory stays valid until every box has left
    } else
    for (int pt = pair; pt < pair_tiles; pt += n_pairs) {
      const TileCoord tc = decode_pair_tile(p, pt, (int)rank);
      const int nt = tc.nt;
      EpiRow row0 = epi_decode_row(p, tc, quad * 32 + lane);
      epi_attach_emb(p, row0);
      EpiRow row1 = row0;                                                                  // kMH = 1: never selected
      if (kMH == 2) { row1 = epi_decode_row(p, tc, 128 + quad * 32 + lane); epi_attach_emb(p, row1); }
      if (p.res0 && sub == 0 && pt + n_pairs < pair_tiles) {
        const TileCoord tn = decode_pair_tile(p, pt + n_pairs, (int)rank);
        epi_prefetch_res(p, epi_decode_row(p, tn, quad * 32 + lane), tn.nt * p.block_n, p.block_n);
        if (kMH == 2) epi_prefetch_res(p, epi_decode_row(p, tn, 128 + quad * 32 + lane), tn.nt * p.block_n, p.block_n);
      }
      uint4 res_cur[4], res_nxt[4];
      if (sub < n_items) {
        const int half = (kMH == 2 && sub >= chunks_per_half) ? 1 : 0;
        epi_load_res(p, epi_pick(row0, row1, half != 0), lane, nt * p.block_n + ((sub - half * chunks_per_half) << 5), res_nxt);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * acc_cols);
      for (int item = sub; item < n_items; item += TC_EPI_WARPS / 4) {
        const int half = (kMH == 2 && item >= chunks_per_half) ? 1 : 0;
        const int c0 = (item - half * chunks_per_half) << 5;
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)(half * p.block_n + c0), v);
#pragma unroll
        for (int j = 0; j < 4; ++j) res_cur[j] = res_nxt[j];
        const int nxt = item + TC_EPI_WARPS / 4;
        if (nxt < n_items) {
          const int nh = (kMH == 2 && nxt >= chunks_per_half) ? 1 : 0;
          epi_load_res(p, epi_pick(row0, row1, nh != 0), lane, nt * p.block_n + ((nxt - nh * chunks_per_half) << 5), res_nxt);
        }
        tmem_ld_wait();
        epi_finish(p, epi_pick(row0, row1, half != 0), lane, nt * p.block_n + c0, v, res_cur, s_bias_addr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tempty_bar[acc]), 0));   // report to the leader
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // neither CTA may retire (smem / TMEM) while its peer can still touch it
  if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TcMaps { CUtensorMap a[3]; CUtensorMap b; CUtensorMap out; };

struct TcConvPlan {
  bf16* w_packed = nullptr;     // [n_phase * total_k * Cout][64], K-iterations in the order the producer walks them
  int n_seg = 0; TcSeg seg[3];
  int seg_tensor[3] = {-1, -1, -1};
  int total_k = 0, block_n = 0;
  int bw = 0, bh = 0, bn = 0, mh = 1;   // per-CTA tile box (128 * mh rows)
  int n_phase = 1, Hg = 0, Wg = 0;      // sub-pixel phases (4 for the folded upsample) and the grid the M tiles walk
  bool pair = false;                    // CTA-pair (cta_group::2) kernel
  int a_slot_bytes = 0, b_slot_bytes = 0, n_a = 0, n_b = 0, a_tile_bytes = 0, a_halo_bytes = 0;
  int kc = 64, valid_rows = 0;
  bool tma_store = false;       // staged epilogue + TMA store (short-K layers on the pair kernel)
  int st_bh = 0, st_bn = 0, st_dh = 0, st_dn = 0;   // store box (128 rows) and the origin shift of the second half
  int ring_bytes = TC_RING_BYTES;
  bool b_stat = false;          // weights resident in shared memory across the M tiles of a pair (n_b == total_k)
  bool nswap = false;           // halo boxes over two whole samples per CTA tile, image rows interleaved (TcParams::nswap)
  // fused GroupNorm epilogue (conv_tc2_kernel<.., true>)
  TcConvPlan* alt = nullptr;   // narrow-N variant of the same conv for launches with few M tiles (own weights copy and maps)
  bool gn_ok = false, gn = false, gn_dual = false; int gn_R = 0, gn_seg = 32, gn_ctas = 1, gn_cpg = 0, gn_total_k = 0;
  int cout_pad = 0;            // GEMM N extent (== Cout, or 32 for the zero-padded network head)
  float* bias_pad = nullptr;
  std::map<int, TcMaps> maps;   // per batch size
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int get_encode(Engine& e) {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t st = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (st != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) { e.err = "cuTensorMapEncodeTiled entry point not available"; return CFM_ERR_CUDA; }
  g_encode = (EncodeTiledFn)fn;
  return 0;
}

static bool env_off(const char* name) { const char* v = tuning_env(name); return v && v[0] == '1'; }

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

static int pick_block_n(int Cout) {
  for (int n : {256, 192, 128, 96, 64, 32})
    if (Cout % n == 0) return n;
  return 0;
}

bool tc_conv_supported(const Engine& e, const Op& op) {
  if (!e.bf16 || op.kind != OP_CONV || op.src_is_input) return false;
  if (op.out_is_output && (op.Cout > 32 || env_off("CFM_DISABLE_TC_HEAD"))) return false;   // head: Cout padded to 32
  if (env_off("CFM_DISABLE_TC")) return false;
  if (op.ups && (op.ks != 3 || op.stride != 1 || op.skip0 >= 0 || op.res0 >= 0 || op.out_is_output || env_off("CFM_DISABLE_TC_UPS"))) return false;
  if (op.stride != 1 && op.stride != 2) return false;
  if (op.stride == 2 && env_off("CFM_DISABLE_TC_STRIDE2")) return false;
  if (op.ks != 1 && op.ks != 3) return false;
  if (!op.out_is_output && (pick_block_n(op.Cout) == 0 || op.Cout > TC_MAX_COUT)) return false;
  const int Hg = op.ups ? op.Hin : op.Hout, Wg = op.ups ? op.Win : op.Wout;      // grid the M tiles walk
  if (Wg > 128 || Hg < 1 || Wg < 1) return false;            // a tile row-block spans the full width when W is not a power of two
  if (op.stride == 2 && (op.Hin != 2 * op.Hout || op.Win != 2 * op.Wout || 2 * std::min(op.Wout, 128) > 256)) return false;
  auto chan_ok = [&](int id) { return id < 0 || e.tensors[id].C % 32 == 0; };   // K-iterations of 64 or 32 channels
  if (op.src1 >= 0) return false;                            // main operand is always a single (GN-output) tensor
  if (!chan_ok(op.src0) || !chan_ok(op.skip0) || !chan_ok(op.skip1)) return false;
  auto res_ok = [&](int id) { return id < 0 || e.tensors[id].C % 32 == 0; };
  if (!res_ok(op.res0) || !res_ok(op.res1)) return false;
  return true;
}

// ring depths: D K-iterations of weights in flight and the A slots that feed them (G K-iterations per A slot)
static bool size_rings(TcConvPlan* pl, int G) {
  for (int D = 12; D >= G; --D) {
    const int n_b = D, n_a = (D + G - 1) / G + (G > 1 ? 1 : 0);
    if (n_a > TC_MAX_A || n_b > TC_MAX_B) continue;
    if ((long long)n_a * pl->a_slot_bytes + (long long)n_b * pl->b_slot_bytes > pl->ring_bytes) continue;
    pl->n_a = n_a; pl->n_b = n_b;
    return true;
  }
  return false;
}

// nswap tensor maps list the sample dimension BEFORE the image-row dimension, i.e. with a larger stride on the lower
// dimension.  Check once that the driver encodes such a map (nothing is dereferenced); otherwise those layers keep the
// per-tap loads.
static bool nswap_encodable(Engine& e) {
  static int state = -1;
  if (state < 0) {
    state = 0;
    if (get_encode(e) == 0) {
      CUtensorMap m;
      cuuint64_t dims[4] = {64, 8, 4, 8};
      cuuint64_t strides[3] = {64 * 2, 8 * 8 * 64 * 2, 8 * 64 * 2};
      cuuint32_t box[4] = {64, 8, 2, 10};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      state = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)(uintptr_t)0x10000, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 1 : 0;
    }
  }
  return state == 1;
}

int tc_conv_prepare(Engine& e, Op& op, const std::vector<float>& w, const std::vector<float>& ws, int force_block_n) {
  TcConvPlan* pl = new TcConvPlan();
  const int Cin = op.Cin, ks = op.ks;
  const int Cout = op.out_is_output ? 32 : op.Cout;      // padded rows of W are zero
  pl->cout_pad = Cout;
  pl->block_n = force_block_n ? force_block_n : pick_block_n(Cout);
  static const int pair_min_n = [] { const char* v = tuning_env("CFM_TC_PAIR_MIN_N"); return v ? atoi(v) : 128; }();
  // N >= 128: CTA pair (half the weight tile per SM).  At N = 128 each CTA of the pair also takes two 128-row M-halves
  // per weight tile (a 512 x 128 pair tile): the K-iteration stays 512 cycles long (a 256-cycle K-iteration is
  // shorter than the issue loop) and the shared-memory bytes per MMA cycle drop from 182 to ~135 of the 128 B/clk port.
  // N <= 96: one CTA with a 256-row tile.
  pl->pair = !op.out_is_output && (pl->block_n >= pair_min_n || force_block_n) && !env_off("CFM_DISABLE_TC_2CTA");
  pl->mh = (pl->block_n <= 128 && !force_block_n && !env_off("CFM_DISABLE_TC_MH2")) ? 2 : 1;      // 2 x 128 rows per CTA share one B tile
  // force_block_n: the NARROW variant of a small-map layer (see the end of this function): smallest tiles, most of them
  const int rows = 128 * pl->mh;
  // K-iteration width: 64 channels (SWIZZLE_128B rows) when every operand allows it, else 32 (SWIZZLE_64B)
  pl->kc = 64;
  for (int id : {op.src0, op.skip0, op.skip1}) if (id >= 0 && e.tensors[id].C % 64) pl->kc = 32;
  const int kc = pl->kc, row_bytes = kc * 2;
  // sub-pixel decomposition of (nearest x2 upsample -> 3x3 conv): the M tiles walk the SOURCE grid, four phases,
  // each a 2x2 conv whose taps are sums of the 3x3 taps that land on the same source pixel (4/9 of the MACs,
  // and the upsampled tensor is never materialised)
  pl->n_phase = op.ups ? 4 : 1;
  const int Hg = op.ups ? op.Hin : op.Hout, Wg = op.ups ? op.Win : op.Wout;
  pl->Hg = Hg; pl->Wg = Wg;
  // tile box: full image rows (bw = W; any W <= 128), bh rows of them, and bn whole images when an image is
  // smaller than the tile.  bw*bh*bn <= rows; the unused accumulator rows (e.g. 16 of 128 at 28x28) are ignored.
  pl->bw = Wg;
  pl->bh = std::min(Hg, rows / pl->bw);
  pl->bn = pl->bh == Hg ? std::max(1, rows / (pl->bw * pl->bh)) : 1;
  pl->valid_rows = pl->bw * pl->bh * pl->bn;
  const int eks = op.ups ? 2 : ks;          // taps per axis the kernel walks
  // Widths that are not a multiple of 8 (28, 14): pad the box width to the next multiple of 8 - the extra columns are
  // out of bounds (zero-filled by TMA, never stored) - so that an image row is a whole number of swizzle atoms and
  // the halo views apply: 3x instead of 9x the activation traffic for ~10 % more (idle) accumulator rows.
  {
    const int bwp = (Wg + 7) / 8 * 8;
    if (Wg % 8 && eks > 1 && op.stride == 1 && rows % bwp == 0 && (long long)Hg * bwp * 10 >= (long long)rows * 6 &&
        !env_off("CFM_DISABLE_TC_HALO") && !env_off("CFM_DISABLE_TC_PADW")) {
      pl->bw = bwp; pl->bh = rows / bwp; pl->bn = 1; pl->valid_rows = rows;
    }
  }
  // Short-K layers (stem, qkv, proj_out: <= 8 K-iterations per tile) are bound by their epilogue's row-per-lane stores
  // (32 sectors of 32 different lines per instruction); they stage the tile in shared memory and let TMA write it.
  // The staging boxes come out of the ring space, which these layers do not need.
  {
    static const int max_k = [] { const char* v = tuning_env("CFM_TC_TMA_STORE_MAX_K"); return v ? atoi(v) : 8; }();
    const int tk_all = eks * eks * (Cin / kc) + op.Cskip / kc;
    const bool halves_ok = pl->mh == 1 || pl->bn % 2 == 0 || (pl->bn == 1 && pl->bh % 2 == 0);
    if (pl->pair && !op.out_is_output && !op.out_f32 && !op.ups && pl->block_n % 64 == 0 && Cout % 64 == 0 && tk_all <= max_k &&
        pl->valid_rows == rows && pl->bw == Wg && halves_ok && !env_off("CFM_DISABLE_TC_TMA_STORE")) {
      pl->tma_store = true;
      pl->ring_bytes = TC_STAGE_OFF;
      if (pl->mh == 1) { pl->st_bh = pl->bh; pl->st_bn = pl->bn; }
      else if (pl->bn % 2 == 0) { pl->st_bh = pl->bh; pl->st_bn = pl->bn / 2; pl->st_dn = pl->bn / 2; }
      else { pl->st_bh = pl->bh / 2; pl->st_bn = 1; pl->st_dh = pl->bh / 2; }
    }
  }
  // Maps half the size of a CTA tile (8x8 with 128-row tiles: two whole samples per CTA).  A halo box per sample would
  // leave a gap between the samples' rows, so the row-shifted views could not be one UMMA operand; with the tensor map's
  // dimensions ordered (C, W, N, H) the box lands as [image row][sample][pixel]: a y tap is again a shift by a whole
  // number of swizzle atoms, and these layers pull their activations 3x instead of 9x (they ran at 68 % tensor activity).
  pl->nswap = pl->pair && pl->mh == 1 && eks > 1 && op.stride == 1 && pl->bn == 2 && pl->bh == Hg && pl->bw == Wg && Wg % 8 == 0 &&
              pl->valid_rows == rows && !pl->tma_store && !env_off("CFM_DISABLE_TC_HALO") && !env_off("CFM_DISABLE_TC_NSWAP") &&
              nswap_encodable(e);
  // GroupNorm of the output applied in the epilogue (requested by the plan for a ResBlock's first conv): the CTA tile must
  // hold whole samples or a whole number of tiles must make up a sample, every warp's rows (or half-warp's: 4x4 maps) must
  // lie in one sample, one N tile must span all channels and a group must be 4, 8, 16 or 32 channels.
  // the two-pass epilogue holds an accumulator stage ~2x longer: it only pays where the K loop of a tile is long enough to
  // cover it (measured: wins from 36 K-iterations, loses at 18)
  static const int gn_min_k = [] { const char* v = tuning_env("CFM_TC_GN_MIN_K"); return v ? atoi(v) : 0; }();
  op.gn_fused = false;
  if (pl->pair && !pl->tma_store && !op.ups && !op.out_is_output && !op.out_f32 && Cout == op.Cout &&
      Cout <= TC_GN_MAX_COUT && pl->block_n % (Cout / 32) == 0 && pl->valid_rows == rows && pl->bw == Wg && !(e.cfg.flags & CFM_FLAG_SEPARATE_GROUPNORM) &&
      eks * eks * (Cin / kc) + op.Cskip / kc >= gn_min_k) {
    const int cpg = Cout / 32, HW = Hg * Wg;
    const bool cpg_ok = cpg >= 4 && pow2(cpg) && cpg <= 32;
    int R = 0, seg = 32, ctas = 1;
    if (HW >= rows) { if (HW % rows == 0) { R = rows; ctas = HW / rows; } }
    else if (rows % HW == 0 && pow2(HW)) {
      R = HW;
      if (HW % 32) { if (HW == 16 && pl->mh == 1) seg = 16; else R = 0; }
    }
    if (cpg_ok && R > 0 && ctas <= 32 && (ctas == 1 || (pl->bn == 1 && pl->block_n == Cout))) {
      // capable: the 12 KB of the epilogue's tables come out of the rings whether or not a GroupNorm ends up attached
      // (tc_conv_attach_gn runs after the plan is complete and must not re-pack the weights)
      if (pl->nswap) seg = 8;     // rows of the two samples interleave in runs of 8 (kernel: il)
      pl->gn_ok = true; pl->gn_R = R; pl->gn_seg = seg; pl->gn_ctas = ctas; pl->gn_cpg = cpg;
      pl->gn_total_k = eks * eks * (Cin / kc) + op.Cskip / kc;
      pl->ring_bytes = TC_GN_OFF;
      if (op.gn_request && op.res0 < 0) { pl->gn = true; op.gn_fused = true; op.gn_ctas = ctas; }
    }
  }
  // halo mode: the tile must lie inside one sample (row-shifted views stay contiguous), fill its rows exactly and
  // an image row must be a whole number of 8-row swizzle atoms
  bool halo = eks > 1 && op.stride == 1 && (pl->bn == 1 || pl->nswap) && pl->bw % 8 == 0 && pl->valid_rows == rows && !env_off("CFM_DISABLE_TC_HALO");
  pl->a_tile_bytes = pl->valid_rows * row_bytes;
  pl->a_halo_bytes = (rows + 2 * pl->bw * pl->bn) * row_bytes;
  pl->b_slot_bytes = (pl->pair ? pl->block_n / 2 : pl->block_n) * row_bytes;
  pl->a_slot_bytes = halo ? pl->a_halo_bytes : rows * row_bytes;
  if (halo && !size_rings(pl, eks)) { halo = false; pl->a_slot_bytes = rows * row_bytes; }
  if (!halo && pl->nswap) { pl->nswap = false; if (pl->gn_seg == 8) pl->gn_seg = 32; }
  if (!halo && !size_rings(pl, 1)) { e.err = "conv tile does not fit the shared-memory rings: " + op.name; delete pl; return CFM_ERR_INVALID; }
  // Short-K 1x1 convs (qkv, proj_out) on the pair kernel: every M tile of a pair uses the same N block (the launch makes
  // the number of pairs a multiple of tiles_n), so its whole K extent of weights is loaded once and stays in shared
  // memory; the ring then only carries activations.  Re-loading 64 KB of weights per 64 KB of activations is what kept
  // these layers at ~45 % of both the tensor pipe and DRAM with L2 -> SM traffic as the limiter.
  {
    const int tk = Cin / kc;
    if (pl->pair && ks == 1 && !op.ups && op.skip0 < 0 && op.skip1 < 0 && tk <= TC_MAX_B && !env_off("CFM_DISABLE_TC_BSTAT") &&
        (long long)tk * pl->b_slot_bytes + 3LL * pl->a_slot_bytes <= pl->ring_bytes) {
      pl->b_stat = true;
      pl->n_b = tk;
      pl->n_a = (int)std::min<long long>(TC_MAX_A, (pl->ring_bytes - (long long)tk * pl->b_slot_bytes) / pl->a_slot_bytes);
    }
  }
  pl->seg[0] = {0, Cin / kc, eks, op.stride, halo ? 1 : 0}; pl->seg_tensor[0] = op.src0; pl->n_seg = 1;
  if (op.skip0 >= 0) { pl->seg[pl->n_seg] = {pl->n_seg, e.tensors[op.skip0].C / kc, 1, 1, 0}; pl->seg_tensor[pl->n_seg] = op.skip0; pl->n_seg++; }
  if (op.skip1 >= 0) { pl->seg[pl->n_seg] = {pl->n_seg, e.tensors[op.skip1].C / kc, 1, 1, 0}; pl->seg_tensor[pl->n_seg] = op.skip1; pl->n_seg++; }
  const int n_chunks = Cin / kc;
  pl->total_k = eks * eks * n_chunks + op.Cskip / kc;
  // pack: [phase] K-iteration-major, [Cout][64] per K-iteration, in the producer's order:
  //   halo:  chunk -> x tap -> y tap        plain:  tap (y-major) -> chunk        then the 1x1 skip chunks
  std::vector<bf16> packed((size_t)pl->n_phase * pl->total_k * Cout * kc);
  // weight of main-operand channel c, output o, for walked tap (yi, xi) of phase ph
  auto tap_weight = [&](int ph, int yi, int xi, int c, int o) -> float {
    if (o >= op.Cout) return 0.f;
    if (!op.ups) return w[((size_t)o * Cin + c) * ks * ks + yi * ks + xi];
    // phase (py, px), tap (a, b): source offset (a - 1 + py, b - 1 + px); the 3x3 taps folding onto it are
    // ky in {0} | {1,2} for py = 0 and {0,1} | {2} for py = 1 (same along x); summed in fp32, rounded once
    auto lo = [](int p, int a) { return p == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2); };
    auto hi = [](int p, int a) { return p == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2); };
    const int py = ph >> 1, px = ph & 1;
    float acc = 0.f;
    for (int ky = lo(py, yi); ky <= hi(py, yi); ++ky)
      for (int kx = lo(px, xi); kx <= hi(px, xi); ++kx) acc += w[((size_t)o * Cin + c) * 9 + ky * 3 + kx];
    return acc;
  };
  size_t kiter = 0;
  auto put = [&](int ph, int yi, int xi, int ch) {
    for (int o = 0; o < Cout; ++o)
      for (int j = 0; j < kc; ++j)
        packed[(kiter * Cout + o) * kc + j] = __float2bfloat16(tap_weight(ph, yi, xi, ch * kc + j, o));
    ++kiter;
  };
  for (int ph = 0; ph < pl->n_phase; ++ph) {
    if (halo) {
      for (int ch = 0; ch < n_chunks; ++ch)
        for (int xi = 0; xi < eks; ++xi)
          for (int yi = 0; yi < eks; ++yi) put(ph, yi, xi, ch);
    } else {
      for (int tap = 0; tap < eks * eks; ++tap)
        for (int ch = 0; ch < n_chunks; ++ch) put(ph, tap / eks, tap % eks, ch);
    }
    for (int ch = 0; ch < op.Cskip / kc; ++ch, ++kiter)
      for (int o = 0; o < Cout; ++o)
        for (int j = 0; j < kc; ++j)
          packed[(kiter * Cout + o) * kc + j] = __float2bfloat16(ws[(size_t)o * op.Cskip + ch * kc + j]);
  }
  void* d = nullptr;
  if (cudaMalloc(&d, packed.size() * sizeof(bf16)) != cudaSuccess) { e.err = "cudaMalloc(packed conv weights) failed"; delete pl; return CFM_ERR_OOM; }
  e.owned.push_back(d);
  if (cudaMemcpy(d, packed.data(), packed.size() * sizeof(bf16), cudaMemcpyHostToDevice) != cudaSuccess) { e.err = "weight upload failed"; delete pl; return CFM_ERR_CUDA; }
  pl->w_packed = (bf16*)d;
  if (op.out_is_output) {
    std::vector<float> hb(op.Cout);
    if (cudaMemcpy(hb.data(), op.bias, sizeof(float) * op.Cout, cudaMemcpyDeviceToHost) != cudaSuccess) { e.err = "bias readback failed"; delete pl; return CFM_ERR_CUDA; }
    hb.resize(Cout, 0.f);
    void* bp = nullptr;
    if (cudaMalloc(&bp, sizeof(float) * Cout) != cudaSuccess) { e.err = "cudaMalloc(bias_pad) failed"; delete pl; return CFM_ERR_OOM; }
    e.owned.push_back(bp);
    cudaMemcpy(bp, hb.data(), sizeof(float) * Cout, cudaMemcpyHostToDevice);
    pl->bias_pad = (float*)bp;
  }
  if (force_block_n) { op.tc = pl; return 0; }
  // NARROW variant for the small maps (8x8 and below, 256 channels): at small batches such a layer has fewer 256-row x
  // 256-channel pair tiles than the GPU has CTA pairs (batch 128 at 4x4: 8 tiles of 72 K-iterations on 8 of 74 pairs - the
  // same ~30 us as at batch 1024).  The same conv with 64- or 128-channel N tiles has 4x / 2x as many tiles with a K loop
  // that is 4x / 2x shorter; tc_conv_launch picks it when the wide tiling would leave more than half of the pairs idle.
  if (pl->pair && !pl->tma_store && !op.ups && !op.out_f32 && Cout == 256 && pl->block_n == 256 && Hg * Wg <= 64 && !env_off("CFM_DISABLE_TC_NARROW")) {
    Op alt = op;
    alt.tc = nullptr;
    const int rc = tc_conv_prepare(e, alt, w, ws, Hg * Wg <= 16 ? 64 : 128);
    if (rc) { delete pl; return rc; }
    if (alt.tc && (alt.tc->gn == pl->gn) && (alt.tc->gn_ok == pl->gn_ok)) pl->alt = alt.tc; else delete alt.tc;
  }
  op.tc = pl;
  static DeviceOnce attr_set;
  if (attr_set.pending(e.device)) {
    if (cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(conv_tc2_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(conv_tc2_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(conv_tc2_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(conv_tc2_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess) {
      e.err = "cudaFuncSetAttribute(conv_tc_kernel, smem) failed"; return CFM_ERR_CUDA;
    }
    attr_set.done(e.device);
  }
  return get_encode(e);
}

static int encode_maps(Engine& e, const Op& op, TcConvPlan* pl, int B, TcMaps* m) {
  std::memset(m, 0, sizeof(*m));
  const CUtensorMapSwizzle swz = pl->kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  for (int s = 0; s < pl->n_seg; ++s) {
    const TensorDesc& t = e.tensors[pl->seg_tensor[s]];
    const int st = pl->seg[s].stride;
    const int box_h = pl->seg[s].halo ? pl->bh + 2 : pl->bh * st;
    cuuint64_t dims[4] = {(cuuint64_t)t.C, (cuuint64_t)t.W, (cuuint64_t)t.H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)t.C * 2, (cuuint64_t)t.W * t.C * 2, (cuuint64_t)t.H * t.W * t.C * 2};
    cuuint32_t box[4] = {(cuuint32_t)pl->kc, (cuuint32_t)(pl->bw * st), (cuuint32_t)box_h, (cuuint32_t)pl->bn};
    cuuint32_t estr[4] = {1, (cuuint32_t)st, (cuuint32_t)st, 1};
    if (pl->nswap) {      // (C, W, N, H): the sample index runs inside the image row - every K segment, so all of them see the same tile rows
      std::swap(dims[2], dims[3]); std::swap(strides[1], strides[2]); std::swap(box[2], box[3]);
    }
    CUresult r = g_encode(&m->a[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, tensor_ptr(e, pl->seg_tensor[s], B), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { e.err = "cuTensorMapEncodeTiled(A) failed for " + op.name + " code " + std::to_string((int)r); return CFM_ERR_CUDA; }
  }
  for (int s = pl->n_seg; s < 3; ++s) m->a[s] = m->a[0];
  cuuint64_t dims[2] = {(cuuint64_t)pl->kc, (cuuint64_t)pl->n_phase * pl->total_k * pl->cout_pad};
  cuuint64_t strides[1] = {(cuuint64_t)pl->kc * 2};
  cuuint32_t box[2] = {(cuuint32_t)pl->kc, (cuuint32_t)(pl->b_slot_bytes / (pl->kc * 2))};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(&m->b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl->w_packed, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { e.err = "cuTensorMapEncodeTiled(B) failed for " + op.name + " code " + std::to_string((int)r); return CFM_ERR_CUDA; }
  m->out = m->a[0];
  if (pl->tma_store) {
    const TensorDesc& t = e.tensors[op.out];
    cuuint64_t odims[4] = {(cuuint64_t)t.C, (cuuint64_t)t.W, (cuuint64_t)t.H, (cuuint64_t)B};
    cuuint64_t ostrides[3] = {(cuuint64_t)t.C * 2, (cuuint64_t)t.W * t.C * 2, (cuuint64_t)t.H * t.W * t.C * 2};
    cuuint32_t obox[4] = {64, (cuuint32_t)pl->bw, (cuuint32_t)pl->st_bh, (cuuint32_t)pl->st_bn};
    cuuint32_t oestr[4] = {1, 1, 1, 1};
    r = g_encode(&m->out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, tensor_ptr(e, op.out, B), odims, ostrides, obox, oestr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { e.err = "cuTensorMapEncodeTiled(out) failed for " + op.name + " code " + std::to_string((int)r); return CFM_ERR_CUDA; }
  }
  return 0;
}

int tc_conv_launch(Engine& e, const Op& op, int B, cudaStream_t st, float* out_nchw) {
  TcConvPlan* pl = op.tc;
  if (pl->alt) {      // fewer wide pair tiles than half of the CTA pairs: take the narrow tiling
    const int tiles128 = ((pl->Wg + pl->bw - 1) / pl->bw) * ((pl->Hg + pl->bh - 1) / pl->bh) * ((B + pl->bn - 1) / pl->bn);
    const int pair_tiles = ((tiles128 + 1) / 2) * (pl->cout_pad / pl->block_n) * pl->n_phase;
    if (pair_tiles * 2 <= e.sm_count / 2) pl = pl->alt;
  }
  auto it = pl->maps.find(B);
  if (it == pl->maps.end()) {
    TcMaps m;
    int rc = encode_maps(e, op, pl, B, &m);
    if (rc) return rc;
    it = pl->maps.emplace(B, m).first;
  }
  TcParams p{};
  p.n_seg = pl->n_seg;
  for (int s = 0; s < 3; ++s) p.seg[s] = pl->seg[s];
  p.total_k = pl->total_k;
  p.B = B; p.H = pl->Hg; p.W = pl->Wg; p.n_phase = pl->n_phase;
  p.bw = pl->bw; p.bh = pl->bh; p.bn = pl->bn; p.mh = pl->mh;
  p.tiles_w = (pl->Wg + pl->bw - 1) / pl->bw; p.tiles_h = (pl->Hg + pl->bh - 1) / pl->bh; p.tiles_b = (B + pl->bn - 1) / pl->bn;
  p.kc = pl->kc; p.valid_rows = pl->valid_rows; p.nswap = pl->nswap ? 1 : 0;
  p.tiles_n = pl->cout_pad / pl->block_n;
  p.d_tiles_n.init(p.tiles_n); p.d_phase.init(p.n_phase); p.d_tiles_w.init(p.tiles_w); p.d_tiles_h.init(p.tiles_h);
  p.d_bw.init(p.bw); p.d_bh.init(p.bh);
  p.n_tiles = p.tiles_w * p.tiles_h * p.tiles_b * p.tiles_n * p.n_phase;
  p.block_n = pl->block_n; p.Cout = pl->cout_pad;
  p.a_slot_bytes = pl->a_slot_bytes; p.b_slot_bytes = pl->b_slot_bytes; p.n_a = pl->n_a; p.n_b = pl->n_b;
  p.a_tile_bytes = pl->a_tile_bytes; p.a_halo_bytes = pl->a_halo_bytes;
  p.bias = pl->bias_pad ? pl->bias_pad : op.bias;
  if (op.emb_off >= 0) { p.emb = e.emb_out + op.emb_off; p.emb_stride = e.emb_total; p.emb_row = e.row_of_sample; }
  p.res0 = (const bf16*)tensor_ptr(e, op.res0, B); p.R0 = op.res0 >= 0 ? e.tensors[op.res0].C : 0;
  p.res1 = (const bf16*)tensor_ptr(e, op.res1, B); p.R1 = op.res1 >= 0 ? e.tensors[op.res1].C : 0;
  p.out = (bf16*)tensor_ptr(e, op.out, B);
  if (op.out_f32) p.out_f32 = (float*)tensor_ptr(e, op.out, B);
  if (op.out_is_output) { p.out_nchw = out_nchw; p.cout_real = op.Cout; }
  if (pl->pair) {
    const int tiles128 = p.tiles_w * p.tiles_h * p.tiles_b;
    const int pair_tiles = ((tiles128 + 1) / 2) * p.tiles_n * p.n_phase;
    int n_pairs = std::min(pair_tiles, e.sm_count / 2);
    if (pl->b_stat) {     // pairs walk pt = pair + k * n_pairs and nt = pt % tiles_n: constant per pair iff tiles_n | n_pairs
      n_pairs = std::min(pair_tiles, e.sm_count / 2 / p.tiles_n * p.tiles_n);
      p.b_stat = n_pairs > 0 && n_pairs % p.tiles_n == 0;
      if (!p.b_stat) n_pairs = std::min(pair_tiles, e.sm_count / 2);
    }
    LaunchCfg lc(dim3(2 * n_pairs), dim3(TC_THREADS), TC_SMEM_BYTES, st, 2, pdl_enabled());
    auto kern = pl->mh == 2 ? (pl->gn ? conv_tc2_kernel<2, true> : conv_tc2_kernel<2, false>)
                            : (pl->gn ? conv_tc2_kernel<1, true> : conv_tc2_kernel<1, false>);
    if (pl->gn) {
      p.gn_gamma = op.gamma; p.gn_beta = op.beta; p.gn_eps = 1e-5f; p.gn_cpg = pl->gn_cpg; p.gn_cpg_log2 = ilog2(pl->gn_cpg);
      p.gn_R = pl->gn_R; p.gn_R_log2 = ilog2(pl->gn_R); p.gn_seg = pl->gn_seg; p.gn_ctas = pl->gn_ctas; p.gn_hw = pl->Hg * pl->Wg;
      p.gn_silu = op.silu;
      if (pl->gn_dual) p.out2 = (bf16*)tensor_ptr(e, op.out2, B);
      if (pl->gn_ctas > 1) {
        if (!e.gn_exch || !e.gn_epoch || op.gn_exch_off < 0) { e.err = "internal: GroupNorm exchange buffers missing for " + op.name; return CFM_ERR_INTERNAL; }
        p.gn_exch = e.gn_exch + (size_t)op.gn_exch_off * B * 64;
        p.gn_epoch = e.gn_epoch;
      }
    }
    p.tma_store = pl->tma_store ? 1 : 0; p.st_dh = pl->st_dh; p.st_dn = pl->st_dn;
    cudaError_t ce = cudaLaunchKernelEx(&lc.cfg, kern, it->second.a[0], it->second.a[1], it->second.a[2], it->second.b, it->second.out, p);
    if (ce != cudaSuccess) { e.err = std::string("conv_tc2_kernel launch failed: ") + cudaGetErrorString(ce); return CFM_ERR_CUDA; }
    return 0;
  }
  const int grid = std::min(p.n_tiles, e.sm_count);
  LaunchCfg lc(dim3(grid), dim3(TC_THREADS), TC_SMEM_BYTES, st, 1, pdl_enabled());
  cudaError_t ce = cudaLaunchKernelEx(&lc.cfg, conv_tc_kernel, it->second.a[0], it->second.a[1], it->second.a[2], it->second.b, p);
  if (ce != cudaSuccess) { e.err = std::string("conv_tc_kernel launch failed: ") + cudaGetErrorString(ce); return CFM_ERR_CUDA; }
  return 0;
}

// Fold the GroupNorm `gn` (single source = conv.out, no FiLM) into the epilogue of the conv that produces its input: the
// conv then writes both tensors, conv.out (un-normalised: it is still a residual / skip operand) and gn.out.
bool tc_conv_attach_gn(Engine& e, Op& conv, const Op& gn) {
  TcConvPlan* pl = conv.tc;
  if (!pl || !pl->gn_ok || pl->gn || gn.kind != OP_GN || gn.src1 >= 0 || gn.film || gn.src0 != conv.out || gn.Cin != conv.Cout) return false;
  // writing two tensors from a two-pass epilogue only pays when the K loop of a tile covers it (>= 30 K-iterations) or
  // the map is small enough that the GroupNorm launch it replaces is pure latency (measured: profiles/r02_gn_fold_ab.txt)
  static const int min_k = [] { const char* v = tuning_env("CFM_TC_GN_DUAL_MIN_K"); return v ? atoi(v) : 30; }();
  if (pl->gn_total_k < min_k && pl->Hg * pl->Wg > 256) return false;
  if (pl->gn_ctas > 4) return false;      // many tiles per sample (64x64 maps and up): the exchange costs more than the pass it saves
  pl->gn = true; pl->gn_dual = true;
  if (pl->alt) { pl->alt->gn = true; pl->alt->gn_dual = true; }
  conv.gamma = gn.gamma; conv.beta = gn.beta; conv.silu = gn.silu;
  conv.out2 = gn.out; conv.gn_fused = true; conv.gn_ctas = pl->gn_ctas;
  if (pl->gn_ctas > 1) { conv.gn_exch_off = e.gn_tiles_per_sample; e.gn_tiles_per_sample += pl->gn_ctas; }
  return true;
}

double tc_conv_executed_flops(const Op& op) {
  const TcConvPlan* pl = op.tc;
  if (!pl) return op.flops;
  return 2.0 * pl->Hg * pl->Wg * pl->n_phase * (double)pl->cout_pad * pl->total_k * pl->kc;
}

void tc_conv_release(Engine& e) {
  for (Op& op : e.ops)
    if (op.tc) { op.tc->maps.clear(); if (op.tc->alt) op.tc->alt->maps.clear(); }
}

}  // namespace cfm

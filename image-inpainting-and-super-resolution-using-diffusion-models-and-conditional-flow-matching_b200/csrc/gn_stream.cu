// GroupNorm32 (+FiLM)(+SiLU) over NHWC bf16 for the large feature maps: a persistent, software-pipelined variant of
// groupnorm_bf16_kernel (kernels_bf16.cu; same arithmetic, same fixed summation order per item).
//
// OPT-IN (CFM_ENABLE_GN_STREAM=1), a measured negative result kept parity-green: the staged kernel loads an item,
// reduces it, writes it, and only other resident CTAs overlap those phases; on the 32x32 maps it reaches 4.5-4.8 TB/s,
// with the arithmetic removed it copies at 5.2-5.6 TB/s and its load phase alone reads at 5.2 TB/s.  Here one CTA per
// SM walks its items with THREE shared-memory buffers: TMA (cp.async.bulk.tensor) fetches items k+1 and k+2 while
// item k is reduced, normalised and stored, so reads stay in flight the whole time and one elected thread issues them.
// Result (CIFAR batch 1024, same box): 0.133 ms against 0.120 ms on the 32x32x128 maps, 0.255 against 0.226 at 256
// channels - all 16 warps of the single CTA sit in the same phase (FP-bound statistics, a one-warp scale/shift step,
// the MUFU-bound SiLU pass: 32 K tanh per 64 KB item = 2 K cycles of the 16/clk unit), so the phases of different
// items never overlap the way three independent CTAs per SM do.  A version that wins has to run statistics and
// normalisation of different items concurrently (warp-specialised stages), not just prefetch.
//
// Item = (sample, channel slab), slab in {32, 64, 128} channels = a whole number of groups, all HW pixels of the
// sample, <= 64 KB; no cluster is needed because an item holds every pixel of its groups.
#include <cstring>
#include <map>
#include <tuple>
#include <algorithm>
#include "engine.h"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cfm {

void* tensor_ptr(const Engine& e, int id, int B);

constexpr int GS_THREADS = 512;
constexpr int GS_NBUF = 3;
constexpr int GS_MAX_SLAB = 128;
constexpr int GS_MAX_ITEM = 64 * 1024;

struct GsArgs {
  int C, C0, HW, cpg, slab, slabs, n_items;
  int item_bytes, box_rows, n_box;
  const float* gamma; const float* beta; float eps; int silu;
  const float* film; int film_stride; const int* film_row;
  bf16* out;
};

__device__ __forceinline__ uint4 gs_lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

__global__ void __launch_bounds__(GS_THREADS, 1)
gn_stream_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const GsArgs a) {
  extern __shared__ uint8_t gs_raw[];
  uint8_t* data = (uint8_t*)(((uintptr_t)gs_raw + 127) & ~(uintptr_t)127);
  // layout: data[NBUF][item_bytes] | red[warps * vpp][17] | ch_sum | ch_sq | ch_scale | ch_shift [128 each] | full[NBUF]
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int vpp = a.slab >> 3;                              // 16-byte vectors per pixel: 4, 8 or 16
  float* red = (float*)(data + GS_NBUF * a.item_bytes);
  float* ch_sum = red + (GS_THREADS / 32) * vpp * 17;
  float* ch_sq = ch_sum + GS_MAX_SLAB;
  float* ch_scale = ch_sq + GS_MAX_SLAB;
  float* ch_shift = ch_scale + GS_MAX_SLAB;
  uint64_t* full = (uint64_t*)(ch_shift + GS_MAX_SLAB);

  pdl_launch_dependents();
  if (t == 0) {
    prefetch_tmap(&map0);
    if (a.C0 < a.C) prefetch_tmap(&map1);
    for (int b = 0; b < GS_NBUF; ++b) mbar_init(&full[b], 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();

  const int first = blockIdx.x, stride = gridDim.x;
  const int n_local = first < a.n_items ? (a.n_items - first + stride - 1) / stride : 0;
  auto issue = [&](int k) {                                  // one thread: all boxes of local item k -> buffer k % NBUF
    const int item = first + k * stride, buf = k % GS_NBUF;
    const int b = item / a.slabs, c_base = (item - b * a.slabs) * a.slab;
    const CUtensorMap* map = c_base < a.C0 ? &map0 : &map1;
    const int col = c_base < a.C0 ? c_base : c_base - a.C0;
    uint8_t* dst = data + buf * a.item_bytes;
    mbar_expect_tx(&full[buf], (uint32_t)a.item_bytes);
    const int box_bytes = a.box_rows * a.slab * 2;
    for (int j = 0; j < a.n_box; ++j) tma_load_2d(dst + j * box_bytes, map, &full[buf], col, b * a.HW + j * a.box_rows);
  };
  if (t == 0)
    for (int k = 0; k < GS_NBUF - 1 && k < n_local; ++k) issue(k);

  const int q = t % vpp;                                    // GS_THREADS % vpp == 0: fixed vector slot per thread
  const int p0 = t / vpp, pstep = GS_THREADS / vpp;
  const int nvec = a.item_bytes >> 4;
  const float inv_n = 1.0f / (float)(a.cpg * a.HW);

  for (int k = 0; k < n_local; ++k) {
    const int buf = k % GS_NBUF;
    // buffer (k - 1) % NBUF was released by the barrier that ended iteration k - 1: refill it two items ahead
    if (t == 0 && k + GS_NBUF - 1 < n_local) { fence_proxy_async(); issue(k + GS_NBUF - 1); }
    const int item = first + k * stride;
    const int b = item / a.slabs, c_base = (item - b * a.slabs) * a.slab;
    const uint32_t base = smem_u32(data) + (uint32_t)(buf * a.item_bytes);
    mbar_wait(&full[buf], (uint32_t)((k / GS_NBUF) & 1));

    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; ss[j] = 0.f; }
#pragma unroll 4
    for (int v = t; v < nvec; v += GS_THREADS) {
      const uint4 r = gs_lds_u4(base + (uint32_t)v * 16u);
      const __nv_bfloat162* h2 = (const __nv_bfloat162*)&r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        s[2 * j] += f.x; ss[2 * j] = fmaf(f.x, f.x, ss[2 * j]);
        s[2 * j + 1] += f.y; ss[2 * j + 1] = fmaf(f.y, f.y, ss[2 * j + 1]);
      }
    }
    // lanes that share a vector slot are vpp apart: butterfly over the lane bits above log2(vpp), then one row per
    // (warp, slot) in shared memory and a fixed-order sum over the warps
    for (int m = vpp; m < 32; m <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += __shfl_xor_sync(0xffffffffu, s[j], m); ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], m); }
    }
    if (lane < vpp) {
      float* rp = red + (warp * vpp + lane) * 17;
#pragma unroll
      for (int j = 0; j < 8; ++j) { rp[j] = s[j]; rp[8 + j] = ss[j]; }
    }
    __syncthreads();
    if (t < a.slab) {
      const int qq = t >> 3, j = t & 7;
      float ts = 0.f, tq = 0.f;
      for (int w = 0; w < GS_THREADS / 32; ++w) { const float* rp = red + (w * vpp + qq) * 17; ts += rp[j]; tq += rp[8 + j]; }
      ch_sum[t] = ts; ch_sq[t] = tq;
    }
    __syncthreads();
    if (t < a.slab) {
      const int g0 = (t / a.cpg) * a.cpg;
      float gs = 0.f, gq = 0.f;
      for (int j = 0; j < a.cpg; ++j) { gs += ch_sum[g0 + j]; gq += ch_sq[g0 + j]; }
      const float mean = gs * inv_n;
      const float var = fmaxf(gq * inv_n - mean * mean, 0.f);
      const float rstd = rsqrtf(var + a.eps);
      float sc_ = rstd * a.gamma[c_base + t];
      float sh_ = a.beta[c_base + t] - mean * sc_;
      if (a.film) {
        const float* f = a.film + (long long)a.film_row[b] * a.film_stride;
        const float m = 1.0f + f[c_base + t];
        sc_ *= m; sh_ = sh_ * m + f[a.C + c_base + t];
      }
      if (a.silu) { sc_ *= 0.5f; sh_ *= 0.5f; }          // the activation works on h = y/2: silu(y) = h*tanh(h) + h
      ch_scale[t] = sc_; ch_shift[t] = sh_;
    }
    __syncthreads();
    float sc8[8], sh8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc8[j] = ch_scale[q * 8 + j]; sh8[j] = ch_shift[q * 8 + j]; }
    bf16* op = a.out + ((long long)b * a.HW + p0) * a.C + c_base + q * 8;
    const long long ostep = (long long)pstep * a.C;
#pragma unroll 4
    for (int v = t; v < nvec; v += GS_THREADS, op += ostep) {
      const uint4 r = gs_lds_u4(base + (uint32_t)v * 16u);
      const __nv_bfloat162* h2 = (const __nv_bfloat162*)&r;
      uint4 o4;
      __nv_bfloat162* o2 = (__nv_bfloat162*)&o4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        float y0 = fmaf(f.x, sc8[2 * j], sh8[2 * j]), y1 = fmaf(f.y, sc8[2 * j + 1], sh8[2 * j + 1]);
        if (a.silu) { y0 = silu_from_half(y0); y1 = silu_from_half(y1); }
        o2[j] = __floats2bfloat162_rn(y0, y1);
      }
      *(uint4*)op = o4;
    }
    __syncthreads();          // every thread is done with buffer `buf` (and with red / ch_*) before it is refilled
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct GsGeom { int cpg, slab, item_bytes, box_rows, n_box; size_t smem; bool ok; };

static GsGeom gs_geometry(const Engine& e, const Op& op) {
  GsGeom g{};
  const int C = op.Cin, HW = op.Hin * op.Win;
  const int C0 = e.tensors[op.src0].C;
  static const int min_bytes = [] { const char* v = getenv("CFM_GN_STREAM_MIN_BYTES"); return v ? atoi(v) : 16 * 1024; }();
  static const int max_slab = [] { const char* v = getenv("CFM_GN_STREAM_MAX_SLAB"); return v ? atoi(v) : GS_MAX_SLAB; }();
  g.cpg = C / 32;
  if (g.cpg < 1 || C % 32) return g;
  g.box_rows = HW <= 256 ? HW : 256;
  if (HW % g.box_rows) return g;
  g.n_box = HW / g.box_rows;
  for (int slab = std::min(GS_MAX_SLAB, max_slab); slab >= 32; slab >>= 1) {
    if (slab % g.cpg || C % slab || C0 % slab) continue;              // whole groups; a slab never straddles the concat
    if ((long long)HW * slab * 2 > GS_MAX_ITEM) continue;
    g.slab = slab;
    break;
  }
  if (!g.slab) return g;
  g.item_bytes = HW * g.slab * 2;
  if (g.item_bytes < min_bytes || g.item_bytes % 128) return g;
  g.smem = 128 + (size_t)GS_NBUF * g.item_bytes + sizeof(float) * ((size_t)(GS_THREADS / 32) * (g.slab / 8) * 17 + 4 * GS_MAX_SLAB) + 64;
  g.ok = g.smem <= 220 * 1024;
  return g;
}

bool gn_stream_supported(const Engine& e, const Op& op) {
  if (!e.bf16 || op.kind != OP_GN) return false;
  const char* on = getenv("CFM_ENABLE_GN_STREAM");
  if (!on || on[0] != '1') return false;
  if (e.tensors[op.src0].C % 8) return false;
  return gs_geometry(e, op).ok;
}

typedef CUresult (*GsEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static GsEncodeFn g_gs_encode = nullptr;
struct GsMaps { CUtensorMap m0, m1; };
static std::map<std::tuple<const void*, const void*, int, int, long long, int, int>, GsMaps> g_gs_maps;

static int gs_encode(Engine& e, CUtensorMap* m, const void* ptr, int Csrc, long long rows, int slab, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)Csrc, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Csrc * 2};
  cuuint32_t box[2] = {(cuuint32_t)slab, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_gs_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { e.err = "cuTensorMapEncodeTiled(groupnorm) failed, code " + std::to_string((int)r); return CFM_ERR_CUDA; }
  return 0;
}

int gn_stream_launch(Engine& e, const Op& op, int B, cudaStream_t st) {
  static DeviceOnce attr;
  if (attr.pending(e.device)) {
    if (cudaFuncSetAttribute(gn_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
      e.err = "cudaFuncSetAttribute(gn_stream_kernel) failed"; return CFM_ERR_CUDA;
    }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { e.err = "cuTensorMapEncodeTiled unavailable"; return CFM_ERR_CUDA; }
    g_gs_encode = (GsEncodeFn)fn;
    attr.done(e.device);
  }
  const GsGeom g = gs_geometry(e, op);
  const int HW = op.Hin * op.Win;
  const void* p0 = tensor_ptr(e, op.src0, B);
  const void* p1 = op.src1 >= 0 ? tensor_ptr(e, op.src1, B) : nullptr;
  const int C0 = e.tensors[op.src0].C, C1 = op.src1 >= 0 ? e.tensors[op.src1].C : 0;
  const long long rows = (long long)B * HW;
  auto key = std::make_tuple(p0, p1, C0, C1, rows, g.slab, g.box_rows);
  auto it = g_gs_maps.find(key);
  if (it == g_gs_maps.end()) {
    GsMaps m;
    std::memset(&m, 0, sizeof(m));
    int rc = gs_encode(e, &m.m0, p0, C0, rows, g.slab, g.box_rows);
    if (rc) return rc;
    if (p1) { if ((rc = gs_encode(e, &m.m1, p1, C1, rows, g.slab, g.box_rows))) return rc; }
    else m.m1 = m.m0;
    it = g_gs_maps.emplace(key, m).first;
  }
  GsArgs a{};
  a.C = op.Cin; a.C0 = C0; a.HW = HW; a.cpg = g.cpg; a.slab = g.slab; a.slabs = op.Cin / g.slab; a.n_items = B * a.slabs;
  a.item_bytes = g.item_bytes; a.box_rows = g.box_rows; a.n_box = g.n_box;
  a.gamma = op.gamma; a.beta = op.beta; a.eps = 1e-5f; a.silu = op.silu;
  if (op.film) { a.film = e.emb_out + op.emb_off; a.film_stride = e.emb_total; a.film_row = e.row_of_sample; }
  a.out = (bf16*)tensor_ptr(e, op.out, B);
  LaunchCfg lc(dim3((unsigned)std::min(a.n_items, e.sm_count)), dim3(GS_THREADS), g.smem, st, 1, pdl_enabled());
  if (cudaLaunchKernelEx(&lc.cfg, gn_stream_kernel, it->second.m0, it->second.m1, a) != cudaSuccess) { e.err = "groupnorm (stream) launch failed"; return CFM_ERR_CUDA; }
  return 0;
}

}  // namespace cfm

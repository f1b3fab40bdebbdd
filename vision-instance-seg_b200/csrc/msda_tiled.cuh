// Tiled kernels for the dense ("encoder") call site of multi-scale deformable attention on B200 (sm_100a).
//
// When every value pixel is also a query (Lq == S: the six pixel-decoder encoder layers of MaskDINO, SURVEY.md §8a rows
// a3/a4/a6/a7 at BASELINE configs 3 and 5) the sampling points of spatially close queries land close together on
// every level.  The direct kernels in msda_kernels.cu cannot use that: every corner row of every point is a 64-byte
// gather through L1TEX (one 128-byte wavefront per row, half of it wasted) and, in the backward, a 64-byte reduction that
// the L2 resolves (11.4 GB of reduction payload per launch at config 3 -- the measured ceiling of round 1).
//
// Here a CTA works on one (image, tile, head): the tile is a 16 x 8 pixel region of the finest level together with
// ALL queries whose pixel centre falls into that region (128 + 32 + 8 + 2 queries for a /8../64 pyramid), and for
// every level the CTA keeps a *window* -- the region scaled to that level plus a halo -- of that head's 64-byte
// channel slices in shared memory.
//
//   forward  (msda_fwd_tiled_kernel)     window filled with cp.async (zero-filled outside the level, which IS the
//                                        zero padding of the operator); sampling points are turned into 16-bit window
//                                        row numbers + packed 16-bit weights once per point (lane = point); the gather loop
//                                        then runs 8 lanes per point: 4 lanes per pixel of an x-adjacent corner pair, so
//                                        that every LDS.128 reads 128 contiguous bytes (bank-conflict free, 128 B/clk/SM).
//   backward (msda_bwd_dots_tiled_kernel) same windows; per corner <value, grad_out> with the mixed-precision FMA, a
//                                        14-shuffle reduce-scatter per 8 points, grad_sampling_loc / grad_attn_weight
//                                        written once, coalesced.  Also reduces max|grad_out| for the fp16 scale.
//            (msda_bwd_scatter_tiled_kernel) grad_value without read-modify-write: the tile's corner rows of one
//                                        level are counting-sorted by destination pixel (integer shared-memory
//                                        atomics only -- float ones are CAS loops on sm_100a), each lane group then
//                                        sums a contiguous range of the sorted list in fp32 registers, reading the
//                                        staged grad_out rows from shared memory, and issues ONE packed fp16 reduction
//                                        per (destination row, tile) into the scaled fp16 accumulator of
//                                        msda_kernels.cu: ~8x fewer reductions reach the L2 than with one per corner row.
//
// A point whose footprint leaves its window (or any point, if the level shapes do not nest the way a pyramid does) takes a
// slow path that gathers / reduces through global memory exactly like the direct kernels, so the result does not
// depend on the sampling pattern -- only the speed does.  Shapes stay on the device (int64 tensors), so the
// grid is persistent and each CTA derives tiles and windows itself.
// This header is included from msda_kernels.cu INSIDE namespace msda, after the helpers it uses (level_of,
// f16_accum_scale, the msda_common.cuh primitives).
#pragma once

namespace tiled {

constexpr int kTW = 16, kTH = 8;           // tile of the finest level, pixels
constexpr int kMaxL = 8;                   // levels the tiled path supports
constexpr int kMaxLP = 64;                 // L*P the tiled path supports
constexpr int kHaloMax = 6;                // pixels of halo per side (3 sigma of N(0, 2^2) px offsets)
constexpr int kWinRowsCap = 1440;          // window rows (64 bytes each) of one head per CTA
constexpr int kOvfPts = 16;                // points per chunk whose footprint may leave the windows (private row copies)
constexpr int kZeroRow = kWinRowsCap;      // rows kZeroRow, kZeroRow + 1: always zero
constexpr int kOvfRow0 = kWinRowsCap + 2;  // 4 rows (tl, tr, bl, br) per overflow point
constexpr int kWinRowsAll = kWinRowsCap + 2 + 4 * kOvfPts;
constexpr int kRowBytes = 64;              // D = 32 channels x 16 bit
constexpr int kD = 32;
constexpr int kThreadsT = 256;
constexpr int kGroups = kThreadsT / 8;     // 8-lane groups per CTA
constexpr int kPF = 4;                     // points per thread per chunk
constexpr int kChunkPts = kThreadsT * kPF; // record slots per chunk
constexpr int kQC = 64;                    // queries per chunk (at most; kChunkPts / LP8 if that is smaller)
// Window fill: 16-byte cp.async (0, default) or one TMA bulk copy per window pixel (1: fill_windows_bulk).  Measured at
// cfg3 with -DMSDA_TILED_FILL_BULK=1: forward 1.24 ms against 0.99 ms, dots 1.50 against 1.25 ms.  UBLKCP takes its
// operands from uniform registers, so per-lane copies compile to an ELECT / R2UR loop of ~8 instructions per 64-byte
// copy (11 K warp-instructions per work item against 1.9 K for the cp.async loop); only a tensor-map copy (one
// instruction per level) would pay, and that needs the level shapes on the host.
#ifndef MSDA_TILED_FILL_BULK
#define MSDA_TILED_FILL_BULK 0
#endif
constexpr bool kFillBulk = MSDA_TILED_FILL_BULK != 0;
constexpr uint32_t kRowFallback = 0xFFFFu; // record marker: footprint outside the windows and no overflow slot left
constexpr uint32_t kRowsNull = static_cast<uint32_t>(kZeroRow) | (static_cast<uint32_t>(kZeroRow) << 16);

struct Geom {
  int L, lf, halo;
  int tiles_x, tiles_y;
  int H[kMaxL], W[kMaxL], start[kMaxL];       // start: level_start_index (value rows)
  int qstart[kMaxL];                          // first query of the level: prefix sums of H*W (queries = pixels)
  int wdx[kMaxL], wdy[kMaxL], base[kMaxL];   // window dims (0: no window) and first window row, per level
  int rows_total;                             // window rows in use
  // per tile
  int wx0[kMaxL], wy0[kMaxL];                 // level pixel of window cell (0, 0)
  int qxa[kMaxL], qya[kMaxL], qnx[kMaxL], qny[kMaxL];
  int qoff[kMaxL + 1];                        // prefix of the tile's query counts per level
  uint32_t fill_bytes;                        // bytes the bulk copies of this tile's windows will deliver (in-level cells)
};

__device__ __forceinline__ int ceil_div_i(int a, int b) { return (a + b - 1) / b; }

// non-negative 64-bit quotient, through a 32-bit division whenever the operands allow it (64-bit division is a ~100
// instruction subroutine, and this runs once per tile)
__device__ __forceinline__ long long div_fast(long long num, long long den) {
  if (((num | den) >> 32) == 0) return static_cast<long long>(static_cast<uint32_t>(num) / static_cast<uint32_t>(den));
  return num / den;
}

// first query column (row) of level extent `wl` that belongs to tile column (row) t of the finest level:
// membership of query x is floor((x + 0.5) * wf / wl / tile), a monotone map, so the tiles partition every level
__device__ __forceinline__ int tile_first(long long t, int tile, int wl, int wf) {
  const long long num = 2ll * tile * t * wl - wf;
  if (num <= 0) return 0;
  const long long v = div_fast(num + 2ll * wf - 1, 2ll * wf);
  return static_cast<int>(v < wl ? v : wl);
}

// once per CTA (thread 0): levels, finest level, tile grid, halo and window dims
__device__ inline void geom_init(Geom& g, const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi, int L, int cap) {
  g.L = L;
  long long best = -1;
  int lf = 0, qs = 0;
  for (int l = 0; l < L; ++l) {
    g.H[l] = static_cast<int>(shapes[2 * l]);
    g.W[l] = static_cast<int>(shapes[2 * l + 1]);
    g.start[l] = static_cast<int>(lsi[l]);
    g.qstart[l] = qs;
    qs += g.H[l] * g.W[l];
    const long long hw = static_cast<long long>(g.H[l]) * g.W[l];
    if (hw > best) { best = hw; lf = l; }
  }
  g.lf = lf;
  const int Wf = g.W[lf] > 0 ? g.W[lf] : 1, Hf = g.H[lf] > 0 ? g.H[lf] : 1;
  g.tiles_x = ceil_div_i(Wf, kTW);
  g.tiles_y = ceil_div_i(Hf, kTH);
  int halo = kHaloMax;
  for (; halo >= 0; --halo) {
    long long rows = 0;
    for (int l = 0; l < L; ++l) {
      const long long rx = (static_cast<long long>(kTW) * g.W[l] + Wf - 1) / Wf, ry = (static_cast<long long>(kTH) * g.H[l] + Hf - 1) / Hf;
      rows += (rx + 2 * halo + 1) * (ry + 2 * halo + 1);
    }
    if (rows <= cap) break;
  }
  g.halo = halo;
  int base = 0;
  for (int l = 0; l < L; ++l) {
    if (halo >= 0 && g.H[l] > 0 && g.W[l] > 0) {
      g.wdx[l] = static_cast<int>((static_cast<long long>(kTW) * g.W[l] + Wf - 1) / Wf) + 2 * halo + 1;
      g.wdy[l] = static_cast<int>((static_cast<long long>(kTH) * g.H[l] + Hf - 1) / Hf) + 2 * halo + 1;
    } else {
      g.wdx[l] = 0; g.wdy[l] = 0;
    }
    g.base[l] = base;
    base += g.wdx[l] * g.wdy[l];
  }
  g.rows_total = base;
}

// once per work item (warp 0, lane = level): window origins and the tile's queries
__device__ __forceinline__ void geom_tile(Geom& g, int tile, int lane) {
  const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
  const int Wf = g.W[g.lf] > 0 ? g.W[g.lf] : 1, Hf = g.H[g.lf] > 0 ? g.H[g.lf] : 1;
  int cnt = 0, cells = 0;
  if (lane < g.L) {
    const int l = lane;
    g.wx0[l] = static_cast<int>(div_fast(2ll * kTW * tx * g.W[l] + Wf, 2ll * Wf)) - g.halo;
    g.wy0[l] = static_cast<int>(div_fast(2ll * kTH * ty * g.H[l] + Hf, 2ll * Hf)) - g.halo;
    const int xa = tile_first(tx, kTW, g.W[l], Wf), xb = tile_first(tx + 1, kTW, g.W[l], Wf);
    const int ya = tile_first(ty, kTH, g.H[l], Hf), yb = tile_first(ty + 1, kTH, g.H[l], Hf);
    g.qxa[l] = xa; g.qya[l] = ya; g.qnx[l] = xb - xa; g.qny[l] = yb - ya;
    cnt = (xb - xa) * (yb - ya);
    // window cells that lie inside the level: the ones fill_windows_bulk copies (the others are zero-filled)
    const int x0 = max(g.wx0[l], 0), x1 = min(g.wx0[l] + g.wdx[l], g.W[l]);
    const int y0 = max(g.wy0[l], 0), y1 = min(g.wy0[l] + g.wdy[l], g.H[l]);
    cells = max(x1 - x0, 0) * max(y1 - y0, 0);
  }
#pragma unroll
  for (int sft = kMaxL / 2; sft >= 1; sft >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, sft);
  if (lane == 0) g.fill_bytes = static_cast<uint32_t>(cells) * kRowBytes;
  int incl = cnt;
#pragma unroll
  for (int sft = 1; sft < kMaxL; sft <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, sft);
    if (lane >= sft) incl += t;
  }
  if (lane < g.L) g.qoff[lane] = incl - cnt;
  if (lane == g.L - 1) g.qoff[g.L] = incl;
}

// global query index of the tile's i-th query (-1 past the end, or if the shape tensor describes more pixels than queries)
__device__ __forceinline__ int tile_query(const Geom& g, int i, int Lq) {
  if (i >= g.qoff[g.L]) return -1;
  int l = 0;
  while (l + 1 < g.L && i >= g.qoff[l + 1]) ++l;
  const int j = i - g.qoff[l];
  const int iy = static_cast<int>(static_cast<uint32_t>(j) / static_cast<uint32_t>(g.qnx[l])), ix = j - iy * g.qnx[l];
  const int q = g.qstart[l] + (g.qya[l] + iy) * g.W[l] + g.qxa[l] + ix;
  return q < Lq ? q : -1;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// per-level constants a thread keeps in registers
struct LevelC {
  float Hf, Wf;
  int H, W, start, wx0, wy0, wdx, wdy, base;
};
__device__ __forceinline__ LevelC level_consts(const Geom& g, int l) {
  LevelC c;
  c.H = g.H[l]; c.W = g.W[l]; c.start = g.start[l];
  c.Hf = static_cast<float>(c.H); c.Wf = static_cast<float>(c.W);
  c.wx0 = g.wx0[l]; c.wy0 = g.wy0[l]; c.wdx = g.wdx[l]; c.wdy = g.wdy[l]; c.base = g.base[l];
  return c;
}

// Fill the CTA's windows with head `head` of image `b`: one warp per window row, 4 lanes per pixel (16 bytes each, so a
// warp's lanes write 512 contiguous bytes); cells outside the level are zero-filled, which implements the operator's
// zero padding.  Byte offsets inside one image fit 32 bits (vec_supported).
template <typename T>
__device__ __forceinline__ void fill_windows(const Geom& g, uint32_t win, const T* __restrict__ value, int b, int S, int vps,
                                             int head, int warp, int lane, int nwarps) {
  int task0 = 0;
  const uint32_t pixb = static_cast<uint32_t>(vps) * static_cast<uint32_t>(sizeof(T));
  const char* img = reinterpret_cast<const char*>(value) + static_cast<size_t>(b) * S * pixb +
                    static_cast<size_t>(head) * kRowBytes + (lane & 3) * 16;
  const int lx = lane >> 2;
  for (int l = 0; l < g.L; ++l) {
    const int wdx = g.wdx[l], wdy = g.wdy[l];
    const int H = g.H[l], W = g.W[l], wx0 = g.wx0[l], wy0 = g.wy0[l];
    const char* lvl = img + static_cast<size_t>(g.start[l]) * pixb;
    const uint32_t dlvl = win + static_cast<uint32_t>(g.base[l]) * kRowBytes + static_cast<uint32_t>(lane) * 16;
    // warp `warp` takes window rows wy with (task0 + wy) % nwarps == warp
    int wy = warp - task0 % nwarps;
    if (wy < 0) wy += nwarps;
    for (; wy < wdy; wy += nwarps) {
      const int gy = wy0 + wy;
      const bool yin = static_cast<unsigned>(gy) < static_cast<unsigned>(H);
      uint32_t dst = dlvl + static_cast<uint32_t>(wy * wdx) * kRowBytes;
      const char* srow = lvl + static_cast<size_t>(static_cast<uint32_t>(yin ? gy : 0) * static_cast<uint32_t>(W)) * pixb;
      int gx = wx0 + lx;
      for (int x = lx; x < wdx; x += 8, gx += 8, dst += 8 * kRowBytes) {
        const bool in = yin && static_cast<unsigned>(gx) < static_cast<unsigned>(W);
        const uint32_t off = static_cast<uint32_t>(in ? gx : 0) * pixb;
        cp_async16_zfill(dst, srow + off, in);
      }
    }
    task0 += wdy;
  }
}

// ---- TMA bulk copies (descriptor-free: shapes stay on the device) + mbarrier completion ----------------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MSDA_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MSDA_DONE;\n"
      "bra MSDA_WAIT;\n"
      "MSDA_DONE:\n"
      "}\n" ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Window fill through the TMA unit: one 64-byte bulk copy (SASS UBLKCP) per window pixel that lies inside its level --
// one lane per pixel, one warp per window row -- completing on `mbar`, whose expected byte count (g.fill_bytes, armed by
// the caller) is exactly the bytes issued here; cells outside the level get plain zero stores (the zero padding).
// ~17x fewer instructions than the 16-byte cp.async loop, and the copies do not pass through the LSU.
template <typename T>
__device__ __forceinline__ void fill_windows_bulk(const Geom& g, unsigned char* win_ptr, uint32_t win, uint32_t mbar,
                                                  const T* __restrict__ value, int b, int S, int vps, int head, int warp,
                                                  int lane, int nwarps) {
  int task0 = 0;
  const uint32_t pixb = static_cast<uint32_t>(vps) * static_cast<uint32_t>(sizeof(T));
  const char* img = reinterpret_cast<const char*>(value) + static_cast<size_t>(b) * S * pixb + static_cast<size_t>(head) * kRowBytes;
  for (int l = 0; l < g.L; ++l) {
    const int wdx = g.wdx[l], wdy = g.wdy[l];
    const int H = g.H[l], W = g.W[l], wx0 = g.wx0[l], wy0 = g.wy0[l];
    const char* lvl = img + static_cast<size_t>(g.start[l]) * pixb;
    int wy = warp - task0 % nwarps;
    if (wy < 0) wy += nwarps;
    for (; wy < wdy; wy += nwarps) {
      const int gy = wy0 + wy;
      const bool yin = static_cast<unsigned>(gy) < static_cast<unsigned>(H);
      const uint32_t row = static_cast<uint32_t>(g.base[l] + wy * wdx);
      const char* srow = lvl + static_cast<size_t>(static_cast<uint32_t>(yin ? gy : 0) * static_cast<uint32_t>(W)) * pixb;
      for (int x = lane; x < wdx; x += 32) {
        const int gx = wx0 + x;
        if (yin && static_cast<unsigned>(gx) < static_cast<unsigned>(W)) {
          bulk_copy_g2s(win + (row + x) * kRowBytes, srow + static_cast<uint32_t>(gx) * pixb, kRowBytes, mbar);
        } else {
          uint4* cell = reinterpret_cast<uint4*>(win_ptr + static_cast<size_t>(row + x) * kRowBytes);
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
          cell[0] = z; cell[1] = z; cell[2] = z; cell[3] = z;
        }
      }
    }
    task0 += wdy;
  }
}

// One sampling point, the part every tiled kernel needs: pixel coordinates of the top-left corner, fractions, gate.
struct PointGeo {
  int ix, iy;
  float lw, lh;
  bool inside;        // upstream's gate: -1 < h_im < H and -1 < w_im < W  (false for NaN locations)
};
__device__ __forceinline__ PointGeo point_geo(float x, float y, float Hf, float Wf) {
  PointGeo p;
  const float h_im = __fsub_rn(__fmul_rn(y, Hf), 0.5f);
  const float w_im = __fsub_rn(__fmul_rn(x, Wf), 0.5f);
  p.inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < Hf) && (w_im < Wf);
  floor_split(h_im, p.lh, p.iy);
  floor_split(w_im, p.lw, p.ix);
  return p;
}

// Window rows (top | bottom << 16) of a point's two corner pairs.  A point that fails the gate reads the zero rows.  A
// footprint that leaves its window gets private copies of its four corner rows in the overflow area (cp.async issued
// here, completed by the caller's wait + barrier; corners outside the level are zero-filled); when the chunk's overflow
// slots are used up the marker tells the gather loop to fetch from global memory itself.
template <typename T>
__device__ __forceinline__ uint32_t resolve_rows(const PointGeo& p, const LevelC& c, uint32_t win, const T* __restrict__ img,
                                                 int vps, int head, int* s_ovf) {
  if (!p.inside) return kRowsNull;
  const int wxr = p.ix - c.wx0, wyr = p.iy - c.wy0;
  if (wxr >= 0 && wyr >= 0 && wxr <= c.wdx - 2 && wyr <= c.wdy - 2) {
    const uint32_t top = static_cast<uint32_t>(c.base + wyr * c.wdx + wxr);
    return top | ((top + static_cast<uint32_t>(c.wdx)) << 16);
  }
  const int slot = atomicAdd(s_ovf, 1);
  if (slot >= kOvfPts) return kRowFallback | (kRowFallback << 16);
  const uint32_t r0 = static_cast<uint32_t>(kOvfRow0 + 4 * slot);
  const size_t pix_bytes = static_cast<size_t>(vps) * sizeof(T);
  const char* base = reinterpret_cast<const char*>(img) + static_cast<size_t>(head) * kRowBytes;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = p.ix + (k & 1), y = p.iy + (k >> 1);
    const bool in = static_cast<unsigned>(x) < static_cast<unsigned>(c.W) && static_cast<unsigned>(y) < static_cast<unsigned>(c.H);
    const char* src = in ? base + (static_cast<size_t>(c.start) + static_cast<size_t>(y) * c.W + x) * pix_bytes
                         : reinterpret_cast<const char*>(img);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) cp_async16_zfill(win + (r0 + k) * kRowBytes + ch * 16, src + ch * 16, in);
  }
  return r0 | ((r0 + 2) << 16);
}

// Slow path: the 16 bytes of this lane's pixel (side s: left / right column, chunk c) of the top and bottom row of a
// footprint, straight from global memory; corners outside the level read as zero (never loaded).
template <typename T>
__device__ __forceinline__ void fetch_global_pair(const T* __restrict__ img, int start, int H, int W, int vps, int head,
                                                  int ix, int iy, int s, int c, uint4& top, uint4& bot) {
  const int x = ix + s;
  const bool xin = x >= 0 && x < W;
  const char* base = reinterpret_cast<const char*>(img) + static_cast<size_t>(head) * kRowBytes + c * 16;
  top = make_uint4(0u, 0u, 0u, 0u);
  bot = top;
  if (xin && iy >= 0 && iy < H)
    top = ldg16(base + (static_cast<size_t>(start) + static_cast<size_t>(iy) * W + x) * static_cast<size_t>(vps) * sizeof(T));
  if (xin && iy + 1 >= 0 && iy + 1 < H)
    bot = ldg16(base + (static_cast<size_t>(start) + static_cast<size_t>(iy + 1) * W + x) * static_cast<size_t>(vps) * sizeof(T));
}

__device__ __forceinline__ int round_up8(int v) { return (v + 7) & ~7; }

// How the kChunkPts record slots of a chunk map to threads: slot = qi * LP8 + lp, thread tid handles slots
// tid + it * kThreadsT.  When LP8 divides the CTA size a thread's lp (hence its level) never changes.
struct SlotMap {
  int LP, LP8, qc;           // points per pair, padded, queries per chunk
  bool fixed;
  int lp_fixed, qi_fixed, qstep;
  __device__ __forceinline__ void init(int L, int P, int tid) {
    LP = L * P; LP8 = round_up8(LP);
    qc = min(kQC, kChunkPts / LP8);
    fixed = (kThreadsT % LP8) == 0;
    lp_fixed = tid % LP8; qi_fixed = tid / LP8; qstep = kThreadsT / LP8;
  }
  __device__ __forceinline__ void slot(int tid, int it, int& qi, int& lp) const {
    if (fixed) { qi = qi_fixed + it * qstep; lp = lp_fixed; }
    else { const int idx = tid + it * kThreadsT; qi = idx / LP8; lp = idx - qi * LP8; }
  }
};

// =====================================================================================================
// Forward
// =====================================================================================================
// shared memory: windows [kWinRowsAll x 64 B] | rows [kChunkPts] u32 | weights [2 sides][kChunkPts] u32
constexpr size_t kFwdSmemBytes = static_cast<size_t>(kWinRowsAll) * kRowBytes + static_cast<size_t>(kChunkPts) * 3 * 4;

template <typename T>
__global__ void __launch_bounds__(kThreadsT, 2)
msda_fwd_tiled_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                      const float* __restrict__ loc, const float* __restrict__ attn, T* __restrict__ out,
                      int N, int S, int M, int Lq, int L, int P, int vps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ Geom g;
  __shared__ int s_qid[2][kQC];
  __shared__ int s_ovf;
  __shared__ __align__(8) unsigned long long s_mbar;       // completion of the window fill's bulk copies
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t mbar = smem_u32(&s_mbar);
  uint32_t fill_parity = 0;
  SlotMap sm;
  sm.init(L, P, tid);
  const int LP = sm.LP, LP8 = sm.LP8;
  unsigned char* win_ptr = smem_raw;
  uint32_t* s_rows = reinterpret_cast<uint32_t*>(smem_raw + static_cast<size_t>(kWinRowsAll) * kRowBytes);
  uint32_t* s_wts = s_rows + kChunkPts;              // [side][slot]
  const uint32_t win = smem_u32(win_ptr);

  if (tid == 0) { geom_init(g, shapes, lsi, L, kWinRowsCap); s_ovf = 0; mbar_init(mbar, 1); }
  if (tid < 2 * kRowBytes / 16)                      // the zero rows: written once, never overwritten
    reinterpret_cast<uint4*>(win_ptr + static_cast<size_t>(kZeroRow) * kRowBytes)[tid] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const int tiles = g.tiles_x * g.tiles_y;
  const long long total = static_cast<long long>(N) * tiles * M;
  const float inv_p = 1.0f / static_cast<float>(P);
  const int grp = tid >> 3;                          // 8-lane group
  const int s = (lane >> 2) & 1, c = lane & 3;
  const uint32_t lane_off = static_cast<uint32_t>(s * kRowBytes + c * 16);
  const float nanf_ = __int_as_float(0x7fc00000);

  for (long long work = blockIdx.x; work < total; work += gridDim.x) {
    const int head = static_cast<int>(work % M);
    const int tile = static_cast<int>((work / M) % tiles);
    const int b = static_cast<int>(work / (static_cast<long long>(M) * tiles));
    __syncthreads();                                 // the previous work item no longer reads windows / geometry / qids
    if (warp == 0) geom_tile(g, tile, lane);
    __syncthreads();
    if (kFillBulk) {
      if (tid == 0) mbar_arrive_expect_tx(mbar, g.fill_bytes);
      fence_proxy_async();               // earlier generic-proxy reads of the windows precede the async-proxy writes
      fill_windows_bulk<T>(g, win_ptr, win, mbar, value, b, S, vps, head, warp, lane, kThreadsT / 32);
    } else {
      fill_windows<T>(g, win, value, b, S, vps, head, warp, lane, kThreadsT / 32);
      cp_async_commit();
    }
    const int nq = g.qoff[L];
    const T* img = value + static_cast<size_t>(b) * S * static_cast<size_t>(vps);
    if (tid < kQC) s_qid[0][tid] = tid < sm.qc ? tile_query(g, tid, Lq) : -1;
    LevelC lc = level_consts(g, level_of(min(sm.lp_fixed, LP - 1), inv_p));
    __syncthreads();

    float2 pxy[kPF];
    float pa[kPF];
    // raw sampling locations / weights of a chunk's points, one slot per (thread, it); NaN marks an empty slot
    auto prefetch = [&](const int* qid) {
#pragma unroll
      for (int it = 0; it < kPF; ++it) {
        int qi, lp;
        sm.slot(tid, it, qi, lp);
        const int q = qi < sm.qc ? qid[qi] : -1;
        pxy[it] = make_float2(nanf_, nanf_);
        pa[it] = 0.f;
        if (q >= 0 && lp < LP) {
          const size_t pair = (static_cast<size_t>(b) * Lq + q) * M + head;
          pxy[it] = __ldg(reinterpret_cast<const float2*>(loc) + pair * LP + lp);
          pa[it] = __ldg(attn + pair * LP + lp);
        }
      }
    };
    prefetch(s_qid[0]);

    int buf = 0;
    for (int chunk0 = 0; chunk0 < nq; chunk0 += sm.qc, buf ^= 1) {
      // ---- phase A: lane = point.  window rows + packed 16-bit weights (top | bottom) per side ----
#pragma unroll
      for (int it = 0; it < kPF; ++it) {
        int qi, lp;
        sm.slot(tid, it, qi, lp);
        if (!sm.fixed) lc = level_consts(g, level_of(min(lp, LP - 1), inv_p));
        const PointGeo pg = point_geo(pxy[it].x, pxy[it].y, lc.Hf, lc.Wf);
        const uint32_t rows = resolve_rows<T>(pg, lc, win, img, vps, head, &s_ovf);
        uint32_t wl = 0u, wr = 0u;
        if (pg.inside) {
          const float ah = (1.f - pg.lh) * pa[it], al = pg.lh * pa[it];
          const float hw = 1.f - pg.lw;
          wl = pack_weight_pair<T>(ah * hw, al * hw);
          wr = pack_weight_pair<T>(ah * pg.lw, al * pg.lw);
        }
        const int slot = qi * LP8 + lp;
        if (slot < kChunkPts) {
          s_rows[slot] = rows;
          s_wts[slot] = wl;
          s_wts[kChunkPts + slot] = wr;
        }
      }
      const bool more = chunk0 + sm.qc < nq;
      if (more && tid < kQC) s_qid[buf ^ 1][tid] = tid < sm.qc ? tile_query(g, chunk0 + sm.qc + tid, Lq) : -1;
      cp_async_commit();
      cp_async_wait_all();
      if (kFillBulk && chunk0 == 0) { mbar_wait(mbar, fill_parity); fill_parity ^= 1u; }
      __syncthreads();                               // records, windows, overflow rows and the next chunk's query ids are visible
      if (tid == 0) s_ovf = 0;                       // slots of this chunk are assigned; nobody touches the counter until the next phase A
      if (more) prefetch(s_qid[buf ^ 1]);            // in flight during the gather

      // ---- gather: 8 lanes per query; lanes 0-3 the left pixel of every corner pair, lanes 4-7 the right one.
      //      Groups without a query run the same code on the zero rows: the shuffles below are full-mask ----
      const int ncq = min(sm.qc, nq - chunk0);
      for (int r0 = 0; r0 < ncq; r0 += kGroups) {     // CTA-uniform trip count: every lane runs the full-mask shuffles
        const int qi = (r0 + grp < sm.qc) ? r0 + grp : 0;   // groups past the chunk shadow query 0 (read-only)
        const int q = (r0 + grp < sm.qc) ? s_qid[buf][qi] : -1;
        const bool active = q >= 0;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        const uint32_t* rrow = s_rows + qi * LP8;
        const uint32_t* wrow = s_wts + s * kChunkPts + qi * LP8;
        const size_t pair = (static_cast<size_t>(b) * Lq + (active ? q : 0)) * M + head;
        for (int lp0 = 0; lp0 < LP8; lp0 += 4) {
          const uint4 r4 = *reinterpret_cast<const uint4*>(rrow + lp0);
          const uint4 w4 = *reinterpret_cast<const uint4*>(wrow + lp0);
          const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
          const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
          const bool slow = ((r4.x & 0xFFFFu) == kRowFallback) | ((r4.y & 0xFFFFu) == kRowFallback) |
                            ((r4.z & 0xFFFFu) == kRowFallback) | ((r4.w & 0xFFFFu) == kRowFallback);
          if (!slow) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 ut = lds128(win + (rr[j] & 0xFFFFu) * kRowBytes + lane_off);
              const uint4 ub = lds128(win + (rr[j] >> 16) * kRowBytes + lane_off);
              axpy16_packed<T>(acc, ut, ww[j], false);
              axpy16_packed<T>(acc, ub, ww[j], true);
            }
          } else {                                   // group-uniform and rare: the chunk ran out of overflow slots
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 ut, ub;
              if ((rr[j] & 0xFFFFu) == kRowFallback) {
                const int lp = lp0 + j;
                const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + pair * LP + lp);
                const int l = level_of(lp, inv_p);
                const PointGeo pg = point_geo(xy.x, xy.y, static_cast<float>(g.H[l]), static_cast<float>(g.W[l]));
                fetch_global_pair<T>(img, g.start[l], g.H[l], g.W[l], vps, head, pg.ix, pg.iy, s, c, ut, ub);
              } else {
                ut = lds128(win + (rr[j] & 0xFFFFu) * kRowBytes + lane_off);
                ub = lds128(win + (rr[j] >> 16) * kRowBytes + lane_off);
              }
              axpy16_packed<T>(acc, ut, ww[j], false);
              axpy16_packed<T>(acc, ub, ww[j], true);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 4);
        if (active && s == 0) __stcs(reinterpret_cast<uint4*>(out + pair * kD + c * 8), pack16<T>(acc));
      }
      __syncthreads();                               // records / overflow rows of this chunk are no longer read
    }
  }
}

// =====================================================================================================
// Backward, part 1: grad_sampling_loc, grad_attn_weight (and max|grad_out| for the fp16 scale of part 2)
// =====================================================================================================
// Gather mapping: the 8 lanes of a group work on two points at a time, one lane per corner ROW (64 bytes = the whole
// head slice), so every lane finishes its corner's <value, grad_out> alone and no channel reduction is needed.  A lane
// reads its row as four 16-byte chunks in the rotated order (c0 + k) mod 4, c0 = 2*point + (corner >> 1): the two x
// neighbours of a footprint always differ in 64-byte parity and rows of the same parity get different c0, so the 8
// lanes of a quarter warp hit 8 different bank groups (conflict-free) whatever the window pitch.  After four such
// steps (8 points) a 4x4 transpose inside each quad (4 shuffles) leaves lane (point, corner) with all four corner dots
// of point 2*corner + point, whose three gradients it then writes.
// shared memory: windows | rows [kChunkPts] u32 | lw, lh [kChunkPts] float each | grad_out rows [kQC][64 B]
constexpr size_t kDotsSmemBytes = static_cast<size_t>(kWinRowsAll) * kRowBytes + static_cast<size_t>(kChunkPts) * 3 * 4 +
                                  static_cast<size_t>(kQC) * kRowBytes;

template <typename T>
__global__ void __launch_bounds__(kThreadsT, 2)
msda_bwd_dots_tiled_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                           const float* __restrict__ loc, const float* __restrict__ attn, const T* __restrict__ grad_out,
                           float* __restrict__ grad_loc, float* __restrict__ grad_attn, uint32_t* __restrict__ ctrl,
                           int N, int S, int M, int Lq, int L, int P, int vps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ Geom g;
  __shared__ int s_qid[2][kQC];
  __shared__ int s_ovf;
  __shared__ __align__(8) unsigned long long s_mbar;       // completion of the window fill's bulk copies
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t mbar = smem_u32(&s_mbar);
  uint32_t fill_parity = 0;
  SlotMap sm;
  sm.init(L, P, tid);
  const int LP = sm.LP, LP8 = sm.LP8;
  unsigned char* win_ptr = smem_raw;
  uint32_t* s_rows = reinterpret_cast<uint32_t*>(smem_raw + static_cast<size_t>(kWinRowsAll) * kRowBytes);
  float* s_lw = reinterpret_cast<float*>(s_rows + kChunkPts);
  float* s_lh = s_lw + kChunkPts;
  unsigned char* s_go = reinterpret_cast<unsigned char*>(s_lh + kChunkPts);
  const uint32_t win = smem_u32(win_ptr);
  const uint32_t go_base = smem_u32(s_go);

  if (tid == 0) { geom_init(g, shapes, lsi, L, kWinRowsCap); s_ovf = 0; mbar_init(mbar, 1); }
  if (tid < 2 * kRowBytes / 16)
    reinterpret_cast<uint4*>(win_ptr + static_cast<size_t>(kZeroRow) * kRowBytes)[tid] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const int tiles = g.tiles_x * g.tiles_y;
  const long long total = static_cast<long long>(N) * tiles * M;
  const float inv_p = 1.0f / static_cast<float>(P);
  const int grp = tid >> 3;
  const int pt = (lane >> 2) & 1, cr = lane & 3;            // point of the pair, corner (bit 0: right, bit 1: bottom)
  const int c0 = 2 * pt + (cr >> 1);
  const uint32_t rsh = (cr & 2) ? 16u : 0u;                 // which half of the rows word holds my row pair
  uint32_t coff[4];                                         // byte offsets of my four chunks inside the corner pair
#pragma unroll
  for (int k = 0; k < 4; ++k) coff[k] = static_cast<uint32_t>(((c0 + k) & 3) * 16 + (cr & 1) * kRowBytes);
  const float nanf_ = __int_as_float(0x7fc00000);
  float go_max = 0.f;

  for (long long work = blockIdx.x; work < total; work += gridDim.x) {
    const int head = static_cast<int>(work % M);
    const int tile = static_cast<int>((work / M) % tiles);
    const int b = static_cast<int>(work / (static_cast<long long>(M) * tiles));
    __syncthreads();
    if (warp == 0) geom_tile(g, tile, lane);
    __syncthreads();
    if (kFillBulk) {
      if (tid == 0) mbar_arrive_expect_tx(mbar, g.fill_bytes);
      fence_proxy_async();               // earlier generic-proxy reads of the windows precede the async-proxy writes
      fill_windows_bulk<T>(g, win_ptr, win, mbar, value, b, S, vps, head, warp, lane, kThreadsT / 32);
    } else {
      fill_windows<T>(g, win, value, b, S, vps, head, warp, lane, kThreadsT / 32);
      cp_async_commit();
    }
    const int nq = g.qoff[L];
    const T* img = value + static_cast<size_t>(b) * S * static_cast<size_t>(vps);
    if (tid < kQC) s_qid[0][tid] = tid < sm.qc ? tile_query(g, tid, Lq) : -1;
    LevelC lc = level_consts(g, level_of(min(sm.lp_fixed, LP - 1), inv_p));
    __syncthreads();

    float2 pxy[kPF];
    auto prefetch = [&](const int* qid) {
#pragma unroll
      for (int it = 0; it < kPF; ++it) {
        int qi, lp;
        sm.slot(tid, it, qi, lp);
        const int q = qi < sm.qc ? qid[qi] : -1;
        pxy[it] = make_float2(nanf_, nanf_);
        if (q >= 0 && lp < LP) {
          const size_t pair = (static_cast<size_t>(b) * Lq + q) * M + head;
          pxy[it] = __ldg(reinterpret_cast<const float2*>(loc) + pair * LP + lp);
        }
      }
    };
    prefetch(s_qid[0]);

    int buf = 0;
    for (int chunk0 = 0; chunk0 < nq; chunk0 += sm.qc, buf ^= 1) {
      // ---- phase A: lane = point.  window rows + fractions; this chunk's grad_out rows ----
#pragma unroll
      for (int it = 0; it < kPF; ++it) {
        int qi, lp;
        sm.slot(tid, it, qi, lp);
        if (!sm.fixed) lc = level_consts(g, level_of(min(lp, LP - 1), inv_p));
        const PointGeo pg = point_geo(pxy[it].x, pxy[it].y, lc.Hf, lc.Wf);
        const uint32_t rows = resolve_rows<T>(pg, lc, win, img, vps, head, &s_ovf);
        const int slot = qi * LP8 + lp;
        if (slot < kChunkPts) {
          s_rows[slot] = rows;
          s_lw[slot] = pg.inside ? pg.lw : nanf_;      // NaN: the point fails the gate, its gradients are zero
          s_lh[slot] = pg.lh;
        }
      }
      if (tid < kQC * 4) {
        const int qi = tid >> 2, ch = tid & 3;
        const int q = qi < sm.qc ? s_qid[buf][qi] : -1;
        const size_t pair = (static_cast<size_t>(b) * Lq + (q >= 0 ? q : 0)) * M + head;
        cp_async16_zfill(go_base + static_cast<uint32_t>(qi * kRowBytes + ch * 16), grad_out + pair * kD + ch * 8, q >= 0);
      }
      const bool more = chunk0 + sm.qc < nq;
      if (more && tid < kQC) s_qid[buf ^ 1][tid] = tid < sm.qc ? tile_query(g, chunk0 + sm.qc + tid, Lq) : -1;
      cp_async_commit();
      cp_async_wait_all();
      if (kFillBulk && chunk0 == 0) { mbar_wait(mbar, fill_parity); fill_parity ^= 1u; }
      __syncthreads();
      if (tid == 0) s_ovf = 0;
      if (more) prefetch(s_qid[buf ^ 1]);
      if (tid < kQC * 4) {                                 // max|grad_out| over the rows this chunk staged
        float f[8];
        unpack16<T>(*reinterpret_cast<const uint4*>(s_go + tid * 16), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) go_max = fmaxf(go_max, fabsf(f[i]));
      }

      const int ncq = min(sm.qc, nq - chunk0);
      for (int r0 = 0; r0 < ncq; r0 += kGroups) {
        const int qi = (r0 + grp < sm.qc) ? r0 + grp : 0;
        const int q = (r0 + grp < sm.qc) ? s_qid[buf][qi] : -1;
        const bool active = q >= 0;
        const size_t pair = (static_cast<size_t>(b) * Lq + (active ? q : 0)) * M + head;
        uint4 gor[4];                                      // the pair's grad_out row, chunks in my rotated order
#pragma unroll
        for (int k = 0; k < 4; ++k) gor[k] = lds128(go_base + static_cast<uint32_t>(qi * kRowBytes + ((c0 + k) & 3) * 16));
        const uint32_t* rrow = s_rows + qi * LP8;
        for (int lp0 = 0; lp0 < LP8; lp0 += 8) {
          const int lp = lp0 + 2 * cr + pt;               // the point this lane owns after the transpose
          float a_own = 0.f;
          if (active && lp < LP) a_own = __ldg(attn + pair * LP + lp);
          const uint4 ra = *reinterpret_cast<const uint4*>(rrow + lp0);
          const uint4 rb = *reinterpret_cast<const uint4*>(rrow + lp0 + 4);
          const uint32_t rw[4] = {pt ? ra.y : ra.x, pt ? ra.w : ra.z, pt ? rb.y : rb.x, pt ? rb.w : rb.z};
          float d[4];                                      // my corner's dot for points lp0 + 2i + pt
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float acc = 0.f;
            if ((rw[i] & 0xFFFFu) != kRowFallback) {
              const uint32_t addr = win + ((rw[i] >> rsh) & 0xFFFFu) * kRowBytes;
#pragma unroll
              for (int k = 0; k < 4; ++k) acc = dot16<T>(lds128(addr + coff[k]), gor[k], acc);
            } else {                                       // rare: no overflow slot was left for this point
              const int lpi = lp0 + 2 * i + pt;
              const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + pair * LP + lpi);
              const int l = level_of(lpi, inv_p);
              const PointGeo pg = point_geo(xy.x, xy.y, static_cast<float>(g.H[l]), static_cast<float>(g.W[l]));
              const int x = pg.ix + (cr & 1), y = pg.iy + (cr >> 1);
              if (static_cast<unsigned>(x) < static_cast<unsigned>(g.W[l]) && static_cast<unsigned>(y) < static_cast<unsigned>(g.H[l])) {
                const char* row = reinterpret_cast<const char*>(img) + static_cast<size_t>(head) * kRowBytes +
                                  (static_cast<size_t>(g.start[l]) + static_cast<size_t>(y) * g.W[l] + x) * static_cast<size_t>(vps) * sizeof(T);
#pragma unroll
                for (int k = 0; k < 4; ++k) acc = dot16<T>(ldg16(row + ((c0 + k) & 3) * 16), gor[k], acc);
              }
            }
            d[i] = acc;
          }
          // 4x4 transpose inside the quad: lane `cr` ends with the four corner dots of point i = cr
          const bool oddc = (cr & 1) != 0, bot = (cr & 2) != 0;
          const float rA0 = __shfl_xor_sync(0xffffffffu, oddc ? d[0] : d[1], 1);
          const float rA1 = __shfl_xor_sync(0xffffffffu, oddc ? d[2] : d[3], 1);
          const float kA0 = oddc ? d[1] : d[0], kA1 = oddc ? d[3] : d[2];
          const float a_lo = oddc ? rA0 : kA0, a_hi = oddc ? kA0 : rA0;       // point (cr & 1), my corner pair (x, x+1)
          const float b_lo = oddc ? rA1 : kA1, b_hi = oddc ? kA1 : rA1;       // point 2 + (cr & 1)
          const float rB_lo = __shfl_xor_sync(0xffffffffu, bot ? a_lo : b_lo, 2);
          const float rB_hi = __shfl_xor_sync(0xffffffffu, bot ? a_hi : b_hi, 2);
          const float kB_lo = bot ? b_lo : a_lo, kB_hi = bot ? b_hi : a_hi;
          const float d00 = bot ? rB_lo : kB_lo, d01 = bot ? rB_hi : kB_hi;
          const float d10 = bot ? kB_lo : rB_lo, d11 = bot ? kB_hi : rB_hi;
          if (active && lp < LP) {
            const int slot = qi * LP8 + lp;
            float lw = s_lw[slot], lh = s_lh[slot];
            float a = a_own;
            if (!(lw == lw)) { lw = 0.f; lh = 0.f; a = 0.f; }   // gate failed: upstream skips the point
            const int l = level_of(lp, inv_p);
            const float hw = 1.f - lw, hh = 1.f - lh;
            const float ga = hh * (hw * d00 + lw * d01) + lh * (hw * d10 + lw * d11);
            const float gx = static_cast<float>(g.W[l]) * a * (hh * (d01 - d00) + lh * (d11 - d10));
            const float gy = static_cast<float>(g.H[l]) * a * (hw * (d10 - d00) + lw * (d11 - d01));
            __stcs(reinterpret_cast<float2*>(grad_loc) + pair * LP + lp, make_float2(gx, gy));
            __stcs(grad_attn + pair * LP + lp, ga);
          }
        }
      }
      __syncthreads();
    }
  }
  if (ctrl != nullptr) {
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) go_max = fmaxf(go_max, __shfl_xor_sync(0xffffffffu, go_max, sft));
    if (lane == 0 && go_max > 0.f) atomicMax(ctrl, __float_as_uint(go_max));
  }
}

// =====================================================================================================
// Backward, part 2: grad_value by counting sort + segmented sums (no read-modify-write in shared memory)
// =====================================================================================================
constexpr int kScThreads = 256;
constexpr int kScLG = 2;                          // levels sorted together per round
constexpr int kScIters = 6;                       // point slots per thread per round
constexpr int kScPoints = kScThreads * kScIters;  // point slots per round
constexpr int kScEntries = 2 * kScPoints;         // pair entries per round (top and bottom row of every point)
constexpr int kScQueries = 192;                   // queries per round (8-bit local index)
constexpr int kScMaxRun = 64;                     // adds per fp16 register accumulator before it is flushed
constexpr int kScClasses = 64;                    // bins are ordered by min(count, 63)

// shared memory: go2 [kScQueries][2][64 B] | entries [kScEntries] uint2 | cnt [kWinRowsCap + 1] u32 | dst [kWinRowsCap] u32 |
//                order [kWinRowsCap] u16 | qid [kScQueries] | hist [kScClasses] | hist_start [kScClasses]
constexpr size_t kScSmemBytes = static_cast<size_t>(kScQueries) * 2 * kRowBytes + static_cast<size_t>(kScEntries) * 8 +
                                static_cast<size_t>(kWinRowsCap + 1) * 4 + static_cast<size_t>(kWinRowsCap) * 4 +
                                static_cast<size_t>(kWinRowsCap) * 2 + static_cast<size_t>(kScQueries) * 4 +
                                static_cast<size_t>(kScClasses) * 8;

__device__ __forceinline__ uint32_t hfma2_u32(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// per-level constants of the count phase
struct ScLevel {
  int l, H, W, wx0, wy0, wdx, wdy, lbase;
  uint32_t acc_row0;
  float Hf, Wf;
};
__device__ __forceinline__ ScLevel sc_level(const Geom& g, const LevelMeta& meta, const int* lst, int lv, int tile) {
  ScLevel s;
  s.l = lst[lv];
  s.H = g.H[s.l]; s.W = g.W[s.l];
  s.Hf = static_cast<float>(s.H); s.Wf = static_cast<float>(s.W);
  s.wx0 = g.wx0[s.l]; s.wy0 = g.wy0[s.l]; s.wdx = g.wdx[s.l]; s.wdy = g.wdy[s.l];
  s.lbase = 0;
  for (int k = 0; k < lv; ++k) s.lbase += g.wdx[lst[k]] * g.wdy[lst[k]];
  const int K = meta.accK[s.l];
  s.acc_row0 = static_cast<uint32_t>(meta.accBase[s.l]) + static_cast<uint32_t>(K > 1 ? tile % K : 0) * static_cast<uint32_t>(s.H * s.W);
  return s;
}
// a footprint row that leaves its window: both 64-byte rows straight into the accumulator (rare)
static __device__ __noinline__ void sc_slow_pair(__half* dst, size_t pix_elems, const unsigned char* go_row, uint32_t wbits, bool do_l, bool do_r) {
  const uint32_t wl2 = __byte_perm(wbits, 0, 0x1010), wr2 = __byte_perm(wbits, 0, 0x3232);
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const uint4 u = *reinterpret_cast<const uint4*>(go_row + ch * 16);
    if (do_l)
      red_add_16bit_x8<__half>(dst + ch * 8, make_uint4(hfma2_u32(wl2, u.x, 0u), hfma2_u32(wl2, u.y, 0u),
                                                       hfma2_u32(wl2, u.z, 0u), hfma2_u32(wl2, u.w, 0u)));
    if (do_r)
      red_add_16bit_x8<__half>(dst + pix_elems + ch * 8, make_uint4(hfma2_u32(wr2, u.x, 0u), hfma2_u32(wr2, u.y, 0u),
                                                                   hfma2_u32(wr2, u.z, 0u), hfma2_u32(wr2, u.w, 0u)));
  }
}

// Unit of the sort: an x-adjacent corner PAIR (left pixel x0, right pixel x0 + 1 of one footprint row) -- both corners
// scale the same grad_out row, so a pair entry (bin = window cell of the left pixel, local query, two fp16 weights) halves
// the ranks, the places and the grad_out row reads of a per-corner sort.  Two levels are sorted together per round.
// After the exclusive scan the non-empty bins are ordered by their entry count (a second, 64-class counting sort), and
// the 4-lane groups take bins in that order: the 8 groups of a warp then work on bins of (nearly) equal length, every
// group sums its bin's left and right pixel in fp16 registers (HFMA2; grad_out staged as fp16 already multiplied by
// the accumulator's power-of-two scale, at both 64-byte parities) and the whole warp reaches the flush -- two packed fp16
// reductions per bin -- together.  A register accumulator takes at most kScMaxRun products between flushes.
template <typename T>
__global__ void __launch_bounds__(kScThreads, 3)
msda_bwd_scatter_tiled_kernel(const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                              const float* __restrict__ loc, const float* __restrict__ attn, const T* __restrict__ grad_out,
                              __half* __restrict__ gv16, const uint32_t* __restrict__ ctrl,
                              int N, int S, int M, int Lq, int L, int P, int depth) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ Geom g;
  __shared__ LevelMeta meta;
  __shared__ uint32_t s_warp_tot[kScThreads / 32];
  __shared__ uint32_t s_total, s_nbins;
  __shared__ int s_lv[kMaxL], s_nlv;         // the levels this kernel handles: all, or meta.sortMask of a hybrid backward
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* s_go = smem_raw;
  uint2* s_ent = reinterpret_cast<uint2*>(smem_raw + static_cast<size_t>(kScQueries) * 2 * kRowBytes);
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_ent + kScEntries);
  uint32_t* s_dst = s_cnt + kWinRowsCap + 1;
  uint16_t* s_order = reinterpret_cast<uint16_t*>(s_dst + kWinRowsCap);
  int* s_qid = reinterpret_cast<int*>(s_order + kWinRowsCap);
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_qid + kScQueries);
  uint32_t* s_hstart = s_hist + kScClasses;
  const uint32_t go_base = smem_u32(s_go);

  load_level_meta(meta, shapes, lsi, L);
  build_accum_layout(meta, L, Lq, P, depth, false);
  if (tid == 0) {
    geom_init(g, shapes, lsi, L, kWinRowsCap);
    int n = 0;
    for (int l = 0; l < L; ++l)
      if (accum_split_of(depth) == 0 || ((meta.sortMask >> l) & 1u)) s_lv[n++] = l;
    s_nlv = n;
  }
  __syncthreads();
  const int nlv = s_nlv;
  const float gv_scale = f16_accum_scale(ctrl, Lq);
  const int tiles = g.tiles_x * g.tiles_y;
  const long long total = static_cast<long long>(N) * tiles * M;
  const int q_round = max(1, min(kScQueries, kScPoints / (P * kScLG)));
  const int grp = tid >> 2, c = tid & 3;                    // 4-lane group: one bin at a time; lane = 8 channels
  const uint32_t go_lane = static_cast<uint32_t>(((grp & 1) * kRowBytes) + c * 16);   // copy of the row in "my" half of the banks
  const size_t pix_elems = static_cast<size_t>(M) * kD;
  const int LP = L * P;
  const float nanf_ = __int_as_float(0x7fc00000);

  for (long long work = blockIdx.x; work < total; work += gridDim.x) {
    const int head = static_cast<int>(work % M);
    const int tile = static_cast<int>((work / M) % tiles);
    const int b = static_cast<int>(work / (static_cast<long long>(M) * tiles));
    __syncthreads();
    if (warp == 0) geom_tile(g, tile, lane);
    __syncthreads();
    const int nq = g.qoff[L];
    __half* acc_img = gv16 + (static_cast<size_t>(b) * meta.accStride * M + head) * kD + c * 8;

    for (int q0 = 0; q0 < nq; q0 += q_round) {
      const int nqr = min(q_round, nq - q0);
      __syncthreads();                      // previous round no longer reads go rows / query ids
      for (int i = tid; i < nqr; i += kScThreads) s_qid[i] = tile_query(g, q0 + i, Lq);
      __syncthreads();
      // ---- stage the round's grad_out rows as scaled fp16, twice: row q at both 64-byte halves of a 128-byte line, so
      //      that the two 4-lane groups of a quarter warp always read from disjoint banks ----
      for (int i = tid; i < nqr * 4; i += kScThreads) {
        const int qi = i >> 2, ch = i & 3;
        const int q = s_qid[qi];
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (q >= 0) {
          const size_t pair = (static_cast<size_t>(b) * Lq + q) * M + head;
          float f[8];
          unpack16<T>(ldg16(grad_out + pair * kD + ch * 8), f);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] *= gv_scale;
          u = pack16<__half>(f);
        }
        *reinterpret_cast<uint4*>(s_go + static_cast<size_t>(qi) * 2 * kRowBytes + ch * 16) = u;
        *reinterpret_cast<uint4*>(s_go + static_cast<size_t>(qi) * 2 * kRowBytes + kRowBytes + ch * 16) = u;
      }

      for (int i0 = 0; i0 < nlv; i0 += kScLG) {
        const int lg = min(kScLG, nlv - i0);                // levels of this round
        const int* lst = s_lv + i0;
        const int ppq = P * lg;                             // point slots per query
        const int npts = nqr * ppq;
        // first window row (bin) of every level of the round, and the round's bin count
        int nrows = 0;
        for (int lv = 0; lv < lg; ++lv) nrows += g.wdx[lst[lv]] * g.wdy[lst[lv]];
        __syncthreads();                    // previous round's entries / counters / order are no longer read; go rows are staged
        for (int i = tid; i <= nrows; i += kScThreads) s_cnt[i] = 0u;
        if (tid < kScClasses) s_hist[tid] = 0u;
        __syncthreads();
        // ---- count: thread = point slot; each footprint row inside the window takes a rank in its pair bin ----
        uint32_t keyrank[kScIters][2];      // key | rank << 16; 0xFFFFFFFF = no entry
        uint32_t payload[kScIters][2];      // fp16 weight of the left | right pixel
        uint32_t qis[kScIters];
        // slot = qi * ppq + lv * P + p.  When ppq divides the CTA size a thread's (lv, p) never changes: decode once and
        // keep the level's constants in registers
        const bool fixed = (kScThreads % ppq) == 0;
        const int remF = tid % ppq, qiF = tid / ppq, qstep = kScThreads / ppq;
        ScLevel sl = sc_level(g, meta, lst, static_cast<int>(static_cast<uint32_t>(remF) / static_cast<uint32_t>(P)), tile);
        const int pF = remF - (remF / P) * P;
#pragma unroll
        for (int it = 0; it < kScIters; ++it) {
          keyrank[it][0] = keyrank[it][1] = 0xFFFFFFFFu;
          qis[it] = 0u;
          int qi, p;
          if (fixed) {
            qi = qiF + it * qstep; p = pF;
          } else {
            const int slot = tid + it * kScThreads;
            qi = static_cast<int>(static_cast<uint32_t>(slot) / static_cast<uint32_t>(ppq));
            const int rem = slot - qi * ppq;
            const int lv = static_cast<int>(static_cast<uint32_t>(rem) / static_cast<uint32_t>(P));
            p = rem - lv * P;
            sl = sc_level(g, meta, lst, lv, tile);
          }
          if (qi >= nqr) continue;
          const int q = s_qid[qi];
          if (q < 0) continue;
          qis[it] = static_cast<uint32_t>(qi);
          const size_t pair = (static_cast<size_t>(b) * Lq + q) * M + head;
          const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + pair * LP + sl.l * P + p);
          const float a = __ldg(attn + pair * LP + sl.l * P + p);
          const PointGeo pg = point_geo(xy.x, xy.y, sl.Hf, sl.Wf);
          if (!pg.inside) continue;
          const int wxr = pg.ix - sl.wx0, wyr0 = pg.iy - sl.wy0;
          const bool lin = static_cast<unsigned>(pg.ix) < static_cast<unsigned>(sl.W);
          const bool rin = static_cast<unsigned>(pg.ix + 1) < static_cast<unsigned>(sl.W);
          const float wx_l = lin ? (1.f - pg.lw) : 0.f, wx_r = rin ? pg.lw : 0.f;
          const float wy[2] = {(1.f - pg.lh) * a, pg.lh * a};
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int y = pg.iy + r;
            if (static_cast<unsigned>(y) >= static_cast<unsigned>(sl.H)) continue;       // zero padding
            const float wl = wy[r] * wx_l, wr = wy[r] * wx_r;
            if (wl == 0.f && wr == 0.f) continue;
            const __half2 w2 = __floats2half2_rn(wl, wr);
            const uint32_t wbits = *reinterpret_cast<const uint32_t*>(&w2);
            const int wyr = wyr0 + r;
            if (wxr >= 0 && wxr <= sl.wdx - 2 && static_cast<unsigned>(wyr) < static_cast<unsigned>(sl.wdy)) {
              const uint32_t key = static_cast<uint32_t>(sl.lbase + wyr * sl.wdx + wxr);
              const uint32_t rank = atomicAdd(&s_cnt[key], 1u);
              keyrank[it][r] = key | (rank << 16);
              payload[it][r] = wbits;
            } else {
              // slow path: the two 64-byte rows of this pair, reduced straight into the accumulator
              sc_slow_pair(acc_img - c * 8 + (static_cast<size_t>(sl.acc_row0) + static_cast<size_t>(y) * sl.W + pg.ix) * pix_elems,
                           pix_elems, s_go + static_cast<size_t>(qi) * 2 * kRowBytes, wbits, lin && wl != 0.f, rin && wr != 0.f);
            }
          }
        }
        __syncthreads();
        // ---- exclusive scan of the bin counts (in place), destination of every bin, histogram of the bin lengths ----
        {
          const int per = (nrows + kScThreads - 1) / kScThreads;       // <= 6
          const int r0 = tid * per, r1 = min(nrows, r0 + per);
          uint32_t sum = 0u;
          for (int r = r0; r < r1; ++r) sum += s_cnt[r];
          uint32_t incl = sum;
#pragma unroll
          for (int sft = 1; sft < 32; sft <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, sft);
            if (lane >= sft) incl += t;
          }
          if (lane == 31) s_warp_tot[warp] = incl;
          __syncthreads();
          uint32_t woff = 0u;
          for (int w2 = 0; w2 < warp; ++w2) woff += s_warp_tot[w2];
          uint32_t run = woff + incl - sum;
          if (r0 < r1) {
            // (level, window cell) of bin r0, then walk
            int lv = 0, rr = r0;
            while (lv + 1 < lg && rr >= g.wdx[lst[lv]] * g.wdy[lst[lv]]) { rr -= g.wdx[lst[lv]] * g.wdy[lst[lv]]; ++lv; }
            int l = lst[lv];
            int wdx = g.wdx[l];
            int wyr = static_cast<int>(static_cast<uint32_t>(rr) / static_cast<uint32_t>(wdx)), wxr = rr - wyr * wdx;
            for (int r = r0; r < r1; ++r) {
              const uint32_t cnt = s_cnt[r];
              s_cnt[r] = run;
              run += cnt;
              if (cnt != 0u) {
                const int H = g.H[l], W = g.W[l];
                const int K = meta.accK[l];
                const uint32_t acc_row0 = static_cast<uint32_t>(meta.accBase[l]) + static_cast<uint32_t>(K > 1 ? tile % K : 0) * static_cast<uint32_t>(H * W);
                const int x = g.wx0[l] + wxr, y = g.wy0[l] + wyr;
                const uint32_t lbit = static_cast<unsigned>(x) < static_cast<unsigned>(W) ? 0x80000000u : 0u;
                const uint32_t rbit = static_cast<unsigned>(x + 1) < static_cast<unsigned>(W) ? 0x40000000u : 0u;
                // row of the left pixel, or of the right one when the left lies outside the level (x == -1)
                s_dst[r] = ((acc_row0 + static_cast<uint32_t>(y * W + (lbit ? x : x + 1))) & 0x3FFFFFFFu) | lbit | rbit;
                atomicAdd(&s_hist[min(cnt, static_cast<uint32_t>(kScClasses - 1))], 1u);
              }
              if (++wxr == wdx) {
                wxr = 0;
                if (++wyr == g.wdy[l]) { wyr = 0; ++lv; l = lst[lv < lg ? lv : lg - 1]; wdx = g.wdx[l]; }
              }
            }
          }
          if (tid == kScThreads - 1) { s_total = woff + incl; s_cnt[nrows] = woff + incl; }
        }
        __syncthreads();
        // ---- order of the non-empty bins: longest first (descending 64-class counting sort) ----
        if (warp == 0) {
          const uint32_t ha = s_hist[kScClasses - 1 - 2 * lane], hb = s_hist[kScClasses - 2 - 2 * lane];
          uint32_t incl = ha + hb;
#pragma unroll
          for (int sft = 1; sft < 32; sft <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, sft);
            if (lane >= sft) incl += t;
          }
          const uint32_t excl = incl - (ha + hb);
          s_hstart[kScClasses - 1 - 2 * lane] = excl;
          s_hstart[kScClasses - 2 - 2 * lane] = excl + ha;
          s_hist[kScClasses - 1 - 2 * lane] = 0u;       // reused as fill counters
          s_hist[kScClasses - 2 - 2 * lane] = 0u;
          if (lane == 31) s_nbins = incl;
        }
        __syncthreads();
        // ---- place the entries and the bins ----
#pragma unroll
        for (int it = 0; it < kScIters; ++it) {
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t kr = keyrank[it][r];
            if (kr != 0xFFFFFFFFu) {
              const uint32_t key = kr & 0xFFFFu;
              s_ent[s_cnt[key] + (kr >> 16)] = make_uint2(key | (qis[it] << 16), payload[it][r]);
            }
          }
        }
        for (int r = tid; r < nrows; r += kScThreads) {
          const uint32_t cnt = s_cnt[r + 1] - s_cnt[r];
          if (cnt != 0u) {
            const uint32_t cls = min(cnt, static_cast<uint32_t>(kScClasses - 1));
            s_order[s_hstart[cls] + atomicAdd(&s_hist[cls], 1u)] = static_cast<uint16_t>(r);
          }
        }
        __syncthreads();
        // ---- sums: one bin per 4-lane group at a time, bins in order of decreasing length ----
        {
          const int nbins = static_cast<int>(s_nbins);
          const int ngroups = kScThreads / 4;
          for (int i = grp; i < nbins; i += ngroups) {
            const uint32_t r = s_order[i];
            const uint32_t e0 = s_cnt[r], e1 = s_cnt[r + 1];
            const uint32_t dst = s_dst[r];
            __half* row = acc_img + static_cast<size_t>(dst & 0x3FFFFFFFu) * pix_elems;
            for (uint32_t ea = e0; ea < e1; ea += kScMaxRun) {
              const uint32_t eb = min(e1, ea + static_cast<uint32_t>(kScMaxRun));
              uint4 aL = make_uint4(0u, 0u, 0u, 0u), aR = aL;
              for (uint32_t e = ea; e < eb; ++e) {
                const uint2 ent = s_ent[e];
                const uint4 u = lds128(go_base + (ent.x >> 16) * (2 * kRowBytes) + go_lane);
                const uint32_t wl2 = __byte_perm(ent.y, 0, 0x1010), wr2 = __byte_perm(ent.y, 0, 0x3232);
                aL.x = hfma2_u32(wl2, u.x, aL.x); aL.y = hfma2_u32(wl2, u.y, aL.y);
                aL.z = hfma2_u32(wl2, u.z, aL.z); aL.w = hfma2_u32(wl2, u.w, aL.w);
                aR.x = hfma2_u32(wr2, u.x, aR.x); aR.y = hfma2_u32(wr2, u.y, aR.y);
                aR.z = hfma2_u32(wr2, u.z, aR.z); aR.w = hfma2_u32(wr2, u.w, aR.w);
              }
              if (dst & 0x80000000u) red_add_16bit_x8<__half>(row, aL);
              if (dst & 0x40000000u) red_add_16bit_x8<__half>((dst & 0x80000000u) ? row + pix_elems : row, aR);
            }
          }
        }
      }
    }
  }
}

}  // namespace tiled

// Stages (a)+(b) as ONE persistent dataflow kernel: dense occupancy grids and/or range images for a batch of ragged frames.
//
// Reference behaviour (bit-exact): voxel_filter data/data_preprocessing.py:172-228 + densify muvo/data/dataset.py:317-327,
// do_range_projection muvo/utils/geometry_utils.py:175-220.
//
// Why one kernel: the multi-launch path (points.cu) runs an issue/atomic-bound point pass and then two DRAM-bound emit
// passes one after the other, and its per-frame winner tables (F x 20 MB of address space) fall out of L2 in between.
// Here the work is cut into units -- P: 1792 points of one frame, ER: 1792 pixels of one range image, ED: 64 KiB of one
// dense grid -- that persistent CTAs draw from a global ticket counter.  Tickets are ordered so that every dependency
// points to a LOWER ticket (P(f) ... L frames of other work ... E(f) ... P(f + R) reuses the table slot of frame f), which
// makes spinning on the per-frame completion counters deadlock free for any co-resident grid.  Consequences:
//   * the winner tables are a ring of R <= 16 frame slots (~1.2 MB touched per slot): they never leave L2, the atomics of
//     one frame run next to the streaming stores of another, and there is a single launch tail;
//   * each CTA = 7 consumer warps + 1 producer warp: the producer draws tickets, waits for dependencies and stages the
//     point tile of the next unit with cp.async.bulk (TMA) into a two-stage shared-memory ring (mbarrier full/empty);
//   * per point the hot loop is f32 only: voxel id from floor(x / res) (exact for a power-of-two res), an APPROXIMATE key
//     class for both arg-mins, one 64-bit atomicMax per table.  Classes are coarse enough that the exact float64 order of
//     two candidates can only disagree with the class order when the classes differ by at most 1; every point whose
//     atomicMax met such a neighbour is queued, and the last CTA to finish a frame's points settles the queue with the
//     exact float64 keys of the reference (tie protocol: lowest index among exactly equal keys).
// Eligibility (everything else runs the multi-launch path of points.cu): float32 points, 16-byte aligned arrays, dense
// and/or range outputs (no sorted sparse list), power-of-two resolution with offset / res integral, sensor position exact
// in float32, H*W % 4 == 0.
#include <string.h>
#include "common.cuh"
#include "points_dev.cuh"

namespace muvo {
namespace {

constexpr int kMegaThreads = 256;
constexpr int kConsumerWarps = 7;
constexpr int kConsumers = kConsumerWarps * 32;            // 224
constexpr int kPPT = 8;                                    // points per consumer thread per P unit
#ifndef MUVO_MEGA_BATCH
#define MUVO_MEGA_BATCH 2
#endif
constexpr int kBatch = MUVO_MEGA_BATCH;                    // points per thread whose claims are in flight together
#ifndef MUVO_MEGA_MINB
#define MUVO_MEGA_MINB 4
#endif
constexpr int kUnit = kConsumers * kPPT;                   // 1792 points
constexpr int kTilePts = kUnit + 16;                       // the copy starts at a multiple of 16 points: up to 15 leading extras
constexpr int kTileXyzBytes = kTilePts * 12;
constexpr int kTileSemBytes = kTilePts;
constexpr int kStageBytes = (kTileXyzBytes + kTileSemBytes + 127) / 128 * 128;
constexpr int kPixUnit = kConsumers * 8;                   // pixels per ER unit (two groups of 4 per thread)
constexpr int kEdWordsPerBlock = kConsumerWarps * 32 * kEmitWords;   // bitmap words per dense block (448 -> 14 KiB of output)
constexpr int kEdBlocksPerUnit = 4;
constexpr int kMaxRing = 64, kDefaultRing = 16;

enum UnitType { U_POINTS = 0, U_RANGE = 1, U_DENSE = 2, U_DONE = 3 };

struct UnitDesc {
  int type, f, j, slot;
  int n, lead, packl, n_units;      // P: points in the unit, leading extras in the tile, label-carrying words, P units of the frame
  int64_t fbeg, a;                  // first point of the frame, first point of the unit
};

struct MegaEmitSmem {
  uint8_t tile[kConsumerWarps][kEmitWords * 1024];
  uint16_t list[kConsumerWarps][kEmitList];
};
static_assert(sizeof(MegaEmitSmem) <= kStageBytes, "dense-emit scratch must fit a stage buffer");

// f32 description of a "regular" grid
struct GridF {
  float inv_res, inv_res2_cls;     // 1/res ; class scale relative to res^2 (see vox_cls_exact)
  float offq[3];
  uint32_t dimu[3];
  uint32_t tiny_m1[3];             // bits(tiny_k) - 1: a coordinate in (-tiny_k, 0) takes the float64 path
  uint32_t sx, sy;                 // bit = ix*sx + iy*sy + iz
  int road;
};

struct SyncWs {
  uint32_t* p_claim;   // [F]   P units of the frame handed out so far
  uint32_t* e_claim;   // [F]   E units handed out so far
  uint32_t* p_done;    // [F]   warps that finished their share of a P unit
  uint32_t* ready;     // [F]   1 = all points are in the tables and the queue is settled
  uint32_t* e_done;    // [F]   warps that finished their share of an E unit
  uint32_t* qn;        // [F]   rare-path queue length of the frame
};

struct MegaArgs {
  const float* xyz; const uint8_t* sem; const int64_t* off; int F; int64_t P;
  GridDev g; GridF gf; RangeDev r; int64_t HW;
  int do_vox, do_range, layout, filter;
  int R, nER, nED;
  uint32_t* bitmap; u64* vtab; u64* pixtab; uint2* queue;
  SyncWs s;
  const uint8_t* remap; uint8_t* dense; int64_t* n_occ;
  float* depth_out; float* xyz_out; uint8_t* sem_out;
  int64_t* diag;
  const double* edges;     // [2 (W + 1)] cos, sin of the column-edge yaw angles, then [H + 1] sin of the row-edge pitch angles
};

constexpr float kVoxClsScale = 262144.0f;      // 2^18 classes per res^2: f32 key error (< 0.2 class) keeps exact order within +-1 class
constexpr uint32_t kVoxTopMax = 0x003fffffu;
constexpr uint32_t kQSlow = 0xffffffffu;       // queue entry .y: the float64 formula decides the cell (else: 1-based index of the in-band holder met)
constexpr uint32_t kQVoxel = 0x80000000u;      // queue entry .x: frame-relative point index, this bit set for a voxel event

// ---------------------------------------------------------------- small PTX helpers
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u64 ld_cg_u64(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_cg_u32x2(const uint2* p) {
  uint2 v;
  asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void bulk_g2s_plain(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Per-CTA cycle counters (SM clock), read back by muvo_debug_mega_stats: [0] consumers waiting for a unit, [1] P units,
// [2] ER units, [3] ED units, [4] queue settling, [5] producer waiting for a free stage, [6] producer waiting for a frame's
// points (ER/ED dependency), [7] producer waiting for a table slot (P dependency), [8..10] units of each type, [11] total.
constexpr int kStatCtas = 1024, kStatSlots = 20;
__device__ unsigned long long g_mega_stats[kStatCtas][kStatSlots];
__device__ __forceinline__ void stat_add(int slot, long long dt) {
  if (blockIdx.x < kStatCtas) g_mega_stats[blockIdx.x][slot] += (unsigned long long)dt;
}

// The winner tables (a ring of frame slots, re-used every R frames) are accessed with an L2 evict_last policy, the output
// streams with evict_first (st.global.cs): the tables stay resident while 0.3 GB of results per step flow past them.
// (l2_evict_last_policy: points_dev.cuh)
// atomicMax on table[cell] unless cell == kCellNone (0xffffffff); returns the old word, or `none` when nothing was done.
// Predicated instead of branched: no divergence bookkeeping around the (very common) claim.
__device__ __forceinline__ u64 atom_max_if(u64* table, uint32_t cell, u64 word, u64 none, uint64_t pol) {
  u64 old = none;
  asm volatile("{\n.reg .pred p;\n.reg .u64 ad;\nsetp.ne.u32 p, %2, 0xffffffff;\nmad.wide.u32 ad, %2, 8, %1;\n"
               "@p atom.global.max.L2::cache_hint.u64 %0, [ad], %3, %4;\n}"
               : "+l"(old) : "l"(table), "r"(cell), "l"(word), "l"(pol) : "memory");
  return old;
}
__device__ __forceinline__ void red_or_if(uint32_t* bitmap, uint32_t bit, bool go, uint64_t pol) {
  asm volatile("{\n.reg .pred p;\n.reg .u64 ad;\n.reg .u32 w, m;\nsetp.ne.u32 p, %2, 0;\nshr.u32 w, %1, 5;\nmad.wide.u32 ad, w, 4, %0;\n"
               "and.b32 w, %1, 31;\nshl.b32 m, 1, w;\n@p red.global.or.L2::cache_hint.b32 [ad], m, %3;\n}"
               :: "l"(bitmap), "r"(bit), "r"((uint32_t)go), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_zero16_last(void* p, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.u64 [%0], {%1,%1}, %2;" ::"l"(p), "l"(0ull), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_zero32_last(void* p, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.u64 [%0], {%1,%1,%1,%1}, %2;" ::"l"(p), "l"(0ull), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_u32_last(uint32_t* p, uint32_t v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t ld_u32_last(const uint32_t* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ u64 ld_u64_last(const u64* p, uint64_t pol) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ ulonglong2 ld_u64x2_last(const u64* p, uint64_t pol) {
  ulonglong2 v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(pol) : "memory");
  return v;
}

// ---------------------------------------------------------------- per-point f32 arithmetic
// voxel of a point on a regular grid (res = 2^k, offset / res integral): q = p / res is exact in f32, so
// floor(q) + offset/res is the reference's floor((p + offset) / res) whenever the float64 sum p + offset is exact, which it
// is unless the coordinate lies in (-tiny, 0) (`slow`: the float64 formula decides).  `top` = inverted approximate key
// class of |p mod res|^2 (road points first): larger = better.
struct VoxF { uint32_t bit, top; bool in, slow; };
__device__ __forceinline__ uint32_t vox_top_from_cls(uint32_t cls, bool notroad) { return kVoxTopMax - (cls + (notroad ? (1u << 20) : 0u)); }
__device__ __forceinline__ VoxF vox_fast32(float x, float y, float z, uint32_t lab, const GridF& g) {
  VoxF v;
  const float qx = __fmul_rn(x, g.inv_res), qy = __fmul_rn(y, g.inv_res), qz = __fmul_rn(z, g.inv_res);
  const float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
  // integer-valued floats; the conversion saturates, so "0 <= t < dim" is one unsigned compare per axis
  const uint32_t ix = (uint32_t)__float2int_rz(__fadd_rn(fx, g.offq[0])), iy = (uint32_t)__float2int_rz(__fadd_rn(fy, g.offq[1])),
                 iz = (uint32_t)__float2int_rz(__fadd_rn(fz, g.offq[2]));
  const float rx = __fsub_rn(qx, fx), ry = __fsub_rn(qy, fy), rz = __fsub_rn(qz, fz);
  const float d = __fmaf_rn(rz, rz, __fmaf_rn(ry, ry, __fmul_rn(rx, rx)));
  v.in = (ix < g.dimu[0]) & (iy < g.dimu[1]) & (iz < g.dimu[2]) & (d < 3.5f);     // d is NaN for NaN / Inf coordinates
  const uint32_t cls = __float2uint_rz(__fmul_rn(d, kVoxClsScale));
  v.top = vox_top_from_cls(cls, (int)lab != g.road);
  v.bit = ix * g.sx + iy * g.sy + iz;
  const uint32_t ux = __float_as_uint(x) - 0x80000001u, uy = __float_as_uint(y) - 0x80000001u, uz = __float_as_uint(z) - 0x80000001u;
  v.slow = (ux < g.tiny_m1[0]) | (uy < g.tiny_m1[1]) | (uz < g.tiny_m1[2]);
  return v;
}
// class of an exactly computed |p mod res|^2 (float64 path): same scale as vox_fast32
__device__ __forceinline__ uint32_t vox_cls_exact(double dis, const GridF& g) {
  const double c = dis * (double)g.inv_res2_cls;
  return c < 4194303.0 ? (uint32_t)c : 4194303u;
}

// Pixel words: [inverted top bits of the float64 squared range | inverted 1-based index], 40 | 24 bits for frames with
// fewer than 2^24 - 1 points (`packl`), else 32 | 32.  The class is an exact truncation of the reference's key (the depth
// orders like its square), so two words only need the exact protocol when their class bits are EQUAL.
struct PixF { uint32_t pix; u64 word; bool slow; };
__device__ __forceinline__ u64 pix_word(double s64, uint32_t me1, bool packl) {
  const u64 inv = ~(u64)__double_as_longlong(s64);
  return packl ? ((inv & ~0xffffffull) | (u64)((~me1) & 0xffffffu)) : ((inv & ~0xffffffffull) | (u64)(uint32_t)(~me1));
}
__device__ __forceinline__ u64 pix_cls(u64 w, bool packl) { return packl ? (w >> 24) : (w >> 32); }
__device__ __forceinline__ uint32_t pix_idx1(u64 w, bool packl) { return packl ? ((~(uint32_t)w) & 0xffffffu) : ~(uint32_t)w; }
// f32 squared range as the hot loop computes it (magnitude test only)
__device__ __forceinline__ float range_s32(float x, float y, float z, const RangeDev& r, float* xf_o, float* yf_o, float* zf_o) {
  const float xf = x - r.Lf[0], yf = -((-y) - r.Lf[1]), zf = z - r.Lf[2];   // same zero signs as geometry_utils.py:177-183
  *xf_o = xf; *yf_o = yf; *zf_o = zf;
  return __fmaf_rn(zf, zf, __fmaf_rn(xf, xf, __fmul_rn(yf, yf)));
}
__device__ __forceinline__ PixF pix_fast32(float x, float y, float z, uint32_t me1, bool packl, const RangeDev& r) {
  PixF k;
  float xf, yf, zf;
  const float s = range_s32(x, y, z, r, &xf, &yf, &zf);
  float pw, ph;
  pix_coords_f32(xf, yf, zf, r, &pw, &ph);
  const float fw = floorf(pw), fh = floorf(ph);
  const bool safe_w = fabsf((pw - fw) - 0.5f) < r.safe_w, safe_h = fabsf((ph - fh) - 0.5f) < r.safe_h;
  // 1e-12 < s < 1e12 (metres^2): inside, no f32 square over/underflows in a way that could move a pixel
  const bool mag_ok = (__float_as_uint(s) - 0x2b8cbcccu) < (0x5368d4a5u - 0x2b8cbcccu);
  k.slow = !(safe_w & safe_h & mag_ok);     // NaN compares false -> slow
  const int iw = (int)fminf(fmaxf(fw, 0.0f), r.w_max);
  const int ih = (int)fminf(fmaxf(fh, 0.0f), r.h_max);
  k.pix = (uint32_t)(ih * r.W + iw);
  double xc, yc, zc;
  k.word = pix_word(range_sq_of(x, y, z, r, &xc, &yc, &zc), me1, packl);      // float64, numpy's order (geometry_utils.py:177-180)
  return k;
}

// A lane whose neighbour lane targets the same cell with a class better by 2 or more can never be the exact winner.
// (shfl_up / shfl_down hand lanes 0 / 31 their own values back: equal class, never "better by 2")
__device__ __forceinline__ bool neighbour_beaten(uint32_t cell, uint32_t top) {
  const uint32_t c_up = __shfl_up_sync(0xffffffffu, cell, 1), t_up = __shfl_up_sync(0xffffffffu, top, 1);
  const uint32_t c_dn = __shfl_down_sync(0xffffffffu, cell, 1), t_dn = __shfl_down_sync(0xffffffffu, top, 1);
  const uint32_t need = top + 2u;
  return ((cell == c_up) & (t_up >= need)) | ((cell == c_dn) & (t_dn >= need));
}

__device__ __forceinline__ void queue_push(uint2* q, uint32_t* qn, uint32_t x, uint32_t y) {
  q[atomicAdd(qn, 1u)] = make_uint2(x, y);
}
// |top_a - top_b| <= 1 (also false for an empty slot: its top is 0 and every real top is >= 2)
__device__ __forceinline__ bool in_band(uint32_t top_a, uint32_t top_b) { return (top_a - top_b + 1u) <= 2u; }

constexpr uint32_t kCellNone = 0xffffffffu;    // no claim (outside the grid, beaten by a neighbour lane, not a point)
constexpr uint32_t kCellSlow = 0xfffffffeu;    // the f32 arithmetic cannot decide the cell

// Pixel of a point whose f32 estimate lies within eps of a bin edge, decided in float64 against the edge itself:
// column edge e <=> yaw = pi (1 - 2e/W): the sign of the cross product of the edge direction with (x_c, -y_c) says on which
// side the point lies; row edge e <=> pitch = fov (1 - e/H) - |fov_down|: compare z_c with depth * sin(edge pitch).
// The reference's own float64 rounding moves a coordinate by < 1e-12 bins, so outside a 1e-9 (relative) zone around the
// edge the geometric side is the reference's bin; inside it (or for wild magnitudes) the answer is "undecided" and the
// point goes to the frame's queue, where the reference formula itself is evaluated (pix_exact).
__device__ __noinline__ uint32_t pix_refine(float x, float y, float z, const RangeDev& r, const double* __restrict__ edges) {
  float xf, yf, zf;
  const float s32 = range_s32(x, y, z, r, &xf, &yf, &zf);
  if (!((__float_as_uint(s32) - 0x2b8cbcccu) < (0x5368d4a5u - 0x2b8cbcccu))) return kCellSlow;
  float pw, ph;
  pix_coords_f32(xf, yf, zf, r, &pw, &ph);
  if (!(pw > -8.f && pw < (float)r.W + 8.f && ph > -1e6f && ph < 1e6f)) return kCellSlow;
  double xc, yc, zc;
  const double s = range_sq_of(x, y, z, r, &xc, &yc, &zc);
  const float fw = floorf(pw), fh = floorf(ph);
  int iw = (int)fw, ih = (int)fh;
  const float dw = pw - fw, dh = ph - fh;
  if (!(fabsf(dw - 0.5f) < r.safe_w)) {
    const int e = dw < 0.5f ? iw : iw + 1;                 // the edge between bins e - 1 and e
    if (e > 0 && e < r.W) {
      const double yy = -yc;
      const double cross = edges[2 * e] * yy - edges[2 * e + 1] * xc;
      if (!(fabs(cross) > 1e-9 * (fabs(xc) + fabs(yy)))) return kCellSlow;
      iw = cross > 0.0 ? e - 1 : e;
    }
  }
  if (!(fabsf(dh - 0.5f) < r.safe_h)) {
    const int e = dh < 0.5f ? ih : ih + 1;
    if (e > 0 && e < r.H) {
      const double depth = sqrt(s);
      const double t = zc - depth * edges[2 * (r.W + 1) + e];
      if (!(fabs(t) > 1e-9 * depth)) return kCellSlow;
      ih = t > 0.0 ? e - 1 : e;
    }
  }
  iw = iw < 0 ? 0 : (iw > r.W - 1 ? r.W - 1 : iw);
  ih = ih < 0 ? 0 : (ih > r.H - 1 ? r.H - 1 : ih);
  return (uint32_t)(ih * r.W + iw);
}

// ---------------------------------------------------------------- P unit
// Per half (4 points per thread, lanes = consecutive points so that a warp's atomics fall into a few 32-byte sectors):
// arithmetic -> [rare] cells the f32 path cannot decide -> neighbour filter -> all atomics in flight -> results.
// Everything rare (undecided cells, in-band holders) is gathered behind ONE branch per stage.
// undecided pixel: refine against the edges; still undecided -> the frame's queue (returns kCellNone then)
__device__ __noinline__ uint32_t pixel_undecided(float x, float y, float z, const RangeDev& r, const double* __restrict__ edges,
                                                 uint2* q, uint32_t* qn, uint32_t rel) {
  const uint32_t c = pix_refine(x, y, z, r, edges);
  if (c != kCellSlow) return c;
  queue_push(q, qn, rel, kQSlow);
  return kCellNone;
}
__device__ __noinline__ void queue_push_cold(uint2* q, uint32_t* qn, uint32_t x, uint32_t y) { queue_push(q, qn, x, y); }

template <bool DO_VOX, bool DO_RANGE, bool FULL>
__device__ __forceinline__ void unit_points(const MegaArgs& a, const UnitDesc& d, const unsigned char* tile, int ct, unsigned& n_in) {
  const float* sx = reinterpret_cast<const float*>(tile) + 3 * (d.lead + ct);
  const uint8_t* ss = tile + kTileXyzBytes + d.lead + ct;
  u64* pixtab_s = a.pixtab + (size_t)d.slot * a.HW;
  u64* vtab_s = a.vtab + (size_t)d.slot * a.g.G;
  uint32_t* bitmap_s = a.bitmap + (size_t)d.slot * a.g.gw;
  asm volatile("" : "+l"(pixtab_s), "+l"(vtab_s), "+l"(bitmap_s));
  uint2* q = a.queue + 2 * d.fbeg;
  uint32_t* qn = a.s.qn + d.f;
  const uint32_t rel0 = (uint32_t)(d.a - d.fbeg) + (uint32_t)ct;     // frame-relative index of this thread's first point
  const bool packl = d.packl != 0;
  const uint64_t pol = l2_evict_last_policy();
#pragma unroll
  for (int h = 0; h < kPPT / kBatch; ++h) {
    uint32_t vcell[kBatch], vtop[kBatch], pcell[kBatch];
    u64 pword[kBatch];
    bool undecided = false;
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int kk = h * kBatch + k;
      const float x = sx[3 * kk * kConsumers], y = sx[3 * kk * kConsumers + 1], z = sx[3 * kk * kConsumers + 2];
      vcell[k] = kCellNone; pcell[k] = kCellNone; vtop[k] = 0u; pword[k] = 0ull;
      if (DO_VOX) {
        const VoxF v = vox_fast32(x, y, z, ss[kk * kConsumers], a.gf);
        vtop[k] = v.top;
        vcell[k] = v.in ? v.bit : kCellNone;
        if (v.slow) vcell[k] = kCellSlow;
      }
      if (DO_RANGE) {
        const PixF p = pix_fast32(x, y, z, rel0 + kk * kConsumers + 1u, packl, a.r);
        pword[k] = p.word;
        pcell[k] = p.slow ? kCellSlow : p.pix;
      }
      if (!FULL && !(ct + kk * kConsumers < d.n)) { vcell[k] = kCellNone; pcell[k] = kCellNone; }   // the frame's last, partial unit
      undecided |= (DO_VOX && vcell[k] == kCellSlow) | (DO_RANGE && pcell[k] == kCellSlow);
    }
    if (undecided) {
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int kk = h * kBatch + k;
        const uint32_t rel = rel0 + kk * kConsumers;
        if (DO_VOX && vcell[k] == kCellSlow) { queue_push_cold(q, qn, rel | kQVoxel, kQSlow); vcell[k] = kCellNone; }
        if (DO_RANGE && pcell[k] == kCellSlow)
          pcell[k] = pixel_undecided(sx[3 * kk * kConsumers], sx[3 * kk * kConsumers + 1], sx[3 * kk * kConsumers + 2], a.r, a.edges, q, qn, rel);
      }
    }
    // voxels: neighbour filter (scan-ordered clouds put runs of points into one voxel); then all claims of the half in flight
    u64 vold[kBatch], pold[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const uint32_t nme = ~(rel0 + (h * kBatch + k) * kConsumers + 1u);
      if (DO_VOX) {
        n_in += vcell[k] != kCellNone ? 1u : 0u;
        if (neighbour_beaten(vcell[k], vtop[k])) vcell[k] = kCellNone;
        const uint32_t lo = packl ? nme * 256u + ss[(h * kBatch + k) * kConsumers] : nme;
        vold[k] = atom_max_if(vtab_s, vcell[k], ((u64)vtop[k] << 32) | lo, ~0ull, pol);
      }
      if (DO_RANGE) pold[k] = atom_max_if(pixtab_s, pcell[k], pword[k], ~0ull, pol);
    }
    bool band = false;
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      if (DO_VOX) {
        red_or_if(bitmap_s, vcell[k], word_top(vold[k]) == 0u, pol);    // first claim marks the voxel occupied
        band |= in_band(word_top(vold[k]), vtop[k]);
      }
      if (DO_RANGE) band |= pix_cls(pold[k] ^ pword[k], packl) == 0ull;
    }
    if (band) {
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const uint32_t rel = rel0 + (h * kBatch + k) * kConsumers;
        if (DO_VOX && in_band(word_top(vold[k]), vtop[k])) queue_push_cold(q, qn, rel | kQVoxel, vox_word_idx1(packl, vold[k]));
        if (DO_RANGE && pix_cls(pold[k] ^ pword[k], packl) == 0ull) queue_push_cold(q, qn, rel, pix_idx1(pold[k], packl));
      }
    }
  }
}

// ---------------------------------------------------------------- queue of a frame (run by the WARP that finished the frame's last P share)
// Exact protocol for candidates whose classes are within 1 of each other.  `cand` = exactly better of (me, partner);
// the slot ends up holding a point that is exactly <= cand, or one whose class is better by 2 or more (which is then
// exactly better as well).  Every point that displaced, or failed to displace, an in-band holder runs this, so the final
// holder is the exact arg-min with the lowest index among equal keys.
template <typename KeyFn, typename PackFn, typename IdxFn, typename ClsFn>
__device__ __noinline__ void band_protocol(u64* slot, uint32_t me1, uint32_t partner1, KeyFn key_of, PackFn pack_of, IdxFn idx_of,
                                           ClsFn cls_of, u64 margin) {
  uint32_t cand1 = me1;
  u64 ck = key_of(me1);
  {
    const u64 pk = key_of(partner1);
    if (pk < ck || (pk == ck && partner1 < cand1)) { cand1 = partner1; ck = pk; }
  }
  const u64 cw = pack_of(cand1);
  u64 cur = ld_cg_u64(slot);
  for (;;) {
    if (cls_of(cur) >= cls_of(cw) + margin) break;        // a class that is certainly better holds the slot
    const uint32_t h1 = idx_of(cur);
    if (h1 == cand1) break;
    const u64 hk = key_of(h1);
    const bool better = ck < hk || (ck == hk && cand1 < h1);
    if (!better) break;
    const u64 prev = atomicCAS(slot, cur, cw);
    if (prev == cur) break;
    cur = prev;
  }
}

__device__ __noinline__ void drain_queue(const MegaArgs& a, const UnitDesc& d) {
  const int ct = (int)lane_id();
  const uint32_t n = ld_cg_u32(a.s.qn + d.f);
  if (n == 0u) return;
  uint2* q = a.queue + 2 * d.fbeg;
  u64* pixtab_s = a.pixtab + (size_t)d.slot * a.HW;
  u64* vtab_s = a.vtab + (size_t)d.slot * a.g.G;
  uint32_t* bitmap_s = a.bitmap + (size_t)d.slot * a.g.gw;
  const float* fx = a.xyz + 3 * d.fbeg;
  const uint8_t* fs = a.sem + d.fbeg;
  const bool packl = d.packl != 0;
  const GridDev& g = a.g; const RangeDev& r = a.r;
  // approximate class of point q1 exactly as the hot loop (or phase A below) packed it
  auto vox_top_of = [&](uint32_t q1, const VoxKey& ex) -> uint32_t {
    const float* p = fx + 3 * (int64_t)(q1 - 1u);
    const uint32_t lab = __ldg(fs + (q1 - 1u));
    const VoxF v = vox_fast32(__ldg(p), __ldg(p + 1), __ldg(p + 2), lab, a.gf);
    return v.slow ? vox_top_from_cls(vox_cls_exact(ex.dis, a.gf), (int)lab != g.road) : v.top;
  };
  unsigned n_drop = 0, n_nw = 0, n_nh = 0, n_in = 0;
  long long td0 = clock64();
  unsigned c_sv = 0, c_sp = 0, c_bv = 0, c_bp = 0;
  // phase A: points whose cell comes from the float64 formula make their claim now
  for (uint32_t e = ct; e < n; e += 32) {
    uint2 ent = ld_cg_u32x2(q + e);
    if (ent.y != kQSlow) { if (ent.x & kQVoxel) ++c_bv; else ++c_bp; continue; }
    if (ent.x & kQVoxel) ++c_sv; else ++c_sp;
    const uint32_t rel = ent.x & ~kQVoxel, me1 = rel + 1u;
    const float x = __ldg(fx + 3 * (int64_t)rel), y = __ldg(fx + 3 * (int64_t)rel + 1), z = __ldg(fx + 3 * (int64_t)rel + 2);
    uint32_t partner = 0u;
    if (ent.x & kQVoxel) {
      const VoxKey v = vox_of<true>((double)x, (double)y, (double)z, g);
      if (v.in) {
        ++n_in;
        const uint32_t lab = __ldg(fs + rel);
        const uint32_t top = vox_top_from_cls(vox_cls_exact(v.dis, a.gf), (int)lab != g.road);
        const u64 mine = vox_word(packl, top, me1, lab);
        const u64 old = atom_max_global(vtab_s + v.bit, mine);
        if (old == 0ull) red_or_global(bitmap_s + (v.bit >> 5), 1u << (v.bit & 31));
        else if (in_band(word_top(old), top)) partner = vox_word_idx1(packl, old);
      }
    } else {
      double xc, yc, zc;
      const double s = range_sq_of(x, y, z, r, &xc, &yc, &zc);
      int iw, ih, flags;
      pix_exact(xc, yc, zc, s, r.H, r.W, r.fda, r.fov, &iw, &ih, &flags);
      if (flags & 4) { ++n_drop; }
      else {
        n_nw += (flags & 1) ? 1u : 0u; n_nh += (flags & 2) ? 1u : 0u;
        const u64 mine = pix_word(s, me1, packl);
        const u64 old = atom_max_global(pixtab_s + (ih * r.W + iw), mine);
        if (old != 0ull && pix_cls(old ^ mine, packl) == 0ull) partner = pix_idx1(old, packl);
      }
    }
    q[e] = make_uint2(ent.x, partner);      // 0 = settled
  }
  __threadfence();
  __syncwarp();
  if (ct == 0) { stat_add(17, clock64() - td0); stat_add(12, n); }
  atomicAdd(&g_mega_stats[blockIdx.x % kStatCtas][13], (unsigned long long)c_sv);
  atomicAdd(&g_mega_stats[blockIdx.x % kStatCtas][14], (unsigned long long)c_sp);
  atomicAdd(&g_mega_stats[blockIdx.x % kStatCtas][15], (unsigned long long)c_bv);
  atomicAdd(&g_mega_stats[blockIdx.x % kStatCtas][16], (unsigned long long)c_bp);
  td0 = clock64();
  // phase B: exact protocol
  for (uint32_t e = ct; e < n; e += 32) {
    const uint2 ent = ld_cg_u32x2(q + e);
    if (ent.y == 0u || ent.y == kQSlow) continue;
    const uint32_t rel = ent.x & ~kQVoxel, me1 = rel + 1u;
    const float x = __ldg(fx + 3 * (int64_t)rel), y = __ldg(fx + 3 * (int64_t)rel + 1), z = __ldg(fx + 3 * (int64_t)rel + 2);
    if (ent.x & kQVoxel) {
      const VoxKey v = vox_of<true>((double)x, (double)y, (double)z, g);
      if (!v.in) continue;
      auto key_of = [&](uint32_t q1) -> u64 {
        const float* p = fx + 3 * (int64_t)(q1 - 1u);
        const VoxKey o = vox_of<true>((double)__ldg(p), (double)__ldg(p + 1), (double)__ldg(p + 2), g);
        return vox_key(o.dis, (int)__ldg(fs + (q1 - 1u)) != g.road);
      };
      auto pack_of = [&](uint32_t q1) -> u64 {
        const float* p = fx + 3 * (int64_t)(q1 - 1u);
        const VoxKey o = vox_of<true>((double)__ldg(p), (double)__ldg(p + 1), (double)__ldg(p + 2), g);
        return vox_word(packl, vox_top_of(q1, o), q1, packl ? (uint32_t)__ldg(fs + (q1 - 1u)) : 0u);
      };
      auto idx_of = [&](u64 wv) -> uint32_t { return vox_word_idx1(packl, wv); };
      auto cls_of = [](u64 wv) -> u64 { return wv >> 32; };
      band_protocol(vtab_s + v.bit, me1, ent.y, key_of, pack_of, idx_of, cls_of, 2ull);
    } else {
      // the pixel: f32 path unless this point itself was decided by the float64 formula (then it is recomputed exactly)
      const PixF pf = pix_fast32(x, y, z, me1, packl, r);
      uint32_t pix = pf.pix;
      if (pf.slow && (pix = pix_refine(x, y, z, r, a.edges)) == kCellSlow) {
        double xc, yc, zc;
        const double s = range_sq_of(x, y, z, r, &xc, &yc, &zc);
        int iw, ih, flags;
        pix_exact(xc, yc, zc, s, r.H, r.W, r.fda, r.fov, &iw, &ih, &flags);
        if (flags & 4) continue;
        pix = (uint32_t)(ih * r.W + iw);
      }
      auto key_of = [&](uint32_t q1) -> u64 {     // exact key: the float64 depth (geometry_utils.py:180)
        const float* p = fx + 3 * (int64_t)(q1 - 1u);
        double aa, bb, cc;
        return (u64)__double_as_longlong(sqrt(range_sq_of(__ldg(p), __ldg(p + 1), __ldg(p + 2), r, &aa, &bb, &cc)));
      };
      auto pack_of = [&](uint32_t q1) -> u64 {
        const float* p = fx + 3 * (int64_t)(q1 - 1u);
        double aa, bb, cc;
        return pix_word(range_sq_of(__ldg(p), __ldg(p + 1), __ldg(p + 2), r, &aa, &bb, &cc), q1, packl);
      };
      auto idx_of = [&](u64 wv) -> uint32_t { return pix_idx1(wv, packl); };
      auto cls_of = [&](u64 wv) -> u64 { return pix_cls(wv, packl); };
      band_protocol(pixtab_s + pix, me1, ent.y, key_of, pack_of, idx_of, cls_of, 1ull);
    }
  }
  if (ct == 0) stat_add(18, clock64() - td0);
  if (a.diag) {
    __syncwarp();
    diag_add(a.diag, MUVO_DIAG_DROPPED_NONFINITE, n_drop);
    diag_add(a.diag, MUVO_DIAG_NEAR_EDGE_W, n_nw);
    diag_add(a.diag, MUVO_DIAG_NEAR_EDGE_H, n_nh);
    diag_add(a.diag, MUVO_DIAG_IN_GRID, n_in);
  }
}

// ---------------------------------------------------------------- ER unit: 1792 pixels of one range image
template <int LAYOUT>
__device__ __forceinline__ void unit_emit_range(const MegaArgs& a, const UnitDesc& d, int ct) {
  const RangeDev& r = a.r;
  const int64_t HW = a.HW;
  u64* pixtab_s = a.pixtab + (size_t)d.slot * HW;
  const float* fx = a.xyz + 3 * d.fbeg;
  const uint8_t* fs = a.sem + d.fbeg;
  const uint64_t pol = l2_evict_last_policy();
  const bool packl = d.packl != 0;
  u64 wv[2][4];
  int64_t pin[2];
#pragma unroll
  for (int gq = 0; gq < 2; ++gq) {              // both groups' table reads first (two round trips in flight)
    pin[gq] = (int64_t)d.j * kPixUnit + (int64_t)(gq * kConsumers + ct) * 4;
    wv[gq][0] = wv[gq][1] = wv[gq][2] = wv[gq][3] = 0ull;
    if (pin[gq] < HW) {
      const ulonglong2 w0 = ld_u64x2_last(pixtab_s + pin[gq], pol), w1 = ld_u64x2_last(pixtab_s + pin[gq] + 2, pol);
      wv[gq][0] = w0.x; wv[gq][1] = w0.y; wv[gq][2] = w1.x; wv[gq][3] = w1.y;
      if (w0.x | w0.y) st_zero16_last(pixtab_s + pin[gq], pol);
      if (w1.x | w1.y) st_zero16_last(pixtab_s + pin[gq] + 2, pol);
    }
  }
#pragma unroll
  for (int gq = 0; gq < 2; ++gq) {
    if (pin[gq] >= HW) continue;
    float px[4], py[4], pz[4], pd[4];
    uint32_t ps = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      px[k] = py[k] = pz[k] = 0.f; pd[k] = -1.f;                 // geometry_utils.py:210-212 initial values
      if (wv[gq][k]) {
        const int64_t qi = (int64_t)(pix_idx1(wv[gq][k], packl) - 1u);
        const float x = __ldg(fx + 3 * qi), y = __ldg(fx + 3 * qi + 1), z = __ldg(fx + 3 * qi + 2);
        double aa, bb, cc;
        pd[k] = (float)range_depth_of(x, y, z, r, &aa, &bb, &cc);   // :217 float32(depth64)
        px[k] = x; py[k] = y; pz[k] = z;                            // :218 the ego-frame input point
        ps |= (uint32_t)__ldg(fs + qi) << (8 * k);                  // :219
      }
    }
    const int64_t p0 = (int64_t)d.f * HW + pin[gq];
    if (LAYOUT == MUVO_RANGE_LAYOUT_HWC) {
      st_stream_f4(reinterpret_cast<float4*>(a.depth_out + p0), make_float4(pd[0], pd[1], pd[2], pd[3]));
      float4* o = reinterpret_cast<float4*>(a.xyz_out + p0 * 3);
      st_stream_f4(o,     make_float4(px[0], py[0], pz[0], px[1]));
      st_stream_f4(o + 1, make_float4(py[1], pz[1], px[2], py[2]));
      st_stream_f4(o + 2, make_float4(pz[2], px[3], py[3], pz[3]));
    } else {
      if (a.depth_out) st_stream_f4(reinterpret_cast<float4*>(a.depth_out + p0), make_float4(pd[0], pd[1], pd[2], pd[3]));
      float* base = a.xyz_out + (size_t)d.f * 4 * HW + pin[gq];
      st_stream_f4(reinterpret_cast<float4*>(base),          make_float4(px[0], px[1], px[2], px[3]));
      st_stream_f4(reinterpret_cast<float4*>(base + HW),     make_float4(py[0], py[1], py[2], py[3]));
      st_stream_f4(reinterpret_cast<float4*>(base + 2 * HW), make_float4(pz[0], pz[1], pz[2], pz[3]));
      st_stream_f4(reinterpret_cast<float4*>(base + 3 * HW), make_float4(pd[0], pd[1], pd[2], pd[3]));
    }
    if (a.sem_out) st_stream_u32(reinterpret_cast<uint32_t*>(a.sem_out + p0), ps);
  }
}

// ---------------------------------------------------------------- ED unit: 4 blocks x 448 bitmap words = 56 KiB of one dense grid
// (the warp-level scheme of k_emit_dense in points.cu: cooperative gather of the set bits' winners, label bytes assembled in
// a swizzled shared-memory tile, one 256-bit store per lane and bitmap word, zeros included)
__device__ __forceinline__ void unit_emit_dense(const MegaArgs& a, const UnitDesc& d, MegaEmitSmem& sm, int ct) {
  const GridDev& g = a.g;
  const unsigned lane = lane_id();
  const int warp = ct >> 5;
  uint32_t* bm = a.bitmap + (size_t)d.slot * g.gw;
  u64* vtab_s = a.vtab + (size_t)d.slot * g.G;
  const uint8_t* sem_f = a.sem + d.fbeg;
  const bool packl = d.packl != 0;
  uint8_t* df = a.dense + (size_t)d.f * g.G;
  const bool aligned = ((reinterpret_cast<uintptr_t>(df) & 31) == 0);
  uint8_t* tile = sm.tile[warp];
  uint16_t* list = sm.list[warp];
  uint32_t occ = 0;
  const uint64_t pol = l2_evict_last_policy();
  for (int blk = 0; blk < kEdBlocksPerUnit; ++blk) {
    const uint32_t warp_w0 = (((uint32_t)d.j * kEdBlocksPerUnit + blk) * kConsumerWarps + warp) * (32 * kEmitWords);
    if (warp_w0 >= (uint32_t)g.gw) break;                         // warp-uniform (gw % 32 == 0)
    uint32_t bits[kEmitWords];
    uint32_t pc = 0;
#pragma unroll
    for (int k = 0; k < kEmitWords; ++k) {
      const uint32_t wi = warp_w0 + 32u * k + lane;
      bits[k] = wi < (uint32_t)g.gw ? ld_u32_last(bm + wi, pol) : 0u;
      pc += __popc(bits[k]);
    }
    __syncwarp();                                                 // previous block's tile reads are done
#pragma unroll
    for (int k = 0; k < kEmitWords; ++k) {                        // zero this lane's rows, clear its bitmap words
      uint4* row = reinterpret_cast<uint4*>(tile + (k * 32 + lane) * 32);
      row[0] = make_uint4(0u, 0u, 0u, 0u); row[1] = make_uint4(0u, 0u, 0u, 0u);
      if (bits[k]) st_u32_last(bm + warp_w0 + 32u * k + lane, 0u, pol);
    }
    uint32_t incl = pc;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, dd); if (lane >= dd) incl += t; }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    occ += total;
    if (total) {
      u64* vt = vtab_s + (size_t)warp_w0 * 32;
      for (uint32_t done = 0; done < total; done += kEmitList) {
        uint32_t n = incl - pc;
#pragma unroll
        for (int k = 0; k < kEmitWords; ++k) {
          uint32_t b = bits[k];
          while (b) {
            const int j = __ffs(b) - 1;
            b &= b - 1;
            if (n - done < (uint32_t)kEmitList) list[n - done] = (uint16_t)((k * 32 + lane) * 32 + j);
            ++n;
          }
        }
        __syncwarp();
        const uint32_t cnt = total - done < (uint32_t)kEmitList ? total - done : (uint32_t)kEmitList;
        for (uint32_t t = lane; t < cnt; t += 32) {
          const uint32_t pos = list[t];
          const u64 wv = ld_u64_last(vt + pos, pol);
          uint32_t lab = vox_word_label(packl, wv, sem_f);
          if (a.remap) lab = __ldg(a.remap + lab);
          tile[tile_swz(pos)] = (uint8_t)lab;
        }
        __syncwarp();
      }
#pragma unroll
      for (int k = 0; k < kEmitWords; ++k) {                      // clear the winners, one 32-byte sector per store
        uint32_t b = bits[k];
        u64* pw = vt + (k * 32 + lane) * 32;
        while (b) {
          const int sct = (__ffs(b) - 1) >> 2;
          b &= ~(0xfu << (4 * sct));
          st_zero32_last(pw + 4 * sct, pol);
        }
      }
    } else {
      __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < kEmitWords; ++k) {
      const uint32_t rr = k * 32 + lane, wi = warp_w0 + rr;
      if (wi >= (uint32_t)g.gw) continue;
      const uint32_t sw = (rr >> 2) & 1u;
      const uint4 h0 = *reinterpret_cast<const uint4*>(tile + rr * 32 + 16 * sw);
      const uint4 h1 = *reinterpret_cast<const uint4*>(tile + rr * 32 + 16 * (sw ^ 1u));
      const uint32_t o[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
      const int64_t vox = (int64_t)wi * 32;
      uint8_t* dst = df + vox;
      if (vox + 32 <= g.G && aligned) {
        st_stream_u8x32(dst, o);
      } else {
        for (int j = 0; j < 32; ++j)
          if (vox + j < g.G) dst[j] = (uint8_t)(o[j >> 2] >> (8 * (j & 3)));
      }
    }
  }
  if (a.n_occ && occ && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(a.n_occ + d.f), (unsigned long long)occ);
}

// ---------------------------------------------------------------- per-call initialisation (one CTA)
__global__ void __launch_bounds__(1024)
k_mega_init(const int64_t* __restrict__ off, int F, SyncWs s, int64_t* __restrict__ n_occ, double* __restrict__ edges, int H, int W,
            double fda, double fov) {
  const int tid = threadIdx.x;
  if (edges) {                                                    // bin edges of the range image (pix_refine)
    for (int e = tid; e <= W; e += 1024) {
      const double th = kPi * (1.0 - 2.0 * (double)e / (double)W);
      edges[2 * e] = cos(th); edges[2 * e + 1] = sin(th);
    }
    for (int e = tid; e <= H; e += 1024) edges[2 * (W + 1) + e] = sin(fov * (1.0 - (double)e / (double)H) - fda);
  }
  for (int f = tid; f < F; f += 1024) {
    s.p_claim[f] = 0u; s.e_claim[f] = 0u; s.p_done[f] = 0u; s.e_done[f] = 0u; s.qn[f] = 0u;
    s.ready[f] = (off[f + 1] - off[f]) > 0 ? 0u : 1u;             // a frame without points has nothing to wait for
    if (n_occ) n_occ[f] = 0;
  }
}

// ---------------------------------------------------------------- the kernel
constexpr int kStages = 2;
constexpr int kSmemDesc = kStages * kStageBytes;
constexpr int kSmemBars = kSmemDesc + kStages * 64;
constexpr int kMegaSmemBytes = kSmemBars + 4 * 8;
static_assert(sizeof(UnitDesc) <= 64, "descriptor slot");

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Producer warp.  Work is claimed per frame and only once it can run: an emit unit of the oldest frame whose points are
// all in (emits free table slots, so they go first), else a P unit of the oldest frame whose table slot has been emitted
// and cleared (frame f uses slot f % R, i.e. waits for the emits of frame f - R).  Nothing is ever held while it waits for
// somebody else, so the scheme cannot deadlock, and at most R frames are in flight.
__device__ __forceinline__ void producer_loop(const MegaArgs& a, unsigned char* smem) {
  UnitDesc* desc = reinterpret_cast<UnitDesc*>(smem + kSmemDesc);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  uint64_t* empty = full + kStages;
  const unsigned lane = lane_id();
  const uint32_t nE = (uint32_t)(a.nER + a.nED);
  int fe = 0, fp = 0;                                            // first frame that may still have E / P units to hand out
  for (uint32_t it = 0;; ++it) {
    const int s = (int)(it & 1u);
    long long tp0 = clock64();
    if (it >= (uint32_t)kStages) mbar_wait(empty + s, ((it >> 1) - 1u) & 1u);
    if (lane == 0) stat_add(5, clock64() - tp0);
    // ---- lane 0 claims a unit: type (U_DONE = everything has been handed out), frame, index
    int type = U_DONE, uf = 0, uj = 0;
    if (lane == 0) {
      tp0 = clock64();
      for (;;) {
        bool got = false;
        for (int f = fe; f < a.F; ++f) {                         // emits, oldest frame first
          if (ld_relaxed_u32(a.s.e_claim + f) >= nE) { if (f == fe) ++fe; continue; }
          if (ld_acquire_u32(a.s.ready + f) == 0u) break;
          const uint32_t t = atomicAdd(a.s.e_claim + f, 1u);
          if (t < nE) { type = t < (uint32_t)a.nER ? U_RANGE : U_DENSE; uf = f; uj = (int)(t < (uint32_t)a.nER ? t : t - a.nER); got = true; break; }
        }
        if (got) break;
        if (fe >= a.F) break;                                    // all emits handed out: done
        for (int f = fp; f < a.F && f < fe + a.R; ++f) {         // points, oldest frame first; its slot must be free
          const int64_t n = __ldg(a.off + f + 1) - __ldg(a.off + f);
          const uint32_t nP = n > 0 ? (uint32_t)((n + kUnit - 1) / kUnit) : 0u;
          if (ld_relaxed_u32(a.s.p_claim + f) >= nP) { if (f == fp) ++fp; continue; }
          if (f >= a.R && ld_acquire_u32(a.s.e_done + (f - a.R)) < nE * kConsumerWarps) break;
          const uint32_t t = atomicAdd(a.s.p_claim + f, 1u);
          if (t < nP) { type = U_POINTS; uf = f; uj = (int)t; got = true; break; }
        }
        if (got) break;
        __nanosleep(200);
      }
      stat_add(6, clock64() - tp0);
    }
    type = __shfl_sync(0xffffffffu, type, 0); uf = __shfl_sync(0xffffffffu, uf, 0); uj = __shfl_sync(0xffffffffu, uj, 0);
    UnitDesc& d = *reinterpret_cast<UnitDesc*>(reinterpret_cast<unsigned char*>(desc) + s * 64);
    if (type == U_DONE) {
      if (lane == 0) { d.type = U_DONE; mbar_arrive(full + s); }
      break;
    }
    unsigned char* tile = smem + s * kStageBytes;
    const int64_t fbeg = __ldg(a.off + uf), fend = __ldg(a.off + uf + 1);
    if (type != U_POINTS) {
      if (lane == 0) {
        d.type = type; d.j = uj; d.f = uf; d.slot = uf % a.R; d.fbeg = fbeg; d.packl = (fend - fbeg) < kPackLimit ? 1 : 0;
        d.n = 0; d.lead = 0; d.a = fbeg; d.n_units = 0;
        mbar_arrive(full + s);
      }
    } else {
      const int64_t pa = fbeg + (int64_t)uj * kUnit;
      const int64_t pb = pa + kUnit < fend ? pa + kUnit : fend;
      const int64_t a0 = pa & ~(int64_t)15;                      // the bulk copies start and end on multiples of 16 points
      const int64_t body_end = (pb & ~(int64_t)15) > a0 ? (pb & ~(int64_t)15) : a0;
      const uint32_t nb = (uint32_t)(body_end - a0);
      // the last (< 16) points of the unit, which a 16-byte granular copy cannot fetch without reading past the frame
      const int64_t tp = body_end + lane;
      if (tp < pb) {
        float* tx = reinterpret_cast<float*>(tile) + 3 * (tp - a0);
        tx[0] = __ldg(a.xyz + 3 * tp); tx[1] = __ldg(a.xyz + 3 * tp + 1); tx[2] = __ldg(a.xyz + 3 * tp + 2);
        (tile + kTileXyzBytes)[tp - a0] = __ldg(a.sem + tp);
      }
      __syncwarp();
      if (lane == 0) {
        d.type = U_POINTS; d.f = uf; d.j = uj; d.slot = uf % a.R; d.fbeg = fbeg; d.a = pa;
        d.n = (int)(pb - pa); d.lead = (int)(pa - a0); d.packl = (fend - fbeg) < kPackLimit ? 1 : 0;
        d.n_units = (int)((fend - fbeg + kUnit - 1) / kUnit);
        if (nb) {
          mbar_expect_tx(full + s, nb * 13u);
          bulk_g2s_plain(tile, a.xyz + 3 * a0, nb * 12u, full + s);
          bulk_g2s_plain(tile + kTileXyzBytes, a.sem + a0, nb, full + s);
        } else {
          mbar_arrive(full + s);
        }
      }
    }
    __syncwarp();
  }
}

__device__ __forceinline__ uint32_t atom_add_acq_rel(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void red_add_release(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Consumer warps run through the units independently of each other (the stage's mbarriers are the only CTA-level
// synchronisation): a warp processes its share of the unit and signals the frame's counter with a release; the warp that
// completes a frame's points settles the frame's queue and publishes `ready`.
template <bool DO_VOX, bool DO_RANGE>
__device__ __forceinline__ void consumer_loop(const MegaArgs& a, unsigned char* smem, int ct) {
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  uint64_t* empty = full + kStages;
  const unsigned lane = lane_id();
  unsigned n_in = 0;
  const long long t_begin = clock64();
  for (uint32_t it = 0;; ++it) {
    const int s = (int)(it & 1u);
    long long tc0 = clock64();
    mbar_wait(full + s, (it >> 1) & 1u);
    long long tc1 = clock64();
    if (ct == 0) stat_add(0, tc1 - tc0);
    const UnitDesc d = *reinterpret_cast<const UnitDesc*>(smem + kSmemDesc + s * 64);
    if (d.type == U_DONE) break;
    unsigned char* tile = smem + s * kStageBytes;
    if (d.type == U_POINTS) {
      if (d.n == kUnit) unit_points<DO_VOX, DO_RANGE, true>(a, d, tile, ct, n_in);
      else unit_points<DO_VOX, DO_RANGE, false>(a, d, tile, ct, n_in);
      __syncwarp();
      uint32_t prev = 0;
      if (lane == 0) prev = atom_add_acq_rel(a.s.p_done + d.f, 1u);
      prev = __shfl_sync(0xffffffffu, prev, 0);
      if (ct == 0) { stat_add(1, clock64() - tc1); stat_add(8, 1); }
      if (prev + 1u == (uint32_t)d.n_units * kConsumerWarps) {    // this warp completed the frame's points: settle its queue
        tc1 = clock64();
        drain_queue(a, d);
        __syncwarp();
        if (lane == 0) { st_release_u32(a.s.ready + d.f, 1u); stat_add(4, clock64() - tc1); }
      }
    } else {
      if (d.type == U_RANGE) {
        if (a.layout == MUVO_RANGE_LAYOUT_HWC) unit_emit_range<MUVO_RANGE_LAYOUT_HWC>(a, d, ct);
        else unit_emit_range<MUVO_RANGE_LAYOUT_XYZD>(a, d, ct);
      } else {
        unit_emit_dense(a, d, *reinterpret_cast<MegaEmitSmem*>(tile), ct);
        fence_proxy_async();                                     // generic writes to the stage before a later bulk copy into it
      }
      __syncwarp();
      if (lane == 0) red_add_release(a.s.e_done + d.f, 1u);
      if (ct == 0) { stat_add(d.type == U_RANGE ? 2 : 3, clock64() - tc1); stat_add(d.type == U_RANGE ? 9 : 10, 1); }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
  }
  if (ct == 0) stat_add(11, clock64() - t_begin);
  if (a.diag && DO_VOX) diag_add(a.diag, MUVO_DIAG_IN_GRID, n_in);
}

template <bool DO_VOX, bool DO_RANGE, int MINB>
__global__ void __launch_bounds__(kMegaThreads, MINB)
k_points_mega(const __grid_constant__ MegaArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  if (threadIdx.x == 0) {
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSmemBars);
    mbar_init(full, 1); mbar_init(full + 1, 1);
    mbar_init(full + 2, kConsumerWarps); mbar_init(full + 3, kConsumerWarps);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();                                                    // the schedule kernel's writes are visible
  const int warp = threadIdx.x >> 5;
  if (warp == kConsumerWarps) producer_loop(a, smem);
  else consumer_loop<DO_VOX, DO_RANGE>(a, smem, (int)threadIdx.x);
}

}  // namespace

// ---------------------------------------------------------------- host side
bool points_mega_eligible(const GridDev* g, const RangeDev* r, const void* xyz, const void* sem, const void* dense,
                          const void* depth_out, const void* xyz_out, const void* sem_out, int layout) {
  if (g_tuning[3] != 2) return false;                            // (work in progress: opt-in with tuning key 3 = 2 until it beats the multi-launch path)
  if ((reinterpret_cast<uintptr_t>(xyz) | reinterpret_cast<uintptr_t>(sem)) & 15) return false;
  if (g) {
    if (!dense || !g->regular) return false;
    const float ir = (float)g->inv_res;
    if ((double)ir != g->inv_res || !(ir > 0.f) || !isfinite(ir)) return false;
    if ((double)(float)g->res != g->res) return false;
    for (int k = 0; k < 3; ++k) {
      const double oq = g->off[k] * g->inv_res;
      if (oq != floor(oq) || fabs(oq) > 4194304.0) return false;
    }
    if (g->dx > 65535 || g->dy > 65535 || g->dz > 65535) return false;
  }
  if (r) {
    if (!r->lf_exact) return false;
    const int64_t HW = (int64_t)r->H * r->W;
    if (HW % 4) return false;
    if ((reinterpret_cast<uintptr_t>(xyz_out) & 15) || (depth_out && (reinterpret_cast<uintptr_t>(depth_out) & 15)) ||
        (sem_out && (reinterpret_cast<uintptr_t>(sem_out) & 3)))
      return false;
    if (layout == MUVO_RANGE_LAYOUT_HWC && (!depth_out || !sem_out)) return false;
  }
  return true;
}

int points_mega_f32(const float* xyz, const uint8_t* sem, const int64_t* off, int F, int64_t P, const GridDev* g,
                    const uint8_t* remap, const RangeDev* r, int layout, uint8_t* dense, int64_t* n_occ, float* depth_out,
                    float* xyz_out, uint8_t* sem_out, int64_t* diag, const PointsWs& w, cudaStream_t st) {
  MegaArgs a{};
  a.xyz = xyz; a.sem = sem; a.off = off; a.F = F; a.P = P;
  a.do_vox = g != nullptr; a.do_range = r != nullptr; a.layout = layout;
  a.filter = (g_tuning[1] & 1) ? 0 : 1;                          // tuning key 1 bit 0: no neighbour filter in front of the atomics
  if (g) {
    a.g = *g;
    a.gf.inv_res = (float)g->inv_res;
    a.gf.inv_res2_cls = (float)(g->inv_res * g->inv_res * (double)kVoxClsScale);
    for (int k = 0; k < 3; ++k) {
      a.gf.offq[k] = (float)(g->off[k] * g->inv_res);
      // p + off is exact in float64 when the lowest bit of p (>= |p| 2^-23) is no more than 52 binary places below the top bit
      // of the sum: |p| >= 2^(e_off - 28) for an offset with at most 22 significant bits; x2 margin; denormals always slow
      int e = 0;
      frexp(fabs(g->off[k]), &e);                                  // |off| in [2^(e-1), 2^e)
      float tiny = g->off[k] != 0.0 ? (float)ldexp(1.0, e - 1 - 28 + 1) : 0.f;
      if (!(tiny > 1e-30f)) tiny = 1e-30f;
      uint32_t tb; memcpy(&tb, &tiny, 4);
      a.gf.tiny_m1[k] = tb - 1u;
    }
    a.gf.dimu[0] = (uint32_t)g->dx; a.gf.dimu[1] = (uint32_t)g->dy; a.gf.dimu[2] = (uint32_t)g->dz;
    a.gf.sx = g->sx; a.gf.sy = g->sy;
    a.gf.road = g->road;
  }
  if (r) { a.r = *r; a.HW = (int64_t)r->H * r->W; }
  // ring of table slots = frames in flight; tuning key 4 overrides
  int ring = g_tuning[4] > 0 ? g_tuning[4] : kDefaultRing;
  if (ring > kMaxRing) ring = kMaxRing;
  a.R = F < ring ? F : ring;
  a.nER = r ? (int)ceil_div64(a.HW, kPixUnit) : 0;
  a.nED = g ? (int)ceil_div64(ceil_div64(g->gw, kEdWordsPerBlock), kEdBlocksPerUnit) : 0;
  a.bitmap = w.bitmap; a.vtab = w.vtab; a.pixtab = w.pixtab; a.queue = w.queue;
  a.s.p_claim = w.sync; a.s.e_claim = a.s.p_claim + F; a.s.p_done = a.s.e_claim + F; a.s.ready = a.s.p_done + F;
  a.s.e_done = a.s.ready + F; a.s.qn = a.s.e_done + F;
  a.remap = remap; a.dense = dense; a.n_occ = n_occ; a.depth_out = depth_out; a.xyz_out = xyz_out; a.sem_out = sem_out;
  a.diag = diag; a.edges = w.edges;

  typedef void (*KFn)(const MegaArgs);
  // CTAs per SM: 3 (80 registers, no spills in the point loop) unless tuning key 6 asks for 4 (64 registers)
  const bool four = g_tuning[6] == 4;
  KFn kern = (g && r) ? (four ? k_points_mega<true, true, 4> : k_points_mega<true, true, 3>)
             : g      ? (four ? k_points_mega<true, false, 4> : k_points_mega<true, false, 3>)
                      : (four ? k_points_mega<false, true, 4> : k_points_mega<false, true, 3>);
  cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMegaSmemBytes);
  int per_sm = 0;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)kern, kMegaThreads, kMegaSmemBytes);
  if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  if (per_sm < 1) per_sm = 1;
  if (g_tuning[0] > 0 && g_tuning[0] < per_sm) per_sm = g_tuning[0];
  int sms = kNumSMsB200;
  { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int64_t max_units = (int64_t)F * (a.nER + a.nED) + ceil_div64(P, kUnit) + F;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > max_units) grid = max_units;
  if (grid < 1) grid = 1;

  // The evict_last hints on the table accesses only pin lines inside the persisting-L2 carve-out: reserve it once per device
  // (tuning key 7 = MB, default 48, negative = leave the device setting alone).
  {
    static int done_for_device[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && done_for_device[dev] != g_tuning[7] + 1) {
      done_for_device[dev] = g_tuning[7] + 1;
      if (g_tuning[7] >= 0) {
        int max_persist = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        size_t want = (size_t)(g_tuning[7] > 0 ? g_tuning[7] : 48) << 20;
        if (want > (size_t)max_persist) want = (size_t)max_persist;
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
        cudaGetLastError();
      }
    }
  }
  k_mega_init<<<1, 1024, 0, st>>>(off, F, a.s, n_occ, w.edges, r ? r->H : 0, r ? r->W : 0, r ? r->fda : 0.0, r ? r->fov : 1.0);
  MUVO_AFTER_LAUNCH("k_mega_init", st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kMegaThreads); cfg.dynamicSmemBytes = kMegaSmemBytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  MUVO_AFTER_LAUNCH("k_points_mega", st);
  return MUVO_OK;
}

}  // namespace muvo

extern "C" MUVO_API int muvo_debug_mega_stats(unsigned long long* out_h, int32_t n_ctas, int32_t reset) {
  using namespace muvo;
  if (n_ctas < 0 || n_ctas > kStatCtas) return MUVO_E_ARG;
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess && out_h && n_ctas)
    e = cudaMemcpyFromSymbol(out_h, g_mega_stats, (size_t)n_ctas * kStatSlots * sizeof(unsigned long long));
  if (e == cudaSuccess && reset) {
    void* p = nullptr;
    e = cudaGetSymbolAddress(&p, g_mega_stats);
    if (e == cudaSuccess) e = cudaMemset(p, 0, sizeof(unsigned long long) * kStatCtas * kStatSlots);
  }
  return e == cudaSuccess ? MUVO_OK : (int)e;
}

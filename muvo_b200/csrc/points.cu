// Stages (a) voxelisation and (b) range-view projection, batched over ragged frames.
//
// Reference behaviour being reproduced (bit-exact):
//   (a) voxel_filter            data/data_preprocessing.py:172-228
//       densify                 muvo/data/dataset.py:317-327
//   (b) do_range_projection     muvo/utils/geometry_utils.py:175-220
//
// Both stages are "arg-min with payload" scatters: per voxel the point nearest to the voxel's lower
// corner (roadline points first), per pixel the point nearest to the sensor; ties go to the lowest point
// index.  No float atomics, no sort of the point stream, deterministic result:
//
//   K1  k_points_tile   : persistent CTAs, each owning a contiguous run of 1024-point tiles that arrive in
//                         shared memory through a two-stage cp.async.bulk (TMA) + mbarrier pipeline.  Per point
//                         (this file is compiled with -fmad=false, numpy's operation order):
//                         * voxel id from an exact floor, |p mod res|^2 in float64, ONE 64-bit atomicMax of
//                           (~top32(key) << 32) | ~(index+1) | label on the voxel's word of a direct table; the first
//                           claim of a voxel sets its bit in the 1-bit-per-voxel frame bitmap;
//                         * range pixel from an f32 polynomial atan2 that is PROVABLY on the same side of every
//                           bin edge as the float64 reference formula (points closer than eps to an edge are
//                           queued), squared range in float64, the same packed atomicMax on the pixel word.
//                         The packed order equals the exact (key, index) order unless two competitors agree in the
//                         top 32 key bits; those (rare) are queued for the exact tie protocol.
//   K2  k_scan_queue    : finishes the queues (float64 atan2/asin pixels, tie protocol); for the sorted sparse list
//                         also the bitmap scan: a cluster of 8 CTAs per frame, totals exchanged through DSMEM.
//   K5  k_emit_*        : bitmap-ordered, fully coalesced write of the dense uint8 grid (zeros included, so no
//                         memset + scatter; winners fetched cooperatively, 256-bit stores) and/or the sorted sparse
//                         (n,4) list; pixel-ordered write of the range image.  Emit kernels put every table entry
//                         they consume back to 0, so the workspace is clean for the next call.
#include <math.h>
#include <cooperative_groups.h>
#include <type_traits>
#include "common.cuh"
#include "points_dev.cuh"

namespace cg = cooperative_groups;

namespace muvo {

// debug / tuning knobs (muvo_debug_set_tuning): [0] CTAs per SM of the point pass (0 = as many as fit),
// [1] bit 0 = disable the neighbour filter in front of the voxel atomicMax; the rest is unused
int g_tuning[8] = {0, 0, 0, 0, 0, 0, 0, 0};

namespace {

#ifndef MUVO_K1_MINB
#define MUVO_K1_MINB 4
#endif


// ---------------------------------------------------------------- point tiles
// A CTA owns a contiguous run of tiles (so the frame of its points advances monotonically); tile t+1 is in flight
// while tile t is processed.  Thread `tid` handles points tid, tid+256, ... of the tile (stride-3 word reads of the
// staged xyz are bank-conflict free).  Tiles that are partial (the last one) or whose source is not 16-byte aligned
// are read straight from global memory.
constexpr int kTileThreads = 256;
#ifndef MUVO_KPL
#define MUVO_KPL 4
#endif
constexpr int kKPL = MUVO_KPL;                   // points per thread per tile
constexpr int kTile = kTileThreads * kKPL;       // 1024 points

template <typename T> struct TileLayout {
  static constexpr size_t xyz_bytes = (size_t)3 * kTile * sizeof(T);
  static constexpr size_t off_sem = 2 * xyz_bytes;
  static constexpr size_t off_bar = off_sem + 2 * (size_t)kTile;
  static constexpr size_t off_qn = off_bar + 16;                 // length of this CTA's segment of the rare-path queue
  static constexpr size_t bytes = off_qn + 16;
};

// Rare-path queue (global memory, one segment per CTA = the CTA's own point range, so appends need no global atomics):
// entry.x = CTA-relative point index, entry.y = kQExact (pixel must come from the float64 formula) or the 1-based
// frame-relative index of the same-class slot holder that the point's atomicMax met (exact tie protocol needed).
// Doing this work inline would put the float64 atan2/asin (or the protocol) on the hot kernel's call graph and
// cost it ~25 registers per thread; k_range_queued / k_voxel_queued finish the queues instead.
constexpr uint32_t kQExact = 0xffffffffu;

template <typename T>
__device__ __forceinline__ void tile_issue(const T* __restrict__ xyz, const uint8_t* __restrict__ sem, int64_t base,
                                           unsigned char* smem, int st) {
  using L = TileLayout<T>;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::off_bar) + st;
  const uint64_t pol = l2_evict_first_policy();
  mbar_expect_tx(bar, (uint32_t)(L::xyz_bytes + kTile));
  bulk_g2s(smem + (size_t)st * L::xyz_bytes, xyz + 3 * base, (uint32_t)L::xyz_bytes, bar, pol);
  bulk_g2s(smem + L::off_sem + (size_t)st * kTile, sem + base, (uint32_t)kTile, bar, pol);
}

// Tiles [t0, t1) of CTA `cta` out of n_ctas for the points [p0, p1): absolute tile indices, contiguous runs.
__device__ __forceinline__ void cta_tile_range(int64_t p0, int64_t p1, int n_ctas, int cta, int* t0, int* t1) {
  const int64_t lo = p0 / kTile, hi = ceil_div64(p1, kTile);
  const int64_t per = ceil_div64(hi - lo, (int64_t)n_ctas);
  const int64_t a = lo + (int64_t)cta * per;
  *t0 = (int)(a < hi ? a : hi);
  *t1 = (int)(a + per < hi ? a + per : hi);
}

// Tile iterator of the point pass (all members are CTA-uniform).
template <typename T>
struct Tiles {
  unsigned char* smem;
  const T* xyz; const uint8_t* sem; const int64_t* off;
  int F; int64_t P0, P; bool vec_ok;   // points [P0, P) of the packed arrays are this launch's (a chunk of the batch)
  int t0, t1, t;           // this CTA's tiles [t0, t1) (absolute tile index = point index / kTile), current tile
  int f;                   // frame of the current tile's first point
  int64_t fbeg, fend;      // its [begin, end) point range
  int64_t base; bool full, one_frame; int st;
  uint32_t phase;          // bit s = parity of the next completion of stage s's mbarrier (flips per waited copy)

  __device__ __forceinline__ bool init(unsigned char* smem_, const T* xyz_, const uint8_t* sem_, const int64_t* off_, int F_,
                                       int64_t P0_, int64_t P_, bool vec_ok_) {
    smem = smem_; xyz = xyz_; sem = sem_; off = off_; F = F_; P0 = P0_; P = P_; vec_ok = vec_ok_;
    cta_tile_range(P0, P, (int)gridDim.x, (int)blockIdx.x, &t0, &t1);
    if (t0 >= t1) return false;
    using L = TileLayout<T>;
    if (threadIdx.x == 0) {
      uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::off_bar);
      mbar_init(bar, 1); mbar_init(bar + 1, 1);
      *reinterpret_cast<uint32_t*>(smem + L::off_qn) = 0u;
      mbar_fence_init();
    }
    __syncthreads();
    const int64_t first = (int64_t)t0 * kTile;
    f = find_frame(off, F, first > P0 ? first : P0);
    if (threadIdx.x == 0 && is_full(t0)) tile_issue<T>(xyz, sem, first, smem, 0);
    t = t0 - 1;
    phase = 0u;
    return true;
  }
  __device__ __forceinline__ bool is_full(int tt) const { return vec_ok && (int64_t)tt * kTile >= P0 && ((int64_t)tt + 1) * kTile <= P; }
  // advance to the next tile; returns false at the end.  The caller ends every tile with __syncthreads().
  __device__ __forceinline__ bool next() {
    ++t;
    if (t >= t1) return false;
    using L = TileLayout<T>;
    st = (t - t0) & 1;
    base = (int64_t)t * kTile;
    full = is_full(t);
    if (threadIdx.x == 0 && t + 1 < t1 && is_full(t + 1)) tile_issue<T>(xyz, sem, base + kTile, smem, st ^ 1);
    while (f + 1 < F && base >= __ldg(off + f + 1)) ++f;
    fbeg = __ldg(off + f); fend = __ldg(off + f + 1);
    one_frame = base >= fbeg && base + kTile <= fend;
    if (full) {   // (not every tile is staged: a chunk's first and last tiles may be partial)
      mbar_wait(reinterpret_cast<uint64_t*>(smem + L::off_bar) + st, (phase >> st) & 1u);
      phase ^= 1u << st;
    }
    return true;
  }
  // point j of the current tile
  __device__ __forceinline__ bool load(int j, T* x, T* y, T* z, uint32_t* lab) const {
    using L = TileLayout<T>;
    if (full) {
      const T* sx = reinterpret_cast<const T*>(smem + (size_t)st * L::xyz_bytes);
      *x = sx[3 * j]; *y = sx[3 * j + 1]; *z = sx[3 * j + 2];
      *lab = (smem + L::off_sem + (size_t)st * kTile)[j];
      return true;
    }
    const int64_t i = base + j;
    *x = *y = *z = (T)0; *lab = 0;
    if (i < P0 || i >= P) return false;
    *x = __ldg(xyz + 3 * i); *y = __ldg(xyz + 3 * i + 1); *z = __ldg(xyz + 3 * i + 2);
    *lab = __ldg(sem + i);
    return true;
  }
  // frame of point i of the current tile: (index, first point)
  __device__ __forceinline__ void frame_of(int64_t i, int* fk, int64_t* fb) const {
    *fk = f; *fb = fbeg;
    if (!one_frame && i >= fend) {
      int ff = f + 1;
      while (ff + 1 < F && i >= __ldg(off + ff + 1)) ++ff;
      *fk = ff; *fb = __ldg(off + ff);
    }
  }
};


// ---------------------------------------------------------------- K2: bitmap scan (+ K1q: range-image queue)
// One thread-block CLUSTER of 8 CTAs per frame: every CTA counts the set bits of its eighth of the frame bitmap,
// the eight totals are exchanged through distributed shared memory, and every CTA then writes the exclusive
// popcount prefix of its 128-bit chunks.  The same launch carries the CTAs that finish K1's rare-path queue
// (the two jobs are independent and both latency bound, so they share the machine instead of serialising).
constexpr int kScanCluster = 8;
constexpr int kScanThreads = 256;

__device__ __forceinline__ void scan_role(const uint32_t* __restrict__ bitmap, uint32_t* __restrict__ prefix, int gw,
                                          int64_t* __restrict__ n_occ_out) {
  __shared__ uint32_t totals[kScanCluster];
  __shared__ uint32_t warp_tot[kScanThreads / 32];
  __shared__ uint32_t carry_s;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int f = blockIdx.x / kScanCluster;
  const int per = (gw / 4) / kScanCluster;                       // gw % 32 == 0 -> chunks % 8 == 0
  const int c_begin = (int)rank * per, c_end = c_begin + per;
  const uint4* bm = reinterpret_cast<const uint4*>(bitmap + (size_t)f * gw);
  uint32_t* pf = prefix + (size_t)f * (gw / 4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // phase 1: set bits of this CTA's range
  uint32_t cnt = 0;
  for (int c = c_begin + tid; c < c_end; c += kScanThreads) {
    const uint4 v = bm[c];
    cnt += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) warp_tot[warp] = cnt;
  __syncthreads();
  if (tid < kScanCluster) {                                      // thread k posts this CTA's total into CTA k
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) tot += warp_tot[w];
    cluster.map_shared_rank(totals, tid)[rank] = tot;
  }
  cluster.sync();
  uint32_t base = 0, all = 0;
#pragma unroll
  for (int k = 0; k < kScanCluster; ++k) { const uint32_t t = totals[k]; all += t; if (k < (int)rank) base += t; }
  if (tid == 0) carry_s = base;
  __syncthreads();
  // phase 2: exclusive prefix per chunk (the second read of the range comes from L1/L2)
  constexpr int kPer = 4;   // chunks per thread per round (64 contiguous bytes)
  for (int c0r = c_begin; c0r < c_end; c0r += kScanThreads * kPer) {
    const int c0 = c0r + tid * kPer;
    uint32_t pc[kPer];
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      uint32_t c = 0;
      if (c0 + k < c_end) { const uint4 v = bm[c0 + k]; c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }
      pc[k] = c; tsum += c;
    }
    uint32_t incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    const uint32_t carry = carry_s;
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) if (w < warp) wbase += warp_tot[w];
    __syncthreads();
    if (tid == kScanThreads - 1) carry_s = carry + wbase + incl;
    uint32_t ex = carry + wbase + (incl - tsum);
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      if (c0 + k < c_end) pf[c0 + k] = ex;
      ex += pc[k];
    }
    __syncthreads();
  }
  if (rank == 0 && tid == 0 && n_occ_out) n_occ_out[f] = (int64_t)all;
}

// Rare-path queue entries.  x = CTA-relative point index, bit 31 set for a voxel event;
//   range : y = kQExact -> the reference's float64 formula decides the pixel and the atomicMax happens here;
//           otherwise y = 1-based frame-relative index of the same-class holder met by the point's atomicMax (the f32
//           pixel stands): exact tie protocol.
//   voxel : y = index of the same-class holder met on the voxel's word: exact tie protocol,
//           key = (not roadline, float64 |p mod res|^2).
constexpr uint32_t kQVoxel = 0x80000000u;
template <typename T>
struct QueueArgs {
  const T* xyz; const uint8_t* sem; const int64_t* off; int F; int64_t P0, P;
  u64* pixtab; u64* vtab; const uint2* queue; const uint32_t* qcount; int n_tile_ctas; int64_t* diag;
};

template <typename T>
__device__ __forceinline__ void queue_role(int cta, const QueueArgs<T>& a, const GridDev& g, const RangeDev& r) {
  if (cta >= a.n_tile_ctas) return;
  const uint32_t n = a.qcount[2 * cta];
  if (n == 0u) return;   // CTA-uniform
  int t0, t1;
  cta_tile_range(a.P0, a.P, a.n_tile_ctas, cta, &t0, &t1);
  const int64_t blk_first = (int64_t)t0 * kTile;
  const uint2* q = a.queue + 2 * (blk_first > a.P0 ? blk_first : a.P0);
  unsigned n_drop = 0, n_nw = 0, n_nh = 0;
  const int f_first = (int)a.qcount[2 * cta + 1];              // a CTA's points span few frames: walk from its first one
  for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
    const uint2 ent = q[e];
    const int64_t i = blk_first + (int64_t)(ent.x & ~kQVoxel);
    T x = __ldg(a.xyz + 3 * i), y = __ldg(a.xyz + 3 * i + 1), z = __ldg(a.xyz + 3 * i + 2);
    if (r.prep) lidar_prep(x, y, z, r);
    int f = f_first;
    while (f + 1 < a.F && i >= __ldg(a.off + f + 1)) ++f;
    const int64_t fb = __ldg(a.off + f);
    const T* fx = a.xyz + 3 * fb;
    const uint32_t me1 = (uint32_t)(i - fb) + 1u;
    if (ent.x & kQVoxel) {
      const uint8_t* fs = a.sem + fb;
      const bool packl = (__ldg(a.off + f + 1) - fb) < kPackLimit;
      auto key_of = [&](uint32_t q1) -> u64 {
        const T* qp = fx + 3 * (int64_t)(q1 - 1u);
        VoxKey o = vox_of<true>((double)__ldg(qp), (double)__ldg(qp + 1), (double)__ldg(qp + 2), g);
        return vox_key(o.dis, (int)__ldg(fs + (q1 - 1u)) != g.road);
      };
      const VoxKey v = vox_of<true>((double)x, (double)y, (double)z, g);
      const uint32_t top = key_top_inv(vox_key(v.dis, (int)__ldg(a.sem + i) != g.road));
      auto pack = [&](uint32_t q1) -> u64 { return vox_word(packl, top, q1, packl ? __ldg(fs + (q1 - 1u)) : 0u); };
      auto idx_of = [&](u64 wv) -> uint32_t { return vox_word_idx1(packl, wv); };
      tie_protocol(a.vtab + (size_t)f * g.G + v.bit, top, me1, pack(me1), pack(ent.y), key_of, pack, idx_of);
      continue;
    }
    double xc, yc, zc;
    const double s = range_sq_of(x, y, z, r, &xc, &yc, &zc);
    int pix;
    if (ent.y == kQExact) {
      const uint32_t pr = pix_refine(x, y, z, r);               // the side of the nearest edge decides (no atan2 / asin) ...
      if (pr != kPixUndecided) pix = (int)pr;
      else {                                                    // ... unless the point is within 1e-9 of it: the reference formula
        int iw, ih, flags;
        pix_exact(xc, yc, zc, s, r.H, r.W, r.fda, r.fov, &iw, &ih, &flags);
        if (flags & 4) { ++n_drop; continue; }
        n_nw += (flags & 1) ? 1u : 0u; n_nh += (flags & 2) ? 1u : 0u;
        pix = ih * r.W + iw;
      }
    } else {
      pix = pix_fast(x, y, z, r).pix;
    }
    u64* slot = a.pixtab + (size_t)f * r.H * r.W + pix;
    const uint32_t top = key_top_inv((u64)__double_as_longlong(s));
    const u64 mine = pack_word(top, me1);
    const u64 old_word = (ent.y == kQExact) ? atomicMax(slot, mine) : pack_word(top, ent.y);
    if (old_word != 0ull && word_top(old_word) == top) {
      auto key_of = [&](uint32_t q1) -> u64 {     // exact key: the float64 depth (geometry_utils.py:180)
        const T* qp = fx + 3 * (int64_t)(q1 - 1u);
        T qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        if (r.prep) lidar_prep(qx, qy, qz, r);
        double aa, bb, cc;
        return (u64)__double_as_longlong(sqrt(range_sq_of(qx, qy, qz, r, &aa, &bb, &cc)));
      };
      auto pack = [&](uint32_t q1) -> u64 { return pack_word(top, q1); };
      auto idx_of = [](u64 wv) -> uint32_t { return word_idx1(wv); };
      tie_protocol(slot, top, me1, mine, old_word, key_of, pack, idx_of);
    }
  }
  if (a.diag) {
    __syncwarp();
    diag_add(a.diag, MUVO_DIAG_DROPPED_NONFINITE, n_drop);
    diag_add(a.diag, MUVO_DIAG_NEAR_EDGE_W, n_nw);
    diag_add(a.diag, MUVO_DIAG_NEAR_EDGE_H, n_nh);
  }
}


// K2: grid = n_scan_frames * 8 scan CTAs (only when the sorted sparse list is wanted), then the queue CTAs (padded to a
// multiple of the cluster size)
template <typename T>
__global__ void __cluster_dims__(kScanCluster, 1, 1) __launch_bounds__(kScanThreads)
k_scan_queue(const uint32_t* __restrict__ bitmap, uint32_t* __restrict__ prefix, int gw, int n_scan_frames,
             int64_t* __restrict__ n_occ_out, const __grid_constant__ QueueArgs<T> qa, const __grid_constant__ GridDev g,
             const __grid_constant__ RangeDev r) {   // (by-value structs whose address is taken were copied to local memory:
                                                     //  408 bytes of stores per thread before the first useful instruction)
  pdl_wait();
  pdl_launch();
  const int scan_ctas = n_scan_frames * kScanCluster;
  if ((int)blockIdx.x < scan_ctas) scan_role(bitmap, prefix, gw, n_occ_out);
  else queue_role<T>((int)blockIdx.x - scan_ctas, qa, g, r);
}


// Scan-ordered clouds put runs of consecutive points (= adjacent lanes) into one voxel.  A lane whose neighbour in
// the same voxel holds a strictly better top-32 key class can never win the voxel: it skips its atomicMax.
template <bool ONE_FRAME>
__device__ __forceinline__ bool pair_filter(bool in, uint32_t bit /* 0xffffffff when !in */, int fk, uint32_t top_inv) {
  const unsigned lane = lane_id();
  const uint32_t bit_up = __shfl_up_sync(0xffffffffu, bit, 1), top_up = __shfl_up_sync(0xffffffffu, top_inv, 1);
  const uint32_t bit_dn = __shfl_down_sync(0xffffffffu, bit, 1), top_dn = __shfl_down_sync(0xffffffffu, top_inv, 1);
  bool beaten = (lane > 0 && bit == bit_up && top_up > top_inv) || (lane < 31 && bit == bit_dn && top_dn > top_inv);
  if (!ONE_FRAME) {
    const int fk_up = __shfl_up_sync(0xffffffffu, fk, 1), fk_dn = __shfl_down_sync(0xffffffffu, fk, 1);
    beaten = (lane > 0 && bit == bit_up && fk == fk_up && top_up > top_inv) ||
             (lane < 31 && bit == bit_dn && fk == fk_dn && top_dn > top_inv);
  }
  return in && !beaten;
}

template <typename T, bool DO_VOX, bool DO_RANGE, bool REG, bool PREP = false>
__global__ void __launch_bounds__(kTileThreads, MUVO_K1_MINB)
k_points_tile(const T* __restrict__ xyz, const uint8_t* __restrict__ sem, const int64_t* __restrict__ off, int F, int64_t P0,
              int64_t P, int f_lo, int f_hi, bool vec_ok, GridDev g, RangeDev r, uint32_t* __restrict__ bitmap, u64* __restrict__ vtab, u64* __restrict__ pixtab,
              uint2* __restrict__ queue, uint32_t* __restrict__ qcount, int64_t* __restrict__ n_occ_zero, int flags,
              int64_t* __restrict__ diag) {
  extern __shared__ __align__(128) unsigned char smem[];
  using L = TileLayout<T>;
  if (n_occ_zero && blockIdx.x == 0) for (int f = f_lo + threadIdx.x; f < f_hi; f += kTileThreads) n_occ_zero[f] = 0;
  Tiles<T> tl;
  if (!tl.init(smem, xyz, sem, off, F, P0, P, vec_ok)) {
    if (threadIdx.x == 0) qcount[2 * blockIdx.x] = 0u;
    return;
  }
  const int f_first = tl.f;                             // frame of this CTA's first point (the queue kernel starts its walks here)
  uint32_t* qn = reinterpret_cast<uint32_t*>(smem + L::off_qn);
  const int tid = threadIdx.x;
  const int64_t blk_first = (int64_t)tl.t0 * kTile;
  uint2* q = queue + 2 * (blk_first > P0 ? blk_first : P0);
  const int64_t HW = (int64_t)r.H * r.W;
  const bool use_filter = (flags & 1) != 0;
  // The voxel-table sectors a frame slot touches are scattered (one 32-byte sector per 1-4 occupied voxels) and are fetched
  // again by the dense emit, where DRAM serves such gathers at ~40 G sectors/s; marked evict_last they survive the streams
  // that flow past them (points in, 0.33 GB of outputs out) and, the workspace being reused, carry over to the next call for
  // every voxel that is occupied again (ground, static structure).  Measured on rotating batches: step 248 -> 239 us.
  const bool keep_vtab = (flags & 8) != 0;
  const uint64_t pol_keep = l2_evict_last_policy();
  unsigned n_drop = 0, n_in = 0;

  while (tl.next()) {
    // FAST = the tile is staged in shared memory and lies inside one frame (all but a handful of tiles): frame
    // bases, table rows and the shared-memory cursor are CTA-uniform and hoisted out of the per-point code.
    const bool packl_tile = (tl.fend - tl.fbeg) < kPackLimit;
    auto body = [&](auto fast_tag, auto packl_tag) {
      constexpr bool FAST = decltype(fast_tag)::value;
      constexpr bool PACKL_T = decltype(packl_tag)::value;          // FAST only: the tile's frame uses the label-carrying word
      const T* sx = reinterpret_cast<const T*>(smem + (size_t)tl.st * L::xyz_bytes) + 3 * tid;
      const uint8_t* ss = smem + L::off_sem + (size_t)tl.st * kTile + tid;
      uint32_t* bitmap_f = bitmap + (size_t)tl.f * g.gw;
      u64* vtab_f = vtab + (size_t)tl.f * g.G;
      u64* pixtab_f = pixtab + (size_t)tl.f * HW;
      // keep the three table rows in registers (the compiler would otherwise re-derive them from f for every point)
      asm volatile("" : "+l"(bitmap_f), "+l"(vtab_f), "+l"(pixtab_f));
      const bool packl_f = FAST ? PACKL_T : packl_tile;
      const uint32_t idx1_0 = (uint32_t)(tl.base - tl.fbeg) + (uint32_t)tid + 1u;     // 1-based frame-relative index, k = 0
      const uint32_t rel_0 = (uint32_t)(tl.base - blk_first) + (uint32_t)tid;         // CTA-relative index, k = 0
      // results of the previous point's atomics, looked at one point later so that their latency is covered by the
      // next point's arithmetic
      struct Pending { u64 vold, pold; uint32_t bit, vtop, ptop, rel; int fk; bool vgo, pgo, packl; };
      auto settle = [&](const Pending& d) {
        if (DO_VOX) {
          if (d.vold == 0ull)                                               // first claim of the voxel marks it occupied
            red_or_global((FAST ? bitmap_f : bitmap + (size_t)d.fk * g.gw) + (d.bit >> 5), 1u << (d.bit & 31));
          else if (d.vgo && word_top(d.vold) == d.vtop)                     // same top-32 class: exact protocol
            q[atomicAdd(qn, 1u)] = make_uint2(d.rel | kQVoxel, vox_word_idx1(d.packl, d.vold));
        }
        if (DO_RANGE) {
          if (d.pgo && d.pold != 0ull && word_top(d.pold) == d.ptop) q[atomicAdd(qn, 1u)] = make_uint2(d.rel, word_idx1(d.pold));
        }
      };
      if (FAST) {
        // Two phases.  (1) arithmetic for the thread's 4 points; the operands of their atomics are kept (table offset +
        // word, ~0 = none).  (2) all 8 atomics are issued back to back and only then are their results looked at: the
        // round trips to L2 / DRAM overlap each other instead of stalling the thread once per point.
        uint32_t vbit[kKPL], ppix[kKPL];
        u64 vword[kKPL], pword[kKPL];
#pragma unroll
        for (int k = 0; k < kKPL; ++k) {
          T x = sx[3 * k * kTileThreads], y = sx[3 * k * kTileThreads + 1], z = sx[3 * k * kTileThreads + 2];
          const uint32_t lab = ss[k * kTileThreads];
          const uint32_t me1 = idx1_0 + k * kTileThreads;
          vbit[k] = 0xffffffffu; ppix[k] = 0xffffffffu; vword[k] = 0ull; pword[k] = 0ull;
          const bool keep = PREP ? lidar_prep(x, y, z, r) : true;                   // N1: raw LiDAR-frame points (range-only calls)
          if (DO_VOX) {
            bool in; double dis; uint32_t bit = 0xffffffffu;
            if (REG) { VoxFast v = vox_regular(x, y, z, g); in = v.in; dis = vox_regular_dis(v, g); if (in) bit = v.bit; }
            else { VoxKey v = vox_of<true>((double)x, (double)y, (double)z, g); in = v.in; dis = v.dis; if (in) bit = v.bit; }
            const uint32_t top = key_top_inv(vox_key(dis, (int)lab != g.road));
            n_in += in ? 1u : 0u;
            const bool go = use_filter ? pair_filter<true>(in, bit, 0, top) : in;
            if (go) { vbit[k] = bit; vword[k] = vox_word(PACKL_T, top, me1, lab); }
          }
          if (DO_RANGE && keep) {
            const PixFast pk = pix_fast(x, y, z, r);
            if (!pk.ok) ++n_drop;
            else if (pk.slow) q[atomicAdd(qn, 1u)] = make_uint2(rel_0 + k * kTileThreads, kQExact);
            else {
              ppix[k] = (uint32_t)pk.pix;
              // s > 0: its bit pattern orders like the value, and like the depth sqrt(s)
              pword[k] = pack_word(key_top_inv((u64)__double_as_longlong(pk.s)), me1);
            }
          }
        }
        u64 vold[kKPL], pold[kKPL];
#pragma unroll
        for (int k = 0; k < kKPL; ++k) {
          vold[k] = 1ull; pold[k] = 0ull;
#ifdef MUVO_TIMING_KNOBS
          // timing experiments only (results are WRONG): flags bit 1 = no table atomics at all, bit 2 = red.max (no return value)
          if (flags & 6) {
            if (flags & 4) {
              if (DO_VOX && vbit[k] != 0xffffffffu) asm volatile("red.global.max.u64 [%0], %1;" ::"l"(vtab_f + vbit[k]), "l"(vword[k]) : "memory");
              if (DO_RANGE && ppix[k] != 0xffffffffu) asm volatile("red.global.max.u64 [%0], %1;" ::"l"(pixtab_f + ppix[k]), "l"(pword[k]) : "memory");
            }
            n_drop += (unsigned)((vword[k] ^ pword[k] ^ vbit[k] ^ ppix[k]) == 0x123456789abcdefull);
            vold[k] = 1ull; pold[k] = 0ull;
            continue;
          }
#endif
          if (DO_VOX && vbit[k] != 0xffffffffu) vold[k] = keep_vtab ? atom_max_global_hint(vtab_f + vbit[k], vword[k], pol_keep) : atom_max_global(vtab_f + vbit[k], vword[k]);
          if (DO_RANGE && ppix[k] != 0xffffffffu) pold[k] = atom_max_global(pixtab_f + ppix[k], pword[k]);
        }
#pragma unroll
        for (int k = 0; k < kKPL; ++k) {
          if (DO_VOX) {
            if (vold[k] == 0ull)                                            // first claim of the voxel marks it occupied
              red_or_global(bitmap_f + (vbit[k] >> 5), 1u << (vbit[k] & 31));
            else if (vbit[k] != 0xffffffffu && word_top(vold[k]) == word_top(vword[k]))   // same top-32 class: exact protocol
              q[atomicAdd(qn, 1u)] = make_uint2((rel_0 + k * kTileThreads) | kQVoxel, vox_word_idx1(PACKL_T, vold[k]));
          }
          if (DO_RANGE) {
            if (ppix[k] != 0xffffffffu && pold[k] != 0ull && word_top(pold[k]) == word_top(pword[k]))
              q[atomicAdd(qn, 1u)] = make_uint2(rel_0 + k * kTileThreads, word_idx1(pold[k]));
          }
        }
        return;
      }
      Pending prev{1ull, 0ull, 0u, 0u, 0u, 0u, 0, false, false, false};
#pragma unroll
      for (int k = 0; k < kKPL; ++k) {
        T x, y, z; uint32_t lab;
        bool valid = true;
        int fk = tl.f; int64_t fb = tl.fbeg;
        bool packl = packl_f;
        if (FAST) {
          x = sx[3 * k * kTileThreads]; y = sx[3 * k * kTileThreads + 1]; z = sx[3 * k * kTileThreads + 2];
          lab = ss[k * kTileThreads];
        } else {
          valid = tl.load(k * kTileThreads + tid, &x, &y, &z, &lab);
          tl.frame_of(tl.base + k * kTileThreads + tid, &fk, &fb);
          packl = (__ldg(off + fk + 1) - fb) < kPackLimit;
        }
        if (PREP && valid) valid = lidar_prep(x, y, z, r);
        const uint32_t me1 = FAST ? idx1_0 + k * kTileThreads : (uint32_t)(tl.base + k * kTileThreads + tid - fb) + 1u;
        Pending cur{1ull, 0ull, 0xffffffffu, 0u, 0u, rel_0 + k * kTileThreads, fk, false, false, packl};
        if (DO_VOX) {
          bool in = false;
          uint32_t top = 0u;
          if (valid) {
            double dis;
            if (REG) { VoxFast v = vox_regular(x, y, z, g); in = v.in; dis = vox_regular_dis(v, g); if (in) cur.bit = v.bit; }
            else { VoxKey v = vox_of<true>((double)x, (double)y, (double)z, g); in = v.in; dis = v.dis; if (in) cur.bit = v.bit; }
            top = key_top_inv(vox_key(dis, (int)lab != g.road));
          }
          n_in += in ? 1u : 0u;
          cur.vgo = use_filter ? pair_filter<FAST>(in, cur.bit, fk, top) : in;
          cur.vtop = top;
          if (cur.vgo) cur.vold = atom_max_global((FAST ? vtab_f : vtab + (size_t)fk * g.G) + cur.bit, vox_word(packl, top, me1, lab));
        }
        if (DO_RANGE && valid) {
          const PixFast pk = pix_fast(x, y, z, r);
          if (!pk.ok) ++n_drop;
          else if (pk.slow) q[atomicAdd(qn, 1u)] = make_uint2(cur.rel, kQExact);
          else {
            cur.pgo = true;
            // s > 0: its bit pattern orders like the value, and like the depth sqrt(s)
            cur.ptop = key_top_inv((u64)__double_as_longlong(pk.s));
            cur.pold = atom_max_global((FAST ? pixtab_f : pixtab + (size_t)fk * HW) + pk.pix, pack_word(cur.ptop, me1));
          }
        }
        if (k > 0) settle(prev);
        prev = cur;
      }
      settle(prev);
    };
    if (tl.full && tl.one_frame) {
      if (packl_tile) body(std::true_type{}, std::true_type{}); else body(std::true_type{}, std::false_type{});
    } else {
      body(std::false_type{}, std::false_type{});
    }
    __syncthreads();                         // tile buffer free for the copy issued by the next next()
  }
  pdl_launch();
  if (tid == 0) { qcount[2 * blockIdx.x] = *qn; qcount[2 * blockIdx.x + 1] = (uint32_t)f_first; }
  if (diag) {
    if (DO_RANGE) diag_add(diag, MUVO_DIAG_DROPPED_NONFINITE, n_drop);
    if (DO_VOX) diag_add(diag, MUVO_DIAG_IN_GRID, n_in);
  }
}

// ---------------------------------------------------------------- K5: emit
// Shared by the sparse emit: lane L owns bitmap word (warp_word0 + L); returns the word and the rank of
// its first bit.  Ranks of the 4 words of a chunk are built with shuffles, so no lane ever re-reads a
// word that its owner may already have cleared.
__device__ __forceinline__ void load_word_and_rank(uint32_t* __restrict__ bitmap, uint32_t* __restrict__ prefix,
                                                   int64_t word_global, bool valid, bool clean, uint32_t* word_o,
                                                   uint32_t* rank_o) {
  uint32_t word = valid ? bitmap[word_global] : 0u;
  uint32_t base = valid ? prefix[word_global >> 2] : 0u;
  // the prefix table is cleared as well (the 4 lanes of a chunk read the entry in the same instruction; its
  // first lane clears it afterwards)
  if (clean && valid && base && (lane_id() & 3u) == 0u) prefix[word_global >> 2] = 0u;
  uint32_t pc = __popc(word);
  unsigned lane = lane_id();
  uint32_t p1 = __shfl_up_sync(0xffffffffu, pc, 1);
  uint32_t p2 = __shfl_up_sync(0xffffffffu, pc, 2);
  uint32_t p3 = __shfl_up_sync(0xffffffffu, pc, 3);
  unsigned k = lane & 3u;
  uint32_t rank = base + (k >= 1 ? p1 : 0u) + (k >= 2 ? p2 : 0u) + (k >= 3 ? p3 : 0u);
  *word_o = word; *rank_o = rank;
}


struct EmitDenseArgs {
  uint32_t* bitmap; u64* vtab; const int64_t* off; const uint8_t* sem; const uint8_t* remap; uint8_t* dense; int64_t* n_occ;
  int f_lo;   // first frame of this launch (grid.y counts frames from here)
  int wait_prev;   // 0 when the previous kernel of the stream is k_emit_range: that one triggers this launch only after its own
                   // pdl_wait (= K1 and K2 complete) and touches none of this kernel's buffers, so the two may overlap
};
// CTA `bx` of frame `f`
__device__ __forceinline__ void emit_dense_body(int bx, int f, EmitSmem& sm, const EmitDenseArgs& a, const GridDev& g) {
  uint32_t* __restrict__ bitmap = a.bitmap; u64* __restrict__ vtab = a.vtab; const int64_t* __restrict__ off = a.off;
  const uint8_t* __restrict__ sem = a.sem; const uint8_t* __restrict__ remap = a.remap; uint8_t* __restrict__ dense = a.dense;
  int64_t* __restrict__ n_occ = a.n_occ;
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  // lane L takes words L, L+32, L+64, L+96 of the warp's 128, so that every load / store instruction of the warp is
  // contiguous; row r = k*32 + L of the tile belongs to word (warp_w0 + r)
  const uint32_t warp_w0 = ((uint32_t)bx * (kBlock / 32) + warp) * (32 * kEmitWords);
  uint32_t* bm = bitmap + (size_t)f * g.gw;
  uint32_t bits[kEmitWords];
  uint32_t pc = 0;
#pragma unroll
  for (int k = 0; k < kEmitWords; ++k) {
    const uint32_t wi = warp_w0 + 32u * k + lane;
    bits[k] = wi < (uint32_t)g.gw ? bm[wi] : 0u;
    pc += __popc(bits[k]);
  }
  uint8_t* tile = sm.tile[warp];
#pragma unroll
  for (int k = 0; k < kEmitWords; ++k) {                        // zero this lane's rows
    uint4* row = reinterpret_cast<uint4*>(tile + (k * 32 + lane) * 32);
    row[0] = make_uint4(0u, 0u, 0u, 0u); row[1] = make_uint4(0u, 0u, 0u, 0u);
    if (bits[k]) bm[warp_w0 + 32u * k + lane] = 0u;
  }
  uint32_t incl = pc;                                           // position of this lane's set bits in the warp's list
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (n_occ) {                                                  // occupied voxels of the frame: one atomic per CTA
    if (lane == 0) sm.occ[warp] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
#pragma unroll
      for (int k = 0; k < kBlock / 32; ++k) t += sm.occ[k];
      if (t) atomicAdd(reinterpret_cast<unsigned long long*>(n_occ + f), (unsigned long long)t);
    }
  }
  if (warp_w0 >= (uint32_t)g.gw) return;                        // whole warp out of range (gw % 32 == 0)
  if (total) {                                                  // warp-uniform
    const int64_t fb = __ldg(off + f);
    const bool packl = (__ldg(off + f + 1) - fb) < kPackLimit;
    const uint8_t* sem_f = sem + fb;
    u64* vt = vtab + (size_t)f * g.G + (size_t)warp_w0 * 32;
    uint16_t* list = sm.list[warp];
    // the set bits of the whole span are gathered cooperatively: lane t takes the t-th, t+32-th, ... set bit, so all
    // gathers of a round are in flight together no matter how they are spread over the words
    for (uint32_t done = 0; done < total; done += kEmitList) {  // rounds only for spans with > kEmitList set bits
      uint32_t n = incl - pc;                                   // list index of this lane's first set bit
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
        const u64 wv = vt[pos];
        uint32_t lab = vox_word_label(packl, wv, sem_f);
        if (remap) lab = __ldg(remap + lab);
        tile[tile_swz(pos)] = (uint8_t)lab;
      }
      __syncwarp();
    }
    // clear the winners, one 32-byte sector (4 voxels) per store, now that every fetch of the span has been consumed
#pragma unroll
    for (int k = 0; k < kEmitWords; ++k) {
      uint32_t b = bits[k];
      u64* pw = vt + (k * 32 + lane) * 32;
      while (b) {
        const int sct = (__ffs(b) - 1) >> 2;
        b &= ~(0xfu << (4 * sct));
        st_zero_sector(pw + 4 * sct);
      }
    }
  }
  uint8_t* df = dense + (size_t)f * g.G;
  const bool aligned = ((reinterpret_cast<uintptr_t>(df) & 31) == 0);
#pragma unroll
  for (int k = 0; k < kEmitWords; ++k) {
    const uint32_t r = k * 32 + lane, wi = warp_w0 + r;
    if (wi >= (uint32_t)g.gw) continue;
    const uint32_t sw = (r >> 2) & 1u;                          // this row's halves are swapped
    const uint4 h0 = *reinterpret_cast<const uint4*>(tile + r * 32 + 16 * sw);
    const uint4 h1 = *reinterpret_cast<const uint4*>(tile + r * 32 + 16 * (sw ^ 1u));
    const uint32_t o[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    const int64_t vox = (int64_t)wi * 32;                       // first voxel of the word inside the frame
    uint8_t* dst = df + vox;
    if (vox + 32 <= g.G && aligned) {
      st_stream_u8x32(dst, o);
    } else {
      for (int j = 0; j < 32; ++j)
        if (vox + j < g.G) dst[j] = (uint8_t)(o[j >> 2] >> (8 * (j & 3)));
    }
  }
}

__global__ void __launch_bounds__(kBlock)
k_emit_dense(EmitDenseArgs a, GridDev g) {
  extern __shared__ __align__(16) unsigned char emit_raw[];
  if (a.wait_prev) pdl_wait();
  emit_dense_body((int)blockIdx.x, a.f_lo + (int)blockIdx.y, *reinterpret_cast<EmitSmem*>(emit_raw), a, g);
}

// Packed sparse output: start[f] = number of occupied voxels in frames < f (start[F] = total).  One CTA.
__global__ void __launch_bounds__(1024)
k_frame_prefix(const int64_t* __restrict__ n_occ, int F, int64_t* __restrict__ start) {
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < F; base += 1024) {
    const int i = base + threadIdx.x;
    const int64_t v = i < F ? n_occ[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    const int64_t carry = carry_s;
    int64_t wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += wsum[w];
    __syncthreads();
    if (i < F) start[i] = carry + wbase + incl - v;
    if (threadIdx.x == 1023) carry_s = carry + wbase + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[F] = carry_s;
}

// Sparse list, bitmap in linear-id order: rows (x,y,z,label) uint16 at sparse[(row0[f] + rank)], row0 = frame_offsets
// (frame f's rows start where its points start) or `start` (packed: frames back to back).
// sparse == nullptr: only clears the tables (n_occ-only calls).
// One warp = 32 consecutive bitmap words.  The 32 voxels of a word are contiguous in the voxel table (bit index =
// linear id here), so a non-empty word is finished by the whole warp at once: lane j fetches (and clears) the winner
// of bit j with one coalesced access and writes its row at rank(word) + popc(bits below j).  Two words per round keep
// two independent fetches in flight.  (One thread per word with a loop over its bits was 700 us at cfg2: ground rows
// fill whole words, i.e. 32 dependent DRAM round trips per thread.)
__global__ void __launch_bounds__(kBlock)
k_emit_sparse(uint32_t* __restrict__ bitmap, uint32_t* __restrict__ prefix, u64* __restrict__ vtab,
              const int64_t* __restrict__ off, const uint8_t* __restrict__ sem, uint16_t* __restrict__ sparse,
              const int64_t* __restrict__ start, GridDev g, int F) {
  const int64_t wg = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const int64_t total = (int64_t)F * g.gw;
  const bool valid = wg < total;
  uint32_t word, rank;
  load_word_and_rank(bitmap, prefix, wg, valid, true, &word, &rank);
  if (valid && word) bitmap[wg] = 0u;
  unsigned todo = __ballot_sync(0xffffffffu, valid && word != 0u);
  if (!todo) return;                                              // warp-uniform
  const unsigned lane = lane_id();
  const int64_t warp_wg = wg - lane;                              // gw % 32 == 0: the warp's words belong to one frame
  const int f = (int)(warp_wg / g.gw);
  const int64_t fbeg = __ldg(off + f);
  const bool packl = (__ldg(off + f + 1) - fbeg) < kPackLimit;
  const uint8_t* sem_f = sem + fbeg;
  const int64_t row0 = start ? __ldg(start + f) : fbeg;
  const uint32_t bit_w0 = (uint32_t)(warp_wg - (int64_t)f * g.gw) * 32u;   // linear id of the warp's first voxel
  u64* vt = vtab + (size_t)f * g.G + bit_w0;
  while (todo) {
    const int s0 = __ffs(todo) - 1; todo &= todo - 1;
    const int s1 = todo ? __ffs(todo) - 1 : s0; const bool two = todo != 0u; todo &= todo - 1;
    const uint32_t w0 = __shfl_sync(0xffffffffu, word, s0), w1 = __shfl_sync(0xffffffffu, word, s1);
    const uint32_t r0 = __shfl_sync(0xffffffffu, rank, s0), r1 = __shfl_sync(0xffffffffu, rank, s1);
    const bool h0 = (w0 >> lane) & 1u, h1 = two && ((w1 >> lane) & 1u);
    u64 e0 = 0ull, e1 = 0ull;
    if (h0) e0 = vt[s0 * 32 + lane];
    if (h1) e1 = vt[s1 * 32 + lane];
    if (sparse) {
      const uint32_t below = (1u << lane) - 1u;
      if (h0) {
        const uint32_t lin = bit_w0 + (uint32_t)s0 * 32u + lane;
        const uint32_t x = lin % (uint32_t)g.dx, yz = lin / (uint32_t)g.dx;
        const uint32_t y = yz % (uint32_t)g.dy, z = yz / (uint32_t)g.dy;
        const uint32_t lab = vox_word_label(packl, e0, sem_f);
        *reinterpret_cast<uint2*>(sparse + (size_t)(row0 + r0 + __popc(w0 & below)) * 4) = make_uint2(x | (y << 16), z | (lab << 16));
      }
      if (h1) {
        const uint32_t lin = bit_w0 + (uint32_t)s1 * 32u + lane;
        const uint32_t x = lin % (uint32_t)g.dx, yz = lin / (uint32_t)g.dx;
        const uint32_t y = yz % (uint32_t)g.dy, z = yz / (uint32_t)g.dy;
        const uint32_t lab = vox_word_label(packl, e1, sem_f);
        *reinterpret_cast<uint2*>(sparse + (size_t)(row0 + r1 + __popc(w1 & below)) * 4) = make_uint2(x | (y << 16), z | (lab << 16));
      }
    }
    // clear the sectors (4 entries = 32 bytes) that held winners: lanes 0-7 own the 8 sectors of word s0, lanes 8-15
    // those of s1.  The shuffles below also order the stores behind the use of e0 / e1.
    const uint32_t used0 = __ballot_sync(0xffffffffu, h0 && e0 != 0ull), used1 = __ballot_sync(0xffffffffu, h1 && e1 != 0ull);
    if (lane < 16) {
      const uint32_t u = lane < 8 ? used0 : used1;
      const int sw = lane < 8 ? s0 : s1, k = lane & 7;
      if ((u >> (4 * k)) & 0xfu) st_zero_sector(vt + sw * 32 + 4 * k);
    }
  }
}

// Dense grid when the bitmap is in linear-id order (both outputs requested): per-voxel lookup, nothing is cleared
// (k_emit_sparse runs afterwards).
__global__ void __launch_bounds__(kBlock)
k_emit_dense_from_linear(const uint32_t* __restrict__ bitmap, const u64* __restrict__ vtab, const int64_t* __restrict__ off,
                         const uint8_t* __restrict__ sem, const uint8_t* __restrict__ remap, uint8_t* __restrict__ dense,
                         GridDev g, int F) {
  int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (t >= (int64_t)F * g.G) return;
  int f = (int)(t / g.G);
  uint32_t d = (uint32_t)(t - (int64_t)f * g.G);
  uint32_t z = d % (uint32_t)g.dz, xy = d / (uint32_t)g.dz;
  uint32_t y = xy % (uint32_t)g.dy, x = xy / (uint32_t)g.dy;
  uint32_t lin = x + (uint32_t)g.dx * (y + (uint32_t)g.dy * z);
  const uint32_t* bm = bitmap + (size_t)f * g.gw;
  uint32_t lab = 0;
  if ((bm[lin >> 5] >> (lin & 31)) & 1u) {
    const int64_t fb = __ldg(off + f);
    const bool packl = (__ldg(off + f + 1) - fb) < kPackLimit;
    lab = vox_word_label(packl, vtab[(size_t)f * g.G + lin], sem + fb);
    if (remap) lab = __ldg(remap + lab);
  }
  dense[t] = (uint8_t)lab;
}

// Range image: one thread per NP pixels (NP = 4: 16-byte stores; NP = 1: generic fallback).
template <typename T>
struct EmitRangeArgs {
  u64* pixtab; const T* xyz; const uint8_t* sem; const int64_t* off; int f_lo, f_hi;   // frames [f_lo, f_hi) of this launch
  float* depth_out; float* xyz_out; uint8_t* sem_out;
  const uint8_t* sem_remap;   // N1: 256-entry label remap applied to the winner's tag (dataset.py:281-283), or nullptr
};
template <typename T, int NP, int LAYOUT>
__device__ __forceinline__ void emit_range_body(int64_t vblock, const EmitRangeArgs<T>& a, const RangeDev& r) {
  u64* __restrict__ pixtab = a.pixtab; const T* __restrict__ xyz = a.xyz; const uint8_t* __restrict__ sem = a.sem;
  const int64_t* __restrict__ off = a.off;
  float* __restrict__ depth_out = a.depth_out; float* __restrict__ xyz_out = a.xyz_out; uint8_t* __restrict__ sem_out = a.sem_out;
  const uint8_t* __restrict__ sem_remap = a.sem_remap;
  constexpr bool clean = true;
  const int64_t HW = (int64_t)r.H * r.W;
  int64_t t = vblock * kBlock + threadIdx.x;
  int64_t p0 = (int64_t)a.f_lo * HW + t * NP;   // global pixel index over [F, H*W]
  if (p0 >= (int64_t)a.f_hi * HW) return;
  int f = (int)(p0 / HW);
  int64_t pin = p0 - (int64_t)f * HW;  // pixel inside the frame
  int64_t fbeg = __ldg(off + f);
  u64 wv[NP];
  if (NP == 4) {
    ulonglong2 a = *reinterpret_cast<ulonglong2*>(pixtab + p0);
    ulonglong2 b = *reinterpret_cast<ulonglong2*>(pixtab + p0 + 2);
    wv[0] = a.x; wv[1 % NP] = a.y; wv[2 % NP] = b.x; wv[3 % NP] = b.y;
    if (clean) {   // (a single 32-byte clear after the gathers was measured slower here: 68 vs 61 us)
      if (a.x | a.y) *reinterpret_cast<ulonglong2*>(pixtab + p0) = make_ulonglong2(0ull, 0ull);
      if (b.x | b.y) *reinterpret_cast<ulonglong2*>(pixtab + p0 + 2) = make_ulonglong2(0ull, 0ull);
    }
  } else {
    wv[0] = pixtab[p0];
    if (clean && wv[0]) pixtab[p0] = 0ull;
  }
  float px[NP], py[NP], pz[NP], pd[NP];
  uint32_t ps = 0;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    px[k] = py[k] = pz[k] = 0.f; pd[k] = -1.f;                 // :210-212 initial values
    if (wv[k]) {
      int64_t qi = fbeg + (int64_t)(word_idx1(wv[k]) - 1u);
      T x = __ldg(xyz + 3 * qi), y = __ldg(xyz + 3 * qi + 1), z = __ldg(xyz + 3 * qi + 2);
      if (r.prep) lidar_prep(x, y, z, r);
      double a, b, c;
      pd[k] = (float)range_depth_of(x, y, z, r, &a, &b, &c);   // :217 float32(depth64)
      px[k] = (float)x; py[k] = (float)y; pz[k] = (float)z;    // :218 the ego-frame input point
      uint32_t sv = __ldg(sem + qi);
      if (sem_remap) sv = __ldg(sem_remap + sv);
      ps |= sv << (8 * k);                                     // :219
    }
  }
  if (NP == 4) {
    if (depth_out) st_stream_f4(reinterpret_cast<float4*>(depth_out + p0), make_float4(pd[0], pd[1 % NP], pd[2 % NP], pd[3 % NP]));
    if (LAYOUT == MUVO_RANGE_LAYOUT_HWC) {
      float4* o = reinterpret_cast<float4*>(xyz_out + p0 * 3);
      st_stream_f4(o,     make_float4(px[0], py[0], pz[0], px[1 % NP]));
      st_stream_f4(o + 1, make_float4(py[1 % NP], pz[1 % NP], px[2 % NP], py[2 % NP]));
      st_stream_f4(o + 2, make_float4(pz[2 % NP], px[3 % NP], py[3 % NP], pz[3 % NP]));
    } else {
      float* base = xyz_out + (size_t)f * 4 * HW + pin;
      st_stream_f4(reinterpret_cast<float4*>(base),          make_float4(px[0], px[1 % NP], px[2 % NP], px[3 % NP]));
      st_stream_f4(reinterpret_cast<float4*>(base + HW),     make_float4(py[0], py[1 % NP], py[2 % NP], py[3 % NP]));
      st_stream_f4(reinterpret_cast<float4*>(base + 2 * HW), make_float4(pz[0], pz[1 % NP], pz[2 % NP], pz[3 % NP]));
      st_stream_f4(reinterpret_cast<float4*>(base + 3 * HW), make_float4(pd[0], pd[1 % NP], pd[2 % NP], pd[3 % NP]));
    }
    if (sem_out) st_stream_u32(reinterpret_cast<uint32_t*>(sem_out + p0), ps);
  } else {
    if (depth_out) depth_out[p0] = pd[0];
    if (LAYOUT == MUVO_RANGE_LAYOUT_HWC) {
      xyz_out[p0 * 3] = px[0]; xyz_out[p0 * 3 + 1] = py[0]; xyz_out[p0 * 3 + 2] = pz[0];
    } else {
      float* base = xyz_out + (size_t)f * 4 * HW + pin;
      base[0] = px[0]; base[HW] = py[0]; base[2 * HW] = pz[0]; base[3 * HW] = pd[0];
    }
    if (sem_out) sem_out[p0] = (uint8_t)ps;
  }
}

#ifndef MUVO_ER_MINB
#define MUVO_ER_MINB 1
#endif
template <typename T, int NP, int LAYOUT>
__global__ void __launch_bounds__(kBlock, MUVO_ER_MINB)
k_emit_range(EmitRangeArgs<T> a, RangeDev r) {
  pdl_wait();
  pdl_launch();
  emit_range_body<T, NP, LAYOUT>((int64_t)blockIdx.x, a, r);
}

// ---------------------------------------------------------------- test hook: f32 pixel path vs float64 formula
__global__ void __launch_bounds__(kBlock)
k_debug_pixel_check(const float* __restrict__ xyz, int64_t n, RangeDev r, unsigned long long* __restrict__ out) {
  unsigned n_fast = 0, n_slow = 0, n_bad = 0, n_drop = 0;
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const float x = __ldg(xyz + 3 * i), y = __ldg(xyz + 3 * i + 1), z = __ldg(xyz + 3 * i + 2);
    const PixFast pk = pix_fast(x, y, z, r);
    if (!pk.ok) { ++n_drop; continue; }
    if (pk.slow) { ++n_slow; continue; }
    double xc, yc, zc;
    const double s = range_sq_of(x, y, z, r, &xc, &yc, &zc);
    int iw, ih, flags;
    pix_exact(xc, yc, zc, s, r.H, r.W, r.fda, r.fov, &iw, &ih, &flags);
    ++n_fast;
    if ((flags & 4) || pk.pix != ih * r.W + iw) ++n_bad;
  }
  unsigned v[4] = {n_fast, n_slow, n_bad, n_drop};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned tot = __reduce_add_sync(0xffffffffu, v[k]);
    if (tot && lane_id() == 0) atomicAdd(out + k, (unsigned long long)tot);
  }
}


static inline unsigned blocks_for(int64_t n) { return (unsigned)ceil_div64(n, kBlock); }

template <typename T>
static int run_points(const T* xyz, const uint8_t* sem, const int64_t* off, int F, int64_t P,
                      const MuvoGrid* grid_h, const uint8_t* remap, const MuvoRangeCfg* cfg_h, int layout, uint8_t* dense,
                      uint16_t* sparse, int64_t* n_occ, int64_t* sparse_start, float* depth_out, float* xyz_out,
                      uint8_t* sem_out, int64_t* diag, void* ws, size_t ws_bytes, cudaStream_t st,
                      const MuvoLidarPrep* prep_h = nullptr) {
  const bool do_vox = grid_h != nullptr, do_range = cfg_h != nullptr;
  if (prep_h && (do_vox || !do_range || !std::is_same<T, float>::value)) return MUVO_E_ARG;   // N1 prep: float32, range-only calls
  if (!do_vox && !do_range) return MUVO_E_ARG;
  if (F < 0 || P < 0) return MUVO_E_ARG;
  if (F == 0) return MUVO_OK;
  if (F > 65535) return MUVO_E_SHAPE;   // frames are a grid dimension of the emit kernels
  if (!off || !ws) return MUVO_E_NULL;
  if (P > 0 && (!xyz || !sem)) return MUVO_E_NULL;
  if (do_vox && !dense && !sparse && !n_occ) return MUVO_E_NULL;
  if (sparse_start && (!sparse || !n_occ)) return MUVO_E_NULL;   // the packed layout needs the per-frame counts
  if (do_range && (!xyz_out || (layout == MUVO_RANGE_LAYOUT_HWC && (!depth_out || !sem_out)))) return MUVO_E_NULL;
  if (layout != MUVO_RANGE_LAYOUT_HWC && layout != MUVO_RANGE_LAYOUT_XYZD) return MUVO_E_ARG;
  if (reinterpret_cast<uintptr_t>(ws) & 255) return MUVO_E_ALIGN;
  if (P >= ((int64_t)1 << 36)) return MUVO_E_SHAPE;              // rare-path queue entries are 32-bit CTA-relative
  GridDev g{}; RangeDev r{};
  int rc;
  const int order = (sparse != nullptr) ? ORDER_LINEAR : ORDER_DENSE;   // bit index = output order of the list that is emitted
  if (do_vox && (rc = make_grid_dev(grid_h, order, &g)) != MUVO_OK) return rc;
  if (do_range && (rc = make_range_dev(cfg_h, &r)) != MUVO_OK) return rc;
  if (prep_h) {
    r.prep = 1; r.prep_box = prep_h->use_ego_box ? 1 : 0;
    for (int k = 0; k < 3; ++k) {
      if (!isfinite(prep_h->add[k])) return MUVO_E_ARG;
      r.padd[k] = prep_h->add[k]; r.blo[k] = prep_h->box_lo[k]; r.bhi[k] = prep_h->box_hi[k];
    }
  }
  PointsWs w = carve(ws, P, F, grid_h, cfg_h);
  if (w.bytes > ws_bytes) return MUVO_E_WORKSPACE;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(xyz) % 16 == 0) && (reinterpret_cast<uintptr_t>(sem) % 16 == 0);
  const size_t tsmem = TileLayout<T>::bytes;
  const int64_t HWr = do_range ? (int64_t)r.H * r.W : 0;
  const bool range_vec4 = do_range && (HWr % 4 == 0) && (reinterpret_cast<uintptr_t>(xyz_out) % 16 == 0) &&
                          (!depth_out || reinterpret_cast<uintptr_t>(depth_out) % 16 == 0) &&
                          (!sem_out || reinterpret_cast<uintptr_t>(sem_out) % 4 == 0);
  // the sorted sparse list needs ranks (bitmap scan); a dense-only call counts n_occ while it emits
  const bool need_scan = do_vox && (sparse != nullptr || dense == nullptr);
  const bool dense_fast = do_vox && !need_scan;                  // dense-order bitmap, dense grid (and n_occ) from k_emit_dense
  int64_t* n_occ_emit = dense_fast ? n_occ : nullptr;
  int flags = (g_tuning[1] & 1) ? 0 : 1;                         // bit 0: neighbour filter before the voxel atomicMax
  if (!(g_tuning[5] & 1)) flags |= 8;                            // bit 3: evict_last hint on the voxel-table atomics (tuning key 5 = 1: off)
#ifdef MUVO_TIMING_KNOBS
  flags |= g_tuning[1] & 6;
#endif
  int sms = kNumSMsB200;
  { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }

  // K1 kernel of this call (persistent tile kernel: one contiguous run of tiles per CTA)
  const bool reg = do_vox ? g.regular != 0 : true;
  typedef void (*K1Fn)(const T*, const uint8_t*, const int64_t*, int, int64_t, int64_t, int, int, bool, GridDev, RangeDev, uint32_t*,
                       u64*, u64*, uint2*, uint32_t*, int64_t*, int, int64_t*);
  K1Fn k1;
  if (do_vox && do_range) k1 = reg ? k_points_tile<T, true, true, true> : k_points_tile<T, true, true, false>;
  else if (do_vox)        k1 = reg ? k_points_tile<T, true, false, true> : k_points_tile<T, true, false, false>;
  else if (prep_h)        k1 = k_points_tile<T, false, true, true, true>;
  else                    k1 = k_points_tile<T, false, true, true>;
  int k1_per_sm = 0;
  {
    cudaError_t e = cudaFuncSetAttribute((const void*)k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k1_per_sm, (const void*)k1, kTileThreads, tsmem);
    if (e != cudaSuccess) return ::muvo::cuda_fail(e);
    if (k1_per_sm < 1) k1_per_sm = 1;
    if (g_tuning[0] > 0 && g_tuning[0] < k1_per_sm) k1_per_sm = g_tuning[0];
  }
  if (dense_fast) {
    cudaError_t e = cudaFuncSetAttribute(k_emit_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EmitSmem));
    if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  }

  // launch with the programmatic-dependent-launch attribute (see pdl_wait)
  auto launch_pdl = [](auto kern, dim3 grid, dim3 block, size_t smem, cudaStream_t cs, auto... args) -> cudaError_t {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = cs;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
  };

  // One chunk = frames [f0, f1) = points [p0, p1): point pass, rare-path queues, emits, all on stream `cs`.
  int chunk_ctas[kMaxChunks] = {0};
  auto run_pass = [&](int chunk, int f0, int f1, int64_t p0, int64_t p1, cudaStream_t cs) -> int {
    int n_tile_ctas = 0;
    uint32_t* qcount = w.qcount + (size_t)chunk * kMaxTileCtas * 2;
    if (p1 > p0) {
      const int64_t n_tiles = ceil_div64(p1, kTile) - p0 / kTile;
      int64_t grid = (int64_t)sms * k1_per_sm;
      if (grid > n_tiles) grid = n_tiles;
      if (grid > kMaxTileCtas) grid = kMaxTileCtas;
      grid = ceil_div64(n_tiles, ceil_div64(n_tiles, grid));     // same tiles per CTA, no idle CTAs at the end
      k1<<<(unsigned)grid, kTileThreads, tsmem, cs>>>(xyz, sem, off, F, p0, p1, f0, f1, vec_ok, g, r, w.bitmap, w.vtab, w.pixtab,
                                                       w.queue, qcount, n_occ_emit, flags, diag);
      MUVO_AFTER_LAUNCH("k_points_tile", cs);
      n_tile_ctas = (int)grid;
    } else if (n_occ_emit) {
      cudaError_t e = cudaMemsetAsync(n_occ_emit + f0, 0, (size_t)(f1 - f0) * sizeof(int64_t), cs);
      if (e != cudaSuccess) return ::muvo::cuda_fail(e);
    }
    chunk_ctas[chunk] = n_tile_ctas;
    return MUVO_OK;
  };
  auto run_emit = [&](int chunk, int f0, int f1, int64_t p0, int64_t p1, cudaStream_t cs) -> int {
    const int n_tile_ctas = chunk_ctas[chunk];
    uint32_t* qcount = w.qcount + (size_t)chunk * kMaxTileCtas * 2;
    // K2: bitmap scan (when ranks are needed) + the rare-path queues, one launch
    {
      const int scan_frames = need_scan ? F : 0;                 // (the scan path is never chunked)
      const int queue_ctas = p1 > p0 ? (n_tile_ctas + kScanCluster - 1) / kScanCluster * kScanCluster : 0;
      const unsigned sgrid = (unsigned)(scan_frames * kScanCluster + queue_ctas);
      if (sgrid > 0) {
        QueueArgs<T> qa{xyz, sem, off, F, p0, p1, w.pixtab, w.vtab, w.queue, qcount, queue_ctas ? n_tile_ctas : 0, diag};
        launch_pdl(k_scan_queue<T>, dim3(sgrid), dim3(kScanThreads), 0, cs, (const uint32_t*)w.bitmap, w.prefix, g.gw, scan_frames, n_occ, qa, g, r);
        MUVO_AFTER_LAUNCH("k_scan_queue", cs);
      }
    }
    bool emitted_range = false;
    if (do_range) {              // right after its producers: the pixel words are still L2 resident
      emitted_range = true;
      EmitRangeArgs<T> era{w.pixtab, xyz, sem, off, f0, f1, depth_out, xyz_out, sem_out, prep_h ? prep_h->remap256 : nullptr};
      const int64_t npix = (int64_t)(f1 - f0) * HWr;
      if (range_vec4) {
        if (layout == MUVO_RANGE_LAYOUT_HWC) launch_pdl(k_emit_range<T, 4, MUVO_RANGE_LAYOUT_HWC>, dim3(blocks_for(npix / 4)), dim3(kBlock), 0, cs, era, r);
        else                                 launch_pdl(k_emit_range<T, 4, MUVO_RANGE_LAYOUT_XYZD>, dim3(blocks_for(npix / 4)), dim3(kBlock), 0, cs, era, r);
      } else {
        if (layout == MUVO_RANGE_LAYOUT_HWC) launch_pdl(k_emit_range<T, 1, MUVO_RANGE_LAYOUT_HWC>, dim3(blocks_for(npix)), dim3(kBlock), 0, cs, era, r);
        else                                 launch_pdl(k_emit_range<T, 1, MUVO_RANGE_LAYOUT_XYZD>, dim3(blocks_for(npix)), dim3(kBlock), 0, cs, era, r);
      }
      MUVO_AFTER_LAUNCH("k_emit_range", cs);
    }
    if (do_vox) {
      // K5 (the last consumer of the tables clears them)
      if (dense_fast) {
        const int bpf = (int)ceil_div64(g.gw, kBlock * kEmitWords);
        EmitDenseArgs eda{w.bitmap, w.vtab, off, sem, remap, dense, n_occ_emit, f0, (do_range && emitted_range) ? 0 : 1};
        launch_pdl(k_emit_dense, dim3((unsigned)bpf, (unsigned)(f1 - f0)), dim3(kBlock), sizeof(EmitSmem), cs, eda, g);
        MUVO_AFTER_LAUNCH("k_emit_dense", cs);
      } else {
        const int64_t words = (int64_t)F * g.gw;
        if (dense) {
          k_emit_dense_from_linear<<<blocks_for((int64_t)F * g.G), kBlock, 0, cs>>>(w.bitmap, w.vtab, off, sem, remap, dense, g, F);
          MUVO_AFTER_LAUNCH("k_emit_dense_from_linear", cs);
        }
        if (sparse && sparse_start) {
          k_frame_prefix<<<1, 1024, 0, cs>>>(n_occ, F, sparse_start);
          MUVO_AFTER_LAUNCH("k_frame_prefix", cs);
        }
        k_emit_sparse<<<blocks_for(words), kBlock, 0, cs>>>(w.bitmap, w.prefix, w.vtab, off, sem, sparse, sparse ? sparse_start : nullptr, g, F);
        MUVO_AFTER_LAUNCH("k_emit_sparse", cs);
      }
    }
    return MUVO_OK;
  };

  prof_mark("<points>", st);
  // float32 clouds -> dense grids / range images: the single-launch dataflow kernel (points_mega.cu)
  if (std::is_same<T, float>::value && !prep_h && (!do_vox || dense_fast) &&
      points_mega_eligible(do_vox ? &g : nullptr, do_range ? &r : nullptr, xyz, sem, dense, depth_out, xyz_out, sem_out, layout))
    return points_mega_f32(reinterpret_cast<const float*>(xyz), sem, off, F, P, do_vox ? &g : nullptr, remap, do_range ? &r : nullptr,
                           layout, dense, n_occ_emit, depth_out, xyz_out, sem_out, diag, w, st);
  // (Cutting the batch into frame chunks and running chunk c+1's point pass on a second stream next to chunk c's emit
  // kernels was measured, also under CUDA-graph replay: 249 us unchunked vs 275 / 342 us with 2 / 4 chunks -- the
  // persistent point pass owns the register file, so the kernels do not actually co-run.  Not kept.)
  if ((rc = run_pass(0, 0, F, 0, P, st)) != MUVO_OK) return rc;
  return run_emit(0, 0, F, 0, P, st);
}

}  // namespace
}  // namespace muvo

using namespace muvo;

extern "C" {

int muvo_debug_pixel_check(const float* xyz, int64_t n_points, const MuvoRangeCfg* cfg_h, int64_t* counts_out, void* stream) {
  if (!cfg_h || !counts_out || (n_points > 0 && !xyz)) return MUVO_E_NULL;
  if (n_points < 0) return MUVO_E_ARG;
  RangeDev r{};
  int rc = make_range_dev(cfg_h, &r);
  if (rc != MUVO_OK) return rc;
  if (n_points == 0) return MUVO_OK;
  k_debug_pixel_check<<<kNumSMsB200 * 8, kBlock, 0, (cudaStream_t)stream>>>(xyz, n_points, r, reinterpret_cast<unsigned long long*>(counts_out));
  MUVO_LAUNCH_CHECK();
  return MUVO_OK;
}

int muvo_debug_set_tuning(int32_t key, int32_t value) {
  if (key < 0 || key >= 8) return MUVO_E_ARG;
  g_tuning[key] = value;
  return MUVO_OK;
}

int muvo_points_workspace_bytes(int64_t n_points_total, int32_t n_frames, const MuvoGrid* grid_h,
                                const MuvoRangeCfg* range_h, size_t* bytes_out_h) {
  if (!bytes_out_h) return MUVO_E_NULL;
  if (n_points_total < 0 || n_frames < 0) return MUVO_E_ARG;
  PointsWs w = carve(nullptr, n_points_total, n_frames, grid_h, range_h);
  *bytes_out_h = w.bytes + 256;
  return MUVO_OK;
}

int muvo_ws_reset(void* ws, size_t ws_bytes, void* stream) {
  if (!ws) return MUVO_E_NULL;
  cudaError_t e = cudaMemsetAsync(ws, 0, ws_bytes, (cudaStream_t)stream);
  return e == cudaSuccess ? MUVO_OK : (int)e;
}

int muvo_voxelize(const void* xyz, int32_t xyz_dtype, const uint8_t* sem, const int64_t* frame_offsets,
                  int32_t n_frames, int64_t n_points_total, const MuvoGrid* grid_h,
                  const uint8_t* remap256, uint8_t* dense_out, uint16_t* sparse_out, int64_t* n_occ_out,
                  int64_t* sparse_start_out, int64_t* diag, void* ws, size_t ws_bytes, void* stream) {
  if (!grid_h) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (xyz_dtype == MUVO_F32)
    return run_points<float>((const float*)xyz, sem, frame_offsets, n_frames, n_points_total, grid_h, remap256,
                             nullptr, 0, dense_out, sparse_out, n_occ_out, sparse_start_out, nullptr, nullptr, nullptr, diag, ws,
                             ws_bytes, st);
  if (xyz_dtype == MUVO_F64)
    return run_points<double>((const double*)xyz, sem, frame_offsets, n_frames, n_points_total, grid_h, remap256,
                              nullptr, 0, dense_out, sparse_out, n_occ_out, sparse_start_out, nullptr, nullptr, nullptr, diag, ws,
                              ws_bytes, st);
  return MUVO_E_ARG;
}

int muvo_range_project(const float* xyz, const uint8_t* sem, const int64_t* frame_offsets,
                       int32_t n_frames, int64_t n_points_total, const MuvoRangeCfg* cfg_h, int32_t layout, float* depth_out,
                       float* xyz_out, uint8_t* sem_out, int64_t* diag, void* ws, size_t ws_bytes, void* stream) {
  if (!cfg_h) return MUVO_E_NULL;
  return run_points<float>(xyz, sem, frame_offsets, n_frames, n_points_total, nullptr, nullptr, cfg_h, layout,
                           nullptr, nullptr, nullptr, nullptr, depth_out, xyz_out, sem_out, diag, ws, ws_bytes,
                           (cudaStream_t)stream);
}

int muvo_range_project_lidar(const float* xyz_raw, const uint8_t* tag, const int64_t* frame_offsets, int32_t n_frames,
                             int64_t n_points_total, const MuvoRangeCfg* cfg_h, const MuvoLidarPrep* prep_h, int32_t layout,
                             float* depth_out, float* xyz_out, uint8_t* sem_out, int64_t* diag, void* ws, size_t ws_bytes,
                             void* stream) {
  if (!cfg_h || !prep_h) return MUVO_E_NULL;
  return run_points<float>(xyz_raw, tag, frame_offsets, n_frames, n_points_total, nullptr, nullptr, cfg_h, layout,
                           nullptr, nullptr, nullptr, nullptr, depth_out, xyz_out, sem_out, diag, ws, ws_bytes,
                           (cudaStream_t)stream, prep_h);
}

int muvo_points_fused(const float* xyz, const uint8_t* sem, const int64_t* frame_offsets,
                      int32_t n_frames, int64_t n_points_total, const MuvoGrid* grid_h, const uint8_t* remap256,
                      const MuvoRangeCfg* cfg_h, int32_t layout, uint8_t* dense_out, uint16_t* sparse_out,
                      int64_t* n_occ_out, int64_t* sparse_start_out, float* depth_out, float* xyz_out, uint8_t* sem_out,
                      int64_t* diag, void* ws, size_t ws_bytes, void* stream) {
  if (!grid_h || !cfg_h) return MUVO_E_NULL;
  return run_points<float>(xyz, sem, frame_offsets, n_frames, n_points_total, grid_h, remap256, cfg_h, layout,
                           dense_out, sparse_out, n_occ_out, sparse_start_out, depth_out, xyz_out, sem_out, diag, ws, ws_bytes,
                           (cudaStream_t)stream);
}

}  // extern "C"

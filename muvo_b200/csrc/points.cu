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
//   K1  k_point_pass    : per point, float64 voxel id + range pixel / depth in numpy's operation order (no
//                         FMA contraction: this file is compiled with -fmad=false; an f32 pre-filter skips
//                         the f64 trigonometry for points that are provably far from a bin edge).
//                         Occupied voxels are marked in a 1-bit-per-voxel frame bitmap; the range pixel is
//                         resolved with ONE 64-bit atomicMax on   (~top32(depth bits) << 32) | ~(index+1).
//                         That packed order equals the exact (depth, index) order unless two points of a
//                         pixel agree in the top 32 key bits; those (rare) run an exact tie protocol.
//   K2  k_bitmap_scan   : popcount prefix per 128-bit bitmap chunk -> every occupied voxel gets a dense
//                         slot id = its rank in output order; n_occ per frame.
//   K3  k_voxel_resolve : second point pass (no trigonometry): same packed atomicMax on slot `rank`,
//                         key = (not roadline, |p mod res|^2).
//   K4  k_slot_labels   : one thread per winner slot: point index -> label byte.
//   K5  k_emit_*        : bitmap-ordered, fully coalesced write of the dense uint8 grid (zeros included,
//                         so no memset + scatter) and/or the sorted sparse (n,4) list; pixel-ordered write
//                         of the range image.  Emit kernels put every table entry they consume back to 0,
//                         so the workspace is clean for the next call.
#include <math.h>
#include "common.cuh"

namespace muvo {
namespace {

#ifndef MUVO_KPL
#define MUVO_KPL 2
#endif
constexpr int kBlock = 256;
constexpr double kPi = 3.141592653589793;            // np.pi
constexpr double kPiOver4 = 0x1.921fb54442d18p-1;    // correctly rounded pi/4 (numpy/glibc value on diagonals)
constexpr double k3PiOver4 = 0x1.2d97c7f3321d2p+1;   // correctly rounded 3pi/4
constexpr double kEdgeEps = 1e-9;                    // diagnostics: "on a bin edge"
constexpr float kFastEps = 2e-3f;                    // f32 pre-filter: distance to a bin edge below which f64 decides

enum BitOrder { ORDER_DENSE = 0 /* (x*Dy+y)*Dz+z */, ORDER_LINEAR = 1 /* x + Dx*(y + Dy*z) */ };

struct GridDev {
  double res, inv_res;
  double off[3], up[3];
  int dx, dy, dz;
  int road;
  int pow2;
  int order;
  int gw;        // bitmap words per frame (multiple of 32)
  int64_t G;     // voxels per frame
};

struct RangeDev {
  int H, W;
  double fda, fov;
  double L[3];
  float inv_pi_f, fda_f, inv_fov_f;   // f32 pre-filter constants
};

typedef unsigned long long u64;

struct PointsWs {
  uint32_t* bitmap = nullptr;   // [F, gw]
  uint32_t* prefix = nullptr;   // [F, gw/4]  exclusive popcount prefix per 128-bit chunk
  u64* pixtab = nullptr;        // [F, H*W]   packed winner per pixel (0 = empty)
  u64* vslot = nullptr;         // [P]        packed winner per occupied voxel, slot = frame_offsets[f] + rank
  size_t bytes = 0;
};

static int bitmap_words(int64_t G) { return (int)(ceil_div64(G, 1024) * 32); }

// Region offsets depend only on (F, grid size, H*W); the per-point region comes last so that ragged batches
// (different P) keep every table at the same address.
static PointsWs carve(void* base, int64_t P, int F, const MuvoGrid* g, const MuvoRangeCfg* r) {
  PointsWs w;
  size_t o = 0;
  char* b = (char*)base;
  if (g) {
    int64_t G = (int64_t)g->size[0] * g->size[1] * g->size[2];
    size_t gw = (size_t)bitmap_words(G);
    w.bitmap = (uint32_t*)(b + o); o = align_up(o + (size_t)F * gw * 4, 256);
    w.prefix = (uint32_t*)(b + o); o = align_up(o + (size_t)F * (gw / 4) * 4, 256);
  }
  if (r) {
    w.pixtab = (u64*)(b + o); o = align_up(o + (size_t)F * r->H * r->W * 8, 256);
  }
  if (g) {
    w.vslot = (u64*)(b + o); o = align_up(o + (size_t)(P > 0 ? P : 1) * 8, 256);
  }
  w.bytes = o;
  return w;
}

// ---------------------------------------------------------------- packed winner words
// word = (~top32(key) << 32) | ~(idx + 1), key = positive float64 bit pattern (optionally with bit 63 set).
// Larger word <=> smaller key, then smaller index; 0 = empty.
__device__ __forceinline__ uint32_t key_top_inv(u64 key) { return ~(uint32_t)(key >> 32); }
__device__ __forceinline__ u64 pack_word(uint32_t top_inv, uint32_t idx1) { return ((u64)top_inv << 32) | (uint32_t)(~idx1); }
__device__ __forceinline__ uint32_t word_idx1(u64 w) { return ~(uint32_t)w; }
__device__ __forceinline__ uint32_t word_top(u64 w) { return (uint32_t)(w >> 32); }
// voxel words also carry the point's label in the low byte (index limited to 24 bits per frame), so the emit
// kernels read the label straight from the slot: word = top_inv << 32 | (~idx1 & 0xffffff) << 8 | label
constexpr uint32_t kVoxIdxMask = 0xffffffu;
__device__ __forceinline__ u64 pack_vox(uint32_t top_inv, uint32_t idx1, uint32_t label) {
  return ((u64)top_inv << 32) | (u64)(((~idx1) & kVoxIdxMask) << 8) | (label & 0xffu);
}
__device__ __forceinline__ uint32_t vox_idx1(u64 w) { return (~((uint32_t)w >> 8)) & kVoxIdxMask; }

// ---------------------------------------------------------------- per-point arithmetic
// numpy's npy_divmod (numpy/_core/src/npymath/npy_math_internal.h.src), the scalar behind np.divmod
// at data_preprocessing.py:183; needed when res is not a power of two.  fmod is exact on both sides.
__device__ __noinline__ double npy_divmod_dev(double a, double b, double* modulus) {
  double mod = fmod(a, b);
  double div = (a - mod) / b;
  if (mod != 0.0) {
    if ((b < 0) != (mod < 0)) { mod += b; div -= 1.0; }
  } else {
    mod = copysign(0.0, b);
  }
  double fl;
  if (div != 0.0) {
    fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
  } else {
    fl = copysign(0.0, a / b);
  }
  *modulus = mod;
  return fl;
}

struct VoxKey {
  uint32_t bit;   // bit index inside the frame bitmap (order per GridDev::order)
  double dis;     // (mx^2 + my^2) + mz^2, float64, np.sum(axis=1) order (:212)
  bool in;
};

template <bool NEED_DIS = true>
__device__ __forceinline__ VoxKey vox_of(double px, double py, double pz, const GridDev& g) {
  VoxKey k;
  double bx = px + g.off[0], by = py + g.off[1], bz = pz + g.off[2];          // :177
  k.in = (bx >= 0.0) && (bx < g.up[0]) && (by >= 0.0) && (by < g.up[1]) && (bz >= 0.0) && (bz < g.up[2]);   // :178
  k.bit = 0; k.dis = 0.0;
  if (!k.in) return k;
  double cx, cy, cz, mx, my, mz;
  if (g.pow2) {  // floor(b/res) and b - floor*res carry no rounding for a power-of-two res
    cx = floor(bx * g.inv_res); cy = floor(by * g.inv_res); cz = floor(bz * g.inv_res);
    mx = bx - cx * g.res; my = by - cy * g.res; mz = bz - cz * g.res;
  } else {
    cx = npy_divmod_dev(bx, g.res, &mx);
    cy = npy_divmod_dev(by, g.res, &my);
    cz = npy_divmod_dev(bz, g.res, &mz);
  }
  if (NEED_DIS) k.dis = (mx * mx + my * my) + mz * mz;
  int ix = (int)cx, iy = (int)cy, iz = (int)cz;
  if (ix < 0 || ix >= g.dx || iy < 0 || iy >= g.dy || iz < 0 || iz >= g.dz) { k.in = false; return k; }
  k.bit = (g.order == ORDER_DENSE) ? (uint32_t)((ix * g.dy + iy) * g.dz + iz)
                                   : (uint32_t)(ix + g.dx * (iy + g.dy * iz));
  return k;
}
// voxel ordering key: bit 63 = "not a roadline point" (dis >= 0 leaves it free), rest = dis bits
__device__ __forceinline__ u64 vox_key(double dis, bool notroad) {
  return (u64)__double_as_longlong(dis) | (notroad ? (1ull << 63) : 0ull);
}

// LiDAR-frame coordinates + depth (for the point itself and for re-deriving a competitor's key)
template <typename T>
__device__ __forceinline__ double range_depth_of(T x, T y, T z, const RangeDev& r, double* xc_o, double* yc_o,
                                                 double* zc_o) {
  double xc = (double)x - r.L[0];          // :177-178  (x * 1) - L0
  double yc = (-(double)y) - r.L[1];       //           (y * -1) - L1   (keeps the sign of zero)
  double zc = (double)z - r.L[2];
  *xc_o = xc; *yc_o = yc; *zc_o = zc;
  return sqrt((xc * xc + yc * yc) + zc * zc);   // :180  np.linalg.norm(., 2, axis=1)
}

__device__ __forceinline__ double atan2_np(double y, double x) {
  // Exact bin edges exist only on the axes and diagonals; numpy returns the correctly rounded
  // multiples of pi/4 there.  CUDA's atan2 is exact on the axes; pin the diagonals explicitly.
  if (fabs(y) == fabs(x) && x != 0.0 && isfinite(x)) return copysign(x > 0.0 ? kPiOver4 : k3PiOver4, y);
  return atan2(y, x);
}

struct PixKey {
  int pix;        // h*W + w
  double s;       // squared range (xc^2 + yc^2) + zc^2: orders like the depth; sqrt only where exactness needs it
  bool ok;
  bool near_w, near_h;
};

// float64 pixel exactly as the reference computes it (geometry_utils.py:180-200)
__device__ __noinline__ void pix_exact(double xc, double yc, double zc, double s, int H, int W, double fda, double fov,
                                       int* pw_o, int* ph_o, int* flags_o) {
  double depth = sqrt(s);                                 // :180
  double yy = -yc;                                        // :183
  double yaw = atan2_np(yy, xc);                          // :186
  double pitch = asin(zc / depth);                        // :187
  double pw = 0.5 * (1.0 - yaw / kPi);                    // :189
  double ph = 1.0 - (pitch + fda) / fov;                  // :190
  pw *= (double)W;                                        // :191
  ph *= (double)H;                                        // :192
  int flags = 0;
  if (!(pw == pw) || !(ph == ph)) { *flags_o = 4; *pw_o = 0; *ph_o = 0; return; }
  double fw = floor(pw), fh = floor(ph);                  // :194,:198
  if ((pw > 0.0 && pw < (double)W) && ((pw - fw) < kEdgeEps || (pw - fw) > 1.0 - kEdgeEps)) flags |= 1;
  if ((ph > 0.0 && ph < (double)H) && ((ph - fh) < kEdgeEps || (ph - fh) > 1.0 - kEdgeEps)) flags |= 2;
  fw = fmax(0.0, fmin((double)(W - 1), fw));              // :195-196
  fh = fmax(0.0, fmin((double)(H - 1), fh));              // :199-200
  *pw_o = (int)fw; *ph_o = (int)fh; *flags_o = flags;
}

template <typename T>
__device__ __forceinline__ double range_sq_of(T x, T y, T z, const RangeDev& r, double* xc_o, double* yc_o, double* zc_o) {
  double xc = (double)x - r.L[0];          // :177-178  (x * 1) - L0
  double yc = (-(double)y) - r.L[1];       //           (y * -1) - L1   (keeps the sign of zero)
  double zc = (double)z - r.L[2];
  *xc_o = xc; *yc_o = yc; *zc_o = zc;
  return (xc * xc + yc * yc) + zc * zc;
}

template <typename T>
__device__ __forceinline__ PixKey pix_of(T x, T y, T z, const RangeDev& r) {
  PixKey k;
  double xc, yc, zc;
  k.s = range_sq_of(x, y, z, r, &xc, &yc, &zc);
  k.ok = isfinite(k.s) && k.s > 0.0;
  k.pix = 0; k.near_w = k.near_h = false;
  if (!k.ok) return k;
  // f32 pre-filter.  |error| of pw_f / ph_f vs the float64 value is < 5e-4 bins (atan2f/asinf <= 3 ulp, a few
  // f32 roundings at magnitude <= W); if both are farther than kFastEps from an interior integer the floor
  // cannot differ from the float64 floor.  Values beyond the image clamp to the border bins on both paths.
  const float xf = (float)xc, yf = (float)(-yc), zf = (float)zc, df = sqrtf((float)k.s);
  const float pw_f = 0.5f * (1.0f - atan2f(yf, xf) * r.inv_pi_f) * (float)r.W;
  const float ph_f = (1.0f - (asinf(zf / df) + r.fda_f) * r.inv_fov_f) * (float)r.H;
  const float fw = floorf(pw_f), fh = floorf(ph_f);
  const bool safe_w = (pw_f <= 0.5f) || (pw_f >= (float)r.W - 0.5f) || (pw_f - fw > kFastEps && pw_f - fw < 1.0f - kFastEps);
  const bool safe_h = (ph_f <= 0.5f) || (ph_f >= (float)r.H - 0.5f) || (ph_f - fh > kFastEps && ph_f - fh < 1.0f - kFastEps);
  int iw, ih;
  if (safe_w && safe_h) {   // NaN compares false -> exact path
    iw = (int)fminf(fmaxf(fw, 0.0f), (float)(r.W - 1));
    ih = (int)fminf(fmaxf(fh, 0.0f), (float)(r.H - 1));
  } else {
    int flags;
    pix_exact(xc, yc, zc, k.s, r.H, r.W, r.fda, r.fov, &iw, &ih, &flags);
    if (flags & 4) { k.ok = false; return k; }
    k.near_w = flags & 1; k.near_h = flags & 2;
  }
  k.pix = ih * r.W + iw;
  return k;
}

__device__ __forceinline__ int find_frame(const int64_t* __restrict__ off, int F, int64_t i) {
  int lo = 0, hi = F;   // largest f with off[f] <= i
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------- point loads
// A warp owns 32*KPL consecutive points; lane L works on points base + 32k + L (k < KPL).  Global loads are
// 16-byte vectors staged through shared memory and read back with a stride of 3 words (conflict free since
// gcd(3, 32) = 1); the KPL points of a lane are processed in lock step so that their atomics are in flight
// together.
constexpr int kWarpsPerBlock = kBlock / 32;
template <int KPL> struct Stage { static constexpr int words = 3 * 32 * KPL + 8 * KPL; };   // xyz floats + packed semantics

template <typename T, int KPL> struct WarpPts { T x[KPL], y[KPL], z[KPL]; uint32_t sem[KPL]; bool valid[KPL]; };

template <typename T, int KPL>
__device__ __forceinline__ void load_warp_points_scalar(const T* __restrict__ xyz, const uint8_t* __restrict__ sem,
                                                        int64_t base, int64_t P, WarpPts<T, KPL>& w) {
  const unsigned lane = lane_id();
#pragma unroll
  for (int k = 0; k < KPL; ++k) {
    int64_t i = base + 32 * k + lane;
    w.valid[k] = i < P;
    w.x[k] = w.y[k] = w.z[k] = (T)0; w.sem[k] = 0;
    if (w.valid[k]) {
      w.x[k] = __ldg(xyz + 3 * i); w.y[k] = __ldg(xyz + 3 * i + 1); w.z[k] = __ldg(xyz + 3 * i + 2);
      w.sem[k] = __ldg(sem + i);
    }
  }
}
template <typename T, int KPL>
__device__ __forceinline__ void load_warp_points(const T* __restrict__ xyz, const uint8_t* __restrict__ sem, int64_t base,
                                                 int64_t P, bool vec_ok, uint32_t* stage, WarpPts<T, KPL>& w) {
  load_warp_points_scalar<T, KPL>(xyz, sem, base, P, w);
}
template <int KPL>
__device__ __forceinline__ void load_warp_points_f32(const float* __restrict__ xyz, const uint8_t* __restrict__ sem,
                                                     int64_t base, int64_t P, bool vec_ok, uint32_t* stage,
                                                     WarpPts<float, KPL>& w) {
  const unsigned lane = lane_id();
  constexpr int kPts = 32 * KPL;
  if (vec_ok && base + kPts <= P) {
    const float4* src = reinterpret_cast<const float4*>(xyz + 3 * base);
    float4* dst = reinterpret_cast<float4*>(stage);
#pragma unroll
    for (int m = 0; m < (24 * KPL + 31) / 32; ++m) {          // 24*KPL float4 per warp
      int q = m * 32 + lane;
      if (q < 24 * KPL) dst[q] = __ldg(src + q);
    }
    if (lane < 8 * KPL) stage[3 * kPts + lane] = __ldg(reinterpret_cast<const uint32_t*>(sem + base) + lane);
    __syncwarp();
    const float* sf = reinterpret_cast<const float*>(stage);
    const uint8_t* sb = reinterpret_cast<const uint8_t*>(stage + 3 * kPts);
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
      int j = 32 * k + lane;
      w.valid[k] = true;
      w.x[k] = sf[3 * j]; w.y[k] = sf[3 * j + 1]; w.z[k] = sf[3 * j + 2];
      w.sem[k] = sb[j];
    }
    __syncwarp();
  } else {
    load_warp_points_scalar<float, KPL>(xyz, sem, base, P, w);
  }
}
template <> __device__ __forceinline__ void load_warp_points<float, 1>(const float* __restrict__ xyz, const uint8_t* __restrict__ sem, int64_t base, int64_t P, bool vec_ok, uint32_t* stage, WarpPts<float, 1>& w) { load_warp_points_f32<1>(xyz, sem, base, P, vec_ok, stage, w); }
template <> __device__ __forceinline__ void load_warp_points<float, 2>(const float* __restrict__ xyz, const uint8_t* __restrict__ sem, int64_t base, int64_t P, bool vec_ok, uint32_t* stage, WarpPts<float, 2>& w) { load_warp_points_f32<2>(xyz, sem, base, P, vec_ok, stage, w); }
template <> __device__ __forceinline__ void load_warp_points<float, 4>(const float* __restrict__ xyz, const uint8_t* __restrict__ sem, int64_t base, int64_t P, bool vec_ok, uint32_t* stage, WarpPts<float, 4>& w) { load_warp_points_f32<4>(xyz, sem, base, P, vec_ok, stage, w); }

// frames of the warp's points: one search for the first point, then a (rare) walk at frame boundaries
template <int KPL>
__device__ __forceinline__ void warp_frames(const int64_t* __restrict__ off, int F, int64_t base, int64_t P, int* fr,
                                            int64_t* fb) {
  int f0 = find_frame(off, F, base < P ? base : P - 1);
  const unsigned lane = lane_id();
#pragma unroll
  for (int k = 0; k < KPL; ++k) {
    int64_t i = base + 32 * k + lane;
    int f = f0;
    if (i < P) { while (i >= __ldg(off + f + 1)) ++f; }
    fr[k] = f; fb[k] = __ldg(off + f);
  }
}

__device__ __forceinline__ void diag_add(int64_t* diag, int slot, unsigned v) {
  unsigned tot = __reduce_add_sync(0xffffffffu, v);
  if (tot && lane_id() == 0) atomicAdd(reinterpret_cast<unsigned long long*>(diag + slot), (unsigned long long)tot);
}

// ---------------------------------------------------------------- exact tie protocol (rare path)
// Called by a point whose atomicMax met a slot holder with the SAME top-32 key bits.  `key_of(idx1)` re-derives
// the exact 64-bit key of point idx1 from its coordinates, `pack(idx1)` builds its slot word, `idx_of(word)`
// extracts the index.  On return the slot holds a point that is exactly <= this point and the one it may have
// displaced (smaller key, then smaller index), or a point from a strictly better top-32 class.  Every tied
// point runs this, so the final holder is the exact arg-min.
template <typename KeyFn, typename PackFn, typename IdxFn>
__device__ __noinline__ void tie_protocol(u64* slot, uint32_t top_inv, uint32_t me1, u64 mine, u64 old_word, KeyFn key_of,
                                          PackFn pack, IdxFn idx_of) {
  uint32_t cand1 = me1;
  u64 cand_key = key_of(me1);
  {
    uint32_t o1 = idx_of(old_word);
    u64 ok = key_of(o1);
    if (ok < cand_key || (ok == cand_key && o1 < cand1)) { cand1 = o1; cand_key = ok; }
  }
  u64 cur = old_word > mine ? old_word : mine;     // content right after this point's atomicMax
  for (;;) {
    if (word_top(cur) != top_inv) break;           // a strictly better class took the slot
    uint32_t h1 = idx_of(cur);
    if (h1 == cand1) break;                        // the slot holds the candidate
    u64 hk = key_of(h1);
    bool cand_better = cand_key < hk || (cand_key == hk && cand1 < h1);
    if (!cand_better) break;                       // holder is exactly better: already in place
    u64 prev = atomicCAS(slot, cur, pack(cand1));
    if (prev == cur) break;
    cur = prev;
  }
}

// ---------------------------------------------------------------- K1: point pass
template <typename T, int KPL, bool DO_VOX, bool DO_RANGE>
__global__ void __launch_bounds__(kBlock)
k_point_pass(const T* __restrict__ xyz, const uint8_t* __restrict__ sem, const int64_t* __restrict__ off, int F, int64_t P,
             bool vec_ok, GridDev g, RangeDev r, uint32_t* __restrict__ bitmap, u64* __restrict__ pixtab,
             int64_t* __restrict__ diag) {
  __shared__ __align__(16) uint32_t stage_all[kWarpsPerBlock * Stage<KPL>::words];
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int64_t base = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * (32 * KPL);
  unsigned n_drop = 0, n_nw = 0, n_nh = 0, n_in = 0;
  if (base < P) {   // warp-uniform
    WarpPts<T, KPL> w;
    load_warp_points<T, KPL>(xyz, sem, base, P, vec_ok, stage_all + warp * Stage<KPL>::words, w);
    int fr[KPL]; int64_t fb[KPL];
    warp_frames<KPL>(off, F, base, P, fr, fb);
    if (DO_VOX) {
#pragma unroll
      for (int k = 0; k < KPL; ++k) {
        if (w.valid[k]) {
          VoxKey v = vox_of<false>((double)w.x[k], (double)w.y[k], (double)w.z[k], g);
          if (v.in) {
            ++n_in;
            atomicOr(bitmap + (size_t)fr[k] * g.gw + (v.bit >> 5), 1u << (v.bit & 31));   // RED, no return value
          }
        }
      }
    }
    if (DO_RANGE) {
      u64* slot[KPL];
      u64 mine[KPL], old[KPL];
      bool act[KPL];
#pragma unroll
      for (int k = 0; k < KPL; ++k) {
        act[k] = false; slot[k] = pixtab; mine[k] = 0;
        if (w.valid[k]) {
          PixKey pk = pix_of(w.x[k], w.y[k], w.z[k], r);
          if (!pk.ok) { ++n_drop; }
          else {
            n_nw += pk.near_w; n_nh += pk.near_h;
            act[k] = true;
            slot[k] = pixtab + (size_t)fr[k] * r.H * r.W + pk.pix;
            // s > 0: its bit pattern orders like the value, and like the depth sqrt(s)
            mine[k] = pack_word(key_top_inv((u64)__double_as_longlong(pk.s)), (uint32_t)(base + 32 * k + lane - fb[k]) + 1u);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < KPL; ++k) old[k] = act[k] ? atomicMax(slot[k], mine[k]) : 0ull;
#pragma unroll
      for (int k = 0; k < KPL; ++k) {
        if (act[k] && old[k] != 0ull && word_top(old[k]) == word_top(mine[k])) {   // same top-32 class: exact protocol
          const T* fx = xyz + 3 * fb[k];
          const uint32_t top = word_top(mine[k]);
          auto key_of = [&](uint32_t q1) -> u64 {     // exact key: the float64 depth (geometry_utils.py:180)
            const T* qp = fx + 3 * (int64_t)(q1 - 1u);
            double a, b, c;
            return (u64)__double_as_longlong(sqrt(range_sq_of(__ldg(qp), __ldg(qp + 1), __ldg(qp + 2), r, &a, &b, &c)));
          };
          auto pack = [&](uint32_t q1) -> u64 { return pack_word(top, q1); };
          auto idx_of = [](u64 wv) -> uint32_t { return word_idx1(wv); };
          tie_protocol(slot[k], top, word_idx1(mine[k]), mine[k], old[k], key_of, pack, idx_of);
        }
      }
    }
  }
  __syncwarp();
  if (diag) {
    if (DO_RANGE) {
      diag_add(diag, MUVO_DIAG_DROPPED_NONFINITE, n_drop);
      diag_add(diag, MUVO_DIAG_NEAR_EDGE_W, n_nw);
      diag_add(diag, MUVO_DIAG_NEAR_EDGE_H, n_nh);
    }
    if (DO_VOX) diag_add(diag, MUVO_DIAG_IN_GRID, n_in);
  }
}

// ---------------------------------------------------------------- K2: bitmap scan (one CTA per frame)
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads)
k_bitmap_scan(const uint32_t* __restrict__ bitmap, uint32_t* __restrict__ prefix, int gw, int64_t* __restrict__ n_occ_out) {
  __shared__ uint32_t warp_tot[kScanThreads / 32];
  __shared__ uint32_t carry_s;
  const int f = blockIdx.x;
  const int chunks = gw / 4;
  const uint4* bm = reinterpret_cast<const uint4*>(bitmap + (size_t)f * gw);
  uint32_t* pf = prefix + (size_t)f * chunks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  constexpr int kPer = 4;  // chunks per thread per tile (64 contiguous bytes)
  for (int base = 0; base < chunks; base += kScanThreads * kPer) {
    int c0 = base + threadIdx.x * kPer;
    uint32_t cnt[kPer];
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      uint32_t c = 0;
      if (c0 + k < chunks) { uint4 v = bm[c0 + k]; c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }
      cnt[k] = c; tsum += c;
    }
    uint32_t incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t carry = carry_s;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_tot[lane];
      uint32_t wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
      warp_tot[lane] = wi - w;   // exclusive warp offsets
      if (lane == 31) carry_s = carry + wi;
    }
    __syncthreads();
    uint32_t ex = carry + warp_tot[warp] + (incl - tsum);
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      if (c0 + k < chunks) pf[c0 + k] = ex;
      ex += cnt[k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && n_occ_out) n_occ_out[f] = (int64_t)carry_s;
}

// rank of set bit `bit` within its frame (= number of set bits before it)
__device__ __forceinline__ uint32_t rank_of(const uint32_t* __restrict__ bitmap_f, const uint32_t* __restrict__ prefix_f,
                                            uint32_t bit) {
  uint32_t chunk = bit >> 7;
  uint4 v = *reinterpret_cast<const uint4*>(bitmap_f + chunk * 4);
  uint32_t w = (bit >> 5) & 3u;
  uint32_t below = 0;
  uint32_t word = v.x;
  if (w >= 1) { below += __popc(v.x); word = v.y; }
  if (w >= 2) { below += __popc(v.y); word = v.z; }
  if (w >= 3) { below += __popc(v.z); word = v.w; }
  below += __popc(word & ((1u << (bit & 31)) - 1u));
  return prefix_f[chunk] + below;
}

// ---------------------------------------------------------------- K3: voxel resolve
// PACKL: the slot word also carries the point's label (needs < 2^24 points per frame); otherwise the label pass
// (k_slot_labels) fills it in afterwards.
template <typename T, int KPL, bool PACKL>
__global__ void __launch_bounds__(kBlock)
k_voxel_resolve(const T* __restrict__ xyz, const uint8_t* __restrict__ sem, const int64_t* __restrict__ off, int F, int64_t P,
                bool vec_ok, GridDev g, const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix,
                u64* __restrict__ vslot) {
  __shared__ __align__(16) uint32_t stage_all[kWarpsPerBlock * Stage<KPL>::words];
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int64_t base = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * (32 * KPL);
  if (base >= P) return;   // warp-uniform
  WarpPts<T, KPL> w;
  load_warp_points<T, KPL>(xyz, sem, base, P, vec_ok, stage_all + warp * Stage<KPL>::words, w);
  int fr[KPL]; int64_t fb[KPL];
  warp_frames<KPL>(off, F, base, P, fr, fb);
  u64* slot[KPL];
  u64 mine[KPL], old[KPL];
  uint32_t bit[KPL];
  bool act[KPL];
#pragma unroll
  for (int k = 0; k < KPL; ++k) {
    act[k] = false; slot[k] = vslot; mine[k] = 0; bit[k] = 0;
    if (w.valid[k]) {
      VoxKey v = vox_of<true>((double)w.x[k], (double)w.y[k], (double)w.z[k], g);
      act[k] = v.in; bit[k] = v.bit;
      const uint32_t top = key_top_inv(vox_key(v.dis, (int)w.sem[k] != g.road));
      const uint32_t me1 = (uint32_t)(base + 32 * k + lane - fb[k]) + 1u;
      mine[k] = PACKL ? pack_vox(top, me1, w.sem[k]) : pack_word(top, me1);
    }
  }
#pragma unroll
  for (int k = 0; k < KPL; ++k) {   // rank lookups of the lane's points are independent loads
    if (act[k]) slot[k] = vslot + fb[k] + rank_of(bitmap + (size_t)fr[k] * g.gw, prefix + (size_t)fr[k] * (g.gw / 4), bit[k]);
  }
#pragma unroll
  for (int k = 0; k < KPL; ++k) old[k] = act[k] ? atomicMax(slot[k], mine[k]) : 0ull;
#pragma unroll
  for (int k = 0; k < KPL; ++k) {
    if (act[k] && old[k] != 0ull && word_top(old[k]) == word_top(mine[k])) {
      const T* fx = xyz + 3 * fb[k];
      const uint8_t* fs = sem + fb[k];
      const uint32_t top = word_top(mine[k]);
      auto key_of = [&](uint32_t q1) -> u64 {
        const T* qp = fx + 3 * (int64_t)(q1 - 1u);
        VoxKey o = vox_of<true>((double)__ldg(qp), (double)__ldg(qp + 1), (double)__ldg(qp + 2), g);
        return vox_key(o.dis, (int)__ldg(fs + (q1 - 1u)) != g.road);
      };
      auto pack = [&](uint32_t q1) -> u64 { return PACKL ? pack_vox(top, q1, __ldg(fs + (q1 - 1u))) : pack_word(top, q1); };
      auto idx_of = [](u64 wv) -> uint32_t { return PACKL ? vox_idx1(wv) : word_idx1(wv); };
      const uint32_t me1 = (uint32_t)(base + 32 * k + lane - fb[k]) + 1u;
      tie_protocol(slot[k], top, me1, mine[k], old[k], key_of, pack, idx_of);
    }
  }
}

// ---------------------------------------------------------------- K4: slot labels
// One thread per winner slot: replace the packed winner by (kLabelTag | raw label of the winning point) so that
// the emit kernels need a single gather per occupied voxel.  Slots that were never claimed stay 0.
constexpr u64 kLabelTag = 0xffffffff00000000ull;
__global__ void __launch_bounds__(kBlock)
k_slot_labels(u64* __restrict__ vslot, const uint8_t* __restrict__ sem, const int64_t* __restrict__ off, int F, int64_t P) {
  int64_t s = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (s >= P) return;
  u64 wv = vslot[s];
  if (!wv) return;
  int f = find_frame(off, F, s);
  uint32_t lab = __ldg(sem + __ldg(off + f) + (int64_t)(word_idx1(wv) - 1u));
  vslot[s] = kLabelTag | lab;
}

// ---------------------------------------------------------------- K5: emit
// Shared by the emit kernels: lane L owns bitmap word (warp_word0 + L); returns the word and the rank of
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

// 16 voxels (a half word) -> 16 label bytes (two 64-bit halves); one gather per set bit
__device__ __forceinline__ uint4 expand_half(uint32_t bits16, uint32_t rank, u64* __restrict__ vslot_f,
                                             const uint8_t* __restrict__ remap, bool clean) {
  u64 lo = 0, hi = 0;
  u64* sl = vslot_f + rank;
  while (bits16) {
    int j = __ffs(bits16) - 1;
    bits16 &= bits16 - 1;
    uint32_t lab = (uint32_t)(*sl) & 0xffu;
    if (clean) *sl = 0ull;
    ++sl;
    if (remap) lab = __ldg(remap + lab);
    u64 add = (u64)lab << (8 * (j & 7));
    if (j < 8) lo |= add; else hi |= add;
  }
  return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}

// Dense grid, bitmap in dense order.  A warp owns 32 words = 1024 voxels = 1 KiB of output, written as
// two fully coalesced 512-byte store instructions (lane j writes 16-byte pieces j and 32+j).
// grid = (gw / kBlock, F): blockIdx.y is the frame, so no 64-bit division is needed.
__global__ void __launch_bounds__(kBlock)
k_emit_dense(uint32_t* __restrict__ bitmap, uint32_t* __restrict__ prefix, u64* __restrict__ vslot,
             const int64_t* __restrict__ off, const uint8_t* __restrict__ remap, uint8_t* __restrict__ dense, GridDev g, int F,
             bool clean) {
  const int f = blockIdx.y;
  const uint32_t wi = blockIdx.x * kBlock + threadIdx.x;        // word index inside the frame
  const bool valid = wi < (uint32_t)g.gw;
  const int64_t wg = (int64_t)f * g.gw + wi;
  uint32_t word, rank;
  load_word_and_rank(bitmap, prefix, wg, valid, clean, &word, &rank);
  const unsigned lane = lane_id();
  const uint32_t warp_w0 = wi - lane;                           // gw % 32 == 0: a warp never straddles frames
  if (warp_w0 >= (uint32_t)g.gw) return;                        // whole warp out of range
  u64* vslot_f = vslot + __ldg(off + f);
  const uint32_t vox0 = warp_w0 * 32u;                          // first voxel of the warp inside the frame
  uint8_t* dst = dense + (size_t)f * g.G + vox0;
  const bool fast = ((g.G & 15) == 0) && ((reinterpret_cast<uintptr_t>(dense) & 15) == 0);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    unsigned src = (unsigned)half * 16u + (lane >> 1);          // lane j handles piece p = half*32 + j -> word p/2
    uint32_t w = __shfl_sync(0xffffffffu, word, src);
    uint32_t rk = __shfl_sync(0xffffffffu, rank, src);
    uint32_t bits = (lane & 1u) ? (w >> 16) : (w & 0xffffu);
    if (lane & 1u) rk += __popc(w & 0xffffu);
    uint4 o = expand_half(bits, rk, vslot_f, remap, clean);
    __syncwarp();   // reconverge after the data-dependent gathers so that the store below is one 512-byte request
    const uint32_t piece = (uint32_t)half * 32u + lane;
    const int64_t v = (int64_t)vox0 + piece * 16;               // first voxel of this piece
    if (fast) {
      if (v + 16 <= g.G) st_stream_u4(reinterpret_cast<uint4*>(dst + piece * 16), o);
    } else {
      uint32_t oo[4] = {o.x, o.y, o.z, o.w};
      for (int j = 0; j < 16; ++j)
        if (v + j < g.G) dense[(size_t)f * g.G + v + j] = (uint8_t)(oo[j >> 2] >> (8 * (j & 3)));
    }
  }
  if (clean && valid && word) bitmap[wg] = 0u;
}

// Sparse list, bitmap in linear-id order: rows (x,y,z,label) uint16 at sparse[(frame_offsets[f] + rank)].
__global__ void __launch_bounds__(kBlock)
k_emit_sparse(uint32_t* __restrict__ bitmap, uint32_t* __restrict__ prefix, u64* __restrict__ vslot,
              const int64_t* __restrict__ off, uint16_t* __restrict__ sparse, GridDev g, int F, bool clean) {
  int64_t wg = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  int64_t total = (int64_t)F * g.gw;
  bool valid = wg < total;
  uint32_t word, rank;
  load_word_and_rank(bitmap, prefix, wg, valid, clean, &word, &rank);
  if (!valid || !word) return;
  int f = (int)(wg / g.gw);
  int64_t fbeg = __ldg(off + f);
  uint32_t bit0 = (uint32_t)(wg - (int64_t)f * g.gw) * 32u;
  uint32_t b = word;
  while (b) {
    int j = __ffs(b) - 1;
    b &= b - 1;
    uint32_t lin = bit0 + (uint32_t)j;
    uint32_t x = lin % (uint32_t)g.dx;
    uint32_t yz = lin / (uint32_t)g.dx;
    uint32_t y = yz % (uint32_t)g.dy, z = yz / (uint32_t)g.dy;
    u64* sl = vslot + fbeg + rank;
    uint32_t lab = (uint32_t)(*sl) & 0xffu;
    if (clean) *sl = 0ull;
    if (sparse) {
      uint2 row = make_uint2(x | (y << 16), z | (lab << 16));
      *reinterpret_cast<uint2*>(sparse + (size_t)(fbeg + rank) * 4) = row;
    }
    ++rank;
  }
  if (clean) bitmap[wg] = 0u;
}

// Dense grid when the bitmap is in linear-id order (both outputs requested): per-voxel lookup.
__global__ void __launch_bounds__(kBlock)
k_emit_dense_from_linear(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix,
                         const u64* __restrict__ vslot, const int64_t* __restrict__ off, const uint8_t* __restrict__ remap,
                         uint8_t* __restrict__ dense, GridDev g, int F) {
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
    uint32_t rank = rank_of(bm, prefix + (size_t)f * (g.gw / 4), lin);
    lab = (uint32_t)vslot[__ldg(off + f) + rank] & 0xffu;
    if (remap) lab = __ldg(remap + lab);
  }
  dense[t] = (uint8_t)lab;
}

// Range image: one thread per NP pixels (NP = 4: 16-byte stores; NP = 1: generic fallback).
template <typename T, int NP, int LAYOUT>
__global__ void __launch_bounds__(kBlock)
k_emit_range(u64* __restrict__ pixtab, const T* __restrict__ xyz, const uint8_t* __restrict__ sem,
             const int64_t* __restrict__ off, RangeDev r, int F, float* __restrict__ depth_out, float* __restrict__ xyz_out,
             uint8_t* __restrict__ sem_out, bool clean) {
  const int64_t HW = (int64_t)r.H * r.W;
  int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  int64_t p0 = t * NP;                 // global pixel index over [F, H*W]
  if (p0 >= (int64_t)F * HW) return;
  int f = (int)(p0 / HW);
  int64_t pin = p0 - (int64_t)f * HW;  // pixel inside the frame
  int64_t fbeg = __ldg(off + f);
  u64 wv[NP];
  if (NP == 4) {
    ulonglong2 a = *reinterpret_cast<ulonglong2*>(pixtab + p0);
    ulonglong2 b = *reinterpret_cast<ulonglong2*>(pixtab + p0 + 2);
    wv[0] = a.x; wv[1 % NP] = a.y; wv[2 % NP] = b.x; wv[3 % NP] = b.y;
    if (clean) {
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
      double a, b, c;
      pd[k] = (float)range_depth_of(x, y, z, r, &a, &b, &c);   // :217 float32(depth64)
      px[k] = (float)x; py[k] = (float)y; pz[k] = (float)z;    // :218 the ego-frame input point
      ps |= (uint32_t)__ldg(sem + qi) << (8 * k);              // :219
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

// ---------------------------------------------------------------- host side
static bool is_pow2_double(double v) {
  if (!(v > 0.0) || !isfinite(v)) return false;
  int e;
  return frexp(v, &e) == 0.5;
}

static int make_grid_dev(const MuvoGrid* g, int order, GridDev* o) {
  if (g->size[0] <= 0 || g->size[1] <= 0 || g->size[2] <= 0 || !(g->res > 0.0)) return MUVO_E_ARG;
  if (g->size[0] > 65535 || g->size[1] > 65535 || g->size[2] > 65535) return MUVO_E_SHAPE;   // uint16 coordinates (:195)
  int64_t G = (int64_t)g->size[0] * g->size[1] * g->size[2];
  if (G > ((int64_t)1 << 31) - 1024) return MUVO_E_SHAPE;
  o->res = g->res; o->inv_res = 1.0 / g->res;
  for (int k = 0; k < 3; ++k) { o->off[k] = g->offset[k]; o->up[k] = g->upper[k]; }
  o->dx = g->size[0]; o->dy = g->size[1]; o->dz = g->size[2];
  o->road = g->roadline_id;
  o->pow2 = is_pow2_double(g->res) ? 1 : 0;
  o->order = order;
  o->gw = bitmap_words(G);
  o->G = G;
  return MUVO_OK;
}

static int make_range_dev(const MuvoRangeCfg* c, RangeDev* o) {
  if (c->H <= 0 || c->W <= 0 || !(c->fov != 0.0)) return MUVO_E_ARG;
  if ((int64_t)c->H * c->W > ((int64_t)1 << 30)) return MUVO_E_SHAPE;
  o->H = c->H; o->W = c->W; o->fda = c->fov_down_abs; o->fov = c->fov;
  o->inv_pi_f = (float)(1.0 / kPi); o->fda_f = (float)c->fov_down_abs; o->inv_fov_f = (float)(1.0 / c->fov);
  for (int k = 0; k < 3; ++k) o->L[k] = c->lidar_pos[k];
  return MUVO_OK;
}

static inline unsigned blocks_for(int64_t n) { return (unsigned)ceil_div64(n, kBlock); }

template <typename T>
static int run_points(const T* xyz, const uint8_t* sem, const int64_t* off, int F, int64_t P, const MuvoGrid* grid_h,
                      const uint8_t* remap, const MuvoRangeCfg* cfg_h, int layout, uint8_t* dense, uint16_t* sparse,
                      int64_t* n_occ, float* depth_out, float* xyz_out, uint8_t* sem_out, int64_t* diag, void* ws,
                      size_t ws_bytes, cudaStream_t st) {
  const bool do_vox = grid_h != nullptr, do_range = cfg_h != nullptr;
  if (!do_vox && !do_range) return MUVO_E_ARG;
  if (F < 0 || P < 0) return MUVO_E_ARG;
  if (F == 0) return MUVO_OK;
  if (F > 65535) return MUVO_E_SHAPE;   // frames are a grid dimension of the emit kernels
  if (!off || !ws) return MUVO_E_NULL;
  if (P > 0 && (!xyz || !sem)) return MUVO_E_NULL;
  if (P >= ((int64_t)1 << 40)) return MUVO_E_SHAPE;
  if (do_vox && !dense && !sparse && !n_occ) return MUVO_E_NULL;
  if (do_range && (!xyz_out || (layout == MUVO_RANGE_LAYOUT_HWC && (!depth_out || !sem_out)))) return MUVO_E_NULL;
  if (layout != MUVO_RANGE_LAYOUT_HWC && layout != MUVO_RANGE_LAYOUT_XYZD) return MUVO_E_ARG;
  if (reinterpret_cast<uintptr_t>(ws) & 255) return MUVO_E_ALIGN;
  GridDev g{}; RangeDev r{};
  int rc;
  const int order = (sparse != nullptr) ? ORDER_LINEAR : ORDER_DENSE;
  if (do_vox && (rc = make_grid_dev(grid_h, order, &g)) != MUVO_OK) return rc;
  if (do_range && (rc = make_range_dev(cfg_h, &r)) != MUVO_OK) return rc;
  PointsWs w = carve(ws, P, F, grid_h, cfg_h);
  if (w.bytes > ws_bytes) return MUVO_E_WORKSPACE;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(xyz) % 16 == 0) && (reinterpret_cast<uintptr_t>(sem) % 4 == 0);
  constexpr int KPL = MUVO_KPL;                                  // points per lane (lock-step ILP vs occupancy)
  const unsigned pblocks = (unsigned)ceil_div64(P, (int64_t)kWarpsPerBlock * 32 * KPL);
  const bool packl = P < ((int64_t)1 << 24) - 1;                 // label rides in the voxel word (24-bit index)

  prof_mark("<points>", st);
  // K1
  if (P > 0) {
    if (do_vox && do_range)
      k_point_pass<T, KPL, true, true><<<pblocks, kBlock, 0, st>>>(xyz, sem, off, F, P, vec_ok, g, r, w.bitmap, w.pixtab, diag);
    else if (do_vox)
      k_point_pass<T, KPL, true, false><<<pblocks, kBlock, 0, st>>>(xyz, sem, off, F, P, vec_ok, g, r, w.bitmap, w.pixtab, diag);
    else
      k_point_pass<T, KPL, false, true><<<pblocks, kBlock, 0, st>>>(xyz, sem, off, F, P, vec_ok, g, r, w.bitmap, w.pixtab, diag);
    MUVO_AFTER_LAUNCH("k_point_pass", st);
  }
  if (do_vox) {
    // K2
    k_bitmap_scan<<<F, kScanThreads, 0, st>>>(w.bitmap, w.prefix, g.gw, n_occ);
    MUVO_AFTER_LAUNCH("k_bitmap_scan", st);
    // K3 (+ K4 when the label does not fit in the slot word)
    if (P > 0) {
      if (packl) {
        k_voxel_resolve<T, KPL, true><<<pblocks, kBlock, 0, st>>>(xyz, sem, off, F, P, vec_ok, g, w.bitmap, w.prefix, w.vslot);
        MUVO_AFTER_LAUNCH("k_voxel_resolve", st);
      } else {
        k_voxel_resolve<T, KPL, false><<<pblocks, kBlock, 0, st>>>(xyz, sem, off, F, P, vec_ok, g, w.bitmap, w.prefix, w.vslot);
        MUVO_AFTER_LAUNCH("k_voxel_resolve", st);
        k_slot_labels<<<blocks_for(P), kBlock, 0, st>>>(w.vslot, sem, off, F, P);
        MUVO_AFTER_LAUNCH("k_slot_labels", st);
      }
    }
    // K5 (the last consumer of the tables clears them)
    const int64_t words = (int64_t)F * g.gw;
    if (order == ORDER_DENSE) {
      if (dense) {
        dim3 grid((unsigned)ceil_div64(g.gw, kBlock), (unsigned)F);
        k_emit_dense<<<grid, kBlock, 0, st>>>(w.bitmap, w.prefix, w.vslot, off, remap, dense, g, F, true);
      } else {  // only n_occ requested: clear through the sparse walker without output
        k_emit_sparse<<<blocks_for(words), kBlock, 0, st>>>(w.bitmap, w.prefix, w.vslot, off, nullptr, g, F, true);
      }
      MUVO_AFTER_LAUNCH(dense ? "k_emit_dense" : "k_emit_sparse", st);
    } else {
      if (dense) {
        k_emit_dense_from_linear<<<blocks_for((int64_t)F * g.G), kBlock, 0, st>>>(w.bitmap, w.prefix, w.vslot, off,
                                                                                 remap, dense, g, F);
        MUVO_AFTER_LAUNCH("k_emit_dense_from_linear", st);
      }
      k_emit_sparse<<<blocks_for(words), kBlock, 0, st>>>(w.bitmap, w.prefix, w.vslot, off, sparse, g, F, true);
      MUVO_AFTER_LAUNCH("k_emit_sparse", st);
    }
  }
  if (do_range) {
    const int64_t HW = (int64_t)r.H * r.W;
    const bool vec4 = (HW % 4 == 0) && (reinterpret_cast<uintptr_t>(xyz_out) % 16 == 0) &&
                      (!depth_out || reinterpret_cast<uintptr_t>(depth_out) % 16 == 0) &&
                      (!sem_out || reinterpret_cast<uintptr_t>(sem_out) % 4 == 0);
    const int64_t npix = (int64_t)F * HW;
    if (vec4) {
      if (layout == MUVO_RANGE_LAYOUT_HWC)
        k_emit_range<T, 4, MUVO_RANGE_LAYOUT_HWC><<<blocks_for(npix / 4), kBlock, 0, st>>>(w.pixtab, xyz, sem, off, r, F, depth_out, xyz_out, sem_out, true);
      else
        k_emit_range<T, 4, MUVO_RANGE_LAYOUT_XYZD><<<blocks_for(npix / 4), kBlock, 0, st>>>(w.pixtab, xyz, sem, off, r, F, depth_out, xyz_out, sem_out, true);
    } else {
      if (layout == MUVO_RANGE_LAYOUT_HWC)
        k_emit_range<T, 1, MUVO_RANGE_LAYOUT_HWC><<<blocks_for(npix), kBlock, 0, st>>>(w.pixtab, xyz, sem, off, r, F, depth_out, xyz_out, sem_out, true);
      else
        k_emit_range<T, 1, MUVO_RANGE_LAYOUT_XYZD><<<blocks_for(npix), kBlock, 0, st>>>(w.pixtab, xyz, sem, off, r, F, depth_out, xyz_out, sem_out, true);
    }
    MUVO_AFTER_LAUNCH("k_emit_range", st);
  }
  return MUVO_OK;
}

}  // namespace
}  // namespace muvo

using namespace muvo;

extern "C" {

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
                  int32_t n_frames, int64_t n_points_total, const MuvoGrid* grid_h, const uint8_t* remap256,
                  uint8_t* dense_out, uint16_t* sparse_out, int64_t* n_occ_out, int64_t* diag, void* ws,
                  size_t ws_bytes, void* stream) {
  if (!grid_h) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (xyz_dtype == MUVO_F32)
    return run_points<float>((const float*)xyz, sem, frame_offsets, n_frames, n_points_total, grid_h, remap256, nullptr,
                             0, dense_out, sparse_out, n_occ_out, nullptr, nullptr, nullptr, diag, ws, ws_bytes, st);
  if (xyz_dtype == MUVO_F64)
    return run_points<double>((const double*)xyz, sem, frame_offsets, n_frames, n_points_total, grid_h, remap256,
                              nullptr, 0, dense_out, sparse_out, n_occ_out, nullptr, nullptr, nullptr, diag, ws,
                              ws_bytes, st);
  return MUVO_E_ARG;
}

int muvo_range_project(const float* xyz, const uint8_t* sem, const int64_t* frame_offsets, int32_t n_frames,
                       int64_t n_points_total, const MuvoRangeCfg* cfg_h, int32_t layout, float* depth_out,
                       float* xyz_out, uint8_t* sem_out, int64_t* diag, void* ws, size_t ws_bytes, void* stream) {
  if (!cfg_h) return MUVO_E_NULL;
  return run_points<float>(xyz, sem, frame_offsets, n_frames, n_points_total, nullptr, nullptr, cfg_h, layout, nullptr,
                           nullptr, nullptr, depth_out, xyz_out, sem_out, diag, ws, ws_bytes, (cudaStream_t)stream);
}

int muvo_points_fused(const float* xyz, const uint8_t* sem, const int64_t* frame_offsets, int32_t n_frames,
                      int64_t n_points_total, const MuvoGrid* grid_h, const uint8_t* remap256,
                      const MuvoRangeCfg* cfg_h, int32_t layout, uint8_t* dense_out, uint16_t* sparse_out,
                      int64_t* n_occ_out, float* depth_out, float* xyz_out, uint8_t* sem_out, int64_t* diag, void* ws,
                      size_t ws_bytes, void* stream) {
  if (!grid_h || !cfg_h) return MUVO_E_NULL;
  return run_points<float>(xyz, sem, frame_offsets, n_frames, n_points_total, grid_h, remap256, cfg_h, layout,
                           dense_out, sparse_out, n_occ_out, depth_out, xyz_out, sem_out, diag, ws, ws_bytes,
                           (cudaStream_t)stream);
}

}  // extern "C"

// Device / host helpers shared by the point kernels (points.cu: sorted sparse lists, generic grids, float64 clouds;
// points_mega.cu: the single-launch dataflow kernel for dense grids + range images).  Internal header.
// Every translation unit that includes this file must be compiled with -fmad=false (numpy's operation order).
#pragma once
#include <math.h>
#include "common.cuh"
#include "tma.cuh"

namespace muvo {

constexpr int kBlock = 256;
constexpr int kMaxChunks = 1;                         // point sub-ranges per call (the kernels take [p0, p1) x [f0, f1))
constexpr int kMaxTileCtas = 4096;                   // upper bound of the persistent tile grid (queue length slots)
constexpr double kPi = 3.141592653589793;            // np.pi
constexpr double kPiOver4 = 0x1.921fb54442d18p-1;    // correctly rounded pi/4 (numpy/glibc value on diagonals)
constexpr double k3PiOver4 = 0x1.2d97c7f3321d2p+1;   // correctly rounded 3pi/4
constexpr double kEdgeEps = 1e-9;                    // diagnostics: "on a bin edge"

enum BitOrder { ORDER_DENSE = 0 /* (x*Dy+y)*Dz+z */, ORDER_LINEAR = 1 /* x + Dx*(y + Dy*z) */ };

struct GridDev {
  double res, inv_res;
  double off[3], up[3];
  int dx, dy, dz;
  int road;
  int pow2;
  int regular;   // pow2 res and upper == size*res exactly: in-grid test and voxel id from one floor per axis
  int order;
  uint32_t sx, sy, sz;   // bit index = ix*sx + iy*sy + iz*sz (per `order`)
  int gw;        // bitmap words per frame (multiple of 32)
  int64_t G;     // voxels per frame
};

struct RangeDev {
  int H, W;
  double fda, fov;
  double L[3];
  // f32 fast path (see pix_fast): pw = (1 - yaw/pi) * half_w ; ph = h_bias - (pitch/pi) * h_scale
  float half_w, h_scale, h_bias;
  float eps_w, eps_h;     // distance to a bin edge (in bins) below which the float64 formula decides
  float w_hi, h_hi;       // W - 0.5, H - 0.5: beyond these (or below 0.5) both paths clamp to the border bin
  float w_max, h_max;     // W - 1, H - 1
  float safe_w, safe_h;   // 0.5 - eps: |frac - 0.5| below this = provably inside the bin
  float Lf[3];            // sensor position in f32
  int lf_exact;           // ... and whether that is exact
  // LiDAR-side prep in front of the projection (muvo/data/dataset.py:278-290), see lidar_prep(); prep = 0: points are taken as is
  int prep, prep_box;
  double padd[3];         // convert_coor_lidar: += lidar_pos (data/data_preprocessing.py:119-122)
  double blo[3], bhi[3];  // ego box (dataset.py:286-288)
};

typedef unsigned long long u64;

struct PointsWs {
  uint32_t* bitmap = nullptr;   // [F, gw]    occupancy, 1 bit per voxel
  uint32_t* prefix = nullptr;   // [F, gw/4]  exclusive popcount prefix per 128-bit chunk (sparse output only)
  u64* vtab = nullptr;          // [F, G]     packed winner per voxel, addressed by the voxel's bit index (0 = empty)
  u64* pixtab = nullptr;        // [F, H*W]   packed winner per pixel (0 = empty)
  uint32_t* qcount = nullptr;   // [kMaxChunks, kMaxTileCtas, 2] (length, frame of the CTA's first point): rare-path queue length per tile CTA (rewritten by every call)
  uint2* queue = nullptr;       // [2P]       rare-path queue, CTA b's segment starts at 2 * (its first point)
  uint32_t* sync = nullptr;     // [6F] per-frame claim / completion counters of points_mega.cu
  double* edges = nullptr;      // [2 (W + 1) + H + 1] range-image bin edges (points_mega.cu)
  size_t bytes = 0;
};

// the dataflow kernel of points_mega.cu (dense grids / range images of float32 clouds); see there for eligibility
bool points_mega_eligible(const GridDev* g, const RangeDev* r, const void* xyz, const void* sem, const void* dense,
                          const void* depth_out, const void* xyz_out, const void* sem_out, int layout);
int points_mega_f32(const float* xyz, const uint8_t* sem, const int64_t* off, int F, int64_t P, const GridDev* g,
                    const uint8_t* remap, const RangeDev* r, int layout, uint8_t* dense, int64_t* n_occ, float* depth_out,
                    float* xyz_out, uint8_t* sem_out, int64_t* diag, const PointsWs& w, cudaStream_t st);

namespace {

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
    w.vtab = (u64*)(b + o); o = align_up(o + (size_t)F * (size_t)G * 8, 256);
  }
  if (r) {
    w.pixtab = (u64*)(b + o); o = align_up(o + (size_t)F * r->H * r->W * 8, 256);
  }
  w.qcount = (uint32_t*)(b + o); o = align_up(o + (size_t)kMaxChunks * kMaxTileCtas * 8, 256);
  w.sync = (uint32_t*)(b + o); o = align_up(o + ((size_t)6 * F + 64) * 4, 256);
  if (r) { w.edges = (double*)(b + o); o = align_up(o + ((size_t)2 * (r->W + 1) + r->H + 1) * 8, 256); }
  w.queue = (uint2*)(b + o); o = align_up(o + (size_t)(P > 0 ? P : 1) * 16, 256);
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
// The label-carrying format is used per FRAME, whenever the frame has fewer than 2^24 - 1 points (decided on the
// device from frame_offsets); larger frames keep the 32-bit index and the emit kernels fetch the label through it.
constexpr int64_t kPackLimit = ((int64_t)1 << 24) - 1;
__device__ __forceinline__ u64 vox_word(bool packl, uint32_t top_inv, uint32_t idx1, uint32_t label) {
  return packl ? pack_vox(top_inv, idx1, label) : pack_word(top_inv, idx1);
}
__device__ __forceinline__ uint32_t vox_word_idx1(bool packl, u64 w) { return packl ? vox_idx1(w) : word_idx1(w); }
__device__ __forceinline__ uint32_t vox_word_label(bool packl, u64 w, const uint8_t* __restrict__ sem_f) {
  return packl ? ((uint32_t)w & 0xffu) : (uint32_t)__ldg(sem_f + (word_idx1(w) - 1u));
}

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

// voxel id for a "regular" grid (power-of-two res, upper == size*res): floor((p+off)/res) is exact, so
// 0 <= b < upper  <=>  0 <= floor(b/res) < size  (:178,:183).  The floor comes from the "add 2^52, round down" trick:
// for 0 <= q < 2^32 the sum stays in [2^52, 2^52 + 2^32), i.e. its high word is exactly 0x43300000 and its low word
// is floor(q); negative, huge, infinite or NaN q change the high word.  No f64 -> int conversions (quarter-rate pipe).
constexpr double kTwo52 = 4503599627370496.0;
struct VoxFast { uint32_t bit; bool in; double bx, by, bz, sx, sy, sz; };
template <typename T>
__device__ __forceinline__ VoxFast vox_regular(T x, T y, T z, const GridDev& g) {
  VoxFast v;
  v.bx = (double)x + g.off[0]; v.by = (double)y + g.off[1]; v.bz = (double)z + g.off[2];   // :177
  v.sx = __dadd_rd(v.bx * g.inv_res, kTwo52); v.sy = __dadd_rd(v.by * g.inv_res, kTwo52); v.sz = __dadd_rd(v.bz * g.inv_res, kTwo52);
  const uint32_t ix = (uint32_t)__double2loint(v.sx), iy = (uint32_t)__double2loint(v.sy), iz = (uint32_t)__double2loint(v.sz);
  v.in = (__double2hiint(v.sx) == 0x43300000) & (__double2hiint(v.sy) == 0x43300000) & (__double2hiint(v.sz) == 0x43300000) &
         (ix < (uint32_t)g.dx) & (iy < (uint32_t)g.dy) & (iz < (uint32_t)g.dz);
  v.bit = ix * g.sx + iy * g.sy + iz * g.sz;
  return v;
}
// |p mod res|^2 for an in-grid point of a regular grid: b - floor(b/res)*res is exact (so the fma equals numpy's
// mul + sub), then numpy's (mx^2 + my^2) + mz^2 with separate roundings (:212)
__device__ __forceinline__ double vox_regular_dis(const VoxFast& v, const GridDev& g) {
  const double mx = __fma_rn(kTwo52 - v.sx, g.res, v.bx), my = __fma_rn(kTwo52 - v.sy, g.res, v.by),
               mz = __fma_rn(kTwo52 - v.sz, g.res, v.bz);
  return (mx * mx + my * my) + mz * mz;
}

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

// LiDAR-side prep (N1), fused into every kernel that loads a raw point: convert_coor_lidar (`pcd += lidar_pos` on a float32
// array = float32(float64(p) + lidar_pos), then `pcd[:, 1] *= -1`, data/data_preprocessing.py:119-122) and the ego-box drop
// (`(lo < p) & (p < hi)` in float64, muvo/data/dataset.py:286-290).  Returns false for a point inside the box (dropped
// before the projection in the reference: the order of the survivors is unchanged, so ties resolve the same way).
__device__ __forceinline__ bool lidar_prep(float& x, float& y, float& z, const RangeDev& r) {
  x = (float)((double)x + r.padd[0]); y = -(float)((double)y + r.padd[1]); z = (float)((double)z + r.padd[2]);
  if (!r.prep_box) return true;
  const double dx = (double)x, dy = (double)y, dz = (double)z;
  return !((r.blo[0] < dx) & (dx < r.bhi[0]) & (r.blo[1] < dy) & (dy < r.bhi[1]) & (r.blo[2] < dz) & (dz < r.bhi[2]));
}
__device__ __forceinline__ bool lidar_prep(double&, double&, double&, const RangeDev&) { return true; }   // (float32 clouds only)

// ---- f32 fast path of the pixel computation
// atan(t)/pi on [0,1] as t*Q(t^2), Q of degree 6 (tools/fit_atan.py: |error| < 1.2e-7 including the f32 Horner
// rounding).  Together with the quotient (2 ulp), the f32 coordinates (0.5 ulp each) and the final scaling the
// column error stays below 2e-4 bins at W = 1024 and the row error below 1e-4 bins at H/fov = 64/40deg; eps_w / eps_h
// (1e-3 bins at those sizes, scaled up for larger images) leave a 5x margin.  tests/test_points_gpu.py checks the
// claim on tens of millions of points through muvo_debug_pixel_check.
__device__ __forceinline__ float atan_over_pi_unit(float t) {
  const float t2 = __fmul_rn(t, t);
  float q = 0.002143867f;
  q = __fmaf_rn(q, t2, -0.010623431f);
  q = __fmaf_rn(q, t2, 0.025261912f);
  q = __fmaf_rn(q, t2, -0.042078603f);
  q = __fmaf_rn(q, t2, 0.063039005f);
  q = __fmaf_rn(q, t2, -0.10605131f);
  q = __fmaf_rn(q, t2, 0.31830862f);
  return __fmul_rn(q, t);
}
// atan2(y, x)/pi in [-1, 1]; NaN when x == y == 0 (the caller then takes the exact path)
__device__ __forceinline__ float atan2_over_pi(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
  float rc;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(mx));
  float p = atan_over_pi_unit(__fmul_rn(mn, rc));
  if (ay > ax) p = 0.5f - p;
  if (x < 0.f) p = 1.0f - p;
  return copysignf(p, y);
}

// fractional pixel coordinates in f32 from the sensor-frame point (xf, yf = -y_c, zf)
__device__ __forceinline__ void pix_coords_f32(float xf, float yf, float zf, const RangeDev& r, float* pw_o, float* ph_o) {
  *pw_o = __fmul_rn(1.0f - atan2_over_pi(yf, xf), r.half_w);
  float rho;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rho) : "f"(__fmaf_rn(xf, xf, __fmul_rn(yf, yf))));
  *ph_o = __fmaf_rn(-atan2_over_pi(zf, rho), r.h_scale, r.h_bias);
}

struct PixFast { int pix; bool ok; bool slow; double s; };
// s = squared range (float64, reference order): orders like the depth; `ok` = finite and non-zero; `slow` = the f32
// pixel is not provably the float64 one (near an edge, NaN, or magnitudes where f32 squares over/underflow).
template <typename T>
__device__ __forceinline__ PixFast pix_fast(T x, T y, T z, const RangeDev& r) {
  PixFast k;
  double xc, yc, zc;
  k.s = range_sq_of(x, y, z, r, &xc, &yc, &zc);
  const uint32_t hi = (uint32_t)__double2hiint(k.s);
  k.ok = (k.s > 0.0) & (hi < 0x7ff00000u);
  float xf, yf, zf;
  if (sizeof(T) == 4 && r.lf_exact) {   // sensor position exact in f32: the f32 difference is the correctly rounded one
    xf = (float)x - r.Lf[0]; yf = -((-(float)y) - r.Lf[1]); zf = (float)z - r.Lf[2];   // same zero signs as :177-183
  } else {
    xf = (float)xc; yf = (float)(-yc); zf = (float)zc;
  }
  float pw, ph;
  pix_coords_f32(xf, yf, zf, r, &pw, &ph);
  const float fw = floorf(pw), fh = floorf(ph);
  const float dw = pw - fw, dh = ph - fh;
  const bool safe_w = (pw <= 0.5f) | (pw >= r.w_hi) | ((dw > r.eps_w) & (dw < 1.0f - r.eps_w));
  const bool safe_h = (ph <= 0.5f) | (ph >= r.h_hi) | ((dh > r.eps_h) & (dh < 1.0f - r.eps_h));
  // 1e-12 < s < 1e12 (metres^2): inside, no f32 square over/underflows in a way that could move a pixel
  const bool mag_ok = (hi - 0x3d719799u) < (0x426d1a94u - 0x3d719799u);
  k.slow = !(safe_w & safe_h & mag_ok);     // NaN compares false -> slow
  const int iw = (int)fminf(fmaxf(fw, 0.0f), r.w_max);
  const int ih = (int)fminf(fmaxf(fh, 0.0f), r.h_max);
  k.pix = ih * r.W + iw;
  return k;
}

// Pixel of a point whose f32 estimate lies within eps of a bin edge (pix_fast's `slow`), decided in float64 against the edge
// itself instead of through atan2 / asin: column edge e <=> yaw = pi (1 - 2e/W), and the sign of the cross product of the
// edge direction with (x_c, -y_c) says on which side the point lies; row edge e <=> pitch = fov (1 - e/H) - |fov_down|:
// compare z_c with depth * sin(edge pitch).  The reference's own float64 rounding moves a coordinate by < 1e-12 bins, so
// outside a 1e-9 (relative) zone around the edge the geometric side IS the reference's bin; inside it, or for magnitudes
// where the f32 estimate itself is not trustworthy, the answer is kPixUndecided and the caller evaluates the reference
// formula (pix_exact: CUDA's atan2 / asin, near-edge diagnostics counted).  Used by the rare-path queue kernel: ~50x
// fewer dependent float64 instructions than the libm path and no local-memory stack traffic.
constexpr uint32_t kPixUndecided = 0xfffffffeu;
template <typename T>
__device__ __forceinline__ uint32_t pix_refine(T x, T y, T z, const RangeDev& r) {
  double xc, yc, zc;
  const double s = range_sq_of(x, y, z, r, &xc, &yc, &zc);
  const uint32_t hi = (uint32_t)__double2hiint(s);
  if (!((hi - 0x3d719799u) < (0x426d1a94u - 0x3d719799u))) return kPixUndecided;      // 1e-12 < s < 1e12, finite
  float pw, ph;
  pix_coords_f32((float)xc, (float)(-yc), (float)zc, r, &pw, &ph);
  if (!(pw > -8.f && pw < (float)r.W + 8.f && ph > -1e6f && ph < 1e6f)) return kPixUndecided;
  const float fw = floorf(pw), fh = floorf(ph);
  int iw = (int)fw, ih = (int)fh;
  const float dw = pw - fw, dh = ph - fh;
  if (!(fabsf(dw - 0.5f) < r.safe_w)) {
    const int e = dw < 0.5f ? iw : iw + 1;                 // the edge between columns e - 1 and e
    if (e > 0 && e < r.W) {
      double sn, cs;
      sincospi(1.0 - 2.0 * (double)e / (double)r.W, &sn, &cs);
      const double yy = -yc;
      const double cross = cs * yy - sn * xc;              // |p| sin(yaw - edge yaw)
      if (!(fabs(cross) > 1e-9 * (fabs(xc) + fabs(yy)))) return kPixUndecided;
      iw = cross > 0.0 ? e - 1 : e;
    }
  }
  if (!(fabsf(dh - 0.5f) < r.safe_h)) {
    const int e = dh < 0.5f ? ih : ih + 1;                 // the edge between rows e - 1 and e
    if (e > 0 && e < r.H) {
      const double depth = sqrt(s);
      const double t = zc - depth * sinpi((r.fov * (1.0 - (double)e / (double)r.H) - r.fda) * (1.0 / kPi));
      if (!(fabs(t) > 1e-9 * depth)) return kPixUndecided;
      ih = t > 0.0 ? e - 1 : e;
    }
  }
  iw = iw < 0 ? 0 : (iw > r.W - 1 ? r.W - 1 : iw);
  ih = ih < 0 ? 0 : (ih > r.H - 1 ? r.H - 1 : ih);
  return (uint32_t)(ih * r.W + iw);
}

__device__ __forceinline__ int find_frame(const int64_t* __restrict__ off, int F, int64_t i) {
  int lo = 0, hi = F;   // largest f with off[f] <= i
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
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

// Programmatic dependent launch: the kernels of one call are chained with cudaLaunchAttributeProgrammaticStreamSerialization.
// `pdl_wait` returns once the previous kernel of the stream has completed and its writes are visible (a no-op for a kernel
// launched without the attribute); `pdl_launch` lets the next kernel's CTAs be scheduled as soon as every CTA of this grid
// has called it or exited, so that launch latency and prologue overlap this grid's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- K1: point pass
// The tables live in global memory; explicit .global atomics keep the compiler from emitting generic-address atomics
// (address-space test + a shared-memory CAS loop per atomic) once a pointer has been through an opaque asm.
__device__ __forceinline__ u64 atom_max_global(u64* p, u64 v) {
  u64 old;
  asm volatile("atom.global.max.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
  return old;
}
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ u64 atom_max_global_hint(u64* p, u64 v, uint64_t pol) {
  u64 old;
  asm volatile("atom.global.max.L2::cache_hint.u64 %0, [%1], %2, %3;" : "=l"(old) : "l"(p), "l"(v), "l"(pol) : "memory");
  return old;
}
__device__ __forceinline__ void red_or_global(uint32_t* p, uint32_t v) {
  asm volatile("red.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Dense grid, bitmap in dense order.  A lane owns one bitmap word = 32 voxels = 32 output bytes, written with ONE
// 256-bit store (sm_100 st.global.v8.b32), so a warp store is a fully coalesced 1 KiB; zeros are part of the
// store (no memset + scatter).  The winner word of every set bit is fetched from (and cleared in) the voxel table,
// up to 4 independent gathers in flight per lane.
// grid = (gw / kBlock, F): blockIdx.y is the frame, so no 64-bit division is needed.
// Table entries are cleared a whole 32-byte sector at a time (one store, no read-modify-write in L2), and only after
// the loads of that sector have been consumed: 8-byte clears issued right behind the gathers were measured at 19x the
// kernel time of the sparse emit.
__device__ __forceinline__ void st_zero_sector(void* p32) {
  asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p32), "r"(0u) : "memory");
}
__device__ __forceinline__ void st_stream_u8x32(void* p, const uint32_t (&o)[8]) {
  asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
               "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
}
constexpr int kEmitWords = 2;       // bitmap words per lane: a warp owns 64 words = 2048 voxels = 2 KiB of output (3+ words per lane is 5x slower: measured)
constexpr int kEmitList = 512;     // set-bit positions staged per warp (more than that: rounds)
struct EmitSmem {
  uint8_t tile[kBlock / 32][kEmitWords * 1024];   // per warp: the 4 KiB it is about to store, assembled from label bytes
  uint16_t list[kBlock / 32][kEmitList];          // per warp: positions (voxel offset inside the warp's span) of set bits
  uint32_t occ[kBlock / 32];
};
// byte `pos` of a warp tile: rows of 32 bytes (one bitmap word); the two 16-byte halves of a row are swapped on every
// other group of 4 rows so that the 128-bit row reads of 8 consecutive lanes hit 8 different bank groups
__device__ __forceinline__ uint32_t tile_swz(uint32_t pos) { return pos ^ ((pos >> 3) & 16u); }

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
  o->regular = o->pow2;
  for (int k = 0; k < 3; ++k)
    if (!(g->upper[k] == (double)g->size[k] * g->res) || !isfinite(g->offset[k])) o->regular = 0;
  o->order = order;
  if (order == ORDER_DENSE) { o->sx = (uint32_t)(g->size[1] * g->size[2]); o->sy = (uint32_t)g->size[2]; o->sz = 1u; }
  else { o->sx = 1u; o->sy = (uint32_t)g->size[0]; o->sz = (uint32_t)(g->size[0] * g->size[1]); }
  o->gw = bitmap_words(G);
  o->G = G;
  return MUVO_OK;
}

static int make_range_dev(const MuvoRangeCfg* c, RangeDev* o) {
  if (c->H <= 0 || c->W <= 0 || !(c->fov != 0.0)) return MUVO_E_ARG;
  if ((int64_t)c->H * c->W > ((int64_t)1 << 30)) return MUVO_E_SHAPE;
  o->H = c->H; o->W = c->W; o->fda = c->fov_down_abs; o->fov = c->fov;
  o->half_w = (float)(0.5 * c->W);
  o->h_scale = (float)(kPi * c->H / c->fov);
  o->h_bias = (float)((double)c->H * (1.0 - c->fov_down_abs / c->fov));
  const double sw = (double)c->W / 1024.0, sh = fabs((double)c->H / c->fov) / (64.0 / (40.0 * kPi / 180.0));
  o->eps_w = (float)(1e-3 * (sw > 1.0 ? sw : 1.0));
  o->eps_h = (float)(1e-3 * (sh > 1.0 ? sh : 1.0));
  o->w_hi = (float)c->W - 0.5f; o->h_hi = (float)c->H - 0.5f;
  o->w_max = (float)(c->W - 1); o->h_max = (float)(c->H - 1);
  o->safe_w = 0.5f - o->eps_w; o->safe_h = 0.5f - o->eps_h;
  o->lf_exact = 1;
  for (int k = 0; k < 3; ++k) { o->Lf[k] = (float)c->lidar_pos[k]; if ((double)o->Lf[k] != c->lidar_pos[k]) o->lf_exact = 0; }
  for (int k = 0; k < 3; ++k) o->L[k] = c->lidar_pos[k];
  o->prep = 0; o->prep_box = 0;
  return MUVO_OK;
}

}  // namespace
}  // namespace muvo

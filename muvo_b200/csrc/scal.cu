// "Next" row N4: the reductions behind SemScalLoss / GeoScalLoss (muvo/losses.py:191-287).
//
// The reference materialises softmax(prediction) ([F,C,X,Y,Z] fp32), then per class boolean-indexes it with the
// valid mask (another copy) and runs ~6 full reductions -- about (10 + 14 C) passes over the grid for the pair of
// losses.  Every term of both losses is a function of 3C+1 scalars:
//     sum_p[i] = sum over valid voxels of p_i        nom[i] = sum over valid voxels with target == i of p_i
//     cnt[i]   = number of valid voxels with target == i            n_valid
// (e.g. sum((1-p_i)(1-[t==i])) = (n_valid - cnt_i) - (sum_p_i - nom_i)).  One streaming pass computes them with the
// softmax kept in registers (k_scal_fwd), and one more pass writes d loss / d logits from the 2C upstream
// derivatives (k_scal_bwd) -- the scalar algebra in between stays in torch on the [3C+1] tensor.
// HBM bound: (C * sizeof(logit) + 1) B/voxel forward, (2 C * sizeof(logit) + 1) B/voxel backward.
// Sums are accumulated in float64 per thread, reduced per CTA and then across CTAs in a fixed order (deterministic).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace muvo {
namespace {

#ifndef MUVO_SCAL_UNR
#define MUVO_SCAL_UNR 1
#endif
constexpr int kScalThreads = 256;
constexpr int kScalMaxC = 32;
constexpr int kScalMaxCtas = kNumSMsB200 * 8;

template <typename LT> __device__ __forceinline__ float lt_to_f32(LT v);
template <> __device__ __forceinline__ float lt_to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float lt_to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float lt_to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename LT> __device__ __forceinline__ LT f32_to_lt(float v);
template <> __device__ __forceinline__ float f32_to_lt<float>(float v) { return v; }
template <> __device__ __forceinline__ __half f32_to_lt<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 f32_to_lt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// VEC consecutive logits of one class plane (VEC = 4: one 16 B / 8 B load, address aligned by the host-side check)
template <typename LT, int VEC> struct Vec;
template <> struct Vec<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
    float4 v = ld_stream_f4(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    st_stream_f4(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};
template <typename LT> struct Vec16x4 {     // __half / __nv_bfloat16
  static __device__ __forceinline__ void load(const LT* p, float (&o)[4]) {
    uint2 raw;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "l"(p));
    LT h[4];
    *reinterpret_cast<uint2*>(h) = raw;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = lt_to_f32<LT>(h[j]);
  }
  static __device__ __forceinline__ void store(LT* p, const float (&v)[4]) {
    LT h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = f32_to_lt<LT>(v[j]);
    uint2 raw = *reinterpret_cast<uint2*>(h);
    asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(raw.x), "r"(raw.y) : "memory");
  }
};
template <> struct Vec<__half, 4> : Vec16x4<__half> {};
template <> struct Vec<__nv_bfloat16, 4> : Vec16x4<__nv_bfloat16> {};
template <typename LT> struct Vec<LT, 1> {
  static __device__ __forceinline__ void load(const LT* p, float (&o)[1]) { o[0] = lt_to_f32<LT>(*p); }
  static __device__ __forceinline__ void store(LT* p, const float (&v)[1]) { *p = f32_to_lt<LT>(v[0]); }
};

template <int VEC> __device__ __forceinline__ void load_target(const uint8_t* p, int (&t)[VEC]) {
  if constexpr (VEC == 4) {
    uint32_t w = ld_stream_u32(reinterpret_cast<const uint32_t*>(p));
    t[0] = w & 255u; t[1] = (w >> 8) & 255u; t[2] = (w >> 16) & 255u; t[3] = w >> 24;
  } else {
    t[0] = __ldg(p);
  }
}

__device__ __forceinline__ float rcp_fast(float v) {      // den is in [1, C]: MUFU.RCP alone, 1 ulp
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ float exp_nonpos(float v) {    // exp of v <= 0: ex2.approx on v * log2(e), no range fix-up needed
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v * 1.4426950408889634f));
  return r;
}

// softmax over the CT register-resident classes of voxel j (F.softmax in fp32): p[k] = exp(x_k - max) / sum.
// ex2.approx + one reciprocal: a few ulp from torch's expf / divide, far inside the 1e-5 bar, and a third of the
// issue slots (the kernels are issue bound at C = 2 otherwise).
template <int CT, int VEC>
__device__ __forceinline__ void softmax_probs(const float (&x)[CT][VEC], int j, float (&p)[CT]) {
  float m = x[0][j];
#pragma unroll
  for (int k = 1; k < CT; ++k) m = fmaxf(m, x[k][j]);
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < CT; ++k) { p[k] = exp_nonpos(x[k][j] - m); den += p[k]; }
  const float inv = rcp_fast(den);
#pragma unroll
  for (int k = 0; k < CT; ++k) p[k] *= inv;
}

// partial[cta][3C+1]: sum_p[C], nom[C], cnt[C], n_valid  (float64).  grid = (CTAs per frame, frames in flight):
// no 64-bit division in the loop.  The VEC voxels of a group are added in fp32 first, then once into the float64
// accumulators.
template <typename LT, int CT, int VEC, int UNR>
__global__ void __launch_bounds__(kScalThreads)
k_scal_fwd(const LT* __restrict__ logits, const uint8_t* __restrict__ target, int F, int64_t S, int ignore,
           double* __restrict__ partial) {
  double sp[CT], nm[CT];
  unsigned cn[CT], nv = 0;
#pragma unroll
  for (int k = 0; k < CT; ++k) { sp[k] = 0.0; nm[k] = 0.0; cn[k] = 0; }
  const int tid = threadIdx.x;
  const int64_t groups = S / VEC, stride = (int64_t)gridDim.x * kScalThreads;
  for (int f = blockIdx.y; f < F; f += gridDim.y) {
    const LT* lf = logits + (size_t)f * CT * S;
    const uint8_t* tf = target + (size_t)f * S;
    for (int64_t g0 = (int64_t)blockIdx.x * kScalThreads + tid; g0 < groups; g0 += stride * UNR) {
      float x[UNR][CT][VEC];
      int t[UNR][VEC];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {                      // all loads of the UNR groups first
        const int64_t g = g0 + u * stride;
        if (g < groups) {
#pragma unroll
          for (int k = 0; k < CT; ++k) Vec<LT, VEC>::load(lf + (size_t)k * S + g * VEC, x[u][k]);
          load_target<VEC>(tf + g * VEC, t[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (g0 + u * stride >= groups) break;
        float gs[CT], gn[CT];
#pragma unroll
        for (int k = 0; k < CT; ++k) { gs[k] = 0.f; gn[k] = 0.f; }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {                     // branch free: selects keep gs / gn / cn in registers
          const bool valid = t[u][j] != ignore;             // mask = target != ignore_index (:208, :273)
          float p[CT];
          softmax_probs<CT, VEC>(x[u], j, p);
          nv += valid ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < CT; ++k) {
            const bool hit = valid && t[u][j] == k;
            gs[k] += valid ? p[k] : 0.f;
            gn[k] += hit ? p[k] : 0.f;
            cn[k] += hit ? 1u : 0u;
          }
        }
#pragma unroll
        for (int k = 0; k < CT; ++k) { sp[k] += (double)gs[k]; nm[k] += (double)gn[k]; }
      }
    }
  }
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  // CTA reduction: warp shuffles, then warp 0 adds the 8 warp rows in order
  __shared__ double red[kScalThreads / 32][3 * CT + 1];
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    double a = sp[k], b = nm[k];
    unsigned c = cn[k];
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); c += __shfl_xor_sync(0xffffffffu, c, d);
    }
    if (lane == 0) { red[warp][k] = a; red[warp][CT + k] = b; red[warp][2 * CT + k] = (double)c; }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, d);
  if (lane == 0) red[warp][3 * CT] = (double)nv;
  __syncthreads();
  if (tid < 3 * CT + 1) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < kScalThreads / 32; ++w) a += red[w][tid];
    partial[(size_t)cta * (3 * CT + 1) + tid] = a;
  }
}

// any C <= 32: classes re-read per pass (L1 resident), private float64 columns in shared memory
template <typename LT>
__global__ void __launch_bounds__(kScalThreads)
k_scal_fwd_any(const LT* __restrict__ logits, const uint8_t* __restrict__ target, int F, int C, int64_t S, int ignore,
               double* __restrict__ partial) {
  extern __shared__ double col[];                     // [(3C+1)][threads]
  const int tid = threadIdx.x, nthr = kScalThreads;
  const int nq = 3 * C + 1;
  for (int i = tid; i < nq * nthr; i += nthr) col[i] = 0.0;
  __syncthreads();
  for (int f = blockIdx.y; f < F; f += gridDim.y)
    for (int64_t s = (int64_t)blockIdx.x * nthr + tid; s < S; s += (int64_t)gridDim.x * nthr) {
      const int t = __ldg(target + (size_t)f * S + s);
      if (t == ignore) continue;
      const LT* lp = logits + (size_t)f * C * S + s;
      float m = lt_to_f32<LT>(lp[0]);
      for (int k = 1; k < C; ++k) m = fmaxf(m, lt_to_f32<LT>(lp[(size_t)k * S]));
      float den = 0.f;
      for (int k = 0; k < C; ++k) den += __expf(lt_to_f32<LT>(lp[(size_t)k * S]) - m);
      const float inv = rcp_fast(den);
      col[3 * C * nthr + tid] += 1.0;
      for (int k = 0; k < C; ++k) {
        const float p = __expf(lt_to_f32<LT>(lp[(size_t)k * S]) - m) * inv;
        col[k * nthr + tid] += (double)p;
        if (t == k) { col[(C + k) * nthr + tid] += (double)p; col[(2 * C + k) * nthr + tid] += 1.0; }
      }
    }
  __syncthreads();
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  for (int q = warp; q < nq; q += nthr / 32) {
    double a = 0.0;
    for (int i = lane; i < nthr; i += 32) a += col[q * nthr + i];
#pragma unroll
    for (int d = 16; d; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if (lane == 0) partial[(size_t)cta * nq + q] = a;
  }
}

// F.binary_cross_entropy(x, 1) = -max(log x, -100) and its derivative as torch evaluates it,
// (x - 1) / max((1 - x) x, 1e-12)   (losses.py:230-249, :284-286)
__device__ __forceinline__ double bce1(double x) { return -fmax(log(x), -100.0); }
__device__ __forceinline__ double dbce1(double x) { return (x - 1.0) / fmax((1.0 - x) * x, 1e-12); }
__device__ __forceinline__ bool in01(double x) { return x >= 0.0 && x <= 1.0; }     // false for NaN, like the reference's test

// sums[q] = sum over CTAs of partial[cta][q], in a fixed order; then (one thread) both losses and their derivatives
// with respect to sum_p / nom:   losses[0] = SemScal, losses[1] = GeoScal, losses[2 .. 2+2C) = d SemScal / d(sum_p, nom),
// losses[2+2C .. 2+4C) = d GeoScal / d(sum_p, nom).
__global__ void k_scal_finish(const double* __restrict__ partial, int n_cta, int C, double* __restrict__ sums,
                              double* __restrict__ losses) {
  __shared__ double sm[3 * kScalMaxC + 1];
  const int nq = 3 * C + 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = warp; q < nq; q += blockDim.x >> 5) {        // a warp per quantity: lane-strided, then a fixed tree
    double a = 0.0;
    for (int c = lane; c < n_cta; c += 32) a += partial[(size_t)c * nq + q];
#pragma unroll
    for (int d = 16; d; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if (lane == 0) { sm[q] = a; sums[q] = a; }
  }
  __syncthreads();
  if (threadIdx.x != 0 || !losses) return;
  const double* sp = sm; const double* nm = sm + C; const double* cn = sm + 2 * C;
  const double nv = sm[3 * C];
  double* dsem = losses + 2; double* dgeo = losses + 2 + 2 * C;
  for (int k = 0; k < 4 * C; ++k) dsem[k] = 0.0;
  // SemScalLoss, losses.py:210-251
  double loss = 0.0, count = 0.0;
  for (int i = 0; i < C; ++i) {
    if (!(cn[i] > 0.0)) continue;                                  // :224
    count += 1.0;
    if (sp[i] > 0.0) {                                             // :229
      const double prec = nm[i] / sp[i];
      if (in01(prec)) { loss += bce1(prec); const double d = dbce1(prec); dsem[C + i] += d / sp[i]; dsem[i] -= d * prec / sp[i]; }
    }
    const double rec = nm[i] / cn[i];                              // :236
    if (in01(rec)) { loss += bce1(rec); dsem[C + i] += dbce1(rec) / cn[i]; }
    const double rest = nv - cn[i];                                // sum(1 - completion_target), :243
    if (rest > 0.0) {
      const double spec = (rest - (sp[i] - nm[i])) / rest;         // sum((1-p)(1-ct)) = rest - (sum_p - nom)
      if (in01(spec)) { loss += bce1(spec); const double d = dbce1(spec) / rest; dsem[i] -= d; dsem[C + i] += d; }
    }
  }
  losses[0] = loss / count;                                        // :251 (NaN where the reference divides by zero)
  for (int k = 0; k < 2 * C; ++k) dsem[k] /= count;
  // GeoScalLoss, losses.py:270-287
  const double ne_t = nv - cn[0];                                  // nonempty_target.sum()
  const double ne_p = nv - sp[0];                                  // nonempty_probs.sum()
  const double inter = ne_t - (sp[0] - nm[0]);                     // sum(nonempty_target * (1 - p_0))
  const double P = inter / ne_p, R = inter / ne_t, Sp = nm[0] / cn[0];
  losses[1] = bce1(P) + bce1(R) + bce1(Sp);
  dgeo[0] = dbce1(P) * (inter - ne_p) / (ne_p * ne_p) - dbce1(R) / ne_t;
  dgeo[C] = dbce1(P) / ne_p + dbce1(R) / ne_t + dbce1(Sp) / cn[0];
}

// d loss / d logit_k = p_k (g_k - sum_j p_j g_j) with g_k = gs[k] + [target == k] gs[C + k] on valid voxels, 0 elsewhere
template <typename LT, int CT, int VEC>
__global__ void __launch_bounds__(kScalThreads)
k_scal_bwd(const LT* __restrict__ logits, const uint8_t* __restrict__ target, int F, int64_t S, int ignore,
           const double* __restrict__ dl, const float* __restrict__ g_sem, const float* __restrict__ g_geo, LT* __restrict__ grad) {
  // d loss / d(sum_p, nom) = g_sem * dSemScal + g_geo * dGeoScal (dl = the [2, 2C] block k_scal_finish wrote)
  const double gs = g_sem ? (double)__ldg(g_sem) : 0.0, gg = g_geo ? (double)__ldg(g_geo) : 0.0;
  float ga[CT], gb[CT];
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    ga[k] = (float)(gs * dl[k] + gg * dl[2 * CT + k]);
    gb[k] = (float)(gs * dl[CT + k] + gg * dl[3 * CT + k]);
  }
  const int64_t groups = S / VEC;
  for (int f = blockIdx.y; f < F; f += gridDim.y) {
    const size_t fo = (size_t)f * CT * S;
    const uint8_t* tf = target + (size_t)f * S;
    for (int64_t g = (int64_t)blockIdx.x * kScalThreads + threadIdx.x; g < groups; g += (int64_t)gridDim.x * kScalThreads) {
      const size_t off = fo + (size_t)g * VEC;
      float x[CT][VEC];
#pragma unroll
      for (int k = 0; k < CT; ++k) Vec<LT, VEC>::load(logits + off + (size_t)k * S, x[k]);
      int t[VEC];
      load_target<VEC>(tf + g * VEC, t);
      float o[CT][VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float p[CT];
        softmax_probs<CT, VEC>(x, j, p);
        const bool valid = t[j] != ignore;
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < CT; ++k) dot += p[k] * (ga[k] + (t[j] == k ? gb[k] : 0.f));
#pragma unroll
        for (int k = 0; k < CT; ++k) o[k][j] = valid ? p[k] * ((ga[k] + (t[j] == k ? gb[k] : 0.f)) - dot) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < CT; ++k) Vec<LT, VEC>::store(grad + off + (size_t)k * S, o[k]);
    }
  }
}

template <typename LT>
__global__ void __launch_bounds__(kScalThreads)
k_scal_bwd_any(const LT* __restrict__ logits, const uint8_t* __restrict__ target, int F, int C, int64_t S, int ignore,
               const double* __restrict__ dl, const float* __restrict__ g_sem, const float* __restrict__ g_geo, LT* __restrict__ grad) {
  __shared__ float g_s[2 * kScalMaxC];
  if (threadIdx.x < 2 * C) {
    const double gs = g_sem ? (double)__ldg(g_sem) : 0.0, gg = g_geo ? (double)__ldg(g_geo) : 0.0;
    g_s[threadIdx.x] = (float)(gs * dl[threadIdx.x] + gg * dl[2 * C + threadIdx.x]);
  }
  __syncthreads();
  for (int f = blockIdx.y; f < F; f += gridDim.y)
    for (int64_t s = (int64_t)blockIdx.x * kScalThreads + threadIdx.x; s < S; s += (int64_t)gridDim.x * kScalThreads) {
      const int t = __ldg(target + (size_t)f * S + s);
      const size_t off = (size_t)f * C * S + s;
      if (t == ignore) {
        for (int k = 0; k < C; ++k) grad[off + (size_t)k * S] = f32_to_lt<LT>(0.f);
        continue;
      }
      float m = lt_to_f32<LT>(logits[off]);
      for (int k = 1; k < C; ++k) m = fmaxf(m, lt_to_f32<LT>(logits[off + (size_t)k * S]));
      float den = 0.f;
      for (int k = 0; k < C; ++k) den += __expf(lt_to_f32<LT>(logits[off + (size_t)k * S]) - m);
      const float inv = rcp_fast(den);
      float dot = 0.f;
      for (int k = 0; k < C; ++k) {
        const float p = __expf(lt_to_f32<LT>(logits[off + (size_t)k * S]) - m) * inv;
        dot += p * (g_s[k] + (t == k ? g_s[C + k] : 0.f));
      }
      for (int k = 0; k < C; ++k) {
        const float p = __expf(lt_to_f32<LT>(logits[off + (size_t)k * S]) - m) * inv;
        grad[off + (size_t)k * S] = f32_to_lt<LT>(p * ((g_s[k] + (t == k ? g_s[C + k] : 0.f)) - dot));
      }
    }
}

// grid = (CTAs per frame, frames in flight): exactly one resident wave of the kernel (occupancy API), at most
// kScalMaxCtas partial rows
template <typename K>
static dim3 scal_grid(K kern, size_t smem, int F, int64_t items_per_frame) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kScalThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  int64_t cap = (int64_t)kNumSMsB200 * per_sm;
  if (cap > kScalMaxCtas) cap = kScalMaxCtas;
  int64_t per_frame = ceil_div64(items_per_frame, kScalThreads);
  if (per_frame < 1) per_frame = 1;
  int64_t gy = F < cap ? F : cap;
  int64_t gx = cap / gy;
  if (gx > per_frame) gx = per_frame;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)gy, 1);
}

static bool vec4_ok(const void* a, const void* b, const uint8_t* t, int64_t S, size_t elem) {
  return S % 4 == 0 && reinterpret_cast<uintptr_t>(a) % (4 * elem) == 0 && (!b || reinterpret_cast<uintptr_t>(b) % (4 * elem) == 0) &&
         reinterpret_cast<uintptr_t>(t) % 4 == 0;
}

template <typename LT>
static int launch_scal_fwd(const void* logits, const uint8_t* target, int F, int C, int64_t S, int ignore, double* sums,
                           double* losses, double* partial, cudaStream_t st) {
  const LT* lg = (const LT*)logits;
  const bool v4 = vec4_ok(logits, nullptr, target, S, sizeof(LT));
  dim3 grid;
  prof_mark("<scal>", st);
  if (C == 2 || C == 9) {
    auto kern = C == 2 ? (v4 ? k_scal_fwd<LT, 2, 4, MUVO_SCAL_UNR> : k_scal_fwd<LT, 2, 1, MUVO_SCAL_UNR>) : (v4 ? k_scal_fwd<LT, 9, 4, 1> : k_scal_fwd<LT, 9, 1, 1>);
    grid = scal_grid(kern, 0, F, v4 ? S / 4 : S);
    kern<<<grid, kScalThreads, 0, st>>>(lg, target, F, S, ignore, partial);
  } else {
    const size_t smem = (size_t)(3 * C + 1) * kScalThreads * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(k_scal_fwd_any<LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ::muvo::cuda_fail(e);
    grid = scal_grid(k_scal_fwd_any<LT>, smem, F, S);
    k_scal_fwd_any<LT><<<grid, kScalThreads, smem, st>>>(lg, target, F, C, S, ignore, partial);
  }
  MUVO_AFTER_LAUNCH("k_scal_fwd", st);
  k_scal_finish<<<1, 1024, 0, st>>>(partial, (int)(grid.x * grid.y), C, sums, losses);
  MUVO_AFTER_LAUNCH("k_scal_finish", st);
  return MUVO_OK;
}

template <typename LT>
static int launch_scal_bwd(const void* logits, const uint8_t* target, int F, int C, int64_t S, int ignore, const double* dl,
                           const float* g_sem, const float* g_geo, void* grad, cudaStream_t st) {
  const LT* lg = (const LT*)logits;
  LT* gr = (LT*)grad;
  const bool v4 = vec4_ok(logits, grad, target, S, sizeof(LT));
  prof_mark("<scal>", st);
  if (C == 2 || C == 9) {
    auto kern = C == 2 ? (v4 ? k_scal_bwd<LT, 2, 4> : k_scal_bwd<LT, 2, 1>) : (v4 ? k_scal_bwd<LT, 9, 4> : k_scal_bwd<LT, 9, 1>);
    kern<<<scal_grid(kern, 0, F, v4 ? S / 4 : S), kScalThreads, 0, st>>>(lg, target, F, S, ignore, dl, g_sem, g_geo, gr);
  } else {
    k_scal_bwd_any<LT><<<scal_grid(k_scal_bwd_any<LT>, 0, F, S), kScalThreads, 0, st>>>(lg, target, F, C, S, ignore, dl, g_sem, g_geo, gr);
  }
  MUVO_AFTER_LAUNCH("k_scal_bwd", st);
  return MUVO_OK;
}

}  // namespace
}  // namespace muvo

using namespace muvo;

extern "C" {

int muvo_scal_workspace_bytes(int32_t n_classes, size_t* bytes_out_h) {
  if (!bytes_out_h) return MUVO_E_NULL;
  if (n_classes <= 0 || n_classes > kScalMaxC) return MUVO_E_ARG;
  *bytes_out_h = (size_t)kScalMaxCtas * (3 * n_classes + 1) * sizeof(double);
  return MUVO_OK;
}

int muvo_scal_sums_fwd(const void* logits, int32_t logits_dtype, const uint8_t* target, int32_t n_frames, int32_t n_classes,
                       int64_t voxels_per_frame, int32_t ignore_index, double* sums_out, double* losses_out, void* ws,
                       size_t ws_bytes, void* stream) {
  if (!sums_out) return MUVO_E_NULL;
  if (n_frames < 0 || n_classes <= 0 || n_classes > kScalMaxC || voxels_per_frame < 0) return MUVO_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_frames == 0 || voxels_per_frame == 0) {
    k_scal_finish<<<1, 1024, 0, st>>>(nullptr, 0, n_classes, sums_out, losses_out);     // all-zero sums -> NaN losses
    MUVO_LAUNCH_CHECK();
    return MUVO_OK;
  }
  if (!logits || !target || !ws) return MUVO_E_NULL;
  if (ws_bytes < (size_t)kScalMaxCtas * (3 * n_classes + 1) * sizeof(double)) return MUVO_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) % 8) return MUVO_E_ARG;
  switch (logits_dtype) {
    case MUVO_F32:  return launch_scal_fwd<float>(logits, target, n_frames, n_classes, voxels_per_frame, ignore_index, sums_out, losses_out, (double*)ws, st);
    case MUVO_F16:  return launch_scal_fwd<__half>(logits, target, n_frames, n_classes, voxels_per_frame, ignore_index, sums_out, losses_out, (double*)ws, st);
    case MUVO_BF16: return launch_scal_fwd<__nv_bfloat16>(logits, target, n_frames, n_classes, voxels_per_frame, ignore_index, sums_out, losses_out, (double*)ws, st);
    default: return MUVO_E_ARG;
  }
}

int muvo_scal_sums_bwd(const void* logits, int32_t logits_dtype, const uint8_t* target, int32_t n_frames, int32_t n_classes,
                       int64_t voxels_per_frame, int32_t ignore_index, const double* dloss, const float* g_sem, const float* g_geo,
                       void* grad_logits, void* stream) {
  if (n_frames < 0 || n_classes <= 0 || n_classes > kScalMaxC || voxels_per_frame < 0) return MUVO_E_ARG;
  if (n_frames == 0 || voxels_per_frame == 0) return MUVO_OK;
  if (!logits || !target || !dloss || !grad_logits) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  switch (logits_dtype) {
    case MUVO_F32:  return launch_scal_bwd<float>(logits, target, n_frames, n_classes, voxels_per_frame, ignore_index, dloss, g_sem, g_geo, grad_logits, st);
    case MUVO_F16:  return launch_scal_bwd<__half>(logits, target, n_frames, n_classes, voxels_per_frame, ignore_index, dloss, g_sem, g_geo, grad_logits, st);
    case MUVO_BF16: return launch_scal_bwd<__nv_bfloat16>(logits, target, n_frames, n_classes, voxels_per_frame, ignore_index, dloss, g_sem, g_geo, grad_logits, st);
    default: return MUVO_E_ARG;
  }
}

}  // extern "C"

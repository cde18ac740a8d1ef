// Stage (d): voxel-occupancy IoU / tp-fp-fn counts.
//
// Replaces SSCMetrics.get_score_completion + get_score_semantic_and_completion
// (muvo/metrics.py:143-216): one streaming pass over (pred, target[, masks]) instead of
// (3 + 3C) masked passes per frame with a host sync each.  Integer counts only -> bit-exact and
// order-independent.  HBM bound: 9 B/voxel (int64 pred + uint8 target).
//
// Per-thread private counters live in shared memory as cnt[bin][thread] (bank = thread -> no
// conflicts, no atomics); one u64 atomicAdd per (block, bin) at the end.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace muvo {
namespace {

constexpr int kTileVox = 512;   // voxels per warp tile: 8 x (32 lanes x 2 voxels)

template <typename PT> struct PredVec;   // 2 consecutive predictions per lane per load
template <> struct PredVec<int64_t> {
  static __device__ __forceinline__ void load2(const int64_t* p, int64_t v, int64_t n, long long& a, long long& b) {
    if (v + 1 < n && ((reinterpret_cast<uintptr_t>(p + v) & 15) == 0)) {
      longlong2 t;
      asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(t.x), "=l"(t.y) : "l"(p + v));
      a = t.x; b = t.y;
    } else { a = v < n ? p[v] : 0; b = v + 1 < n ? p[v + 1] : 0; }
  }
};
template <> struct PredVec<int32_t> {
  static __device__ __forceinline__ void load2(const int32_t* p, int64_t v, int64_t n, long long& a, long long& b) {
    a = v < n ? __ldg(p + v) : 0; b = v + 1 < n ? __ldg(p + v + 1) : 0;
  }
};
template <> struct PredVec<int16_t> {
  static __device__ __forceinline__ void load2(const int16_t* p, int64_t v, int64_t n, long long& a, long long& b) {
    a = v < n ? __ldg(p + v) : 0; b = v + 1 < n ? __ldg(p + v + 1) : 0;
  }
};
template <> struct PredVec<uint8_t> {
  static __device__ __forceinline__ void load2(const uint8_t* p, int64_t v, int64_t n, long long& a, long long& b) {
    a = v < n ? __ldg(p + v) : 0; b = v + 1 < n ? __ldg(p + v + 1) : 0;
  }
};

__device__ __forceinline__ void load2_u8(const uint8_t* p, int64_t v, int64_t n, uint32_t& a, uint32_t& b) {
  if (v + 1 < n && ((reinterpret_cast<uintptr_t>(p + v) & 1) == 0)) {
    uint16_t t = __ldg(reinterpret_cast<const uint16_t*>(p + v));
    a = t & 0xffu; b = t >> 8;
  } else { a = v < n ? __ldg(p + v) : 0; b = v + 1 < n ? __ldg(p + v + 1) : 0; }
}

struct Comp { unsigned tp, fp, fn; };

// counters: cnt[(bin) * blockDim.x + tid]; bins: tp[0..C), fp[C..2C), fn[2C..3C)
__device__ __forceinline__ void tally(long long p, uint32_t t, bool sem_valid, bool comp_valid, int C, uint32_t* cnt,
                                      int nthr, int tid, Comp& c) {
  const bool is255 = (t == 255u);
  if (is255) { p = 0; t = 0; }                       // :150-151, :184-185
  const bool bt = t > 0u, bp = p > 0;                // :157-160
  if (comp_valid) { c.tp += (bt && bp); c.fp += (!bt && bp); c.fn += (bt && !bp); }   // :170-172
  if (sem_valid) {                                   // :207-214
    if (p == (long long)t) {
      if ((int)t < C) cnt[(int)t * nthr + tid] += 1u;
    } else {
      if (p >= 0 && p < C) cnt[(C + (int)p) * nthr + tid] += 1u;
      if ((int)t < C) cnt[(2 * C + (int)t) * nthr + tid] += 1u;
    }
  }
}

__device__ __forceinline__ void flush_counts(uint32_t* cnt, int C, Comp c, int64_t* out) {
  const int nthr = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  __syncthreads();
  for (int bin = warp; bin < 3 * C; bin += nwarp) {
    unsigned long long s = 0;
    for (int t = lane; t < nthr; t += 32) s += cnt[bin * nthr + t];
#pragma unroll
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0 && s) atomicAdd(reinterpret_cast<unsigned long long*>(out + 3 + bin), s);
  }
  unsigned long long a = c.tp, b = c.fp, d2 = c.fn;
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); d2 += __shfl_xor_sync(0xffffffffu, d2, d);
  }
  if (lane == 0) {
    if (a) atomicAdd(reinterpret_cast<unsigned long long*>(out + 0), a);
    if (b) atomicAdd(reinterpret_cast<unsigned long long*>(out + 1), b);
    if (d2) atomicAdd(reinterpret_cast<unsigned long long*>(out + 2), d2);
  }
}

template <typename PT>
__global__ void k_ssc_counts(const PT* __restrict__ pred, const uint8_t* __restrict__ target, const uint8_t* __restrict__ nonempty,
                             const uint8_t* __restrict__ nonsurface, int ignore255, int64_t n, int C, int64_t* __restrict__ out) {
  extern __shared__ uint32_t cnt[];
  const int nthr = blockDim.x, tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 3 * C * nthr; i += nthr) cnt[i] = 0;
  __syncthreads();
  Comp c{0, 0, 0};
  const int64_t warp_global = ((int64_t)blockIdx.x * nthr + tid) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * nthr) >> 5;
  const int64_t n_tiles = ceil_div64(n, kTileVox);
  for (int64_t tile = warp_global; tile < n_tiles; tile += n_warps) {
    const int64_t base = tile * kTileVox;
    long long pa[8], pb[8];
    uint32_t ta[8], tb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {   // all loads first: 8 x 16 B (pred) + 8 x 2 B (target) in flight per lane
      int64_t v = base + k * 64 + lane * 2;
      PredVec<PT>::load2(pred, v, n, pa[k], pb[k]);
      load2_u8(target, v, n, ta[k], tb[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int64_t v = base + k * 64 + lane * 2;
      uint32_t ea = 1, eb = 1, sa = 1, sb = 1;
      if (nonempty) load2_u8(nonempty, v, n, ea, eb);
      if (nonsurface) load2_u8(nonsurface, v, n, sa, sb);
      bool va = v < n && ea && !(ignore255 && ta[k] == 255u);
      bool vb = v + 1 < n && eb && !(ignore255 && tb[k] == 255u);
      tally(pa[k], ta[k], va, va && sa, C, cnt, nthr, tid, c);
      tally(pb[k], tb[k], vb, vb && sb, C, cnt, nthr, tid, c);
    }
  }
  flush_counts(cnt, C, c, out);
}

// ---- pair-histogram path (C <= 16): ONE private counter update per voxel.  Each thread owns a column of
// u16 counters in shared memory, bin = tb * (C + 2) + pb with
//   tb in {0..C-1, C = "other target (>= C)"},  pb in {0..C-1, C = "prediction >= C", C+1 = "prediction < 0"},
// plus one dummy bin for voxels that are masked out, for the per-class mask and (when a nonsurface mask is
// given) a second histogram for the completion mask.  tp/fp/fn per class and the completion counts are linear
// in these bins and are derived once per block.
constexpr int kPairTile = 256;   // voxels per warp step: 4 x (32 lanes x 2 voxels)

template <typename PT>
__device__ __forceinline__ int pred_bucket(PT p, int C) {     // class, C = ">= C", C + 1 = "negative"
  long long v = (long long)p;
  return v < 0 ? C + 1 : (v < (long long)C ? (int)v : C);
}

template <typename PT, bool ALIGNED, bool TWO>
__global__ void __launch_bounds__(256, 3)
k_ssc_pairhist(const PT* __restrict__ pred, const uint8_t* __restrict__ target, const uint8_t* __restrict__ nonempty,
               const uint8_t* __restrict__ nonsurface, int ignore255, int64_t n, int C, int64_t* __restrict__ out) {
  extern __shared__ uint16_t hist[];                    // [TWO ? 2 : 1][nbins + 1][256]
  __shared__ unsigned long long red[3 * 16 + 3];
  const int tid = threadIdx.x, lane = tid & 31;
  const int PB = C + 2, nbins = (C + 1) * PB, dummy = nbins;
  const int nh = TWO ? 2 : 1;
  for (int i = tid; i < nh * (nbins + 1) * 256; i += 256) hist[i] = 0;
  for (int i = tid; i < 51; i += 256) red[i] = 0ull;
  __syncthreads();
  uint16_t* mine = hist + tid;
  uint16_t* mine2 = hist + (size_t)(nbins + 1) * 256 + tid;
  const int bin255 = ignore255 ? dummy : 0;             // target == 255: rewritten to (0,0) (:150-151) or ignored (:79)
  const int64_t warp_global = ((int64_t)blockIdx.x * 256 + tid) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * 256) >> 5;
  const int64_t n_tiles = ceil_div64(n, kPairTile);
  for (int64_t tile = warp_global; tile < n_tiles; tile += n_warps) {
    const int64_t base = tile * kPairTile;
    if (ALIGNED && !TWO && nonempty == nullptr && base + kPairTile <= n) {
      // fast path: full tile, no masks -> no bounds checks, no predicates
      PT pv[8];
      uint32_t tv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t v = base + q * 64 + lane * 2;
        if (sizeof(PT) == 8) {
          longlong2 t2;
          asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(t2.x), "=l"(t2.y) : "l"(pred + v));
          pv[2 * q] = (PT)t2.x; pv[2 * q + 1] = (PT)t2.y;
        } else { pv[2 * q] = pred[v]; pv[2 * q + 1] = pred[v + 1]; }
        tv[q] = __ldg(reinterpret_cast<const uint16_t*>(target + v));
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t t = (tv[e >> 1] >> (8 * (e & 1))) & 0xffu;
        const int bin = (t == 255u) ? bin255 : min((int)t, C) * PB + pred_bucket<PT>(pv[e], C);
        mine[bin * 256] += 1;
      }
    } else {
#pragma unroll 1
      for (int e = 0; e < 8; ++e) {
        const int64_t v = base + (e >> 1) * 64 + lane * 2 + (e & 1);
        if (v >= n) continue;
        const uint32_t t = target[v];
        const bool valid = !(nonempty && !nonempty[v]);
        int bin = (t == 255u) ? bin255 : min((int)t, C) * PB + pred_bucket<PT>(pred[v], C);
        if (!valid) bin = dummy;
        mine[bin * 256] += 1;
        if (TWO) mine2[((valid && nonsurface[v]) ? bin : dummy) * 256] += 1;
      }
    }
  }
  __syncthreads();
  // block reduction: warp w sums bins w, w+8, ...; contributions go to red[] (tp | fp | fn | completion)
  const int warp = tid >> 5;
  for (int hsel = 0; hsel < nh; ++hsel) {
    const uint16_t* hh = hist + (size_t)hsel * (nbins + 1) * 256;
    for (int bin = warp; bin < nbins; bin += 8) {
      unsigned sum = 0;
      for (int t = lane; t < 256; t += 32) sum += hh[bin * 256 + t];
      sum = __reduce_add_sync(0xffffffffu, sum);
      if (lane == 0 && sum) {
        const int tb = bin / PB, pb = bin % PB;
        const unsigned long long v = sum;
        if (hsel == 0) {                                                // per-class counts (:207-214)
          if (tb == pb && tb < C) atomicAdd(&red[tb], v);
          else {
            if (pb < C) atomicAdd(&red[16 + pb], v);
            if (tb < C) atomicAdd(&red[32 + tb], v);
          }
        }
        if (hsel == nh - 1) {                                           // completion counts (:157-172)
          const bool bt = tb > 0, bp = (pb > 0 && pb <= C);
          if (bt && bp) atomicAdd(&red[48], v);
          else if (!bt && bp) atomicAdd(&red[49], v);
          else if (bt && !bp) atomicAdd(&red[50], v);
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 51; i += 256) {
    unsigned long long v = red[i];
    if (!v) continue;
    if (i >= 48) atomicAdd(reinterpret_cast<unsigned long long*>(out + (i - 48)), v);
    else {
      int fam = i >> 4, j = i & 15;
      if (j < C) atomicAdd(reinterpret_cast<unsigned long long*>(out + 3 + fam * C + j), v);
    }
  }
}

// ---- warp-vote path (C <= 32): the warp's 32 voxels of a step are classified together -- per class j,
// ballot(t == j) and ballot(p == j) give tp/fp/fn of that class with three popcounts, and lane j keeps the
// counters of class j.  Cost ~ (13 C + 20) warp instructions per 32 voxels instead of ~60 per voxel.
struct VoteAcc { unsigned tp, fp, fn; unsigned ctp, cfp, cfn; };

__device__ __forceinline__ void vote_step(long long p, uint32_t t, bool sem_valid, bool comp_valid, int C, unsigned lane,
                                          VoteAcc& a) {
  const bool is255 = (t == 255u);
  if (is255) { p = 0; t = 0; }                                   // :150-151, :184-185
  const int pc = (p >= 0 && p < (long long)C) ? (int)p : -1;      // class of the prediction, -1 = none
  const int tc = ((int)t < C) ? (int)t : -2;                      // class of the target, -2 = none
  const unsigned vs = __ballot_sync(0xffffffffu, sem_valid);
  const unsigned vc = __ballot_sync(0xffffffffu, comp_valid);
  const unsigned bt = __ballot_sync(0xffffffffu, t > 0u) & vc;    // :157-160
  const unsigned bp = __ballot_sync(0xffffffffu, p > 0) & vc;
  a.ctp += __popc(bt & bp); a.cfp += __popc(~bt & bp); a.cfn += __popc(bt & ~bp);   // :170-172 (lane-uniform)
  for (int j = 0; j < C; ++j) {                                   // :207-214
    const unsigned mt = __ballot_sync(0xffffffffu, tc == j) & vs;
    const unsigned mp = __ballot_sync(0xffffffffu, pc == j) & vs;
    if (lane == (unsigned)j) { a.tp += __popc(mt & mp); a.fp += __popc(~mt & mp & vs); a.fn += __popc(mt & ~mp); }
  }
}

template <typename PT>
__global__ void __launch_bounds__(256)
k_ssc_counts_vote(const PT* __restrict__ pred, const uint8_t* __restrict__ target, const uint8_t* __restrict__ nonempty,
                  const uint8_t* __restrict__ nonsurface, int ignore255, int64_t n, int C, int64_t* __restrict__ out) {
  __shared__ unsigned long long red[3 * 32 + 3];
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31;
  for (int i = tid; i < 3 * 32 + 3; i += blockDim.x) red[i] = 0ull;
  __syncthreads();
  VoteAcc a{0, 0, 0, 0, 0, 0};
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + tid) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_tiles = ceil_div64(n, kTileVox);
  for (int64_t tile = warp_global; tile < n_tiles; tile += n_warps) {
    const int64_t base = tile * kTileVox;
    long long pa[8], pb[8];
    uint32_t ta[8], tb[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {   // all loads first: 8 x 16 B (pred) + 8 x 2 B (target) in flight per lane
      int64_t v = base + q * 64 + lane * 2;
      PredVec<PT>::load2(pred, v, n, pa[q], pb[q]);
      load2_u8(target, v, n, ta[q], tb[q]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int64_t v = base + q * 64 + lane * 2;
      uint32_t ea = 1, eb = 1, sa = 1, sb = 1;
      if (nonempty) load2_u8(nonempty, v, n, ea, eb);
      if (nonsurface) load2_u8(nonsurface, v, n, sa, sb);
      bool va = v < n && ea && !(ignore255 && ta[q] == 255u);
      bool vb = v + 1 < n && eb && !(ignore255 && tb[q] == 255u);
      vote_step(pa[q], ta[q], va, va && sa, C, lane, a);
      vote_step(pb[q], tb[q], vb, vb && sb, C, lane, a);
    }
  }
  // lane j holds class j; completion counters are lane-uniform.  Block-level integer reduction, then one global
  // atomic per (block, bin).
  if ((int)lane < C) {
    if (a.tp) atomicAdd(&red[lane], (unsigned long long)a.tp);
    if (a.fp) atomicAdd(&red[32 + lane], (unsigned long long)a.fp);
    if (a.fn) atomicAdd(&red[64 + lane], (unsigned long long)a.fn);
  }
  if (lane == 0) {
    if (a.ctp) atomicAdd(&red[96], (unsigned long long)a.ctp);
    if (a.cfp) atomicAdd(&red[97], (unsigned long long)a.cfp);
    if (a.cfn) atomicAdd(&red[98], (unsigned long long)a.cfn);
  }
  __syncthreads();
  for (int i = tid; i < 99; i += blockDim.x) {
    unsigned long long v = red[i];
    if (!v) continue;
    if (i >= 96) atomicAdd(reinterpret_cast<unsigned long long*>(out + (i - 96)), v);
    else {
      int fam = i >> 5, j = i & 31;
      if (j < C) atomicAdd(reinterpret_cast<unsigned long long*>(out + 3 + fam * C + j), v);
    }
  }
}

// Fused argmax + counts: logits [F, C, S]; lane-coalesced along S for every class plane.
template <typename LT> __device__ __forceinline__ float to_f32(LT v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename LT>
__global__ void k_ssc_from_logits(const LT* __restrict__ logits, const uint8_t* __restrict__ target, int F, int C, int64_t S,
                                  int ignore255, int64_t* __restrict__ out) {
  extern __shared__ uint32_t cnt[];
  const int nthr = blockDim.x, tid = threadIdx.x;
  for (int i = tid; i < 3 * C * nthr; i += nthr) cnt[i] = 0;
  __syncthreads();
  Comp c{0, 0, 0};
  const int64_t total = (int64_t)F * S;
  for (int64_t v = (int64_t)blockIdx.x * nthr + tid; v < total; v += (int64_t)gridDim.x * nthr) {
    const int64_t f = v / S, s = v - f * S;
    const LT* lp = logits + (size_t)f * C * S + s;
    float best = to_f32<LT>(lp[0]);
    int arg = 0;
    for (int k = 1; k < C; ++k) {      // torch.argmax: first maximum; NaN counts as maximal
      float x = to_f32<LT>(lp[(size_t)k * S]);
      if (x > best || (x != x && best == best)) { best = x; arg = k; }
    }
    uint32_t t = __ldg(target + v);
    bool valid = !(ignore255 && t == 255u);
    tally((long long)arg, t, valid, valid, C, cnt, nthr, tid, c);
  }
  flush_counts(cnt, C, c, out);
}

static int pick_threads(int C, size_t* smem) {
  for (int thr = 256; thr >= 64; thr >>= 1) {
    size_t b = (size_t)3 * C * thr * 4;
    if (b <= 200 * 1024) { *smem = b; return thr; }
  }
  return 0;
}

static int blocks_per_sm(size_t smem, int thr) {
  int by_smem = (int)((220 * 1024) / (smem + 1024));
  int by_thr = 2048 / thr;
  int v = by_smem < by_thr ? by_smem : by_thr;
  return v < 1 ? 1 : v;
}

template <typename K>
static int set_smem(K kern, size_t smem) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  }
  return 0;
}

template <typename PT>
static int launch_counts(const void* pred, const uint8_t* target, const uint8_t* ne, const uint8_t* ns, int ignore255,
                         int64_t n, int C, int64_t* out, cudaStream_t st) {
  if (C <= 16) {
    // u16 private counters: at most 2^31 voxels per launch (<= 7 k voxels per thread)
    const int nbins = (C + 1) * (C + 2);
    const bool two = ns != nullptr;
    const size_t smem = (size_t)(two ? 2 : 1) * (nbins + 1) * 256 * sizeof(uint16_t);
    if (smem <= 200 * 1024) {
      const bool aligned = (reinterpret_cast<uintptr_t>(pred) % 16 == 0) && (reinterpret_cast<uintptr_t>(target) % 2 == 0);
      auto kern = two ? (aligned ? k_ssc_pairhist<PT, true, true> : k_ssc_pairhist<PT, false, true>)
                      : (aligned ? k_ssc_pairhist<PT, true, false> : k_ssc_pairhist<PT, false, false>);
      int rc = set_smem(kern, smem);
      if (rc) return rc;
      const int per_sm = blocks_per_sm(smem, 256);
      const int64_t chunk = (int64_t)1 << 31;
      for (int64_t o = 0; o < n; o += chunk) {
        const int64_t m = (n - o) < chunk ? (n - o) : chunk;
        int64_t want = ceil_div64(ceil_div64(m, kPairTile), 8);
        int64_t cap = (int64_t)kNumSMsB200 * per_sm;
        unsigned grid = (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
        prof_mark("<ssc>", st);
        kern<<<grid, 256, smem, st>>>((const PT*)pred + o, target + o, ne ? ne + o : nullptr, ns ? ns + o : nullptr, ignore255, m,
                                      C, out);
        MUVO_AFTER_LAUNCH("k_ssc_pairhist", st);
      }
      return MUVO_OK;
    }
  }
  if (C <= 32 && n < ((int64_t)1 << 40)) {                      // u32 per-warp counters cannot overflow below 2^40 voxels
    const int64_t tiles = ceil_div64(n, kTileVox);
    int64_t want = ceil_div64(tiles, 8);                       // 8 warps per block, >= 1 tile per warp
    int64_t cap = (int64_t)kNumSMsB200 * 8;
    unsigned grid = (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
    prof_mark("<ssc>", st);
    k_ssc_counts_vote<PT><<<grid, 256, 0, st>>>((const PT*)pred, target, ne, ns, ignore255, n, C, out);
    MUVO_AFTER_LAUNCH("k_ssc_counts_vote", st);
    return MUVO_OK;
  }
  size_t smem; int thr = pick_threads(C, &smem);
  if (!thr) return MUVO_E_ARG;
  int rc = set_smem(k_ssc_counts<PT>, smem);
  if (rc) return rc;
  int per_sm = blocks_per_sm(smem, thr);
  int64_t want = ceil_div64(ceil_div64(n, kTileVox) * 32, thr);
  int64_t cap = (int64_t)kNumSMsB200 * per_sm;
  unsigned grid = (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
  prof_mark("<ssc>", st);
  k_ssc_counts<PT><<<grid, thr, smem, st>>>((const PT*)pred, target, ne, ns, ignore255, n, C, out);
  MUVO_AFTER_LAUNCH("k_ssc_counts", st);
  return MUVO_OK;
}

template <typename LT>
static int launch_logits(const void* logits, const uint8_t* target, int F, int C, int64_t S, int ignore255, int64_t* out,
                         cudaStream_t st) {
  size_t smem; int thr = pick_threads(C, &smem);
  if (!thr) return MUVO_E_ARG;
  int rc = set_smem(k_ssc_from_logits<LT>, smem);
  if (rc) return rc;
  int per_sm = blocks_per_sm(smem, thr);
  int64_t want = ceil_div64((int64_t)F * S, thr);
  int64_t cap = (int64_t)kNumSMsB200 * per_sm;
  unsigned grid = (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
  prof_mark("<ssc>", st);
  k_ssc_from_logits<LT><<<grid, thr, smem, st>>>((const LT*)logits, target, F, C, S, ignore255, out);
  MUVO_AFTER_LAUNCH("k_ssc_from_logits", st);
  return MUVO_OK;
}

}  // namespace
}  // namespace muvo

using namespace muvo;

extern "C" {

int muvo_ssc_counts(const void* pred, int32_t pred_dtype, const uint8_t* target, const uint8_t* nonempty,
                    const uint8_t* nonsurface, int32_t ignore255, int64_t n_voxels, int32_t n_classes,
                    int64_t* counts_out, void* stream) {
  if (!counts_out) return MUVO_E_NULL;
  if (n_voxels < 0 || n_classes <= 0) return MUVO_E_ARG;
  if (n_voxels == 0) return MUVO_OK;
  if (!pred || !target) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  switch (pred_dtype) {
    case MUVO_I64: return launch_counts<int64_t>(pred, target, nonempty, nonsurface, ignore255, n_voxels, n_classes, counts_out, st);
    case MUVO_I32: return launch_counts<int32_t>(pred, target, nonempty, nonsurface, ignore255, n_voxels, n_classes, counts_out, st);
    case MUVO_I16: return launch_counts<int16_t>(pred, target, nonempty, nonsurface, ignore255, n_voxels, n_classes, counts_out, st);
    case MUVO_U8:  return launch_counts<uint8_t>(pred, target, nonempty, nonsurface, ignore255, n_voxels, n_classes, counts_out, st);
    default: return MUVO_E_ARG;
  }
}

int muvo_ssc_counts_from_logits(const void* logits, int32_t logits_dtype, const uint8_t* target, int32_t n_frames,
                                int32_t n_classes, int64_t voxels_per_frame, int32_t ignore255, int64_t* counts_out,
                                void* stream) {
  if (!counts_out) return MUVO_E_NULL;
  if (n_frames < 0 || n_classes <= 0 || voxels_per_frame < 0) return MUVO_E_ARG;
  if (n_frames == 0 || voxels_per_frame == 0) return MUVO_OK;
  if (!logits || !target) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  switch (logits_dtype) {
    case MUVO_F32:  return launch_logits<float>(logits, target, n_frames, n_classes, voxels_per_frame, ignore255, counts_out, st);
    case MUVO_F16:  return launch_logits<__half>(logits, target, n_frames, n_classes, voxels_per_frame, ignore255, counts_out, st);
    case MUVO_BF16: return launch_logits<__nv_bfloat16>(logits, target, n_frames, n_classes, voxels_per_frame, ignore255, counts_out, st);
    default: return MUVO_E_ARG;
  }
}

}  // extern "C"

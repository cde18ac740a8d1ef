// Stage (c), streamed forward: out[b, c, cell] = sum of x[b, p, c] over the kept frustum points p of the cell
// (FrustumPooling.voxel_pooling, muvo/models/frustum_pooling.py:131-187, for the lifted tensor of mile.py:517-521 whose
// memory is (B, C, D, H, W): every channel row is contiguous in the point index).
//
// The gather kernels of bev.cu read only the kept elements, but a top-k depth mask leaves ~92 % of the 32-byte sectors of
// the tensor touched, so they move the whole tensor anyway, one scattered sector at a time (3.6x the algorithmic bytes,
// 36 % of the HBM rate).  Here the tensor is STREAMED: a producer warp moves [8 channels x 2048 points] tiles into a
// shared-memory ring with cp.async.bulk (TMA) + mbarriers, and the segment sums are formed from shared memory:
//
//   L  k_chunk_lists : per (frame, chunk of 2048 points): keys (cell << 11 | position) of the kept points, mask folded in,
//                      bitonic-sorted in shared memory (= stable by cell, ascending point order inside a cell) and written
//                      row-interleaved per warp portion (write_rows).  k_chunk_compact: the same from a cached,
//                      mask-independent sorted plan by a stable compaction (no sort per call).
//   P  k_pool_stream : one CTA per (frame, 8 channels), walking the chunks in order; a chunk's list is cut into 8 warp
//                      portions at cell boundaries, every lane owns a contiguous run of its warp's portion and sums it for
//                      all 8 channels at once (accumulators: 8 x n_cells floats in shared memory, no atomics); the segment a
//                      run shares with the previous lanes (its first cell) is combined by a fixed shuffle tree; one named
//                      barrier per chunk.  The summation order is a function of (geometry, mask) only: deterministic, and
//                      within the fp32 tolerance of the reference's cumsum differencing.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "bev_stream.cuh"
#include "tma.cuh"

namespace muvo {
namespace {

constexpr int kSCh = 8;                            // channels per CTA = consumer warps
constexpr int kSThreads = (kSCh + 1) * 32;         // + one producer warp
constexpr uint32_t kNone = 0xffffffffu;
constexpr uint32_t kPosMask = (1u << kStreamPosBits) - 1u;
constexpr int kListThreads = kStreamChunk / 2;
constexpr int kSVirt = kSCh * 32;                  // consumer threads = virtual lanes of a chunk list
__host__ __device__ constexpr size_t align_up16(size_t v) { return (v + 15) & ~(size_t)15; }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------- L: per-chunk sorted lists
// Layout of one chunk's list (kListStride u32 = 74 rows of 32): row 0 is a header, word w < 8 = (first row << 16 | rows) of
// consumer warp w, word 8 = total rows; the data rows follow.  The chunk's sorted keys [0, n) are cut into 8 warp portions
// AT CELL BOUNDARIES (the first cell start at or after w * n / 8), so no cell is shared between two warps of the pool kernel
// (no cross-warp ordering, no barrier); inside a portion of length L, lane l owns the contiguous run [l*m, (l+1)*m),
// m = ceil(L / 32), stored row-interleaved: row i holds entry l*m + i of every lane.  Rows <= n/32 + 8 <= 72.
// steps[b][k] = 1 + data rows = rows the pool kernel's producer copies.
constexpr int kListRows = kStreamChunk / 32 + kSCh + 2;          // header + data rows (+1 spare): 74
constexpr int kListStride = kListRows * 32;                      // u32 per chunk

// `keys` = the chunk's n sorted keys in shared memory (visible to the block); called by every thread of a block of >= 256 threads
__device__ __forceinline__ void write_rows(const uint32_t* keys, int n, uint32_t* __restrict__ out, uint32_t* __restrict__ steps_slot,
                                           uint32_t* hdr_s /* [kSCh + 1] shared scratch */) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const unsigned lane = lane_id();
  if (warp < kSCh) {                                             // warp w: portion start b_w = first cell start >= w * n / 8
    int bnd = (int)(((int64_t)warp * n) / kSCh);
    if (warp > 0 && bnd > 0) {
      for (;;) {                                                 // 32 candidates per round
        const int p = bnd + (int)lane;
        const bool hit = p >= n || (keys[p] >> kStreamPosBits) != (keys[p - 1] >> kStreamPosBits);
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal) { bnd += __ffs(bal) - 1; break; }
        bnd += 32;
      }
      if (bnd > n) bnd = n;
    }
    if (lane == 0) hdr_s[warp] = (uint32_t)bnd;
  }
  if (tid == 0) hdr_s[kSCh] = (uint32_t)n;
  __syncthreads();
  uint32_t row = 1, my_hdr = 0;                                  // every thread derives the 8 (first row, rows) pairs itself
  for (int w = 0; w < kSCh; ++w) {
    const uint32_t b0 = hdr_s[w], b1 = hdr_s[w + 1], m = (b1 - b0 + 31u) >> 5;
    if (tid == w) my_hdr = (row << 16) | m;
    for (int idx = tid; idx < (int)m * 32; idx += blockDim.x) {
      const uint32_t src = b0 + (uint32_t)(idx & 31) * m + (uint32_t)(idx >> 5);
      out[(row + (idx >> 5)) * 32 + (idx & 31)] = src < b1 ? keys[src] : kNone;
    }
    row += m;
  }
  if (tid < 32) out[tid] = tid < kSCh ? my_hdr : (tid == kSCh ? row : 0u);
  if (tid == 0) *steps_slot = row;                               // header row + data rows
}

__global__ void __launch_bounds__(kListThreads)
k_chunk_lists(const int32_t* __restrict__ cell0, const uint8_t* __restrict__ mask, int32_t* __restrict__ cell_out, int64_t n_pts,
              int n_cells, int n_chunks, uint32_t* __restrict__ lists, uint32_t* __restrict__ steps, int plan_mode) {
  __shared__ uint32_t s[kStreamChunk];
  __shared__ uint32_t hdr_s[kSCh + 1];
  __shared__ int n_s;
  const int b = blockIdx.y, k = blockIdx.x, tid = threadIdx.x;
  const int64_t base = (int64_t)k * kStreamChunk;
  if (tid == 0) n_s = 0;
#pragma unroll
  for (int r = 0; r < kStreamChunk / kListThreads; ++r) {
    const int i = tid + r * kListThreads;
    const int64_t p = base + i;
    uint32_t key = kNone;
    if (p < n_pts) {
      int32_t c = __ldg(cell0 + (int64_t)b * n_pts + p);
      if (mask && !__ldg(mask + (int64_t)b * n_pts + p)) c = -1;
      if (cell_out) cell_out[(int64_t)b * n_pts + p] = c;
      if (c >= 0 && c < n_cells) key = ((uint32_t)c << kStreamPosBits) | (uint32_t)i;
    }
    s[i] = key;
  }
  __syncthreads();
  // bitonic sort, ascending (dropped points = 0xffffffff end up last)
  for (int kk = 2; kk <= kStreamChunk; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      const int t = tid;
      const int i = 2 * t - (t & (j - 1));             // lower element of the pair (bit j clear)
      const uint32_t a = s[i], c = s[i + j];
      const bool up = (i & kk) == 0;
      if ((a > c) == up) { s[i] = c; s[i + j] = a; }
      __syncthreads();
    }
  }
#pragma unroll
  for (int r = 0; r < kStreamChunk / kListThreads; ++r) {
    const int i = tid + r * kListThreads;
    if (s[i] != kNone && (i == kStreamChunk - 1 || s[i + 1] == kNone)) n_s = i + 1;
  }
  __syncthreads();
  const int n = n_s;
  if (plan_mode) {                // the mask-independent plan: the sorted keys as they are + their count (k_chunk_compact's input)
    uint32_t* out = lists + ((size_t)b * n_chunks + k) * kStreamChunk;
    for (int idx = tid; idx < kStreamChunk; idx += kListThreads) out[idx] = s[idx];
    if (tid == 0) steps[(size_t)b * n_chunks + k] = (uint32_t)n;
    return;
  }
  write_rows(s, n, lists + ((size_t)b * n_chunks + k) * kListStride, steps + (size_t)b * n_chunks + k, hdr_s);
}

// Per call, with a cached plan: keep the plan entries whose mask byte is set (a stable compaction: the order by cell, then
// by point, survives), write them lane-interleaved like k_chunk_lists, and write the folded cell ids the backward needs.
constexpr int kCompactThreads = 256;
constexpr int kCompactPer = kStreamChunk / kCompactThreads;
__global__ void __launch_bounds__(kCompactThreads)
k_chunk_compact(const uint32_t* __restrict__ plan, const uint32_t* __restrict__ plan_n, const int32_t* __restrict__ cell0,
                const uint8_t* __restrict__ mask, int32_t* __restrict__ cell_out, int64_t n_pts, int n_chunks,
                uint32_t* __restrict__ lists, uint32_t* __restrict__ steps) {
  __shared__ __align__(16) uint8_t ms[kStreamChunk];
  __shared__ uint32_t cs[kStreamChunk];
  __shared__ uint32_t wsum[kCompactThreads / 32];
  __shared__ uint32_t hdr_s[kSCh + 1];
  const int b = blockIdx.y, k = blockIdx.x, tid = threadIdx.x;
  const unsigned lane = lane_id();
  const int64_t base = (int64_t)b * n_pts + (int64_t)k * kStreamChunk;
  const int64_t left = n_pts - (int64_t)k * kStreamChunk;
  const int npt = left < kStreamChunk ? (int)left : kStreamChunk;
  for (int i = tid; i < npt; i += kCompactThreads) {
    const uint8_t mk = mask ? __ldg(mask + base + i) : (uint8_t)1;
    ms[i] = mk;
    if (cell_out) cell_out[base + i] = mk ? __ldg(cell0 + base + i) : -1;
  }
  __syncthreads();
  const size_t slot = (size_t)b * n_chunks + k;
  const int n0 = (int)__ldg(plan_n + slot);
  const uint32_t* pl = plan + slot * kStreamChunk;
  uint32_t e[kCompactPer];
  uint32_t keep = 0;
#pragma unroll
  for (int u = 0; u < kCompactPer; ++u) {
    const int idx = tid * kCompactPer + u;
    e[u] = idx < n0 ? __ldg(pl + idx) : kNone;
    if (e[u] != kNone && ms[e[u] & kPosMask]) keep |= 1u << u;
  }
  const uint32_t cnt = __popc(keep);
  uint32_t incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (unsigned)d) incl += t; }
  if (lane == 31) wsum[tid >> 5] = incl;
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kCompactThreads / 32; ++w) { const uint32_t t = wsum[w]; if (w < (tid >> 5)) wbase += t; total += t; }
  uint32_t o = wbase + incl - cnt;
#pragma unroll
  for (int u = 0; u < kCompactPer; ++u) if (keep & (1u << u)) cs[o++] = e[u];
  __syncthreads();
  write_rows(cs, (int)total, lists + slot * kListStride, steps + slot, hdr_s);
}

// ---------------------------------------------------------------- P: streamed pool
template <typename T, int NS>
__global__ void __launch_bounds__(kSThreads, 1)
k_pool_stream(const T* __restrict__ x, int64_t sb, int64_t sc, const uint32_t* __restrict__ lists, const uint32_t* __restrict__ steps,
              int B, int64_t n_pts, int C, int n_cells, int n_chunks, float* __restrict__ out, int flags) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr size_t kChanBytes = (size_t)kStreamChunk * sizeof(T);
  constexpr size_t kTileBytes = kChanBytes * kSCh;
  constexpr size_t kListBytes = (size_t)kListStride * 4;
  constexpr size_t kStageBytes = kTileBytes + kListBytes;
  // [NS stages: 8 channel tiles + the chunk's list] [acc: 8 x n_cells floats] [barriers]
  float* acc = reinterpret_cast<float*>(smem + NS * kStageBytes);
  const size_t acc_bytes = align_up16((size_t)kSCh * n_cells * 4);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NS * kStageBytes + acc_bytes);
  uint64_t* empty = full + NS;
  const int tid = threadIdx.x, warp = tid >> 5;
  const unsigned lane = lane_id();
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kSCh); }
    mbar_fence_init();
  }
  __syncthreads();
  const int n_cg = (C + kSCh - 1) / kSCh;
  const int items = B * n_cg;
  if (warp == kSCh) {
    // ---- producer: one elected lane issues every bulk copy of this CTA
    if (lane == 0) {
      const uint64_t pol = l2_evict_first_policy(), pol_keep = l2_default_policy();
      uint32_t it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = item / n_cg, c0 = (item % n_cg) * kSCh;
        const int nc = C - c0 < kSCh ? C - c0 : kSCh;
        const T* xb = x + (int64_t)b * sb + (int64_t)c0 * sc;
        const uint32_t* mb = steps + (size_t)b * n_chunks;
        uint32_t rows_next = __ldg(mb);
        for (int k = 0; k < n_chunks; ++k, ++it) {
          const int s = it % NS;
          const uint32_t use = it / NS;
          const uint32_t rows = rows_next;                               // header row + data rows of this chunk's list
          if (k + 1 < n_chunks) rows_next = __ldg(mb + k + 1);
          if (use > 0) mbar_wait(empty + s, (use - 1) & 1u);
          unsigned char* sbase = smem + s * kStageBytes;
          const int64_t p0 = (int64_t)k * kStreamChunk;
          const uint32_t bytes = (uint32_t)((n_pts - p0 < kStreamChunk ? n_pts - p0 : kStreamChunk) * sizeof(T));
          mbar_expect_tx(full + s, bytes * nc + rows * 128u);
          bulk_g2s(sbase + kTileBytes, lists + ((size_t)b * n_chunks + k) * kListStride, rows * 128u, full + s, pol_keep);
          for (int c = 0; c < nc; ++c) bulk_g2s(sbase + c * kChanBytes, xb + (int64_t)c * sc + p0, bytes, full + s, pol);
        }
      }
    }
    return;
  }
  // ---- consumers.  Warp w walks ITS portion of every chunk list (portions start at cell boundaries: the cells of two warps
  // are disjoint within a chunk, so the warps never synchronise with each other); lane l owns a contiguous run of the
  // portion and sums it for all 8 channels at once.  A run's segments that START inside the run are added to the
  // accumulators directly (nobody else holds points of those cells in this chunk); the run's first segment may continue the
  // previous lane's last one: those go through a segmented shuffle scan over the lanes (fixed tree) and are added once.
  uint32_t it = 0;
  for (int i = tid; i < kSCh * n_cells; i += kSVirt) acc[i] = 0.f;
  asm volatile("bar.sync 1, %0;" ::"n"(kSVirt) : "memory");
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / n_cg, c0 = (item % n_cg) * kSCh;
    const int nc = C - c0 < kSCh ? C - c0 : kSCh;
    for (int k = 0; k < n_chunks; ++k, ++it) {
      const int s = it % NS;
      mbar_wait(full + s, (it / NS) & 1u);
      const unsigned char* sbase = smem + s * kStageBytes;
      const T* st = reinterpret_cast<const T*>(sbase);
      const uint32_t* L = reinterpret_cast<const uint32_t*>(sbase + kTileBytes);
      const uint32_t h = L[warp];
      const int m = (flags & 1) ? 0 : (int)(h & 0xffffu);
      const uint32_t* Lw = L + (h >> 16) * 32 + lane;
      uint32_t cur = kNone, fcell = kNone;
      float run[kSCh], fp[kSCh];
#pragma unroll
      for (int c = 0; c < kSCh; ++c) { run[c] = 0.f; fp[c] = 0.f; }
      bool first = true;
      auto close = [&]() {
        if (cur != kNone) {
          if (first) {
#pragma unroll
            for (int c = 0; c < kSCh; ++c) fp[c] = run[c];
            fcell = cur; first = false;
          } else {
            float a[kSCh];
#pragma unroll
            for (int c = 0; c < kSCh; ++c) a[c] = acc[c * n_cells + cur];
#pragma unroll
            for (int c = 0; c < kSCh; ++c) acc[c * n_cells + cur] = a[c] + run[c];
          }
        }
      };
      for (int i = 0; i < m; ++i) {
        const uint32_t e = Lw[i * 32];
        if (e != kNone) {
          const uint32_t cell = e >> kStreamPosBits, pos = e & kPosMask;
          float v[kSCh];
#pragma unroll
          for (int c = 0; c < kSCh; ++c) v[c] = to_f32<T>(st[c * kStreamChunk + pos]);
          if (cell != cur) {
            close();
            cur = cell;
#pragma unroll
            for (int c = 0; c < kSCh; ++c) run[c] = 0.f;
          }
#pragma unroll
          for (int c = 0; c < kSCh; ++c) run[c] += v[c];
        }
      }
      close();
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);                             // the tile is consumed: the producer may refill the stage
      if (m > 0) {
        // first cells of the lanes: a non-decreasing sequence over the lanes that hold entries (a prefix); inclusive
        // segmented scan, stopped as soon as no lane has an equal first cell 2^r lanes below
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t cu = __shfl_up_sync(0xffffffffu, fcell, off);
          const bool add = lane >= (unsigned)off && cu == fcell && fcell != kNone;
          if (!__any_sync(0xffffffffu, add)) break;
#pragma unroll
          for (int c = 0; c < kSCh; ++c) {
            const float t = __shfl_up_sync(0xffffffffu, fp[c], off);
            if (add) fp[c] += t;
          }
        }
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, fcell, 1);
        if (fcell != kNone && (lane == 31u || nxt != fcell)) {
          float a[kSCh];
#pragma unroll
          for (int c = 0; c < kSCh; ++c) a[c] = acc[c * n_cells + fcell];
#pragma unroll
          for (int c = 0; c < kSCh; ++c) acc[c * n_cells + fcell] = a[c] + fp[c];
        }
      }
      // the portions are re-cut for every chunk, so a cell may change hands between chunks: every warp's additions of this
      // chunk must have landed before anyone starts the next one (the only cross-warp synchronisation of the loop)
      asm volatile("bar.sync 1, %0;" ::"n"(kSVirt) : "memory");
    }
    // the item's 8 x n_cells sums -> out (contiguous: channels c0 .. c0 + nc - 1 of frame b), accumulators back to zero
    float* o = out + ((size_t)b * C + c0) * n_cells;
    const int n_out = nc * n_cells;
    for (int i = tid; i < kSCh * n_cells; i += kSVirt) {
      if (i < n_out) o[i] = acc[i];
      acc[i] = 0.f;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kSVirt) : "memory");
  }
}

// ---------------------------------------------------------------- B: streamed backward
// grad_x[b, p, c] = kept(p) ? grad_out[b, c, cell(p)] : 0 for (B, C, D, H, W) gradient memory: every element is written
// once, 79 % of them zeros.  The gather kernel of bev.cu spends 176 M warp instructions on address arithmetic and predicated
// loads for 1.42 GB of stores (ncu round 1: 288 us, issue-active 55 %).  Here one CTA per (frame, 8 channels) keeps its 8
// grad_out rows in shared memory, builds [8 channels x 2048 points] tiles there -- zero fill with 128-bit stores, then the
// kept points scattered from the forward pass's chunk lists -- and hands every tile to the TMA (cp.async.bulk shared ->
// global, 8 KiB per channel row); two tiles alternate, a tile is reused once its bulk group has finished reading it.
// 16-bit gradients take 16 channels per CTA (the same 64 KiB tiles as float32 with 8): the per-chunk costs -- zero fill, list wait,
// three barriers -- are per tile, so halving the tiles per byte took the fp16 backward from 0.21 to the float32 rate.  The
// grad_out rows are kept already converted to T (one conversion per (channel, cell) instead of one per point).
constexpr int kBThreads = 512;     // (256: 0.250 ms at cfg3, 512: 0.243, 1024: 0.241)
template <typename T> constexpr int bwd_channels() { return sizeof(T) == 4 ? kSCh : 2 * kSCh; }
template <typename T, int CH>
__global__ void __launch_bounds__(kBThreads, 1)
k_pool_bwd_stream(const float* __restrict__ gout, const uint32_t* __restrict__ lists, const uint32_t* __restrict__ steps, int B,
                  int64_t n_pts, int C, int n_cells, int n_chunks, T* __restrict__ gx, int64_t sb, int64_t sc) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr size_t kChanBytes = (size_t)kStreamChunk * sizeof(T);
  constexpr size_t kTileBytes = kChanBytes * CH;
  constexpr size_t kListBytes = (size_t)kListStride * 4;
  // [2 tiles] [2 list buffers] [grad_out rows: CH x n_cells, as T] [2 mbarriers]
  unsigned char* tiles = smem;
  uint32_t* lbuf = reinterpret_cast<uint32_t*>(smem + 2 * kTileBytes);
  T* g_s = reinterpret_cast<T*>(smem + 2 * kTileBytes + 2 * kListBytes);
  uint64_t* lfull = reinterpret_cast<uint64_t*>(smem + 2 * kTileBytes + 2 * kListBytes + align_up16((size_t)CH * n_cells * sizeof(T)));
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(lfull, 1); mbar_init(lfull + 1, 1); mbar_fence_init(); }
  __syncthreads();
  const int n_cg = (C + CH - 1) / CH;
  const int items = B * n_cg;
  const uint64_t pol_keep = l2_default_policy(), pol_stream = l2_evict_first_policy();
  uint32_t it = 0;                                               // chunks processed by this CTA so far (list barrier phases)
  auto request_list = [&](int b, int k, int buf) {               // thread 0 only
    const uint32_t rows = __ldg(steps + (size_t)b * n_chunks + k);
    mbar_expect_tx(lfull + buf, rows * 128u);
    bulk_g2s(lbuf + (size_t)buf * kListStride, lists + ((size_t)b * n_chunks + k) * kListStride, rows * 128u, lfull + buf, pol_keep);
  };
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / n_cg, c0 = (item % n_cg) * CH;
    const int nc = C - c0 < CH ? C - c0 : CH;
    __syncthreads();                                             // previous item's scatters have read g_s
    for (int i = tid; i < CH * n_cells; i += kBThreads) {
      const int c = i / n_cells;
      g_s[i] = from_f32<T>(c < nc ? __ldg(gout + ((size_t)b * C + c0) * n_cells + i) : 0.f);
    }
    if (tid == 0) { request_list(b, 0, (int)(it & 1u)); if (n_chunks > 1) request_list(b, 1, (int)((it + 1) & 1u)); }
    T* gxb = gx + (int64_t)b * sb + (int64_t)c0 * sc;
    for (int k = 0; k < n_chunks; ++k, ++it) {
      const int s = (int)(it & 1u);
      unsigned char* tile = tiles + s * kTileBytes;
      if (tid == 0) bulk_wait_read<1>();                         // the store that used this tile two chunks ago has read it
      __syncthreads();
      {                                                          // zero fill, 16 bytes per store
        uint4* t4 = reinterpret_cast<uint4*>(tile);
        for (int i = tid; i < (int)(kTileBytes / 16); i += kBThreads) t4[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(lfull + s, (it >> 1) & 1u);
      __syncthreads();
      const uint32_t* L = lbuf + (size_t)s * kListStride;
      const int n_ent = ((int)L[kSCh] - 1) * 32;                 // data rows x 32 (padding entries = kNone)
      T* tt = reinterpret_cast<T*>(tile);
      for (int i = tid; i < n_ent; i += kBThreads) {
        const uint32_t e = L[32 + i];
        if (e == kNone) continue;
        const uint32_t cell = e >> kStreamPosBits, pos = e & kPosMask;
#pragma unroll
        for (int c = 0; c < CH; ++c) tt[c * kStreamChunk + pos] = g_s[c * n_cells + cell];
      }
      fence_proxy_async();                                       // generic-proxy writes of the tile -> visible to the bulk copy
      __syncthreads();
      if (tid == 0) {
        const int64_t p0 = (int64_t)k * kStreamChunk;
        const uint32_t bytes = (uint32_t)((n_pts - p0 < kStreamChunk ? n_pts - p0 : kStreamChunk) * sizeof(T));
        for (int c = 0; c < nc; ++c) bulk_s2g(gxb + (int64_t)c * sc + p0, tile + c * kChanBytes, bytes, pol_stream);
        bulk_commit();
        if (k + 2 < n_chunks) request_list(b, k + 2, s);          // everybody is past this list buffer
      }
    }
  }
  if (tid == 0) bulk_wait_all<0>();                              // shared memory must outlive the last bulk reads
}

// ring depth: 2 x 64 KiB tiles for float32, 3 x 32 KiB for 16-bit inputs (leaves room for 8 x 2304 accumulators either way)
template <typename T> constexpr int stages_for() { return sizeof(T) == 4 ? 2 : 3; }

template <typename T>
size_t stream_smem_bytes(int n_cells) {
  return (size_t)stages_for<T>() * ((size_t)kSCh * kStreamChunk * sizeof(T) + (size_t)kListStride * 4) + align_up16((size_t)kSCh * n_cells * 4) +
         2 * stages_for<T>() * 8;
}

template <typename T>
int launch_stream(const T* x, int64_t sb, int64_t sc, const uint32_t* lists, const uint32_t* steps, int B, int64_t n_pts, int C,
                  int n_cells, int n_chunks, float* out, cudaStream_t st) {
  constexpr int NS = stages_for<T>();
  const size_t smem = stream_smem_bytes<T>(n_cells);
  cudaError_t e = cudaFuncSetAttribute(k_pool_stream<T, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  int sms = kNumSMsB200;
  { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int items = B * ((C + kSCh - 1) / kSCh);
  k_pool_stream<T, NS><<<(unsigned)(items < sms ? items : sms), kSThreads, smem, st>>>(x, sb, sc, lists, steps, B, n_pts, C, n_cells,
                                                                                      n_chunks, out, g_tuning[3]);
  MUVO_AFTER_LAUNCH("k_pool_stream", st);
  return MUVO_OK;
}

template <typename T>
size_t bwd_stream_smem_bytes(int n_cells) {
  constexpr int CH = bwd_channels<T>();
  return 2 * (size_t)CH * kStreamChunk * sizeof(T) + 2 * (size_t)kListStride * 4 + align_up16((size_t)CH * n_cells * sizeof(T)) + 16;
}

template <typename T>
int launch_bwd_stream(const float* gout, const uint32_t* lists, const uint32_t* steps, int B, int64_t n_pts, int C, int n_cells,
                      T* gx, int64_t sb, int64_t sc, cudaStream_t st) {
  const size_t smem = bwd_stream_smem_bytes<T>(n_cells);
  constexpr int CH = bwd_channels<T>();
  cudaError_t e = cudaFuncSetAttribute(k_pool_bwd_stream<T, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  int sms = kNumSMsB200;
  { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int items = B * ((C + CH - 1) / CH);
  const int n_chunks = (int)ceil_div64(n_pts, kStreamChunk);
  k_pool_bwd_stream<T, CH><<<(unsigned)(items < sms ? items : sms), kBThreads, smem, st>>>(gout, lists, steps, B, n_pts, C, n_cells, n_chunks,
                                                                                     gx, sb, sc);
  MUVO_AFTER_LAUNCH("k_pool_bwd_stream", st);
  return MUVO_OK;
}

}  // namespace

bool pool_bwd_stream_eligible(int elem_bytes, const void* gx, int64_t sb, int64_t sp, int64_t sc, int B, int64_t n_pts, int C, int n_cells) {
  if (!pool_stream_eligible(elem_bytes, gx, sb, sp, sc, B, n_pts, C, n_cells)) return false;
  const size_t smem = elem_bytes == 4 ? bwd_stream_smem_bytes<float>(n_cells) : bwd_stream_smem_bytes<__half>(n_cells);
  return smem <= 227 * 1024;
}

int pool_stream_bwd(const float* gout, const uint32_t* lists, const uint32_t* steps, int B, int64_t n_pts, int C, int n_cells, void* gx,
                    int32_t gx_dtype, int64_t sb, int64_t sc, cudaStream_t st) {
  switch (gx_dtype) {
    case MUVO_F32:  return launch_bwd_stream<float>(gout, lists, steps, B, n_pts, C, n_cells, (float*)gx, sb, sc, st);
    case MUVO_F16:  return launch_bwd_stream<__half>(gout, lists, steps, B, n_pts, C, n_cells, (__half*)gx, sb, sc, st);
    case MUVO_BF16: return launch_bwd_stream<__nv_bfloat16>(gout, lists, steps, B, n_pts, C, n_cells, (__nv_bfloat16*)gx, sb, sc, st);
    default: return MUVO_E_ARG;
  }
}

size_t stream_lists_bytes(int B, int64_t n_pts) { return (size_t)B * (size_t)ceil_div64(n_pts, kStreamChunk) * kListStride * 4; }   // (the plan uses the first 2048 of every 2368)
size_t stream_steps_bytes(int B, int64_t n_pts) { return (size_t)B * (size_t)ceil_div64(n_pts, kStreamChunk) * 4; }

bool pool_stream_eligible(int elem_bytes, const void* x, int64_t sb, int64_t sp, int64_t sc, int B, int64_t n_pts, int C, int n_cells) {
  if (g_tuning[2] == 1 || g_tuning[2] == 2) return false;     // tuning key 2: 1 / 2 = the gather kernels of bev.cu, 3 = always stream
  if (sp != 1 || n_pts <= 0 || B <= 0 || C <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (sb * elem_bytes) % 16 || (sc * elem_bytes) % 16 || (n_pts * elem_bytes) % 16) return false;
  if (n_cells >= (1 << (32 - kStreamPosBits)) - 1) return false;
  const size_t smem = elem_bytes == 4 ? stream_smem_bytes<float>(n_cells) : stream_smem_bytes<__half>(n_cells);
  if (smem > 227 * 1024) return false;
  // one CTA per (frame, 8 channels): with fewer items than ~half the SMs the row kernels (one CTA per channel row) are faster
  return g_tuning[2] == 3 || (int64_t)B * ((C + kSCh - 1) / kSCh) >= 64;
}

int pool_stream_plan(const int32_t* cell0, int B, int64_t n_pts, int n_cells, uint32_t* plan, uint32_t* plan_n, cudaStream_t st) {
  const int n_chunks = (int)ceil_div64(n_pts, kStreamChunk);
  k_chunk_lists<<<dim3((unsigned)n_chunks, (unsigned)B), kListThreads, 0, st>>>(cell0, nullptr, nullptr, n_pts, n_cells, n_chunks, plan, plan_n, 1);
  MUVO_AFTER_LAUNCH("k_chunk_lists", st);
  return MUVO_OK;
}

int pool_stream_fwd(const void* x, int32_t x_dtype, int64_t sb, int64_t sc, const int32_t* cell0, const uint8_t* mask,
                    int32_t* cell_out, const uint32_t* plan, const uint32_t* plan_n, int B, int64_t n_pts, int C, int n_cells,
                    float* out, uint32_t* lists, uint32_t* steps, cudaStream_t st) {
  const int n_chunks = (int)ceil_div64(n_pts, kStreamChunk);
  if (plan && plan_n) {
    k_chunk_compact<<<dim3((unsigned)n_chunks, (unsigned)B), kCompactThreads, 0, st>>>(plan, plan_n, cell0, mask, cell_out, n_pts, n_chunks, lists, steps);
    MUVO_AFTER_LAUNCH("k_chunk_compact", st);
  } else {
    k_chunk_lists<<<dim3((unsigned)n_chunks, (unsigned)B), kListThreads, 0, st>>>(cell0, mask, cell_out, n_pts, n_cells, n_chunks, lists, steps, 0);
    MUVO_AFTER_LAUNCH("k_chunk_lists", st);
  }
  switch (x_dtype) {
    case MUVO_F32:  return launch_stream<float>((const float*)x, sb, sc, lists, steps, B, n_pts, C, n_cells, n_chunks, out, st);
    case MUVO_F16:  return launch_stream<__half>((const __half*)x, sb, sc, lists, steps, B, n_pts, C, n_cells, n_chunks, out, st);
    case MUVO_BF16: return launch_stream<__nv_bfloat16>((const __nv_bfloat16*)x, sb, sc, lists, steps, B, n_pts, C, n_cells, n_chunks, out, st);
    default: return MUVO_E_ARG;
  }
}

}  // namespace muvo

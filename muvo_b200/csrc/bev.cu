// Stage (c): lift-splat BEV pooling (forward + backward) and the sorted-rank segment sum.
//
// Replaces FrustumPooling.voxel_pooling / QuickCumsum / cumsum_trick
// (muvo/models/frustum_pooling.py:23-60,131-187) and VoxelsSumming (muvo/layers/layers.py:326-357).
//
// The reference reshapes the lifted tensor to channels-last (a 236 MB/frame copy), compacts it twice,
// argsorts and gathers the *feature rows*, prefix-sums every row and differences the prefix sums.
// Here only point INDICES are sorted (stable counting sort by BEV cell, 4 B/point), and each cell's
// features are summed directly in ascending point order from wherever they live (any strides): the
// result is deterministic, more accurate than the cumsum trick, and x is read exactly once.
//
// Forward kernels
//   S1 k_cell_hist   : per warp-chunk (1024 consecutive points) histogram over cells (shared memory)
//   S2 k_cell_scan   : per (frame, cell) exclusive scan over warp-chunks  -> chunk bases + cell totals
//   S3 k_cell_starts : per frame exclusive scan over cells                -> cell_start[n_cells+1]
//   S4 k_cell_place  : stable placement of point ids                      -> sorted[b][...]
//   P  k_pool_*      : per (frame, cell) ordered sum over its points, all C channels; writes every
//                      output element (zeros for empty cells), so `out` needs no memset.
// Backward: pure gather, every grad_x element written once, in the producer's memory format.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "common.cuh"
#include "bev_stream.cuh"

namespace muvo {
namespace {

constexpr int kWarpChunk = 1024;   // points per warp-chunk (32 rounds of 32 lanes)
constexpr int kSortWarps = 4;      // warps per block in S1/S4

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }
template <typename T> __device__ __forceinline__ float ldf_reg(T v);
template <> __device__ __forceinline__ float ldf_reg<float>(float v) { return v; }
template <> __device__ __forceinline__ float ldf_reg<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float ldf_reg<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T cvt(float v);
template <> __device__ __forceinline__ float cvt<float>(float v) { return v; }
template <> __device__ __forceinline__ __half cvt<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

struct BevWs {
  uint32_t* chunk_base = nullptr;   // [B, n_wc, n_cells]  exclusive offset of (chunk, cell) inside its cell
  uint32_t* cell_total = nullptr;   // [B, n_cells]
  uint32_t* cell_start = nullptr;   // [B, n_cells + 1]
  int32_t* sorted = nullptr;        // [B, n_pts]  point ids grouped by cell, ascending inside a cell
  int32_t* kp = nullptr;            // [B, n_pts]  kept points in ascending point order ...
  int32_t* dest = nullptr;          // [B, n_pts]  ... and their position in `sorted`
  uint16_t* dest16 = nullptr;       // [B, n_pts]  `dest` as uint16 (saturated; read only for frames with <= 32768 kept points)
  uint32_t* chunk_kept = nullptr;   // [B, n_wc]   kept points per warp-chunk, scanned in place by S3
  uint32_t* lists = nullptr;        // [B, n_chunks, 2048]  streamed pool: per-chunk sorted (cell, position) lists, lane-interleaved
  uint32_t* steps = nullptr;        // [B, n_chunks]        ... and their length in warp steps
  int n_wc = 0;
  size_t bytes = 0;
};

static BevWs carve_bev(void* base, int B, int64_t n_pts, int n_cells) {
  BevWs w;
  char* b = (char*)base;
  size_t o = 0;
  w.n_wc = (int)ceil_div64(n_pts, kWarpChunk);
  w.chunk_base = (uint32_t*)(b + o); o = align_up(o + (size_t)B * w.n_wc * n_cells * 4, 256);
  w.cell_total = (uint32_t*)(b + o); o = align_up(o + (size_t)B * n_cells * 4, 256);
  w.cell_start = (uint32_t*)(b + o); o = align_up(o + (size_t)B * (n_cells + 1) * 4, 256);
  w.sorted = (int32_t*)(b + o);      o = align_up(o + (size_t)B * n_pts * 4, 256);
  w.kp = (int32_t*)(b + o);          o = align_up(o + (size_t)B * n_pts * 4, 256);
  w.dest = (int32_t*)(b + o);        o = align_up(o + (size_t)B * n_pts * 4, 256);
  w.dest16 = (uint16_t*)(b + o);     o = align_up(o + (size_t)B * n_pts * 2, 256);
  w.chunk_kept = (uint32_t*)(b + o); o = align_up(o + (size_t)B * w.n_wc * 4, 256);
  w.lists = (uint32_t*)(b + o);      o = align_up(o + stream_lists_bytes(B, n_pts), 256);
  w.steps = (uint32_t*)(b + o);      o = align_up(o + stream_steps_bytes(B, n_pts), 256);
  w.bytes = o;
  return w;
}

// S1: histogram of one warp-chunk; counts (<= 1024) written as u32 into chunk_base (scanned in place by S2).
__global__ void __launch_bounds__(kSortWarps * 32)
k_cell_hist(const int32_t* __restrict__ cell, int B, int64_t n_pts, int n_cells, int n_wc, uint32_t* __restrict__ chunk_base,
            uint32_t* __restrict__ chunk_kept) {
  extern __shared__ uint32_t sm[];                 // [kSortWarps][n_cells]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t wc_global = (int64_t)blockIdx.x * kSortWarps + warp;
  uint32_t* h = sm + (size_t)warp * n_cells;
  for (int i = lane; i < n_cells; i += 32) h[i] = 0;
  __syncwarp();
  if (wc_global >= (int64_t)B * n_wc) return;
  const int b = (int)(wc_global / n_wc), wc = (int)(wc_global % n_wc);
  const int32_t* cp = cell + (size_t)b * n_pts;
  const int64_t p0 = (int64_t)wc * kWarpChunk;
  uint32_t kept = 0;
  for (int r0 = 0; r0 < kWarpChunk / 32; r0 += 16) {
    int cc[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {          // 16 independent loads in flight (the kernel is latency bound: ~6 warps per SM)
      int64_t p = p0 + (r0 + u) * 32 + lane;
      int c = (p < n_pts) ? __ldg(cp + p) : -1;
      cc[u] = (c >= n_cells) ? -1 : c;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      unsigned m = __match_any_sync(0xffffffffu, cc[u]);
      kept += __popc(__ballot_sync(0xffffffffu, cc[u] >= 0));
      if (cc[u] >= 0 && lane == (__ffs(m) - 1)) h[cc[u]] += __popc(m);
      __syncwarp();
    }
  }
  if (lane == 0) chunk_kept[(size_t)b * n_wc + wc] = kept;
  uint32_t* dst = chunk_base + ((size_t)b * n_wc + wc) * n_cells;
  for (int i = lane; i < n_cells; i += 32) dst[i] = h[i];
}

// S2: per (frame, cell) exclusive scan over the warp-chunks (in place) + total.
__global__ void k_cell_scan(uint32_t* __restrict__ chunk_base, uint32_t* __restrict__ cell_total, int B, int n_cells, int n_wc) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)B * n_cells) return;
  int b = (int)(t / n_cells), c = (int)(t % n_cells);
  uint32_t* col = chunk_base + (size_t)b * n_wc * n_cells + c;
  uint32_t run = 0;
  int k = 0;
  for (; k + 32 <= n_wc; k += 32) {         // 32 independent loads, then the dependent stores (latency bound: few threads)
    uint32_t v[32];
#pragma unroll
    for (int u = 0; u < 32; ++u) v[u] = col[(size_t)(k + u) * n_cells];
#pragma unroll
    for (int u = 0; u < 32; ++u) { col[(size_t)(k + u) * n_cells] = run; run += v[u]; }
  }
  for (; k + 8 <= n_wc; k += 8) {
    uint32_t v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = col[(size_t)(k + u) * n_cells];
#pragma unroll
    for (int u = 0; u < 8; ++u) { col[(size_t)(k + u) * n_cells] = run; run += v[u]; }
  }
  for (; k < n_wc; ++k) { uint32_t v = col[(size_t)k * n_cells]; col[(size_t)k * n_cells] = run; run += v; }
  cell_total[t] = run;
}

// S3: per frame exclusive scan over cells (one block per frame).
__global__ void __launch_bounds__(1024)
k_cell_starts(const uint32_t* __restrict__ cell_total, uint32_t* __restrict__ cell_start, int n_cells,
              uint32_t* __restrict__ chunk_kept, int n_wc) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t* tot = cell_total + (size_t)b * n_cells;
  uint32_t* st = cell_start + (size_t)b * (n_cells + 1);
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_cells; base += 1024) {
    int i = base + threadIdx.x;
    uint32_t v = i < n_cells ? tot[i] : 0;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t carry = carry_s;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = wsum[lane], wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
      wsum[lane] = wi - w;
      if (lane == 31) carry_s = carry + wi;
    }
    __syncthreads();
    if (i < n_cells) st[i] = carry + wsum[warp] + incl - v;
    __syncthreads();
  }
  if (threadIdx.x == 0) st[n_cells] = carry_s;
  // exclusive scan of the kept-per-chunk counts (in place): rank of every kept point in ascending point order
  __syncthreads();
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  uint32_t* ck = chunk_kept + (size_t)b * n_wc;
  for (int base = 0; base < n_wc; base += 1024) {
    int i = base + threadIdx.x;
    uint32_t v = i < n_wc ? ck[i] : 0;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t carry = carry_s;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = wsum[lane], wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
      wsum[lane] = wi - w;
      if (lane == 31) carry_s = carry + wi;
    }
    __syncthreads();
    if (i < n_wc) ck[i] = carry + wsum[warp] + incl - v;
    __syncthreads();
  }
}

// S4: stable placement.
__global__ void __launch_bounds__(kSortWarps * 32)
k_cell_place(const int32_t* __restrict__ cell, int B, int64_t n_pts, int n_cells, int n_wc,
             const uint32_t* __restrict__ chunk_base, const uint32_t* __restrict__ cell_start, int32_t* __restrict__ sorted,
             const uint32_t* __restrict__ chunk_kept, int32_t* __restrict__ kp, int32_t* __restrict__ dest,
             uint16_t* __restrict__ dest16) {
  extern __shared__ uint32_t sm[];                 // running count per cell inside this warp-chunk
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t wc_global = (int64_t)blockIdx.x * kSortWarps + warp;
  uint32_t* run = sm + (size_t)warp * n_cells;
  for (int i = lane; i < n_cells; i += 32) run[i] = 0;
  __syncwarp();
  if (wc_global >= (int64_t)B * n_wc) return;
  const int b = (int)(wc_global / n_wc), wc = (int)(wc_global % n_wc);
  const int32_t* cp = cell + (size_t)b * n_pts;
  const uint32_t* cb = chunk_base + ((size_t)b * n_wc + wc) * n_cells;
  const uint32_t* cs = cell_start + (size_t)b * (n_cells + 1);
  int32_t* out = sorted + (size_t)b * n_pts;
  int32_t* kpo = kp + (size_t)b * n_pts;
  int32_t* dso = dest + (size_t)b * n_pts;
  uint16_t* d16 = dest16 + (size_t)b * n_pts;
  uint32_t krun = chunk_kept[(size_t)b * n_wc + wc];   // kept points before this chunk (frame-relative)
  const int64_t p0 = (int64_t)wc * kWarpChunk;
  for (int r0 = 0; r0 < kWarpChunk / 32; r0 += 8) {
    int cc[8];
    uint32_t basepos[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      int64_t p = p0 + (r0 + u) * 32 + lane;
      int c = (p < n_pts) ? __ldg(cp + p) : -1;
      cc[u] = (c >= n_cells) ? -1 : c;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) basepos[u] = cc[u] >= 0 ? __ldg(cs + cc[u]) + __ldg(cb + cc[u]) : 0u;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = cc[u];
      unsigned m = __match_any_sync(0xffffffffu, c);
      const int64_t pp = p0 + (r0 + u) * 32 + lane;
      const unsigned vb = __ballot_sync(0xffffffffu, c >= 0);
      if (c >= 0) {
        uint32_t pos = basepos[u] + run[c] + __popc(m & ((1u << lane) - 1u));
        uint32_t kr = krun + __popc(vb & ((1u << lane) - 1u));
        out[pos] = (int32_t)pp;
        kpo[kr] = (int32_t)pp;
        dso[kr] = (int32_t)pos;
        d16[kr] = (uint16_t)(pos < 65535u ? pos : 65535u);
      }
      krun += __popc(vb);
      __syncwarp();
      if (c >= 0 && lane == (__ffs(m) - 1)) run[c] += __popc(m);
      __syncwarp();
    }
  }
}

// (Streaming the whole row through shared memory with cp.async.bulk, 2-3 stages of 32-40 KiB next to the 128 KiB staging
// buffer, was measured at 436-501 us against 385 us for the gathers below: with one CTA per SM the bulk copies keep
// < 100 KB in flight per SM, the 16-deep gathers keep ~260 KB.)
// P (point-major, x_stride_p == 1): one CTA per (frame, channel).  The channel's row x[b, c, :] is streamed in
// ADDRESS ORDER (HBM friendly; only sectors that hold a kept point are requested), every kept value is dropped
// into shared memory at its position in the cell-sorted order, and each cell's now contiguous segment is summed
// (lane-strided partials + fixed xor tree -> deterministic).  Frames with more kept points than fit in shared
// memory are processed in windows of whole cells.
constexpr int kRowThreads = 1024;
constexpr int kRowCap = 52 * 1024;   // floats of staging (208 KB)
constexpr int kBigCell = 64;         // cells above this many points are summed by the whole warp
template <typename T>
__global__ void __launch_bounds__(kRowThreads, 1)
k_pool_rows(const T* __restrict__ x, int64_t sb, int64_t sc, const int32_t* __restrict__ kp_all, const int32_t* __restrict__ dest_all,
            const uint32_t* __restrict__ cell_start, const int32_t* __restrict__ sorted, int B, int64_t n_pts, int C,
            int n_cells, float* __restrict__ out) {
  extern __shared__ float buf[];
  __shared__ int n_big;
  uint16_t* big_list = reinterpret_cast<uint16_t*>(buf + kRowCap);   // [n_cells] cells of the current window with > kBigCell points
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / C, c = blockIdx.x % C;
  const uint32_t* cs = cell_start + (size_t)b * (n_cells + 1);
  const int32_t* kp = kp_all + (size_t)b * n_pts;
  const int32_t* ds = dest_all + (size_t)b * n_pts;
  const int n_kept = (int)cs[n_cells];
  const T* xr = x + (size_t)b * sb + (size_t)c * sc;
  float* o = out + ((size_t)b * C + c) * n_cells;
  int c0 = 0;
  while (c0 < n_cells) {
    const uint32_t w0 = cs[c0];
    int lo = c0 + 1, hi = n_cells;            // largest c1 in [c0+1, n_cells] with cs[c1] - w0 <= kRowCap
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (cs[mid] - w0 <= (uint32_t)kRowCap) lo = mid; else hi = mid - 1; }
    const int c1 = lo;
    const uint32_t w1 = cs[c1];
    if (w1 - w0 <= (uint32_t)kRowCap) {
      if (w1 > w0) {
        const int iw0 = (int)w0, iw1 = (int)w1;
        for (int j0 = 0; j0 < n_kept; j0 += kRowThreads * 8) {
          int d[8], pp[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {          // coalesced index loads, 16 in flight
            int j = j0 + u * kRowThreads + tid;
            bool in = j < n_kept;
            d[u] = in ? __ldg(ds + j) : -1;
            pp[u] = in ? __ldg(kp + j) : 0;
          }
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)            // address-ordered gathers of the kept points only
            v[u] = (d[u] >= iw0 && d[u] < iw1) ? ldf<T>(xr + pp[u]) : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (d[u] >= iw0 && d[u] < iw1) buf[d[u] - iw0] = v[u];
        }
      }
      __syncthreads();
      // Segment sums: one THREAD per cell for the common small cells (sequential, ascending point order); cells above
      // kBigCell points (they are neighbours: the cells next to the camera) are collected and then dealt round-robin to
      // ALL warps of the CTA (lane-strided partials + fixed xor tree), instead of leaving them to the one warp that
      // owns their index range.  A warp step of the first pass covers 32 consecutive cells -> one coalesced store.
      if (tid == 0) n_big = 0;
      __syncthreads();
      for (int cb = c0 + warp * 32; cb < c1; cb += (kRowThreads / 32) * 32) {
        const int cell = cb + lane;
        uint32_t s0 = 0, s1 = 0;
        if (cell < c1) { s0 = cs[cell] - w0; s1 = cs[cell + 1] - w0; }
        const bool big = (s1 - s0) > (uint32_t)kBigCell;
        const unsigned bm = __ballot_sync(0xffffffffu, big);    // one shared-memory atomic per warp step, not per big cell
        if (bm) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&n_big, __popc(bm));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (big) big_list[base + __popc(bm & ((1u << lane) - 1u))] = (uint16_t)(cell - c0);
        }
        float acc = 0.f;
        if (!big) for (uint32_t j = s0; j < s1; ++j) acc += buf[j];
        if (cell < c1 && !big) o[cell] = acc;
      }
      __syncthreads();
      for (int i = warp; i < n_big; i += kRowThreads / 32) {
        const int cell = c0 + (int)big_list[i];
        const uint32_t b0 = cs[cell] - w0, b1 = cs[cell + 1] - w0;
        float part = 0.f;
        for (uint32_t j = b0 + lane; j < b1; j += 32) part += buf[j];
#pragma unroll
        for (int dd = 16; dd; dd >>= 1) part += __shfl_xor_sync(0xffffffffu, part, dd);
        if (lane == 0) o[cell] = part;
      }
      __syncthreads();
    } else {
      // a single cell larger than the staging buffer: gather it straight from global memory (rare)
      if (warp == 0) {
        const int32_t* list = sorted + (size_t)b * n_pts;
        float acc = 0.f;
        for (uint32_t j = w0 + lane; j < w1; j += 32) acc += ldf<T>(xr + list[j]);
#pragma unroll
        for (int dd = 16; dd; dd >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, dd);
        if (lane == 0) o[c0] = acc;
      }
    }
    c0 = c1;
  }
}

// (Streaming the row with coalesced 16-byte loads next to a dense int32 position map was also measured: 584 us.  The map
// doubles the L2 -> SM traffic, 4.4 GB per forward, and that fabric tops out near 7.5 TB/s.)
//
// Persistent rows (the default): one CTA per SM walks ~16 consecutive (frame, channel) rows.  A thread owns up to
// kPipeItems kept points of a row (address-ordered pairs, one 8-byte index load per pair), issues all its gathers at once
// and does so for row r+1 BEFORE the segment sums of row r; the frame's cell starts and its list of big cells are staged
// in shared memory once per frame; the positions of a thread's pairs arrive in one batch of 32-bit loads (uint16 copy of
// `dest`).  Rows of frames with > 32768 kept points take the windowed code of k_pool_rows via pool_row_windowed.
//
// Measured at cfg3 (6 frames, C = 384, 30 k kept points per row; forward incl. the 62 us index sort):
//   k_pool_rows, one CTA per row, 8 gathers in flight per thread                                   427 us  (16 in flight: 540 us)
//   this kernel                                                                                    402 us
//   ... without the scatter into the row buffer: -24 us; without the segment sums: -94 us; without both: 299 us
//   gathering in cell-sorted order (coalesced stores, no position list)                           1038 us
//   streaming the row with 16-byte loads next to a dense int32 position map                         644 us
//   warp-specialised (8 producer warps gather half a row, 8 consumer warps sum the other half)      771 us
//   one CTA per (row, half of the cells), 64 KB staging, three CTAs per SM                          513 us
// ncu (profiles/r1_bev_pool_rows_ncu.txt): DRAM 36 % of peak, no unit above 50 %, long-scoreboard stalls.  The gather rate
// follows the number of WARPS that issue loads (8 warps: half the rate of 16 or 32), not the loads each thread has in
// flight, so the shared-memory phases cannot be hidden behind the gathers by the same warps and taking warps away from
// the gathers costs more than the overlap returns; splitting a row by cells so that several small CTAs share an SM makes
// every half fetch the sectors it shares with the other one (neighbouring points fall into different cells).  The
// fused lift-splat (N2) avoids the problem: it never reads `x`.
constexpr int kPipeThreads = 512;                      // 128 registers per thread: 64 of them hold the row in flight
constexpr int kPipeItems = 64;                         // kept points per thread: kPipeItems * kPipeThreads = 32768 per row

// segment sums of the cells [c0, c1) whose values sit in buf[cs[cell] - w0 ...): see k_pool_rows
__device__ __forceinline__ void pool_segments(const float* buf, uint16_t* big_list, int* n_big, const uint32_t* __restrict__ cs,
                                              int c0, int c1, uint32_t w0, float* __restrict__ o) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int cb = c0 + warp * 32; cb < c1; cb += (kPipeThreads / 32) * 32) {
    const int cell = cb + lane;
    uint32_t s0 = 0, s1 = 0;
    if (cell < c1) { s0 = cs[cell] - w0; s1 = cs[cell + 1] - w0; }
    const bool big = (s1 - s0) > (uint32_t)kBigCell;
    const unsigned bm = __ballot_sync(0xffffffffu, big);        // one shared-memory atomic per warp step, not per big cell
    if (bm) {
      int base = 0;
      if (lane == 0) base = atomicAdd(n_big, __popc(bm));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (big) big_list[base + __popc(bm & ((1u << lane) - 1u))] = (uint16_t)(cell - c0);
    }
    float acc = 0.f;
    if (!big) for (uint32_t j = s0; j < s1; ++j) acc += buf[j];
    if (cell < c1 && !big) o[cell] = acc;
  }
  __syncthreads();
  const int nb = *n_big;
  for (int i = warp; i < nb; i += kPipeThreads / 32) {
    const int cell = c0 + (int)big_list[i];
    const uint32_t b0 = cs[cell] - w0, b1 = cs[cell + 1] - w0;
    float part = 0.f;
    for (uint32_t j = b0 + lane; j < b1; j += 32) part += buf[j];
#pragma unroll
    for (int dd = 16; dd; dd >>= 1) part += __shfl_xor_sync(0xffffffffu, part, dd);
    if (lane == 0) o[cell] = part;
  }
}

// The same sums with the big-cell list of the frame built once (pool_big_list) and reused by every channel row: no atomics,
// no barrier between the two parts.
__device__ __forceinline__ void pool_big_list(uint16_t* big_list, int* n_big, const uint32_t* cs, int n_cells) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) *n_big = 0;
  __syncthreads();
  for (int cb = warp * 32; cb < n_cells; cb += (kPipeThreads / 32) * 32) {
    const int cell = cb + lane;
    const bool big = cell < n_cells && (cs[cell + 1] - cs[cell]) > (uint32_t)kBigCell;
    const unsigned bm = __ballot_sync(0xffffffffu, big);
    if (bm) {
      int base = 0;
      if (lane == 0) base = atomicAdd(n_big, __popc(bm));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (big) big_list[base + __popc(bm & ((1u << lane) - 1u))] = (uint16_t)cell;
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void pool_segments_pre(const float* buf, const uint16_t* big_list, int nb, const uint32_t* cs, int n_cells,
                                                  float* __restrict__ o) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = warp; i < nb; i += kPipeThreads / 32) {        // big cells first: they are the long poles
    const int cell = (int)big_list[i];
    const uint32_t b0 = cs[cell], b1 = cs[cell + 1];
    float part = 0.f;
    for (uint32_t j = b0 + lane; j < b1; j += 32) part += buf[j];
#pragma unroll
    for (int dd = 16; dd; dd >>= 1) part += __shfl_xor_sync(0xffffffffu, part, dd);
    if (lane == 0) o[cell] = part;
  }
  for (int cell = tid; cell < n_cells; cell += kPipeThreads) {
    const uint32_t s0 = cs[cell], s1 = cs[cell + 1];
    if (s1 - s0 > (uint32_t)kBigCell) continue;
    float acc = 0.f;
    for (uint32_t j = s0; j < s1; ++j) acc += buf[j];
    o[cell] = acc;
  }
}

// one row with the windowed gather code (any number of kept points); all threads of the CTA
template <typename T>
__device__ __noinline__ void pool_row_windowed(const T* __restrict__ xr, const int32_t* __restrict__ kp, const int32_t* __restrict__ ds,
                                  const uint32_t* __restrict__ cs, const int32_t* __restrict__ list, int n_cells, float* buf,
                                  uint16_t* big_list, int* n_big, float* __restrict__ o) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_kept = (int)cs[n_cells];
  int c0 = 0;
  while (c0 < n_cells) {
    const uint32_t w0 = cs[c0];
    int lo = c0 + 1, hi = n_cells;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (cs[mid] - w0 <= (uint32_t)kRowCap) lo = mid; else hi = mid - 1; }
    const int c1 = lo;
    const uint32_t w1 = cs[c1];
    if (w1 - w0 <= (uint32_t)kRowCap) {
      const int iw0 = (int)w0, iw1 = (int)w1;
      for (int j0 = 0; j0 < n_kept && w1 > w0; j0 += kPipeThreads * 8) {
        int d[8], pp[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          int j = j0 + u * kPipeThreads + tid;
          bool in = j < n_kept;
          d[u] = in ? __ldg(ds + j) : -1;
          pp[u] = in ? __ldg(kp + j) : 0;
        }
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (d[u] >= iw0 && d[u] < iw1) ? ldf<T>(xr + pp[u]) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (d[u] >= iw0 && d[u] < iw1) buf[d[u] - iw0] = v[u];
      }
      if (tid == 0) *n_big = 0;
      __syncthreads();
      pool_segments(buf, big_list, n_big, cs, c0, c1, w0, o);
      __syncthreads();
    } else {
      if (warp == 0) {
        float acc = 0.f;
        for (uint32_t j = w0 + lane; j < w1; j += 32) acc += ldf<T>(xr + list[j]);
#pragma unroll
        for (int dd = 16; dd; dd >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, dd);
        if (lane == 0) o[c0] = acc;
      }
    }
    c0 = c1;
  }
}

template <typename T>
__global__ void __launch_bounds__(kPipeThreads, 1)
k_pool_rows_pipe(const T* __restrict__ x, int64_t sb, int64_t sc, const int32_t* __restrict__ kp_all,
                 const int32_t* __restrict__ dest_all, const uint16_t* __restrict__ dest16_all,
                 const uint32_t* __restrict__ cell_start, const int32_t* __restrict__ sorted_all, int B, int64_t n_pts, int C,
                 int n_cells, float* __restrict__ out) {
  extern __shared__ float buf[];
  __shared__ int n_big;
  uint16_t* big_list = reinterpret_cast<uint16_t*>(buf + kRowCap);
  uint32_t* cs_s = reinterpret_cast<uint32_t*>(buf + kRowCap) + (n_cells + 1) / 2 + 1;   // [n_cells + 1] cell starts of the current frame
  const int tid = threadIdx.x;
  const int n_rows = B * C;
  // contiguous rows per CTA: ~n_rows / gridDim.x consecutive channels of (mostly) one frame
  const int r_lo = (int)((int64_t)n_rows * blockIdx.x / gridDim.x), r_hi = (int)((int64_t)n_rows * (blockIdx.x + 1) / gridDim.x);
  if (r_lo >= r_hi) return;
  constexpr int kPairs = kPipeItems / 2;
  T v[kPipeItems];                                       // raw elements: converting here would wait for every load
  // All gathers of row `row` into v[], in ADDRESS order (kept list): thread t owns the pairs t, t + 512, ...; v[2i], v[2i+1]
  // = items 2 (i * 512 + t) and the next one, fetched with one 8-byte index load.  Asynchronous: nothing here waits for
  // the values.  (Gathering in cell-sorted order instead -- no position list, coalesced stores -- was 2.5x slower: neighbours
  // in a sector belong to different cells, so every sector is requested again and again.)
  auto issue = [&](int row, int n_kept) {
    const int b = row / C, c = row - b * C;
    const int2* kp2 = reinterpret_cast<const int2*>(kp_all + (size_t)b * n_pts);
    const T* xr = x + (size_t)b * sb + (size_t)c * sc;
    int2 pp[kPairs];
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      const int k = i * kPipeThreads + tid;
      pp[i] = 2 * k < n_kept ? __ldg(kp2 + k) : make_int2(-1, -1);
      if (2 * k + 1 >= n_kept) pp[i].y = -1;
    }
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      v[2 * i] = pp[i].x >= 0 ? __ldg(xr + pp[i].x) : cvt<T>(0.f);
      v[2 * i + 1] = pp[i].y >= 0 ? __ldg(xr + pp[i].y) : cvt<T>(0.f);
    }
  };
  auto kept_of = [&](int row) { return (int)__ldg(cell_start + (size_t)(row / C) * (n_cells + 1) + n_cells); };

  int row = r_lo;
  int n_kept = kept_of(row);
  bool piped = n_kept <= kPipeItems * kPipeThreads;
  if (piped) issue(row, n_kept);
  int b_staged = -1;
  while (row < r_hi) {
    const int b = row / C, c = row - b * C;
    if (b != b_staged) {                                  // (the previous row's segment sums ended with a barrier)
      const uint32_t* cs = cell_start + (size_t)b * (n_cells + 1);
      for (int i = tid; i <= n_cells; i += kPipeThreads) cs_s[i] = __ldg(cs + i);
      __syncthreads();
      pool_big_list(big_list, &n_big, cs_s, n_cells);      // once per frame, shared by its channel rows
      b_staged = b;
    }
    float* o = out + ((size_t)b * C + c) * n_cells;
    const int next = row + 1;
    const int next_kept = next < r_hi ? (next / C == b ? n_kept : kept_of(next)) : 0;
    const bool next_piped = next < r_hi && next_kept <= kPipeItems * kPipeThreads;
    if (piped) {
      // positions of this thread's pairs: ALL loads first (one L2 round trip, the gathers are still in flight), then the
      // scatter into the row buffer
      const uint32_t* d2 = reinterpret_cast<const uint32_t*>(dest16_all + (size_t)b * n_pts);
      uint32_t d[kPairs];
#pragma unroll
      for (int i = 0; i < kPairs; ++i) { const int k = i * kPipeThreads + tid; d[i] = 2 * k < n_kept ? __ldg(d2 + k) : 0u; }
#pragma unroll
      for (int i = 0; i < kPairs; ++i) {
        const int k = i * kPipeThreads + tid;
        if (2 * k < n_kept) buf[d[i] & 0xffffu] = ldf_reg<T>(v[2 * i]);
        if (2 * k + 1 < n_kept) buf[d[i] >> 16] = ldf_reg<T>(v[2 * i + 1]);
      }
      __syncthreads();
      if (next_piped) issue(next, next_kept);             // next row's DRAM round trips run under this row's segment sums
      pool_segments_pre(buf, big_list, n_big, cs_s, n_cells, o);
      __syncthreads();
    } else {
      __syncthreads();
      b_staged = -1;                                      // the windowed code rebuilds the big-cell list per window
      pool_row_windowed<T>(x + (size_t)b * sb + (size_t)c * sc, kp_all + (size_t)b * n_pts, dest_all + (size_t)b * n_pts, cs_s,
                           sorted_all + (size_t)b * n_pts, n_cells, buf, big_list, &n_big, o);
      if (next_piped) issue(next, next_kept);
    }
    row = next; n_kept = next_kept; piped = next_piped;
  }
}

// P (generic strides; coalesced when x_stride_c == 1): one block per (frame, cell), threads over channels,
// sequential ascending-point sum.
template <typename T>
__global__ void __launch_bounds__(128)
k_pool_channel_major(const T* __restrict__ x, int64_t sb, int64_t sp, int64_t sc, const uint32_t* __restrict__ cell_start,
                     const int32_t* __restrict__ sorted, int B, int64_t n_pts, int C, int n_cells, float* __restrict__ out) {
  const int b = blockIdx.x / n_cells, c = blockIdx.x % n_cells;
  const uint32_t* cs = cell_start + (size_t)b * (n_cells + 1);
  const uint32_t s0 = cs[c], s1 = cs[c + 1];
  const int32_t* list = sorted + (size_t)b * n_pts;
  float* o = out + (size_t)b * C * n_cells + c;
  const T* xb = x + (size_t)b * sb;
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    float acc = 0.f;
    for (uint32_t j = s0; j < s1; ++j) acc += ldf<T>(xb + (size_t)list[j] * sp + (size_t)ch * sc);
    o[(size_t)ch * n_cells] = acc;
  }
}

// Backward: grad_x[b,p,c] = cell >= 0 ? grad_out[b,c,cell] : 0.
// FAST_P: grad_x point-contiguous (stride_p == 1): thread per 4 consecutive points of one channel.
template <typename T>
__global__ void __launch_bounds__(256)
k_pool_bwd_point_major(const float* __restrict__ gout, const int32_t* __restrict__ cell, int B, int64_t n_pts, int C,
                       int n_cells, T* __restrict__ gx, int64_t sb, int64_t sc) {
  const int64_t quads = ceil_div64(n_pts, 4);
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)B * C * quads) return;
  const int64_t q = t % quads;
  const int64_t bc = t / quads;
  const int ch = (int)(bc % C), b = (int)(bc / C);
  const int64_t p0 = q * 4;
  const int32_t* cp = cell + (size_t)b * n_pts + p0;
  const float* g = gout + ((size_t)b * C + ch) * n_cells;
  T* dst = gx + (size_t)b * sb + (size_t)ch * sc + p0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (p0 + k < n_pts) {
      int c = __ldg(cp + k);
      dst[k] = cvt<T>((c >= 0 && c < n_cells) ? __ldg(g + c) : 0.f);
    }
  }
}
// float32 fast path: a thread owns 4 consecutive points and walks kBwdCh channels with them, so the cell ids are
// read once per kBwdCh channels (not once per channel); the grad_out planes of those channels (9 KB each) stay
// in L1; every store is a coalesced 16-byte streaming store.
constexpr int kBwdCh = 8;
__global__ void __launch_bounds__(256)
k_pool_bwd_rows_f32(const float* __restrict__ gout, const int32_t* __restrict__ cell, int B, int64_t n_pts, int C,
                    int n_cells, float* __restrict__ gx, int64_t sb, int64_t sc) {
  const int64_t quads = n_pts / 4;                      // host guarantees n_pts % 4 == 0 and 16-byte alignment
  const int n_cg = (C + kBwdCh - 1) / kBwdCh;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)B * n_cg * quads) return;
  const int64_t q = t % quads;
  const int64_t bg = t / quads;
  const int cg = (int)(bg % n_cg), b = (int)(bg / n_cg);
  const int64_t p0 = q * 4;
  const int4 c4 = __ldg(reinterpret_cast<const int4*>(cell + (size_t)b * n_pts + p0));
  const bool k0 = c4.x >= 0 && c4.x < n_cells, k1 = c4.y >= 0 && c4.y < n_cells, k2 = c4.z >= 0 && c4.z < n_cells,
             k3 = c4.w >= 0 && c4.w < n_cells;
  const int ch0 = cg * kBwdCh;
#pragma unroll
  for (int u = 0; u < kBwdCh; ++u) {
    const int ch = ch0 + u;
    if (ch < C) {
      const float* g = gout + ((size_t)b * C + ch) * n_cells;
      float4 v;
      v.x = k0 ? __ldg(g + c4.x) : 0.f;
      v.y = k1 ? __ldg(g + c4.y) : 0.f;
      v.z = k2 ? __ldg(g + c4.z) : 0.f;
      v.w = k3 ? __ldg(g + c4.w) : 0.f;
      st_stream_f4(reinterpret_cast<float4*>(gx + (size_t)b * sb + (size_t)ch * sc + p0), v);
    }
  }
}

// generic strides: thread per (b, p, c) with c fastest (coalesced when stride_c == 1)
template <typename T>
__global__ void __launch_bounds__(256)
k_pool_bwd_generic(const float* __restrict__ gout, const int32_t* __restrict__ cell, int B, int64_t n_pts, int C, int n_cells,
                   T* __restrict__ gx, int64_t sb, int64_t sp, int64_t sc) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)B * n_pts * C) return;
  const int ch = (int)(t % C);
  const int64_t bp = t / C;
  const int64_t p = bp % n_pts;
  const int b = (int)(bp / n_pts);
  int c = __ldg(cell + (size_t)b * n_pts + p);
  float v = (c >= 0 && c < n_cells) ? __ldg(gout + ((size_t)b * C + ch) * n_cells + c) : 0.f;
  gx[(size_t)b * sb + (size_t)p * sp + (size_t)ch * sc] = cvt<T>(v);
}

// ---------------------------------------------------------------- sorted-rank segment sum (QuickCumsum API)
constexpr int kScanBlock = 1024;

// kept[i] = (i == n-1) || ranks[i+1] != ranks[i]  (frustum_pooling.py:38-39); per-block count of kept flags
__global__ void __launch_bounds__(kScanBlock)
k_seg_block_count(const int64_t* __restrict__ ranks, int64_t n, uint32_t* __restrict__ block_cnt) {
  __shared__ uint32_t ws[32];
  int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  uint32_t k = 0;
  if (i < n) k = (i == n - 1) || (ranks[i + 1] != ranks[i]);
  unsigned bal = __ballot_sync(0xffffffffu, k);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = __popc(bal);
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t v = ws[threadIdx.x];
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = v;
  }
}
// single block: exclusive scan of block counts (in place) + n_seg
__global__ void __launch_bounds__(1024)
k_seg_scan_blocks(uint32_t* __restrict__ block_cnt, int nblocks, int32_t* __restrict__ n_seg_out) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    int i = base + threadIdx.x;
    uint32_t v = i < nblocks ? block_cnt[i] : 0, incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t carry = carry_s;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = wsum[lane], wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
      wsum[lane] = wi - w;
      if (lane == 31) carry_s = carry + wi;
    }
    __syncthreads();
    if (i < nblocks) block_cnt[i] = carry + wsum[warp] + incl - v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_seg_out = (int32_t)carry_s;
}
// seg_id[i] = #kept before i ; last_row[seg_id] = i for kept rows
__global__ void __launch_bounds__(kScanBlock)
k_seg_assign(const int64_t* __restrict__ ranks, int64_t n, const uint32_t* __restrict__ block_off, int32_t* __restrict__ seg_id,
             int64_t* __restrict__ last_row) {
  __shared__ uint32_t ws[32];
  int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t k = 0;
  if (i < n) k = (i == n - 1) || (ranks[i + 1] != ranks[i]);
  unsigned bal = __ballot_sync(0xffffffffu, k);
  if (lane == 0) ws[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    uint32_t v = ws[lane], incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    ws[lane] = incl - v;
  }
  __syncthreads();
  if (i < n) {
    uint32_t id = block_off[blockIdx.x] + ws[warp] + __popc(bal & ((1u << lane) - 1u));
    seg_id[i] = (int32_t)id;
    if (k) last_row[id] = i;
  }
}
// x_seg[s, :] = sum of rows (last_row[s-1], last_row[s]] in ascending row order; threads over channels
__global__ void __launch_bounds__(256)
k_seg_sum(const float* __restrict__ x, const int64_t* __restrict__ last_row, const int32_t* __restrict__ n_seg, int C,
          float* __restrict__ x_seg) {
  const int ns = *n_seg;
  for (int s = blockIdx.x; s < ns; s += gridDim.x) {
    const int64_t r1 = last_row[s], r0 = s ? last_row[s - 1] + 1 : 0;
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      float acc = 0.f;
      for (int64_t r = r0; r <= r1; ++r) acc += __ldg(x + (size_t)r * C + ch);
      x_seg[(size_t)s * C + ch] = acc;
    }
  }
}
__global__ void __launch_bounds__(256)
k_seg_bwd(const float* __restrict__ gseg, const int32_t* __restrict__ seg_id, int64_t n, int C, float* __restrict__ gx) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * C) return;
  int64_t i = t / C;
  int ch = (int)(t - i * C);
  gx[t] = __ldg(gseg + (size_t)__ldg(seg_id + i) * C + ch);
}


// ---------------------------------------------------------------- N2: fused lift-splat (mile.py:508-523 + frustum_pooling.py:131-187)
// The reference materialises x = depth (B,D,H,W) (outer) feat (B,C,H,W) -- 236 MB per frame at muvo.yml shapes -- and
// pools it.  Fused: out[b,c,cell] = sum over the cell's kept frustum points p = (d, hw), ascending p, of
// fl(depth[b,d,hw] * feat[b,c,hw]) -- the same products and the same summation order as lifting and then pooling,
// without ever writing the product.  feat is taken channels-last ([B, HW, C]) so a point reads one contiguous C-vector.
//   forward : one CTA per (frame, 8 consecutive cells), threads over channels; every thread owns out[b, c, cell0..cell0+7]
//             (one 32-byte sector) and walks the cells' sorted point lists.
//   backward: one warp per pixel (b, hw), lanes over channels; it owns grad_feat[b, hw, :] (sum over the pixel's kept
//             depth bins in ascending d) and produces grad_depth[b, d, hw] by a fixed xor-tree reduction: no atomics.
#ifndef MUVO_LS_CELLS
#define MUVO_LS_CELLS 4
#endif
constexpr int kLsCells = MUVO_LS_CELLS;   // cells per CTA (multiple of 4: the outputs of 4 cells form one 16-byte store)
constexpr int kLsThreads = 512;
constexpr int kLsStage = 2048;     // points of a cell staged in shared memory per round
// C % 4 == 0: a thread owns 4 consecutive channels (one 16-byte load per point); the CTA's threads form
// G = 512 / (C/4) groups that split every cell's point list round-robin (group g takes points g, g+G, ...) -- the
// heavy cells near the camera hold > 1000 points -- and the G partial sums are added in group order, so the result is
// a fixed function of the inputs.  The point list and the depth values are staged through shared memory first, so the
// only global loads in the inner loop are the independent feat vectors.  kLsCells = 4 instead of 8 cut the kernel from 97 to
// 65 us: the cells next to the camera hold a hundred times the points of the far ones, and the heavy CTAs are the tail.
__global__ void __launch_bounds__(kLsThreads)
k_lift_splat_fwd(const float* __restrict__ feat_cl, const float* __restrict__ depth, const uint32_t* __restrict__ cell_start,
                 const int32_t* __restrict__ sorted, int B, int64_t n_pts, int HW, int C, int n_cells, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char ls_raw[];
  int* hw_s = reinterpret_cast<int*>(ls_raw);                       // [kLsStage]
  float* dep_s = reinterpret_cast<float*>(ls_raw) + kLsStage;       // [kLsStage]
  float4* red = reinterpret_cast<float4*>(dep_s + kLsStage);        // [kLsCells][G][C4]
  __shared__ uint32_t cs_s[kLsCells + 1];
  const int C4 = C >> 2;
  const int G = kLsThreads / C4 > 0 ? kLsThreads / C4 : 1;
  const int tid = threadIdx.x;
  const int g = tid / C4, q = tid - g * C4;                          // group, channel quad
  const bool active = g < G;
  const int groups = (n_cells + kLsCells - 1) / kLsCells;
  const int b = blockIdx.x / groups, cell0 = (blockIdx.x % groups) * kLsCells;
  const uint32_t* cs = cell_start + (size_t)b * (n_cells + 1);
  const int32_t* list = sorted + (size_t)b * n_pts;
  const float4* fb = reinterpret_cast<const float4*>(feat_cl + (size_t)b * HW * C);
  const float* db = depth + (size_t)b * n_pts;
  if (tid <= kLsCells) cs_s[tid] = cs[cell0 + tid < n_cells ? cell0 + tid : n_cells];
  __syncthreads();
  // The point lists of the CTA's cells are one contiguous range of `sorted`: it is staged (point -> pixel, depth) in rounds of
  // kLsStage points that span cell boundaries, so the two dependent global loads and the barriers are paid once per
  // round, not once per cell; inside a round every cell's sub-range is split round-robin over the G groups.
  float4 acc[kLsCells];
#pragma unroll
  for (int k = 0; k < kLsCells; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t S0 = cs_s[0], S1 = cs_s[kLsCells];
  for (uint32_t jb = S0; jb < S1; jb += kLsStage) {
    const uint32_t je = S1 - jb < (uint32_t)kLsStage ? S1 : jb + kLsStage;
    if (jb != S0) __syncthreads();                                   // previous round's readers are done
    for (int t = tid; t < (int)(je - jb); t += kLsThreads) {
      const int p = list[jb + t];
      hw_s[t] = p % HW;
      dep_s[t] = __ldg(db + p);
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int k = 0; k < kLsCells; ++k) {
        const uint32_t a0 = cs_s[k] > jb ? cs_s[k] : jb, a1 = cs_s[k + 1] < je ? cs_s[k + 1] : je;   // this cell's part of the round
        if (a1 <= a0) continue;
        const int n = (int)(a1 - jb);
        int t = (int)(a0 - jb) + g;
        float4 a = acc[k];
        for (; t + 3 * G < n; t += 4 * G) {                          // 4 independent feat loads in flight (8 were not faster: most cells are short)
          float4 f[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) f[u] = __ldg(fb + (size_t)hw_s[t + u * G] * C4 + q);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float d = dep_s[t + u * G];
            a.x = __fadd_rn(a.x, __fmul_rn(d, f[u].x)); a.y = __fadd_rn(a.y, __fmul_rn(d, f[u].y));
            a.z = __fadd_rn(a.z, __fmul_rn(d, f[u].z)); a.w = __fadd_rn(a.w, __fmul_rn(d, f[u].w));
          }
        }
        for (; t < n; t += G) {
          const float4 f0 = __ldg(fb + (size_t)hw_s[t] * C4 + q);
          const float d0 = dep_s[t];
          a.x = __fadd_rn(a.x, __fmul_rn(d0, f0.x)); a.y = __fadd_rn(a.y, __fmul_rn(d0, f0.y));
          a.z = __fadd_rn(a.z, __fmul_rn(d0, f0.z)); a.w = __fadd_rn(a.w, __fmul_rn(d0, f0.w));
        }
        acc[k] = a;
      }
    }
  }
  // the G partial sums of every cell are added in group order (deterministic); one barrier for all cells
  if (active) {
#pragma unroll
    for (int k = 0; k < kLsCells; ++k) red[((size_t)k * G + g) * C4 + q] = acc[k];
  }
  __syncthreads();
  const bool vec = (cell0 + kLsCells <= n_cells) && (n_cells % 4 == 0);
  for (int qq = tid; qq < C4; qq += kLsThreads) {
    float4 res[kLsCells];
#pragma unroll
    for (int k = 0; k < kLsCells; ++k) {
      float4 r = red[((size_t)k * G) * C4 + qq];
      for (int gg = 1; gg < G; ++gg) {
        const float4 v = red[((size_t)k * G + gg) * C4 + qq];
        r.x = __fadd_rn(r.x, v.x); r.y = __fadd_rn(r.y, v.y); r.z = __fadd_rn(r.z, v.z); r.w = __fadd_rn(r.w, v.w);
      }
      res[k] = r;
    }
    float* o = out + ((size_t)b * C + 4 * qq) * n_cells + cell0;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float v[kLsCells];
#pragma unroll
      for (int k = 0; k < kLsCells; ++k) v[k] = ch == 0 ? res[k].x : ch == 1 ? res[k].y : ch == 2 ? res[k].z : res[k].w;
      float* oc = o + (size_t)ch * n_cells;
      if (vec) {
#pragma unroll
        for (int k4 = 0; k4 < kLsCells / 4; ++k4)
          reinterpret_cast<float4*>(oc)[k4] = make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]);
      } else {
#pragma unroll
        for (int k = 0; k < kLsCells; ++k) if (cell0 + k < n_cells) oc[k] = v[k];
      }
    }
  }
}

// C % 4 != 0 or C > 2048: one thread per channel, sequential over the cell's points.
__global__ void __launch_bounds__(512)
k_lift_splat_fwd_scalar(const float* __restrict__ feat_cl, const float* __restrict__ depth, const uint32_t* __restrict__ cell_start,
                        const int32_t* __restrict__ sorted, int B, int64_t n_pts, int HW, int C, int n_cells, float* __restrict__ out) {
  const int groups = (n_cells + kLsCells - 1) / kLsCells;
  const int b = blockIdx.x / groups, cell0 = (blockIdx.x % groups) * kLsCells;
  const uint32_t* cs = cell_start + (size_t)b * (n_cells + 1);
  const int32_t* list = sorted + (size_t)b * n_pts;
  const float* fb = feat_cl + (size_t)b * HW * C;
  const float* db = depth + (size_t)b * n_pts;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    for (int k = 0; k < kLsCells; ++k) {
      const int cell = cell0 + k;
      if (cell >= n_cells) break;
      float acc = 0.f;
      for (uint32_t j = cs[cell]; j < cs[cell + 1]; ++j) {
        const int p = list[j];
        acc = __fadd_rn(acc, __fmul_rn(__ldg(db + p), __ldg(fb + (size_t)(p % HW) * C + c)));
      }
      out[((size_t)b * C + c) * n_cells + cell] = acc;
    }
  }
}

// grad_depth[b,d,hw] = kept ? sum_c gout[b,cell,c] * feat[b,hw,c] : 0 ; grad_feat[b,hw,c] = sum_{d kept} depth[b,d,hw] * gout[b,cell,c]
// gout_cl is the output gradient in [B, n_cells, C] order.  One warp per pixel; lane L owns channels L, L+32, ...
constexpr int kLsMaxC = 1024;    // channels held in registers per lane: kLsMaxC / 32
template <int CPL>
__global__ void __launch_bounds__(256)
k_lift_splat_bwd(const float* __restrict__ gout_cl, const float* __restrict__ feat_cl, const float* __restrict__ depth,
                 const int32_t* __restrict__ cell, int B, int D, int HW, int C, int n_cells, float* __restrict__ grad_depth,
                 float* __restrict__ grad_feat_cl) {
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (wid >= (int64_t)B * HW) return;
  const int b = (int)(wid / HW), hw = (int)(wid % HW);
  const float* f = feat_cl + ((size_t)b * HW + hw) * C;
  float fv[CPL], gf[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) { const int c = lane + 32 * k; fv[k] = c < C ? __ldg(f + c) : 0.f; gf[k] = 0.f; }
  const int32_t* cp = cell + (size_t)b * D * HW + hw;
  const float* dp = depth + (size_t)b * D * HW + hw;
  float* gd = grad_depth + (size_t)b * D * HW + hw;
  // Depth bins in chunks of 32: lane L fetches the cell id and the depth weight of bin d0 + L (one round trip for the whole
  // chunk instead of one per bin), then the warp walks only the KEPT bins (top-k mask: ~10 of 37), two gout rows in flight.
  for (int d0 = 0; d0 < D; d0 += 32) {
    const int dl = d0 + (int)lane;
    int cl = -1;
    float dv = 0.f;
    if (dl < D) {
      cl = __ldg(cp + (size_t)dl * HW);
      if (cl >= n_cells) cl = -1;
      if (cl >= 0) dv = __ldg(dp + (size_t)dl * HW);
    }
    float my_dot = 0.f;
    unsigned kept = __ballot_sync(0xffffffffu, cl >= 0);
    while (kept) {
      const int da = __ffs(kept) - 1;
      kept &= kept - 1;
      const int db2 = kept ? __ffs(kept) - 1 : -1;
      if (db2 >= 0) kept &= kept - 1;
      const int ca = __shfl_sync(0xffffffffu, cl, da), cb = __shfl_sync(0xffffffffu, cl, db2 >= 0 ? db2 : da);
      const float wa = __shfl_sync(0xffffffffu, dv, da), wb = __shfl_sync(0xffffffffu, dv, db2 >= 0 ? db2 : da);
      const float* ga = gout_cl + ((size_t)b * n_cells + ca) * C;
      const float* gb = gout_cl + ((size_t)b * n_cells + cb) * C;
      float va[CPL], vb[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k) { const int c = lane + 32 * k; va[k] = c < C ? __ldg(ga + c) : 0.f; }
#pragma unroll
      for (int k = 0; k < CPL; ++k) { const int c = lane + 32 * k; vb[k] = (db2 >= 0 && c < C) ? __ldg(gb + c) : 0.f; }
      float dota = 0.f, dotb = 0.f;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {                    // bin da first, then db2: ascending depth order, as before
        dota = __fadd_rn(dota, __fmul_rn(va[k], fv[k]));
        gf[k] = __fadd_rn(gf[k], __fmul_rn(wa, va[k]));
      }
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        dotb = __fadd_rn(dotb, __fmul_rn(vb[k], fv[k]));
        if (db2 >= 0) gf[k] = __fadd_rn(gf[k], __fmul_rn(wb, vb[k]));
      }
#pragma unroll
      for (int dd = 16; dd; dd >>= 1) {
        dota += __shfl_xor_sync(0xffffffffu, dota, dd);
        dotb += __shfl_xor_sync(0xffffffffu, dotb, dd);
      }
      if ((int)lane == da) my_dot = dota;
      if ((int)lane == db2) my_dot = dotb;
    }
    if (dl < D) gd[(size_t)dl * HW] = my_dot;
  }
  float* o = grad_feat_cl + ((size_t)b * HW + hw) * C;
#pragma unroll
  for (int k = 0; k < CPL; ++k) { const int c = lane + 32 * k; if (c < C) o[c] = gf[k]; }
}

template <typename T>
static int run_pool_fwd(const T* x, int64_t sb, int64_t sp, int64_t sc, const int32_t* cell, int B, int64_t n_pts, int C,
                        int n_cells, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
  BevWs w = carve_bev(ws, B, n_pts, n_cells);
  if (w.bytes > ws_bytes) return MUVO_E_WORKSPACE;
  if (pool_stream_eligible((int)sizeof(T), x, sb, sp, sc, B, n_pts, C, n_cells)) {     // (B, C, D, H, W) memory: stream it (bev_stream.cu)
    prof_mark("<bev_fwd>", st);
    return pool_stream_fwd(x, sizeof(T) == 4 ? MUVO_F32 : (std::is_same<T, __half>::value ? MUVO_F16 : MUVO_BF16), sb, sc, cell, nullptr,
                           nullptr, nullptr, nullptr, B, n_pts, C, n_cells, out, w.lists, w.steps, st);
  }
  const size_t smem = (size_t)kSortWarps * n_cells * 4;
  if (smem > 200 * 1024) return MUVO_E_SHAPE;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_cell_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_cell_place, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  }
  const int64_t n_chunks = (int64_t)B * w.n_wc;
  const unsigned sort_blocks = (unsigned)ceil_div64(n_chunks, kSortWarps);
  prof_mark("<bev_fwd>", st);
  k_cell_hist<<<sort_blocks, kSortWarps * 32, smem, st>>>(cell, B, n_pts, n_cells, w.n_wc, w.chunk_base, w.chunk_kept);
  MUVO_AFTER_LAUNCH("k_cell_hist", st);
  k_cell_scan<<<(unsigned)ceil_div64((int64_t)B * n_cells, 64), 64, 0, st>>>(w.chunk_base, w.cell_total, B, n_cells, w.n_wc);
  MUVO_AFTER_LAUNCH("k_cell_scan", st);
  k_cell_starts<<<B, 1024, 0, st>>>(w.cell_total, w.cell_start, n_cells, w.chunk_kept, w.n_wc);
  MUVO_AFTER_LAUNCH("k_cell_starts", st);
  k_cell_place<<<sort_blocks, kSortWarps * 32, smem, st>>>(cell, B, n_pts, n_cells, w.n_wc, w.chunk_base, w.cell_start, w.sorted,
                                                             w.chunk_kept, w.kp, w.dest, w.dest16);
  MUVO_AFTER_LAUNCH("k_cell_place", st);
  // (B, C, D, H, W) memory: the pipelined row kernel while its cell tables fit next to the staging buffer (<= 3072 cells), the
  // one-CTA-per-row kernel up to 9208 cells, above that (and for every other layout) the kernel that walks cells
  const size_t rsmem = (size_t)kRowCap * sizeof(float) + (size_t)n_cells * 2 + 16;
  const size_t psmem = (size_t)kRowCap * sizeof(float) + (size_t)((n_cells + 1) / 2 + 1) * 4 + (size_t)(n_cells + 1) * 4;
  constexpr size_t kSmemCap = 226 * 1024;            // 227 KiB per CTA minus the kernels' static shared memory
  const bool rows_ok = n_cells <= 65535 && rsmem <= kSmemCap;
  const bool pipe_ok = rows_ok && psmem <= kSmemCap && g_tuning[2] != 1 && n_pts % 2 == 0;
  if (sp == 1 && rows_ok) {
    if (!pipe_ok) {        // tuning key 2 = 1: the one-CTA-per-row gather kernel (also: odd row stride, grids past the pipelined kernel)
      cudaError_t e = cudaFuncSetAttribute(k_pool_rows<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
      if (e != cudaSuccess) return ::muvo::cuda_fail(e);
      k_pool_rows<T><<<(unsigned)((int64_t)B * C), kRowThreads, rsmem, st>>>(x, sb, sc, w.kp, w.dest, w.cell_start, w.sorted, B, n_pts,
                                                                           C, n_cells, out);
    } else {
      cudaError_t e = cudaFuncSetAttribute(k_pool_rows_pipe<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);
      if (e != cudaSuccess) return ::muvo::cuda_fail(e);
      int sms = kNumSMsB200;
      { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
      const int64_t rows = (int64_t)B * C;
      k_pool_rows_pipe<T><<<(unsigned)(rows < sms ? rows : sms), kPipeThreads, psmem, st>>>(x, sb, sc, w.kp, w.dest, w.dest16, w.cell_start,
                                                                                        w.sorted, B, n_pts, C, n_cells, out);
    }
  } else {
    k_pool_channel_major<T><<<(unsigned)((int64_t)B * n_cells), 128, 0, st>>>(x, sb, sp, sc, w.cell_start, w.sorted, B, n_pts,
                                                                             C, n_cells, out);
  }
  MUVO_AFTER_LAUNCH(sp == 1 && rows_ok ? "k_pool_rows" : "k_pool_channel_major", st);
  return MUVO_OK;
}

template <typename T>
static int run_pool_bwd(const float* gout, const int32_t* cell, int B, int64_t n_pts, int C, int n_cells, T* gx, int64_t sb,
                        int64_t sp, int64_t sc, cudaStream_t st) {
  prof_mark("<bev_bwd>", st);
  if (sp == 1 && sizeof(T) == 4 && n_pts % 4 == 0 && sb % 4 == 0 && sc % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(gx) & 15) == 0 && (reinterpret_cast<uintptr_t>(cell) & 15) == 0) {
    const int n_cg = (C + kBwdCh - 1) / kBwdCh;
    int64_t n = (int64_t)B * n_cg * (n_pts / 4);
    k_pool_bwd_rows_f32<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(gout, cell, B, n_pts, C, n_cells, (float*)gx, sb, sc);
    MUVO_AFTER_LAUNCH("k_pool_bwd_rows_f32", st);
    return MUVO_OK;
  }
  if (sp == 1) {
    int64_t n = (int64_t)B * C * ceil_div64(n_pts, 4);
    k_pool_bwd_point_major<T><<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(gout, cell, B, n_pts, C, n_cells, gx, sb, sc);
  } else {
    int64_t n = (int64_t)B * n_pts * C;
    k_pool_bwd_generic<T><<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(gout, cell, B, n_pts, C, n_cells, gx, sb, sp, sc);
  }
  MUVO_AFTER_LAUNCH(sp == 1 ? "k_pool_bwd_point_major" : "k_pool_bwd_generic", st);
  return MUVO_OK;
}

// cell[i] = mask[i] ? cell0[i] : -1  (frustum_pooling.py:153-156 applied to the cached, mask-independent cell ids)
__global__ void __launch_bounds__(256)
k_fold_mask(const int32_t* __restrict__ cell0, const uint8_t* __restrict__ mask, int64_t n4, int64_t n, int32_t* __restrict__ cell) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t < n4) {
    const int4 c = __ldg(reinterpret_cast<const int4*>(cell0) + t);
    const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(mask) + t);
    int4 o;
    o.x = (m & 0x000000ffu) ? c.x : -1; o.y = (m & 0x0000ff00u) ? c.y : -1;
    o.z = (m & 0x00ff0000u) ? c.z : -1; o.w = (m & 0xff000000u) ? c.w : -1;
    reinterpret_cast<int4*>(cell)[t] = o;
  }
  if (t == 0) for (int64_t i = n4 * 4; i < n; ++i) cell[i] = mask[i] ? cell0[i] : -1;
}

struct SegWs { uint32_t* block_cnt; size_t bytes; int nblocks; };
static SegWs carve_seg(void* base, int64_t n) {
  SegWs w;
  w.nblocks = (int)ceil_div64(n > 0 ? n : 1, kScanBlock);
  w.block_cnt = (uint32_t*)base;
  w.bytes = align_up((size_t)w.nblocks * 4, 256);
  return w;
}


// ---------------------------------------------------------------- N2: cached cell sort for the fused lift-splat
// The cell of every frustum point is fixed per camera rig; only the top-k depth mask changes per call.  The plan holds the
// mask-independent sort (point ids grouped by cell, cell starts, and the cell of every sorted entry); a call filters it by
// the mask with two small kernels (a stable compaction: per-chunk counts, then placement + re-based cell starts) instead
// of the four-kernel counting sort (50 us at cfg3 -> 12 us).
struct LsPlan {
  uint32_t* cell_start0 = nullptr;  // [B, n_cells + 1]
  int32_t* sorted0 = nullptr;       // [B, n_pts]   first cell_start0[b][n_cells] entries valid
  int32_t* scell0 = nullptr;        // [B, n_pts]   cell of sorted0[i]
  size_t bytes = 0;
};
static LsPlan carve_ls_plan(void* base, int B, int64_t n_pts, int n_cells) {
  LsPlan p;
  char* b = (char*)base;
  size_t o = 0;
  p.cell_start0 = (uint32_t*)(b + o); o = align_up(o + (size_t)B * (n_cells + 1) * 4, 256);
  p.sorted0 = (int32_t*)(b + o);      o = align_up(o + (size_t)B * n_pts * 4, 256);
  p.scell0 = (int32_t*)(b + o);       o = align_up(o + (size_t)B * n_pts * 4, 256);
  p.bytes = o;
  return p;
}

__global__ void __launch_bounds__(128)
k_ls_plan_cells(const uint32_t* __restrict__ cell_start0, int B, int64_t n_pts, int n_cells, int32_t* __restrict__ scell0) {
  const int64_t t = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (t >= (int64_t)B * n_cells) return;
  const int b = (int)(t / n_cells), c = (int)(t % n_cells);
  const uint32_t* cs = cell_start0 + (size_t)b * (n_cells + 1);
  int32_t* o = scell0 + (size_t)b * n_pts;
  for (uint32_t i = cs[c]; i < cs[c + 1]; ++i) o[i] = c;
}

constexpr int kLfThreads = 256;
constexpr int kLfPer = 8;
constexpr int kLfChunk = kLfThreads * kLfPer;      // sorted entries per CTA

__global__ void __launch_bounds__(kLfThreads)
k_ls_filter_count(const int32_t* __restrict__ sorted0, const uint32_t* __restrict__ cell_start0, const uint8_t* __restrict__ mask,
                  int64_t n_pts, int n_cells, int n_fc, uint32_t* __restrict__ counts) {
  __shared__ uint32_t ws[kLfThreads / 32];
  const int b = blockIdx.y, k = blockIdx.x;
  const uint32_t n0 = cell_start0[(size_t)b * (n_cells + 1) + n_cells];
  const int32_t* so = sorted0 + (size_t)b * n_pts;
  const uint8_t* mb = mask + (size_t)b * n_pts;
  uint32_t cnt = 0;
  const uint32_t i0 = (uint32_t)k * kLfChunk + threadIdx.x * kLfPer;
#pragma unroll
  for (int e = 0; e < kLfPer; ++e)
    if (i0 + e < n0) cnt += __ldg(mb + __ldg(so + i0 + e)) != 0 ? 1u : 0u;
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kLfThreads / 32; ++w) t += ws[w];
    counts[(size_t)b * n_fc + k] = t;
  }
}

__global__ void __launch_bounds__(kLfThreads)
k_ls_filter_write(const int32_t* __restrict__ sorted0, const int32_t* __restrict__ scell0, const uint32_t* __restrict__ cell_start0,
                  const uint8_t* __restrict__ mask, int64_t n_pts, int n_cells, int n_fc, const uint32_t* __restrict__ counts,
                  int32_t* __restrict__ sorted, uint32_t* __restrict__ cell_start) {
  __shared__ uint32_t ws[kLfThreads / 32];
  __shared__ uint32_t prefix_s;
  const int b = blockIdx.y, k = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n0 = cell_start0[(size_t)b * (n_cells + 1) + n_cells];
  uint32_t* cs = cell_start + (size_t)b * (n_cells + 1);
  if (n0 == 0u) {                                                // nothing in bounds in this frame
    if (k == 0) for (int c = tid; c <= n_cells; c += kLfThreads) cs[c] = 0u;
    return;
  }
  if ((uint32_t)k * kLfChunk >= n0) return;
  // kept entries in the chunks before this one
  uint32_t pre = 0;
  for (int j = tid; j < k; j += kLfThreads) pre += counts[(size_t)b * n_fc + j];
  pre = __reduce_add_sync(0xffffffffu, pre);
  if (lane == 0) ws[warp] = pre;
  __syncthreads();
  if (tid == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kLfThreads / 32; ++w) t += ws[w];
    prefix_s = t;
  }
  __syncthreads();
  const uint32_t prefix = prefix_s;
  const int32_t* so = sorted0 + (size_t)b * n_pts;
  const int32_t* sc = scell0 + (size_t)b * n_pts;
  const uint8_t* mb = mask + (size_t)b * n_pts;
  int32_t* out = sorted + (size_t)b * n_pts;
  const uint32_t i0 = (uint32_t)k * kLfChunk + tid * kLfPer;
  int32_t pid[kLfPer];
  uint32_t keep = 0, cnt = 0;
#pragma unroll
  for (int e = 0; e < kLfPer; ++e) {
    pid[e] = 0;
    if (i0 + e < n0) {
      pid[e] = __ldg(so + i0 + e);
      if (__ldg(mb + pid[e]) != 0) { keep |= 1u << e; ++cnt; }
    }
  }
  // block-wide exclusive scan of the per-thread counts
  uint32_t incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  __syncthreads();                                               // ws is reused
  if (lane == 31) ws[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int w = 0; w < kLfThreads / 32; ++w) if (w < warp) wbase += ws[w];
  uint32_t ex = prefix + wbase + (incl - cnt);                   // new index of this thread's first kept entry
  int32_t prev = i0 > 0u && i0 < n0 ? __ldg(sc + i0 - 1) : -1;
#pragma unroll
  for (int e = 0; e < kLfPer; ++e) {
    const uint32_t g = i0 + e;
    if (g >= n0) break;
    const int32_t cur = __ldg(sc + g);
    if (cur != prev) for (int32_t c = prev + 1; c <= cur; ++c) cs[c] = ex;      // cells (prev, cur] start here (empty ones included)
    prev = cur;
    if ((keep >> e) & 1u) out[ex++] = pid[e];
    if (g == n0 - 1u) for (int32_t c = cur + 1; c <= n_cells; ++c) cs[c] = ex;  // the cells behind the last entry, and the total
  }
}

static int run_cell_sort(const int32_t* cell, int B, int64_t n_pts, int n_cells, const BevWs& w, cudaStream_t st) {
  const size_t smem = (size_t)kSortWarps * n_cells * 4;
  if (smem > 200 * 1024) return MUVO_E_SHAPE;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_cell_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_cell_place, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  }
  const int64_t n_chunks = (int64_t)B * w.n_wc;
  const unsigned sort_blocks = (unsigned)ceil_div64(n_chunks, kSortWarps);
  k_cell_hist<<<sort_blocks, kSortWarps * 32, smem, st>>>(cell, B, n_pts, n_cells, w.n_wc, w.chunk_base, w.chunk_kept);
  MUVO_AFTER_LAUNCH("k_cell_hist", st);
  k_cell_scan<<<(unsigned)ceil_div64((int64_t)B * n_cells, 64), 64, 0, st>>>(w.chunk_base, w.cell_total, B, n_cells, w.n_wc);
  MUVO_AFTER_LAUNCH("k_cell_scan", st);
  k_cell_starts<<<B, 1024, 0, st>>>(w.cell_total, w.cell_start, n_cells, w.chunk_kept, w.n_wc);
  MUVO_AFTER_LAUNCH("k_cell_starts", st);
  k_cell_place<<<sort_blocks, kSortWarps * 32, smem, st>>>(cell, B, n_pts, n_cells, w.n_wc, w.chunk_base, w.cell_start, w.sorted,
                                                             w.chunk_kept, w.kp, w.dest, w.dest16);
  MUVO_AFTER_LAUNCH("k_cell_place", st);
  return MUVO_OK;
}

static int launch_lift_fwd(const float* feat_cl, const float* depth, const uint32_t* cell_start, const int32_t* sorted, int B, int64_t n_pts,
                           int HW, int C, int n_cells, float* out, cudaStream_t st) {
  const int groups = (n_cells + kLsCells - 1) / kLsCells;
  if (C % 4 == 0 && C / 4 <= kLsThreads && (reinterpret_cast<uintptr_t>(feat_cl) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int C4 = C / 4, G = kLsThreads / C4;
    const size_t lsmem = (size_t)kLsStage * 8 + (size_t)kLsCells * G * C4 * 16;
    cudaError_t e = cudaFuncSetAttribute(k_lift_splat_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem);
    if (e != cudaSuccess) return ::muvo::cuda_fail(e);
    k_lift_splat_fwd<<<(unsigned)((int64_t)B * groups), kLsThreads, lsmem, st>>>(feat_cl, depth, cell_start, sorted, B, n_pts, HW, C, n_cells, out);
  } else {
    int threads = ((C + 31) / 32) * 32;
    if (threads > 512) threads = 512;
    k_lift_splat_fwd_scalar<<<(unsigned)((int64_t)B * groups), threads, 0, st>>>(feat_cl, depth, cell_start, sorted, B, n_pts, HW, C, n_cells, out);
  }
  MUVO_AFTER_LAUNCH("k_lift_splat_fwd", st);
  return MUVO_OK;
}

}  // namespace
}  // namespace muvo

using namespace muvo;

extern "C" {

int muvo_bev_pool_workspace_bytes(int32_t B, int64_t n_pts, int32_t n_cells, size_t* bytes_out_h) {
  if (!bytes_out_h) return MUVO_E_NULL;
  if (B < 0 || n_pts < 0 || n_cells <= 0) return MUVO_E_ARG;
  *bytes_out_h = carve_bev(nullptr, B, n_pts, n_cells).bytes + 256;
  return MUVO_OK;
}

int muvo_bev_pool_max_cells(void) { return (200 * 1024) / (kSortWarps * 4); }

int muvo_bev_fold_mask(const int32_t* cell0, const uint8_t* mask, int64_t n, int32_t* cell_out, void* stream) {
  if (n < 0) return MUVO_E_ARG;
  if (n == 0) return MUVO_OK;
  if (!cell0 || !mask || !cell_out) return MUVO_E_NULL;
  const bool vec = ((reinterpret_cast<uintptr_t>(cell0) | reinterpret_cast<uintptr_t>(cell_out)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(mask) & 3) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  k_fold_mask<<<(unsigned)ceil_div64(n4 > 0 ? n4 : 1, 256), 256, 0, (cudaStream_t)stream>>>(cell0, mask, n4, n, cell_out);
  MUVO_AFTER_LAUNCH("k_fold_mask", (cudaStream_t)stream);
  return MUVO_OK;
}

int muvo_bev_pool_fwd(const void* x, int32_t x_dtype, int64_t x_stride_b, int64_t x_stride_p, int64_t x_stride_c,
                      const int32_t* cell, int32_t B, int64_t n_pts, int32_t C, int32_t n_cells, float* out, void* ws,
                      size_t ws_bytes, void* stream) {
  if (B < 0 || n_pts < 0 || C < 0 || n_cells <= 0) return MUVO_E_ARG;
  if (B == 0 || C == 0) return MUVO_OK;
  if (!out || !ws) return MUVO_E_NULL;
  if (n_pts > 0 && (!x || !cell)) return MUVO_E_NULL;
  if (n_pts >= ((int64_t)1 << 31) || (int64_t)B * n_cells >= ((int64_t)1 << 31)) return MUVO_E_SHAPE;
  if (reinterpret_cast<uintptr_t>(ws) & 255) return MUVO_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  switch (x_dtype) {
    case MUVO_F32:  return run_pool_fwd<float>((const float*)x, x_stride_b, x_stride_p, x_stride_c, cell, B, n_pts, C, n_cells, out, ws, ws_bytes, st);
    case MUVO_F16:  return run_pool_fwd<__half>((const __half*)x, x_stride_b, x_stride_p, x_stride_c, cell, B, n_pts, C, n_cells, out, ws, ws_bytes, st);
    case MUVO_BF16: return run_pool_fwd<__nv_bfloat16>((const __nv_bfloat16*)x, x_stride_b, x_stride_p, x_stride_c, cell, B, n_pts, C, n_cells, out, ws, ws_bytes, st);
    default: return MUVO_E_ARG;
  }
}

int muvo_bev_plan_bytes(int32_t B, int64_t n_pts, size_t* bytes_out_h) {
  if (!bytes_out_h) return MUVO_E_NULL;
  if (B < 0 || n_pts < 0) return MUVO_E_ARG;
  *bytes_out_h = align_up(stream_lists_bytes(B, n_pts), 256) + align_up(stream_steps_bytes(B, n_pts), 256) + 256;
  return MUVO_OK;
}

int muvo_bev_plan_build(const int32_t* cell0, int32_t B, int64_t n_pts, int32_t n_cells, void* plan, size_t plan_bytes, void* stream) {
  if (B < 0 || n_pts < 0 || n_cells <= 0) return MUVO_E_ARG;
  if (B == 0 || n_pts == 0) return MUVO_OK;
  if (!cell0 || !plan) return MUVO_E_NULL;
  if (reinterpret_cast<uintptr_t>(plan) & 255) return MUVO_E_ALIGN;
  if (n_cells >= (1 << (32 - kStreamPosBits)) - 1) return MUVO_E_SHAPE;
  size_t need = 0;
  muvo_bev_plan_bytes(B, n_pts, &need);
  if (plan_bytes < need - 256) return MUVO_E_WORKSPACE;
  uint32_t* keys = (uint32_t*)plan;
  uint32_t* cnt = (uint32_t*)((char*)plan + align_up(stream_lists_bytes(B, n_pts), 256));
  return pool_stream_plan(cell0, B, n_pts, n_cells, keys, cnt, (cudaStream_t)stream);
}

int muvo_bev_pool_fwd_masked(const void* x, int32_t x_dtype, int64_t x_stride_b, int64_t x_stride_p, int64_t x_stride_c,
                             const int32_t* cell0, const uint8_t* mask, int32_t* cell_out, const void* plan, int32_t B,
                             int64_t n_pts, int32_t C, int32_t n_cells, float* out, void* ws, size_t ws_bytes, void* stream) {
  if (B < 0 || n_pts < 0 || C < 0 || n_cells <= 0) return MUVO_E_ARG;
  if (B == 0 || C == 0) return MUVO_OK;
  if (!out || !ws || !cell_out) return MUVO_E_NULL;
  if (n_pts > 0 && (!x || !cell0)) return MUVO_E_NULL;
  if (x_dtype != MUVO_F32 && x_dtype != MUVO_F16 && x_dtype != MUVO_BF16) return MUVO_E_ARG;
  if (n_pts >= ((int64_t)1 << 31) || (int64_t)B * n_cells >= ((int64_t)1 << 31)) return MUVO_E_SHAPE;
  if (reinterpret_cast<uintptr_t>(ws) & 255) return MUVO_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const int eb = x_dtype == MUVO_F32 ? 4 : 2;
  if (pool_stream_eligible(eb, x, x_stride_b, x_stride_p, x_stride_c, B, n_pts, C, n_cells)) {
    BevWs w = carve_bev(ws, B, n_pts, n_cells);
    if (w.bytes > ws_bytes) return MUVO_E_WORKSPACE;
    prof_mark("<bev_fwd>", st);
    const uint32_t* keys = (const uint32_t*)plan;
    const uint32_t* cnt = plan ? (const uint32_t*)((const char*)plan + align_up(stream_lists_bytes(B, n_pts), 256)) : nullptr;
    return pool_stream_fwd(x, x_dtype, x_stride_b, x_stride_c, cell0, mask, cell_out, keys, cnt, B, n_pts, C, n_cells, out, w.lists,
                           w.steps, st);
  }
  // gather path: fold the mask first (frustum_pooling.py:153-156), then the index sort + row kernels
  const int64_t n = (int64_t)B * n_pts;
  if (mask) {
    int rc = muvo_bev_fold_mask(cell0, mask, n, cell_out, stream);
    if (rc != MUVO_OK) return rc;
  } else if (n > 0) {
    cudaError_t e = cudaMemcpyAsync(cell_out, cell0, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  }
  return muvo_bev_pool_fwd(x, x_dtype, x_stride_b, x_stride_p, x_stride_c, cell_out, B, n_pts, C, n_cells, out, ws, ws_bytes, stream);
}

int muvo_bev_pool_bwd(const float* grad_out, const int32_t* cell, int32_t B, int64_t n_pts, int32_t C, int32_t n_cells,
                      void* grad_x, int32_t gx_dtype, int64_t gx_stride_b, int64_t gx_stride_p, int64_t gx_stride_c,
                      void* stream) {
  if (B < 0 || n_pts < 0 || C < 0 || n_cells <= 0) return MUVO_E_ARG;
  if (B == 0 || C == 0 || n_pts == 0) return MUVO_OK;
  if (!grad_out || !cell || !grad_x) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  switch (gx_dtype) {
    case MUVO_F32:  return run_pool_bwd<float>(grad_out, cell, B, n_pts, C, n_cells, (float*)grad_x, gx_stride_b, gx_stride_p, gx_stride_c, st);
    case MUVO_F16:  return run_pool_bwd<__half>(grad_out, cell, B, n_pts, C, n_cells, (__half*)grad_x, gx_stride_b, gx_stride_p, gx_stride_c, st);
    case MUVO_BF16: return run_pool_bwd<__nv_bfloat16>(grad_out, cell, B, n_pts, C, n_cells, (__nv_bfloat16*)grad_x, gx_stride_b, gx_stride_p, gx_stride_c, st);
    default: return MUVO_E_ARG;
  }
}

int muvo_bev_pool_is_streamed(int32_t elem_bytes, const void* x, int64_t x_stride_b, int64_t x_stride_p, int64_t x_stride_c, int32_t B,
                              int64_t n_pts, int32_t C, int32_t n_cells) {
  return pool_stream_eligible(elem_bytes, x, x_stride_b, x_stride_p, x_stride_c, B, n_pts, C, n_cells) ? 1 : 0;
}

int muvo_bev_pool_bwd_streamed(const float* grad_out, const int32_t* cell, int32_t B, int64_t n_pts, int32_t C, int32_t n_cells,
                               void* grad_x, int32_t gx_dtype, int64_t gx_stride_b, int64_t gx_stride_p, int64_t gx_stride_c,
                               const void* fwd_ws, size_t fwd_ws_bytes, void* stream) {
  if (B < 0 || n_pts < 0 || C < 0 || n_cells <= 0) return MUVO_E_ARG;
  if (B == 0 || C == 0 || n_pts == 0) return MUVO_OK;
  if (!grad_out || !cell || !grad_x) return MUVO_E_NULL;
  const int eb = gx_dtype == MUVO_F32 ? 4 : 2;
  if (fwd_ws && (gx_dtype == MUVO_F32 || gx_dtype == MUVO_F16 || gx_dtype == MUVO_BF16) &&
      pool_bwd_stream_eligible(eb, grad_x, gx_stride_b, gx_stride_p, gx_stride_c, B, n_pts, C, n_cells)) {
    BevWs w = carve_bev(const_cast<void*>(fwd_ws), B, n_pts, n_cells);
    if (w.bytes > fwd_ws_bytes) return MUVO_E_WORKSPACE;
    prof_mark("<bev_bwd>", (cudaStream_t)stream);
    return pool_stream_bwd(grad_out, w.lists, w.steps, B, n_pts, C, n_cells, grad_x, gx_dtype, gx_stride_b, gx_stride_c, (cudaStream_t)stream);
  }
  return muvo_bev_pool_bwd(grad_out, cell, B, n_pts, C, n_cells, grad_x, gx_dtype, gx_stride_b, gx_stride_p, gx_stride_c, stream);
}

int muvo_lift_splat_fwd(const float* feat_cl, const float* depth, const int32_t* cell, int32_t B, int32_t D, int32_t HW,
                        int32_t C, int32_t n_cells, float* out, void* ws, size_t ws_bytes, void* stream) {
  if (B < 0 || D <= 0 || HW <= 0 || C < 0 || n_cells <= 0) return MUVO_E_ARG;
  if (B == 0 || C == 0) return MUVO_OK;
  if (!feat_cl || !depth || !cell || !out || !ws) return MUVO_E_NULL;
  const int64_t n_pts = (int64_t)D * HW;
  if (n_pts >= ((int64_t)1 << 31) || (int64_t)B * n_cells >= ((int64_t)1 << 31)) return MUVO_E_SHAPE;
  if (reinterpret_cast<uintptr_t>(ws) & 255) return MUVO_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  BevWs w = carve_bev(ws, B, n_pts, n_cells);
  if (w.bytes > ws_bytes) return MUVO_E_WORKSPACE;
  prof_mark("<lift_splat_fwd>", st);
  const int rc = run_cell_sort(cell, B, n_pts, n_cells, w, st);
  if (rc != MUVO_OK) return rc;
  return launch_lift_fwd(feat_cl, depth, w.cell_start, w.sorted, B, n_pts, HW, C, n_cells, out, st);
}

int muvo_lift_splat_plan_bytes(int32_t B, int64_t n_pts, int32_t n_cells, size_t* bytes_out_h) {
  if (!bytes_out_h) return MUVO_E_NULL;
  if (B < 0 || n_pts < 0 || n_cells <= 0) return MUVO_E_ARG;
  *bytes_out_h = carve_ls_plan(nullptr, B, n_pts, n_cells).bytes + 256;
  return MUVO_OK;
}

int muvo_lift_splat_plan_build(const int32_t* cell0, int32_t B, int64_t n_pts, int32_t n_cells, void* plan, size_t plan_bytes, void* ws,
                               size_t ws_bytes, void* stream) {
  if (B < 0 || n_pts < 0 || n_cells <= 0) return MUVO_E_ARG;
  if (B == 0) return MUVO_OK;
  if (!cell0 || !plan || !ws) return MUVO_E_NULL;
  if (n_pts >= ((int64_t)1 << 31) || (int64_t)B * n_cells >= ((int64_t)1 << 31)) return MUVO_E_SHAPE;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) || (reinterpret_cast<uintptr_t>(plan) & 255)) return MUVO_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  BevWs w = carve_bev(ws, B, n_pts, n_cells);
  LsPlan pl = carve_ls_plan(plan, B, n_pts, n_cells);
  if (w.bytes > ws_bytes || pl.bytes > plan_bytes) return MUVO_E_WORKSPACE;
  int rc = run_cell_sort(cell0, B, n_pts, n_cells, w, st);
  if (rc != MUVO_OK) return rc;
  cudaError_t e = cudaMemcpyAsync(pl.cell_start0, w.cell_start, (size_t)B * (n_cells + 1) * 4, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(pl.sorted0, w.sorted, (size_t)B * n_pts * 4, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  k_ls_plan_cells<<<(unsigned)ceil_div64((int64_t)B * n_cells, 128), 128, 0, st>>>(pl.cell_start0, B, n_pts, n_cells, pl.scell0);
  MUVO_AFTER_LAUNCH("k_ls_plan_cells", st);
  return MUVO_OK;
}

int muvo_lift_splat_fwd_planned(const float* feat_cl, const float* depth, const void* plan, size_t plan_bytes, const uint8_t* mask,
                                int32_t B, int32_t D, int32_t HW, int32_t C, int32_t n_cells, float* out, void* ws, size_t ws_bytes,
                                void* stream) {
  if (B < 0 || D <= 0 || HW <= 0 || C < 0 || n_cells <= 0) return MUVO_E_ARG;
  if (B == 0 || C == 0) return MUVO_OK;
  if (!feat_cl || !depth || !plan || !out || !ws) return MUVO_E_NULL;
  const int64_t n_pts = (int64_t)D * HW;
  if (n_pts >= ((int64_t)1 << 31) || (int64_t)B * n_cells >= ((int64_t)1 << 31)) return MUVO_E_SHAPE;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) || (reinterpret_cast<uintptr_t>(plan) & 255)) return MUVO_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  BevWs w = carve_bev(ws, B, n_pts, n_cells);
  LsPlan pl = carve_ls_plan(const_cast<void*>(plan), B, n_pts, n_cells);
  if (w.bytes > ws_bytes || pl.bytes > plan_bytes) return MUVO_E_WORKSPACE;
  prof_mark("<lift_splat_fwd>", st);
  if (!mask)                                                     // nothing to drop: the plan IS the sort
    return launch_lift_fwd(feat_cl, depth, pl.cell_start0, pl.sorted0, B, n_pts, HW, C, n_cells, out, st);
  const int n_fc = (int)ceil_div64(n_pts, kLfChunk);
  k_ls_filter_count<<<dim3((unsigned)n_fc, (unsigned)B), kLfThreads, 0, st>>>(pl.sorted0, pl.cell_start0, mask, n_pts, n_cells, n_fc, w.chunk_kept);
  MUVO_AFTER_LAUNCH("k_ls_filter_count", st);
  k_ls_filter_write<<<dim3((unsigned)n_fc, (unsigned)B), kLfThreads, 0, st>>>(pl.sorted0, pl.scell0, pl.cell_start0, mask, n_pts, n_cells, n_fc,
                                                                              w.chunk_kept, w.sorted, w.cell_start);
  MUVO_AFTER_LAUNCH("k_ls_filter_write", st);
  return launch_lift_fwd(feat_cl, depth, w.cell_start, w.sorted, B, n_pts, HW, C, n_cells, out, st);
}

int muvo_lift_splat_bwd(const float* gout_cl, const float* feat_cl, const float* depth, const int32_t* cell, int32_t B,
                        int32_t D, int32_t HW, int32_t C, int32_t n_cells, float* grad_depth, float* grad_feat_cl,
                        void* stream) {
  if (B < 0 || D <= 0 || HW <= 0 || C < 0 || n_cells <= 0) return MUVO_E_ARG;
  if (C > kLsMaxC) return MUVO_E_SHAPE;
  if (B == 0) return MUVO_OK;
  if (!gout_cl || !feat_cl || !depth || !cell || !grad_depth || !grad_feat_cl) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t warps = (int64_t)B * HW;
  const unsigned blocks = (unsigned)ceil_div64(warps * 32, 256);
  prof_mark("<lift_splat_bwd>", st);
  const int cpl = (C + 31) / 32;
  if (cpl <= 2)       k_lift_splat_bwd<2><<<blocks, 256, 0, st>>>(gout_cl, feat_cl, depth, cell, B, D, HW, C, n_cells, grad_depth, grad_feat_cl);
  else if (cpl <= 4)  k_lift_splat_bwd<4><<<blocks, 256, 0, st>>>(gout_cl, feat_cl, depth, cell, B, D, HW, C, n_cells, grad_depth, grad_feat_cl);
  else if (cpl <= 8)  k_lift_splat_bwd<8><<<blocks, 256, 0, st>>>(gout_cl, feat_cl, depth, cell, B, D, HW, C, n_cells, grad_depth, grad_feat_cl);
  else if (cpl <= 12) k_lift_splat_bwd<12><<<blocks, 256, 0, st>>>(gout_cl, feat_cl, depth, cell, B, D, HW, C, n_cells, grad_depth, grad_feat_cl);
  else if (cpl <= 16) k_lift_splat_bwd<16><<<blocks, 256, 0, st>>>(gout_cl, feat_cl, depth, cell, B, D, HW, C, n_cells, grad_depth, grad_feat_cl);
  else                k_lift_splat_bwd<32><<<blocks, 256, 0, st>>>(gout_cl, feat_cl, depth, cell, B, D, HW, C, n_cells, grad_depth, grad_feat_cl);
  MUVO_AFTER_LAUNCH("k_lift_splat_bwd", st);
  return MUVO_OK;
}

int muvo_segment_sum_workspace_bytes(int64_t n, size_t* bytes_out_h) {
  if (!bytes_out_h) return MUVO_E_NULL;
  if (n < 0) return MUVO_E_ARG;
  *bytes_out_h = carve_seg(nullptr, n).bytes + 256;
  return MUVO_OK;
}

int muvo_segment_sum_fwd(const float* x, const int64_t* ranks, int64_t n, int32_t C, int32_t* seg_id_out,
                         int32_t* n_seg_out, float* x_seg_out, int64_t* last_row_out, void* ws, size_t ws_bytes,
                         void* stream) {
  if (n < 0 || C < 0) return MUVO_E_ARG;
  if (!n_seg_out) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) { cudaError_t e = cudaMemsetAsync(n_seg_out, 0, 4, st); return e == cudaSuccess ? MUVO_OK : (int)e; }
  if (!ranks || !seg_id_out || !last_row_out || !ws || (C > 0 && (!x || !x_seg_out))) return MUVO_E_NULL;
  if (n >= ((int64_t)1 << 31)) return MUVO_E_SHAPE;
  SegWs w = carve_seg(ws, n);
  if (w.bytes > ws_bytes) return MUVO_E_WORKSPACE;
  prof_mark("<seg>", st);
  k_seg_block_count<<<w.nblocks, kScanBlock, 0, st>>>(ranks, n, w.block_cnt);
  MUVO_AFTER_LAUNCH("k_seg_block_count", st);
  k_seg_scan_blocks<<<1, 1024, 0, st>>>(w.block_cnt, w.nblocks, n_seg_out);
  MUVO_AFTER_LAUNCH("k_seg_scan_blocks", st);
  k_seg_assign<<<w.nblocks, kScanBlock, 0, st>>>(ranks, n, w.block_cnt, seg_id_out, last_row_out);
  MUVO_AFTER_LAUNCH("k_seg_assign", st);
  if (C > 0) {
    int64_t grid = n < (int64_t)kNumSMsB200 * 16 ? n : (int64_t)kNumSMsB200 * 16;
    k_seg_sum<<<(unsigned)grid, 256, 0, st>>>(x, last_row_out, n_seg_out, C, x_seg_out);
    MUVO_AFTER_LAUNCH("k_seg_sum", st);
  }
  return MUVO_OK;
}

int muvo_segment_sum_bwd(const float* grad_seg, const int32_t* seg_id, int64_t n, int32_t C, float* grad_x, void* stream) {
  if (n < 0 || C < 0) return MUVO_E_ARG;
  if (n == 0 || C == 0) return MUVO_OK;
  if (!grad_seg || !seg_id || !grad_x) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  k_seg_bwd<<<(unsigned)ceil_div64(n * C, 256), 256, 0, st>>>(grad_seg, seg_id, n, C, grad_x);
  MUVO_AFTER_LAUNCH("k_seg_bwd", st);
  return MUVO_OK;
}

}  // extern "C"

// N1, second half: the label pyramids PreProcess builds behind stages (a)/(b)
// (muvo/models/preprocess.py:151-186): range view / LIDAR_RE.SCALE, then nearest-neighbour halvings of the range view,
// its segmentation and the voxel grid (factor 2, then factor 2 of that).  One launch per tensor family; every level is
// written from level 1 directly (nearest of nearest = one composed index), so level 1 is read once.
// Index rule = PyTorch's `nearest` (torchvision resize NEAREST / F.interpolate): src = min(floor(dst * (float)in / out), in - 1).
#include "common.cuh"

namespace muvo {
namespace {

__device__ __forceinline__ int nearest_src(int dst, int in, int out) {
  const float scale = (float)in / (float)out;
  const int s = (int)floorf((float)dst * scale);
  return s < in - 1 ? s : in - 1;
}

// range view: xyzd [F,4,H,W] f32 -> l1 = x / scale [F,4,H,W], l2 [F,4,H2,W2], l4 [F,4,H4,W4]; sem [F,H,W] u8 -> s2, s4
__global__ void __launch_bounds__(256)
k_range_pyramid(const float* __restrict__ xyzd, const uint8_t* __restrict__ sem, int F, int H, int W, float scale,
                float* __restrict__ l1, float* __restrict__ l2, float* __restrict__ l4, uint8_t* __restrict__ s2,
                uint8_t* __restrict__ s4) {
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
  const int64_t n1 = (int64_t)F * 4 * H * W, n2 = (int64_t)F * 4 * H2 * W2, n4 = (int64_t)F * 4 * H4 * W4;
  const int64_t m2 = (int64_t)F * H2 * W2, m4 = (int64_t)F * H4 * W4;
  const int64_t total = n1 + n2 + n4 + (sem ? m2 + m4 : 0);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    if (t < n1) { l1[t] = __fdiv_rn(__ldg(xyzd + t), scale); continue; }                 // :152
    int64_t u = t - n1;
    if (u < n2) {                                                                         // :155-162, factor 2
      const int w = (int)(u % W2), h = (int)((u / W2) % H2); const int64_t fc = u / ((int64_t)W2 * H2);
      l2[u] = __fdiv_rn(__ldg(xyzd + (fc * H + nearest_src(h, H, H2)) * W + nearest_src(w, W, W2)), scale);
      continue;
    }
    u -= n2;
    if (u < n4) {                                                                         // factor 4 = nearest of the factor-2 level
      const int w = (int)(u % W4), h = (int)((u / W4) % H4); const int64_t fc = u / ((int64_t)W4 * H4);
      const int h1 = nearest_src(nearest_src(h, H2, H4), H, H2), w1 = nearest_src(nearest_src(w, W2, W4), W, W2);
      l4[u] = __fdiv_rn(__ldg(xyzd + (fc * H + h1) * W + w1), scale);
      continue;
    }
    u -= n4;
    if (u < m2) {                                                                         // :164-174
      const int w = (int)(u % W2), h = (int)((u / W2) % H2); const int64_t f = u / ((int64_t)W2 * H2);
      s2[u] = __ldg(sem + (f * H + nearest_src(h, H, H2)) * W + nearest_src(w, W, W2));
      continue;
    }
    u -= m2;
    {
      const int w = (int)(u % W4), h = (int)((u / W4) % H4); const int64_t f = u / ((int64_t)W4 * H4);
      const int h1 = nearest_src(nearest_src(h, H2, H4), H, H2), w1 = nearest_src(nearest_src(w, W2, W4), W, W2);
      s4[u] = __ldg(sem + (f * H + h1) * W + w1);
    }
  }
}

// voxel [F,X,Y,Z] u8 -> v2 [F,X/2,Y/2,Z/2], v4 [F,X/4,Y/4,Z/4]   (:176-186)
__global__ void __launch_bounds__(256)
k_voxel_pyramid(const uint8_t* __restrict__ vox, int F, int X, int Y, int Z, uint8_t* __restrict__ v2, uint8_t* __restrict__ v4) {
  const int X2 = X / 2, Y2 = Y / 2, Z2 = Z / 2, X4 = X / 4, Y4 = Y / 4, Z4 = Z / 4;
  const int64_t n2 = (int64_t)F * X2 * Y2 * Z2, n4 = (int64_t)F * X4 * Y4 * Z4;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2 + n4; t += (int64_t)gridDim.x * blockDim.x) {
    if (t < n2) {
      const int z = (int)(t % Z2), y = (int)((t / Z2) % Y2), x = (int)((t / ((int64_t)Z2 * Y2)) % X2);
      const int64_t f = t / ((int64_t)Z2 * Y2 * X2);
      v2[t] = __ldg(vox + ((f * X + nearest_src(x, X, X2)) * Y + nearest_src(y, Y, Y2)) * Z + nearest_src(z, Z, Z2));
    } else {
      const int64_t u = t - n2;
      const int z = (int)(u % Z4), y = (int)((u / Z4) % Y4), x = (int)((u / ((int64_t)Z4 * Y4)) % X4);
      const int64_t f = u / ((int64_t)Z4 * Y4 * X4);
      const int x1 = nearest_src(nearest_src(x, X2, X4), X, X2), y1 = nearest_src(nearest_src(y, Y2, Y4), Y, Y2),
                z1 = nearest_src(nearest_src(z, Z2, Z4), Z, Z2);
      v4[u] = __ldg(vox + ((f * X + x1) * Y + y1) * Z + z1);
    }
  }
}

// ---------------------------------------------------------------- vectorised paths (every size a multiple of 4)
// in % 2 == 0 makes the factor-2 index rule exact (scale = 2.0f: src = 2 * dst), so the levels are strided picks of level 1
// and every load / store can be 8-16 bytes wide with 32-bit index arithmetic.  (The generic kernels above decode a 64-bit
// linear index per ELEMENT: ncu round 2, 48 / 71 us for 24 frames = 5-7 % of the DRAM rate.)
// Thread = 16 consecutive w of one row of one plane: level 1 (4 x 16 B), the 8 even columns on even rows, the 4 columns
// w % 4 == 0 on rows h % 4 == 0.  The tail of the grid handles the segmentation rows (h even only: odd rows feed nothing).
__global__ void __launch_bounds__(256)
k_range_pyramid_vec(const float* __restrict__ xyzd, const uint8_t* __restrict__ sem, int planes, int F, int H, int W, float scale,
                    float* __restrict__ l1, float* __restrict__ l2, float* __restrict__ l4, uint8_t* __restrict__ s2,
                    uint8_t* __restrict__ s4) {
  const int W16 = W >> 4;
  const uint32_t n_main = (uint32_t)planes * H * W16;
  uint32_t t = blockIdx.x * 256u + threadIdx.x;
  if (t < n_main) {
    const uint32_t wq = t % W16, row = t / W16;               // row = plane * H + h
    const uint32_t h = row % H, pl = row / H;
    const float4* src = reinterpret_cast<const float4*>(xyzd + (size_t)row * W) + wq * 4;
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ld_stream_f4(src + k);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k].x = __fdiv_rn(v[k].x, scale); v[k].y = __fdiv_rn(v[k].y, scale); v[k].z = __fdiv_rn(v[k].z, scale); v[k].w = __fdiv_rn(v[k].w, scale);
    }
    float4* d1 = reinterpret_cast<float4*>(l1 + (size_t)row * W) + wq * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) d1[k] = v[k];
    if ((h & 1u) == 0u) {
      float4* d2 = reinterpret_cast<float4*>(l2 + ((size_t)pl * (H >> 1) + (h >> 1)) * (W >> 1)) + wq * 2;
      d2[0] = make_float4(v[0].x, v[0].z, v[1].x, v[1].z);
      d2[1] = make_float4(v[2].x, v[2].z, v[3].x, v[3].z);
      if ((h & 3u) == 0u)
        reinterpret_cast<float4*>(l4 + ((size_t)pl * (H >> 2) + (h >> 2)) * (W >> 2))[wq] = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
    }
    return;
  }
  t -= n_main;
  if (!sem || t >= (uint32_t)F * (H >> 1) * W16) return;
  const uint32_t wq = t % W16, r2 = t / W16;                  // r2 = f * (H / 2) + h / 2
  const uint32_t h2 = r2 % (H >> 1), f = r2 / (H >> 1);
  const uint4 b = ld_stream_u4(reinterpret_cast<const uint4*>(sem + ((size_t)f * H + 2 * h2) * W) + wq);
  reinterpret_cast<uint2*>(s2 + (size_t)r2 * (W >> 1))[wq] = make_uint2(__byte_perm(b.x, b.y, 0x6420), __byte_perm(b.z, b.w, 0x6420));
  if ((h2 & 1u) == 0u)
    reinterpret_cast<uint32_t*>(s4 + ((size_t)f * (H >> 2) + (h2 >> 1)) * (W >> 2))[wq] =
        __byte_perm(__byte_perm(b.x, b.y, 0x4040), __byte_perm(b.z, b.w, 0x4040), 0x5410);
}

// Thread = 16 consecutive z of the voxel column (2 x2, 2 y2): only those columns feed the pyramid (a quarter of the grid is read)
__global__ void __launch_bounds__(256)
k_voxel_pyramid_vec(const uint8_t* __restrict__ vox, int F, int X, int Y, int Z, uint8_t* __restrict__ v2, uint8_t* __restrict__ v4) {
  const int Z16 = Z >> 4, X2 = X >> 1, Y2 = Y >> 1;
  const uint32_t t = blockIdx.x * 256u + threadIdx.x;
  if (t >= (uint32_t)F * X2 * Y2 * Z16) return;
  const uint32_t zq = t % Z16, col = t / Z16;                 // col = (f * X2 + x2) * Y2 + y2
  const uint32_t y2 = col % Y2, fx = col / Y2, x2 = fx % X2, f = fx / X2;
  const uint4 b = ld_stream_u4(reinterpret_cast<const uint4*>(vox + (((size_t)f * X + 2 * x2) * Y + 2 * y2) * Z) + zq);
  reinterpret_cast<uint2*>(v2 + (size_t)col * (Z >> 1))[zq] = make_uint2(__byte_perm(b.x, b.y, 0x6420), __byte_perm(b.z, b.w, 0x6420));
  if (((x2 | y2) & 1u) == 0u)
    reinterpret_cast<uint32_t*>(v4 + (((size_t)f * (X >> 2) + (x2 >> 1)) * (Y >> 2) + (y2 >> 1)) * (Z >> 2))[zq] =
        __byte_perm(__byte_perm(b.x, b.y, 0x4040), __byte_perm(b.z, b.w, 0x4040), 0x5410);
}

// ---------------------------------------------------------------- saved sparse voxels -> dense grid (dataset.py:317-327)
// rows [n,4] uint16 (x, y, z, label) of F files back to back (row_offsets [F+1]); label 255 -> 0, then remap; numpy's fancy
// assignment `voxels[x, y, z] = labels` lets the LAST row of a voxel win.  Files written by voxelize_one hold every voxel
// once, so the common case is one marking pass + one writing pass; voxels that occur more than once are detected by the
// marking pass (state byte 3) and only their rows search the rest of the frame for a later occurrence.
__device__ __forceinline__ bool row_voxel(const uint16_t* __restrict__ rows, int64_t i, int dx, int dy, int dz, int64_t* lin) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(rows) + i);
  const uint32_t x = r.x & 0xffffu, y = r.x >> 16, z = r.y & 0xffffu;
  if (x >= (uint32_t)dx || y >= (uint32_t)dy || z >= (uint32_t)dz) return false;        // (numpy: IndexError)
  *lin = ((int64_t)x * dy + y) * dz + z;
  return true;
}
__device__ __forceinline__ int frame_of_row(const int64_t* __restrict__ off, int F, int64_t i) {
  int lo = 0, hi = F;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(off + mid) <= i) lo = mid; else hi = mid; }
  return lo;
}

__global__ void __launch_bounds__(256)
k_densify_mark(const uint16_t* __restrict__ rows, const int64_t* __restrict__ off, int F, int64_t n, int dx, int dy, int dz,
               uint8_t* __restrict__ dense, int64_t* __restrict__ n_bad) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  int64_t lin;
  if (!row_voxel(rows, i, dx, dy, dz, &lin)) { if (n_bad) atomicAdd(reinterpret_cast<unsigned long long*>(n_bad), 1ull); return; }
  const int64_t a = (int64_t)frame_of_row(off, F, i) * dx * dy * dz + lin;
  uint32_t* word = reinterpret_cast<uint32_t*>(dense) + (a >> 2);
  const int sh = 8 * (int)(a & 3);
  const uint32_t old = atomicOr(word, 1u << sh);                      // state 1: seen once
  if ((old >> sh) & 1u) atomicOr(word, 2u << sh);                     // state 3: seen more than once
}

__global__ void __launch_bounds__(256)
k_densify_last(const uint16_t* __restrict__ rows, const int64_t* __restrict__ off, int F, int64_t n, int dx, int dy, int dz,
               const uint8_t* __restrict__ dense, uint8_t* __restrict__ writer) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  int64_t lin;
  uint8_t w = 0;
  if (row_voxel(rows, i, dx, dy, dz, &lin)) {
    const int f = frame_of_row(off, F, i);
    const uint8_t state = dense[(int64_t)f * dx * dy * dz + lin];
    w = 1;
    if (state == 3) {                                                 // (rare) is there a later row of this frame for the same voxel?
      const int64_t end = __ldg(off + f + 1);
      const uint2 me = __ldg(reinterpret_cast<const uint2*>(rows) + i);
      for (int64_t j = i + 1; j < end; ++j) {
        const uint2 o = __ldg(reinterpret_cast<const uint2*>(rows) + j);
        if (o.x == me.x && (o.y & 0xffffu) == (me.y & 0xffffu)) { w = 0; break; }
      }
    }
  }
  writer[i] = w;
}

__global__ void __launch_bounds__(256)
k_densify_write(const uint16_t* __restrict__ rows, const int64_t* __restrict__ off, int F, int64_t n, int dx, int dy, int dz,
                const uint8_t* __restrict__ remap, const uint8_t* __restrict__ writer, uint8_t* __restrict__ dense) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n || !writer[i]) return;
  int64_t lin;
  if (!row_voxel(rows, i, dx, dy, dz, &lin)) return;
  uint32_t lab = (uint32_t)(__ldg(reinterpret_cast<const uint2*>(rows) + i).y >> 16) & 0xffu;   // saved as uint16, values < 256
  if (lab == 255u) lab = 0u;                                          // dataset.py:322
  if (remap) lab = __ldg(remap + lab);                                // :323
  dense[(int64_t)frame_of_row(off, F, i) * dx * dy * dz + lin] = (uint8_t)lab;     // :325 (the state byte becomes the label)
}

}  // namespace
}  // namespace muvo

using namespace muvo;

extern "C" {

int muvo_densify_sparse(const uint16_t* rows, const int64_t* row_offsets, int32_t n_frames, int64_t n_rows, int32_t dx, int32_t dy,
                        int32_t dz, const uint8_t* remap256, uint8_t* dense_out, uint8_t* scratch_rows, int64_t* n_bad, void* stream) {
  if (n_frames < 0 || n_rows < 0 || dx <= 0 || dy <= 0 || dz <= 0) return MUVO_E_ARG;
  if (n_frames == 0) return MUVO_OK;
  if (!dense_out || !row_offsets || (n_rows > 0 && (!rows || !scratch_rows))) return MUVO_E_NULL;
  const int64_t G = (int64_t)dx * dy * dz;
  if (G > ((int64_t)1 << 31) || dx > 65535 || dy > 65535 || dz > 65535) return MUVO_E_SHAPE;
  if ((reinterpret_cast<uintptr_t>(dense_out) & 3) || (reinterpret_cast<uintptr_t>(rows) & 7)) return MUVO_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dense_out, 0, (size_t)n_frames * G, st);               // :324 np.zeros
  if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  if (n_rows == 0) return MUVO_OK;
  const unsigned grid = (unsigned)ceil_div64(n_rows, 256);
  k_densify_mark<<<grid, 256, 0, st>>>(rows, row_offsets, n_frames, n_rows, dx, dy, dz, dense_out, n_bad);
  MUVO_AFTER_LAUNCH("k_densify_mark", st);
  k_densify_last<<<grid, 256, 0, st>>>(rows, row_offsets, n_frames, n_rows, dx, dy, dz, dense_out, scratch_rows);
  MUVO_AFTER_LAUNCH("k_densify_last", st);
  k_densify_write<<<grid, 256, 0, st>>>(rows, row_offsets, n_frames, n_rows, dx, dy, dz, remap256, scratch_rows, dense_out);
  MUVO_AFTER_LAUNCH("k_densify_write", st);
  return MUVO_OK;
}

int muvo_label_pyramids(const float* range_xyzd, const uint8_t* range_sem, const uint8_t* voxel, int32_t F, int32_t H, int32_t W,
                        int32_t X, int32_t Y, int32_t Z, float scale, float* rv1, float* rv2, float* rv4, uint8_t* seg2,
                        uint8_t* seg4, uint8_t* vox2, uint8_t* vox4, void* stream) {
  if (F < 0 || H < 0 || W < 0 || X < 0 || Y < 0 || Z < 0) return MUVO_E_ARG;
  if (F == 0) return MUVO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (range_xyzd) {
    if (!(scale != 0.f)) return MUVO_E_ARG;
    if (!rv1 || !rv2 || !rv4 || (range_sem && (!seg2 || !seg4))) return MUVO_E_NULL;
    if (H < 4 || W < 4) return MUVO_E_SHAPE;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const int64_t items = (int64_t)4 * F * H * (W / 16) + (range_sem ? (int64_t)F * (H / 2) * (W / 16) : 0);
    if (H % 4 == 0 && W % 16 == 0 && items < ((int64_t)1 << 31) && al16(range_xyzd) && al16(rv1) && al16(rv2) && al16(rv4) &&
        (!range_sem || (al16(range_sem) && al16(seg2) && al16(seg4)))) {
      k_range_pyramid_vec<<<(unsigned)ceil_div64(items, 256), 256, 0, st>>>(range_xyzd, range_sem, 4 * F, F, H, W, scale, rv1, rv2, rv4, seg2, seg4);
      MUVO_AFTER_LAUNCH("k_range_pyramid_vec", st);
    } else {
      k_range_pyramid<<<kNumSMsB200 * 8, 256, 0, st>>>(range_xyzd, range_sem, F, H, W, scale, rv1, rv2, rv4, seg2, seg4);
      MUVO_AFTER_LAUNCH("k_range_pyramid", st);
    }
  }
  if (voxel) {
    if (!vox2 || !vox4) return MUVO_E_NULL;
    if (X < 4 || Y < 4 || Z < 4) return MUVO_E_SHAPE;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const int64_t items = (int64_t)F * (X / 2) * (Y / 2) * (Z / 16);
    if (X % 4 == 0 && Y % 4 == 0 && Z % 16 == 0 && items < ((int64_t)1 << 31) && al16(voxel) && al16(vox2) && al16(vox4)) {
      k_voxel_pyramid_vec<<<(unsigned)ceil_div64(items, 256), 256, 0, st>>>(voxel, F, X, Y, Z, vox2, vox4);
      MUVO_AFTER_LAUNCH("k_voxel_pyramid_vec", st);
    } else {
      k_voxel_pyramid<<<kNumSMsB200 * 8, 256, 0, st>>>(voxel, F, X, Y, Z, vox2, vox4);
      MUVO_AFTER_LAUNCH("k_voxel_pyramid", st);
    }
  }
  return MUVO_OK;
}

}  // extern "C"

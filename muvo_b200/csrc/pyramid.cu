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
  if (e != cudaSuccess) return (int)e;
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
    k_range_pyramid<<<kNumSMsB200 * 8, 256, 0, st>>>(range_xyzd, range_sem, F, H, W, scale, rv1, rv2, rv4, seg2, seg4);
    MUVO_AFTER_LAUNCH("k_range_pyramid", st);
  }
  if (voxel) {
    if (!vox2 || !vox4) return MUVO_E_NULL;
    if (X < 4 || Y < 4 || Z < 4) return MUVO_E_SHAPE;
    k_voxel_pyramid<<<kNumSMsB200 * 8, 256, 0, st>>>(voxel, F, X, Y, Z, vox2, vox4);
    MUVO_AFTER_LAUNCH("k_voxel_pyramid", st);
  }
  return MUVO_OK;
}

}  // extern "C"

// N1: the camera + LiDAR cloud that feeds voxel_filter, built on the device.
//
// Replaces merge_pcd (data/data_preprocessing.py:125-139) = read_img's depth decode (:72-77) + depth2pcd (:87-106)
// + convert_coor_img (:109-119) + convert_coor_lidar (:121-123) + the ego-box mask (:133-138), bit-exact in float64
// (compiled with -fmad=false, numpy's operation order).  Output order = the reference's: valid camera pixels in
// row-major order, then the LiDAR points, ego-box points removed -- an ORDER-PRESERVING compaction, because
// voxel_filter breaks exact key ties by the lowest index.
//   M1 k_merge_count : keep flag per element (pixel or LiDAR point), count per 256-element block
//   M2 k_merge_scan  : exclusive scan of the block counts (one CTA) -> n_out
//   M3 k_merge_write : recompute, write float64 xyz + uint8 label at block offset + ballot rank
#include <math.h>
#include "common.cuh"

namespace muvo {
namespace {

constexpr int kMergeBlock = 256;

struct MergeDev {
  int H, W;
  int64_t n_img, n_lidar;
  double f, cx, cy, range;
  double cam[3];      // forward, right, up (float32 values of the 4x4 matrix, :111-116)
  double lid[3];      // lidar position (:122)
  double lo[3], hi[3];
  int mask_ego;
};

// element e -> (kept, xyz, label)
__device__ __forceinline__ bool merge_element(const MergeDev& m, const uint8_t* __restrict__ img, const float* __restrict__ lxyz,
                                              const uint8_t* __restrict__ lsem, int64_t e, double* X, double* Y, double* Z,
                                              uint8_t* lab) {
  double px, py, pz;
  if (e < m.n_img) {
    const uchar4 p = reinterpret_cast<const uchar4*>(img)[e];                     // [B, G, R, semantic] as cv2 returns it
    const double code = (65536.0 * (double)p.z + 256.0 * (double)p.y) + (double)p.x;   // :76, exact
    const double depth = 1000.0 * (code / 16777215.0);
    if (!(depth < 1000.0)) return false;                                          // :94
    const int v = (int)(e / m.W), u = (int)(e - (int64_t)v * m.W);
    const double x = (((double)u - m.cx) * depth) / m.f, y = (((double)v - m.cy) * depth) / m.f;   // :101
    if (!(sqrt((x * x + y * y) + depth * depth) < m.range)) return false;         // :104
    px = depth + m.cam[0]; py = (-x) + (-m.cam[1]); pz = (-y) + m.cam[2];          // :111-118 (one +-1 coefficient per row)
    *lab = p.w;
  } else {
    const int64_t i = e - m.n_img;
    const float x = (float)((double)__ldg(lxyz + 3 * i) + m.lid[0]);              // :122 float32 += float64
    const float y = -(float)((double)__ldg(lxyz + 3 * i + 1) + m.lid[1]);         // :123
    const float z = (float)((double)__ldg(lxyz + 3 * i + 2) + m.lid[2]);
    px = (double)x; py = (double)y; pz = (double)z;
    *lab = __ldg(lsem + i);
  }
  if (m.mask_ego && m.lo[0] < px && px < m.hi[0] && m.lo[1] < py && py < m.hi[1] && m.lo[2] < pz && pz < m.hi[2]) return false;
  *X = px; *Y = py; *Z = pz;
  return true;
}

__global__ void __launch_bounds__(kMergeBlock)
k_merge_count(MergeDev m, const uint8_t* __restrict__ img, const float* __restrict__ lxyz, const uint8_t* __restrict__ lsem,
              uint32_t* __restrict__ block_cnt) {
  __shared__ uint32_t ws[kMergeBlock / 32];
  const int64_t e = (int64_t)blockIdx.x * kMergeBlock + threadIdx.x;
  double X, Y, Z; uint8_t lab;
  const bool keep = e < m.n_img + m.n_lidar && merge_element(m, img, lxyz, lsem, e, &X, &Y, &Z, &lab);
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int k = 0; k < kMergeBlock / 32; ++k) t += ws[k];
    block_cnt[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024)
k_merge_scan(uint32_t* __restrict__ block_cnt, int nblocks, int64_t* __restrict__ n_out, int64_t* __restrict__ row_offsets, int frame) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? block_cnt[i] : 0;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    const uint32_t carry = carry_s;
    uint32_t wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += wsum[w];
    __syncthreads();
    if (i < nblocks) block_cnt[i] = carry + wbase + incl - v;
    if (threadIdx.x == 1023) carry_s = carry + wbase + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (n_out) *n_out = (int64_t)carry_s;
    if (row_offsets) row_offsets[frame + 1] = row_offsets[frame] + (int64_t)carry_s;    // batched: frames packed back to back
  }
}

__global__ void __launch_bounds__(kMergeBlock)
k_merge_write(MergeDev m, const uint8_t* __restrict__ img, const float* __restrict__ lxyz, const uint8_t* __restrict__ lsem,
              const uint32_t* __restrict__ block_off, double* __restrict__ xyz_out, uint8_t* __restrict__ sem_out,
              const int64_t* __restrict__ row_offsets, int frame, int64_t capacity) {
  __shared__ uint32_t ws[kMergeBlock / 32];
  const int64_t e = (int64_t)blockIdx.x * kMergeBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double X = 0, Y = 0, Z = 0; uint8_t lab = 0;
  const bool keep = e < m.n_img + m.n_lidar && merge_element(m, img, lxyz, lsem, e, &X, &Y, &Z, &lab);
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) ws[warp] = __popc(bal);
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int k = 0; k < kMergeBlock / 32; ++k) if (k < warp) wbase += ws[k];
  if (keep) {
    const int64_t r = (row_offsets ? row_offsets[frame] : 0) + (int64_t)block_off[blockIdx.x] + wbase + __popc(bal & ((1u << lane) - 1u));
    if (capacity >= 0 && r >= capacity) return;                        // (batched output buffer too small: the host checks the offsets)
    xyz_out[3 * r] = X; xyz_out[3 * r + 1] = Y; xyz_out[3 * r + 2] = Z;
    sem_out[r] = lab;
  }
}

}  // namespace
}  // namespace muvo

using namespace muvo;

extern "C" {

int muvo_merge_pcd_workspace_bytes(int32_t H, int32_t W, int64_t n_lidar, size_t* bytes_out_h) {
  if (!bytes_out_h) return MUVO_E_NULL;
  if (H < 0 || W < 0 || n_lidar < 0) return MUVO_E_ARG;
  const int64_t n = (int64_t)H * W + n_lidar;
  *bytes_out_h = align_up((size_t)ceil_div64(n > 0 ? n : 1, kMergeBlock) * 4, 256) + 256;
  return MUVO_OK;
}

static int run_merge(const uint8_t* img_bgra, int32_t H, int32_t W, double focal, double range, const double* camera_pos_h,
                     const float* lidar_xyz, const uint8_t* lidar_sem, int64_t n_lidar, const double* lidar_pos_h,
                     const double* ego_box_h, double* xyz_out, uint8_t* sem_out, int64_t* n_out, int64_t* row_offsets, int frame,
                     int64_t capacity, void* ws, size_t ws_bytes, void* stream) {
  if (H < 0 || W < 0 || n_lidar < 0 || ((int64_t)H * W > 0 && !(focal > 0.0))) return MUVO_E_ARG;
  if (!camera_pos_h || !lidar_pos_h || (!n_out && !row_offsets) || !ws) return MUVO_E_NULL;
  const int64_t n_img = (int64_t)H * W, n = n_img + n_lidar;
  if ((n_img > 0 && !img_bgra) || (n_lidar > 0 && (!lidar_xyz || !lidar_sem)) || (n > 0 && (!xyz_out || !sem_out))) return MUVO_E_NULL;
  if (n >= ((int64_t)1 << 31) * kMergeBlock) return MUVO_E_SHAPE;
  if (reinterpret_cast<uintptr_t>(img_bgra) & 3) return MUVO_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nblocks = ceil_div64(n > 0 ? n : 1, kMergeBlock);
  if ((size_t)nblocks * 4 > ws_bytes) return MUVO_E_WORKSPACE;
  MergeDev m{};
  m.H = H; m.W = W; m.n_img = n_img; m.n_lidar = n_lidar;
  m.f = focal; m.cx = W / 2.0; m.cy = H / 2.0; m.range = range;
  for (int k = 0; k < 3; ++k) { m.cam[k] = camera_pos_h[k]; m.lid[k] = lidar_pos_h[k]; }
  m.mask_ego = ego_box_h != nullptr;
  if (ego_box_h) for (int k = 0; k < 3; ++k) { m.lo[k] = ego_box_h[k]; m.hi[k] = ego_box_h[3 + k]; }
  uint32_t* block_cnt = (uint32_t*)ws;
  prof_mark("<merge_pcd>", st);
  k_merge_count<<<(unsigned)nblocks, kMergeBlock, 0, st>>>(m, img_bgra, lidar_xyz, lidar_sem, block_cnt);
  MUVO_AFTER_LAUNCH("k_merge_count", st);
  k_merge_scan<<<1, 1024, 0, st>>>(block_cnt, (int)nblocks, n_out, row_offsets, frame);
  MUVO_AFTER_LAUNCH("k_merge_scan", st);
  k_merge_write<<<(unsigned)nblocks, kMergeBlock, 0, st>>>(m, img_bgra, lidar_xyz, lidar_sem, block_cnt, xyz_out, sem_out, row_offsets,
                                                            frame, capacity);
  MUVO_AFTER_LAUNCH("k_merge_write", st);
  return MUVO_OK;
}

int muvo_merge_pcd(const uint8_t* img_bgra, int32_t H, int32_t W, double focal, double range, const double* camera_pos_h,
                   const float* lidar_xyz, const uint8_t* lidar_sem, int64_t n_lidar, const double* lidar_pos_h,
                   const double* ego_box_h, double* xyz_out, uint8_t* sem_out, int64_t* n_out, void* ws, size_t ws_bytes,
                   void* stream) {
  if (!n_out) return MUVO_E_NULL;
  return run_merge(img_bgra, H, W, focal, range, camera_pos_h, lidar_xyz, lidar_sem, n_lidar, lidar_pos_h, ego_box_h, xyz_out, sem_out,
                   n_out, nullptr, 0, -1, ws, ws_bytes, stream);
}

int muvo_merge_pcd_at(const uint8_t* img_bgra, int32_t H, int32_t W, double focal, double range, const double* camera_pos_h,
                      const float* lidar_xyz, const uint8_t* lidar_sem, int64_t n_lidar, const double* lidar_pos_h,
                      const double* ego_box_h, double* xyz_out, uint8_t* sem_out, int64_t capacity_rows, int64_t* row_offsets,
                      int32_t frame, void* ws, size_t ws_bytes, void* stream) {
  if (!row_offsets) return MUVO_E_NULL;
  if (frame < 0 || capacity_rows < 0) return MUVO_E_ARG;
  return run_merge(img_bgra, H, W, focal, range, camera_pos_h, lidar_xyz, lidar_sem, n_lidar, lidar_pos_h, ego_box_h, xyz_out, sem_out,
                   nullptr, row_offsets, frame, capacity_rows, ws, ws_bytes, stream);
}

}  // extern "C"
